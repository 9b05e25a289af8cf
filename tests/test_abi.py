"""The C-ABI boundary: the built library exports every symbol include/hfb200.h declares, the product has no
CPU fallback, and nothing in the product reaches into oracle/."""
import os
import re
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    h = open(os.path.join(ROOT, "include", "hfb200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(hfb200_[a-z0-9_]+)\s*\(", h)))


def test_header_and_binding_agree(pkg):
    assert _declared() == sorted(pkg.EXPORTS)


def test_library_exports_every_declared_symbol(pkg):
    import __graft_entry__
    __graft_entry__.build()
    out = subprocess.check_output(["nm", "-D", "--defined-only", pkg.LIB_PATH]).decode()
    exported = set(re.findall(r"\bT (hfb200_[a-z0-9_]+)", out))
    assert set(_declared()) <= exported
    lib = pkg.load_library()
    assert b"sm_100a" in lib.hfb200_version()


def test_library_is_sm100a_and_has_no_emulator(pkg):
    out = subprocess.check_output(["cuobjdump", "-lelf", pkg.LIB_PATH]).decode()
    assert "sm_100a" in out
    assert b"EMULATOR" not in open(pkg.LIB_PATH, "rb").read()


def test_product_never_touches_the_oracle():
    pkg_dir = os.path.join(ROOT, "hyperfridge-r0_b200")
    for dp, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".inc", ".rs", ".toml")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert '#include "../../oracle' not in txt and "liboracle" not in txt, f


def test_no_cpu_fallback_without_gpu(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pkg.Hfb200Error, match="no CUDA device|CUDA"):
        pkg.Context(0, 12, (8, 16, 8))


def test_public_headers_compile_standalone(tmp_path):
    """include/hfb200.h is plain C (C99, no C++ or CUDA types); include/hfb200_prover.hpp is C++17 on top of it, warning-free."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    c = tmp_path / "h.c"
    c.write_text('#include "hfb200.h"\nint main(void) { return 0; }\n')
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), "-c", str(c), "-o", str(tmp_path / "h.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cpp = tmp_path / "h.cpp"
    cpp.write_text('#include "hfb200_prover.hpp"\nint main() { return 0; }\n')
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), "-c", str(cpp), "-o", str(tmp_path / "h2.o")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
