"""C++ host mirror of the reference's prover surface (include/hfb200_prover.hpp: default_prover / prove / Receipt JSON / verify)
above the C ABI.  The reference is compiled Rust and there is no Rust toolchain here, so the host side is C++; the demo program is
compiled with g++ and linked against the host-emulator build of the kernel sources (CPU tier) or libhfb200.so (GPU tier).  The
receipt it writes must parse with the Python mirror, carry the oracle's seals bit for bit, and verify."""
import os
import subprocess
import sys
import numpy as np
import pytest
from conftest import SMALL, make_segment

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build_demo(lib_path, out):
    cmd = ["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_demo.cpp"), "-o", out,
           lib_path, "-Wl,-rpath," + os.path.dirname(lib_path)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return out


def _run_and_check(pkg, orc, exe, widths, po2s, tmp_path, device=0, lib=None):
    out_json = str(tmp_path / "receipt.json")
    r = subprocess.run([exe, str(device)] + [str(w) for w in widths] + [out_json] + [str(p) for p in po2s], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.startswith("OK segments=%d" % len(po2s))
    rec = pkg.Receipt.from_json(open(out_json).read())           # the Python mirror reads what the C++ mirror wrote
    assert rec.journal.decode() == '{"iban":"CH4308307000289537312"}'
    ids = {}
    made = [make_segment(orc, widths, po2, trace_seed=500 + i, blind_seed=9 + i) for i, po2 in enumerate(po2s)]
    # the claims the C++ mirror wrote into the globals are the ones the Python mirror computes for the same session
    session = pkg.Session([pkg.Segment(i, po2, m[1], m[2], m[3], 9 + i) for i, (po2, m) in enumerate(zip(po2s, made))], journal='{"iban":"CH4308307000289537312"}')
    claimed, _ = pkg.bind_claims(session, lib=lib)
    for i, (po2, s) in enumerate(zip(po2s, rec.inner.segments)):
        cir, g, code, data = made[i]
        oseal, ocps, _ = cir.prove(po2, claimed[i], code, data, 9 + i)    # same trace, same claims, same seeds: the oracle's seal
        assert s.index == i and np.array_equal(np.asarray(s.seal, dtype=np.uint32), oseal)
        ids[po2] = ocps["code_root"]
    return rec, ids


def test_cpp_host_mirror_on_emulator(pkg, emu_lib, orc, tmp_path):
    emu_path = os.path.join(ROOT, "tests", "emu", "libhfb200_emu.so")
    exe = _build_demo(emu_path, str(tmp_path / "host_demo"))
    rec, ids = _run_and_check(pkg, orc, exe, SMALL, [12, 13, 12], tmp_path, lib=emu_lib)
    rec.verify(pkg.default_image_id(emu_lib), ids, circuit=SMALL, lib=emu_lib)


@pytest.mark.gpu
def test_cpp_host_mirror_on_gpu(pkg, gpu_lib, orc, tmp_path):
    exe = _build_demo(pkg.LIB_PATH, str(tmp_path / "host_demo"))
    rec, ids = _run_and_check(pkg, orc, exe, (16, 64, 16), [14, 13, 14, 12], tmp_path, lib=gpu_lib)
    rec.verify(pkg.default_image_id(gpu_lib), ids, circuit=(16, 64, 16))
