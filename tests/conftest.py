import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# Parity tests compare seals with the CPU oracle, so they need reproducible blinding: the binding switches every context it
# creates to HFB200_BLIND_DETERMINISTIC.  The C ABI's default stays OS entropy (tests/test_blinding.py checks both).
os.environ.setdefault("HFB200_DETERMINISTIC_BLINDING", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    import oracle
    oracle.build()
    return oracle


@pytest.fixture(scope="session")
def pkg():
    import hfb200_loader
    return hfb200_loader.load()


@pytest.fixture(scope="session")
def emu_lib(pkg):
    """Host emulator of the CUDA kernel sources (tests only)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("hfb200_build", os.path.join(ROOT, "hyperfridge-r0_b200", "build.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    return pkg.load_library(b.build_emu())


@pytest.fixture(scope="session")
def gpu_lib(pkg):
    return pkg.load_library()


SMALL = (8, 16, 8)       # narrow circuit for fast CPU-side runs
DEFAULT = (16, 192, 48)  # the declared W = 256 shape
TRACE_SEED = 0x48595046


def make_segment(orc, widths, po2, trace_seed=TRACE_SEED, blind_seed=1):
    cir = orc.Circuit(*widths)
    code = cir.gen_code(po2)
    g = cir.gen_globals(trace_seed)
    data = cir.gen_data(po2, code, g, trace_seed, blind_seed)
    return cir, g, code, data


def rand_elems(rng, shape):
    return rng.integers(0, 2013265921, size=shape, dtype=np.uint32)
