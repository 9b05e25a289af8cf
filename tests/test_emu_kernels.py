"""The CUDA kernel SOURCES compiled for the host (tests/emu, -DHFB200_EMU) against the oracle.

This tier exists because the build container has no GPU: it checks the kernels' index arithmetic, twiddle
schedules and the whole device-side pipeline (trace-domain DEEP quotient etc.) bit-for-bit before GPU minutes
are spent.  The emulator is test infrastructure: libhfb200.so never contains it and the package never loads it.
"""
import json
import os
import numpy as np
import pytest
from conftest import SMALL, TRACE_SEED, make_segment, rand_elems

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx(pkg, emu_lib):
    c = pkg.Context(0, 14, SMALL, lib=emu_lib)
    yield c
    c.close()


def test_is_emulator(ctx):
    assert "EMULATOR" in ctx.version


def test_poseidon2(ctx, orc):
    st = rand_elems(np.random.default_rng(0), (7, 24))
    assert (ctx.op_poseidon2(st) == np.stack([orc.poseidon2_mix(s) for s in st])).all()


@pytest.mark.parametrize("lg", [1, 2, 5, 10, 11, 12, 14, 16, 18])
def test_ntt_ops(ctx, orc, lg):
    x = rand_elems(np.random.default_rng(lg), (3, 1 << lg))
    coeffs = orc.interpolate_ntt(x)
    assert (ctx.op_interpolate_ntt(x, False) == coeffs).all()
    assert (ctx.op_interpolate_ntt(x, True) == orc.zk_shift(coeffs)).all()
    assert (ctx.op_expand_ntt(x, 2) == orc.expand_ntt(x, 2)).all()
    assert (ctx.op_expand_ntt(x, 0) == orc.expand_ntt(x, 0)).all()
    assert (ctx.op_lde(x) == orc.expand_ntt(orc.zk_shift(coeffs), 2)).all()


@pytest.mark.parametrize("a_env,lg", [("6", 16), ("11", 13), ("12", 14), ("2", 12)])
def test_ntt_alternate_plans(pkg, emu_lib, orc, a_env, lg, monkeypatch):
    monkeypatch.setenv("HFB200_NTT_A", a_env)
    with pkg.Context(0, 12, SMALL, lib=emu_lib) as c:
        x = rand_elems(np.random.default_rng(5), (2, 1 << lg))
        assert (c.op_lde(x) == orc.expand_ntt(orc.zk_shift(orc.interpolate_ntt(x)), 2)).all()
        assert (c.op_expand_ntt(x, 0) == orc.expand_ntt(x, 0)).all()


@pytest.mark.parametrize("rows,cols", [(2, 1), (16, 3), (64, 16), (1024, 17), (2048, 64)])
def test_merkle(ctx, orc, rows, cols):
    m = rand_elems(np.random.default_rng(rows + cols), (cols, rows))
    _, nodes = orc.merkle(m, True)
    assert (ctx.op_merkle(m)[1:] == nodes[1:]).all()


def test_full_segment_matches_oracle_and_golden(pkg, emu_lib, orc, monkeypatch):
    monkeypatch.setenv("HFB200_DEBUG_CHECKPOINTS", "1")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_small.json")))["segment"]
    for po2 in (12, 13):
        cir, g, code, data = make_segment(orc, SMALL, po2)
        oseal, ocps, _ = cir.prove(po2, g, code, data, 1)
        with pkg.Context(0, po2, SMALL, lib=emu_lib) as c:
            # witness stand-in parity
            assert (c.witgen_synth(po2, TRACE_SEED, 1) == g).all()
            assert (c.read_group(1) == code).all() and (c.read_group(2) == data).all()
            seal = c.prove_segment(po2, g, code, data, 1)
            cps = c.checkpoints()
            for k, v in ocps.items():
                assert (cps[k] == v).all(), k
            assert len(seal) == c.seal_words(po2) == len(oseal) and (seal == oseal).all()
            assert (c.read_group(0) == cir.step_accum(po2, data, ocps["accum_mix"], 1)).all()
            assert (c.prove_resident(1) == oseal).all()
            assert cir.verify(seal, ocps["code_root"]) == po2
            if po2 == gold["po2"]:
                for k, v in gold["checkpoints"].items():
                    assert cps[k].tolist() == v, k
            # two-phase API with an externally computed accum (oracle's step_accum)
            mix = c.segment_begin(po2, g, code, data, 1)
            assert (mix == ocps["accum_mix"]).all()
            assert (c.segment_finish(cir.step_accum(po2, data, mix, 1)) == oseal).all()


def test_error_paths(pkg, emu_lib):
    with pytest.raises(pkg.Hfb200Error):
        pkg.Context(0, 30, SMALL, lib=emu_lib)
    with pytest.raises(pkg.Hfb200Error):
        pkg.Context(0, 12, (3, 16, 8), lib=emu_lib)
    with pkg.Context(0, 12, SMALL, lib=emu_lib) as c:
        with pytest.raises(pkg.Hfb200Error):
            c._po2 = 12
            c.prove_resident(1)  # no resident trace yet
        with pytest.raises(pkg.Hfb200Error):
            c.witgen_synth(13, 1, 1)  # above max_po2


def test_abi_buffer_contract(pkg, emu_lib, orc):
    """Seal buffer too small -> error and the required size; mix buffer too small -> error; NULL trace -> error."""
    import ctypes as C
    cir, g, code, data = make_segment(orc, SMALL, 12)
    with pkg.Context(0, 12, SMALL, lib=emu_lib) as c:
        need = c.seal_words(12)
        seal = np.zeros(16, np.uint32)
        got = C.c_size_t()
        e = emu_lib.hfb200_prove_segment(c._h, 12, g.ctypes.data, code.ctypes.data, data.ctypes.data, 1, seal.ctypes.data, seal.size, C.byref(got))
        assert e and b"too small" in C.cast(e, C.c_char_p).value and got.value == need
        emu_lib.hfb200_free_error(e)
        e = emu_lib.hfb200_prove_segment(c._h, 12, g.ctypes.data, None, data.ctypes.data, 1, seal.ctypes.data, seal.size, C.byref(got))
        assert e and b"NULL" in C.cast(e, C.c_char_p).value
        emu_lib.hfb200_free_error(e)
        bad = g.copy(); bad[3] = 0xFFFFFFFF  # INVALID marker is never a valid input
        with pytest.raises(pkg.Hfb200Error, match="non-canonical"):
            c.prove_segment(12, bad, code, data, 1)
        with pytest.raises(pkg.Hfb200Error):
            c.segment_finish(None)  # finish without begin


@pytest.mark.parametrize("widths,po2", [((21, 40, 12), 12), ((5, 8, 4), 12), ((16, 200, 52), 12)])
def test_odd_circuit_widths(pkg, emu_lib, orc, widths, po2):
    """Ragged shapes: widths that are not multiples of the sponge rate (partial absorb blocks), the minimum circuit,
    and column-group sizes that do not divide the per-block column counts of the NTT / dot kernels."""
    cir, g, code, data = make_segment(orc, widths, po2)
    oseal, ocps, _ = cir.prove(po2, g, code, data, 3)
    with pkg.Context(0, po2, widths, lib=emu_lib) as c:
        seal = c.prove_segment(po2, g, code, data, 3)
        assert len(seal) == len(oseal) and (seal == oseal).all()
        assert cir.verify(seal, ocps["code_root"]) == po2


@pytest.mark.parametrize("widths,po2", [((8, 16, 8), 12), ((16, 64, 16), 13), ((21, 40, 12), 12)])
def test_device_transcript_gives_the_same_seal(pkg, emu_lib, orc, widths, po2):
    """hfb200_set_transcript(1): the Fiat-Shamir transcript as one-warp kernels (csrc/transcript.cuh), seal assembled on the device,
    ONE synchronisation per segment -- the same seal and the same checkpoints as the host transcript and the oracle."""
    cir = orc.Circuit(*widths)
    code = cir.gen_code(po2); g = cir.gen_globals(11); data = cir.gen_data(po2, code, g, 11, 4)
    oseal, ocps, _ = cir.prove(po2, g, code, data, 4)
    with pkg.Context(0, po2, widths, lib=emu_lib) as c:
        host = c.prove_segment(po2, g, code, data, 4)
        assert c.last_stats()["host_syncs"] >= 8   # 4 + FRI rounds roots, tap evaluations, final coefficients, openings
        c.set_transcript(True)
        dev = c.prove_segment(po2, g, code, data, 4)
        assert c.last_stats()["host_syncs"] == 1
        cps = c.checkpoints()
        for k, v in ocps.items():
            if k in cps:
                assert (cps[k] == v).all(), k
        assert len(dev) == len(oseal) and (dev == oseal).all() and (host == oseal).all()
        root = c.control_root(po2, code)                      # cached control group + device transcript
        assert (c.prove_segment(po2, g, None, data, 4) == oseal).all() and (root == ocps["code_root"]).all()
        c.witgen_synth(po2, 11, 4)
        assert (c.prove_resident(4) == oseal).all()           # resident entry
        c.set_transcript(2)                                   # graph-replay mode: the device-resident key / globals[0] path (the emulator has no graphs)
        for _ in range(3):
            assert (c.prove_segment(po2, g, code, data, 4) == oseal).all()
        assert c.graph_launches() == 0
        with pytest.raises(pkg.Hfb200Error, match="mode must be"):
            c.set_transcript(3)
