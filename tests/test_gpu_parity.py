"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle (bit-exact: all
arithmetic on this path is integer mod p) and, at BASELINE.json's full size (po2 = 20), through size-independent
properties (verifier acceptance, linearity of the LDE, determinism, seal-length model)."""
import json
import os
import numpy as np
import pytest
from conftest import SMALL, DEFAULT, TRACE_SEED, make_segment, rand_elems

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2013265921


@pytest.fixture(scope="module")
def ctx(pkg, gpu_lib):
    c = pkg.Context(0, 16, DEFAULT, lib=gpu_lib)
    assert "sm_100a" in c.version
    yield c
    c.close()


def test_poseidon2_permutation(ctx, orc):
    st = rand_elems(np.random.default_rng(0), (1000, 24))
    st[0] = 0
    st[1] = orc.encode(list(range(24)))
    got = ctx.op_poseidon2(st)
    exp = np.stack([orc.poseidon2_mix(s) for s in st])
    assert (got == exp).all()
    # upstream's own known-answer vector (risc0-zkp `poseidon2_test_vectors`, see tests/test_poseidon2.py) on the GPU kernel
    from test_poseidon2 import UPSTREAM_KAT_0_TO_23
    assert [int(v) for v in orc.decode(got[1])] == UPSTREAM_KAT_0_TO_23


@pytest.mark.parametrize("lg", [1, 4, 10, 11, 12, 14, 15, 16, 17, 18, 19, 20, 21])
def test_ntt_ops(ctx, orc, lg):
    ncols = 3 if lg < 19 else 2
    x = rand_elems(np.random.default_rng(lg), (ncols, 1 << lg))
    x[0, :] = np.uint32(P - 1)  # extreme values
    coeffs = orc.interpolate_ntt(x)
    shifted = orc.zk_shift(coeffs)
    assert (ctx.op_interpolate_ntt(x, False) == coeffs).all()
    assert (ctx.op_interpolate_ntt(x, True) == shifted).all()
    assert (ctx.op_expand_ntt(x, 2) == orc.expand_ntt(x, 2)).all()
    assert (ctx.op_expand_ntt(x, 0) == orc.expand_ntt(x, 0)).all()
    assert (ctx.op_lde(x) == orc.expand_ntt(shifted, 2)).all()


def test_ntt_size_22_inverse(ctx, orc):
    # the check polynomial of a po2 = 20 segment is interpolated over 2^22 points
    x = rand_elems(np.random.default_rng(22), (1, 1 << 22))
    assert (ctx.op_interpolate_ntt(x, False) == orc.interpolate_ntt(x)).all()


@pytest.mark.parametrize("rows,cols", [(2, 1), (16, 3), (64, 16), (1024, 17), (4096, 64), (1 << 16, 48), (1 << 14, 272)])
def test_merkle(ctx, orc, rows, cols):
    m = rand_elems(np.random.default_rng(rows + cols), (cols, rows))
    _, nodes = orc.merkle(m, True)
    assert (ctx.op_merkle(m)[1:] == nodes[1:]).all()


def test_fri_fold_matches_definition(ctx, orc):
    rng = np.random.default_rng(7)
    n = 1 << 12
    x = rand_elems(rng, (4, n))
    mix = rand_elems(rng, 4)
    got = ctx.op_fri_fold(x, mix)
    # definition on natural-order coefficients: out_m = sum_i mix^i c[16 m + i]; buffers are bit-reversed
    L = orc.lib()

    def e4mul(a, b):
        out = np.zeros(4, np.uint32)
        a = np.ascontiguousarray(a, np.uint32); b = np.ascontiguousarray(b, np.uint32)
        L.orc_fp4_mul(a.ctypes.data, b.ctypes.data, out.ctypes.data)
        return out
    nat = orc.bit_reverse(x)
    outnat = orc.bit_reverse(got)
    one = np.array([orc.encode([1])[0], 0, 0, 0], np.uint32)
    for m in (0, 1, 17, n // 16 - 1):
        tot = np.zeros(4, np.uint64)
        cur = one
        for i in range(16):
            tot = (tot + e4mul(cur, nat[:, 16 * m + i]).astype(np.uint64)) % P
            cur = e4mul(cur, mix)
        assert (outnat[:, m] == tot.astype(np.uint32)).all()


# (SMALL, 20): the headline size itself, bit for bit against the oracle (the narrow circuit keeps the oracle at ~20 s): the
# 2^10 x 2^10 NTT plan, the 1024-thread strided tiles and the compile-time fused middle kernel are the ones bench.py times
@pytest.mark.parametrize("widths,po2", [(SMALL, 12), (SMALL, 13), (SMALL, 15), (SMALL, 17), (DEFAULT, 12), (DEFAULT, 14), (DEFAULT, 16), (DEFAULT, 18), (SMALL, 20)])
def test_segment_seal_bit_exact(pkg, gpu_lib, orc, widths, po2, monkeypatch):
    monkeypatch.setenv("HFB200_DEBUG_CHECKPOINTS", "1")
    cir, g, code, data = make_segment(orc, widths, po2)
    oseal, ocps, _ = cir.prove(po2, g, code, data, 1)
    with pkg.Context(0, po2, widths, lib=gpu_lib) as c:
        assert (c.witgen_synth(po2, TRACE_SEED, 1) == g).all()
        assert (c.read_group(1) == code).all() and (c.read_group(2) == data).all()
        seal_res = c.prove_resident(1)                      # inputs resident in HBM
        cps = c.checkpoints()
        for k, v in ocps.items():
            assert (cps[k] == v).all(), "checkpoint %s differs" % k
        assert len(seal_res) == c.seal_words(po2) == len(oseal)
        assert (seal_res == oseal).all()
        assert (c.read_group(0) == cir.step_accum(po2, data, ocps["accum_mix"], 1)).all()
        seal_host = c.prove_segment(po2, g, code, data, 1)  # host buffers through the C ABI
        assert (seal_host == oseal).all()
        mix = c.segment_begin(po2, g, code, data, 1)        # two-phase, external accum
        assert (mix == ocps["accum_mix"]).all()
        assert (c.segment_finish(cir.step_accum(po2, data, mix, 1)) == oseal).all()
        assert cir.verify(seal_res, ocps["code_root"]) == po2
        st = c.last_stats()
        assert st["launches"] > 20 and st["ms_total"] > 0


def test_golden_fixture(pkg, gpu_lib, monkeypatch):
    monkeypatch.setenv("HFB200_DEBUG_CHECKPOINTS", "1")
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "golden_small.json")))["segment"]
    with pkg.Context(0, gold["po2"], tuple(gold["widths"]), lib=gpu_lib) as c:
        c.witgen_synth(gold["po2"], gold["trace_seed"], gold["blind_seed"])
        seal = c.prove_resident(gold["blind_seed"])
        assert len(seal) == gold["seal_words"]
        cps = c.checkpoints()
        for k, v in gold["checkpoints"].items():
            assert cps[k].tolist() == v, k


def _check_against_pins(orc, gold, seal, cps):
    for k, v in gold["checkpoints"].items():
        assert cps[k].tolist() == v, "checkpoint %s differs from the pinned oracle value" % k
    assert len(seal) == gold["seal_words"]
    assert seal[:64].tolist() == gold["seal_head"] and seal[-64:].tolist() == gold["seal_tail"]
    assert int(seal.astype(np.uint64).sum()) == gold["seal_sum_u64"]
    assert orc.hash_elems(seal % orc.P).tolist() == gold["seal_hash"]


def _load_pins(po2):
    return json.load(open(os.path.join(ROOT, "tests", "golden", "golden_po2_%d_w256.json" % po2)))


def test_headline_po2_20_w256_bit_exact_pinned(pkg, gpu_lib, orc, monkeypatch):
    """BASELINE.json configs[1] at its largest size, the exact shape bench.py times (W = 256, po2 = 20: 192-column
    chunked H2D, 12-permutation leaves, the 2^10 x 2^10 NTT plan): every transcript checkpoint, the seal's hash, length,
    head and tail against the pins generated from the CPU oracle (tests/golden/make_golden_large.py, 131 s on 8
    threads) -- resident path, host-buffer path through hfb200_prove_segment, and FOUR contexts in flight on one GPU
    (bench.py's configuration), all against the same pins."""
    import threading
    monkeypatch.setenv("HFB200_DEBUG_CHECKPOINTS", "1")
    gold = _load_pins(20)
    po2, W = gold["po2"], tuple(gold["widths"])
    with pkg.Context(0, po2, W, lib=gpu_lib) as c:
        g = c.witgen_synth(po2, gold["trace_seed"], gold["blind_seed"])
        seal = c.prove_resident(gold["blind_seed"])
        _check_against_pins(orc, gold, seal, c.checkpoints())
        code, data = c.read_group(1), c.read_group(2)
        seal_host = c.prove_segment(po2, g, code, data, gold["blind_seed"])     # host buffers, chunked H2D
        _check_against_pins(orc, gold, seal_host, c.checkpoints())
    ctxs = [pkg.Context(0, po2, W, lib=gpu_lib) for _ in range(4)]
    out = [None] * 4

    def work(k):
        for _ in range(2):
            out[k] = ctxs[k].prove_segment(po2, g, code, data, gold["blind_seed"])
    th = [threading.Thread(target=work, args=(k,)) for k in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    for k in range(4):
        _check_against_pins(orc, gold, out[k], ctxs[k].checkpoints())
    [c.close() for c in ctxs]


def test_headline_po2_20_w256_bit_exact_live_oracle(pkg, gpu_lib, orc):
    """The same comparison against a LIVE oracle run (about two minutes of host time on 8-16 threads), so the pin file
    itself cannot go stale unnoticed: whole seal, word for word."""
    cir, g, code, data = make_segment(orc, DEFAULT, 20)
    oseal, ocps, _ = cir.prove(20, g, code, data, 1)
    with pkg.Context(0, 20, DEFAULT, lib=gpu_lib) as c:
        seal = c.prove_segment(20, g, code, data, 1)
        assert len(seal) == len(oseal) and (seal == oseal).all()
        cps = c.checkpoints()
        for k, v in ocps.items():
            if k in cps:
                assert (cps[k] == v).all(), k


def test_tampered_gpu_seal_rejected(pkg, gpu_lib, orc):
    with pkg.Context(0, 12, SMALL, lib=gpu_lib) as c:
        c.witgen_synth(12, TRACE_SEED, 1)
        seal = c.prove_resident(1)
        cir = orc.Circuit(*SMALL)
        cid = cir.control_id(12)
        assert cir.verify(seal, cid) == 12
        for pos in (3, 40, len(seal) // 2, len(seal) - 5):
            bad = seal.copy(); bad[pos] ^= 1
            with pytest.raises(RuntimeError):
                cir.verify(bad, cid)


def test_full_size_po2_20_properties(pkg, gpu_lib, orc):
    """BASELINE.json's full size: too slow for a seal-vs-oracle comparison (the oracle needs minutes), so check
    size-independent properties instead."""
    po2 = 20
    with pkg.Context(0, po2, DEFAULT, lib=gpu_lib) as c:
        c.witgen_synth(po2, TRACE_SEED, 1)
        seal = c.prove_resident(1)
        cir = orc.Circuit(*DEFAULT)
        assert len(seal) == c.seal_words(po2) == cir.seal_words_model(po2)
        assert (c.prove_resident(1) == seal).all()            # deterministic
        assert cir.verify(seal, c.checkpoint("code_root")) == po2  # the independent verifier accepts
        assert (cir.control_id(po2) == c.checkpoint("code_root")).all()
        # linearity of the fused trace -> LDE pipeline, and LDE of a constant column
        rng = np.random.default_rng(20)
        x = rand_elems(rng, (2, 1 << po2))
        s = ((x[0].astype(np.uint64) + x[1]) % P).astype(np.uint32)
        const = np.full(1 << po2, 12345, np.uint32)
        out = c.op_lde(np.stack([x[0], x[1], s, const]))
        assert (((out[0].astype(np.uint64) + out[1]) % P).astype(np.uint32) == out[2]).all()
        assert (out[3] == 12345).all()
        # inverse o forward round trip at 2^20
        assert (c.op_expand_ntt(c.op_interpolate_ntt(x, False), 0) == x).all()


def test_po2_22_large_segment(pkg, gpu_lib, orc):
    """BASELINE.json configs[4]: one po2 = 22 segment per GPU (29 GB arena).  Size-independent checks only."""
    po2 = 22
    with pkg.Context(0, po2, DEFAULT, lib=gpu_lib) as c:
        c.witgen_synth(po2, TRACE_SEED, 1)
        seal = c.prove_resident(1)
        cir = orc.Circuit(*DEFAULT)
        assert len(seal) == c.seal_words(po2) == cir.seal_words_model(po2)
        assert cir.verify(seal, c.checkpoint("code_root")) == po2
        st = c.last_stats()
        print("po2=22 segment: %.1f ms (ntt %.1f, hash %.1f)" % (st["ms_total"], st["ms_ntt_main"], st["ms_hash_main"]))
        # the po2 = 22 LDE plan itself (2^10 chunk, 2^12-row strided stages on 2^15-word tiles), bit for bit
        x = rand_elems(np.random.default_rng(2222), (2, 1 << po2))
        x[1, ::7] = np.uint32(P - 1)
        assert (c.op_lde(x) == orc.expand_ntt(orc.zk_shift(orc.interpolate_ntt(x)), 2)).all()


def test_po2_22_w256_bit_exact_pinned(pkg, gpu_lib, orc, monkeypatch):
    """configs[4]: the po2 = 22, W = 256 seal against pins generated from the CPU oracle (make_golden_large.py)."""
    monkeypatch.setenv("HFB200_DEBUG_CHECKPOINTS", "1")
    gold = _load_pins(22)
    po2, W = gold["po2"], tuple(gold["widths"])
    with pkg.Context(0, po2, W, lib=gpu_lib) as c:
        c.witgen_synth(po2, gold["trace_seed"], gold["blind_seed"])
        seal = c.prove_resident(gold["blind_seed"])
        _check_against_pins(orc, gold, seal, c.checkpoints())


def test_pool_and_concurrent_contexts_on_gpu(pkg, gpu_lib, orc):
    """hfb200_pool on real devices (every visible GPU, two contexts each) and two contexts driven from two host
    threads: seals equal the oracle's regardless of which device / context proved them."""
    import threading
    import torch
    ndev = max(1, torch.cuda.device_count())
    jobs, expect = [], []
    for i, po2 in enumerate([12, 13, 12, 13, 12, 12]):
        cir, g, code, data = make_segment(orc, SMALL, po2, trace_seed=300 + i)
        jobs.append((po2, g, code, data, 40 + i))
        expect.append(cir.prove(po2, g, code, data, 40 + i)[0])
    with pkg.Pool(devices=tuple(range(ndev)), contexts_per_device=2, max_po2=13, circuit=SMALL, lib=gpu_lib) as pool:
        seals, devs, ms = pool.prove(jobs, 40000)
        assert all(len(a) == len(b) and (a == b).all() for a, b in zip(seals, expect))
        assert set(devs) <= set(range(ndev))
        # opt-in control-group reuse through the pool (mixed po2: a worker re-commits the group when its context changes size)
        pool.load_control(12, jobs[0][2])
        pool.load_control(13, jobs[1][2])
        seals2, _, _ = pool.prove([(p_, g_, None, d_, s_) for (p_, g_, c_, d_, s_) in jobs], 40000)
        assert all((a == b).all() for a, b in zip(seals2, expect))
    ctxs = [pkg.Context(0, 13, SMALL, lib=gpu_lib) for _ in range(2)]
    out = [None, None]

    def work(k):
        po2, g, code, data, seed = jobs[k]
        for _ in range(3):
            out[k] = ctxs[k].prove_segment(po2, g, code, data, seed)
    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert (out[0] == expect[0]).all() and (out[1] == expect[1]).all()
    [c.close() for c in ctxs]


@pytest.mark.parametrize("widths,po2", [((21, 40, 12), 13), ((5, 8, 4), 12), ((16, 200, 52), 14)])
def test_odd_circuit_widths(pkg, gpu_lib, orc, widths, po2):
    """Ragged shapes on the GPU: partial sponge blocks, the minimum circuit, non-divisible column groups."""
    cir, g, code, data = make_segment(orc, widths, po2)
    oseal, ocps, _ = cir.prove(po2, g, code, data, 3)
    with pkg.Context(0, po2, widths, lib=gpu_lib) as c:
        seal = c.prove_segment(po2, g, code, data, 3)
        assert len(seal) == len(oseal) and (seal == oseal).all()
        assert cir.verify(seal, ocps["code_root"]) == po2


@pytest.mark.parametrize("widths,po2,variant,nest", [(SMALL, 12, 0, False), ((16, 64, 16), 13, 1, True), (DEFAULT, 14, 1, False)])
def test_data_defined_circuit_on_gpu(pkg, gpu_lib, orc, widths, po2, variant, nest):
    """hfb200_init_ir: tap table + PolyExtStep-shaped constraint list compiled to bytecode, interpreter kernel, generic
    DEEP kernels (tap sets {0},{0,1},{0,1,2}); seal equal to the oracle interpreting the same data."""
    from oracle import synth_ir
    cir = orc.Circuit(*widths, variant=variant)
    code = cir.gen_code(po2); g = cir.gen_globals(5); data = cir.gen_data(po2, code, g, 5, 1)
    ir = synth_ir.build(widths, variant, nest=nest)
    cir_ir = orc.Circuit(*widths, variant=variant)
    cir_ir.set_ir(ir["taps"], ir["steps"], ir["ret"], ir.get("info"))
    oseal, ocps, _ = cir_ir.prove(po2, g, code, data, 1)
    with pkg.Context(0, po2, widths, lib=gpu_lib, ir=ir) as c:
        mix = c.segment_begin(po2, g, code, data, 1)
        assert (mix == ocps["accum_mix"]).all()
        seal = c.segment_finish(cir.step_accum(po2, data, mix, 1))
        cps = c.checkpoints()
        for k, v in ocps.items():
            if k in cps:
                assert (cps[k] == v).all(), k
        assert len(seal) == len(oseal) and (seal == oseal).all()
        assert cir_ir.verify(seal, ocps["code_root"]) == po2


@pytest.mark.gpu
def test_jit_and_interpreter_agree_on_gpu(pkg, gpu_lib, orc, monkeypatch):
    """The NVRTC-specialised eval_check (default) and the interpreter kernel (HFB200_IR_JIT=0) give the same seal, equal
    to the oracle's; the JIT really is the path in use."""
    from oracle import synth_ir
    widths, po2, variant = (16, 64, 16), 13, 1
    cir = orc.Circuit(*widths, variant=variant)
    code = cir.gen_code(po2); g = cir.gen_globals(5); data = cir.gen_data(po2, code, g, 5, 1)
    ir = synth_ir.build(widths, variant, nest=True)
    cir_ir = orc.Circuit(*widths, variant=variant)
    cir_ir.set_ir(ir["taps"], ir["steps"], ir["ret"], ir.get("info"))
    oseal, ocps, _ = cir_ir.prove(po2, g, code, data, 1)
    seals = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("HFB200_IR_JIT", mode)
        with pkg.Context(0, po2, widths, lib=gpu_lib, ir=ir) as c:
            active, ms = c.ir_jit_active()
            assert active == (mode == "1"), (mode, active)
            mix = c.segment_begin(po2, g, code, data, 1)
            seals[mode] = c.segment_finish(cir.step_accum(po2, data, mix, 1))
    assert (seals["1"] == oseal).all() and (seals["0"] == oseal).all()


@pytest.mark.gpu
def test_data_defined_circuit_at_rv32im_scale_on_gpu(pkg, gpu_lib, orc, monkeypatch):
    """~51 k PolyExtSteps, W = 400, ~1100 taps at the declared limits (4 back values, 8 tap sets).  The NVRTC-specialised
    chunk kernels (default) and the interpreter kernel (HFB200_IR_JIT=0) both give the oracle's seal; the pool over a data-defined circuit
    (hfb200_pool_create_ir) refuses one-shot jobs with the documented reason."""
    from oracle import synth_ir
    ir = synth_ir.build_scaled(n_groups=560)
    W, po2 = ir["widths"], 12
    assert len(ir["steps"]) > 50000
    rng = np.random.default_rng(3)
    cir = orc.Circuit(*W)
    code = cir.gen_code(po2); g = cir.gen_globals(9)
    data = rng.integers(0, orc.P, size=(W[1], 1 << po2), dtype=np.uint32)
    cir_ir = orc.Circuit(*W)
    cir_ir.set_ir(ir["taps"], ir["steps"], ir["ret"], ir.get("info"))
    oseal, ocps, _ = cir_ir.prove(po2, g, code, data, 1)
    for mode in ("1", "0"):
        monkeypatch.setenv("HFB200_IR_JIT", mode)
        with pkg.Context(0, po2, W, lib=gpu_lib, ir=ir) as c:
            active, ms = c.ir_jit_active()
            assert active == (mode == "1")
            if active:
                print("rv32im-scale IR: %d steps, NVRTC %.1f s" % (len(ir["steps"]), ms / 1e3))
            mix = c.segment_begin(po2, g, code, data, 1)
            seal = c.segment_finish(cir.step_accum(po2, data, mix, 1))
            assert len(seal) == len(oseal) and (seal == oseal).all(), mode
    monkeypatch.setenv("HFB200_IR_JIT", "0")
    with pkg.Pool(devices=(0,), contexts_per_device=1, max_po2=po2, circuit=W, lib=gpu_lib, ir=ir) as pool:
        _, _, _, errs = pool.prove([(po2, g, code, data, 1)], 1 << 18, return_errors=True)
        assert errs[0] is not None and "step_accum is the caller's" in errs[0]


@pytest.mark.gpu
@pytest.mark.parametrize("widths,po2", [(SMALL, 12), (DEFAULT, 16), ((21, 40, 12), 13)])
def test_device_transcript_on_gpu(pkg, gpu_lib, orc, widths, po2, monkeypatch):
    """The opt-in device-side transcript (warp-cooperative Poseidon2 RNG, seal assembled in HBM, one sync per segment): seal and
    checkpoints equal the oracle's, through the host-buffer entry, the cached control group and the resident entry."""
    monkeypatch.setenv("HFB200_DEBUG_CHECKPOINTS", "1")
    cir, g, code, data = make_segment(orc, widths, po2)
    oseal, ocps, _ = cir.prove(po2, g, code, data, 1)
    with pkg.Context(0, po2, widths, lib=gpu_lib) as c:
        c.set_transcript(True)
        seal = c.prove_segment(po2, g, code, data, 1)
        assert c.last_stats()["host_syncs"] == 1              # the seal (the debug checkpoint's read-back shares that synchronisation)
        cps = c.checkpoints()
        for k, v in ocps.items():
            assert (cps[k] == v).all(), k
        assert len(seal) == len(oseal) and (seal == oseal).all()
        c.control_root(po2, code)
        assert (c.prove_segment(po2, g, None, data, 1) == oseal).all()
        c.witgen_synth(po2, TRACE_SEED, 1)
        assert (c.prove_resident(1) == oseal).all()


@pytest.mark.gpu
def test_device_transcript_headline_pins(pkg, gpu_lib, orc, monkeypatch):
    monkeypatch.setenv("HFB200_DEBUG_CHECKPOINTS", "1")
    gold = _load_pins(20)
    with pkg.Context(0, 20, tuple(gold["widths"]), lib=gpu_lib) as c:
        c.set_transcript(True)
        c.witgen_synth(20, gold["trace_seed"], gold["blind_seed"])
        seal = c.prove_resident(gold["blind_seed"])
        _check_against_pins(orc, gold, seal, c.checkpoints())


@pytest.mark.gpu
def test_cuda_graph_replay_on_gpu(pkg, gpu_lib, orc):
    """Transcript mode 2: the device-transcript segment captured once as a CUDA graph and replayed.  Every per-segment value
    (globals, blinding key, trace, control reuse) must reach the replayed kernels: segments with DIFFERENT traces, globals and
    blind seeds proved through one instantiated graph equal the oracle's seals word for word, for host buffers, the cached
    control group (a second graph) and the resident entry, across a po2 switch and back."""
    widths = (16, 64, 16)
    cir = orc.Circuit(*widths)
    def inputs(po2, trace_seed, blind):
        _, g, code, data = make_segment(orc, widths, po2, trace_seed=trace_seed, blind_seed=blind)
        return g, code, data, cir.prove(po2, g, code, data, blind)[0]
    with pkg.Context(0, 13, widths, lib=gpu_lib, deterministic=True) as c:
        c.set_transcript(2)
        for i, (po2, ts, blind) in enumerate([(13, 11, 1), (13, 12, 2), (13, 13, 3), (13, 14, 4), (12, 21, 5), (12, 22, 6), (12, 23, 7), (13, 15, 8)]):
            g, code, data, oseal = inputs(po2, ts, blind)
            seal = c.prove_segment(po2, g, code, data, blind)
            assert len(seal) == len(oseal) and (seal == oseal).all(), (i, po2)
            assert c.last_stats()["host_syncs"] == 1
        n_host = c.graph_launches()
        assert n_host >= 5                                  # per shape: one plain run, one capture (already a graph launch), then replays
        g, code, data, oseal = inputs(13, 31, 9)
        c.control_root(13, code)
        for blind in (9, 10, 11):                           # cached control group: its own graph
            _, _, _, oseal = inputs(13, 31, blind)
            _, g2, code2, data2 = make_segment(orc, widths, 13, trace_seed=31, blind_seed=blind)
            assert (c.prove_segment(13, g2, None, data2, blind) == oseal).all()
        assert c.graph_launches() >= n_host + 2
        c.witgen_synth(13, TRACE_SEED, 1)                   # resident entry: the data columns carry blind seed 1, the accum noise the call's seed
        _, g3, code3, data3 = make_segment(orc, widths, 13)
        for blind in (1, 1):
            assert (c.prove_resident(blind) == cir.prove(13, g3, code3, data3, blind)[0]).all()
        c.set_transcript(0)                                 # back to the host transcript: same seal
        assert (c.prove_resident(1) == cir.prove(13, g3, code3, data3, 1)[0]).all()


@pytest.mark.gpu
def test_chunk_stage_kernels_agree():
    """The fused chunk NTT stage exists twice: `mid_warp_kernel` (one warp per transform, four-warp teams; the default) and
    `MiddleKernel2` (the host emulator's kernel, HFB200_MID_WARP=0).  Exact field arithmetic: the LDE of the same random columns
    must be bit-identical through both, for full and ragged quads of columns, 2^11 .. 2^21 rows (tools/lde_stress.py --digests)."""
    import subprocess
    import sys
    tool = os.path.join(ROOT, "tools", "lde_stress.py")
    out = {}
    for mode in ("0", "1"):
        env = dict(os.environ, HFB200_MID_WARP=mode)
        out[mode] = subprocess.check_output([sys.executable, tool, "--digests"], env=env, timeout=600).decode().split()
    assert len(out["1"]) == 8 and out["0"] == out["1"]
