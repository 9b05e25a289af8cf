import sys
import os; os.environ.setdefault("HFB200_DETERMINISTIC_BLINDING", "1"); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, hfb200_loader
pkg = hfb200_loader.load()
lib = pkg.load_library("/tmp/libhfb200_emu_asan.so")
rng = np.random.default_rng(0)
with pkg.Context(0, 14, (16, 64, 16), lib=lib) as ctx:
    for lg in (1, 3, 5, 8, 10, 11, 12, 14, 16):
        x = rng.integers(0, pkg.P, size=(2, 1 << lg), dtype=np.uint32)
        ctx.op_lde(x); ctx.op_interpolate_ntt(x, True); ctx.op_expand_ntt(x, 0); ctx.op_expand_ntt(x, 2)
    for rows, cols in ((2, 1), (16, 3), (1024, 17), (4096, 64)):
        ctx.op_merkle(rng.integers(0, pkg.P, size=(cols, rows), dtype=np.uint32))
    for po2 in (12, 13, 14):
        g = ctx.witgen_synth(po2, 7, 1)
        seal = ctx.prove_resident(1)
        code, data = ctx.read_group(1), ctx.read_group(2)
        assert (seal == ctx.prove_segment(po2, g, code, data, 1)).all()
print("asan probe ok")
# wider circuits (several DotKernel column groups, ragged widths), control-group reuse, the product verifier on good and bad seals
for widths, po2 in (((16, 192, 48), 12), ((21, 72, 20), 12)):
    with pkg.Context(0, po2, widths, lib=lib) as ctx:
        g = ctx.witgen_synth(po2, 9, 1)
        code, data = ctx.read_group(1), ctx.read_group(2)
        seal = ctx.prove_segment(po2, g, code, data, 3)
        root = ctx.control_root(po2, code)
        assert (ctx.prove_segment(po2, g, None, data, 3) == seal).all()
        assert pkg.verify_segment(seal, root, widths, lib=lib) == po2
        for k in range(40):
            bad = seal.copy()
            if k % 2:
                bad = bad[:int(rng.integers(0, len(bad)))]
            else:
                bad[int(rng.integers(0, len(bad)))] = np.uint32(rng.integers(0, 1 << 32, dtype=np.uint64))
            try:
                pkg.verify_segment(bad, root, widths, lib=lib)
            except pkg.Hfb200Error:
                pass
print("asan probe 2 ok")
# round 2: device transcript, pool fault handling + control generations, claims, chunked data-defined circuits
os.environ["HFB200_IR_CHUNK"] = "200"
from oracle import synth_ir   # table builder only
with pkg.Context(0, 13, (16, 64, 16), lib=lib, deterministic=True) as ctx:
    g = ctx.witgen_synth(13, 5, 2)
    code, data = ctx.read_group(1), ctx.read_group(2)
    host = ctx.prove_segment(13, g, code, data, 2)
    ctx.set_transcript(True)
    assert (ctx.prove_segment(13, g, code, data, 2) == host).all() and (ctx.prove_resident(2) == host).all()
with pkg.Pool(devices=(0, 0), contexts_per_device=2, max_po2=13, circuit=(16, 64, 16), lib=lib, deterministic=True) as pool:
    pool.inject_fault(0, 0, 0)
    pool.load_control(13, code)
    seals, _, _ = pool.prove([(13, g, code, data, 2), (13, g, None, data, 2), (13, g, code, data, 2)], 1 << 17)
    assert all((s == host).all() for s in seals)
    keep = code.copy()
    pool.load_control(13, keep)
    seals, _, _ = pool.prove([(13, g, None, data, 2)], 1 << 17)
    assert (seals[0] == host).all()
ir = synth_ir.build((16, 64, 16), 1, nest=True)
with pkg.Context(0, 12, (16, 64, 16), lib=lib, ir=ir, deterministic=True) as ctx:
    rng2 = np.random.default_rng(2)
    code = rng2.integers(0, pkg.P, size=(16, 4096), dtype=np.uint32); data = rng2.integers(0, pkg.P, size=(64, 4096), dtype=np.uint32)
    accum = rng2.integers(0, pkg.P, size=(16, 4096), dtype=np.uint32)
    ctx.segment_begin(12, g, code, data, 1); ctx.segment_finish(accum)
big = synth_ir.build_scaled(n_groups=40)
with pkg.Context(0, 12, big["widths"], lib=lib, ir=big, deterministic=True) as ctx:
    W = big["widths"]
    code = rng2.integers(0, pkg.P, size=(W[0], 4096), dtype=np.uint32); data = rng2.integers(0, pkg.P, size=(W[1], 4096), dtype=np.uint32)
    accum = rng2.integers(0, pkg.P, size=(W[2], 4096), dtype=np.uint32)
    ctx.segment_begin(12, g, code, data, 1); ctx.segment_finish(accum)
d = pkg.digest_bytes(b"journal", lib=lib); pkg.digest_pair(d, d, lib=lib); pkg.claim_next_state(d, 3, 20, lib=lib)
# round 2, second half: batch verifier on host threads, graph-replay mode (plain device transcript in the emulator), circuit info
cir_seals = [host, host.copy(), host]
with pkg.Context(0, 13, (16, 64, 16), lib=lib, deterministic=True) as ctx:
    g = ctx.witgen_synth(13, 5, 2)
    code = ctx.read_group(1)
    root = ctx.control_root(13, code)
    assert pkg.verify_segments(cir_seals, [root] * 3, (16, 64, 16), threads=3, lib=lib) == [13, 13, 13]
    bad = host.copy(); bad[300] ^= 1
    try:
        pkg.verify_segments([host, bad], [root] * 2, (16, 64, 16), threads=2, lib=lib)
        raise SystemExit("tampered seal accepted")
    except pkg.Hfb200Error:
        pass
    ctx.set_transcript(2)
    ctx.witgen_synth(13, 5, 2)
    assert (ctx.prove_resident(2) == host).all() and (ctx.prove_resident(2) == host).all()
print("asan probe 3 ok")
