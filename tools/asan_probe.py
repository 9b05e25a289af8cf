import sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
