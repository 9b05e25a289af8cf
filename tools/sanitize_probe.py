#!/usr/bin/env python3
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): ops + two segments at po2 = 12, 13."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hfb200_loader

pkg = hfb200_loader.load()
rng = np.random.default_rng(0)
with pkg.Context(0, 13, (16, 64, 16)) as ctx:
    for lg in (5, 10, 12, 14):
        x = rng.integers(0, pkg.P, size=(2, 1 << lg), dtype=np.uint32)
        ctx.op_lde(x); ctx.op_interpolate_ntt(x, True); ctx.op_expand_ntt(x, 0)
    ctx.op_merkle(rng.integers(0, pkg.P, size=(17, 1024), dtype=np.uint32))
    for po2 in (12, 13):
        g = ctx.witgen_synth(po2, 7, 1)
        seal = ctx.prove_resident(1)
        code, data = ctx.read_group(1), ctx.read_group(2)
        seal2 = ctx.prove_segment(po2, g, code, data, 1)
        assert (seal == seal2).all()
print("sanitize probe ok")
