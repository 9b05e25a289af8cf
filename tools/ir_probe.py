#!/usr/bin/env python3
"""Built-in kernels vs the data-defined (interpreted) circuit on the same po2 segment: seal equality and stage times."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hfb200_loader
from oracle import synth_ir   # test infrastructure: only builds the tables

pkg = hfb200_loader.load()
po2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
W = (16, 192, 48)
ir = synth_ir.build(W, 0)
with pkg.Context(0, po2, W) as a, pkg.Context(0, po2, W, ir=ir) as b:
    g = a.witgen_synth(po2, 0x48595046, 1)
    code, data = a.read_group(1), a.read_group(2)
    for _ in range(2):
        sa = a.prove_resident(1)
    accum = a.read_group(0)
    for _ in range(2):
        b.segment_begin(po2, g, code, data, 1)
        sb = b.segment_finish(accum)
    print("seal equal:", (sa == sb).all(), "steps", len(ir["steps"]), "jit (active, compile ms):", b.ir_jit_active())
    print("built-in :", {k: round(v, 2) for k, v in a.last_stats().items() if k.startswith("ms_")})
    print("data-def.:", {k: round(v, 2) for k, v in b.last_stats().items() if k.startswith("ms_")})
