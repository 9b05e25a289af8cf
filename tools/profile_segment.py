#!/usr/bin/env python3
"""Proves `reps` resident po2 segments (used under ncu: launch list / --set full captures)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hfb200_loader

pkg = hfb200_loader.load()
po2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
with pkg.Context(0, po2, (16, 192, 48)) as ctx:
    ctx.witgen_synth(po2, 0x48595046, 1)
    for i in range(reps):
        seal = ctx.prove_resident(1 + i)
    print("ok", len(seal), ctx.last_stats())
