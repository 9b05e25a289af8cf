#!/usr/bin/env python3
"""One Poseidon2 Merkle build over the 192-column data group at po2 = 20 (ncu target).  usage: merkle_once.py [iters]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hfb200_loader
pkg = hfb200_loader.load()
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 1
with pkg.Context(0, 20, (16, 192, 48), deterministic=True) as ctx:
    ctx.witgen_synth(20, 0x48595046, 1)
    print("merkle192 ms", ctx.bench_merkle(20, 192, iters))
