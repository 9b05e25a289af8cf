#!/usr/bin/env python3
"""One fused trace -> LDE pass over 192 resident columns at po2 = 20 (ncu target).  usage: lde_once.py [lib.so] [iters]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hfb200_loader
pkg = hfb200_loader.load()
lib = pkg.load_library(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1] else pkg.load_library()
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with pkg.Context(0, 20, (16, 192, 48), lib=lib, deterministic=True) as ctx:
    ctx.witgen_synth(20, 0x48595046, 1)
    print("lde192 ms", ctx.bench_lde(20, 192, iters))
