#!/usr/bin/env python3
"""Times the LDE pipeline and the Merkle build of the data group on resident columns (kernel experiments)."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hfb200_loader

pkg = hfb200_loader.load()
libs = sys.argv[1:] or [pkg.LIB_PATH]
for path in libs:
    lib = pkg.load_library(path)
    with pkg.Context(0, 20, (16, 192, 48), lib=lib) as ctx:
        ctx.witgen_synth(20, 0x48595046, 1)
        ctx.bench_lde(20, 192, 1)
        print(os.path.basename(path), "lde192 ms", round(ctx.bench_lde(20, 192, 3), 3), "merkle192 ms", round(ctx.bench_merkle(20, 192, 3), 3),
              "merkle16 ms", round(ctx.bench_merkle(20, 16, 3), 3))
