#!/usr/bin/env python3
"""Times the LDE pipeline and the Merkle build of the data group on resident columns (kernel experiments).
Usage: bench_kernels.py [lib.so ...]  -- one subprocess per library (two copies of the library cannot share a process)."""
import os
import subprocess
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one(path):
    import hfb200_loader
    pkg = hfb200_loader.load()
    lib = pkg.load_library(path) if path else pkg.load_library()
    with pkg.Context(0, 20, (16, 192, 48), lib=lib) as ctx:
        ctx.witgen_synth(20, 0x48595046, 1)
        ctx.bench_lde(20, 192, 1)
        print(os.path.basename(path or pkg.LIB_PATH), "lde192 ms", round(ctx.bench_lde(20, 192, 5), 3), "merkle192 ms", round(ctx.bench_merkle(20, 192, 3), 3),
              "merkle16 ms", round(ctx.bench_merkle(20, 16, 3), 3), flush=True)
        ctx.prove_resident(1)
        ctx.prove_resident(1)
        print("   segment:", {k: round(v, 2) for k, v in ctx.last_stats().items()}, flush=True)


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--one":
        one(sys.argv[2])
    else:
        for p in sys.argv[1:] or [""]:
            subprocess.call([sys.executable, os.path.abspath(__file__), "--one", p])
