#!/usr/bin/env python3
"""Race hunt for the chunk NTT stage (compute-sanitizer is closed on the pool): the fused LDE of random columns, repeated many times
on several contexts at once, must be bit-identical every time and identical to the MiddleKernel2 path (HFB200_MID_WARP=0, run in a
child process).  Usage: lde_stress.py [seconds=30] [contexts=4]"""
import hashlib
import os
import subprocess
import sys
import threading
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

SHAPES = [(20, 7), (20, 16), (19, 5), (21, 3), (16, 48), (12, 9), (11, 4), (13, 192)]


def digests():
    import hfb200_loader
    pkg = hfb200_loader.load()
    out = []
    with pkg.Context(0, 20, (16, 192, 48)) as c:
        for lg, ncols in SHAPES:
            x = np.random.default_rng(lg * 1000 + ncols).integers(0, pkg.P, size=(ncols, 1 << lg), dtype=np.uint32)
            out.append(hashlib.sha256(c.op_lde(x).tobytes()).hexdigest())
    return out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--digests":
        print("\n".join(digests()))
        sys.exit(0)
    seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 30.0
    F = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    env = dict(os.environ, HFB200_MID_WARP="0")
    ref = subprocess.check_output([sys.executable, os.path.abspath(__file__), "--digests"], env=env).decode().split()
    import hfb200_loader
    pkg = hfb200_loader.load()
    errs, counts = [], [0] * F

    def work(slot):
        try:
            with pkg.Context(0, 20, (16, 192, 48)) as c:
                t_end = time.time() + seconds
                k = slot
                while time.time() < t_end:
                    lg, ncols = SHAPES[k % len(SHAPES)]
                    x = np.random.default_rng(lg * 1000 + ncols).integers(0, pkg.P, size=(ncols, 1 << lg), dtype=np.uint32)
                    d = hashlib.sha256(c.op_lde(x).tobytes()).hexdigest()
                    if d != ref[k % len(SHAPES)]:
                        errs.append((slot, k, lg, ncols))
                    counts[slot] += 1
                    k += 1
        except Exception as e:  # noqa: BLE001
            errs.append((slot, repr(e)))

    th = [threading.Thread(target=work, args=(i,)) for i in range(F)]
    [t.start() for t in th]; [t.join() for t in th]
    print("lde stress: %d LDEs on %d contexts in %.0f s, mismatches against the MiddleKernel2 path: %d %s" % (sum(counts), F, seconds, len(errs), errs[:5]))
    sys.exit(1 if errs else 0)
