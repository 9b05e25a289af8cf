#!/usr/bin/env python3
"""Soak test: several contexts proving concurrently for a while; every seal must equal the first seal produced for the
same (context, po2, blinding seed) -- catches rare races between contexts (stream / attribute / table sharing).
Usage: soak.py [seconds=40] [contexts=4]"""
import os
os.environ.setdefault("HFB200_DETERMINISTIC_BLINDING", "1")  # seals must be reproducible for the comparison
import sys
import threading
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hfb200_loader

pkg = hfb200_loader.load()
seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 40.0
F = int(sys.argv[2]) if len(sys.argv) > 2 else 4
W = (16, 192, 48)
PO2S = (20, 18, 20, 16)
errs, counts = [], [0] * F


def work(slot):
    try:
        with pkg.Context(0, 20, W) as c:
            ref = {}
            host = {}
            t_end = time.time() + seconds
            k = 0
            while time.time() < t_end:
                po2 = PO2S[(k // 6) % len(PO2S)]
                if k % 6 == 0:
                    g = c.witgen_synth(po2, 1000 + slot, 1)
                    host[po2] = (g, c.read_group(1), c.read_group(2))
                    c.witgen_synth(po2, 1000 + slot, 1)
                seed = 1 + (k % 2)
                if k % 3 == 2:
                    g, code, data = host[po2]
                    seal = c.prove_segment(po2, g, code, data, seed)   # host path (chunked H2D)
                    c.witgen_synth(po2, 1000 + slot, 1)
                else:
                    seal = c.prove_resident(seed)
                key = (po2, seed)
                if key not in ref:
                    ref[key] = seal.copy()
                elif not np.array_equal(ref[key], seal):
                    raise AssertionError("context %d: seal %r differs at iteration %d" % (slot, key, k))
                counts[slot] += 1
                k += 1
    except Exception as e:  # noqa: BLE001
        errs.append(e)


th = [threading.Thread(target=work, args=(i,)) for i in range(F)]
[t.start() for t in th]
[t.join() for t in th]
if errs:
    raise errs[0]
print("soak ok: %d contexts, %s segments in %.0f s, all seals reproducible" % (F, counts, seconds))
