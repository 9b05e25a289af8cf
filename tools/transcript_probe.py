#!/usr/bin/env python3
"""Wall-clock milliseconds per resident segment for the three transcript modes (0 host, 1 device, 2 device + CUDA-graph replay),
one context, and segments/s at po2 = 20 with 1 / 2 / 4 contexts in flight.  Usage: transcript_probe.py [po2 ...]"""
import os
import sys
import threading
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hfb200_loader
pkg = hfb200_loader.load()
W = (16, 192, 48)
po2s = [int(a) for a in sys.argv[1:]] or [12, 14, 16, 18, 20]
for po2 in po2s:
    row = []
    with pkg.Context(0, po2, W) as c:
        c.witgen_synth(po2, 0x48595046, 1)
        for mode in (0, 1, 2):
            c.set_transcript(mode)
            for k in range(4):
                c.prove_resident(1 + k)
            K = 40 if po2 <= 16 else 10
            t0 = time.perf_counter()
            for k in range(K):
                c.prove_resident(1 + k)
            row.append((time.perf_counter() - t0) / K * 1e3)
        print("po2 %2d  ms per segment: host transcript %.3f  device %.3f  device + graph %.3f  (graph launches %d)" % (po2, row[0], row[1], row[2], c.graph_launches()), flush=True)
po2 = 20
for S in (1, 2, 4):
    ctxs = [pkg.Context(0, po2, W) for _ in range(S)]
    for i, c in enumerate(ctxs):
        c.witgen_synth(po2, 0x48595046 + i, 1)
    out = []
    for mode in (0, 1, 2):
        for c in ctxs:
            c.set_transcript(mode)
            for k in range(3):
                c.prove_resident(1 + k)
        K = 6
        def work(c):
            for k in range(K):
                c.prove_resident(1 + k)
        t0 = time.perf_counter()
        th = [threading.Thread(target=work, args=(c,)) for c in ctxs]
        [t.start() for t in th]; [t.join() for t in th]
        out.append(S * K / (time.perf_counter() - t0))
    print("po2 20, %d contexts in flight: segments/s host %.3f  device %.3f  device + graph %.3f" % (S, out[0], out[1], out[2]), flush=True)
    for c in ctxs:
        c.close()
