import sys, time, threading
sys.path.insert(0, "/root/repo")
import numpy as np, hfb200_loader
pkg = hfb200_loader.load()
po2 = 20
for S in (1, 2, 3):
    ctxs = [pkg.Context(0, po2, (16, 192, 48)) for _ in range(S)]
    for i, c in enumerate(ctxs):
        c.witgen_synth(po2, 0x48595046 + i, 1)
        c.prove_resident(1)
    K = 6
    def work(c):
        for k in range(K):
            c.prove_resident(1 + k)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work, args=(c,)) for c in ctxs]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print("contexts", S, "segments/s", S * K / dt, "ms per segment", dt / (S * K) * 1e3)
    # e2e with host buffers
    hb = []
    for c in ctxs:
        code = c.host_alloc((16, 1 << po2)); data = c.host_alloc((192, 1 << po2))
        code[...] = c.read_group(1); data[...] = c.read_group(2)
        hb.append((code, data))
    g = ctxs[0].witgen_synth(po2, 0x48595046, 1)
    def work2(c, code, data):
        for k in range(K):
            c.prove_segment(po2, g, code, data, 1 + k)
    t0 = time.perf_counter()
    th = [threading.Thread(target=work2, args=(c, *hb[i])) for i, c in enumerate(ctxs)]
    [t.start() for t in th]; [t.join() for t in th]
    dt = time.perf_counter() - t0
    print("   e2e contexts", S, "segments/s", S * K / dt)
    for c in ctxs: c.close()
