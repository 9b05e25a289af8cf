import os, sys, time
sys.path.insert(0, "/root/repo")
import hfb200_loader
pkg = hfb200_loader.load()
for po2 in (14, 16, 18, 20):
    with pkg.Context(0, po2, (16, 192, 48)) as c:
        c.witgen_synth(po2, 0x48595046, 1)
        for k in range(3): c.prove_resident(1)
        K = 30 if po2 <= 16 else 8
        t0 = time.perf_counter()
        for k in range(K): c.prove_resident(1)
        print("warp_max", os.environ.get("HFB200_FOLD_WARP_MAX"), "po2", po2, "ms per segment %.3f" % ((time.perf_counter() - t0) / K * 1e3), flush=True)
