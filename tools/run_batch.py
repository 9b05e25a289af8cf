#!/usr/bin/env python3
"""BASELINE.json configs[3]: a batch of 64 synthetic camt53 statements (statement i = 30 + (i mod 16) segments of
po2 = 20) proved through ONE host process with hfb200_pool over all visible GPUs.  All segments share one host trace
(872 MB pinned) and differ by their blinding seed, which is what varies between segments of equal shape anyway.

  python tools/run_batch.py [--statements 64] [--po2 20] [--contexts 2] [--limit N]
"""
import argparse
import importlib
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hfb200_loader

ap = argparse.ArgumentParser()
ap.add_argument("--statements", type=int, default=64)
ap.add_argument("--po2", type=int, default=20)
ap.add_argument("--contexts", type=int, default=2)
ap.add_argument("--limit", type=int, default=0, help="prove only the first N segments of the batch")
ap.add_argument("--reuse-control", action="store_true", help="opt-in: commit the control group once per context (hfb200_pool_load_control), jobs pass code = NULL")
args = ap.parse_args()

import torch
pkg = hfb200_loader.load()
sched = importlib.import_module("hyperfridge_r0_b200.scheduler")
ndev = torch.cuda.device_count()
W = (16, 192, 48)
batch = sched.make_batch(args.statements, 30, 16, args.po2)
if args.limit:
    batch = batch[:args.limit]
with pkg.Context(0, args.po2, W) as c:
    g = c.witgen_synth(args.po2, 0x48595046, 1)
    code_h = c.host_alloc((W[0], 1 << args.po2)); data_h = c.host_alloc((W[1], 1 << args.po2))
    code_h[...] = c.read_group(1); data_h[...] = c.read_group(2)
    cap = c.seal_words(args.po2)
jobs = [(j["po2"], g, code_h, data_h, sched.job_seed(1, j["statement"], j["segment"])) for j in batch]
with pkg.Pool(devices=tuple(range(ndev)), contexts_per_device=args.contexts, max_po2=args.po2, circuit=W) as pool:
    if args.reuse_control:
        pool.load_control(args.po2, code_h)
        jobs = [(p_, g_, None, d_, s_) for (p_, g_, c_, d_, s_) in jobs]
    pool.prove(jobs[:2 * ndev * args.contexts], cap)  # warm-up
    t0 = time.perf_counter()
    seals, devs, ms = pool.prove(jobs, cap)
    dt = time.perf_counter() - t0
per_dev = {d: devs.count(d) for d in sorted(set(devs))}
print(json.dumps({"workload": "configs[3]: %d statements, %d segments of po2=%d" % (args.statements, len(jobs), args.po2), "gpus": ndev,
                  "contexts_per_gpu": args.contexts, "reuse_control": bool(args.reuse_control), "seconds": dt, "segments_per_s": len(jobs) / dt, "segments_per_gpu": per_dev,
                  "seal_words": int(len(seals[0])), "mean_job_ms": float(np.mean(ms))}))
