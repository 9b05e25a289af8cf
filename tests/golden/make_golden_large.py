#!/usr/bin/env python3
"""Regenerates tests/golden/golden_po2_<P>_w256.json from the CPU oracle: the HEADLINE configuration of bench.py
(BASELINE.json configs[1]: synthetic segment, W = 256 = code 16 + data 192 + accum 48, trace seed 0x48595046, blind
seed 1) at po2 = 20 and, for configs[4], po2 = 22.  Each file pins every transcript checkpoint (Merkle roots,
challenges, hash_u, final_poly_hash, FRI roots/mixes, query positions), the Poseidon2 hash of the whole seal, the seal
length and the first / last 64 seal words, so the GPU test compares the CUDA path with the oracle at full size in
milliseconds instead of re-running the oracle (~1 min at po2 = 20, several minutes and ~35 GB at po2 = 22).

Like golden_small.json these are SELF-GENERATED regression pins of the oracle (the reference holds no STARK vector of
this path: both receipts under /root/reference/data/test are dev-mode fakes) -- "parity unpinned" against risc0 3.0.5
still applies.

usage: python tests/golden/make_golden_large.py [po2 ...]   (default: 20 22)"""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

W = (16, 192, 48)
TRACE_SEED = 0x48595046
BLIND_SEED = 1


def make(po2):
    cir = oracle.Circuit(*W)
    t0 = time.time()
    code = cir.gen_code(po2)
    g = cir.gen_globals(TRACE_SEED)
    data = cir.gen_data(po2, code, g, TRACE_SEED, BLIND_SEED)
    t1 = time.time()
    seal, cps, _ = cir.prove(po2, g, code, data, BLIND_SEED)
    t2 = time.time()
    out = {"widths": list(W), "po2": po2, "trace_seed": TRACE_SEED, "blind_seed": BLIND_SEED,
           "seal_words": int(len(seal)),
           "seal_hash": oracle.hash_elems(seal % oracle.P).tolist(),
           "seal_head": seal[:64].tolist(), "seal_tail": seal[-64:].tolist(),
           "seal_sum_u64": int(seal.astype(np.uint64).sum()),
           "checkpoints": {k: v.tolist() for k, v in cps.items()},
           "oracle_seconds": {"witgen": round(t1 - t0, 2), "prove": round(t2 - t1, 2), "threads": os.cpu_count()}}
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_po2_%d_w256.json" % po2)
    with open(path, "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote %s (prove %.1f s)" % (path, t2 - t1), flush=True)


if __name__ == "__main__":
    oracle.build()
    for p in ([int(a) for a in sys.argv[1:]] or [20, 22]):
        make(p)
