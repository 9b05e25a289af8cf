#!/usr/bin/env python3
"""Regenerates tests/golden/golden_small.json from the CPU oracle.

The reference ships NO golden vector for this path (both receipts under /root/reference/data/test are dev-mode
fakes), and the upstream crates are not importable here, so these vectors are SELF-GENERATED regression pins:
they freeze today's oracle output so that oracle, emulator and CUDA path can all be checked against the same
committed numbers.  They do not pin parity with risc0 3.0.5 ("parity unpinned")."""
import json
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402


def main():
    out = {}
    out["poseidon2_zero"] = oracle.poseidon2_mix(np.zeros(24, np.uint32)).tolist()
    out["poseidon2_iota_mont"] = oracle.poseidon2_mix(oracle.encode(np.arange(24))).tolist()
    out["hash_empty"] = oracle.hash_elems(np.zeros(0, np.uint32)).tolist()
    out["hash_16"] = oracle.hash_elems(oracle.encode(np.arange(16))).tolist()
    out["hash_17"] = oracle.hash_elems(oracle.encode(np.arange(17))).tolist()
    x = oracle.encode(np.arange(1, 17))
    out["intt16_of_1_to_16"] = oracle.interpolate_ntt(x).tolist()
    out["lde16_of_1_to_16"] = oracle.expand_ntt(oracle.zk_shift(oracle.interpolate_ntt(x)), 2).ravel().tolist()
    W, po2 = (8, 16, 8), 12
    cir = oracle.Circuit(*W)
    code = cir.gen_code(po2)
    g = cir.gen_globals(0x48595046)
    data = cir.gen_data(po2, code, g, 0x48595046, 1)
    seal, cps, _ = cir.prove(po2, g, code, data, 1)
    out["segment"] = {"widths": list(W), "po2": po2, "trace_seed": 0x48595046, "blind_seed": 1, "seal_words": int(len(seal)),
                      "seal_hash": oracle.hash_elems(seal % oracle.P).tolist(),
                      "checkpoints": {k: v.tolist() for k, v in cps.items()}}
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_small.json"), "w") as f:
        json.dump(out, f, indent=0, sort_keys=True)
    print("wrote golden_small.json")


if __name__ == "__main__":
    main()
