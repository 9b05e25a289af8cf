// Integer-multiply roofline probe: sustained throughput of the modular-multiply instruction sequences the prover is
// made of (the Montgomery product of field.cuh, the Shoup constant product, the signed S-box chain of poseidon2.cuh),
// 8 independent chains per thread.  bench.py reports Poseidon2 against this measured peak: nothing on this path is a
// dense float contraction, and on sm_100a every 32-bit integer multiply runs on the heavy half of the FMA pipe only.
#pragma once
#include "poseidon2.cuh"

namespace hf {

struct ModmulProbeKernel {
    static constexpr bool kBarrier = false;
    static constexpr int ILP = 8;
    // kind 0: Montgomery fmul; 1: Shoup fmul_shoup; 2: (s + rc)^7 signed chain (4 products)
    HD static void run(const KCtx& cx, uint32_t*, uint32_t* out, uint32_t seed, uint32_t iters, int kind) {
        uint32_t x[ILP], y[ILP];
#pragma unroll
        for (int i = 0; i < ILP; i++) { x[i] = (seed + (uint32_t)cx.tid * 7u + (uint32_t)i * 13u) % P; y[i] = (seed * 3u + (uint32_t)i + (uint32_t)cx.tid * 11u + cx.bx) % P; }
        if (kind == 0) {
#pragma unroll 1
            for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
                for (int i = 0; i < ILP; i++) x[i] = fmul(x[i], y[i]);
            }
        } else if (kind == 1) {
#pragma unroll 1
            for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
                for (int i = 0; i < ILP; i++) x[i] = fmul_shoup(x[i], y[i], y[(i + 1) % ILP]);
            }
        } else {
#pragma unroll 1
            for (uint32_t it = 0; it < iters; it++) {
#pragma unroll
                for (int i = 0; i < ILP; i++) x[i] = sbox7_rc(x[i], y[i]);
            }
        }
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < ILP; i++) acc ^= x[i];
        out[(uint64_t)cx.bx * cx.nt + cx.tid] = acc;
    }
};

}  // namespace hf
