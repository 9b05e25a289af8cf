// Fused middle stage of the main-group LDE (chunk iNTT . zk_shift . x4 expand . chunk NTT, 2^10 -> 2^12 per chunk) with ONE WARP
// per (column, chunk) and radix-32 register rounds (sm_100a only).
//
// MiddleKernel2<10, 2> (ntt.cuh) gives a (column, chunk) to a 64-thread unit with 16 values per thread: radix-8 / radix-16 rounds,
// five shared-memory round trips, 294 executed instructions per trace element.  Its one-warp experiment (MID_R5) cut that to 249 but
// kept all four cosets of the expanded chunk (16.5 KB) per warp: 8 warps per SM, issue rate down, no gain (profiles/r2_mid_r5_*).
// This kernel keeps the radix-32 structure and gets 16 warps per SM:
//   * the four cosets of the x4 expansion (index = 4 c + r) are INDEPENDENT 2^10-point transforms once the two replication levels
//     are skipped, so they are produced two at a time: the per-warp buffer is 2 x 1024 words (8.4 KB) instead of 4 x 1024;
//   * phase 1  inverse levels 10..6 on elements lane + 32 j (global -> registers, x w^-(rev(hi) lo)), twiddles from the per-level table;
//     phase 2  inverse levels 5..1 on elements 32 lane + j: every twiddle is a compile-time constant -> __constant__ operands, then
//              x n^-1 3^j (zk_shift);
//     phase 3  per coset r: forward levels 3..7 on the same 32 registers, constant twiddles again;
//     phase 4  per coset r: forward levels 8..12 on elements lane + 32 j, per-level table, x w^(rev(hi) lo'), and both cosets of the
//              pair leave as one 8-byte store per element (two passes fill every 16-byte group of the output);
//   * butterfly outputs that only feed a product (the last level of phases 2 and 4) skip their range correction.
// Same tables, same arithmetic, same output as MiddleKernel2 (exact field arithmetic: the LDE is bit-identical); the host emulator
// keeps MiddleKernel2.  Replaces the same risc0-zkp `Hal` ops as ntt.cuh.
#pragma once
#ifndef HFB200_EMU
#include "ntt.cuh"

namespace hf {

static constexpr int MW_WARPS = 16;
static constexpr uint32_t MW_PLANE = 1024 + 32;                                    // one coset, one pad word per 32
static constexpr uint32_t MW_GS_PAIRS = 1024 + 32;                                  // one pad pair per 32: lane L reads pairs 33 L + j
static constexpr uint32_t MW_TABLE_WORDS = 2 * (1024 + 1024 + MW_GS_PAIRS + 4096 + 4096);  // twI, G3, Gs, tw4, G2 as (w, w') pairs
static constexpr size_t MW_SMEM = (size_t)(MW_TABLE_WORDS + MW_WARPS * 2 * MW_PLANE) * 4;

// constant twiddles of the register-local levels, canonical (w, w' = floor(w 2^32 / p)) pairs:
//   inv[s]    = w_{2^q}^-x        for slot s = 2^(q-1) + x, x < 2^(q-1), q = 1..5      (inverse levels 5..1)
//   fwd[r][s] = w_{2^(q+2)}^(4x+r)  same slots                                          (forward levels 3..7 of coset r)
struct MidWarpConsts { uint32_t inv[32][2]; uint32_t fwd[4][32][2]; };
__constant__ MidWarpConsts g_mwc;

struct MidWarpArgs {
    const uint32_t* in;
    uint32_t* out;
    uint64_t in_stride, out_stride;
    uint32_t ncols, cols_per_block;
    int n, b;        // 2^n coefficients per column = 2^b chunks of 2^10
    uint32_t n_inv;  // Montgomery form of (2^n)^-1
    RootTables rt;
};

// The two coset pairs of an element reach global memory as two 8-byte stores ~10 us apart: the first half is written with an L2
// evict_last policy so that the half-filled sector waits in L2 for its other half instead of going to DRAM twice (without the
// hint ncu showed +2.1 GB of DRAM reads and +1.9 GB of writes per 192 columns: read-modify-write of partial sectors).
#ifndef MW_L2_HINT
#define MW_L2_HINT 1
#endif
__device__ __forceinline__ uint64_t mw_policy_evict_last() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void mw_store2(uint32_t* ptr, uint32_t a, uint32_t b, bool keep, uint64_t pol) {
#if MW_L2_HINT
    if (keep) { asm volatile("st.global.L2::cache_hint.v2.b32 [%0], {%1, %2}, %3;" ::"l"(ptr), "r"(a), "r"(b), "l"(pol) : "memory"); return; }
#endif
    *reinterpret_cast<uint2*>(ptr) = make_uint2(a, b);
}
__device__ __forceinline__ uint32_t mw_mulc(uint32_t x, const uint32_t (&c)[2]) { return fmul_shoup(x, c[0], c[1]); }

__global__ void __launch_bounds__(32 * MW_WARPS, 1) mid_warp_kernel(MidWarpArgs p) {
    extern __shared__ uint4 mw_smem_[];
    uint32_t* sm = reinterpret_cast<uint32_t*>(mw_smem_);
    // tw4 / G2 are stored COSET-MAJOR (entry of expanded index 4 c + r at [r][c]): phase 4 walks one coset with c = lane + 32 j, so
    // consecutive lanes read consecutive pairs (the natural order would put them 32 bytes apart: 4-way bank conflicts)
    uint32_t *twI = sm, *G3 = twI + 2048, *Gs = G3 + 2048, *tw4 = Gs + 2 * MW_GS_PAIRS, *G2 = tw4 + 8192;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t hi = blockIdx.x, rb = brev(hi, p.b);
    auto put = [](uint32_t* t, uint32_t i, uint32_t wm) { t[2 * i] = from_mont(wm); t[2 * i + 1] = shoup_quot_mont(wm); };
    // per-level layout: index i in [2^(l-1), 2^l) holds w_{2^l}^(+-(i - 2^(l-1)))
    for (uint32_t i = (uint32_t)tid; i < 1024; i += 32 * MW_WARPS) {
        uint32_t v = ONE;
        if (i >= 1) { const int l = 32 - __clz((int)i); v = tab_pow(p.rt.i_lo, p.rt.i_hi, (i - (1u << (l - 1))) << (24 - l)); }
        put(twI, i, v);
        put(G3, i, tab_pow(p.rt.i_lo, p.rt.i_hi, (rb * i) << (24 - p.n)));                                      // w_{2^n}^-(rev(hi) i)
        put(Gs, i + (i >> 5), fmul(p.n_inv, tab_pow(p.rt.p3_lo, p.rt.p3_hi, (brev(i, 10) << p.b) + rb)));       // n^-1 3^(coefficient index)
    }
    for (uint32_t i = (uint32_t)tid; i < 4096; i += 32 * MW_WARPS) {
        const uint32_t r = i >> 10, c = i & 1023u;
        put(G2, i, tab_pow(p.rt.f_lo, p.rt.f_hi, (rb * (4u * c + r)) << (24 - p.n - 2)));                       // w_{2^(n+2)}^(rev(hi) (4 c + r))
        // levels 8..12 of the 2^12 transform, coset r: slot 32 h + x', h = 2^(q-1), x' < 32 h, holds w_{2^(7+q)}^(4 x' + r)
        uint32_t v = ONE;
        if (c >= 32) { const int q = 32 - __clz((int)(c >> 5)); v = tab_pow(p.rt.f_lo, p.rt.f_hi, (4u * (c - (32u << (q - 1))) + r) << (24 - 7 - q)); }
        put(tw4, i, v);
    }
    __syncthreads();
    const uint64_t keep_policy = mw_policy_evict_last();
    uint32_t* buf = sm + MW_TABLE_WORDS + (uint32_t)warp * 2 * MW_PLANE;
    const uint32_t col_end = (blockIdx.y + 1) * p.cols_per_block < p.ncols ? (blockIdx.y + 1) * p.cols_per_block : p.ncols;
    for (uint32_t col = blockIdx.y * p.cols_per_block + (uint32_t)warp; col < col_end; col += MW_WARPS) {
        const uint32_t* src = p.in + (uint64_t)col * p.in_stride + ((uint64_t)hi << 10);
        uint32_t* dst = p.out + (uint64_t)col * p.out_stride + ((uint64_t)hi << 12);
        uint32_t v[32];
        // ---- phase 1: elements lane + 32 j; inverse (Gentleman-Sande) levels 10..6 ----
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = src[lane + 32 * j];
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = fmul_pair(v[j], G3, (uint32_t)(lane + 32 * j));
#pragma unroll
        for (int q = 5; q >= 1; q--) {
            const int h = 1 << (q - 1);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                if (j & h) continue;
                const uint32_t a = v[j], b = v[j + h];
                v[j] = fadd(a, b);
                v[j + h] = fmul_pair(fsub_lazy(a, b), twI, (1u << (5 + q - 1)) + ((uint32_t)(j & (h - 1)) << 5) + (uint32_t)lane);
            }
        }
#pragma unroll
        for (int j = 0; j < 32; j++) buf[lane + 33 * j] = v[j];  // padi(lane + 32 j)
        __syncwarp();
        // ---- phase 2: elements 32 lane + j; inverse levels 5..1 with constant twiddles, then x n^-1 3^(...) ----
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = buf[33 * lane + j];  // padi(32 lane + j)
        __syncwarp();  // the buffer becomes the coset planes below
#pragma unroll
        for (int q = 5; q >= 1; q--) {
            const int h = 1 << (q - 1);
#pragma unroll
            for (int j = 0; j < 32; j++) {
                if (j & h) continue;
                const uint32_t a = v[j], b = v[j + h];
                // the last level feeds the scaling product, which takes any 32-bit operand: no range corrections there
                v[j] = q == 1 ? fadd_lazy(a, b) : fadd(a, b);
                if ((j & (h - 1)) != 0) v[j + h] = mw_mulc(fsub_lazy(a, b), g_mwc.inv[h + (j & (h - 1))]);
                else v[j + h] = q == 1 ? fsub_lazy(a, b) : fsub(a, b);
            }
        }
#pragma unroll
        for (int j = 0; j < 32; j++) v[j] = fmul_pair(v[j], Gs, (uint32_t)(33 * lane + j));
        // ---- the four cosets, two at a time (loops over pair / coset are NOT unrolled: the code stays inside the instruction cache) ----
#pragma unroll 1
        for (int pr = 0; pr < 2; pr++) {
            // phase 3: forward (Cooley-Tukey) levels 3..7 of coset r on coefficients 32 lane + j
#pragma unroll 1
            for (int dr = 0; dr < 2; dr++) {
                const int r = 2 * pr + dr;
                uint32_t w[32];
#pragma unroll
                for (int j = 0; j < 32; j++) w[j] = v[j];
#pragma unroll
                for (int q = 1; q <= 5; q++) {
                    const int h = 1 << (q - 1);
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if (j & h) continue;
                        const uint32_t x = mw_mulc(w[j + h], g_mwc.fwd[r][h + (j & (h - 1))]);
                        const int hn = h << 1;  // an output the next level multiplies stays in [0, 2p)
                        const bool lz0 = q < 5 && (j & hn), lz1 = q < 5 && ((j + h) & hn);
                        w[j + h] = lz1 ? fsub_lazy(w[j], x) : fsub(w[j], x);
                        w[j] = lz0 ? fadd_lazy(w[j], x) : fadd(w[j], x);
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j++) buf[dr * MW_PLANE + 33 * lane + j] = w[j];
            }
            __syncwarp();
            // phase 4: forward levels 8..12 of coset r on elements c = lane + 32 j (expanded index 4 c + r), x w^(rev(hi) (4 c + r))
            uint32_t o[32];
#pragma unroll 1
            for (int dr = 0; dr < 2; dr++) {
                const uint32_t r = 2u * (uint32_t)pr + (uint32_t)dr;
                const uint32_t* t4 = tw4 + 2048u * r;
                const uint32_t* g2 = G2 + 2048u * r;
#pragma unroll
                for (int j = 0; j < 32; j++) o[j] = buf[dr * MW_PLANE + lane + 33 * j];
#pragma unroll
                for (int q = 1; q <= 5; q++) {
                    const int h = 1 << (q - 1);
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        if (j & h) continue;
                        const uint32_t x = fmul_pair(o[j + h], t4, (uint32_t)(32 * h + 32 * (j & (h - 1)) + lane));
                        const int hn = h << 1;
                        const bool lz0 = q == 5 || (j & hn), lz1 = q == 5 || ((j + h) & hn);  // last level: the G2 product follows
                        o[j + h] = lz1 ? fsub_lazy(o[j], x) : fsub(o[j], x);
                        o[j] = lz0 ? fadd_lazy(o[j], x) : fadd(o[j], x);
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; j++) o[j] = fmul_pair(o[j], g2, (uint32_t)(lane + 32 * j));
                if (dr == 0) {
                    // park the even coset where it came from (each lane re-reads only its own words)
#pragma unroll
                    for (int j = 0; j < 32; j++) buf[lane + 33 * j] = o[j];
                }
            }
#pragma unroll
            for (int j = 0; j < 32; j++) mw_store2(dst + 4u * (uint32_t)(lane + 32 * j) + 2u * (uint32_t)pr, buf[lane + 33 * j], o[j], pr == 0, keep_policy);
            __syncwarp();
        }
    }
}

static inline MidWarpConsts mid_warp_make_consts() {
    MidWarpConsts c{};
    auto rou = [](int k) { uint32_t g = to_mont(137); for (int i = k; i < 27; i++) g = fmul(g, g); return g; };  // Montgomery form
    auto set = [](uint32_t (&d)[2], uint32_t wm) { d[0] = from_mont(wm); d[1] = shoup_quot_mont(wm); };
    for (int q = 1; q <= 5; q++) {
        const uint32_t wi = finv(rou(q)), wf = rou(q + 2);
        for (int x = 0; x < (1 << (q - 1)); x++) {
            const int s = (1 << (q - 1)) + x;
            set(c.inv[s], fpow(wi, (uint64_t)x));
            for (int r = 0; r < 4; r++) set(c.fwd[r][s], fpow(wf, (uint64_t)(4 * x + r)));
        }
    }
    set(c.inv[0], ONE);
    for (int r = 0; r < 4; r++) set(c.fwd[r][0], ONE);
    return c;
}

// Returns false when the shape is not this kernel's (the caller falls back to MiddleKernel2).
static inline bool mid_warp(Dev* dev, const RootTables& rt, const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t ncols, int n, int a, int e,
                            uint32_t flags) {
    static const bool enabled = [] { const char* v = std::getenv("HFB200_MID_WARP"); return !v || std::atoi(v) != 0; }();
    const uint32_t want = MID_INTT | MID_SHIFT | MID_FWD;
    if (!enabled || a != 10 || e != 2 || (flags & ~MID_GFLY) != want || n <= a || n + 2 > 24 || ncols == 0) return false;
    if ((reinterpret_cast<uintptr_t>(out) & 7u) || ((out_stride * 4) & 7u)) return false;  // 8-byte stores
    if (dev->replay) { dev->launches++; return true; }
    static std::atomic<bool> configured[Dev::MAX_DEVICES];
    static std::mutex mu;
    const int d = dev->device >= 0 && dev->device < Dev::MAX_DEVICES ? dev->device : 0;
    if (!configured[d].load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(mu);
        if (!configured[d].load(std::memory_order_relaxed)) {
            const MidWarpConsts c = mid_warp_make_consts();
            CUDA_CHECK(cudaMemcpyToSymbol(g_mwc, &c, sizeof c));
            CUDA_CHECK(cudaFuncSetAttribute(mid_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MW_SMEM));
            configured[d].store(true, std::memory_order_release);
        }
    }
    MidWarpArgs p{};
    p.in = in; p.out = out; p.in_stride = in_stride; p.out_stride = out_stride; p.ncols = ncols;
    p.n = n; p.b = n - a;
    p.n_inv = finv(to_mont((uint32_t)((1ull << n) % P)));
    p.rt = rt;
    const uint64_t chunks = 1ull << p.b;
    uint32_t groups = (uint32_t)((2ull * dev->sm_count + chunks - 1) / chunks);
    if (groups < 1) groups = 1;
    const uint32_t max_groups = (ncols + MW_WARPS - 1) / MW_WARPS;
    if (groups > max_groups) groups = max_groups;
    p.cols_per_block = (ncols + groups - 1) / groups;
    groups = (ncols + p.cols_per_block - 1) / p.cols_per_block;
    mid_warp_kernel<<<dim3((unsigned)chunks, groups), 32 * MW_WARPS, MW_SMEM, dev->stream>>>(p);
    CUDA_CHECK(cudaGetLastError());
    dev->launches++;
    return true;
}

}  // namespace hf
#endif  // !HFB200_EMU
