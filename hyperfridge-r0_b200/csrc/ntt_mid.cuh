// Fused middle stage of the main-group LDE (chunk iNTT . zk_shift . x4 expand . chunk NTT, 2^10 -> 2^12 per chunk) with radix-32
// register rounds: ONE WARP per transform, FOUR-WARP TEAMS per four columns (sm_100a only).
//
// MiddleKernel2<10, 2> (ntt.cuh) gives a (column, chunk) to a 64-thread unit with 16 values per thread: radix-8 / radix-16 rounds,
// five shared-memory round trips, 294 executed instructions per trace element.  Its one-warp experiment (MID_R5) cut that to 249 but
// kept all four cosets of the expanded chunk (16.5 KB) per warp: 8 warps per SM, issue rate down, no gain (profiles/r2_mid_r5_*).
// Here every warp works on 32 values per lane and the SM still holds 16 warps:
//   * the four cosets of the x4 expansion (index = 4 c + r) are INDEPENDENT 2^10-point transforms once the two replication levels
//     are skipped.  A team of four warps takes four columns of the chunk: warp w first runs the inverse part of column w
//       phase 1  inverse levels 10..6 on elements lane + 32 j (global -> registers, x w^-(rev(hi) lo)), twiddles from the per-level table
//       phase 2  inverse levels 5..1 on elements 32 lane + j: every twiddle is a compile-time constant -> __constant__ operands, then
//                x n^-1 3^j (zk_shift); the 1024 coefficients are left in the team's plane V[w];
//     then, column by column, warp w produces COSET w of that column
//       phase 3  forward levels 3..7 on coefficients 32 lane + j (constant twiddles, indexed by the coset)
//       phase 4  forward levels 8..12 on elements lane + 32 j, per-level table, x w^(rev(hi) (4 c + w)), result parked in plane R[w];
//     and after a team barrier the four planes leave as 16-byte stores (c-major, all four cosets of an element together): every
//     32-byte sector is written whole.  (Two cosets per warp with 8-byte stores was measured first: half-filled sectors cost
//     +2.0 GB of DRAM reads and +1.8 GB of writes per 192 columns, L2 evict_last hints did not change that.)
//   * 8 planes of 1088 words per team = 34.8 KB (two pad words per 32: 64-bit shared accesses, conflict-free both ways), four teams
//     per CTA: 139 KB + 88 KB of tables, one CTA of 16 warps per SM;
//   * tables for phase 4 are stored coset-major so that consecutive lanes read consecutive (w, w') pairs;
//   * butterfly outputs that only feed a product (the last level of phases 2 and 4) skip their range correction.
// Same arithmetic and output as MiddleKernel2 (exact field arithmetic: the LDE is bit-identical); the host emulator keeps
// MiddleKernel2.  Replaces the same risc0-zkp `Hal` ops as ntt.cuh.
#pragma once
#ifndef HFB200_EMU
#include "ntt.cuh"

namespace hf {

static constexpr int MW_WARPS = 16;
static constexpr uint32_t MW_PLANE = 1024 + 64;                                    // one coset, TWO pad words per 32: element i at i + 2 (i >> 5), so a lane's
                                                                                    // 32 consecutive elements start 8-byte aligned (64-bit shared-memory accesses,
                                                                                    // conflict-free both ways: 34 L + j across lanes, L + 34 j across lanes)
static constexpr uint32_t MW_GS_PAIRS = 1024 + 32;                                  // one pad pair per 32: lane L reads pairs 33 L + j
static constexpr uint32_t MW_TABLE_WORDS = 2 * (1024 + 1024 + MW_GS_PAIRS + 4096 + 4096);  // twI, G3, Gs, tw4, G2 as (w, w') pairs
static constexpr int MW_TEAMS = MW_WARPS / 4;
static constexpr size_t MW_SMEM = (size_t)(MW_TABLE_WORDS + MW_TEAMS * 8 * MW_PLANE) * 4;

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

// the 32 consecutive elements 32 L .. 32 L + 31 of a plane, as 16 eight-byte accesses
__device__ __forceinline__ void mw_load_row(uint32_t (&v)[32], const uint32_t* plane, int lane) {
    const uint2* q = reinterpret_cast<const uint2*>(plane + 34 * lane);
#pragma unroll
    for (int j = 0; j < 16; j++) { const uint2 t = q[j]; v[2 * j] = t.x; v[2 * j + 1] = t.y; }
}
__device__ __forceinline__ void mw_store_row(uint32_t* plane, int lane, const uint32_t (&v)[32]) {
    uint2* q = reinterpret_cast<uint2*>(plane + 34 * lane);
#pragma unroll
    for (int j = 0; j < 16; j++) q[j] = make_uint2(v[2 * j], v[2 * j + 1]);
}
__device__ __forceinline__ void mw_team_sync(int team) { asm volatile("bar.sync %0, 128;" ::"r"(team + 1) : "memory"); }
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
    const int team = warp >> 2, w = warp & 3;  // warp w of a team: inverse part of column w of a quad, then coset w of every column
    uint32_t* V = sm + MW_TABLE_WORDS + (uint32_t)team * 8 * MW_PLANE;  // V[4]: coefficients of the quad's columns
    uint32_t* R = V + 4 * MW_PLANE;                                       // R[4]: the four cosets of the column being expanded
    uint32_t* Vw = V + (uint32_t)w * MW_PLANE;
    uint32_t* Rw = R + (uint32_t)w * MW_PLANE;
    const uint32_t* t4 = tw4 + 2048u * (uint32_t)w;
    const uint32_t* g2 = G2 + 2048u * (uint32_t)w;
    const uint32_t col_begin = blockIdx.y * p.cols_per_block;
    const uint32_t col_end = col_begin + p.cols_per_block < p.ncols ? col_begin + p.cols_per_block : p.ncols;
    for (uint32_t quad = col_begin + 4u * (uint32_t)team; quad < col_end; quad += 4u * MW_TEAMS) {
        uint32_t v[32];
        if (quad + (uint32_t)w < col_end) {
            const uint32_t* src = p.in + (uint64_t)(quad + (uint32_t)w) * p.in_stride + ((uint64_t)hi << 10);
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
            for (int j = 0; j < 32; j++) Vw[lane + 34 * j] = v[j];  // element lane + 32 j
            __syncwarp();
            // ---- phase 2: elements 32 lane + j; inverse levels 5..1 with constant twiddles, then x n^-1 3^(...) ----
            mw_load_row(v, Vw, lane);  // elements 32 lane + j: each lane rewrites exactly the words it read
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
            mw_store_row(Vw, lane, v);
        }
        mw_team_sync(team);  // the coefficients of the quad are in V[0..3]
        // ---- column by column: warp w expands coset w (loop not unrolled: the code stays inside the instruction cache) ----
#pragma unroll 1
        for (uint32_t J = 0; J < 4 && quad + J < col_end; J++) {
            const uint32_t* VJ = V + J * MW_PLANE;
            mw_load_row(v, VJ, lane);
            // phase 3: forward (Cooley-Tukey) levels 3..7 of coset w on coefficients 32 lane + j
#pragma unroll
            for (int q = 1; q <= 5; q++) {
                const int h = 1 << (q - 1);
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    if (j & h) continue;
                    const uint32_t x = mw_mulc(v[j + h], g_mwc.fwd[w][h + (j & (h - 1))]);
                    const int hn = h << 1;  // an output the next level multiplies stays in [0, 2p)
                    const bool lz0 = q < 5 && (j & hn), lz1 = q < 5 && ((j + h) & hn);
                    v[j + h] = lz1 ? fsub_lazy(v[j], x) : fsub(v[j], x);
                    v[j] = lz0 ? fadd_lazy(v[j], x) : fadd(v[j], x);
                }
            }
            if (J > 0) mw_team_sync(team);  // the previous column's planes have been stored
            mw_store_row(Rw, lane, v);
            __syncwarp();
            // phase 4: forward levels 8..12 of coset w on elements c = lane + 32 j (expanded index 4 c + w), x w^(rev(hi) (4 c + w))
#pragma unroll
            for (int j = 0; j < 32; j++) v[j] = Rw[lane + 34 * j];
#pragma unroll
            for (int q = 1; q <= 5; q++) {
                const int h = 1 << (q - 1);
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    if (j & h) continue;
                    const uint32_t x = fmul_pair(v[j + h], t4, (uint32_t)(32 * h + 32 * (j & (h - 1)) + lane));
                    const int hn = h << 1;
                    const bool lz0 = q == 5 || (j & hn), lz1 = q == 5 || ((j + h) & hn);  // last level: the G2 product follows
                    v[j + h] = lz1 ? fsub_lazy(v[j], x) : fsub(v[j], x);
                    v[j] = lz0 ? fadd_lazy(v[j], x) : fadd(v[j], x);
                }
            }
#pragma unroll
            for (int j = 0; j < 32; j++) Rw[lane + 34 * j] = fmul_pair(v[j], g2, (uint32_t)(lane + 32 * j));  // each lane rewrites the words it read
            mw_team_sync(team);  // all four cosets of column quad + J are parked
            // 16-byte stores, c-major: warp w writes c in [256 w, 256 w + 256)
            uint32_t* dst = p.out + (uint64_t)(quad + J) * p.out_stride + ((uint64_t)hi << 12);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                const uint32_t c = 256u * (uint32_t)w + 32u * (uint32_t)k + (uint32_t)lane;
                const uint32_t pc = c + 2u * (c >> 5);
                const uint4 four = make_uint4(R[pc], R[MW_PLANE + pc], R[2 * MW_PLANE + pc], R[3 * MW_PLANE + pc]);
                *reinterpret_cast<uint4*>(dst + 4u * c) = four;
            }
        }
        mw_team_sync(team);  // V and R are free for the next quad
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
    if ((reinterpret_cast<uintptr_t>(out) & 15u) || ((out_stride * 4) & 15u)) return false;  // 16-byte stores
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
    const uint32_t max_groups = (ncols + 4 * MW_TEAMS - 1) / (4 * MW_TEAMS);
    if (groups > max_groups) groups = max_groups;
    p.cols_per_block = ((ncols + groups - 1) / groups + 3u) & ~3u;  // whole quads per CTA
    groups = (ncols + p.cols_per_block - 1) / p.cols_per_block;
    mid_warp_kernel<<<dim3((unsigned)chunks, groups), 32 * MW_WARPS, MW_SMEM, dev->stream>>>(p);
    CUDA_CHECK(cudaGetLastError());
    dev->launches++;
    return true;
}

}  // namespace hf
#endif  // !HFB200_EMU
