// Poseidon2 (BabyBear, t = 24, x^7, 4+21+4 rounds) permutation, sponge and Merkle kernels.
// Replaces risc0-zkp 3.0.4 `core::hash::poseidon2` + `Hal::hash_rows` / `Hal::hash_fold` (upstream GPU:
// sppark poseidon2_rows/fold in risc0-sys 1.5.0; /root/reference/Cargo.lock:3174-3223, not vendored;
// SURVEY.md Appendix A.3/A.4).  Integer-pipe bound: one thread owns one 24-word sponge state in
// registers; a warp owns 32 consecutive rows so every column read is one coalesced 128-byte line.
#pragma once
#include "dev.cuh"

namespace hf {

#include "poseidon2_consts.inc"

struct P2Consts {
    uint32_t rc_first[96], rc_partial[21], rc_last[96], diag[24];  // Montgomery form
    uint32_t diag_std[24], diag_shoup[24];  // canonical d_i and floor(d_i * 2^32 / p): constant multiplier in Shoup form
};

static inline P2Consts p2_make_consts() {
    P2Consts c;
    for (int i = 0; i < 96; i++) { c.rc_first[i] = to_mont(P2_RC_FULL_FIRST[i]); c.rc_last[i] = to_mont(P2_RC_FULL_LAST[i]); }
    for (int i = 0; i < 21; i++) c.rc_partial[i] = to_mont(P2_RC_PARTIAL[i]);
    for (int i = 0; i < 24; i++) {
        c.diag[i] = to_mont(P2_M_INT_DIAG[i]);
        c.diag_std[i] = P2_M_INT_DIAG[i];
        c.diag_shoup[i] = (uint32_t)(((uint64_t)P2_M_INT_DIAG[i] << 32) / P);
    }
    return c;
}

#if !defined(HFB200_EMU)
__constant__ P2Consts g_p2c;
#define P2C_DEV g_p2c
#endif
static P2Consts g_p2c_host;
static inline const P2Consts& p2_host_consts() {
    static const bool once = (g_p2c_host = p2_make_consts(), true);  // thread-safe one-time initialisation
    (void)once;
    return g_p2c_host;
}

HD const P2Consts& p2c() {
#if defined(__CUDA_ARCH__)
    return P2C_DEV;
#else
    return g_p2c_host;
#endif
}

// Pipe steering experiment (B200): 32-bit integer multiplies run only on the heavy half of the FMA pipe, the binding
// unit of this kernel (ncu: sm__pipe_fmaheavy_cycles_active ~92 %).  ptxas splits plain adds between the ALU pipe
// (IADD3) and the FMA pipe (IMAD.IADD).  P2_FADD_ALU=1 pins them to the ALU pipe by writing the add as VIADDMNMX
// (min(a + b, UINT_MAX)); measured on B200 this is 3.7 % SLOWER (28.9 -> 30.0 ms for the 192-column tree), i.e. both
// pipes are already saturated together, so the default stays 0.
#ifndef P2_FADD_ALU
#define P2_FADD_ALU 0
#endif
HD uint32_t padd(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && P2_FADD_ALU
    const uint32_t x = __viaddmin_u32(a, b, 0xFFFFFFFFu);
    return __viaddmin_u32(x, 0u - P, x);
#else
    return fadd(a, b);
#endif
}
// Montgomery product left in (0, 2p): one IADD3 instead of subtract + conditional correction.  Valid as ONE operand of a
// following fmul/fmul_lazy whose other operand is canonical (2p * p < p * 2^32).
HD uint32_t fmul_lazy(uint32_t a, uint32_t b) {
    uint64_t o = (uint64_t)a * b;
    uint32_t m = (uint32_t)o * P_INV;
#ifdef __CUDA_ARCH__
    uint32_t mp = __umulhi(m, P);
#else
    uint32_t mp = (uint32_t)(((uint64_t)m * P) >> 32);
#endif
    return (uint32_t)(o >> 32) - mp + P;
}
// x^7 with 4 multiplies; x4 stays lazy (it only meets the canonical x3).
HD uint32_t sbox7(uint32_t x) { uint32_t x2 = fmul(x, x), x3 = fmul(x2, x), x4 = fmul_lazy(x2, x2); return fmul(x3, x4); }
// Signed Montgomery product: for |a * b| < 2^31 * p the value hi(a*b) - hi(m*p), m = lo(a*b) * p^-1 (all signed), is the
// exact quotient (a*b - m*p) / 2^32 and lies in (-p, p): no correction step at all.
HD int32_t smul(int32_t a, int32_t b) {
    const int64_t o = (int64_t)a * b;
    const int32_t m = (int32_t)((uint32_t)o * P_INV);
#ifdef __CUDA_ARCH__
    const int32_t mp = __mulhi(m, (int32_t)P);
#else
    const int32_t mp = (int32_t)(((int64_t)m * (int64_t)P) >> 32);
#endif
    return (int32_t)(o >> 32) - mp;
}
// (s + rc)^7 for canonical s, rc: the sum enters the chain as the signed representative s + rc - p in [-p, p) (one IADD3),
// the four products stay signed in (-p, p) and only the result is made canonical: 4 * 4 + 2 instructions instead of 21.
#ifndef P2_SIGNED_SBOX
#define P2_SIGNED_SBOX 1
#endif
HD uint32_t sbox7_rc(uint32_t s, uint32_t rc) {
#if P2_SIGNED_SBOX
    const int32_t x = (int32_t)(s + rc - P);
    const int32_t x2 = smul(x, x), x3 = smul(x2, x), x4 = smul(x2, x2);
    const uint32_t r = (uint32_t)smul(x3, x4);
    return umin32(r, r + P);
#else
    return sbox7(padd(s, rc));
#endif
}
// x * d mod p for a constant d given as (d, d' = floor(d 2^32 / p)) -- Shoup: no 64-bit product, result already in [0, 2p).
// x may be any residue representation (here Montgomery), d is the canonical constant.
HD uint32_t fmul_const(uint32_t x, uint32_t d, uint32_t dp) {
#ifdef __CUDA_ARCH__
    const uint32_t q = __umulhi(x, dp);
#else
    const uint32_t q = (uint32_t)(((uint64_t)x * dp) >> 32);
#endif
    const uint32_t r = x * d - q * P;
    return umin32(r, r - P);
}

#ifndef P2_DEFERRED_SUM
#define P2_DEFERRED_SUM 1  // 0: plain internal layer; 1: carried row sum, 23 S by a Shoup product; 2: 23 S by doublings
#endif
// External layer: M4 on each 4-chunk (Poseidon2 add/double chain) then add the cross-chunk column sums.
HD void p2_m_ext(uint32_t* s) {
#pragma unroll
    for (int c = 0; c < 24; c += 4) {
        uint32_t a = s[c], b = s[c + 1], cc = s[c + 2], d = s[c + 3];
        uint32_t t0 = padd(a, b), t1 = padd(cc, d);
        uint32_t t2 = padd(padd(b, b), t1), t3 = padd(padd(d, d), t0);
        uint32_t t4 = padd(t1, t1); t4 = padd(padd(t4, t4), t3);
        uint32_t t5 = padd(t0, t0); t5 = padd(padd(t5, t5), t2);
        s[c] = padd(t3, t5); s[c + 1] = t5; s[c + 2] = padd(t2, t4); s[c + 3] = t4;
    }
    uint32_t sum[4];
#pragma unroll
    for (int j = 0; j < 4; j++) sum[j] = padd(padd(padd(s[j], s[4 + j]), padd(s[8 + j], s[12 + j])), padd(s[16 + j], s[20 + j]));
#pragma unroll
    for (int i = 0; i < 24; i++) s[i] = padd(s[i], sum[i & 3]);
}

HD void p2_m_int(uint32_t* s, const P2Consts& k) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 24; i++) sum = padd(sum, s[i]);
#pragma unroll
    for (int i = 0; i < 24; i++) s[i] = padd(sum, fmul_const(s[i], k.diag_std[i], k.diag_shoup[i]));
}

HD void p2_mix(uint32_t* s) {
    const P2Consts& k = p2c();
    p2_m_ext(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 24; i++) s[i] = sbox7_rc(s[i], k.rc_first[r * 24 + i]);
        p2_m_ext(s);
    }
#if P2_DEFERRED_SUM
    // Partial rounds with the row sum carried instead of recomputed.  The round maps y_i -> mu_i y_i + S with S = sum_i y_i
    // (after the S-box on y_0).  With U = sum_{i>=1} y_i kept canonical:  S = y_0 + U,  t_i = mu_i y_i (Shoup: accepts ANY
    // 32-bit y_i),  U' = sum_{i>=1} t_i + 23 S,  y_0' = t_0 + S (canonical: it meets the S-box),  y_i' = t_i + S left in
    // [0, 2p) for i >= 1 (one add, no range correction: its only reader is the next Shoup product).
    {
        uint32_t U = s[1];
#pragma unroll
        for (int i = 2; i < 24; i++) U = padd(U, s[i]);
#pragma unroll 1
        for (int r = 0; r < 21; r++) {
            s[0] = sbox7_rc(s[0], k.rc_partial[r]);
            const uint32_t S = padd(s[0], U);
            uint32_t t[24];
#pragma unroll
            for (int i = 0; i < 24; i++) t[i] = fmul_const(s[i], k.diag_std[i], k.diag_shoup[i]);
            uint32_t T = t[1];
#pragma unroll
            for (int i = 2; i < 24; i++) T = padd(T, t[i]);
#if P2_DEFERRED_SUM == 2
            uint32_t S2 = padd(S, S), S4 = padd(S2, S2), S8 = padd(S4, S4), S16 = padd(S8, S8);
            const uint32_t S23 = fsub(padd(S16, S8), S);
#else
            const uint32_t S23 = fmul_const(S, 23u, (uint32_t)((23ull << 32) / P));
#endif
            U = padd(T, S23);
            s[0] = padd(t[0], S);
#pragma unroll
            for (int i = 1; i < 24; i++) s[i] = fadd_lazy(t[i], S);
        }
#pragma unroll
        for (int i = 1; i < 24; i++) s[i] = umin32(s[i], s[i] - P);
    }
#else
#pragma unroll 1
    for (int r = 0; r < 21; r++) {
        s[0] = sbox7_rc(s[0], k.rc_partial[r]);
        p2_m_int(s, k);
    }
#endif
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 24; i++) s[i] = sbox7_rc(s[i], k.rc_last[r * 24 + i]);
        p2_m_ext(s);
    }
}

// ---- kernels ------------------------------------------------------------------------------------
// Hal::hash_rows: leaf r = unpadded sponge over matrix[c * col_stride + r], c < cols.  Digest -> nodes[rows + r].
struct HashRowsKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, const uint32_t* matrix, uint64_t col_stride, uint32_t rows, uint32_t cols, uint32_t* nodes) {
        const uint64_t r = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (r >= rows) return;
        uint32_t s[24];
#pragma unroll
        for (int i = 0; i < 24; i++) s[i] = 0;
        const uint32_t* src = matrix + r;
        uint32_t c = 0;
        for (; c + 16 <= cols; c += 16) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = src[(uint64_t)(c + i) * col_stride];
            p2_mix(s);
        }
        if (c < cols || cols == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) s[i] = (c + i < cols) ? src[(uint64_t)(c + i) * col_stride] : 0u;
            p2_mix(s);
        }
        uint4* dst = reinterpret_cast<uint4*>(nodes + ((uint64_t)rows + r) * 8);
        dst[0] = make_uint4(s[0], s[1], s[2], s[3]);
        dst[1] = make_uint4(s[4], s[5], s[6], s[7]);
    }
};

// Hal::hash_fold for one level: nodes[i] = H(nodes[2i] || nodes[2i+1]) for i in [level_size, 2*level_size).
struct HashFoldKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, uint32_t* nodes, uint32_t level_size) {
        const uint64_t t = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (t >= level_size) return;
        const uint64_t i = level_size + t;
        const uint4* in = reinterpret_cast<const uint4*>(nodes + 2 * i * 8);
        uint32_t s[24];
        uint4 v0 = in[0], v1 = in[1], v2 = in[2], v3 = in[3];
        s[0] = v0.x; s[1] = v0.y; s[2] = v0.z; s[3] = v0.w; s[4] = v1.x; s[5] = v1.y; s[6] = v1.z; s[7] = v1.w;
        s[8] = v2.x; s[9] = v2.y; s[10] = v2.z; s[11] = v2.w; s[12] = v3.x; s[13] = v3.y; s[14] = v3.z; s[15] = v3.w;
#pragma unroll
        for (int k = 16; k < 24; k++) s[k] = 0;
        p2_mix(s);
        uint4* dst = reinterpret_cast<uint4*>(nodes + i * 8);
        dst[0] = make_uint4(s[0], s[1], s[2], s[3]);
        dst[1] = make_uint4(s[4], s[5], s[6], s[7]);
    }
};

// The last levels of the tree in one block: levels of size top, top/2, ..., 1 (top <= block threads).
struct HashFoldTailKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t*, uint32_t* nodes, uint32_t top) {
        for (uint32_t level = top; level >= 1; level >>= 1) {
            for (uint32_t t = cx.tid; t < level; t += cx.nt) {
                const uint64_t i = level + t;
                uint32_t s[24];
#pragma unroll
                for (int k = 0; k < 16; k++) s[k] = nodes[2 * i * 8 + k];
#pragma unroll
                for (int k = 16; k < 24; k++) s[k] = 0;
                p2_mix(s);
#pragma unroll
                for (int k = 0; k < 8; k++) nodes[i * 8 + k] = s[k];
            }
#ifdef __CUDA_ARCH__
            __threadfence_block();
#endif
            cx.sync();
        }
    }
};

// ---- warp-cooperative permutation for the small upper levels of a tree --------------------------------------------
// With one thread per node a level of <= a few thousand nodes is pure latency (one permutation ~17 us, one launch per
// level).  Here a WARP computes one node: lane L < 24 owns state cell L, the S-boxes run 24-wide, the linear layers use
// shuffles (4-lane M4 groups, xor-butterfly column sums, warp all-reduce for the internal layer): ~3.5 us per node.
// 4x more instructions per node than the thread version, so it is used only where latency, not throughput, binds.
#ifdef __CUDA_ARCH__
struct WarpConsts { uint32_t rcf[4], rcl[4], d, dp; };
__device__ __forceinline__ WarpConsts wp2_consts(int lane) {
    const P2Consts& k = p2c();
    WarpConsts w;
    const int l = lane < 24 ? lane : 0;
#pragma unroll
    for (int r = 0; r < 4; r++) { w.rcf[r] = lane < 24 ? k.rc_first[r * 24 + l] : 0u; w.rcl[r] = lane < 24 ? k.rc_last[r * 24 + l] : 0u; }
    w.d = k.diag_std[l]; w.dp = k.diag_shoup[l];
    return w;
}
__device__ __forceinline__ uint32_t wp2_m_ext(uint32_t x, int lane) {
    const int q = lane & ~3;
    const uint32_t a = __shfl_sync(0xffffffffu, x, q), b = __shfl_sync(0xffffffffu, x, q + 1);
    const uint32_t c = __shfl_sync(0xffffffffu, x, q + 2), d = __shfl_sync(0xffffffffu, x, q + 3);
    const uint32_t t0 = fadd(a, b), t1 = fadd(c, d);
    const uint32_t t2 = fadd(fadd(b, b), t1), t3 = fadd(fadd(d, d), t0);
    uint32_t t4 = fadd(t1, t1); t4 = fadd(fadd(t4, t4), t3);
    uint32_t t5 = fadd(t0, t0); t5 = fadd(fadd(t5, t5), t2);
    const int r = lane & 3;
    const uint32_t out = r == 0 ? fadd(t3, t5) : r == 1 ? t5 : r == 2 ? fadd(t2, t4) : t4;
    uint32_t s = out;
    s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 4));
    s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 8));
    s = fadd(s, __shfl_xor_sync(0xffffffffu, s, 16));
    return lane < 24 ? fadd(out, s) : 0u;
}
__device__ __forceinline__ uint32_t wp2_mix(uint32_t x, int lane, const WarpConsts& w) {
    const P2Consts& k = p2c();
    x = wp2_m_ext(x, lane);
#pragma unroll
    for (int r = 0; r < 4; r++) x = wp2_m_ext(sbox7_rc(x, w.rcf[r]), lane);
#pragma unroll 1
    for (int r = 0; r < 21; r++) {
        const uint32_t y = sbox7_rc(x, k.rc_partial[r]);
        if (lane == 0) x = y;
        uint32_t s = x;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) s = fadd(s, __shfl_xor_sync(0xffffffffu, s, off));
        x = lane < 24 ? fadd(s, fmul_const(x, w.d, w.dp)) : 0u;
    }
#pragma unroll
    for (int r = 0; r < 4; r++) x = wp2_m_ext(sbox7_rc(x, w.rcl[r]), lane);
    return x;
}
__device__ __forceinline__ void wp2_hash_pair(uint32_t* nodes, uint64_t i, int lane, const WarpConsts& w) {
    uint32_t x = lane < 16 ? nodes[2 * i * 8 + lane] : 0u;
    x = wp2_mix(x, lane, w);
    if (lane < 8) nodes[i * 8 + lane] = x;
}
#endif
// one level, one warp per node (device only; the emulator build keeps the thread-per-node path)
struct HashFoldWarpKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t*, uint32_t* nodes, uint32_t level_size) {
#ifdef __CUDA_ARCH__
        const int lane = cx.tid & 31;
        const uint64_t wid = ((uint64_t)cx.bx * cx.nt + cx.tid) >> 5;
        if (wid >= level_size) return;
        const WarpConsts w = wp2_consts(lane);
        wp2_hash_pair(nodes, (uint64_t)level_size + wid, lane, w);
#else
        (void)cx; (void)nodes; (void)level_size;
#endif
    }
};
// levels top, top/2, ..., 1 in one block of 32 warps
struct HashFoldWarpTailKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t*, uint32_t* nodes, uint32_t top) {
#ifdef __CUDA_ARCH__
        const int lane = cx.tid & 31, warp = cx.tid >> 5, nwarps = cx.nt >> 5;
        const WarpConsts w = wp2_consts(lane);
        for (uint32_t level = top; level >= 1; level >>= 1) {
            for (uint32_t t = warp; t < level; t += nwarps) wp2_hash_pair(nodes, (uint64_t)level + t, lane, w);
            __threadfence_block();
            __syncthreads();
        }
#else
        (void)cx; (void)nodes; (void)top;
#endif
    }
};

// poseidon2_mix on n independent states (parity probe for the permutation itself).
struct PermuteKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, uint32_t* states, uint32_t n) {
        const uint64_t t = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (t >= n) return;
        uint32_t s[24];
        for (int i = 0; i < 24; i++) s[i] = states[t * 24 + i];
        p2_mix(s);
        for (int i = 0; i < 24; i++) states[t * 24 + i] = s[i];
    }
};

struct Merkle {
    Dev* dev = nullptr;
    void init(Dev* d) {
        dev = d;
        p2_host_consts();
#if !defined(HFB200_EMU)
        CUDA_CHECK(cudaMemcpyToSymbol(g_p2c, &g_p2c_host, sizeof(P2Consts)));
#endif
    }
    // nodes: 2*rows digests (heap layout).  matrix element (r, c) at matrix[c*col_stride + r].
    void build(const uint32_t* matrix, uint64_t col_stride, uint32_t rows, uint32_t cols, uint32_t* nodes) {
        const int T = 128;
        dev->launch<HashRowsKernel, 128, 1>((rows + T - 1) / T, 1, T, 0, matrix, col_stride, rows, cols, nodes);
        uint32_t level = rows / 2;
#ifndef HFB200_EMU
        // big levels: one thread per node (throughput); <= 4096 nodes: one warp per node (latency); <= 32: one block
        static const uint32_t warp_max = [] { const char* e = std::getenv("HFB200_FOLD_WARP_MAX"); const int v = e ? std::atoi(e) : 0; return v >= 64 ? (uint32_t)v : 4096u; }();
        for (; level > warp_max; level >>= 1) dev->launch<HashFoldKernel, 128, 1>((level + T - 1) / T, 1, T, 0, nodes, level);
        for (; level > 32; level >>= 1) dev->launch<HashFoldWarpKernel, 256, 1>((level * 32 + 255) / 256, 1, 256, 0, nodes, level);
        if (level >= 1) dev->launch<HashFoldWarpTailKernel, 1024, 1>(1, 1, 1024, 0, nodes, level);
#else
        for (; level >= 512; level >>= 1) dev->launch<HashFoldKernel, 128, 1>((level + T - 1) / T, 1, T, 0, nodes, level);
        if (level >= 1) dev->launch<HashFoldTailKernel, 256, 1>(1, 1, 256, 0, nodes, level);
#endif
    }
};

// ---- host-side sponge / RNG for the Fiat-Shamir transcript (WriteIOP) ---------------------------------
struct Digest8 { uint32_t w[8]; };

static inline Digest8 host_hash_elems(const uint32_t* in, size_t n) {
    p2_host_consts();
    uint32_t s[24] = {0};
    size_t used = 0; bool any = false;
    for (size_t i = 0; i < n; i++) {
        s[used++] = in[i];
        if (used == 16) { p2_mix(s); used = 0; any = true; }
    }
    if (used != 0 || !any) { for (size_t k = used; k < 16; k++) s[k] = 0; p2_mix(s); }
    Digest8 d; for (int i = 0; i < 8; i++) d.w[i] = s[i];
    return d;
}

struct HostRng {
    uint32_t cells[24] = {0};
    int pool_used = 0;
    // Poseidon2Rng::mix: a permutation first when elements were drawn since the last mix (upstream: "if switching from squeezing,
    // do a poseidon2 mix"), then add the 8 digest words into the rate and permute
    void mix(const uint32_t* d8) {
        p2_host_consts();
        if (pool_used != 0) { p2_mix(cells); pool_used = 0; }
        for (int i = 0; i < 8; i++) cells[i] = fadd(cells[i], d8[i]);
        p2_mix(cells);
    }
    uint32_t random_elem() { if (pool_used == 16) { p2_mix(cells); pool_used = 0; } return cells[pool_used++]; }
    E4 random_ext() { uint32_t a = random_elem(), b = random_elem(), c = random_elem(), d = random_elem(); return e4(a, b, c, d); }
    uint32_t random_bits(unsigned bits) {
        uint32_t v = from_mont(random_elem());
        for (int i = 0; i < 3; i++) { uint32_t n = from_mont(random_elem()); if (v == 0) v = n; }
        return v & (uint32_t)((1ull << bits) - 1);
    }
};

// ---- transcript header ------------------------------------------------------------------------------------------------
// risc0-circuit-rv32im `SegmentProver::prove` / risc0-zkp `verify`: "At the start of the protocol, seed the Fiat-Shamir
// transcript with context information about the proof system and circuit":
//   commit(H(PROOF_SYSTEM_INFO.encode())); commit(H(CIRCUIT_INFO.encode()));         (`ProtocolInfo`: 16 bytes, one element per byte)
//   header = globals ++ [po2 as a raw word]; commit(H(header)); write header
// Returns the header digest.  Host code, shared by the prover (both transcript modes) and the verifier.
static constexpr char PROOF_SYSTEM_INFO[17] = "RISC0_STARK:v1__";      // risc0-zkp PROOF_SYSTEM_INFO
static constexpr char BUILTIN_CIRCUIT_INFO[17] = "SYNTH_RV32IM:v1_";   // the declared stand-in circuit names itself
static constexpr char DEFAULT_IR_CIRCUIT_INFO[17] = "RV32IM:v2_______"; // data-defined circuits: upstream's rv32im-v2 string unless given
// dst[17] <- the circuit's info string: built-in circuit, data-defined with its own 16 bytes, or data-defined default
static inline void set_circuit_info(char* dst, bool data_defined, const uint8_t* info16) {
    bool given = false;
    if (info16) for (int i = 0; i < 16; i++) given = given || info16[i] != 0;
    if (!data_defined) std::memcpy(dst, BUILTIN_CIRCUIT_INFO, 17);
    else if (given) { std::memcpy(dst, info16, 16); dst[16] = 0; }
    else std::memcpy(dst, DEFAULT_IR_CIRCUIT_INFO, 17);
}
static inline Digest8 protocol_info_digest(const char* info16) {
    uint32_t e[16];
    for (int i = 0; i < 16; i++) e[i] = to_mont((uint32_t)(uint8_t)info16[i]);
    return host_hash_elems(e, 16);
}
static inline Digest8 transcript_header(HostRng& rng, const char* circuit_info16, const uint32_t* globals, size_t n_globals, uint32_t po2) {
    rng.mix(protocol_info_digest(PROOF_SYSTEM_INFO).w);
    rng.mix(protocol_info_digest(circuit_info16).w);
    std::vector<uint32_t> header(globals, globals + n_globals);
    header.push_back(po2);
    const Digest8 d = host_hash_elems(header.data(), header.size());
    rng.mix(d.w);
    return d;
}

}  // namespace hf
