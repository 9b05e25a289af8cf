// DEEP-ALI and FRI kernels.
// Replaces the risc0-zkp 3.0.4 `Hal` ops used by `Prover::finalize` and `fri_prove`
// (`batch_evaluate_any`, `mix_poly_coeffs`, combos prepare/divide, `eltwise_sum_extelem`,
// `batch_bit_reverse`, `fri_fold`, Merkle openings; /root/reference/Cargo.lock:3195-3223, not vendored;
// SURVEY.md Appendix A.7).  B200-first restructuring: upstream works on COEFFICIENTS (Horner-style
// evaluation, coefficient mixing, synthetic division by (x - z w^-b)); here everything is done POINT-WISE on
// the trace domain D1 = { w_N^i / 3 } where the committed polynomials g(y) = f(3y) take the raw trace
// values, so the main groups never need their coefficient form:
//   * g(z w^-b)  = sum_i trace[i-b] * L_i,   L_i = ((3z)^N - 1)/N * w^i / (3z - w^i)     (barycentric)
//   * FRI input  = iNTT+zk_shift of  sum_c (mix-combination_c(i) - U_c(y_i)) / prod_b (y_i - z w^-b)
// Field arithmetic is exact, so the seal is bit-identical to the coefficient-space formulation.
#pragma once
#include "ntt.cuh"
#include "circuit.cuh"

namespace hf {

// INV[i] = 1/(3z - w_N^i), L[i] = A * w_N^i * INV[i], INV4[i] = 1/(w_N^i/3 - z^4)
struct DeepWeightsKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, E4* INV, E4* L, E4* INV4, E4 z3, E4 z4, E4 A, uint32_t po2, RootTables rt) {
        const uint64_t n = 1ull << po2;
        const uint64_t i = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (i >= n) return;
        const uint32_t w = tab_pow(rt.f_lo, rt.f_hi, (uint32_t)(i << (24 - po2)));
        E4 d = z3; d.c[0] = fsub(d.c[0], w);
        const E4 inv = e4_inv(d);
        INV[i] = inv;
        L[i] = e4_scale(e4_mul(A, inv), w);
        E4 d4 = e4_neg(z4); d4.c[0] = fadd(d4.c[0], fmul(w, INV3));
        INV4[i] = e4_inv(d4);
    }
};

// W[i] = x^(bitrev_n(i)) from the squarings xs[k] = x^(2^k): weights for evaluating bit-reversed coefficients.
struct PowBitrevKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, E4* W, const E4* xs, uint32_t po2) {
        const uint64_t n = 1ull << po2;
        const uint64_t i = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (i >= n) return;
        const uint32_t j = brev((uint32_t)i, po2);
        E4 r = e4_one();
        for (uint32_t k = 0; k < po2; k++) if ((j >> k) & 1u) r = e4_mul(r, xs[k]);
        W[i] = r;
    }
};

static constexpr uint32_t DOT_RPB = 4096;  // rows per block (= per partial sum) of the generic kernel and the upper bound of DotKernel's
// DotKernel: rows per block chosen by the host so that small segments still spread over the SMs (2^16 cycles: 512 rows per block,
// 128 row blocks x 3 column groups instead of 16 x 3)
static inline uint32_t dot_rows_per_block(uint32_t po2) {
    const uint64_t r = (1ull << po2) / 128;
    return r < 512 ? 512u : (r > DOT_RPB ? DOT_RPB : (uint32_t)r);
}
// TMA-staged tile: DT_CG columns x DT_R rows of the trace + the DT_R (+1) weights of those rows per pipeline stage.
static constexpr uint32_t DT_CG = 64, DT_R = 128, DT_STAGES = 3, DT_Q = 4, DT_T = DT_CG * DT_Q;
static constexpr uint32_t DT_CSTRIDE = DT_R + 4;                                // words; 16-byte aligned column starts
static constexpr uint32_t DT_STAGE_WORDS = DT_CG * DT_CSTRIDE + (DT_R + 4) * 4;  // columns, then the weights (E4)
static constexpr size_t DT_SMEM = (size_t)DT_STAGES * DT_STAGE_WORDS * 4 + 64;

// partial[(col * nblk + blk) * 2 + b] = sum over the block's `rpb` rows of cols[col][r] * Wt[(r + b) mod n]
// (b = 1 only for col < n_back1).  grid.x = row blocks, grid.y = column groups of DT_CG.
//
// B200 structure: the column segments (512 B each) and the weight slice of a stage arrive by TMA bulk copies
// (cp.async.bulk, one mbarrier per stage, 3 stages in flight) issued by warp 0, so no thread ever waits on a global
// load.  Thread (c, q) owns column c and a quarter q of the stage's rows: its two lazy 64-bit Fp4 accumulators
// (field.cuh) are the only live state (no cross-thread reduction until the very end), the column word is a conflict-free
// LDS (rows visited in an order rotated by q: bank = 4 c + i + q) and the weight is a 16-byte LDS shared by the 8 lanes
// of the same q.  Bound: one IMAD.WIDE per Fp4 component and term.
struct DotKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t* sm, const uint32_t* cols, uint64_t col_stride, uint32_t ncols, uint32_t n_back1, const E4* Wt, uint32_t po2, E4* partial, uint32_t rpb) {
        const uint64_t n = 1ull << po2;
        const uint32_t nblk = cx.gx, c0 = cx.by * DT_CG;
        const uint32_t ncg = ncols - c0 < DT_CG ? ncols - c0 : DT_CG;
        const uint64_t row0 = (uint64_t)cx.bx * rpb;
        const uint32_t n_iter = rpb / DT_R;
#ifdef __CUDA_ARCH__
        constexpr uint32_t NV = 1;
        uint64_t* bars = reinterpret_cast<uint64_t*>(sm + DT_STAGES * DT_STAGE_WORDS);
        if (cx.tid == 0) {
            for (uint32_t s = 0; s < DT_STAGES; s++) mbar_init(&bars[s], 1);
            mbar_init_fence();
        }
        cx.sync();
        auto issue = [&](uint32_t it) {  // warp 0
            uint32_t* st = sm + (it % DT_STAGES) * DT_STAGE_WORDS;
            uint64_t* bar = &bars[it % DT_STAGES];
            const uint64_t r0 = row0 + (uint64_t)it * DT_R;
            const bool wrap = r0 + DT_R == n;  // the shifted weight of the last row is Wt[0]
            if (cx.tid == 0) {
                mbar_expect_tx(bar, ncg * DT_R * 4 + (DT_R + 1) * 16);
                tma_bulk_g2s(st + DT_CG * DT_CSTRIDE, Wt + r0, (wrap ? DT_R : DT_R + 1) * 16, bar);
                if (wrap) tma_bulk_g2s(st + DT_CG * DT_CSTRIDE + DT_R * 4, Wt, 16, bar);
            }
            __syncwarp();
            for (uint32_t c = (uint32_t)cx.tid; c < ncg; c += 32) tma_bulk_g2s(st + c * DT_CSTRIDE, cols + (uint64_t)(c0 + c) * col_stride + r0, DT_R * 4, bar);
        };
        if (cx.tid < 32) for (uint32_t it = 0; it + 1 < DT_STAGES && it < n_iter; it++) issue(it);
#else
        constexpr uint32_t NV = DT_T;
        (void)n;
#endif
        E4A acc[NV][2];
        for (uint32_t v = 0; v < NV; v++) { acc[v][0] = e4a_zero(); acc[v][1] = e4a_zero(); }
        for (uint32_t it = 0; it < n_iter; it++) {
#ifdef __CUDA_ARCH__
            if (cx.tid < 32 && it + DT_STAGES - 1 < n_iter) issue(it + DT_STAGES - 1);  // its stage was drained in iteration it - 1
            mbar_wait(&bars[it % DT_STAGES], (it / DT_STAGES) & 1u);
            const uint32_t* st = sm + (it % DT_STAGES) * DT_STAGE_WORDS;
#else
            uint32_t* st = sm;
            {
                const uint64_t r0 = row0 + (uint64_t)it * DT_R;
                for (uint32_t c = 0; c < ncg; c++) for (uint32_t r = 0; r < DT_R; r++) st[c * DT_CSTRIDE + r] = cols[(uint64_t)(c0 + c) * col_stride + r0 + r];
                E4* wdst = reinterpret_cast<E4*>(st + DT_CG * DT_CSTRIDE);
                for (uint32_t r = 0; r <= DT_R; r++) wdst[r] = Wt[(r0 + r) & ((1ull << po2) - 1)];
            }
#endif
            const E4* ws = reinterpret_cast<const E4*>(st + DT_CG * DT_CSTRIDE);
            for (uint32_t t = cx.tid; t < DT_T; t += cx.nt) {
                const uint32_t c = t / DT_Q, q = t % DT_Q;
                if (c >= ncg) continue;
                E4A* a = acc[NV == 1 ? 0 : t];
                const uint32_t* cs = st + c * DT_CSTRIDE + q * (DT_R / DT_Q);
                const E4* wq = ws + q * (DT_R / DT_Q);
                const bool b1 = c0 + c < n_back1;
#pragma unroll 4
                for (uint32_t i = 0; i < DT_R / DT_Q; i++) {
                    const uint32_t rr = (i + q) % (DT_R / DT_Q);
                    const uint32_t x = cs[rr];
                    e4a_mac(a[0], wq[rr], x);
                    if (b1) e4a_mac(a[1], wq[rr + 1], x);
                }
            }
            cx.sync();  // the stage is free for the copy issued in the next iteration
        }
        // reduce the DT_Q row quarters of every column
        E4* red = reinterpret_cast<E4*>(sm);  // [DT_T][2]
        for (uint32_t t = cx.tid; t < DT_T; t += cx.nt) {
            const E4A* a = acc[NV == 1 ? 0 : t];
            red[t * 2] = e4a_redc(a[0]); red[t * 2 + 1] = e4a_redc(a[1]);
        }
        cx.sync();
        for (uint32_t k = cx.tid; k < DT_CG * 2; k += cx.nt) {
            const uint32_t c = k >> 1, b = k & 1;
            if (c >= ncg) continue;
            E4 sum = red[(c * DT_Q) * 2 + b];
            for (uint32_t q = 1; q < DT_Q; q++) sum = e4_add(sum, red[(c * DT_Q + q) * 2 + b]);
            partial[((uint64_t)(c0 + c) * nblk + cx.bx) * 2 + b] = sum;
        }
    }
};
// out[col*2 + b] = sum_blk partial[(col*nblk + blk)*2 + b]
struct DotReduceKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, const E4* partial, uint32_t ncols, uint32_t nblk, E4* out) {
        const uint64_t t = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (t >= (uint64_t)ncols * 2) return;
        const uint32_t col = (uint32_t)(t >> 1), b = (uint32_t)(t & 1);
        E4 acc = e4_zero();
        for (uint32_t k = 0; k < nblk; k++) acc = e4_add(acc, partial[((uint64_t)col * nblk + k) * 2 + b]);
        out[t] = acc;
    }
};

// S[k][i] = (sum_c mixpow[c] * check_coeffs[c][i])_k * 3^-bitrev(i)   (bit-reversed coefficient order kept)
struct CheckMixKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, const uint32_t* check_coeffs /*[16][n]*/, const E4* mixpow /*16*/, uint32_t* S /*[4][n]*/, uint32_t po2, RootTables rt) {
        const uint64_t n = 1ull << po2;
        const uint64_t i = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (i >= n) return;
        E4A la = e4a_zero();
        for (uint32_t c = 0; c < 16; c++) e4a_mac(la, mixpow[c], check_coeffs[(uint64_t)c * n + i]);
        const E4 acc = e4a_redc(la);
        const uint32_t un = tab_pow(rt.ip3_lo, rt.ip3_hi, brev((uint32_t)i, po2));
        for (int k = 0; k < 4; k++) S[(uint64_t)k * n + i] = fmul(acc.c[k], un);
    }
};

struct DeepMixArgs {
    const uint32_t* tr[3];   // accum, code, data traces [w][n]
    uint32_t w[3];
    uint32_t n_back1[3];     // columns [0, n_back1) of the group are in combo {0,1}, the rest in combo {0}
    const E4* mixpow;        // per register in taps order (accum, code, data), device
    const uint32_t* S;       // [4][n] check combination on D1
    const E4 *INV, *INV4;
    uint32_t* out;           // [4][n]
    E4 U0, U1a, U1b, Vc;     // combo_u: {0} -> U0 ; {0,1} -> U1a + U1b*y ; check -> Vc
    const E4* uvec;          // when non-NULL: the same four values in device memory (transcript on the device)
    uint32_t omega;          // w_N (Montgomery)
    uint32_t po2;
    RootTables rt;
};
// One thread per trace row: mix-combine all registers per tap-set, subtract the U polynomials, divide point-wise.
struct DeepMixKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, DeepMixArgs p) {
        const uint64_t n = 1ull << p.po2;
        const uint64_t i = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (i >= n) return;
        E4A l0 = e4a_zero(), l1 = e4a_zero();  // lazy 64-bit accumulators: one wide multiply-add per (register, component)
        uint32_t reg = 0;
        for (int g = 0; g < 3; g++) {
            const uint32_t* base = p.tr[g] + i;
            const uint32_t nb = p.n_back1[g] < p.w[g] ? p.n_back1[g] : p.w[g];
#pragma unroll 4
            for (uint32_t c = 0; c < nb; c++, reg++) e4a_mac(l1, p.mixpow[reg], base[(uint64_t)c * n]);
#pragma unroll 4
            for (uint32_t c = nb; c < p.w[g]; c++, reg++) e4a_mac(l0, p.mixpow[reg], base[(uint64_t)c * n]);
        }
        const E4 c0 = e4a_redc(l0), c1 = e4a_redc(l1);
        const uint32_t w = tab_pow(p.rt.f_lo, p.rt.f_hi, (uint32_t)(i << (24 - p.po2)));
        const uint32_t y = fmul(w, INV3);
        // 1/(y - z) = -3/(3z - w^i) ; 1/(y - z w^-1) = -3 w /(3z - w^(i+1))
        const uint32_t m3 = fneg(THREE);
        const E4 d0 = e4_scale(p.INV[i], m3);
        const E4 d1 = e4_scale(p.INV[(i + 1) & (n - 1)], fmul(m3, p.omega));
        const E4 U0 = p.uvec ? p.uvec[0] : p.U0, U1a = p.uvec ? p.uvec[1] : p.U1a, U1b = p.uvec ? p.uvec[2] : p.U1b, Vc = p.uvec ? p.uvec[3] : p.Vc;
        E4 r = e4_mul(e4_sub(c0, U0), d0);
        const E4 u1 = e4_add(U1a, e4_scale(U1b, y));
        r = e4_add(r, e4_mul(e4_mul(e4_sub(c1, u1), d0), d1));
        const E4 s = e4(p.S[i], p.S[n + i], p.S[2 * n + i], p.S[3 * n + i]);
        r = e4_add(r, e4_mul(e4_sub(s, Vc), p.INV4[i]));
        for (int k = 0; k < 4; k++) p.out[(uint64_t)k * n + i] = r.c[k];
    }
};

// Hal::fri_fold: out[k][idx] = (sum_{i<16} mix^i * in[.][bitrev4(i)*m + idx])_k , m = n/16.
struct FriFoldArgs { const uint32_t* in; uint32_t* out; uint32_t n; E4 mixpow[16]; };
struct FriFoldKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, FriFoldArgs p) {
        const uint32_t m = p.n / 16;
        const uint64_t idx = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (idx >= m) return;
        E4 tot = e4_zero();
#pragma unroll
        for (uint32_t i = 0; i < 16; i++) {
            const uint64_t src = (uint64_t)brev(i, 4) * m + idx;
            const E4 e = e4(p.in[src], p.in[(uint64_t)p.n + src], p.in[2ull * p.n + src], p.in[3ull * p.n + src]);
            tot = e4_add(tot, e4_mul(p.mixpow[i], e));
        }
        for (int k = 0; k < 4; k++) p.out[(uint64_t)k * m + idx] = tot.c[k];
    }
};

struct BitRevKernel {  // out[c][i] = in[c][bitrev(i)]
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, const uint32_t* in, uint32_t* out, uint32_t ncols, uint32_t lg) {
        const uint64_t t = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (t >= ((uint64_t)ncols << lg)) return;
        const uint32_t c = (uint32_t)(t >> lg), i = (uint32_t)(t & ((1u << lg) - 1));
        out[t] = in[((uint64_t)c << lg) + brev(i, lg)];
    }
};

// Merkle openings (MerkleTreeProver::prove) for all queries of all trees in one launch.
struct OpenDesc {
    const uint32_t* matrix; const uint32_t* nodes;
    uint64_t col_stride;
    uint32_t rows, cols, idx, top_size, out_off;
};
struct OpenKernel {  // grid.x = descriptors
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t*, const OpenDesc* descs, uint32_t* out) {
        const OpenDesc d = descs[cx.bx];
        uint32_t* o = out + d.out_off;
        for (uint32_t c = cx.tid; c < d.cols; c += cx.nt) o[c] = d.matrix[(uint64_t)c * d.col_stride + d.idx];
        o += d.cols;
        // sibling path: i = idx + rows; while i >= 2*top: emit nodes[i^1]; i >>= 1
        uint32_t steps = 0;
        for (uint64_t i = (uint64_t)d.idx + d.rows; i >= 2ull * d.top_size; i >>= 1) steps++;
        for (uint32_t w = cx.tid; w < steps * 8; w += cx.nt) {
            const uint32_t s = w >> 3;
            const uint64_t i = ((uint64_t)d.idx + d.rows) >> s;
            o[w] = d.nodes[(i ^ 1) * 8 + (w & 7)];
        }
    }
};

}  // namespace hf

// ---- DEEP kernels for data-defined circuits: arbitrary tap sets (<= 4 distinct backs, <= 8 distinct back-sets) -------
namespace hf {

static constexpr uint32_t DOTG_T = 128, DOTG_CPB = 2;  // 2 columns x 4 backs x 8 accumulator words per thread
struct Backs4 { uint32_t nb; uint32_t back[GEN_MAX_BACKS]; };

// partial[((col * nblk + blk) * 4) + s] = sum over the block's rows of cols[col][r] * Wt[(r + back[s]) mod n]
// for every slot s whose bit is set in colmask[col].
struct DotKernelG {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t* sm, const uint32_t* cols, uint64_t col_stride, uint32_t ncols, const uint8_t* colmask, Backs4 bk,
                       const E4* Wt, uint32_t po2, E4* partial) {
        const uint64_t n = 1ull << po2;
        const uint32_t nblk = cx.gx, c0 = cx.by * DOTG_CPB;
        const uint64_t row0 = (uint64_t)cx.bx * DOT_RPB;
        E4* red = reinterpret_cast<E4*>(sm);  // [DOTG_T][DOTG_CPB * 4]
        for (uint32_t it = cx.tid; it < DOTG_T; it += cx.nt) {
            E4A acc[DOTG_CPB][GEN_MAX_BACKS];  // lazy 64-bit accumulators (field.cuh)
            uint32_t mask[DOTG_CPB];
            for (uint32_t c = 0; c < DOTG_CPB; c++) {
                mask[c] = c0 + c < ncols ? colmask[c0 + c] : 0u;
                for (uint32_t s = 0; s < GEN_MAX_BACKS; s++) acc[c][s] = e4a_zero();
            }
            for (uint64_t r = row0 + it; r < row0 + DOT_RPB && r < n; r += DOTG_T) {
                E4 w[GEN_MAX_BACKS];
#pragma unroll
                for (uint32_t s = 0; s < GEN_MAX_BACKS; s++) w[s] = s < bk.nb ? Wt[(r + bk.back[s]) & (n - 1)] : e4_zero();
#pragma unroll
                for (uint32_t c = 0; c < DOTG_CPB; c++) {
                    if (c0 + c >= ncols) break;
                    const uint32_t t = cols[(uint64_t)(c0 + c) * col_stride + r];
#pragma unroll
                    for (uint32_t s = 0; s < GEN_MAX_BACKS; s++)
                        if ((mask[c] >> s) & 1u) e4a_mac(acc[c][s], w[s], t);
                }
            }
            for (uint32_t c = 0; c < DOTG_CPB; c++)
                for (uint32_t s = 0; s < GEN_MAX_BACKS; s++) red[(it * DOTG_CPB + c) * GEN_MAX_BACKS + s] = e4a_redc(acc[c][s]);
        }
        cx.sync();
        const uint32_t K = DOTG_CPB * GEN_MAX_BACKS;
        for (uint32_t stride = DOTG_T / 2; stride >= 1; stride >>= 1) {
            for (uint32_t w = cx.tid; w < stride * K; w += cx.nt) {
                const uint32_t it = w / K, k = w % K;
                red[it * K + k] = e4_add(red[it * K + k], red[(it + stride) * K + k]);
            }
            cx.sync();
        }
        for (uint32_t k = cx.tid; k < K; k += cx.nt) {
            const uint32_t c = k / GEN_MAX_BACKS, s = k % GEN_MAX_BACKS;
            if (c0 + c < ncols) partial[((uint64_t)(c0 + c) * nblk + cx.bx) * GEN_MAX_BACKS + s] = red[k];
        }
    }
};
struct DotReduceKernelG {  // out[col * 4 + s] = sum_blk partial[(col * nblk + blk) * 4 + s]
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, const E4* partial, uint32_t ncols, uint32_t nblk, E4* out) {
        const uint64_t t = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (t >= (uint64_t)ncols * GEN_MAX_BACKS) return;
        const uint32_t col = (uint32_t)(t / GEN_MAX_BACKS), s = (uint32_t)(t % GEN_MAX_BACKS);
        E4 acc = e4_zero();
        for (uint32_t k = 0; k < nblk; k++) acc = e4_add(acc, partial[((uint64_t)col * nblk + k) * GEN_MAX_BACKS + s]);
        out[t] = acc;
    }
};

struct DeepMixGArgs {
    const uint32_t* tr[3];     // accum, code, data traces [w][n]
    const uint32_t* regcol;    // registers sorted by combo: group << 28 | offset
    const E4* regmix;          // their mix powers, same order
    uint32_t combo_start[GEN_MAX_COMBOS + 1];
    uint32_t n_combos;
    uint32_t combo_nb[GEN_MAX_COMBOS];
    uint32_t combo_back[GEN_MAX_COMBOS][GEN_MAX_BACKS];   // back values of the combo
    uint32_t combo_fb[GEN_MAX_COMBOS][GEN_MAX_BACKS];     // -3 w^back (Montgomery): 1/(y_i - z w^-b) = fb * INV[(i + b) mod n]
    E4 U[GEN_MAX_COMBOS][GEN_MAX_BACKS];                  // combo_u: U_c(y) = sum_k U[c][k] y^k
    E4 Vc;
    const uint32_t* S; const E4 *INV, *INV4; uint32_t* out;
    uint32_t po2;
    RootTables rt;
};
struct DeepMixKernelG {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, DeepMixGArgs p) {
        const uint64_t n = 1ull << p.po2;
        const uint64_t i = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (i >= n) return;
        const uint32_t w = tab_pow(p.rt.f_lo, p.rt.f_hi, (uint32_t)(i << (24 - p.po2)));
        const uint32_t y = fmul(w, INV3);
        E4 r = e4_zero();
        for (uint32_t c = 0; c < p.n_combos; c++) {
            E4A lt = e4a_zero();
#pragma unroll 4
            for (uint32_t k = p.combo_start[c]; k < p.combo_start[c + 1]; k++) {
                const uint32_t rc = p.regcol[k];
                e4a_mac(lt, p.regmix[k], p.tr[rc >> 28][(uint64_t)(rc & 0x0FFFFFFFu) * n + i]);
            }
            const E4 tot = e4a_redc(lt);
            E4 u = e4_zero();
            for (uint32_t k = p.combo_nb[c]; k-- > 0;) u = e4_add(e4_scale(u, y), p.U[c][k]);
            E4 q = e4_sub(tot, u);
            for (uint32_t k = 0; k < p.combo_nb[c]; k++) q = e4_scale(e4_mul(q, p.INV[(i + p.combo_back[c][k]) & (n - 1)]), p.combo_fb[c][k]);
            r = e4_add(r, q);
        }
        const E4 s = e4(p.S[i], p.S[n + i], p.S[2 * n + i], p.S[3 * n + i]);
        r = e4_add(r, e4_mul(e4_sub(s, p.Vc), p.INV4[i]));
        for (int k = 0; k < 4; k++) p.out[(uint64_t)k * n + i] = r.c[k];
    }
};

}  // namespace hf
