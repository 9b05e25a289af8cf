// Fiat-Shamir transcript ON THE DEVICE (risc0-zkp 3.0.4 `WriteIOP` + `Poseidon2Rng`; upstream
// runs it on the host between HAL calls).  Every transcript step of a segment -- commit a tree's top layer, mix its root, draw
// challenges, derive the parameters the next kernels need, interpolate the tap polynomials, hash what was written, draw the 50
// query positions and build the opening descriptors -- is a ONE-WARP kernel on the prover's stream that reads and writes device
// memory only.  The seal is assembled in a device buffer and leaves the GPU with a single copy at the end, so the host enqueues
// the whole segment without waiting for the GPU once (1 stream synchronisation per segment instead of 10).
//
// State and parameters live in two small structs in the context's arena:
//   TxState  : the 24 sponge cells of the RNG + the pool cursor
//   TxParams : the challenges and what is derived from them (powers of the mixes, DEEP point data, query positions)
// Kernels: block = 32 threads.  A permutation is warp-cooperative (poseidon2.cuh wp2_mix: lane L < 24 owns cell L); the host
// emulator runs the same bodies with one thread and the scalar permutation.  Work-item loops `for (i = cx.tid; ...; i += cx.nt)`
// between cx.sync() as everywhere else.
#pragma once
#include "poseidon2.cuh"
#include "deep.cuh"

namespace hf {

static constexpr uint32_t TX_MAX_FRI_ROUNDS = 5, TX_QUERIES = 50;

struct TxState { uint32_t cells[24]; uint32_t pool_used; uint32_t pad_[7]; };
struct TxParams {
    E4 poly_mix, z, z3, z4, A, deep_mix;
    E4 uvec[4];                         // U0, U1a, U1b, Vc of the DEEP quotient (DeepMixKernel)
    E4 xs[24];                          // z4^(2^k): PowBitrevKernel
    E4 fri_mixpow[TX_MAX_FRI_ROUNDS][16];
    uint32_t positions[TX_QUERIES];
    uint32_t pad_[2];
};
// where the parity checkpoints land in the device checkpoint buffer (words)
enum : uint32_t { CP_CODE_ROOT = 0, CP_DATA_ROOT = 8, CP_ACCUM_ROOT = 16, CP_CHECK_ROOT = 24, CP_POLY_MIX = 32, CP_Z = 36, CP_HASH_U = 40, CP_DEEP_MIX = 48,
                  CP_FRI_FINAL_HASH = 52, CP_FRI_ROOT0 = 60 /* 8 per round */, CP_FRI_MIX0 = 100 /* 4 per round */, CP_POSITIONS = 120, CP_ACCUM_MIX = 176, CP_WORDS = 176 + 4096 };

// ---- primitives (uniform across the warp: every lane follows the same control flow) ---------------------------------
HD void tx_permute(const KCtx& cx, uint32_t* cells) {
#ifdef __CUDA_ARCH__
    const int lane = cx.tid & 31;
    const WarpConsts w = wp2_consts(lane);
    uint32_t x = lane < 24 ? cells[lane] : 0u;
    x = wp2_mix(x, lane, w);
    if (lane < 24) cells[lane] = x;
    __syncwarp();
#else
    (void)cx;
    p2_host_consts();
    p2_mix(cells);
#endif
}
// Poseidon2Rng::mix(digest)
HD void tx_mix(const KCtx& cx, TxState* t, const uint32_t* d8) {
    if (t->pool_used != 0) {  // switching from squeezing: one permutation first (uniform across the warp)
        cx.sync();
        tx_permute(cx, t->cells);
        cx.sync();
        if (cx.tid == 0) t->pool_used = 0;
        cx.sync();
    }
    for (uint32_t i = cx.tid; i < 8; i += cx.nt) t->cells[i] = fadd(t->cells[i], d8[i]);
    cx.sync();
    tx_permute(cx, t->cells);
    if (cx.tid == 0) t->pool_used = 0;
    cx.sync();
}
// Poseidon2Rng::random_elem (every lane returns the same value)
HD uint32_t tx_elem(const KCtx& cx, TxState* t) {
    if (t->pool_used == 16) {
        cx.sync();
        tx_permute(cx, t->cells);
        if (cx.tid == 0) t->pool_used = 0;
        cx.sync();
    }
    const uint32_t u = t->pool_used;
    const uint32_t v = t->cells[u];
    cx.sync();
    if (cx.tid == 0) t->pool_used = u + 1;
    cx.sync();
    return v;
}
HD E4 tx_ext(const KCtx& cx, TxState* t) { const uint32_t a = tx_elem(cx, t), b = tx_elem(cx, t), c = tx_elem(cx, t), d = tx_elem(cx, t); return e4(a, b, c, d); }
HD uint32_t tx_bits(const KCtx& cx, TxState* t, unsigned bits) {
    uint32_t v = from_mont(tx_elem(cx, t));
    for (int i = 0; i < 3; i++) { const uint32_t n = from_mont(tx_elem(cx, t)); if (v == 0) v = n; }
    return v & (uint32_t)((1ull << bits) - 1);
}
// unpadded_hash over n words of src (sponge state st[24] in shared memory), digest -> out8 (may be global)
HD void tx_hash(const KCtx& cx, uint32_t* st, const uint32_t* src, uint32_t n, uint32_t* out8) {
    for (uint32_t k = cx.tid; k < 24; k += cx.nt) st[k] = 0;
    cx.sync();
    const uint32_t blocks = n == 0 ? 1u : (n + 15u) / 16u;
    for (uint32_t b = 0; b < blocks; b++) {
        for (uint32_t k = cx.tid; k < 16; k += cx.nt) st[k] = b * 16 + k < n ? src[b * 16 + k] : 0u;
        cx.sync();
        tx_permute(cx, st);
        cx.sync();
    }
    for (uint32_t k = cx.tid; k < 8; k += cx.nt) out8[k] = st[k];
    cx.sync();
}
// out[j] = base^j, j < n: lane L takes j = L, L + nt, ... (start base^L, step base^nt)
HD void tx_powers(const KCtx& cx, const E4& base, E4* out, uint32_t n) {
    E4 start = e4_one(), step = e4_one();
    for (int i = 0; i < cx.nt; i++) { if (i < cx.tid) start = e4_mul(start, base); step = e4_mul(step, base); }
    for (uint32_t j = cx.tid; j < n; j += cx.nt) { out[j] = start; start = e4_mul(start, step); }
}

// ---- kernels -----------------------------------------------------------------------------------------------------------
// MerkleTreeProver::commit: the tree's top layer into the seal, the root into the transcript (+ the parity checkpoint)
struct TxCommitKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t*, TxState* t, const uint32_t* nodes, uint32_t top_size, uint32_t* seal_dst, uint32_t* cp_root) {
        for (uint32_t i = cx.tid; i < top_size * 8; i += cx.nt) seal_dst[i] = nodes[top_size * 8 + i];
        for (uint32_t i = cx.tid; i < 8; i += cx.nt) cp_root[i] = nodes[8 + i];
        tx_mix(cx, t, nodes + 8);
    }
};
// the accum mix: n elements (after the DATA commit)
struct TxDrawElemsKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t*, TxState* t, uint32_t* out, uint32_t n, uint32_t* cp) {
        for (uint32_t i = 0; i < n; i++) { const uint32_t v = tx_elem(cx, t); if (cx.tid == 0) { out[i] = v; cp[i] = v; } }
    }
};
// poly_mix and its powers for eval_check (after the ACCUM commit)
struct TxPolyMixKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t*, TxState* t, TxParams* p, E4* mixpow, uint32_t n, uint32_t* cp) {
        const E4 m = tx_ext(cx, t);
        if (cx.tid == 0) { p->poly_mix = m; for (int k = 0; k < 4; k++) cp[k] = m.c[k]; }
        tx_powers(cx, m, mixpow, n);
    }
};
// z and what DeepWeightsKernel / PowBitrevKernel need (after the CHECK commit).  n_inv = Montgomery form of (2^po2)^-1.
struct TxDeepPointKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t*, TxState* t, TxParams* p, uint32_t po2, uint32_t n_inv, uint32_t* cp) {
        const E4 z = tx_ext(cx, t);
        if (cx.tid != 0) return;
        const E4 z3 = e4_scale(z, THREE), z4 = e4_pow(z, 4);
        E4 A = z3;
        for (uint32_t k = 0; k < po2; k++) A = e4_sqr(A);  // (3z)^N
        A.c[0] = fsub(A.c[0], ONE);
        A = e4_scale(A, n_inv);
        p->z = z; p->z3 = z3; p->z4 = z4; p->A = A;
        E4 cur = z4;
        for (uint32_t k = 0; k < po2; k++) { p->xs[k] = cur; cur = e4_sqr(cur); }
        for (int k = 0; k < 4; k++) cp[k] = z.c[k];
    }
};
struct DeepWeightsKernelD {  // DeepWeightsKernel with the point read from TxParams
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t* sm, E4* INV, E4* L, E4* INV4, const TxParams* p, uint32_t po2, RootTables rt) {
        DeepWeightsKernel::run(cx, sm, INV, L, INV4, p->z3, p->z4, p->A, po2, rt);
    }
};

// coeff_u (per register the interpolant through its tap evaluations; the 16 check evaluations), written to the seal; its hash
// into the transcript; then deep_mix, the per-register mix powers and the combination constants of the DEEP quotient.
// Built-in circuit: tap sets {0} and {0,1} only.  evals[goff[g] + 2 c + b], check evals at evals[goff[3] + 2 c].
struct TxCoeffUArgs {
    TxState* t; TxParams* p;
    const E4* evals;
    uint32_t goff[4], w[3], back1[3];
    uint32_t n_taps;          // T
    uint32_t back_one;        // w_N^-1 (Montgomery)
    E4* coeff_u;              // [T + 16] (16-byte aligned scratch)
    uint32_t* seal_dst;       // 4 (T + 16) words of the seal
    E4* regmix;               // [W + 16]
    uint32_t* cp;             // checkpoint buffer base
};
struct TxCoeffUKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t* sm, TxCoeffUArgs a) {
        uint32_t* st = sm;                                  // 24 sponge cells
        E4* part = reinterpret_cast<E4*>(sm + 32);          // [nt][4] partial U sums
        const uint32_t W = a.w[0] + a.w[1] + a.w[2], T = a.n_taps;
        const E4 z = a.p->z;
        const E4 x0 = z, x1 = e4_scale(z, a.back_one);
        const E4 dinv = e4_inv(e4_sub(x0, x1));
        // tap index of (group g, column c): two-tap columns come first in every group
        uint32_t tbase[3];
        { uint32_t tt = 0; for (int g = 0; g < 3; g++) { tbase[g] = tt; const uint32_t nb = a.back1[g] < a.w[g] ? a.back1[g] : a.w[g]; tt += 2 * nb + (a.w[g] - nb); } }
        uint32_t rbase[3] = {0, a.w[0], a.w[0] + a.w[1]};
        for (uint32_t reg = cx.tid; reg < W; reg += cx.nt) {
            const int g = reg < rbase[1] ? 0 : (reg < rbase[2] ? 1 : 2);
            const uint32_t c = reg - rbase[g];
            const uint32_t nb = a.back1[g] < a.w[g] ? a.back1[g] : a.w[g];
            const E4 u0 = a.evals[a.goff[g] + 2 * c], u1 = a.evals[a.goff[g] + 2 * c + 1];
            if (c < nb) {
                const uint32_t tap = tbase[g] + 2 * c;
                const E4 c1 = e4_mul(e4_sub(u0, u1), dinv);
                a.coeff_u[tap] = e4_sub(u0, e4_mul(c1, x0));
                a.coeff_u[tap + 1] = c1;
            } else {
                a.coeff_u[tbase[g] + 2 * nb + (c - nb)] = u0;
            }
        }
        for (uint32_t c = cx.tid; c < 16; c += cx.nt) a.coeff_u[T + c] = a.evals[a.goff[3] + 2 * c];
#ifdef __CUDA_ARCH__
        __threadfence_block();
#endif
        cx.sync();
        for (uint32_t i = cx.tid; i < 4 * (T + 16); i += cx.nt) a.seal_dst[i] = reinterpret_cast<const uint32_t*>(a.coeff_u)[i];
        tx_hash(cx, st, reinterpret_cast<const uint32_t*>(a.coeff_u), 4 * (T + 16), a.cp + CP_HASH_U);
        tx_mix(cx, a.t, a.cp + CP_HASH_U);
        const E4 dmix = tx_ext(cx, a.t);
        if (cx.tid == 0) { a.p->deep_mix = dmix; for (int k = 0; k < 4; k++) a.cp[CP_DEEP_MIX + k] = dmix.c[k]; }
        tx_powers(cx, dmix, a.regmix, W + 16);
#ifdef __CUDA_ARCH__
        __threadfence_block();
#endif
        cx.sync();
        // U0 = sum over one-tap registers of mix * coeff; U1a / U1b over two-tap registers; Vc over the check evaluations
        E4 s0 = e4_zero(), s1a = e4_zero(), s1b = e4_zero(), sv = e4_zero();
        for (uint32_t reg = cx.tid; reg < W; reg += cx.nt) {
            const int g = reg < rbase[1] ? 0 : (reg < rbase[2] ? 1 : 2);
            const uint32_t c = reg - rbase[g];
            const uint32_t nb = a.back1[g] < a.w[g] ? a.back1[g] : a.w[g];
            const E4 m = a.regmix[reg];
            if (c < nb) {
                const uint32_t tap = tbase[g] + 2 * c;
                s1a = e4_add(s1a, e4_mul(m, a.coeff_u[tap])); s1b = e4_add(s1b, e4_mul(m, a.coeff_u[tap + 1]));
            } else s0 = e4_add(s0, e4_mul(m, a.coeff_u[tbase[g] + 2 * nb + (c - nb)]));
        }
        for (uint32_t c = cx.tid; c < 16; c += cx.nt) sv = e4_add(sv, e4_mul(a.regmix[W + c], a.coeff_u[T + c]));
        part[cx.tid * 4 + 0] = s0; part[cx.tid * 4 + 1] = s1a; part[cx.tid * 4 + 2] = s1b; part[cx.tid * 4 + 3] = sv;
        cx.sync();
        if (cx.tid == 0) {
            for (int k = 0; k < 4; k++) {
                E4 acc = e4_zero();
                for (int l = 0; l < cx.nt; l++) acc = e4_add(acc, part[l * 4 + k]);
                a.p->uvec[k] = acc;
            }
        }
    }
};
// a FRI round's commit: top layer, root, then the fold mix and its 16 powers
struct TxFriCommitKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t* sm, TxState* t, TxParams* p, const uint32_t* nodes, uint32_t top_size, uint32_t* seal_dst, uint32_t round, uint32_t* cp) {
        TxCommitKernel::run(cx, sm, t, nodes, top_size, seal_dst, cp + CP_FRI_ROOT0 + 8 * round);
        const E4 fm = tx_ext(cx, t);
        if (cx.tid == 0) for (int k = 0; k < 4; k++) cp[CP_FRI_MIX0 + 4 * round + k] = fm.c[k];
        tx_powers(cx, fm, p->fri_mixpow[round], 16);
    }
};
struct FriFoldKernelD {  // FriFoldKernel with the mix powers read from TxParams
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, const uint32_t* in, uint32_t* out, uint32_t n, const E4* mixpow) {
        const uint32_t m = n / 16;
        const uint64_t idx = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (idx >= m) return;
        E4 tot = e4_zero();
#pragma unroll
        for (uint32_t i = 0; i < 16; i++) {
            const uint64_t src = (uint64_t)brev(i, 4) * m + idx;
            const E4 e = e4(in[src], in[(uint64_t)n + src], in[2ull * n + src], in[3ull * n + src]);
            tot = e4_add(tot, e4_mul(mixpow[i], e));
        }
        for (int k = 0; k < 4; k++) out[(uint64_t)k * m + idx] = tot.c[k];
    }
};
// final coefficients into the seal + their hash into the transcript; the 50 query positions; one opening descriptor per
// (query, tree): the four main trees at `pos`, every FRI round at pos mod rows_r (the seal offsets are static)
struct TxTreeInfo { const uint32_t* matrix; const uint32_t* nodes; uint64_t col_stride; uint32_t rows, cols, top_size, pad_; };
struct TxQueriesArgs {
    TxState* t; TxParams* p;
    const uint32_t* final_nat; uint32_t final_words;   // natural-order final coefficients (4 * n_final words)
    uint32_t* seal_final;                               // where they go in the seal
    const TxTreeInfo* trees; uint32_t n_main, n_rounds; // n_main = 4 main trees first, then the rounds
    uint32_t domain_bits;                               // log2(4N)
    OpenDesc* descs;                                    // [50 * (n_main + n_rounds)]
    uint32_t query_words;                               // seal words per query
    uint32_t* cp;
};
struct TxQueriesKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t* sm, TxQueriesArgs a) {
        uint32_t* st = sm;
        for (uint32_t i = cx.tid; i < a.final_words; i += cx.nt) a.seal_final[i] = a.final_nat[i];
#ifdef __CUDA_ARCH__
        __threadfence_block();
#endif
        cx.sync();
        tx_hash(cx, st, a.final_nat, a.final_words, a.cp + CP_FRI_FINAL_HASH);
        tx_mix(cx, a.t, a.cp + CP_FRI_FINAL_HASH);
        for (uint32_t q = 0; q < TX_QUERIES; q++) {
            const uint32_t pos = tx_bits(cx, a.t, a.domain_bits);
            if (cx.tid == 0) { a.p->positions[q] = pos; a.cp[CP_POSITIONS + q] = pos; }
        }
        cx.sync();
        const uint32_t per_q = a.n_main + a.n_rounds;
        for (uint32_t q = cx.tid; q < TX_QUERIES; q += cx.nt) {
            uint32_t pos = a.p->positions[q];
            uint32_t off = q * a.query_words;
            for (uint32_t k = 0; k < per_q; k++) {
                const TxTreeInfo ti = a.trees[k];
                if (k >= a.n_main) pos %= ti.rows;
                uint32_t layers = 0; while ((1u << layers) < ti.rows) layers++;
                uint32_t top_layer = 0; while ((1u << top_layer) < ti.top_size) top_layer++;
                a.descs[q * per_q + k] = OpenDesc{ti.matrix, ti.nodes, ti.col_stride, ti.rows, ti.cols, pos, ti.top_size, off};
                off += ti.cols + 8 * (layers - top_layer);
            }
        }
    }
};

}  // namespace hf
