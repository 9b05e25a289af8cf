// BabyBear NTT / iNTT / coset-LDE kernels for sm_100a.
//
// Replaces the risc0-zkp 3.0.4 `Hal` ops behind `Prover::commit_group` and `PolyGroup::new`
// (`batch_interpolate_ntt`, `zk_shift`, `batch_expand_into_evaluate_ntt`; upstream GPU impl: sppark
// batch_NTT/iNTT inside risc0-sys 1.5.0 -- /root/reference/Cargo.lock:3174-3223, not vendored;
// conventions per SURVEY.md Appendix A.2).  Not a port: the transform is restructured for B200 as
//
//   size 2^n = 2^a (contiguous "chunk" dimension, lo)  x  2^b (strided dimension, hi)
//
//   iNTT  : [strided DIF over hi] -> x w^-(rev(hi) lo) -> [chunk DIF over lo] -> x n^-1 3^j
//   LDE x4: replicate x4 -> [chunk DIT over lo, first 2 levels skipped] -> x w^(rev(hi) lo) -> [strided DIT over hi]
//
// so that (1) every global access is a coalesced run (whole 2^a chunks, or 2^c-word segments of 2^b
// strided rows staged through shared memory), (2) all butterflies run out of shared memory / registers
// in radix-2^R rounds, and (3) the two chunk stages of iNTT and LDE are FUSED into one kernel ("middle"):
// the coefficient form never goes to HBM for the main groups (the DEEP step reads the trace instead).
// HBM traffic per trace element: 8 B (strided DIF) + 20 B (middle) + 32 B (strided DIT) = 60 B.
#pragma once
#include "dev.cuh"

namespace hf {

// Two-level power tables (4096 entries each, global memory, L1/L2 resident):
//   w24 = primitive 2^24-th root: x^E = hi[E >> 12] * lo[E & 4095] for E < 2^24.
struct RootTables {
    const uint32_t *f_lo, *f_hi;      // w24^i, w24^(4096 i)
    const uint32_t *i_lo, *i_hi;      // w24^-i, ...
    const uint32_t *p3_lo, *p3_hi;    // 3^i, 3^(4096 i)
    const uint32_t *ip3_lo, *ip3_hi;  // 3^-i, ...
};
HD uint32_t tab_pow(const uint32_t* lo, const uint32_t* hi, uint32_t E) { return fmul(hi[E >> 12], lo[E & 4095u]); }

// One pad word per 32 keeps the radix rounds (stride 2^l0 register gathers) off the same bank.
HD uint32_t padi(uint32_t i) { return i + (i >> 5); }
HD uint32_t padded_words(uint32_t n) { return n + (n >> 5) + 1; }

// One radix-2^R register round over the `pos` dimension of a [2^k][2^c] shared-memory tile
// (element (pos, batch) at padi(pos << c | batch)).  Covers butterfly levels l0+1 .. l0+R.
// tw[i] = w_{2^k}^(+-i), i < 2^(k-1).  INV: Gentleman-Sande (a+b, (a-b)w); else Cooley-Tukey (a+bw, a-bw).
template <int R, bool INV>
HD void ntt_round(const KCtx& cx, uint32_t* s, const uint32_t* tw, int k, int c, int l0) {
    const uint32_t items = 1u << (k - R + c);
    const uint32_t cmask = (1u << c) - 1u, lmask = (1u << l0) - 1u;
    for (uint32_t it = cx.tid; it < items; it += cx.nt) {
        const uint32_t batch = it & cmask, t = it >> c;
        const uint32_t low = t & lmask, high = t >> l0;
        const uint32_t base = (high << (l0 + R)) | low;
        uint32_t v[1 << R];
#pragma unroll
        for (int j = 0; j < (1 << R); j++) v[j] = s[padi(((base + ((uint32_t)j << l0)) << c) + batch)];
        if (!INV) {
#pragma unroll
            for (int q = 1; q <= R; q++) {
                const int h = 1 << (q - 1);
#pragma unroll
                for (int j = 0; j < (1 << R); j++) {
                    if (j & h) continue;
                    const uint32_t w = tw[(low + ((uint32_t)(j & (h - 1)) << l0)) << (k - l0 - q)];
                    const uint32_t x = fmul(v[j + h], w);
                    v[j + h] = fsub(v[j], x);
                    v[j] = fadd(v[j], x);
                }
            }
        } else {
#pragma unroll
            for (int q = R; q >= 1; q--) {
                const int h = 1 << (q - 1);
#pragma unroll
                for (int j = 0; j < (1 << R); j++) {
                    if (j & h) continue;
                    const uint32_t w = tw[(low + ((uint32_t)(j & (h - 1)) << l0)) << (k - l0 - q)];
                    const uint32_t a = v[j], b = v[j + h];
                    v[j] = fadd(a, b);
                    v[j + h] = fmul(fsub(a, b), w);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < (1 << R); j++) s[padi(((base + ((uint32_t)j << l0)) << c) + batch)] = v[j];
    }
}

template <bool INV>
HD void ntt_round_dyn(const KCtx& cx, uint32_t* s, const uint32_t* tw, int k, int c, int l0, int R) {
    switch (R) {
        case 1: ntt_round<1, INV>(cx, s, tw, k, c, l0); break;
        case 2: ntt_round<2, INV>(cx, s, tw, k, c, l0); break;
        case 3: ntt_round<3, INV>(cx, s, tw, k, c, l0); break;
        default: ntt_round<4, INV>(cx, s, tw, k, c, l0); break;
    }
}

// Full transform of the tile: levels skip+1..k (forward) or k..1 (inverse), in rounds of <= rmax levels.
// Every round ends with a barrier.
template <bool INV>
HD void ntt_tile(const KCtx& cx, uint32_t* s, const uint32_t* tw, int k, int c, int skip, int rmax) {
    const int L = k - skip;
    if (L <= 0) return;
    const int nr = (L + rmax - 1) / rmax;
    int done = 0;
    for (int r = 0; r < nr; r++) {
        const int R = L / nr + (r < L % nr ? 1 : 0);
        const int l0 = INV ? (k - done - R) : (skip + done);
        ntt_round_dyn<INV>(cx, s, tw, k, c, l0, R);
        cx.sync();
        done += R;
    }
}

enum : uint32_t { MID_INTT = 1, MID_SHIFT = 2, MID_FWD = 4, MID_GFLY = 8 };

struct MidArgs {
    const uint32_t* in;
    uint32_t* out;
    uint64_t in_stride, out_stride;  // column strides (words)
    uint32_t ncols, cols_per_block;
    int n, a, b;   // 2^n coefficients per column = 2^b chunks of 2^a
    int e;         // expand bits of the forward part (0 or 2)
    uint32_t flags;
    int rmax_inv, rmax_fwd;
    uint32_t n_inv;  // Montgomery form of (2^n)^-1
    RootTables rt;
};

struct MidLayout {
    uint32_t A, B, twI, twF, G3, Gs, G2, total;  // word offsets
    HD MidLayout(const MidArgs& p) {
        uint32_t o = 0;
        const bool intt = p.flags & MID_INTT, fwd = p.flags & MID_FWD, fly = p.flags & MID_GFLY;
        A = o; if (intt) o += padded_words(1u << p.a);
        B = o; if (fwd) o += padded_words(1u << (p.a + p.e));
        twI = o; if (intt) o += (p.a > 0 ? 1u << (p.a - 1) : 1u);
        twF = o; if (fwd) o += 1u << (p.a + p.e - 1);
        G3 = o; if (intt && p.b > 0 && !fly) o += 1u << p.a;
        Gs = o; if (intt && !fly) o += 1u << p.a;
        G2 = o; if (fwd && p.b > 0 && !fly) o += 1u << (p.a + p.e);
        total = o;
    }
};

// grid.x = 2^b chunks, grid.y = column groups.
struct MiddleKernel {
    static constexpr bool kBarrier = true;
    HD static uint32_t g3(const MidArgs& p, uint32_t rb, uint32_t i) { return tab_pow(p.rt.i_lo, p.rt.i_hi, (rb * i) << (24 - p.n)); }
    HD static uint32_t gs(const MidArgs& p, uint32_t rb, uint32_t i) {
        uint32_t g = p.n_inv;
        if (p.flags & MID_SHIFT) g = fmul(g, tab_pow(p.rt.p3_lo, p.rt.p3_hi, (brev(i, p.a) << p.b) + rb));
        return g;
    }
    HD static uint32_t g2(const MidArgs& p, uint32_t rb, uint32_t i) { return tab_pow(p.rt.f_lo, p.rt.f_hi, (rb * i) << (24 - p.n - p.e)); }

    HD static void run(const KCtx& cx, uint32_t* sm, MidArgs p) {
        const MidLayout L(p);
        const bool intt = p.flags & MID_INTT, fwd = p.flags & MID_FWD, fly = p.flags & MID_GFLY;
        const uint32_t hi = cx.bx, rb = brev(hi, p.b);
        const uint32_t na = 1u << p.a, nf = 1u << (p.a + p.e);
        uint32_t *A = sm + L.A, *B = sm + L.B, *twI = sm + L.twI, *twF = sm + L.twF, *G3 = sm + L.G3, *Gs = sm + L.Gs, *G2 = sm + L.G2;
        if (intt) {
            for (uint32_t i = cx.tid; i < (na >> 1); i += cx.nt) twI[i] = tab_pow(p.rt.i_lo, p.rt.i_hi, i << (24 - p.a));
            if (!fly) {
                if (p.b > 0) for (uint32_t i = cx.tid; i < na; i += cx.nt) G3[i] = g3(p, rb, i);
                for (uint32_t i = cx.tid; i < na; i += cx.nt) Gs[i] = gs(p, rb, i);
            }
        }
        if (fwd) {
            for (uint32_t i = cx.tid; i < (nf >> 1); i += cx.nt) twF[i] = tab_pow(p.rt.f_lo, p.rt.f_hi, i << (24 - p.a - p.e));
            if (!fly && p.b > 0) for (uint32_t i = cx.tid; i < nf; i += cx.nt) G2[i] = g2(p, rb, i);
        }
        cx.sync();
        for (uint32_t cc = 0; cc < p.cols_per_block; cc++) {
            const uint32_t col = cx.by * p.cols_per_block + cc;
            if (col >= p.ncols) break;
            const uint32_t* src = p.in + (uint64_t)col * p.in_stride + ((uint64_t)hi << p.a);
            if (intt) {
                for (uint32_t i = cx.tid; i < na; i += cx.nt) {
                    uint32_t v = src[i];
                    if (p.b > 0) v = fmul(v, fly ? g3(p, rb, i) : G3[i]);
                    A[padi(i)] = v;
                }
                cx.sync();
                ntt_tile<true>(cx, A, twI, p.a, 0, 0, p.rmax_inv);
                if (!fwd) {
                    uint32_t* dst = p.out + (uint64_t)col * p.out_stride + ((uint64_t)hi << p.a);
                    for (uint32_t i = cx.tid; i < na; i += cx.nt) dst[i] = fmul(A[padi(i)], fly ? gs(p, rb, i) : Gs[i]);
                } else {
                    for (uint32_t i = cx.tid; i < na; i += cx.nt) {
                        const uint32_t v = fmul(A[padi(i)], fly ? gs(p, rb, i) : Gs[i]);
                        for (uint32_t r = 0; r < (1u << p.e); r++) B[padi((i << p.e) + r)] = v;
                    }
                }
            } else {
                for (uint32_t i = cx.tid; i < na; i += cx.nt) {
                    const uint32_t v = src[i];
                    for (uint32_t r = 0; r < (1u << p.e); r++) B[padi((i << p.e) + r)] = v;
                }
            }
            if (fwd) {
                cx.sync();
                ntt_tile<false>(cx, B, twF, p.a + p.e, 0, p.e, p.rmax_fwd);
                uint32_t* dst = p.out + (uint64_t)col * p.out_stride + ((uint64_t)hi << (p.a + p.e));
                for (uint32_t i = cx.tid; i < nf; i += cx.nt) {
                    uint32_t v = B[padi(i)];
                    if (p.b > 0) v = fmul(v, fly ? g2(p, rb, i) : G2[i]);
                    dst[i] = v;
                }
            }
            cx.sync();
        }
    }
};

// =====================================================================================================
// v2 kernels: the first round of every stage reads straight from global memory into registers and the last
// round writes straight back (no staging pass, many loads in flight per thread), at most 16 values per
// thread, and the chunk-stage CTA is split into 64-thread units that share the per-chunk twiddle tables.
// =====================================================================================================
// IO functors of a radix round.  An element is addressed either by (pos, batch) or, when the round's stride maps
// to a constant index step ("affine"), by base index + j * step -- with compile-time sizes the steps become
// immediate offsets of the load/store instructions.
template <int PS>
HD uint32_t padx(uint32_t i) { return i + (i >> PS); }
template <int PS = 5>
struct SmemIOT {  // one pad word per 2^PS
    uint32_t* s;
    HD uint32_t base(uint32_t pos, uint32_t batch, int c) const { return padx<PS>((pos << c) + batch); }
    HD uint32_t step(int l0, int c) const { const uint32_t S = 1u << (l0 + c); return S + (S >> PS); }
    HD bool affine(int l0, int c) const { return l0 + c >= PS; }
    HD uint32_t ld_i(uint32_t i) const { return s[i]; }
    HD void st_i(uint32_t i, uint32_t v) const { s[i] = v; }
};
using SmemIO = SmemIOT<5>;
template <int A>
struct GlobStridedIO {  // element (pos, batch) of a strided tile: row pos at stride 2^a words, column batch
    const uint32_t* src; uint32_t* dst; int a_;
    HD int a() const { return A >= 0 ? A : a_; }
    HD uint32_t base(uint32_t pos, uint32_t batch, int) const { return (pos << a()) + batch; }
    HD uint32_t step(int l0, int) const { return 1u << (l0 + a()); }
    HD bool affine(int, int) const { return true; }
    HD uint32_t ld_i(uint32_t i) const { return src[i]; }
    HD void st_i(uint32_t i, uint32_t v) const { dst[i] = v; }
};

// One radix-2^R round over levels l0+1..l0+R of a [2^k][2^c] tile.  K, C, L0 >= 0 fix the sizes at compile time
// (-1 = runtime).  With L0 == 0 the twiddle exponents are compile-time and the unit twiddles (15 of the 32
// butterflies of a radix-16 round) are skipped.
// TWL: twiddle table layout. false: tw[i] = w_{2^k}^i (i < 2^(k-1)), level l reads stride 2^(k-l);  true: per-level
// blocks tw[2^(l-1) + x] = w_{2^l}^x, so consecutive lanes read consecutive words (no bank conflicts when c = 0).
// HF: lanes enumerate `high` first (only for c = 0); keeps rounds with l0 < 5 off the same banks.
// SHP: the table holds canonical (w, w' = floor(w 2^32 / p)) pairs at tw[2 i], tw[2 i + 1] and multiplies use Shoup's form.
// PF: software-pipelined loads (two register sets): the loads of the thread's next item are issued before the
// butterflies of the current one, so a global-facing round does not sit on the long scoreboard.
// geometry of one item of a radix-2^R round
template <int R, int K, int C, int L0, bool HF>
struct RoundGeom {
    int k, c, l0;
    uint32_t items, cmask, lmask, nhigh_mask;
    HD RoundGeom(int k_, int c_, int l0_) : k(K >= 0 ? K : k_), c(C >= 0 ? C : c_), l0(L0 >= 0 ? L0 : l0_) {
        items = 1u << (k - R + c);
        cmask = (1u << c) - 1u; lmask = (1u << l0) - 1u;
        nhigh_mask = (items >> (l0 + c)) - 1u;
    }
    HD void split(uint32_t it, uint32_t& batch, uint32_t& low, uint32_t& base) const {
        batch = it & cmask;
        const uint32_t t = it >> c;
        low = HF ? (t >> (k - R - l0)) : (t & lmask);
        const uint32_t high = HF ? (t & nhigh_mask) : (t >> l0);
        base = (high << (l0 + R)) | low;
    }
};
template <int R, int K, int C, int L0, bool HF, typename LD>
HD void round_load_item(const RoundGeom<R, K, C, L0, HF>& g, uint32_t it, uint32_t* v, const LD& L) {
    uint32_t batch, low, base;
    g.split(it, batch, low, base);
    if (L.affine(g.l0, g.c)) {
        const uint32_t i0 = L.base(base, batch, g.c), lstep = L.step(g.l0, g.c);
#pragma unroll
        for (int j = 0; j < (1 << R); j++) v[j] = L.ld_i(i0 + (uint32_t)j * lstep);
    } else {
#pragma unroll
        for (int j = 0; j < (1 << R); j++) v[j] = L.ld_i(L.base(base + ((uint32_t)j << g.l0), batch, g.c));
    }
}
// butterflies of levels l0+1..l0+R on the 2^R values of item `it` (already in v), then the store through S
template <int R, bool INV, int K, int C, int L0, bool TWL, bool HF, bool SHP, typename ST>
HD void round_compute_store_item(const RoundGeom<R, K, C, L0, HF>& g, const uint32_t* tw, uint32_t it, uint32_t* v, const ST& S) {
    const int k = g.k, c = g.c, l0 = g.l0;
    uint32_t batch, low, base;
    g.split(it, batch, low, base);
    if (!INV) {
#pragma unroll
        for (int q = 1; q <= R; q++) {
            const int h = 1 << (q - 1);
#pragma unroll
            for (int j = 0; j < (1 << R); j++) {
                if (j & h) continue;
                uint32_t x = v[j + h];
                if (!(L0 == 0 && (j & (h - 1)) == 0)) {
                    const uint32_t ti = TWL ? (1u << (l0 + q - 1)) + ((uint32_t)(j & (h - 1)) << l0) + low : (low + ((uint32_t)(j & (h - 1)) << l0)) << (k - l0 - q);
                    x = SHP ? fmul_pair(x, tw, ti) : fmul(x, tw[ti]);
                }
                // an output that the next level multiplies by a twiddle stays in [0, 2p): both product forms take any
                // 32-bit first operand, so its range correction is dropped (decided at compile time)
                const int hn = h << 1;
                const bool lz0 = q < R && (j & hn) && !(L0 == 0 && (j & (hn - 1)) == 0);
                const bool lz1 = q < R && ((j + h) & hn) && !(L0 == 0 && ((j + h) & (hn - 1)) == 0);
                v[j + h] = lz1 ? fsub_lazy(v[j], x) : fsub(v[j], x);
                v[j] = lz0 ? fadd_lazy(v[j], x) : fadd(v[j], x);
            }
        }
    } else {
#pragma unroll
        for (int q = R; q >= 1; q--) {
            const int h = 1 << (q - 1);
#pragma unroll
            for (int j = 0; j < (1 << R); j++) {
                if (j & h) continue;
                const uint32_t a = v[j], b = v[j + h];
                v[j] = fadd(a, b);
                uint32_t d;
                if (!(L0 == 0 && (j & (h - 1)) == 0)) {
                    const uint32_t ti = TWL ? (1u << (l0 + q - 1)) + ((uint32_t)(j & (h - 1)) << l0) + low : (low + ((uint32_t)(j & (h - 1)) << l0)) << (k - l0 - q);
                    d = fsub_lazy(a, b);  // in (0, 2p): the product takes any 32-bit first operand
                    d = SHP ? fmul_pair(d, tw, ti) : fmul(d, tw[ti]);
                } else d = fsub(a, b);
                v[j + h] = d;
            }
        }
    }
    if (S.affine(l0, c)) {
        const uint32_t i0 = S.base(base, batch, c), sstep = S.step(l0, c);
#pragma unroll
        for (int j = 0; j < (1 << R); j++) S.st_i(i0 + (uint32_t)j * sstep, v[j]);
    } else {
#pragma unroll
        for (int j = 0; j < (1 << R); j++) S.st_i(S.base(base + ((uint32_t)j << l0), batch, c), v[j]);
    }
}
template <int R, bool INV, int K, int C, int L0, bool TWL = false, bool HF = false, bool SHP = false, bool PF = false, typename LD, typename ST>
HD void round_t(const KCtx& cx, const uint32_t* tw, int k_, int c_, int l0_, const LD& L, const ST& S) {
    const RoundGeom<R, K, C, L0, HF> g(k_, c_, l0_);
    const uint32_t items = g.items;
    if constexpr (PF) {
        const uint32_t nt = (uint32_t)cx.nt;
        uint32_t v0[1 << R], v1[1 << R];
        uint32_t it = (uint32_t)cx.tid;
        if (it < items) round_load_item(g, it, v0, L);
        for (; it < items; it += 2 * nt) {
            if (it + nt < items) round_load_item(g, it + nt, v1, L);
            round_compute_store_item<R, INV, K, C, L0, TWL, HF, SHP>(g, tw, it, v0, S);
            if (it + 2 * nt < items) round_load_item(g, it + 2 * nt, v0, L);
            if (it + nt < items) round_compute_store_item<R, INV, K, C, L0, TWL, HF, SHP>(g, tw, it + nt, v1, S);
        }
    } else {
        for (uint32_t it = cx.tid; it < items; it += cx.nt) {
            uint32_t v[1 << R];
            round_load_item(g, it, v, L);
            round_compute_store_item<R, INV, K, C, L0, TWL, HF, SHP>(g, tw, it, v, S);
        }
    }
}
template <bool INV, bool TWL = false, typename LD, typename ST>
HD void round_io_dyn(const KCtx& cx, const uint32_t* tw, int k, int c, int l0, int R, const LD& L, const ST& S) {
    switch (R) {
        case 1: round_t<1, INV, -1, -1, -1, TWL>(cx, tw, k, c, l0, L, S); break;
        case 2: round_t<2, INV, -1, -1, -1, TWL>(cx, tw, k, c, l0, L, S); break;
        case 3: round_t<3, INV, -1, -1, -1, TWL>(cx, tw, k, c, l0, L, S); break;
        default: round_t<4, INV, -1, -1, -1, TWL>(cx, tw, k, c, l0, L, S); break;
    }
}

#ifndef STR_R5
#define STR_R5 1
#endif
#ifndef STR_MINB
#define STR_MINB 2
#endif
#ifndef STR_PF
#define STR_PF false  // software-pipelined global loads in the first round of the strided kernels: measured no gain on B200 (5.35 vs 5.41 ms per 192-column LDE, 120 vs 64 registers), kept as an option
#endif
struct Str2Args {
    const uint32_t* in;
    uint32_t* out;                   // may alias `in`
    uint64_t in_stride, out_stride;
    uint32_t ncols;
    int a, b, c, inv;
    int nr, R[4];
    RootTables rt;
};
// SB/SC/SA >= 0: sizes fixed at compile time (hot po2 configurations); -1: generic.
template <int SB, int SC, int SA>
struct StridedKernel2 {
    static constexpr bool kBarrier = true;
    template <bool INV>
    HD static void tile(const KCtx& cx, const Str2Args& p, uint32_t* s, const uint32_t* tw, const uint32_t* src, uint32_t* dst) {
        const GlobStridedIO<SA> G{src, dst, p.a};
        const SmemIO S{s};
#if STR_R5
        if constexpr (SB >= 6 && SB <= 10 && SB != 8) {  // 2^8 rows: 4+2+2 levels measured faster (po2 = 18 NTT stage 1.71 against 1.76 ms)
            // up to 2^10 rows: TWO rounds (one shared-memory round trip), radix up to 32 (32 values per thread, 124 registers).
            // 2^10 rows, 512 threads on the 2^15-word tile: 4.82 -> 4.71 ms per 192-column LDE against 4+3+3 levels with 1024 threads
            constexpr int R0 = (SB + 1) / 2, R1 = SB - R0;
            if (INV) {
                round_t<R0, true, SB, SC, SB - R0, false, false, true>(cx, tw, SB, SC, SB - R0, G, S); cx.sync();
                round_t<R1, true, SB, SC, 0, false, false, true>(cx, tw, SB, SC, 0, S, G);
            } else {
                round_t<R0, false, SB, SC, 0, false, false, true>(cx, tw, SB, SC, 0, G, S); cx.sync();
                round_t<R1, false, SB, SC, R0, false, false, true>(cx, tw, SB, SC, R0, S, G);
            }
            cx.sync();
        } else
#endif
        if constexpr (SB >= 6) {
            // fixed schedule: SB = R0 + R1 + R2 with R0 = 4 on the global-facing first round
            constexpr int R0 = 4, R1 = (SB - 4 + 1) / 2, R2 = SB - 4 - R1;
            if (INV) {
                round_t<R0, true, SB, SC, SB - R0, false, false, true, STR_PF>(cx, tw, SB, SC, SB - R0, G, S); cx.sync();
                round_t<R1, true, SB, SC, R2, false, false, true>(cx, tw, SB, SC, R2, S, S); cx.sync();
                round_t<(R2 > 0 ? R2 : 1), true, SB, SC, 0, false, false, true>(cx, tw, SB, SC, 0, S, G);
            } else {
                round_t<R0, false, SB, SC, 0, false, false, true, STR_PF>(cx, tw, SB, SC, 0, G, S); cx.sync();
                round_t<R1, false, SB, SC, R0, false, false, true>(cx, tw, SB, SC, R0, S, S); cx.sync();
                round_t<(R2 > 0 ? R2 : 1), false, SB, SC, R0 + R1, false, false, true>(cx, tw, SB, SC, R0 + R1, S, G);
            }
            cx.sync();
        } else {
        int done = 0;
        for (int r = 0; r < p.nr; r++) {
            const int R = p.R[r];
            const int l0 = INV ? (p.b - done - R) : done;
            const bool first = r == 0, last = r == p.nr - 1;
            if (first && last) round_io_dyn<INV>(cx, tw, p.b, p.c, l0, R, G, G);
            else if (first) round_io_dyn<INV>(cx, tw, p.b, p.c, l0, R, G, S);
            else if (last) round_io_dyn<INV>(cx, tw, p.b, p.c, l0, R, S, G);
            else round_io_dyn<INV>(cx, tw, p.b, p.c, l0, R, S, S);
            if (!last) cx.sync();
            done += R;
        }
        if (p.nr > 1) cx.sync();  // the tile buffer is reused by the next tile
        }
    }
    HD static void run(const KCtx& cx, uint32_t* sm, Str2Args p) {
        uint32_t* s = sm;
        uint32_t* tw = sm + ((padded_words(1u << (p.b + p.c)) + 1u) & ~1u);  // 8-byte aligned (w, w') pairs
        const uint32_t half = 1u << (p.b - 1);
        for (uint32_t i = cx.tid; i < half; i += cx.nt) {
            const uint32_t w = p.inv ? tab_pow(p.rt.i_lo, p.rt.i_hi, i << (24 - p.b)) : tab_pow(p.rt.f_lo, p.rt.f_hi, i << (24 - p.b));
            if (SB >= 6) { tw[2 * i] = from_mont(w); tw[2 * i + 1] = shoup_quot_mont(w); }  // Shoup pairs
            else tw[i] = w;
        }
        cx.sync();
        const uint32_t tiles_per_col = 1u << (p.a - p.c);
        const uint64_t total = (uint64_t)p.ncols * tiles_per_col;
        for (uint64_t t = cx.bx; t < total; t += cx.gx) {
            const uint32_t col = (uint32_t)(t >> (p.a - p.c));
            const uint32_t lo0 = ((uint32_t)t & (tiles_per_col - 1u)) << p.c;
            const uint32_t* src = p.in + (uint64_t)col * p.in_stride + lo0;
            uint32_t* dst = p.out + (uint64_t)col * p.out_stride + lo0;
            if (p.inv) tile<true>(cx, p, s, tw, src, dst); else tile<false>(cx, p, s, tw, src, dst);
        }
    }
};

static constexpr int MID_UT = 64;  // threads per unit of the chunk-stage kernel for chunks up to 2^10 (larger chunks: 128 / 256)
// MID_R5 = 1 (experiment, kept as a build option): the main-group chunk stage (a = 10, e = 2) as ONE WARP per (column, chunk):
// radix-32 register rounds (32 values per lane), one shared-memory round trip for the inverse part and one for the x4 forward
// part (instead of two each with radix 8/16), __syncwarp instead of named barriers, next column prefetched into registers.
// Measured on B200 (profiles/r2_mid_r5_ncu_summary.txt): 14 % fewer warp instructions (1564 M against 1813 M per 192 columns)
// but 254 registers -> 8 warps per SM; issue-active falls from 66 % to 56 % (stall "wait": two warps per scheduler cannot keep
// the heavy-multiply and ALU pipes both fed) and the stage takes 4.86 ms per 192-column LDE against 4.74 ms.  Default stays 0.
#ifndef MID_R5
#define MID_R5 0
#endif
#ifndef MID_R5_UNITS
#define MID_R5_UNITS 8
#endif

struct Mid2Args {
    const uint32_t* in;
    uint32_t* out;
    uint64_t in_stride, out_stride;
    uint32_t ncols, cols_per_block;
    int n, a, b, e;
    uint32_t flags;
    uint32_t n_inv;
    int units, alias, ut;
    int shp;         // twI / twF / G2 hold Shoup (w, w') pairs (the compile-time main-group instantiation)
    int nri, Ri[4];  // inverse rounds, top level first; the last one (RF levels) runs in registers
    int nrf, Rf[4];  // forward rounds; the first one (RF levels) runs in registers, fused with the inverse tail
    RootTables rt;
};
struct Mid2Layout {
    uint32_t twI, twF, G3, Gs, G2, unit0, unit_words, A, B, total;
    HD Mid2Layout(const Mid2Args& p) {
        uint32_t o = 0;
        const bool intt = p.flags & MID_INTT, fwd = p.flags & MID_FWD, fly = p.flags & MID_GFLY;
        const uint32_t pw = p.shp ? 2u : 1u;        // words per twiddle entry
        twI = o; if (intt) o += pw << p.a;          // per-level layout: tw[2^(l-1) + x] = w_{2^l}^-x
        twF = o; if (fwd) o += pw << (p.a + p.e);
        G3 = o; if (intt && p.b > 0 && !fly) o += 1u << p.a;
        Gs = o; if (intt && !fly) o += (padded_words(1u << p.a) + 1u) & ~1u;  // padded like the data: the register tail reads 2^RF consecutive entries per thread
        G2 = o; if (fwd && p.b > 0 && !fly) o += pw << (p.a + p.e);
        unit0 = o;
        uint32_t u = 0;
        A = u; if (intt && !(fwd && p.alias)) u += padded_words(1u << p.a);
        B = u; if (fwd) u += padded_words(1u << (p.a + p.e));
        unit_words = u;
        total = o + u * (uint32_t)p.units;
    }
};

// MA/ME >= 0: chunk size / expand bits fixed at compile time (the main-group configuration a = 10, e = 2).
template <int MA, int ME>
struct MiddleKernel2 {
    static constexpr bool kBarrier = true;
    static constexpr int PSB = (MA == 10 && ME == 2) ? (MID_R5 ? 7 : 6) : 5;  // pad shift of the forward buffer B
    static constexpr bool SHP = (MA == 10 && ME == 2);         // Shoup (w, w') pairs in twI / twF / G2 (host sets p.shp alike)
    HD static int A_(const Mid2Args& p) { return MA >= 0 ? MA : p.a; }
    HD static int E_(const Mid2Args& p) { return ME >= 0 ? ME : p.e; }
    HD static uint32_t g3(const Mid2Args& p, uint32_t rb, uint32_t i) { return tab_pow(p.rt.i_lo, p.rt.i_hi, (rb * i) << (24 - p.n)); }
    HD static uint32_t gs(const Mid2Args& p, uint32_t rb, uint32_t i) {
        uint32_t g = p.n_inv;
        if (p.flags & MID_SHIFT) g = fmul(g, tab_pow(p.rt.p3_lo, p.rt.p3_hi, (brev(i, p.a) << p.b) + rb));
        return g;
    }
    HD static uint32_t g2(const Mid2Args& p, uint32_t rb, uint32_t i) { return tab_pow(p.rt.f_lo, p.rt.f_hi, (rb * i) << (24 - p.n - p.e)); }

    struct SrcIO {  // chunk element from global, times the inter-stage twiddle w^-(rev(hi) lo)
        const uint32_t* src; const uint32_t* G3; const uint32_t *lo, *hi; uint32_t rb; int shift; int mode;  // 0 none, 1 table, 2 fly
        HD uint32_t base(uint32_t pos, uint32_t, int) const { return pos; }
        HD uint32_t step(int l0, int) const { return 1u << l0; }
        HD bool affine(int, int) const { return true; }
        HD uint32_t ld_i(uint32_t pos) const {
            uint32_t v = src[pos];
            if (mode == 1) v = fmul(v, G3[pos]); else if (mode == 2) v = fmul(v, tab_pow(lo, hi, (rb * pos) << shift));
            return v;
        }
        HD void st_i(uint32_t, uint32_t) const {}
    };
    struct DstIO {  // LDE chunk element to global, times the inter-stage twiddle w^(rev(hi) lo')
        uint32_t* dst; const uint32_t* G2; const uint32_t *lo, *hi; uint32_t rb; int shift; int mode;
        HD uint32_t base(uint32_t pos, uint32_t, int) const { return pos; }
        HD uint32_t step(int l0, int) const { return 1u << l0; }
        HD bool affine(int, int) const { return true; }
        HD uint32_t ld_i(uint32_t) const { return 0; }
        HD void st_i(uint32_t pos, uint32_t v) const {
            if (mode == 1) v = SHP ? fmul_pair(v, G2, pos) : fmul(v, G2[pos]); else if (mode == 2) v = fmul(v, tab_pow(lo, hi, (rb * pos) << shift));
            dst[pos] = v;
        }
    };

    // The register-resident tail of the inverse transform (last RF levels, scale by n^-1 3^j) fused with the
    // register-resident head of the forward transform (replicate x 2^e, first RF levels) for one item of 2^RF
    // consecutive coefficients.
    template <int RF>
    HD static void tail_load(uint32_t* v, uint32_t base, bool from_global, const SrcIO& G, const uint32_t* A) {
#pragma unroll
        for (int j = 0; j < (1 << RF); j++) v[j] = from_global ? G.ld_i(base + j) : A[padi(base + j)];
    }
    template <int RF>
    HD static void tail_compute_store(uint32_t* v, uint32_t base, const Mid2Args& p, uint32_t rb, const uint32_t* twI, const uint32_t* twF,
                                      const uint32_t* Gs, uint32_t* B, uint32_t* dst_coef, const DstIO& D) {
        const bool intt = p.flags & MID_INTT, fwd = p.flags & MID_FWD, fly = p.flags & MID_GFLY;
        if (intt) {
#pragma unroll
            for (int q = RF; q >= 1; q--) {
                const int h = 1 << (q - 1);
#pragma unroll
                for (int j = 0; j < (1 << RF); j++) {
                    if (j & h) continue;
                    const uint32_t a = v[j], b = v[j + h];
                    v[j] = fadd(a, b);
                    uint32_t d;
                    if ((j & (h - 1)) != 0) { const uint32_t ti = (1u << (q - 1)) + (uint32_t)(j & (h - 1)); d = fsub_lazy(a, b); d = SHP ? fmul_pair(d, twI, ti) : fmul(d, twI[ti]); }
                    else d = fsub(a, b);
                    v[j + h] = d;
                }
            }
#pragma unroll
            for (int j = 0; j < (1 << RF); j++) v[j] = fmul(v[j], fly ? gs(p, rb, base + j) : Gs[padi(base + j)]);
        }
        if (!fwd) {
#pragma unroll
            for (int j = 0; j < (1 << RF); j++) dst_coef[base + j] = v[j];
            return;
        }
        head<RF, true>(v, base, 0, p, twF, B, D);
        for (uint32_t r = 1; r < (1u << E_(p)); r++) head<RF, false>(v, base, r, p, twF, B, D);
    }
    // forward head for coset r: first RF levels of the x 2^e expanded chunk (levels e+1 .. e+RF)
    template <int RF, bool RZ>
    HD static void head(const uint32_t* v, uint32_t base, uint32_t r, const Mid2Args& p, const uint32_t* twF, uint32_t* B, const DstIO& D) {
        const int e = E_(p), kf = A_(p) + e;
        uint32_t w[1 << RF];
#pragma unroll
        for (int j = 0; j < (1 << RF); j++) w[j] = v[j];
#pragma unroll
        for (int q = 1; q <= RF; q++) {
            const int h = 1 << (q - 1);
#pragma unroll
            for (int j = 0; j < (1 << RF); j++) {
                if (j & h) continue;
                uint32_t x = w[j + h];
                if (!(RZ && (j & (h - 1)) == 0)) { const uint32_t ti = (1u << (e + q - 1)) + ((uint32_t)(j & (h - 1)) << e) + r; x = SHP ? fmul_pair(x, twF, ti) : fmul(x, twF[ti]); }
                const int hn = h << 1;  // outputs the next level multiplies stay in [0, 2p) (see round_t)
                const bool lz0 = q < RF && (j & hn) && !(RZ && (j & (hn - 1)) == 0);
                const bool lz1 = q < RF && ((j + h) & hn) && !(RZ && ((j + h) & (hn - 1)) == 0);
                w[j + h] = lz1 ? fsub_lazy(w[j], x) : fsub(w[j], x);
                w[j] = lz0 ? fadd_lazy(w[j], x) : fadd(w[j], x);
            }
        }
        if (p.nrf > 1) {
#pragma unroll
            for (int j = 0; j < (1 << RF); j++) B[padx<PSB>(((base + j) << e) + r)] = w[j];
        } else {
#pragma unroll
            for (int j = 0; j < (1 << RF); j++) D.st_i(((base + j) << e) + r, w[j]);
        }
    }
    template <int RF>
    HD static void tail(const KCtx& ux, const Mid2Args& p, uint32_t rb, bool from_global, const SrcIO& G, const uint32_t* A, const uint32_t* twI,
                        const uint32_t* twF, const uint32_t* Gs, uint32_t* B, uint32_t* dst_coef, const DstIO& D) {
        const uint32_t items = 1u << (A_(p) - RF);
        if (p.alias && (p.flags & MID_INTT) && (p.flags & MID_FWD)) {
            // A aliases B: every thread owns at most one item; all loads complete before any store
            uint32_t v[1 << RF];
            const uint32_t it = ux.tid;
            if (it < items) tail_load<RF>(v, it << RF, from_global, G, A);
            ux.sync();
            if (it < items) tail_compute_store<RF>(v, it << RF, p, rb, twI, twF, Gs, B, dst_coef, D);
        } else {
            for (uint32_t it = ux.tid; it < items; it += ux.nt) {
                uint32_t v[1 << RF];
                tail_load<RF>(v, it << RF, from_global, G, A);
                tail_compute_store<RF>(v, it << RF, p, rb, twI, twF, Gs, B, dst_coef, D);
            }
        }
    }

    HD static void run(const KCtx& cx, uint32_t* sm, Mid2Args p) {
        const Mid2Layout L(p);
        const bool intt = p.flags & MID_INTT, fwd = p.flags & MID_FWD, fly = p.flags & MID_GFLY;
        const uint32_t hi = cx.bx, rb = brev(hi, p.b);
        const uint32_t na = 1u << p.a, nf = 1u << (p.a + p.e);
        uint32_t *twI = sm + L.twI, *twF = sm + L.twF, *G3 = sm + L.G3, *Gs = sm + L.Gs, *G2 = sm + L.G2;
        if (intt) {
            // per-level blocks: index i in [2^(l-1), 2^l) holds w_{2^l}^-(i - 2^(l-1))
            for (uint32_t i = cx.tid; i < na; i += cx.nt) {
                uint32_t v = ONE;
                if (i >= 1) { const int l = 32 - clz32(i); v = tab_pow(p.rt.i_lo, p.rt.i_hi, (i - (1u << (l - 1))) << (24 - l)); }
                if (SHP) { twI[2 * i] = from_mont(v); twI[2 * i + 1] = shoup_quot_mont(v); } else twI[i] = v;
            }
            if (!fly) {
                if (p.b > 0) for (uint32_t i = cx.tid; i < na; i += cx.nt) G3[i] = g3(p, rb, i);
                for (uint32_t i = cx.tid; i < na; i += cx.nt) Gs[padi(i)] = gs(p, rb, i);
            }
        }
        if (fwd) {
            for (uint32_t i = cx.tid; i < nf; i += cx.nt) {
                uint32_t v = ONE;
                if (i >= 1) { const int l = 32 - clz32(i); v = tab_pow(p.rt.f_lo, p.rt.f_hi, (i - (1u << (l - 1))) << (24 - l)); }
                if (SHP) { twF[2 * i] = from_mont(v); twF[2 * i + 1] = shoup_quot_mont(v); } else twF[i] = v;
            }
            if (!fly && p.b > 0) for (uint32_t i = cx.tid; i < nf; i += cx.nt) {
                const uint32_t g = g2(p, rb, i);
                if (SHP) { G2[2 * i] = from_mont(g); G2[2 * i + 1] = shoup_quot_mont(g); } else G2[i] = g;
            }
        }
        cx.sync();
        const int gmode = p.b == 0 ? 0 : (fly ? 2 : 1);
        const int RF = intt ? p.Ri[p.nri - 1] : p.Rf[0];
        const uint32_t col_end = (cx.by + 1) * p.cols_per_block < p.ncols ? (cx.by + 1) * p.cols_per_block : p.ncols;
        for (int u = unit_first(cx, p.ut); u < p.units; u += unit_step(cx, p.ut)) {
            const KCtx ux = unit_ctx(cx, u, p.ut);
            uint32_t* ubase = sm + L.unit0 + (uint32_t)u * L.unit_words;
            uint32_t* B = ubase + L.B;
            uint32_t* A = (fwd && p.alias) ? B : ubase + L.A;
#if MID_R5 && defined(__CUDA_ARCH__)
            uint32_t pf[32];
            if constexpr (MA == 10 && ME == 2) {
                const uint32_t col0 = cx.by * p.cols_per_block + (uint32_t)u;
                if (intt && col0 < col_end) {
                    const uint32_t* src0 = p.in + (uint64_t)col0 * p.in_stride + ((uint64_t)hi << p.a);
#pragma unroll
                    for (int j = 0; j < 32; j++) pf[j] = src0[(uint32_t)ux.tid + 32u * j];
                }
            }
#endif
            for (uint32_t col = cx.by * p.cols_per_block + (uint32_t)u; col < col_end; col += (uint32_t)p.units) {
                const SrcIO G{p.in + (uint64_t)col * p.in_stride + ((uint64_t)hi << p.a), G3, p.rt.i_lo, p.rt.i_hi, rb, 24 - p.n, intt ? gmode : 0};
                const DstIO D{p.out + (uint64_t)col * p.out_stride + ((uint64_t)hi << (p.a + p.e)), G2, p.rt.f_lo, p.rt.f_hi, rb, 24 - p.n - p.e, gmode};
                uint32_t* dst_coef = p.out + (uint64_t)col * p.out_stride + ((uint64_t)hi << p.a);
                const SmemIO SA{A}, SB{B};
                if constexpr (MA == 10 && ME == 2) {  // host dispatches this instantiation for the fused iNTT+LDE and the expand-only LDE
                    // main-group schedule, all sizes compile-time: DIF 3+3 (+4 in registers), DIT (4 in registers +) 3+3
                    const SmemIOT<PSB> SB6{B};
#if MID_R5
                    if (intt) {
                        const RoundGeom<5, 10, 0, 5, false> g1(10, 0, 5);
#ifdef __CUDA_ARCH__
                        // one item per lane: raw values of THIS column were fetched during the previous column (pf), the next
                        // column's are requested before the butterflies start
                        uint32_t v[32];
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] = gmode == 1 ? fmul(pf[j], G3[(uint32_t)ux.tid + 32u * j]) : (gmode == 2 ? fmul(pf[j], g3(p, rb, (uint32_t)ux.tid + 32u * j)) : pf[j]);
                        if (col + (uint32_t)p.units < col_end) {
                            const uint32_t* nsrc = p.in + (uint64_t)(col + (uint32_t)p.units) * p.in_stride + ((uint64_t)hi << p.a);
#pragma unroll
                            for (int j = 0; j < 32; j++) pf[j] = nsrc[(uint32_t)ux.tid + 32u * j];
                        }
                        round_compute_store_item<5, true, 10, 0, 5, true, false, true>(g1, twI, (uint32_t)ux.tid, v, SA);
#else
                        round_t<5, true, 10, 0, 5, true, false, true>(ux, twI, 10, 0, 5, G, SA);
#endif
                        ux.sync();
                    }
                    tail<5>(ux, p, rb, !intt, G, A, twI, twF, Gs, B, dst_coef, D); ux.sync();  // expand-only: coefficients straight from global
                    round_t<5, false, 12, 0, 7, true, false, true>(ux, twF, 12, 0, 7, SB6, D);
#else
                    if (intt) {
                        round_t<3, true, 10, 0, 7, true, false, true>(ux, twI, 10, 0, 7, G, SA); ux.sync();
                        round_t<3, true, 10, 0, 4, true, true, true>(ux, twI, 10, 0, 4, SA, SA); ux.sync();
                    }
                    tail<4>(ux, p, rb, !intt, G, A, twI, twF, Gs, B, dst_coef, D); ux.sync();  // expand-only: coefficients straight from global
                    round_t<3, false, 12, 0, 6, true, false, true>(ux, twF, 12, 0, 6, SB6, SB6); ux.sync();
                    round_t<3, false, 12, 0, 9, true, false, true>(ux, twF, 12, 0, 9, SB6, D);
#endif
                } else {
                bool from_global = true;
                if (intt) {
                    int done = 0;
                    for (int r = 0; r + 1 < p.nri; r++) {
                        const int R = p.Ri[r], l0 = p.a - done - R;
                        if (r == 0) round_io_dyn<true, true>(ux, twI, p.a, 0, l0, R, G, SA);
                        else round_io_dyn<true, true>(ux, twI, p.a, 0, l0, R, SA, SA);
                        ux.sync();
                        done += R;
                        from_global = false;
                    }
                }
                switch (RF) {
                    case 1: tail<1>(ux, p, rb, from_global, G, A, twI, twF, Gs, B, dst_coef, D); break;
                    case 2: tail<2>(ux, p, rb, from_global, G, A, twI, twF, Gs, B, dst_coef, D); break;
                    case 3: tail<3>(ux, p, rb, from_global, G, A, twI, twF, Gs, B, dst_coef, D); break;
                    default: tail<4>(ux, p, rb, from_global, G, A, twI, twF, Gs, B, dst_coef, D); break;
                }
                if (fwd && p.nrf > 1) {
                    ux.sync();
                    int done = RF;
                    for (int r = 1; r < p.nrf; r++) {
                        const int R = p.Rf[r], l0 = p.e + done;
                        const bool last = r == p.nrf - 1;
                        if (last) round_io_dyn<false, true>(ux, twF, p.a + p.e, 0, l0, R, SB, D);
                        else round_io_dyn<false, true>(ux, twF, p.a + p.e, 0, l0, R, SB, SB);
                        if (!last) ux.sync();
                        done += R;
                    }
                }
                }
                ux.sync();  // unit buffers are reused by the next column
            }
        }
    }
};

}  // namespace hf
#include "ntt_tma.cuh"  // strided stages of the 2^10-row plans through TMA tensor maps (device builds only)
#include "ntt_mid.cuh"  // fused middle stage of the main-group LDE, one warp per (column, chunk) (device builds only)
namespace hf {

// ---- host-side planning / launching ---------------------------------------------------------------
struct NttPlan { int a, b, c; };

static inline int ilog2(uint64_t x) { int k = 0; while ((1ull << k) < x) k++; return k; }

// n = log2(coefficients per column); e = expand bits of the forward part (0 when only the inverse runs).
static inline NttPlan ntt_plan(int n, int e) {
    NttPlan pl;
    int a = n <= 10 ? n : (n - 10 > 10 ? n - 10 : 10);
    // po2 21 / 22 LDE: keep the 2^10 chunk so that the fused middle stage is the compile-time main-group kernel (tables in
    // shared memory, Shoup pairs); the strided stages then cover 2^11 / 2^12 rows with 32- / 16-byte row segments.  Measured
    // at po2 = 22, 192 columns: middle 38.8 -> ~10 ms against a few ms more in the strided stages.
    if (e == 2 && n > 20 && n <= 22) a = 10;
    if (const char* env = std::getenv("HFB200_NTT_A")) { int v = std::atoi(env); if (v >= 1 && v <= n) a = v; }
    if (a + e > 14) a = 14 - e;
    if (n - a > 12) a = n - 12;
    pl.a = a;
    pl.b = n - a;
    pl.c = pl.b == 0 ? 0 : (14 - pl.b < 5 ? 14 - pl.b : 5);
    // 2^10 .. 2^12 rows: 2^15-word tile on ONE 512-thread CTA per SM (128- / 64- / 32-byte row segments) instead of two
    // 2^14-word tiles: po2 = 20 NTT stage 7.19 -> 6.85 ms, po2 = 21 17.1 -> 14.1 ms, po2 = 22 38.4 -> 29.9 ms
    if (pl.b >= 10 && pl.b <= 12) pl.c = 15 - pl.b;
    if (const char* env = std::getenv("HFB200_NTT_C")) { int v = std::atoi(env); if (v >= 0 && v <= pl.c) pl.c = v; }
    if (pl.c > pl.a) pl.c = pl.a;
    if (pl.c < 0) pl.c = 0;
    return pl;
}

struct Ntt {
    Dev* dev = nullptr;
    RootTables rt{};
    uint32_t* tab_mem = nullptr;  // 8 x 4096 words
    int mid_threads = 256, str_threads = 256;

    void init(Dev* d) {
        dev = d;
        std::vector<uint32_t> h(8 * 4096);
        auto rou = [](int k) { uint32_t g = to_mont(137); for (int i = k; i < 27; i++) g = fmul(g, g); return g; };
        const uint32_t w = rou(24), wi = finv(w);
        const uint32_t bases[4] = {w, wi, THREE, INV3};
        for (int t = 0; t < 4; t++) {
            uint32_t* lo = &h[(2 * t) * 4096];
            uint32_t* hi = &h[(2 * t + 1) * 4096];
            uint32_t cur = ONE;
            for (int i = 0; i < 4096; i++) { lo[i] = cur; cur = fmul(cur, bases[t]); }
            const uint32_t step = cur;  // base^4096
            cur = ONE;
            for (int i = 0; i < 4096; i++) { hi[i] = cur; cur = fmul(cur, step); }
        }
        tab_mem = (uint32_t*)dev->alloc(h.size() * 4);
        dev->h2d(tab_mem, h.data(), h.size() * 4);
        dev->sync();
        rt.f_lo = tab_mem; rt.f_hi = tab_mem + 4096;
        rt.i_lo = tab_mem + 2 * 4096; rt.i_hi = tab_mem + 3 * 4096;
        rt.p3_lo = tab_mem + 4 * 4096; rt.p3_hi = tab_mem + 5 * 4096;
        rt.ip3_lo = tab_mem + 6 * 4096; rt.ip3_hi = tab_mem + 7 * 4096;
        if (const char* env = std::getenv("HFB200_NTT_THREADS")) { int v = std::atoi(env); if (v >= 32 && v <= 1024) mid_threads = str_threads = v; }
    }
    void destroy() { if (dev) dev->free(tab_mem); tab_mem = nullptr; }

    static int pick_rmax(int k, int c, int threads) {
        int r = k + c - ilog2((uint64_t)threads);
        return r < 2 ? 2 : (r > 5 ? 5 : r);
    }

    // rounds of <= 4 levels, as even as possible
    static int split_rounds(int L, int* R) {
        if (L <= 0) return 0;
        const int nr = (L + 3) / 4;
        for (int r = 0; r < nr; r++) R[r] = L / nr + (r < L % nr ? 1 : 0);
        return nr;
    }

    void strided(const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t ncols, int a, int b, int c, bool inv) {
        if (b == 0) { if (in != out) throw Err("ntt: strided pass with b = 0 must be in place"); return; }
#ifndef HFB200_EMU
        if (strided_tma(dev, rt, in, in_stride, out, out_stride, ncols, a, b, inv)) return;  // 2^10 rows: TMA-pipelined kernel (ntt_tma.cuh)
#endif
        Str2Args p{};
        p.in = in; p.out = out; p.in_stride = in_stride; p.out_stride = out_stride; p.ncols = ncols;
        p.a = a; p.b = b; p.c = c; p.inv = inv ? 1 : 0; p.rt = rt;
        p.nr = split_rounds(b, p.R);
        const size_t smem = (size_t)(((padded_words(1u << (b + c)) + 1u) & ~1u) + (1u << b)) * 4;  // tile + twiddle (w, w') pairs
        const uint64_t tiles = (uint64_t)ncols << (a - c);
        unsigned per_sm = (unsigned)((220 * 1024) / (smem + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
        if (const char* env = std::getenv("HFB200_NTT_PER_SM")) { int v = std::atoi(env); if (v >= 1 && v <= 8) per_sm = (unsigned)v; }
        uint64_t grid = (uint64_t)dev->sm_count * per_sm;
        if (grid > tiles) grid = tiles;
        const int key = b * 10000 + c * 100 + a;
        static const int big_threads = std::getenv("HFB200_STR_T") ? std::atoi(std::getenv("HFB200_STR_T")) : 1024;
        switch (key) {
            case 100310: dev->launch<StridedKernel2<10, 3, 10>, 256, 4>((unsigned)grid, 1, 256, smem, p); break;
            case 100312: dev->launch<StridedKernel2<10, 3, 12>, 256, 4>((unsigned)grid, 1, 256, smem, p); break;
            case 100410: dev->launch<StridedKernel2<10, 4, 10>, 256, STR_MINB>((unsigned)grid, 1, 256, smem, p); break;
            case 100412: dev->launch<StridedKernel2<10, 4, 12>, 256, STR_MINB>((unsigned)grid, 1, 256, smem, p); break;
            case 100411: dev->launch<StridedKernel2<10, 4, 11>, 256, STR_MINB>((unsigned)grid, 1, 256, smem, p); break;
            case 100413: dev->launch<StridedKernel2<10, 4, 13>, 256, STR_MINB>((unsigned)grid, 1, 256, smem, p); break;
            case 100414: dev->launch<StridedKernel2<10, 4, 14>, 256, STR_MINB>((unsigned)grid, 1, 256, smem, p); break;
            case 110310: dev->launch<StridedKernel2<11, 3, 10>, 256, STR_MINB>((unsigned)grid, 1, 256, smem, p); break;
            case 110312: dev->launch<StridedKernel2<11, 3, 12>, 256, STR_MINB>((unsigned)grid, 1, 256, smem, p); break;
            // 2^15-word tiles: one 1024-thread CTA per SM (512 threads: 6.85 ms, 1024: 6.58 ms for the po2 = 20 NTT stage)
#if STR_R5
            case 100510: dev->launch<StridedKernel2<10, 5, 10>, 512, 1>((unsigned)grid, 1, 512, smem, p); break;
            case 100512: dev->launch<StridedKernel2<10, 5, 12>, 512, 1>((unsigned)grid, 1, 512, smem, p); break;
#else
            case 100510: dev->launch<StridedKernel2<10, 5, 10>, 1024, 1>((unsigned)grid, 1, big_threads, smem, p); break;
            case 100512: dev->launch<StridedKernel2<10, 5, 12>, 1024, 1>((unsigned)grid, 1, big_threads, smem, p); break;
#endif
            case 110410: dev->launch<StridedKernel2<11, 4, 10>, 1024, 1>((unsigned)grid, 1, big_threads, smem, p); break;
            case 110412: dev->launch<StridedKernel2<11, 4, 12>, 1024, 1>((unsigned)grid, 1, big_threads, smem, p); break;
            case 120310: dev->launch<StridedKernel2<12, 3, 10>, 1024, 1>((unsigned)grid, 1, big_threads, smem, p); break;
            case 120312: dev->launch<StridedKernel2<12, 3, 12>, 1024, 1>((unsigned)grid, 1, big_threads, smem, p); break;
            case 120210: dev->launch<StridedKernel2<12, 2, 10>, 256, STR_MINB>((unsigned)grid, 1, 256, smem, p); break;
            case 120212: dev->launch<StridedKernel2<12, 2, 12>, 256, STR_MINB>((unsigned)grid, 1, 256, smem, p); break;
            case 90510: dev->launch<StridedKernel2<9, 5, 10>, 256, 2>((unsigned)grid, 1, 256, smem, p); break;
            case 90512: dev->launch<StridedKernel2<9, 5, 12>, 256, 2>((unsigned)grid, 1, 256, smem, p); break;
            case 70510: dev->launch<StridedKernel2<7, 5, 10>, 256, 2>((unsigned)grid, 1, 256, smem, p); break;
            case 70512: dev->launch<StridedKernel2<7, 5, 12>, 256, 2>((unsigned)grid, 1, 256, smem, p); break;
            case 80510: dev->launch<StridedKernel2<8, 5, 10>, 256, 2>((unsigned)grid, 1, 256, smem, p); break;
            case 80512: dev->launch<StridedKernel2<8, 5, 12>, 256, 2>((unsigned)grid, 1, 256, smem, p); break;
            case 60510: dev->launch<StridedKernel2<6, 5, 10>, 256, 2>((unsigned)grid, 1, 256, smem, p); break;
            case 60512: dev->launch<StridedKernel2<6, 5, 12>, 256, 2>((unsigned)grid, 1, 256, smem, p); break;
            default: dev->launch<StridedKernel2<-1, -1, -1>, 256, 2>((unsigned)grid, 1, 256, smem, p); break;
        }
    }

    void middle(const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t ncols, int n, int a, int e, uint32_t flags) {
        if (a < 5) { middle_small(in, in_stride, out, out_stride, ncols, n, a, e, flags); return; }
#ifndef HFB200_EMU
        if (mid_warp(dev, rt, in, in_stride, out, out_stride, ncols, n, a, e, flags)) return;  // fused iNTT + x4 LDE of 2^10 chunks (ntt_mid.cuh)
#endif
        Mid2Args p{};
        p.in = in; p.out = out; p.in_stride = in_stride; p.out_stride = out_stride; p.ncols = ncols;
        p.n = n; p.a = a; p.b = n - a; p.e = e; p.flags = flags;
        if (a + e > 12) p.flags |= MID_GFLY;
        p.n_inv = finv(to_mont((uint32_t)((1ull << n) % P)));
        p.rt = rt;
        const bool intt = flags & MID_INTT, fwd = flags & MID_FWD;
        const int RF = a < 4 ? a : 4;
        if (intt) { p.nri = split_rounds(a - RF, p.Ri); p.Ri[p.nri++] = RF; }
        if (fwd) { p.Rf[0] = RF; p.nrf = 1 + split_rounds(a - RF, p.Rf + 1); }
        // A may alias B when every thread of a unit owns at most one register-tail item
#ifdef HFB200_EMU
        p.alias = 0;
#else
        p.ut = a <= 10 ? MID_UT : (a == 11 ? 128 : 256);
        p.alias = (intt && fwd && (1 << (a - RF)) <= p.ut) ? 1 : 0;
#if MID_R5
        if (a == 10 && e == 2 && fwd && !(p.flags & MID_GFLY)) { p.ut = 32; p.alias = intt ? 1 : 0; }  // one warp per (column, chunk), 32 values per lane
#endif
#endif
#ifdef HFB200_EMU
        p.ut = MID_UT;
#endif
        // the compile-time kernel serves the fused iNTT+LDE and the expand-only LDE (check group, FRI rounds) of 2^10 chunks
        const bool main_cfg = a == 10 && e == 2 && fwd && !(p.flags & MID_GFLY);
        p.shp = main_cfg ? 1 : 0;
        // main group: ONE 512-thread CTA per SM (8 units sharing one set of Shoup-pair tables, 215 KB); otherwise two
        // 256-thread CTAs per SM for chunks up to 2^10
        const size_t budget = main_cfg ? 226 * 1024 : (a <= 10 ? 112 * 1024 : 200 * 1024);
        int units = (main_cfg ? 512 : 256) / p.ut;
#if MID_R5 && !defined(HFB200_EMU)
        if (main_cfg) units = MID_R5_UNITS;
#endif
        for (;; units >>= 1) {
            p.units = units;
            if ((size_t)Mid2Layout(p).total * 4 <= budget || units == 1) break;
        }
        const uint64_t chunks = 1ull << p.b;
        uint32_t groups = (uint32_t)((2ull * dev->sm_count + chunks - 1) / chunks);
        if (groups < 1) groups = 1;
        uint32_t max_groups = (ncols + (uint32_t)p.units - 1) / (uint32_t)p.units;
        if (groups > max_groups) groups = max_groups;
        p.cols_per_block = (ncols + groups - 1) / groups;
        groups = (ncols + p.cols_per_block - 1) / p.cols_per_block;
#if MID_R5
        if (main_cfg) dev->launch<MiddleKernel2<10, 2>, 32 * MID_R5_UNITS, 1>((unsigned)chunks, groups, p.units * p.ut, (size_t)Mid2Layout(p).total * 4, p);
#else
        if (main_cfg) dev->launch<MiddleKernel2<10, 2>, 512, 1>((unsigned)chunks, groups, p.units * p.ut, (size_t)Mid2Layout(p).total * 4, p);
#endif
        else dev->launch<MiddleKernel2<-1, -1>, 256, 2>((unsigned)chunks, groups, p.units * p.ut, (size_t)Mid2Layout(p).total * 4, p);
    }

    void middle_small(const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t ncols, int n, int a, int e, uint32_t flags) {
        MidArgs p{};
        p.in = in; p.out = out; p.in_stride = in_stride; p.out_stride = out_stride; p.ncols = ncols;
        p.n = n; p.a = a; p.b = n - a; p.e = e; p.flags = flags;
        if (a + e > 12) p.flags |= MID_GFLY;
        const int threads = 64;
        p.rmax_inv = pick_rmax(a, 0, threads); if (p.rmax_inv > a) p.rmax_inv = a > 0 ? a : 1;
        p.rmax_fwd = pick_rmax(a + e, 0, threads); if (p.rmax_fwd > a) p.rmax_fwd = a > 0 ? a : 1;
        if (p.rmax_inv > 4) p.rmax_inv = 4;
        if (p.rmax_fwd > 4) p.rmax_fwd = 4;
        p.n_inv = finv(to_mont((uint32_t)((1ull << n) % P)));
        p.rt = rt;
        const uint64_t chunks = 1ull << p.b;
        uint32_t groups = (uint32_t)((4ull * dev->sm_count + chunks - 1) / chunks);
        if (groups < 1) groups = 1;
        if (groups > ncols) groups = ncols;
        p.cols_per_block = (ncols + groups - 1) / groups;
        groups = (ncols + p.cols_per_block - 1) / p.cols_per_block;
        const MidLayout L(p);
        dev->launch<MiddleKernel, 64, 1>((unsigned)chunks, groups, threads, (size_t)L.total * 4, p);
    }

    // Hal::batch_interpolate_ntt (+ zk_shift): natural evaluations -> bit-reversed coefficients.
    void interpolate(const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t ncols, int n, bool shift) {
        const NttPlan pl = ntt_plan(n, 0);
        const uint32_t* mid_in = in; uint64_t mid_stride = in_stride;
        if (pl.b > 0) { strided(in, in_stride, out, out_stride, ncols, pl.a, pl.b, pl.c, true); mid_in = out; mid_stride = out_stride; }
        middle(mid_in, mid_stride, out, out_stride, ncols, n, pl.a, 0, MID_INTT | (shift ? MID_SHIFT : 0));
    }
    // Hal::batch_expand_into_evaluate_ntt: bit-reversed coefficients -> natural evaluations on the 2^e larger domain.
    void expand_evaluate(const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t ncols, int n, int e) {
        const NttPlan pl = ntt_plan(n, e);
        middle(in, in_stride, out, out_stride, ncols, n, pl.a, e, MID_FWD);
        if (pl.b > 0) strided(out, out_stride, out, out_stride, ncols, pl.a + e, pl.b, pl.c, false);
    }
    // Fused commit_group path: trace evaluations -> x4 LDE, coefficients never leave the SM.
    // `scratch` holds ncols * 2^n words (unused when the transform fits one chunk).
    void lde(const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t* scratch, uint32_t ncols, int n) {
        const NttPlan pl = ntt_plan(n, 2);
        const uint32_t* mid_in = in; uint64_t mid_stride = in_stride;
        if (pl.b > 0) { strided(in, in_stride, scratch, 1ull << n, ncols, pl.a, pl.b, pl.c, true); mid_in = scratch; mid_stride = 1ull << n; }
        middle(mid_in, mid_stride, out, out_stride, ncols, n, pl.a, 2, MID_INTT | MID_SHIFT | MID_FWD);
        if (pl.b > 0) strided(out, out_stride, out, out_stride, ncols, pl.a + 2, pl.b, pl.c, false);
    }
};

}  // namespace hf
