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
        case 4: ntt_round<4, INV>(cx, s, tw, k, c, l0); break;
        default: ntt_round<5, INV>(cx, s, tw, k, c, l0); break;
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

struct StrArgs {
    const uint32_t* in;
    uint32_t* out;                   // may alias `in`
    uint64_t in_stride, out_stride;  // column strides (words)
    uint32_t ncols;
    int a, b, c;  // rows hi < 2^b at stride 2^a words; tile = 2^b x 2^c
    int inv, rmax;
    RootTables rt;
};

// grid.x = any (grid-stride over ncols * 2^(a-c) tiles).
struct StridedKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t* sm, StrArgs p) {
        uint32_t* s = sm;
        uint32_t* tw = sm + padded_words(1u << (p.b + p.c));
        const uint32_t half = p.b > 0 ? 1u << (p.b - 1) : 1u;
        for (uint32_t i = cx.tid; i < half; i += cx.nt)
            tw[i] = p.inv ? tab_pow(p.rt.i_lo, p.rt.i_hi, i << (24 - p.b)) : tab_pow(p.rt.f_lo, p.rt.f_hi, i << (24 - p.b));
        cx.sync();
        const uint32_t tiles_per_col = 1u << (p.a - p.c);
        const uint64_t total = (uint64_t)p.ncols * tiles_per_col;
        const uint32_t cmask = (1u << p.c) - 1u;
        for (uint64_t tile = cx.bx; tile < total; tile += cx.gx) {
            const uint32_t col = (uint32_t)(tile >> (p.a - p.c));
            const uint32_t lo0 = ((uint32_t)tile & (tiles_per_col - 1u)) << p.c;
            const uint32_t* src = p.in + (uint64_t)col * p.in_stride + lo0;
            uint32_t* dst = p.out + (uint64_t)col * p.out_stride + lo0;
            if (p.c >= 2) {
                const uint32_t n4 = 1u << (p.b + p.c - 2), c4 = p.c - 2, m4 = (1u << c4) - 1u;
                for (uint32_t i4 = cx.tid; i4 < n4; i4 += cx.nt) {
                    const uint32_t h = i4 >> c4, l = (i4 & m4) << 2;
                    const uint4 v = *reinterpret_cast<const uint4*>(src + ((uint64_t)h << p.a) + l);
                    const uint32_t o = (h << p.c) + l;
                    s[padi(o)] = v.x; s[padi(o + 1)] = v.y; s[padi(o + 2)] = v.z; s[padi(o + 3)] = v.w;
                }
            } else {
                for (uint32_t i = cx.tid; i < (1u << (p.b + p.c)); i += cx.nt) s[padi(i)] = src[((uint64_t)(i >> p.c) << p.a) + (i & cmask)];
            }
            cx.sync();
            if (p.inv) ntt_tile<true>(cx, s, tw, p.b, p.c, 0, p.rmax);
            else ntt_tile<false>(cx, s, tw, p.b, p.c, 0, p.rmax);
            if (p.c >= 2) {
                const uint32_t n4 = 1u << (p.b + p.c - 2), c4 = p.c - 2, m4 = (1u << c4) - 1u;
                for (uint32_t i4 = cx.tid; i4 < n4; i4 += cx.nt) {
                    const uint32_t h = i4 >> c4, l = (i4 & m4) << 2;
                    const uint32_t o = (h << p.c) + l;
                    uint4 v;
                    v.x = s[padi(o)]; v.y = s[padi(o + 1)]; v.z = s[padi(o + 2)]; v.w = s[padi(o + 3)];
                    *reinterpret_cast<uint4*>(dst + ((uint64_t)h << p.a) + l) = v;
                }
            } else {
                for (uint32_t i = cx.tid; i < (1u << (p.b + p.c)); i += cx.nt) dst[((uint64_t)(i >> p.c) << p.a) + (i & cmask)] = s[padi(i)];
            }
            cx.sync();
        }
    }
};

// ---- host-side planning / launching ---------------------------------------------------------------
struct NttPlan { int a, b, c; };

static inline int ilog2(uint64_t x) { int k = 0; while ((1ull << k) < x) k++; return k; }

// n = log2(coefficients per column); e = expand bits of the forward part (0 when only the inverse runs).
static inline NttPlan ntt_plan(int n, int e) {
    NttPlan pl;
    int a = n <= 10 ? n : (n - 10 > 10 ? n - 10 : 10);
    if (const char* env = std::getenv("HFB200_NTT_A")) { int v = std::atoi(env); if (v >= 1 && v <= n) a = v; }
    if (a + e > 14) a = 14 - e;
    if (n - a > 12) a = n - 12;
    pl.a = a;
    pl.b = n - a;
    pl.c = pl.b == 0 ? 0 : (14 - pl.b < 5 ? 14 - pl.b : 5);
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

    void strided(const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t ncols, int a, int b, int c, bool inv) {
        if (b == 0) { if (in != out) throw Err("ntt: strided pass with b = 0 must be in place"); return; }
        StrArgs p{in, out, in_stride, out_stride, ncols, a, b, c, inv ? 1 : 0, pick_rmax(b, c, str_threads), rt};
        if (p.rmax > b) p.rmax = b;
        const size_t smem = (size_t)(padded_words(1u << (b + c)) + (1u << (b > 0 ? b - 1 : 0))) * 4;
        const uint64_t tiles = (uint64_t)ncols << (a - c);
        const unsigned ctas_per_sm = (unsigned)(200 * 1024 / (smem + 1024)) ? (unsigned)(200 * 1024 / (smem + 1024)) : 1u;
        uint64_t grid = (uint64_t)dev->sm_count * (ctas_per_sm > 8 ? 8 : ctas_per_sm);
        if (grid > tiles) grid = tiles;
        dev->launch<StridedKernel, 256, 1>((unsigned)grid, 1, str_threads > 256 ? 256 : str_threads, smem, p);
    }

    void middle(const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t ncols, int n, int a, int e, uint32_t flags) {
        MidArgs p{};
        p.in = in; p.out = out; p.in_stride = in_stride; p.out_stride = out_stride; p.ncols = ncols;
        p.n = n; p.a = a; p.b = n - a; p.e = e; p.flags = flags;
        if (a + e > 12) p.flags |= MID_GFLY;
        const int threads = mid_threads > 256 ? 256 : mid_threads;
        p.rmax_inv = pick_rmax(a, 0, threads); if (p.rmax_inv > a) p.rmax_inv = a > 0 ? a : 1;
        p.rmax_fwd = pick_rmax(a + e, 0, threads); if (p.rmax_fwd > a) p.rmax_fwd = a > 0 ? a : 1;
        p.n_inv = finv(to_mont((uint32_t)((1ull << n) % P)));
        p.rt = rt;
        // enough column groups to fill the machine ~4x, but keep per-CTA table setup amortised
        const uint64_t chunks = 1ull << p.b;
        uint32_t groups = (uint32_t)((4ull * dev->sm_count + chunks - 1) / chunks);
        if (groups < 1) groups = 1;
        if (groups > ncols) groups = ncols;
        p.cols_per_block = (ncols + groups - 1) / groups;
        groups = (ncols + p.cols_per_block - 1) / p.cols_per_block;
        const MidLayout L(p);
        dev->launch<MiddleKernel, 256, 1>((unsigned)chunks, groups, threads, (size_t)L.total * 4, p);
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
