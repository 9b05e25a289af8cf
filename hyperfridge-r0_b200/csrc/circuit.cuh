// Circuit plug-in "synth-rv32im-shape v1" on the device: witness stand-in, step_accum, eval_check.
// Plays the role of risc0-circuit-rv32im-sys 4.0.2's generated kernels (witgen steps, `step_accum`,
// `eval_check`/`poly_fp`; /root/reference/Cargo.lock:3121-3132) -- generated code that cannot be obtained
// here, so the circuit is a DECLARED synthetic one with the same shape (DESIGN.md section "circuit").
// Definition (independently restated by oracle/circuit.h):
//   groups: ACCUM=0 (w_accum = 4*n_chains), CODE=1 (w_code), DATA=2 (w_data = 2*n_free)
//   code:   c0 active (rows < N-1994), c1 first-row flag, c2 last-active flag, c3 cycle, c4.. pseudo-random
//   data:   free columns f_0..f_{F-1}; derived columns g_k = expr_k(f taps), k < F
//             k%4==0: A*B + C      1: A*B*C + P'     2: (A+X)*B*C*D     3: P'*B + C*D + X      (P' = back 1)
//   accum:  Fp4 chains: acc_r(i) = (first ? 1 : acc_r(i-1)) * (data[src_r](i) + mix_r)
//   constraints (all gated by `active`, degree <= 5), mixed with successive powers of poly_mix:
//     F derived, 4*n_chains accum components, 1 global tie  first*(f_0 - global_0)
#pragma once
#include "dev.cuh"

namespace hf {

static constexpr uint32_t ZK_CYCLES = 1994, N_GLOBAL = 32, CODE_FIXED = 4;
enum { GROUP_ACCUM = 0, GROUP_CODE = 1, GROUP_DATA = 2 };
static constexpr uint64_t CODE_SEED = 0x636F6465ull;

HD uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
HD uint32_t synth_value(uint64_t seed, uint32_t col, uint32_t row) {
    return to_mont((uint32_t)(splitmix64(seed ^ (((uint64_t)col << 32) | row)) % P));
}
HD uint32_t blind_value(uint64_t seed, uint32_t group, uint32_t col, uint32_t row) {
    const uint64_t key = splitmix64(seed ^ 0x6E6F697365ull) ^ ((uint64_t)group << 60) ^ ((uint64_t)col << 32) ^ row;
    uint64_t v = 0;
    for (int i = 0; i < 3; i++) {
        const uint64_t d = splitmix64(key + (uint64_t)i * 0xD1342543DE82EF95ull);
        v = ((v << 32) + (uint32_t)d) % P;
        v = ((v << 32) + (uint32_t)(d >> 32)) % P;
    }
    return to_mont((uint32_t)v);
}

struct CircuitDev {
    uint32_t w_code, w_data, w_accum, n_free, n_prev, n_chains;
    const uint16_t* picks;      // [n_free][6] : a, b, c, d, p, x   (device memory)
    const uint16_t* chain_src;  // [n_chains]
    HD uint32_t n_constraints() const { return n_free + 4 * n_chains + 1; }
};

HD uint32_t derived_expr(uint32_t k, uint32_t A, uint32_t B, uint32_t C, uint32_t D, uint32_t Pp, uint32_t X) {
    switch (k & 3u) {
        case 0: return fadd(fmul(A, B), C);
        case 1: return fadd(fmul(fmul(A, B), C), Pp);
        case 2: return fmul(fmul(fmul(fadd(A, X), B), C), D);
        default: return fadd(fadd(fmul(Pp, B), fmul(C, D)), X);
    }
}

// ---- witness stand-in ----------------------------------------------------------------------------
struct GenCodeKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, uint32_t* code, uint32_t w_code, uint32_t po2) {
        const uint64_t n = 1ull << po2, act = n - ZK_CYCLES;
        const uint64_t t = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (t >= (uint64_t)w_code * n) return;
        const uint32_t c = (uint32_t)(t >> po2), r = (uint32_t)(t & (n - 1));
        uint32_t v = 0;
        if (r < act) {
            if (c == 0) v = ONE;
            else if (c == 1) v = r == 0 ? ONE : 0u;
            else if (c == 2) v = r == act - 1 ? ONE : 0u;
            else if (c == 3) v = to_mont(r % P);
            else v = synth_value(CODE_SEED, c, r);
        }
        code[t] = v;
    }
};
struct GenFreeKernel {  // free columns (active rows) + blinding rows of every data column
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, uint32_t* data, CircuitDev cd, uint32_t po2, uint64_t trace_seed, uint64_t blind_seed, uint32_t global0) {
        const uint64_t n = 1ull << po2, act = n - ZK_CYCLES;
        const uint64_t t = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (t >= (uint64_t)cd.w_data * n) return;
        const uint32_t c = (uint32_t)(t >> po2), r = (uint32_t)(t & (n - 1));
        if (r >= act) data[t] = blind_value(blind_seed, GROUP_DATA, c, r);
        else if (c < cd.n_free) data[t] = (c == 0 && r == 0) ? global0 : synth_value(trace_seed, c, r);
    }
};
struct GenDerivedKernel {
    static constexpr bool kBarrier = false;
    HD static void run(const KCtx& cx, uint32_t*, uint32_t* data, const uint32_t* code, CircuitDev cd, uint32_t po2) {
        const uint64_t n = 1ull << po2, act = n - ZK_CYCLES;
        const uint64_t t = (uint64_t)cx.bx * cx.nt + cx.tid;
        if (t >= (uint64_t)cd.n_free * n) return;
        const uint32_t k = (uint32_t)(t >> po2), r = (uint32_t)(t & (n - 1));
        if (r >= act) return;
        const uint16_t* pk = cd.picks + 6 * k;
        const uint64_t rp = (r + n - 1) & (n - 1);
        data[(uint64_t)(cd.n_free + k) * n + r] = derived_expr(k, data[(uint64_t)pk[0] * n + r], data[(uint64_t)pk[1] * n + r], data[(uint64_t)pk[2] * n + r],
                                                               data[(uint64_t)pk[3] * n + r], data[(uint64_t)pk[4] * n + rp], code[(uint64_t)pk[5] * n + r]);
    }
};

// ---- step_accum: Fp4 running products (three-phase scan) -----------------------------------------------
// In-place inclusive scan (Fp4 product) of buf0[0..n) in shared memory; buf1 is scratch. Result in the returned buffer.
HD E4* block_scan_e4(const KCtx& cx, E4* buf0, E4* buf1, uint32_t n) {
    E4 *src = buf0, *dst = buf1;
    for (uint32_t d = 1; d < n; d <<= 1) {
        for (uint32_t i = cx.tid; i < n; i += cx.nt) dst[i] = i >= d ? e4_mul(src[i - d], src[i]) : src[i];
        cx.sync();
        E4* t = src; src = dst; dst = t;
    }
    return src;
}

static constexpr uint32_t ACC_ITEMS = 256, ACC_PER = 8, ACC_RPB = ACC_ITEMS * ACC_PER;  // rows per block

// mode 0: write the block's total product to partial[chain * nblk + blk]
// mode 1: read the exclusive block offset from partial[...] and write the running products (+ blinding rows)
struct AccumKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t* sm, uint32_t* accum, const uint32_t* data, const uint32_t* mix, E4* partial, CircuitDev cd, uint32_t po2, uint64_t blind_seed, int mode) {
        const uint64_t n = 1ull << po2, act = n - ZK_CYCLES;
        const uint32_t chain = cx.by, blk = cx.bx, nblk = cx.gx;
        const uint32_t* src = data + (uint64_t)cd.chain_src[chain] * n;
        const E4 m = e4(mix[4 * chain], mix[4 * chain + 1], mix[4 * chain + 2], mix[4 * chain + 3]);
        E4* buf0 = reinterpret_cast<E4*>(sm);
        E4* buf1 = buf0 + ACC_ITEMS;
        const uint64_t row0 = (uint64_t)blk * ACC_RPB;
        for (uint32_t it = cx.tid; it < ACC_ITEMS; it += cx.nt) {
            E4 pr = e4_one();
            for (uint32_t j = 0; j < ACC_PER; j++) {
                const uint64_t r = row0 + (uint64_t)it * ACC_PER + j;
                if (r < act) { E4 t = m; t.c[0] = fadd(t.c[0], src[r]); pr = e4_mul(pr, t); }
            }
            buf0[it] = pr;
        }
        cx.sync();
        E4* sc = block_scan_e4(cx, buf0, buf1, ACC_ITEMS);
        if (mode == 0) {
            if (cx.tid == 0) partial[(uint64_t)chain * nblk + blk] = sc[ACC_ITEMS - 1];
            return;
        }
        const E4 off = partial[(uint64_t)chain * nblk + blk];
        for (uint32_t it = cx.tid; it < ACC_ITEMS; it += cx.nt) {
            E4 pr = it == 0 ? off : e4_mul(off, sc[it - 1]);
            for (uint32_t j = 0; j < ACC_PER; j++) {
                const uint64_t r = row0 + (uint64_t)it * ACC_PER + j;
                if (r >= n) break;
                if (r < act) {
                    E4 t = m; t.c[0] = fadd(t.c[0], src[r]); pr = e4_mul(pr, t);
                    for (int k = 0; k < 4; k++) accum[(uint64_t)(4 * chain + k) * n + r] = pr.c[k];
                } else {
                    for (int k = 0; k < 4; k++) accum[(uint64_t)(4 * chain + k) * n + r] = blind_value(blind_seed, GROUP_ACCUM, 4 * chain + k, (uint32_t)r);
                }
            }
        }
    }
};
// Exclusive scan of the per-block products of one chain (grid.y = chain, one block): partial[i] <- prod_{j<i}.
struct AccumOffsetsKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t* sm, E4* partial, uint32_t nblk) {
        E4* buf0 = reinterpret_cast<E4*>(sm);
        E4* buf1 = buf0 + nblk;
        E4* p = partial + (uint64_t)cx.by * nblk;
        for (uint32_t i = cx.tid; i < nblk; i += cx.nt) buf0[i] = p[i];
        cx.sync();
        E4* sc = block_scan_e4(cx, buf0, buf1, nblk);
        for (uint32_t i = cx.tid; i < nblk; i += cx.nt) p[i] = i == 0 ? e4_one() : sc[i - 1];
    }
};

// ---- eval_check (CircuitHal::eval_check): constraint polynomial / vanishing polynomial on the LDE domain ----
struct EvalCheckArgs {
    const uint32_t *ev_accum, *ev_code, *ev_data;  // [w][domain]
    uint32_t* check;                               // [4][domain]
    const E4* mixpow;                              // poly_mix^j, j < n_constraints (device)
    const uint32_t* mix;                           // accum mix elems (device) [4*n_chains]
    uint32_t global0;
    uint32_t yinv[4];                              // 1/((3 w_4N^i)^N - 1) depends on i mod 4 only
    uint32_t po2;
    uint32_t rows_per_block;
    CircuitDev cd;
};
// One thread per LDE row; every operand is a coalesced 128-byte line across the warp (served by L2 after the first
// touch).  A shared-memory row-tile variant (stage all 272 columns of 64 rows, two threads per row) was measured at
// 20.2 ms against 3.8 ms for this kernel: 78 KB of tile per 128 threads leaves 8 warps per SM (profiles/README.md).
struct EvalCheckKernel {
    static constexpr bool kBarrier = true;
    HD static void run(const KCtx& cx, uint32_t* sm, EvalCheckArgs p) {
        const CircuitDev& cd = p.cd;
        const uint32_t nc = cd.n_constraints();
        E4* mp = reinterpret_cast<E4*>(sm);
        uint16_t* spk = reinterpret_cast<uint16_t*>(mp + nc);  // column picks staged in shared memory: the operand addresses
        for (uint32_t i = cx.tid; i < nc; i += cx.nt) mp[i] = p.mixpow[i];       // no longer wait on a global load
        for (uint32_t i = cx.tid; i < 6 * cd.n_free; i += cx.nt) spk[i] = cd.picks[i];
        cx.sync();
        const uint64_t domain = 4ull << p.po2, dmask = domain - 1;
        for (uint32_t rr = cx.tid; rr < p.rows_per_block; rr += cx.nt) {
            const uint64_t i = (uint64_t)cx.bx * p.rows_per_block + rr;
            if (i >= domain) break;
            const uint64_t ib = (i + domain - 4) & dmask;  // back 1 on the x4 domain
            const uint32_t active = p.ev_code[i], first = p.ev_code[domain + i];
            E4 tot = e4_zero();
            uint32_t j = 0;
#pragma unroll 4
            for (uint32_t k = 0; k < cd.n_free; k++, j++) {
                const uint16_t* pk = spk + 6 * k;
                const uint32_t e = derived_expr(k, p.ev_data[(uint64_t)pk[0] * domain + i], p.ev_data[(uint64_t)pk[1] * domain + i], p.ev_data[(uint64_t)pk[2] * domain + i],
                                                p.ev_data[(uint64_t)pk[3] * domain + i], p.ev_data[(uint64_t)pk[4] * domain + ib], p.ev_code[(uint64_t)pk[5] * domain + i]);
                const uint32_t cv = fmul(active, fsub(p.ev_data[(uint64_t)(cd.n_free + k) * domain + i], e));
                tot = e4_add(tot, e4_scale(mp[j], cv));
            }
            const uint32_t nf = fsub(ONE, first);
            for (uint32_t r = 0; r < cd.n_chains; r++) {
                E4 acc, s, t;
                for (int k = 0; k < 4; k++) {
                    acc.c[k] = p.ev_accum[(uint64_t)(4 * r + k) * domain + i];
                    s.c[k] = fmul(nf, p.ev_accum[(uint64_t)(4 * r + k) * domain + ib]);
                    t.c[k] = p.mix[4 * r + k];
                }
                s.c[0] = fadd(s.c[0], first);
                t.c[0] = fadd(t.c[0], p.ev_data[(uint64_t)cd.chain_src[r] * domain + i]);
                const E4 pr = e4_mul(s, t);
                for (int k = 0; k < 4; k++, j++) tot = e4_add(tot, e4_scale(mp[j], fmul(active, fsub(acc.c[k], pr.c[k]))));
            }
            tot = e4_add(tot, e4_scale(mp[j], fmul(first, fsub(p.ev_data[i], p.global0))));
            const uint32_t yi = p.yinv[i & 3];
            for (int k = 0; k < 4; k++) p.check[(uint64_t)k * domain + i] = fmul(tot.c[k], yi);
        }
    }
};

struct CircuitHost {
    CircuitDev cd{};
    uint32_t n_taps = 0;
    uint16_t* tab_mem = nullptr;
    std::vector<uint16_t> h_picks, h_chain_src;
    void init(Dev* dev, uint32_t wc, uint32_t wd, uint32_t wa) {
        if (wc < CODE_FIXED + 1 || wd < 8 || (wd & 3) || wa < 4 || (wa & 3) || wc > 4096 || wd > 4096 || wa > 4096) throw Err("circuit: unsupported widths");
        cd.w_code = wc; cd.w_data = wd; cd.w_accum = wa;
        cd.n_free = wd / 2; cd.n_prev = cd.n_free / 2; cd.n_chains = wa / 4;
        const uint32_t F = cd.n_free;
        h_picks.resize(6 * F);
        for (uint32_t k = 0; k < F; k++) {
            h_picks[6 * k + 0] = (uint16_t)(k % F);
            h_picks[6 * k + 1] = (uint16_t)((5 * k + 1) % F);
            h_picks[6 * k + 2] = (uint16_t)((11 * k + 2) % F);
            h_picks[6 * k + 3] = (uint16_t)((17 * k + 3) % F);
            h_picks[6 * k + 4] = (uint16_t)((7 * k + 1) % cd.n_prev);
            h_picks[6 * k + 5] = (uint16_t)(CODE_FIXED + k % (wc - CODE_FIXED));
        }
        h_chain_src.resize(cd.n_chains);
        for (uint32_t r = 0; r < cd.n_chains; r++) h_chain_src[r] = (uint16_t)((13 * r + 5) % wd);
        tab_mem = (uint16_t*)dev->alloc((h_picks.size() + h_chain_src.size()) * 2 + 16);
        dev->h2d(tab_mem, h_picks.data(), h_picks.size() * 2);
        dev->h2d(tab_mem + h_picks.size(), h_chain_src.data(), h_chain_src.size() * 2);
        dev->sync();
        cd.picks = tab_mem;
        cd.chain_src = tab_mem + h_picks.size();
        // taps: accum all {0,1}; code all {0}; data: columns < n_prev {0,1}, others {0}
        n_taps = 2 * wa + wc + wd + cd.n_prev;
    }
    void destroy(Dev* dev) { dev->free(tab_mem); tab_mem = nullptr; }
    uint32_t group_width(int g) const { return g == GROUP_ACCUM ? cd.w_accum : g == GROUP_CODE ? cd.w_code : cd.w_data; }
    uint32_t group_back1(int g) const { return g == GROUP_ACCUM ? cd.w_accum : g == GROUP_CODE ? 0u : cd.n_prev; }
    uint32_t n_regs() const { return cd.w_accum + cd.w_code + cd.w_data; }
    uint32_t n_mix() const { return 4 * cd.n_chains; }
};

}  // namespace hf
