// ORACLE (test infrastructure only).  PARITY UNPINNED vs upstream for M_INT_DIAG (see
// tools/gen_poseidon2_consts.py); round constants are pinned to the published Grain-LFSR procedure.
// Restates risc0-zkp 3.0.4 `core::hash::poseidon2::{poseidon2_mix, unpadded_hash, hash_pair}` and
// `Poseidon2Rng` (/root/reference/Cargo.lock:3195-3198; not vendored).
#pragma once
#include <vector>
#include <cstring>
#include "fp.h"

namespace orc {

#include "poseidon2_consts.inc"

static constexpr int CELLS = 24, CELLS_RATE = 16, CELLS_OUT = 8;
static constexpr int ROUNDS_HALF_FULL = 4, ROUNDS_PARTIAL = 21;

struct Digest {
    uint32_t w[8];
    bool operator==(const Digest& o) const { return std::memcmp(w, o.w, 32) == 0; }
    bool operator!=(const Digest& o) const { return !(*this == o); }
};

struct P2Tables {
    Fp rc_first[96], rc_partial[21], rc_last[96], diag[24];
    P2Tables() {
        for (int i = 0; i < 96; i++) { rc_first[i] = Fp::from_u32(P2_RC_FULL_FIRST[i]); rc_last[i] = Fp::from_u32(P2_RC_FULL_LAST[i]); }
        for (int i = 0; i < 21; i++) rc_partial[i] = Fp::from_u32(P2_RC_PARTIAL[i]);
        for (int i = 0; i < 24; i++) diag[i] = Fp::from_u32(P2_M_INT_DIAG[i]);
    }
};
static inline const P2Tables& p2_tables() { static P2Tables t; return t; }

static inline Fp sbox7(Fp x) { Fp x2 = x * x; Fp x4 = x2 * x2; return x4 * x2 * x; }

// External linear layer: M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]] on each 4-chunk
// (Poseidon2 paper, appendix B add/double chain), then add the column sums across chunks.
static inline void m_ext(Fp* s) {
    for (int c = 0; c < CELLS; c += 4) {
        Fp a = s[c], b = s[c + 1], cc = s[c + 2], d = s[c + 3];
        Fp t0 = a + b, t1 = cc + d;
        Fp t2 = b + b + t1, t3 = d + d + t0;
        Fp t4 = t1 + t1; t4 = t4 + t4 + t3;
        Fp t5 = t0 + t0; t5 = t5 + t5 + t2;
        Fp t6 = t3 + t5, t7 = t2 + t4;
        s[c] = t6; s[c + 1] = t5; s[c + 2] = t7; s[c + 3] = t4;
    }
    Fp sum[4];
    for (int j = 0; j < 4; j++) { Fp t = fp_zero(); for (int c = 0; c < CELLS; c += 4) t += s[c + j]; sum[j] = t; }
    for (int i = 0; i < CELLS; i++) s[i] += sum[i & 3];
}

static inline void m_int(Fp* s) {
    const P2Tables& T = p2_tables();
    Fp sum = fp_zero();
    for (int i = 0; i < CELLS; i++) sum += s[i];
    for (int i = 0; i < CELLS; i++) s[i] = sum + T.diag[i] * s[i];
}

static inline void poseidon2_mix(Fp* s) {
    const P2Tables& T = p2_tables();
    m_ext(s);
    for (int r = 0; r < ROUNDS_HALF_FULL; r++) {
        for (int i = 0; i < CELLS; i++) s[i] = sbox7(s[i] + T.rc_first[r * CELLS + i]);
        m_ext(s);
    }
    for (int r = 0; r < ROUNDS_PARTIAL; r++) {
        s[0] = sbox7(s[0] + T.rc_partial[r]);
        m_int(s);
    }
    for (int r = 0; r < ROUNDS_HALF_FULL; r++) {
        for (int i = 0; i < CELLS; i++) s[i] = sbox7(s[i] + T.rc_last[r * CELLS + i]);
        m_ext(s);
    }
}

// Sponge without padding: overwrite-mode absorb of 16 cells per permutation; a trailing partial
// block (or an empty input) is zero-filled; digest = first 8 cells (raw Montgomery words).
static inline Digest unpadded_hash_stride(const Fp* in, size_t count, size_t stride) {
    Fp st[CELLS];
    size_t used = 0;
    bool any = false;
    for (size_t i = 0; i < count; i++) {
        st[used++] = in[i * stride];
        if (used == CELLS_RATE) { poseidon2_mix(st); used = 0; any = true; }
    }
    if (used != 0 || !any) {
        for (size_t k = used; k < (size_t)CELLS_RATE; k++) st[k] = fp_zero();
        poseidon2_mix(st);
    }
    Digest d;
    for (int i = 0; i < CELLS_OUT; i++) d.w[i] = st[i].v;
    return d;
}
static inline Digest hash_elems(const Fp* in, size_t count) { return unpadded_hash_stride(in, count, 1); }
static inline Digest hash_ext_elems(const Fp4* in, size_t count) { return hash_elems(reinterpret_cast<const Fp*>(in), count * 4); }

static inline Digest hash_pair(const Digest& a, const Digest& b) {
    Fp st[CELLS];
    for (int i = 0; i < 8; i++) { st[i] = Fp::raw(a.w[i]); st[8 + i] = Fp::raw(b.w[i]); }
    poseidon2_mix(st);
    Digest d;
    for (int i = 0; i < CELLS_OUT; i++) d.w[i] = st[i].v;
    return d;
}

struct Poseidon2Rng {
    Fp cells[CELLS];
    int pool_used = 0;
    // risc0-zkp 3.0.4 `core/hash/poseidon2/rng.rs`, `Poseidon2Rng::mix`: "if switching from squeezing, do a poseidon2 mix";
    // "Add in CELLS_OUT elements (also # of digest words)"; "Mix".  The first step (a permutation when elements were drawn
    // since the last mix) was missing from SURVEY.md Appendix A.3's restatement and was added in round 2 from recollection
    // of that source file.
    void mix(const Digest& d) {
        if (pool_used != 0) { poseidon2_mix(cells); pool_used = 0; }
        for (int i = 0; i < CELLS_OUT; i++) cells[i] += Fp::raw(d.w[i]);
        poseidon2_mix(cells);
        pool_used = 0;
    }
    Fp random_elem() {
        if (pool_used == CELLS_RATE) { poseidon2_mix(cells); pool_used = 0; }
        return cells[pool_used++];
    }
    Fp4 random_ext_elem() { Fp a = random_elem(), b = random_elem(), c = random_elem(), d = random_elem(); return Fp4(a, b, c, d); }
    uint32_t random_bits(unsigned bits) {
        uint32_t v = random_elem().as_u32();
        for (int i = 0; i < 3; i++) { uint32_t n = random_elem().as_u32(); if (v == 0) v = n; }
        return v & (uint32_t)(((uint64_t)1 << bits) - 1);
    }
};

}  // namespace orc
