// ORACLE (test infrastructure only).  PARITY UNPINNED (upstream not vendored).
// Restates risc0-zkp 3.0.4 `core::ntt::{interpolate_ntt, evaluate_ntt, bit_reverse, expand}` and the
// CPU HAL ops `batch_interpolate_ntt`, `zk_shift`, `batch_expand_into_evaluate_ntt`, `batch_bit_reverse`
// (/root/reference/Cargo.lock:3195-3198).  Plain radix-2 loops on purpose.
#pragma once
#include <vector>
#include "fp.h"

namespace orc {

template <typename T>
static inline void bit_reverse_inplace(T* io, size_t n) {
    unsigned bits = log2_exact(n);
    for (size_t i = 0; i < n; i++) {
        size_t j = bit_rev((uint32_t)i, bits);
        if (i < j) { T t = io[i]; io[i] = io[j]; io[j] = t; }
    }
}

// Inverse NTT: natural-order evaluations on <w_n> -> coefficients in BIT-REVERSED order, scaled 1/n.
// DIF (Gentleman-Sande) butterflies from the top level down: a' = a + b ; b' = (a - b) * w^-k.
template <typename T>
static inline void interpolate_ntt(T* io, size_t n) {
    unsigned lg = log2_exact(n);
    for (unsigned level = lg; level >= 1; level--) {
        size_t half = (size_t)1 << (level - 1);
        Fp step = rou_rev(level);
        for (size_t blk = 0; blk < n; blk += 2 * half) {
            Fp cur = fp_one();
            for (size_t k = 0; k < half; k++) {
                T a = io[blk + k], b = io[blk + k + half];
                io[blk + k] = a + b;
                io[blk + k + half] = (a - b) * cur;
                cur *= step;
            }
        }
    }
    Fp norm = Fp::from_u64(n).inv();
    for (size_t i = 0; i < n; i++) io[i] = io[i] * norm;
}

// Forward NTT: BIT-REVERSED coefficients -> natural-order evaluations on <w_n>.  DIT
// (Cooley-Tukey) butterflies from level 1 up, skipping the first `expand_bits` levels (valid
// when the input was produced by `expand`, i.e. each coefficient replicated 2^expand_bits times).
template <typename T>
static inline void evaluate_ntt(T* io, size_t n, unsigned expand_bits) {
    unsigned lg = log2_exact(n);
    for (unsigned level = expand_bits + 1; level <= lg; level++) {
        size_t half = (size_t)1 << (level - 1);
        Fp step = rou_fwd(level);
        for (size_t blk = 0; blk < n; blk += 2 * half) {
            Fp cur = fp_one();
            for (size_t k = 0; k < half; k++) {
                T a = io[blk + k], b = io[blk + k + half] * cur;
                io[blk + k] = a + b;
                io[blk + k + half] = a - b;
                cur *= step;
            }
        }
    }
}

// coeff at bit-reversed position i (true index j = bitrev(i)) *= 3^j   =>  f(x) -> f(3x)
static inline void zk_shift(Fp* io, size_t n) {
    unsigned bits = log2_exact(n);
    std::vector<Fp> pw(n);
    Fp three = Fp::from_u32(3), cur = fp_one();
    for (size_t j = 0; j < n; j++) { pw[j] = cur; cur *= three; }
    for (size_t i = 0; i < n; i++) io[i] *= pw[bit_rev((uint32_t)i, bits)];
}

// out[i] = in[i >> bits] (replicate) followed by evaluate_ntt skipping `bits` levels:
// low-degree extension of a bit-reversed coefficient vector onto the 2^bits-times larger domain.
static inline void expand_into_evaluate_ntt(Fp* out, const Fp* in, size_t n_in, unsigned bits) {
    size_t n_out = n_in << bits;
    for (size_t i = 0; i < n_out; i++) out[i] = in[i >> bits];
    evaluate_ntt(out, n_out, bits);
}

}  // namespace orc
