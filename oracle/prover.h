// ORACLE (test infrastructure only).  PARITY UNPINNED (upstream not vendored, no golden seal exists
// under /root/reference: both shipped receipts are dev-mode fakes, data/test/test.xml-Receipt-test.json:1).
// Restates, step by step and in upstream's order of operations, risc0-zkp 3.0.4
// `prove::Prover::{commit_group, finalize}`, `prove::poly_group::PolyGroup`, `prove::fri::fri_prove`
// and the rv32im `SegmentProver::prove` wrapper around them (/root/reference/Cargo.lock:3087-3223;
// entered from /root/reference/host/src/main.rs:423).
// Coefficient-space DEEP quotient with synthetic division, like upstream's CPU HAL -- deliberately a
// different algorithm from the CUDA product (which forms the quotient point-wise on the trace domain),
// so seal equality is a real cross-check.
// Pinned to upstream where that is possible without the sources: the Poseidon2 permutation reproduces upstream's known-answer
// vector (tests/test_poseidon2.py), the seal-length model the reference's five published seal sizes; whole seals stay unpinned.
#pragma once
#include <map>
#include <functional>
#include <string>
#include "ntt.h"
#include "circuit.h"

namespace orc {

struct Checkpoints {
    // named transcript values, in order of appearance (for bit-exact GPU-vs-oracle comparison)
    std::vector<std::pair<std::string, std::vector<uint32_t>>> items;
    void add(const std::string& name, const uint32_t* w, size_t n) { items.emplace_back(name, std::vector<uint32_t>(w, w + n)); }
    void add(const std::string& name, const Digest& d) { add(name, d.w, 8); }
    void add(const std::string& name, const Fp4& e) { uint32_t w[4] = {e.c[0].v, e.c[1].v, e.c[2].v, e.c[3].v}; add(name, w, 4); }
};

// A committed group of `count` polynomials of degree < size: coefficients (natural order after
// construction), their x4 low-degree extension on 3<w_4N>, and the Merkle tree over LDE rows.
struct PolyGroup {
    size_t count, size, domain;
    std::vector<Fp> coeffs;     // [count][size], natural order, of g(y) = f(3y)
    std::vector<Fp> evaluated;  // [count][domain]
    MerkleTreeProver* merkle = nullptr;
    // `bitrev_coeffs` = output of interpolate_ntt (+ zk_shift): bit-reversed order.
    PolyGroup(std::vector<Fp>&& bitrev_coeffs, size_t count_, size_t size_) : count(count_), size(size_), domain(size_ * INV_RATE), coeffs(std::move(bitrev_coeffs)), evaluated(count_ * size_ * INV_RATE) {
        #pragma omp parallel for schedule(dynamic)
        for (long c = 0; c < (long)count; c++) {
            expand_into_evaluate_ntt(&evaluated[c * domain], &coeffs[c * size], size, 2);
            bit_reverse_inplace(&coeffs[c * size], size);
        }
        merkle = new MerkleTreeProver(evaluated.data(), domain, count);
    }
    ~PolyGroup() { delete merkle; }
    PolyGroup(const PolyGroup&) = delete;
};

// Prover::commit_group's make_coeffs: iNTT each column then zk_shift.
static inline std::vector<Fp> make_coeffs(const Fp* cols, size_t count, size_t size, bool shift) {
    std::vector<Fp> c(cols, cols + count * size);
    #pragma omp parallel for schedule(dynamic)
    for (long i = 0; i < (long)count; i++) {
        interpolate_ntt(&c[i * size], size);
        if (shift) zk_shift(&c[i * size], size);
    }
    return c;
}

static inline Fp4 poly_eval(const Fp4* coeffs, size_t n, const Fp4& x) {
    Fp4 tot = Fp4::zero();
    for (size_t i = n; i-- > 0;) tot = tot * x + coeffs[i];
    return tot;
}
static inline Fp4 poly_eval_base(const Fp* coeffs, size_t n, const Fp4& x) {
    Fp4 tot = Fp4::zero();
    for (size_t i = n; i-- > 0;) tot = tot * x + Fp4(coeffs[i]);
    return tot;
}
// Lagrange interpolation of n points (n <= a handful) into coefficients.
static inline void poly_interpolate(Fp4* out, const Fp4* xs, const Fp4* ys, size_t n) {
    for (size_t i = 0; i < n; i++) out[i] = Fp4::zero();
    for (size_t i = 0; i < n; i++) {
        // basis_i(x) = prod_{j != i} (x - x_j) / (x_i - x_j)
        std::vector<Fp4> b(1, Fp4::one());
        Fp4 den = Fp4::one();
        for (size_t j = 0; j < n; j++) {
            if (j == i) continue;
            std::vector<Fp4> nb(b.size() + 1, Fp4::zero());
            for (size_t k = 0; k < b.size(); k++) { nb[k + 1] += b[k]; nb[k] -= b[k] * xs[j]; }
            b.swap(nb);
            den *= xs[i] - xs[j];
        }
        Fp4 s = ys[i] * den.inv();
        for (size_t k = 0; k < b.size(); k++) out[k] += b[k] * s;
    }
}
// In-place synthetic division of a natural-order polynomial by (x - z); returns the remainder.
static inline Fp4 poly_divide(Fp4* p, size_t n, const Fp4& z) {
    Fp4 cur = Fp4::zero();
    for (size_t i = n; i-- > 0;) {
        Fp4 next = z * cur + p[i];
        p[i] = cur;
        cur = next;
    }
    return cur;
}

struct FriRoundTree { std::vector<Fp> evaluated; MerkleTreeProver* merkle; size_t domain; };

struct SegmentProof {
    std::vector<uint32_t> seal;
    Checkpoints cp;
};

// Timings (seconds) of the oracle's stages, for the CPU-baseline report.
struct OracleTimes { double commit = 0, accum = 0, check = 0, deep = 0, fri = 0, total = 0; };

SegmentProof prove_segment(const Circuit& cir, unsigned po2, const Fp* globals, const Fp* code, const Fp* data,
                           uint64_t blind_seed, OracleTimes* times = nullptr);

// Throws std::runtime_error with a reason on rejection.  `code_root` is the control id of (circuit, po2).
void verify_segment(const Circuit& cir, const uint32_t* seal, size_t seal_words, const Digest& code_root, unsigned* po2_out = nullptr);

Digest control_id(const Circuit& cir, unsigned po2);
size_t seal_words_model(const Circuit& cir, unsigned po2);

}  // namespace orc
