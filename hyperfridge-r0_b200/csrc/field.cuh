// BabyBear field arithmetic for sm_100a kernels and the host-side transcript.
// Replaces risc0-core 3.0.1 `field::baby_bear::{Elem, ExtElem}` (/root/reference/Cargo.lock:3135-3138,
// crate not vendored; semantics per SURVEY.md Appendix A.1).  u32 Montgomery residues, R = 2^32.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define HD __host__ __device__ __forceinline__

namespace hf {

static constexpr uint32_t P = 2013265921u;       // 15 * 2^27 + 1
static constexpr uint32_t P_INV = 0x88000001u;   // P * P_INV == 1 (mod 2^32)
static constexpr uint32_t R2 = 1172168163u;      // 2^64 mod P
static constexpr uint32_t ONE = 268435454u;      // 2^32 mod P  (Montgomery 1)

HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

HD uint32_t fadd(uint32_t a, uint32_t b) { uint32_t x = a + b; return umin32(x, x - P); }
HD uint32_t fsub(uint32_t a, uint32_t b) { uint32_t x = a - b; return umin32(x, x + P); }
// Uncorrected sum / difference of canonical residues, in [0, 2p) < 2^32: valid ONLY as the first operand of fmul /
// fmul_shoup (both accept any 32-bit value there), never as an operand of fadd / fsub.
HD uint32_t fadd_lazy(uint32_t a, uint32_t b) { return a + b; }
HD uint32_t fsub_lazy(uint32_t a, uint32_t b) { return a - b + P; }
HD uint32_t fneg(uint32_t a) { return a == 0 ? 0u : P - a; }
// Montgomery product, subtractive form: hi(a*b) - hi(m*P) with m = lo(a*b) * P^-1 lies in (-P, P).
HD uint32_t fmul(uint32_t a, uint32_t b) {
    uint64_t o = (uint64_t)a * b;
    uint32_t m = (uint32_t)o * P_INV;
#ifdef __CUDA_ARCH__
    uint32_t mp = __umulhi(m, P);
#else
    uint32_t mp = (uint32_t)(((uint64_t)m * P) >> 32);
#endif
    uint32_t r = (uint32_t)(o >> 32) - mp;
    return umin32(r, r + P);
}
HD uint32_t fsqr(uint32_t a) { return fmul(a, a); }
HD uint32_t to_mont(uint32_t x) { return fmul(x % P, R2); }
HD uint32_t from_mont(uint32_t x) { return fmul(x, 1u); }
HD uint32_t fpow(uint32_t b, uint64_t e) {
    uint32_t r = ONE;
    while (e) { if (e & 1) r = fmul(r, b); b = fmul(b, b); e >>= 1; }
    return r;
}
HD uint32_t finv(uint32_t a) { return fpow(a, P - 2); }

// x * d mod p for a constant d known as (d, d' = floor(d 2^32 / p)) in CANONICAL form -- Shoup's method: one high
// multiply + two low multiplies, no 64-bit product, no separate subtraction.  x may be in any residue representation
// (Montgomery here): the result is in the same representation.  Canonical output.
HD uint32_t fmul_shoup(uint32_t x, uint32_t d, uint32_t dp) {
#ifdef __CUDA_ARCH__
    const uint32_t q = __umulhi(x, dp);
#else
    const uint32_t q = (uint32_t)(((uint64_t)x * dp) >> 32);
#endif
    const uint32_t r = x * d - q * P;
    return umin32(r, r - P);
}
HD uint32_t shoup_quot(uint32_t d) { return (uint32_t)(((uint64_t)d << 32) / P); }
// The same quotient from the Montgomery form wm = d 2^32 mod p, without a division: d 2^32 = q p + wm exactly, so
// q = (d 2^32 - wm) / p = -wm p^-1 (mod 2^32).
HD uint32_t shoup_quot_mont(uint32_t wm) { return (0u - wm) * P_INV; }
// (w, w') pair at word offset 2 t of an 8-byte aligned table
HD void ld_pair(const uint32_t* tw, uint32_t t, uint32_t& w, uint32_t& wp) {
    const uint2 v = *reinterpret_cast<const uint2*>(tw + 2 * t);
    w = v.x; wp = v.y;
}
HD uint32_t fmul_pair(uint32_t x, const uint32_t* tw, uint32_t t) { uint32_t w, wp; ld_pair(tw, t, w, wp); return fmul_shoup(x, w, wp); }

struct __align__(16) E4 { uint32_t c[4]; };

HD E4 e4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { E4 r; r.c[0] = a; r.c[1] = b; r.c[2] = c; r.c[3] = d; return r; }
HD E4 e4_zero() { return e4(0, 0, 0, 0); }
HD E4 e4_one() { return e4(ONE, 0, 0, 0); }
HD E4 e4_from(uint32_t a) { return e4(a, 0, 0, 0); }
HD E4 e4_add(const E4& a, const E4& b) { return e4(fadd(a.c[0], b.c[0]), fadd(a.c[1], b.c[1]), fadd(a.c[2], b.c[2]), fadd(a.c[3], b.c[3])); }
HD E4 e4_sub(const E4& a, const E4& b) { return e4(fsub(a.c[0], b.c[0]), fsub(a.c[1], b.c[1]), fsub(a.c[2], b.c[2]), fsub(a.c[3], b.c[3])); }
HD E4 e4_neg(const E4& a) { return e4(fneg(a.c[0]), fneg(a.c[1]), fneg(a.c[2]), fneg(a.c[3])); }
HD E4 e4_scale(const E4& a, uint32_t s) { return e4(fmul(a.c[0], s), fmul(a.c[1], s), fmul(a.c[2], s), fmul(a.c[3], s)); }
HD bool e4_eq(const E4& a, const E4& b) { return a.c[0] == b.c[0] && a.c[1] == b.c[1] && a.c[2] == b.c[2] && a.c[3] == b.c[3]; }
// x^4 = -11
static constexpr uint32_t NBETA = 1073741848u;  // Montgomery form of P - 11
static constexpr uint32_t BETA = 939524073u;    // Montgomery form of 11
static constexpr uint32_t THREE = 805306362u;   // Montgomery form of 3
static constexpr uint32_t INV3 = 760567125u;    // Montgomery form of 3^-1
HD E4 e4_mul(const E4& a, const E4& b) {
    uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    uint32_t b0 = b.c[0], b1 = b.c[1], b2 = b.c[2], b3 = b.c[3];
    uint32_t t0 = fadd(fadd(fmul(a1, b3), fmul(a2, b2)), fmul(a3, b1));
    uint32_t t1 = fadd(fmul(a2, b3), fmul(a3, b2));
    uint32_t t2 = fmul(a3, b3);
    E4 r;
    r.c[0] = fadd(fmul(a0, b0), fmul(NBETA, t0));
    r.c[1] = fadd(fadd(fmul(a0, b1), fmul(a1, b0)), fmul(NBETA, t1));
    r.c[2] = fadd(fadd(fadd(fmul(a0, b2), fmul(a1, b1)), fmul(a2, b0)), fmul(NBETA, t2));
    r.c[3] = fadd(fadd(fmul(a0, b3), fmul(a1, b2)), fadd(fmul(a2, b1), fmul(a3, b0)));
    return r;
}
HD E4 e4_sqr(const E4& a) { return e4_mul(a, a); }
HD E4 e4_pow(E4 b, uint64_t e) {
    E4 r = e4_one();
    while (e) { if (e & 1) r = e4_mul(r, b); b = e4_mul(b, b); e >>= 1; }
    return r;
}
// Inverse through the quadratic tower (y = x^2, y^2 = -11): a = A + xB, a^-1 = (A - xB)/(A^2 - yB^2).
HD E4 e4_inv(const E4& a) {
    auto mul11 = [](uint32_t x) { return fmul(x, BETA); };
    uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3];
    uint32_t a1a3 = fmul(a1, a3), a0a2 = fmul(a0, a2);
    uint32_t d0 = fadd(fsub(fsqr(a0), mul11(fsqr(a2))), mul11(fadd(a1a3, a1a3)));
    uint32_t d1 = fadd(fsub(fadd(a0a2, a0a2), fsqr(a1)), mul11(fsqr(a3)));
    uint32_t n = fadd(fsqr(d0), mul11(fsqr(d1)));
    uint32_t ni = finv(n);
    uint32_t e0 = fmul(d0, ni), e1 = fneg(fmul(d1, ni));
    E4 r;
    r.c[0] = fsub(fmul(a0, e0), mul11(fmul(a2, e1)));
    r.c[2] = fadd(fmul(a0, e1), fmul(a2, e0));
    r.c[1] = fneg(fsub(fmul(a1, e0), mul11(fmul(a3, e1))));
    r.c[3] = fneg(fadd(fmul(a1, e1), fmul(a3, e0)));
    return r;
}

// ---- lazy dot products ---------------------------------------------------------------------------------------------
// sum_i x_i * y_i of canonical residues accumulated as a 64-bit integer with ONE wide multiply-add per term and no
// per-term Montgomery reduction.  Invariant: hi < p.  A product is < p^2 < 2^62, so hi grows by at most 2^30 + 1 per
// term and the conditional subtraction of p * 2^32 (one min on the high word) restores hi < p.  redc() then maps the
// sum to sum * 2^-32 mod p: exactly what the chain fadd(acc, fmul(x_i, y_i)) computes, at a third of the multiplier work.
struct A64 { uint32_t lo, hi; };
HD A64 a64_zero() { A64 a; a.lo = 0; a.hi = 0; return a; }
HD void mac(A64& a, uint32_t x, uint32_t y) {
    const uint64_t t = (uint64_t)x * y + (((uint64_t)a.hi << 32) | a.lo);
    const uint32_t h = (uint32_t)(t >> 32);
    a.lo = (uint32_t)t;
    a.hi = umin32(h, h - P);
}
HD uint32_t redc(const A64& a) {
    const uint32_t m = a.lo * P_INV;
#ifdef __CUDA_ARCH__
    const uint32_t mp = __umulhi(m, P);
#else
    const uint32_t mp = (uint32_t)(((uint64_t)m * P) >> 32);
#endif
    const uint32_t r = a.hi - mp;
    return umin32(r, r + P);
}
struct E4A { A64 c[4]; };
HD E4A e4a_zero() { E4A a; a.c[0] = a.c[1] = a.c[2] = a.c[3] = a64_zero(); return a; }
HD void e4a_mac(E4A& a, const E4& w, uint32_t t) { mac(a.c[0], w.c[0], t); mac(a.c[1], w.c[1], t); mac(a.c[2], w.c[2], t); mac(a.c[3], w.c[3], t); }
HD E4 e4a_redc(const E4A& a) { return e4(redc(a.c[0]), redc(a.c[1]), redc(a.c[2]), redc(a.c[3])); }

HD int clz32(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __clz((int)x);
#else
    return x ? __builtin_clz(x) : 32;
#endif
}
HD uint32_t brev(uint32_t x, unsigned bits) {
#ifdef __CUDA_ARCH__
    return bits == 0 ? 0u : (__brev(x) >> (32 - bits));
#else
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

}  // namespace hf
