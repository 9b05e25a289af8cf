// ORACLE (test infrastructure only -- never linked into the product library).
// PARITY UNPINNED: restates risc0-core 3.0.1 `field::baby_bear::{Elem, ExtElem}`
// (/root/reference/Cargo.lock:3135-3138; crate source not vendored under /root/reference, so no
// file:line exists; reference call site is /root/reference/host/src/main.rs:420-423).
// Algorithm source: SURVEY.md Appendix A.1 (constants re-derived numerically in tests/test_field.py).
//
// BabyBear: p = 15 * 2^27 + 1.  Elements are u32 Montgomery residues (R = 2^32).
// Fp4 = Fp[x] / (x^4 + 11).
#pragma once
#include <cstdint>
#include <cstddef>

namespace orc {

static constexpr uint32_t P = 2013265921u;        // 0x78000001
static constexpr uint32_t MONT_M = 0x88000001u;   // P * MONT_M == 1 (mod 2^32)
static constexpr uint32_t MONT_R2 = 1172168163u;  // 2^64 mod P
static constexpr uint32_t INVALID = 0xFFFFFFFFu;

struct Fp {
    uint32_t v;  // Montgomery form
    Fp() : v(0) {}
    static Fp raw(uint32_t r) { Fp x; x.v = r; return x; }
    static uint32_t mont_mul(uint32_t a, uint32_t b) {
        uint64_t o = (uint64_t)a * b;
        uint32_t low = 0u - (uint32_t)o;
        uint32_t red = MONT_M * low;
        o += (uint64_t)red * P;
        uint32_t r = (uint32_t)(o >> 32);
        return r >= P ? r - P : r;
    }
    static Fp from_u32(uint32_t x) { return raw(mont_mul(MONT_R2, x % P)); }
    static Fp from_u64(uint64_t x) { return from_u32((uint32_t)(x % P)); }
    uint32_t as_u32() const { return mont_mul(1, v); }
    Fp operator+(Fp o) const { uint32_t x = v + o.v; return raw(x >= P ? x - P : x); }
    Fp operator-(Fp o) const { uint32_t x = v - o.v; return raw(x > P ? x + P : x); }
    Fp operator*(Fp o) const { return raw(mont_mul(v, o.v)); }
    Fp operator-() const { return raw(v == 0 ? 0 : P - v); }
    Fp& operator+=(Fp o) { *this = *this + o; return *this; }
    Fp& operator-=(Fp o) { *this = *this - o; return *this; }
    Fp& operator*=(Fp o) { *this = *this * o; return *this; }
    bool operator==(Fp o) const { return v == o.v; }
    bool operator!=(Fp o) const { return v != o.v; }
    Fp pow(uint64_t e) const {
        Fp r = from_u32(1), b = *this;
        while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
        return r;
    }
    Fp inv() const { return pow(P - 2); }
};

static inline Fp fp_zero() { return Fp::raw(0); }
static inline Fp fp_one() { return Fp::from_u32(1); }

// 137 has multiplicative order 2^27; ROU_FWD[k] = 137^(2^(27-k)) has order 2^k.
static inline Fp rou_fwd(unsigned k) {
    Fp g = Fp::from_u32(137);
    for (unsigned i = k; i < 27; i++) g *= g;
    return g;
}
static inline Fp rou_rev(unsigned k) { return rou_fwd(k).inv(); }

struct Fp4 {
    Fp c[4];
    Fp4() {}
    Fp4(Fp a, Fp b, Fp cc, Fp d) { c[0] = a; c[1] = b; c[2] = cc; c[3] = d; }
    explicit Fp4(Fp a) { c[0] = a; }
    static Fp4 zero() { return Fp4(); }
    static Fp4 one() { return Fp4(fp_one()); }
    Fp4 operator+(const Fp4& o) const { return Fp4(c[0] + o.c[0], c[1] + o.c[1], c[2] + o.c[2], c[3] + o.c[3]); }
    Fp4 operator-(const Fp4& o) const { return Fp4(c[0] - o.c[0], c[1] - o.c[1], c[2] - o.c[2], c[3] - o.c[3]); }
    Fp4 operator-() const { return Fp4(-c[0], -c[1], -c[2], -c[3]); }
    Fp4 operator*(Fp s) const { return Fp4(c[0] * s, c[1] * s, c[2] * s, c[3] * s); }
    Fp4 operator*(const Fp4& o) const {
        const Fp nbeta = Fp::from_u32(P - 11);
        const Fp *a = c, *b = o.c;
        Fp4 r;
        r.c[0] = a[0] * b[0] + nbeta * (a[1] * b[3] + a[2] * b[2] + a[3] * b[1]);
        r.c[1] = a[0] * b[1] + a[1] * b[0] + nbeta * (a[2] * b[3] + a[3] * b[2]);
        r.c[2] = a[0] * b[2] + a[1] * b[1] + a[2] * b[0] + nbeta * (a[3] * b[3]);
        r.c[3] = a[0] * b[3] + a[1] * b[2] + a[2] * b[1] + a[3] * b[0];
        return r;
    }
    Fp4& operator+=(const Fp4& o) { *this = *this + o; return *this; }
    Fp4& operator-=(const Fp4& o) { *this = *this - o; return *this; }
    Fp4& operator*=(const Fp4& o) { *this = *this * o; return *this; }
    Fp4& operator*=(Fp s) { *this = *this * s; return *this; }
    bool operator==(const Fp4& o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2] && c[3] == o.c[3]; }
    bool operator!=(const Fp4& o) const { return !(*this == o); }
    Fp4 pow(uint64_t e) const {
        Fp4 r = one(), b = *this;
        while (e) { if (e & 1) r *= b; b *= b; e >>= 1; }
        return r;
    }
    // Inverse through the tower Fp4 = Fp2[x]/(x^2 - y), y^2 = -11:
    // a = A(y) + x*B(y) with A = a0 + a2*y, B = a1 + a3*y ;  a^-1 = (A - xB) / (A^2 - y B^2).
    Fp4 inv() const {
        const Fp beta = Fp::from_u32(11);
        Fp a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
        // (A^2 - y*B^2) in Fp2 with y^2 = -beta : D = d0 + d1*y
        Fp d0 = a0 * a0 - beta * (a2 * a2) + beta * ((a1 * a3) + (a1 * a3));  // a0^2 - b a2^2 - (-b)(2 a1 a3)
        Fp d1 = (a0 * a2) + (a0 * a2) - a1 * a1 + beta * (a3 * a3);
        // norm to Fp: d0^2 + beta d1^2
        Fp n = d0 * d0 + beta * (d1 * d1);
        Fp ni = n.inv();
        Fp e0 = d0 * ni, e1 = -(d1 * ni);  // D^-1 = (d0 - d1 y)/n
        // (A - xB) * (e0 + e1 y):  A*E = (a0 e0 - b a2 e1) + (a0 e1 + a2 e0) y ; B*E similarly
        Fp r0 = a0 * e0 - beta * (a2 * e1);
        Fp r2 = a0 * e1 + a2 * e0;
        Fp r1 = -(a1 * e0 - beta * (a3 * e1));
        Fp r3 = -(a1 * e1 + a3 * e0);
        return Fp4(r0, r1, r2, r3);
    }
};

static inline unsigned log2_exact(size_t n) { unsigned k = 0; while (((size_t)1 << k) < n) k++; return k; }
static inline uint32_t bit_rev(uint32_t x, unsigned bits) {
    uint32_t r = 0;
    for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

}  // namespace orc
