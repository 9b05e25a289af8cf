// Zero-knowledge blinding noise: a 256-bit key expanded by the ChaCha20 block function (RFC 8439 section 2.3), counter
// based so that every (group, column, row) element is computed independently on the device.
//
// Upstream draws the blinding rows (the last ZK_CYCLES = 1994 rows of DATA and ACCUM) from a cryptographic RNG per proof
// (risc0-zkp `Elem::random(&mut rand::thread_rng())`).  Here:
//   * production (default): the key is 32 bytes of OS entropy (getrandom) drawn per segment; the caller's `blind_seed`
//     is only mixed in;
//   * deterministic mode (tests / bench, opt-in through hfb200_set_blinding): the key is derived from the 64-bit
//     `blind_seed` alone, so that seals are reproducible and comparable with the CPU oracle.  NOT zero-knowledge against
//     anyone who can guess the seed -- never the mode for real statements.
// Element shape follows `Elem::random`: six u32 draws folded mod p.
#pragma once
#include "dev.cuh"
#include <cerrno>
#include <sys/random.h>

namespace hf {

struct BlindKey { uint32_t k[8]; };

HD uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#define HF_QR(a, b, c, d) a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); a += b; d ^= a; d = rotl32(d, 8); c += d; b ^= c; b = rotl32(b, 7);
// out = first `n_out` (<= 16) words of the ChaCha20 block for (key, counter, nonce)
HD void chacha20_block(const BlindKey& key, uint32_t counter, uint32_t n0, uint32_t n1, uint32_t n2, uint32_t* out, int n_out) {
    uint32_t s[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u, key.k[0], key.k[1], key.k[2], key.k[3],
                      key.k[4], key.k[5], key.k[6], key.k[7], counter, n0, n1, n2};
    uint32_t x0 = s[0], x1 = s[1], x2 = s[2], x3 = s[3], x4 = s[4], x5 = s[5], x6 = s[6], x7 = s[7];
    uint32_t x8 = s[8], x9 = s[9], x10 = s[10], x11 = s[11], x12 = s[12], x13 = s[13], x14 = s[14], x15 = s[15];
    for (int i = 0; i < 10; i++) {
        HF_QR(x0, x4, x8, x12) HF_QR(x1, x5, x9, x13) HF_QR(x2, x6, x10, x14) HF_QR(x3, x7, x11, x15)
        HF_QR(x0, x5, x10, x15) HF_QR(x1, x6, x11, x12) HF_QR(x2, x7, x8, x13) HF_QR(x3, x4, x9, x14)
    }
    const uint32_t x[16] = {x0, x1, x2, x3, x4, x5, x6, x7, x8, x9, x10, x11, x12, x13, x14, x15};
    for (int i = 0; i < n_out; i++) out[i] = x[i] + s[i];
}
#undef HF_QR

// nonce = (group, column, "blnd"), block counter = row; the first six words of the block fold into one element
HD uint32_t blind_value(const BlindKey& key, uint32_t group, uint32_t col, uint32_t row) {
    uint32_t w[6];
    chacha20_block(key, row, group, col, 0x646e6c62u, w, 6);
    uint64_t v = 0;
    for (int i = 0; i < 6; i++) v = ((v << 32) + w[i]) % P;
    return to_mont((uint32_t)v);
}

enum { BLIND_OS_ENTROPY = 0, BLIND_DETERMINISTIC = 1 };

// deterministic mode: key = seed (64 bits) padded with a domain tag
static inline BlindKey blind_key_from_seed(uint64_t seed) {
    BlindKey k;
    k.k[0] = (uint32_t)seed; k.k[1] = (uint32_t)(seed >> 32);
    k.k[2] = 0x32626668u; k.k[3] = 0x74656430u; k.k[4] = 0x696d7265u; k.k[5] = 0x7473696eu; k.k[6] = 0x73206369u; k.k[7] = 0x64656573u;  // "hfb20det" "erminist" "ic seed"
    return k;
}
// production: 32 bytes from the OS (getrandom blocks until the pool is initialised, never returns weak bytes)
static inline BlindKey blind_key_from_os(uint64_t mix_in) {
    BlindKey k;
    uint8_t* p = reinterpret_cast<uint8_t*>(k.k);
    size_t got = 0;
    while (got < sizeof k.k) {
        const ssize_t r = getrandom(p + got, sizeof k.k - got, 0);
        if (r < 0) { if (errno == EINTR) continue; throw Err("getrandom failed: no entropy source for the zero-knowledge blinding"); }
        got += (size_t)r;
    }
    k.k[0] ^= (uint32_t)mix_in; k.k[1] ^= (uint32_t)(mix_in >> 32);
    return k;
}

}  // namespace hf
