// TEST INFRASTRUCTURE (CPU oracle): the blinding-noise expansion, restated on the CPU.
// ChaCha20 block function exactly as RFC 8439 section 2.3 states it (state = 4 constants, 8 key words, block counter,
// 3 nonce words; 10 double rounds; add the input state), pinned by the RFC's own test vector (section 2.3.2) in
// tests/test_oracle_properties.py.  Element shape follows upstream's `Elem::random` (risc0-core 3.0.1, SURVEY.md
// Appendix A.1): six u32 draws folded mod p.  Deterministic test mode only: key = (seed, domain tag).
#pragma once
#include <cstdint>
#include "fp.h"

namespace orc {

static inline uint32_t rol(uint32_t v, int n) { return (v << n) | (v >> (32 - n)); }
static inline void quarter_round(uint32_t* st, int a, int b, int c, int d) {
    st[a] += st[b]; st[d] ^= st[a]; st[d] = rol(st[d], 16);
    st[c] += st[d]; st[b] ^= st[c]; st[b] = rol(st[b], 12);
    st[a] += st[b]; st[d] ^= st[a]; st[d] = rol(st[d], 8);
    st[c] += st[d]; st[b] ^= st[c]; st[b] = rol(st[b], 7);
}
static inline void chacha20_block(const uint32_t key[8], uint32_t counter, const uint32_t nonce[3], uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) in[4 + i] = key[i];
    in[12] = counter;
    for (int i = 0; i < 3; i++) in[13 + i] = nonce[i];
    uint32_t st[16];
    for (int i = 0; i < 16; i++) st[i] = in[i];
    for (int r = 0; r < 10; r++) {
        quarter_round(st, 0, 4, 8, 12); quarter_round(st, 1, 5, 9, 13); quarter_round(st, 2, 6, 10, 14); quarter_round(st, 3, 7, 11, 15);
        quarter_round(st, 0, 5, 10, 15); quarter_round(st, 1, 6, 11, 12); quarter_round(st, 2, 7, 8, 13); quarter_round(st, 3, 4, 9, 14);
    }
    for (int i = 0; i < 16; i++) out[i] = st[i] + in[i];
}

struct BlindKey {
    uint32_t k[8];
    explicit BlindKey(uint64_t seed) {
        // deterministic mode of the product (csrc/blind.cuh blind_key_from_seed): seed || "hfb20deterministic seed"
        static const uint32_t tag[6] = {0x32626668u, 0x74656430u, 0x696d7265u, 0x7473696eu, 0x73206369u, 0x64656573u};
        k[0] = (uint32_t)seed; k[1] = (uint32_t)(seed >> 32);
        for (int i = 0; i < 6; i++) k[2 + i] = tag[i];
    }
};
// nonce = (group, column, "blnd"), counter = row
static inline Fp blind_value(const BlindKey& key, uint32_t group, uint32_t col, uint32_t row) {
    const uint32_t nonce[3] = {group, col, 0x646e6c62u};
    uint32_t blk[16];
    chacha20_block(key.k, row, nonce, blk);
    uint64_t v = 0;
    for (int i = 0; i < 6; i++) v = ((v << 32) + blk[i]) % P;
    return Fp::from_u32((uint32_t)v);
}

}  // namespace orc
