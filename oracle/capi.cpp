// ORACLE (test infrastructure only): C ABI over the CPU restatement so tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline leg can drive it through ctypes.  Nothing in the product links this.
#include "prover.h"
#include <cstring>
#include <omp.h>

using namespace orc;

static thread_local std::string g_err;
#define ORC_TRY try {
#define ORC_CATCH } catch (const std::exception& e) { g_err = e.what(); return -1; } return 0;

struct ProofHandle { SegmentProof proof; OracleTimes times; };

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }
int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_threads(int n) { omp_set_num_threads(n); }

// ---- blinding PRF (RFC 8439 block function; pinned by the RFC's test vector in tests/) ----
void orc_chacha20_block(const uint32_t* key8, uint32_t counter, const uint32_t* nonce3, uint32_t* out16) { chacha20_block(key8, counter, nonce3, out16); }
uint32_t orc_blind_value(uint64_t seed, uint32_t group, uint32_t col, uint32_t row) { return blind_value(BlindKey(seed), group, col, row).v; }

// ---- field ----
uint32_t orc_mont_mul(uint32_t a, uint32_t b) { return Fp::mont_mul(a, b); }
uint32_t orc_encode(uint32_t x) { return Fp::from_u32(x).v; }
uint32_t orc_decode(uint32_t x) { return Fp::raw(x).as_u32(); }
uint32_t orc_add(uint32_t a, uint32_t b) { return (Fp::raw(a) + Fp::raw(b)).v; }
uint32_t orc_sub(uint32_t a, uint32_t b) { return (Fp::raw(a) - Fp::raw(b)).v; }
uint32_t orc_inv(uint32_t a) { return Fp::raw(a).inv().v; }
uint32_t orc_rou_fwd(uint32_t k) { return rou_fwd(k).v; }
uint32_t orc_rou_rev(uint32_t k) { return rou_rev(k).v; }
void orc_fp4_mul(const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fp4 r = *reinterpret_cast<const Fp4*>(a) * *reinterpret_cast<const Fp4*>(b);
    std::memcpy(out, &r, 16);
}
void orc_fp4_inv(const uint32_t* a, uint32_t* out) { Fp4 r = reinterpret_cast<const Fp4*>(a)->inv(); std::memcpy(out, &r, 16); }

// ---- poseidon2 ----
void orc_poseidon2_mix(uint32_t* state24) { poseidon2_mix(reinterpret_cast<Fp*>(state24)); }
void orc_hash_elems(const uint32_t* in, size_t n, uint32_t* out8) { Digest d = hash_elems(reinterpret_cast<const Fp*>(in), n); std::memcpy(out8, d.w, 32); }
void orc_hash_pair(const uint32_t* a, const uint32_t* b, uint32_t* out8) {
    Digest d = hash_pair(*reinterpret_cast<const Digest*>(a), *reinterpret_cast<const Digest*>(b));
    std::memcpy(out8, d.w, 32);
}
// Mixes `n_digests` digests then draws `n_out` elems (raw Montgomery) -- exercises Poseidon2Rng.
void orc_rng_draw(const uint32_t* digests, size_t n_digests, uint32_t* out, size_t n_out) {
    Poseidon2Rng r;
    for (size_t i = 0; i < n_digests; i++) r.mix(*reinterpret_cast<const Digest*>(digests + 8 * i));
    for (size_t i = 0; i < n_out; i++) out[i] = r.random_elem().v;
}
// mix(d1), draw n_skip elements, mix(d2), draw n_out elements: the squeeze -> absorb switch of Poseidon2Rng::mix
void orc_rng_mix_draw_mix(const uint32_t* d1, size_t n_skip, const uint32_t* d2, uint32_t* out, size_t n_out) {
    Poseidon2Rng r;
    r.mix(*reinterpret_cast<const Digest*>(d1));
    for (size_t i = 0; i < n_skip; i++) (void)r.random_elem();
    r.mix(*reinterpret_cast<const Digest*>(d2));
    for (size_t i = 0; i < n_out; i++) out[i] = r.random_elem().v;
}
uint32_t orc_rng_bits(const uint32_t* digest, unsigned bits, size_t skip) {
    Poseidon2Rng r; r.mix(*reinterpret_cast<const Digest*>(digest));
    uint32_t v = 0;
    for (size_t i = 0; i <= skip; i++) v = r.random_bits(bits);
    return v;
}

// ---- NTT (column-major batches: io[c*n + i]) ----
void orc_interpolate_ntt(uint32_t* io, size_t count, size_t n) {
    #pragma omp parallel for
    for (long c = 0; c < (long)count; c++) interpolate_ntt(reinterpret_cast<Fp*>(io) + c * n, n);
}
void orc_evaluate_ntt(uint32_t* io, size_t count, size_t n, unsigned expand_bits) {
    #pragma omp parallel for
    for (long c = 0; c < (long)count; c++) evaluate_ntt(reinterpret_cast<Fp*>(io) + c * n, n, expand_bits);
}
void orc_zk_shift(uint32_t* io, size_t count, size_t n) {
    #pragma omp parallel for
    for (long c = 0; c < (long)count; c++) zk_shift(reinterpret_cast<Fp*>(io) + c * n, n);
}
void orc_expand_ntt(uint32_t* out, const uint32_t* in, size_t count, size_t n_in, unsigned bits) {
    #pragma omp parallel for
    for (long c = 0; c < (long)count; c++)
        expand_into_evaluate_ntt(reinterpret_cast<Fp*>(out) + (c * n_in << bits), reinterpret_cast<const Fp*>(in) + c * n_in, n_in, bits);
}
void orc_bit_reverse(uint32_t* io, size_t count, size_t n) {
    for (size_t c = 0; c < count; c++) bit_reverse_inplace(io + c * n, n);
}

// ---- Merkle ----
// nodes_out (optional): 2*rows digests in heap layout.
int orc_merkle(const uint32_t* matrix, size_t rows, size_t cols, uint32_t* root_out, uint32_t* nodes_out) {
    ORC_TRY
    MerkleTreeProver t(reinterpret_cast<const Fp*>(matrix), rows, cols);
    std::memcpy(root_out, t.root().w, 32);
    if (nodes_out) std::memcpy(nodes_out, t.nodes.data(), 2 * rows * 32);
    ORC_CATCH
}

// ---- circuit ----
int orc_circuit_info(uint32_t wc, uint32_t wd, uint32_t wa, uint32_t* n_taps, uint32_t* n_mix, uint32_t* n_constraints) {
    ORC_TRY
    Circuit c(wc, wd, wa);
    *n_taps = (uint32_t)c.taps.size(); *n_mix = c.n_mix(); *n_constraints = c.n_constraints();
    ORC_CATCH
}
int orc_gen_code(uint32_t wc, uint32_t wd, uint32_t wa, unsigned po2, uint32_t* code) {
    ORC_TRY Circuit c(wc, wd, wa); c.gen_code(reinterpret_cast<Fp*>(code), po2); ORC_CATCH
}
int orc_gen_globals(uint32_t wc, uint32_t wd, uint32_t wa, uint64_t seed, uint32_t* globals) {
    ORC_TRY Circuit c(wc, wd, wa); c.gen_globals(reinterpret_cast<Fp*>(globals), seed); ORC_CATCH
}
int orc_gen_data(uint32_t wc, uint32_t wd, uint32_t wa, unsigned po2, const uint32_t* code, const uint32_t* globals,
                 uint64_t trace_seed, uint64_t blind_seed, uint32_t* data) {
    ORC_TRY
    Circuit c(wc, wd, wa);
    c.gen_data(reinterpret_cast<Fp*>(data), reinterpret_cast<const Fp*>(code), reinterpret_cast<const Fp*>(globals), po2, trace_seed, blind_seed);
    ORC_CATCH
}
int orc_step_accum(uint32_t wc, uint32_t wd, uint32_t wa, unsigned po2, const uint32_t* data, const uint32_t* mix, uint64_t blind_seed, uint32_t* accum) {
    ORC_TRY
    Circuit c(wc, wd, wa);
    c.step_accum(reinterpret_cast<Fp*>(accum), reinterpret_cast<const Fp*>(data), reinterpret_cast<const Fp*>(mix), po2, blind_seed);
    ORC_CATCH
}
int orc_control_id(uint32_t wc, uint32_t wd, uint32_t wa, unsigned po2, uint32_t* root_out) {
    ORC_TRY Circuit c(wc, wd, wa); Digest d = control_id(c, po2); std::memcpy(root_out, d.w, 32); ORC_CATCH
}
size_t orc_seal_words_model(uint32_t wc, uint32_t wd, uint32_t wa, unsigned po2) {
    try { Circuit c(wc, wd, wa); return seal_words_model(c, po2); } catch (...) { return 0; }
}

// ---- handle-based circuit API (variants, user tap sets, constraint polynomial as PolyStep data) ----
void* orc_circuit_new(uint32_t wc, uint32_t wd, uint32_t wa, uint32_t variant) {
    try { return new Circuit(wc, wd, wa, variant); } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void orc_circuit_free(void* h) { delete static_cast<Circuit*>(h); }
// taps: n_taps triples (group, offset, back) sorted; steps: n_steps quads (op, a, b, c); ret: mix var holding the result
// info16: the circuit's 16-byte CIRCUIT_INFO (NULL = upstream's "RV32IM:v2_______")
int orc_circuit_set_ir(void* h, const uint32_t* taps, size_t n_taps, const uint32_t* steps, size_t n_steps, uint32_t ret, const uint8_t* info16) {
    ORC_TRY
    Circuit* c = static_cast<Circuit*>(h);
    if (taps) {
        std::vector<Tap> t(n_taps);
        for (size_t i = 0; i < n_taps; i++) t[i] = Tap{taps[3 * i], taps[3 * i + 1], taps[3 * i + 2], 0};
        c->set_taps(t);
    }
    std::vector<PolyStep> st(n_steps);
    for (size_t i = 0; i < n_steps; i++) st[i] = PolyStep{steps[4 * i], steps[4 * i + 1], steps[4 * i + 2], steps[4 * i + 3]};
    c->set_ir(st, ret, info16);
    ORC_CATCH
}
size_t orc_h_n_taps(void* h) { return static_cast<Circuit*>(h)->taps.size(); }
void orc_h_taps(void* h, uint32_t* out) {
    const Circuit* c = static_cast<Circuit*>(h);
    for (size_t i = 0; i < c->taps.size(); i++) { out[3 * i] = c->taps[i].group; out[3 * i + 1] = c->taps[i].offset; out[3 * i + 2] = c->taps[i].back; }
}
int orc_h_gen_data(void* h, unsigned po2, const uint32_t* code, const uint32_t* globals, uint64_t trace_seed, uint64_t blind_seed, uint32_t* data) {
    ORC_TRY
    static_cast<Circuit*>(h)->gen_data(reinterpret_cast<Fp*>(data), reinterpret_cast<const Fp*>(code), reinterpret_cast<const Fp*>(globals), po2, trace_seed, blind_seed);
    ORC_CATCH
}
int orc_h_control_id(void* h, unsigned po2, uint32_t* root_out) {
    ORC_TRY Digest d = control_id(*static_cast<Circuit*>(h), po2); std::memcpy(root_out, d.w, 32); ORC_CATCH
}
size_t orc_h_seal_words_model(void* h, unsigned po2) {
    try { return seal_words_model(*static_cast<Circuit*>(h), po2); } catch (...) { return 0; }
}
int orc_h_prove_segment(void* h, unsigned po2, const uint32_t* globals, const uint32_t* code, const uint32_t* data, uint64_t blind_seed, void** handle_out) {
    ORC_TRY
    ProofHandle* ph = new ProofHandle();
    try {
        ph->proof = prove_segment(*static_cast<Circuit*>(h), po2, reinterpret_cast<const Fp*>(globals), reinterpret_cast<const Fp*>(code),
                                  reinterpret_cast<const Fp*>(data), blind_seed, &ph->times);
    } catch (...) { delete ph; throw; }
    *handle_out = ph;
    ORC_CATCH
}
int orc_h_verify_segment(void* h, const uint32_t* seal, size_t seal_words, const uint32_t* code_root, unsigned* po2_out) {
    ORC_TRY
    verify_segment(*static_cast<Circuit*>(h), seal, seal_words, *reinterpret_cast<const Digest*>(code_root), po2_out);
    ORC_CATCH
}

// ---- prove / verify ----
int orc_prove_segment(uint32_t wc, uint32_t wd, uint32_t wa, unsigned po2, const uint32_t* globals, const uint32_t* code,
                      const uint32_t* data, uint64_t blind_seed, void** handle_out) {
    ORC_TRY
    Circuit c(wc, wd, wa);
    ProofHandle* h = new ProofHandle();
    try {
        h->proof = prove_segment(c, po2, reinterpret_cast<const Fp*>(globals), reinterpret_cast<const Fp*>(code),
                                 reinterpret_cast<const Fp*>(data), blind_seed, &h->times);
    } catch (...) { delete h; throw; }
    *handle_out = h;
    ORC_CATCH
}
size_t orc_proof_seal_words(void* h) { return static_cast<ProofHandle*>(h)->proof.seal.size(); }
void orc_proof_seal(void* h, uint32_t* out) { auto& s = static_cast<ProofHandle*>(h)->proof.seal; std::memcpy(out, s.data(), s.size() * 4); }
size_t orc_proof_n_checkpoints(void* h) { return static_cast<ProofHandle*>(h)->proof.cp.items.size(); }
size_t orc_proof_checkpoint(void* h, size_t i, char* name, size_t name_cap, uint32_t* words, size_t cap) {
    auto& it = static_cast<ProofHandle*>(h)->proof.cp.items[i];
    std::snprintf(name, name_cap, "%s", it.first.c_str());
    size_t n = it.second.size() < cap ? it.second.size() : cap;
    std::memcpy(words, it.second.data(), n * 4);
    return it.second.size();
}
void orc_proof_times(void* h, double* out6) {
    OracleTimes& t = static_cast<ProofHandle*>(h)->times;
    out6[0] = t.commit; out6[1] = t.accum; out6[2] = t.check; out6[3] = t.deep; out6[4] = t.fri; out6[5] = t.total;
}
void orc_proof_free(void* h) { delete static_cast<ProofHandle*>(h); }

int orc_verify_segment(uint32_t wc, uint32_t wd, uint32_t wa, const uint32_t* seal, size_t seal_words, const uint32_t* code_root, unsigned* po2_out) {
    ORC_TRY
    Circuit c(wc, wd, wa);
    verify_segment(c, seal, seal_words, *reinterpret_cast<const Digest*>(code_root), po2_out);
    ORC_CATCH
}

}  // extern "C"
