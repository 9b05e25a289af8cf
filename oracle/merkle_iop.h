// ORACLE (test infrastructure only).  PARITY UNPINNED (upstream not vendored).
// Restates risc0-zkp 3.0.4 `prove::merkle::MerkleTreeProver`, `verify::merkle::MerkleTreeVerifier`,
// `merkle::MerkleTreeParams`, `prove::write_iop::WriteIOP`, `verify::read_iop::ReadIOP`
// (/root/reference/Cargo.lock:3195-3198).
#pragma once
#include <vector>
#include <stdexcept>
#include <string>
#include "poseidon2.h"

namespace orc {

static constexpr size_t QUERIES = 50, INV_RATE = 4, FRI_FOLD = 16, FRI_MIN_DEGREE = 256;
static constexpr size_t ZK_CYCLES = 1994, EXT_SIZE = 4, CHECK_SIZE = INV_RATE * EXT_SIZE;

struct MerkleParams {
    size_t row_size, col_size, queries, layers, top_layer, top_size;
    MerkleParams(size_t rows, size_t cols, size_t q = QUERIES) : row_size(rows), col_size(cols), queries(q) {
        layers = log2_exact(rows);
        if (((size_t)1 << layers) != rows) throw std::runtime_error("merkle: row_size must be a power of two");
        top_layer = 0;
        for (size_t i = 1; i < layers; i++) if (((size_t)1 << i) > queries) break; else top_layer = i;
        top_size = (size_t)1 << top_layer;
    }
};

struct WriteIOP {
    std::vector<uint32_t> proof;
    Poseidon2Rng rng;
    void write_u32s(const uint32_t* p, size_t n) { proof.insert(proof.end(), p, p + n); }
    void write_elems(const Fp* p, size_t n) { for (size_t i = 0; i < n; i++) proof.push_back(p[i].v); }
    void write_ext_elems(const Fp4* p, size_t n) { write_elems(reinterpret_cast<const Fp*>(p), n * 4); }
    void write_digests(const Digest* d, size_t n) { for (size_t i = 0; i < n; i++) write_u32s(d[i].w, 8); }
    void commit(const Digest& d) { rng.mix(d); }
    Fp random_elem() { return rng.random_elem(); }
    Fp4 random_ext_elem() { return rng.random_ext_elem(); }
    uint32_t random_bits(unsigned b) { return rng.random_bits(b); }
};

struct ReadIOP {
    const uint32_t* p;
    size_t len, pos = 0;
    Poseidon2Rng rng;
    ReadIOP(const uint32_t* proof, size_t n) : p(proof), len(n) {}
    void need(size_t n) { if (pos + n > len) throw std::runtime_error("verify: seal truncated"); }
    void read_u32s(uint32_t* out, size_t n) { need(n); for (size_t i = 0; i < n; i++) out[i] = p[pos++]; }
    void read_elems(Fp* out, size_t n) {
        need(n);
        for (size_t i = 0; i < n; i++) { uint32_t w = p[pos++]; if (w >= P) throw std::runtime_error("verify: non-canonical field element"); out[i] = Fp::raw(w); }
    }
    void read_ext_elems(Fp4* out, size_t n) { read_elems(reinterpret_cast<Fp*>(out), n * 4); }
    void read_digests(Digest* d, size_t n) { for (size_t i = 0; i < n; i++) read_u32s(d[i].w, 8); }
    void commit(const Digest& d) { rng.mix(d); }
    Fp random_elem() { return rng.random_elem(); }
    Fp4 random_ext_elem() { return rng.random_ext_elem(); }
    uint32_t random_bits(unsigned b) { return rng.random_bits(b); }
    void verify_complete() { if (pos != len) throw std::runtime_error("verify: trailing words in seal"); }
};

// Heap-layout Merkle tree over the rows of a column-major matrix (matrix[c*rows + r]).
struct MerkleTreeProver {
    MerkleParams params;
    std::vector<Digest> nodes;  // [0, 2*rows); leaves at [rows, 2*rows); root = nodes[1]
    const Fp* matrix;
    MerkleTreeProver(const Fp* m, size_t rows, size_t cols) : params(rows, cols), nodes(2 * rows), matrix(m) {
        #pragma omp parallel for schedule(static)
        for (long r = 0; r < (long)rows; r++) nodes[rows + r] = unpadded_hash_stride(m + r, cols, rows);
        for (size_t level_size = rows / 2; level_size >= 1; level_size /= 2) {
            #pragma omp parallel for schedule(static) if (level_size > 1024)
            for (long i = 0; i < (long)level_size; i++) nodes[level_size + i] = hash_pair(nodes[2 * (level_size + i)], nodes[2 * (level_size + i) + 1]);
        }
    }
    const Digest& root() const { return nodes[1]; }
    void commit(WriteIOP& iop) const {
        iop.write_digests(&nodes[params.top_size], params.top_size);
        iop.commit(root());
    }
    // Opens row `idx`: the row's values then the sibling path up to (not including) the top layer.
    void prove(WriteIOP& iop, size_t idx) const {
        if (idx >= params.row_size) throw std::runtime_error("merkle prove: index out of range");
        for (size_t c = 0; c < params.col_size; c++) iop.proof.push_back(matrix[c * params.row_size + idx].v);
        size_t i = idx + params.row_size;
        while (i >= 2 * params.top_size) { iop.write_digests(&nodes[i ^ 1], 1); i >>= 1; }
    }
};

struct MerkleTreeVerifier {
    MerkleParams params;
    std::vector<Digest> top;  // [0, 2*top_size); top[1] = root
    MerkleTreeVerifier(ReadIOP& iop, size_t rows, size_t cols) : params(rows, cols), top(2 * params.top_size) {
        iop.read_digests(&top[params.top_size], params.top_size);
        for (size_t i = params.top_size - 1; i >= 1; i--) top[i] = hash_pair(top[2 * i], top[2 * i + 1]);
        iop.commit(top[1]);
    }
    const Digest& root() const { return top[1]; }
    std::vector<Fp> verify(ReadIOP& iop, size_t idx) const {
        if (idx >= params.row_size) throw std::runtime_error("merkle verify: index out of range");
        std::vector<Fp> row(params.col_size);
        iop.read_elems(row.data(), params.col_size);
        Digest cur = hash_elems(row.data(), row.size());
        size_t i = idx + params.row_size;
        while (i >= 2 * params.top_size) {
            Digest other; iop.read_digests(&other, 1);
            cur = (i & 1) ? hash_pair(other, cur) : hash_pair(cur, other);
            i >>= 1;
        }
        if (top[i] != cur) throw std::runtime_error("verify: merkle path does not match the committed top layer");
        return row;
    }
};

}  // namespace orc
