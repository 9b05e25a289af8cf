// ORACLE (test infrastructure only).
// The real rv32im-v2 circuit (taps, constraint polynomial, witgen, step_accum) lives in
// risc0-circuit-rv32im 4.0.4 / risc0-circuit-rv32im-sys 4.0.2 (/root/reference/Cargo.lock:3087-3132),
// generated code that is not vendored and not reproducible here.  This header defines the DECLARED
// stand-in used by every parity / bench configuration: "synth-rv32im-shape v1" (SURVEY.md section 8d,
// config 2).  It has the rv32im SHAPE (three register groups, {0} and {0,1} tap sets, degree <= 5
// constraints gated by a code-group selector, Fp4 grand-product accumulators fed by post-commit mix
// randomness, ZK blinding rows) and plays the role of the `CircuitHal` plug-in:
//   poly_fp / poly_ext  <->  CircuitHal::eval_check's generated poly_fp and verify-side poly_ext
//   step_accum          <->  rv32im-sys step_accum
//   gen_code / gen_data <->  control columns / WitnessGenerator stand-in
// The CUDA product implements the same definition independently (hyperfridge-r0_b200/csrc/circuit.cuh).
#pragma once
#include <cstring>
#include <vector>
#include <algorithm>
#include <tuple>
#include "merkle_iop.h"
#include "blind.h"

namespace orc {

enum { GROUP_ACCUM = 0, GROUP_CODE = 1, GROUP_DATA = 2, NUM_GROUPS = 3 };

static inline uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline Fp synth_value(uint64_t seed, uint32_t col, uint32_t row) {
    return Fp::from_u64(splitmix64(seed ^ (((uint64_t)col << 32) | row)));
}
struct Tap { uint32_t group, offset, back, combo; };

// Constraint polynomial as data: upstream's verifier-side representation `risc0_zkp::adapter::PolyExtStepDef`
// (block of PolyExtStep { Const, Get, GetGlobal, Add, Sub, Mul, True, AndEqz, AndCond } + ret; recollection, crate not
// vendored).  fp vars and mix vars live in two separate SSA index spaces, every step pushes one value.
enum PolyOp : uint32_t { OP_CONST = 0, OP_GET = 1, OP_GET_GLOBAL = 2, OP_ADD = 3, OP_SUB = 4, OP_MUL = 5, OP_TRUE = 6, OP_AND_EQZ = 7, OP_AND_COND = 8 };
struct PolyStep { uint32_t op, a, b, c; };
struct Reg { uint32_t group, offset, combo, tap_begin, size; };

struct Circuit {
    uint32_t w_code, w_data, w_accum;  // widths
    uint32_t n_free, n_prev;           // data: free columns [0,n_free), of which [0,n_prev) also tapped at back 1
    uint32_t n_chains;                 // accum: Fp4 chains
    static constexpr uint32_t N_GLOBAL = 32;
    static constexpr uint32_t CODE_FIXED = 4;  // active, first, last, cycle
    uint64_t code_seed = 0x636F6465ull;
    uint32_t variant = 0;              // 1 = "synth-rv32im-shape v2": the k%4==1 constraints tap P at back 2 (tap sets {0},{0,1},{0,1,2})
    std::vector<PolyStep> ir;          // when non-empty, poly() interprets this instead of the built-in formula
    uint32_t ir_ret = 0;

    std::vector<Tap> taps;             // sorted (group, offset, back)
    std::vector<Reg> regs;             // sorted (group, offset)
    std::vector<std::vector<uint32_t>> combos;  // distinct back-sets: {0}, {0,1}
    std::vector<uint32_t> combo_begin;
    uint32_t tot_combo_backs = 0;
    uint32_t group_tap_begin[NUM_GROUPS + 1];

    Circuit(uint32_t wc, uint32_t wd, uint32_t wa, uint32_t variant_ = 0) : w_code(wc), w_data(wd), w_accum(wa), variant(variant_) {
        if (wc < CODE_FIXED + 1 || wd < 8 || (wd & 3) || wa < 4 || (wa & 3) || variant_ > 1) throw std::runtime_error("circuit: unsupported widths");
        n_free = wd / 2; n_prev = n_free / 2; n_chains = wa / 4;
        std::vector<Tap> t;
        for (uint32_t c = 0; c < w_accum; c++) { t.push_back(Tap{GROUP_ACCUM, c, 0, 0}); t.push_back(Tap{GROUP_ACCUM, c, 1, 0}); }
        for (uint32_t c = 0; c < w_code; c++) t.push_back(Tap{GROUP_CODE, c, 0, 0});
        for (uint32_t c = 0; c < w_data; c++) {
            t.push_back(Tap{GROUP_DATA, c, 0, 0});
            if (c < n_prev) { t.push_back(Tap{GROUP_DATA, c, 1, 0}); if (variant) t.push_back(Tap{GROUP_DATA, c, 2, 0}); }
        }
        set_taps(t);
    }
    // Rebuilds registers / combos from a tap list sorted by (group, offset, back) -- the shape of upstream's TapSet.
    void set_taps(const std::vector<Tap>& user) {
        taps = user; regs.clear(); combos.clear(); combo_begin.clear();
        for (size_t i = 1; i < taps.size(); i++) {
            const Tap &a = taps[i - 1], &b = taps[i];
            if (std::make_tuple(a.group, a.offset, a.back) >= std::make_tuple(b.group, b.offset, b.back)) throw std::runtime_error("circuit: taps must be strictly sorted by (group, offset, back)");
        }
        std::vector<std::vector<uint32_t>> backsets;
        for (size_t i = 0; i < taps.size();) {
            if (taps[i].group >= NUM_GROUPS || taps[i].offset >= group_width(taps[i].group)) throw std::runtime_error("circuit: tap out of range");
            size_t j = i; std::vector<uint32_t> b;
            while (j < taps.size() && taps[j].group == taps[i].group && taps[j].offset == taps[i].offset) b.push_back(taps[j++].back);
            regs.push_back(Reg{taps[i].group, taps[i].offset, 0, (uint32_t)i, (uint32_t)(j - i)});
            backsets.push_back(b);
            i = j;
        }
        combos = backsets;
        std::sort(combos.begin(), combos.end());
        combos.erase(std::unique(combos.begin(), combos.end()), combos.end());
        combo_begin.push_back(0);
        for (auto& c : combos) combo_begin.push_back(combo_begin.back() + (uint32_t)c.size());
        tot_combo_backs = combo_begin.back();
        for (size_t r = 0; r < regs.size(); r++) {
            regs[r].combo = (uint32_t)(std::lower_bound(combos.begin(), combos.end(), backsets[r]) - combos.begin());
            for (uint32_t k = 0; k < regs[r].size; k++) taps[regs[r].tap_begin + k].combo = regs[r].combo;
        }
        for (uint32_t g = 0; g <= NUM_GROUPS; g++) {
            uint32_t k = 0;
            while (k < taps.size() && taps[k].group < g) k++;
            group_tap_begin[g] = k;
        }
    }
    // upstream `CircuitImpl::CIRCUIT_INFO` (risc0-zkp `ProtocolInfo`, 16 bytes): hashed into the transcript before anything else.
    // The built-in stand-in names itself; a data-defined circuit (upstream's tables) carries upstream's rv32im-v2 string.
    char circuit_info[17] = "SYNTH_RV32IM:v1_";
    void set_ir(const std::vector<PolyStep>& steps, uint32_t ret, const uint8_t* info16 = nullptr) {
        ir = steps; ir_ret = ret;
        if (info16) { std::memcpy(circuit_info, info16, 16); circuit_info[16] = 0; } else std::memcpy(circuit_info, "RV32IM:v2_______", 17);
    }
    uint32_t group_width(uint32_t g) const { return g == GROUP_ACCUM ? w_accum : g == GROUP_CODE ? w_code : w_data; }
    uint32_t n_constraints() const { return n_free + 4 * n_chains + 1; }
    uint32_t n_mix() const { return 4 * n_chains; }

    // ---- column pickers of the derived-column constraints (k < n_free) ----
    uint32_t pick_a(uint32_t k) const { return k % n_free; }
    uint32_t pick_b(uint32_t k) const { return (5 * k + 1) % n_free; }
    uint32_t pick_c(uint32_t k) const { return (11 * k + 2) % n_free; }
    uint32_t pick_d(uint32_t k) const { return (17 * k + 3) % n_free; }
    uint32_t pick_p(uint32_t k) const { return (7 * k + 1) % n_prev; }
    uint32_t pick_x(uint32_t k) const { return CODE_FIXED + k % (w_code - CODE_FIXED); }
    uint32_t chain_src(uint32_t r) const { return (13 * r + 5) % w_data; }

    // expr_k over any commutative ring V that Fp embeds into.
    template <typename V>
    V derived_expr(uint32_t k, const V& A, const V& B, const V& C, const V& D, const V& Pp, const V& X) const {
        switch (k & 3) {
            case 0: return A * B + C;
            case 1: return A * B * C + Pp;  // Pp = P at back 1 (variant 0) or back 2 (variant 1), chosen by the caller
            case 2: return (A + X) * B * C * D;
            default: return Pp * B + C * D + X;
        }
    }

    // ---- control columns (depend on the circuit and po2 only; their Merkle root is the control id) ----
    void gen_code(Fp* code, unsigned po2) const {
        size_t n = (size_t)1 << po2, act = n - ZK_CYCLES;
        #pragma omp parallel for schedule(static)
        for (long c = 0; c < (long)w_code; c++) {
            Fp* col = code + (size_t)c * n;
            for (size_t r = 0; r < n; r++) {
                Fp v = fp_zero();
                if (r < act) {
                    if (c == 0) v = fp_one();
                    else if (c == 1) v = r == 0 ? fp_one() : fp_zero();
                    else if (c == 2) v = r == act - 1 ? fp_one() : fp_zero();
                    else if (c == 3) v = Fp::from_u64(r);
                    else v = synth_value(code_seed, (uint32_t)c, (uint32_t)r);
                }
                col[r] = v;
            }
        }
    }
    void gen_globals(Fp* g, uint64_t seed) const {
        for (uint32_t i = 0; i < N_GLOBAL; i++) g[i] = synth_value(seed ^ 0x676C6F62ull, 0xFFFFu, i);
    }
    // Witness stand-in: free columns pseudo-random (row 0 of column 0 pinned to global[0]), the last
    // ZK_CYCLES rows of every column blinding noise, derived columns solved from the constraints.
    void gen_data(Fp* data, const Fp* code, const Fp* globals, unsigned po2, uint64_t trace_seed, uint64_t blind_seed) const {
        size_t n = (size_t)1 << po2, act = n - ZK_CYCLES;
        #pragma omp parallel for schedule(static)
        for (long c = 0; c < (long)w_data; c++) {
            Fp* col = data + (size_t)c * n;
            if ((uint32_t)c < n_free) for (size_t r = 0; r < act; r++) col[r] = synth_value(trace_seed, (uint32_t)c, (uint32_t)r);
            for (size_t r = act; r < n; r++) col[r] = blind_value(BlindKey(blind_seed), GROUP_DATA, (uint32_t)c, (uint32_t)r);
        }
        data[0] = globals[0];
        #pragma omp parallel for schedule(static)
        for (long k = 0; k < (long)n_free; k++) {
            Fp* out = data + (size_t)(n_free + k) * n;
            const Fp *A = data + (size_t)pick_a(k) * n, *B = data + (size_t)pick_b(k) * n, *C = data + (size_t)pick_c(k) * n;
            const Fp *D = data + (size_t)pick_d(k) * n, *Pp = data + (size_t)pick_p(k) * n, *X = code + (size_t)pick_x(k) * n;
            const size_t pb = (variant && (k & 3) == 1) ? 2 : 1;
            for (size_t r = 0; r < act; r++) out[r] = derived_expr<Fp>((uint32_t)k, A[r], B[r], C[r], D[r], Pp[(r + n - pb) & (n - 1)], X[r]);
        }
    }
    // step_accum stand-in: chain r is the running product over active rows of (data[src_r] + mix_r).
    void step_accum(Fp* accum, const Fp* data, const Fp* mix, unsigned po2, uint64_t blind_seed) const {
        size_t n = (size_t)1 << po2, act = n - ZK_CYCLES;
        #pragma omp parallel for schedule(static)
        for (long r = 0; r < (long)n_chains; r++) {
            const Fp* src = data + (size_t)chain_src(r) * n;
            Fp4 m(mix[4 * r], mix[4 * r + 1], mix[4 * r + 2], mix[4 * r + 3]);
            Fp4 acc = Fp4::one();
            for (size_t i = 0; i < n; i++) {
                if (i < act) {
                    acc = acc * (Fp4(src[i]) + m);
                    for (int k = 0; k < 4; k++) accum[(size_t)(4 * r + k) * n + i] = acc.c[k];
                } else {
                    for (int k = 0; k < 4; k++) accum[(size_t)(4 * r + k) * n + i] = blind_value(BlindKey(blind_seed), GROUP_ACCUM, (uint32_t)(4 * r + k), (uint32_t)i);
                }
            }
        }
    }

    // ---- the constraint polynomial ----
    // V = Fp on the LDE domain (prover, `poly_fp`), V = Fp4 at the DEEP point z (verifier, `poly_ext`).
    // get(group, offset, back) returns the tapped value.  Result = sum_j poly_mix^j * constraint_j.
    // Interpreter of the PolyStep list (MixState {tot, mul} semantics of upstream's PolyExtStep::step).
    template <typename V, typename Get>
    Fp4 poly_ir(const Fp4& poly_mix, const Fp* globals, const Fp* mix, Get get) const {
        struct MixState { Fp4 tot, mul; };
        std::vector<V> fp; fp.reserve(ir.size());
        std::vector<MixState> ms;
        for (const PolyStep& st : ir) {
            switch (st.op) {
                case OP_CONST: fp.push_back(V(Fp::from_u32(st.a))); break;
                case OP_GET: { const Tap& t = taps.at(st.a); fp.push_back(get(t.group, t.offset, t.back)); break; }
                case OP_GET_GLOBAL: fp.push_back(V(st.a == 0 ? globals[st.b] : mix[st.b])); break;
                case OP_ADD: fp.push_back(fp.at(st.a) + fp.at(st.b)); break;
                case OP_SUB: fp.push_back(fp.at(st.a) - fp.at(st.b)); break;
                case OP_MUL: fp.push_back(fp.at(st.a) * fp.at(st.b)); break;
                case OP_TRUE: ms.push_back(MixState{Fp4::zero(), Fp4::one()}); break;
                case OP_AND_EQZ: { const MixState x = ms.at(st.a); ms.push_back(MixState{x.tot + x.mul * fp.at(st.b), x.mul * poly_mix}); break; }
                case OP_AND_COND: { const MixState x = ms.at(st.a), in = ms.at(st.c); ms.push_back(MixState{x.tot + (in.tot * x.mul) * fp.at(st.b), x.mul * in.mul}); break; }
                default: throw std::runtime_error("circuit: bad poly op");
            }
        }
        return ms.at(ir_ret).tot;
    }

    template <typename V, typename Get>
    Fp4 poly(const Fp4& poly_mix, const Fp* globals, const Fp* mix, Get get) const {
        if (!ir.empty()) return poly_ir<V>(poly_mix, globals, mix, get);
        auto emb = [](Fp f) -> V { return V(f); };
        Fp4 tot = Fp4::zero(), cur = Fp4::one();
        V active = get(GROUP_CODE, 0, 0), first = get(GROUP_CODE, 1, 0);
        for (uint32_t k = 0; k < n_free; k++) {
            V e = derived_expr<V>(k, get(GROUP_DATA, pick_a(k), 0), get(GROUP_DATA, pick_b(k), 0), get(GROUP_DATA, pick_c(k), 0),
                                  get(GROUP_DATA, pick_d(k), 0), get(GROUP_DATA, pick_p(k), (variant && (k & 3) == 1) ? 2 : 1), get(GROUP_CODE, pick_x(k), 0));
            V cv = active * (get(GROUP_DATA, n_free + k, 0) - e);
            tot += cur * cv; cur *= poly_mix;
        }
        V one = emb(fp_one());
        for (uint32_t r = 0; r < n_chains; r++) {
            // Fp4-valued constraint over V-valued columns: components are handled as a 4-vector of V with
            // the x^4 = -11 reduction, so the same code serves V = Fp and V = Fp4.
            V acc[4], prev[4], s[4], t[4];
            for (int k = 0; k < 4; k++) { acc[k] = get(GROUP_ACCUM, 4 * r + k, 0); prev[k] = get(GROUP_ACCUM, 4 * r + k, 1); }
            // s = first*1 + (1-first)*prev
            V nf = one - first;
            for (int k = 0; k < 4; k++) s[k] = nf * prev[k];
            s[0] = s[0] + first;
            // t = d + mix_r
            V d = get(GROUP_DATA, chain_src(r), 0);
            for (int k = 0; k < 4; k++) t[k] = emb(mix[4 * r + k]);
            t[0] = t[0] + d;
            V nb = emb(Fp::from_u32(P - 11));
            V pr[4];
            pr[0] = s[0] * t[0] + nb * (s[1] * t[3] + s[2] * t[2] + s[3] * t[1]);
            pr[1] = s[0] * t[1] + s[1] * t[0] + nb * (s[2] * t[3] + s[3] * t[2]);
            pr[2] = s[0] * t[2] + s[1] * t[1] + s[2] * t[0] + nb * (s[3] * t[3]);
            pr[3] = s[0] * t[3] + s[1] * t[2] + s[2] * t[1] + s[3] * t[0];
            for (int k = 0; k < 4; k++) {
                V cv = active * (acc[k] - pr[k]);
                tot += cur * cv; cur *= poly_mix;
            }
        }
        {
            V cv = first * (get(GROUP_DATA, 0, 0) - emb(globals[0]));
            tot += cur * cv; cur *= poly_mix;
        }
        return tot;
    }
};

}  // namespace orc
