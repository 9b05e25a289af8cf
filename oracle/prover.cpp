// ORACLE (test infrastructure only).  See prover.h for provenance; PARITY UNPINNED vs upstream.
#include "prover.h"
#include <chrono>
#include <cstdio>

namespace orc {

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static void fri_prove(WriteIOP& iop, Checkpoints& cp, std::vector<Fp> coeffs /* [4][n] bit-reversed */, size_t n,
                      const std::function<void(WriteIOP&, size_t)>& inner) {
    const size_t orig_domain = n * INV_RATE;
    std::vector<FriRoundTree*> rounds;
    while (n > FRI_MIN_DEGREE) {
        size_t domain = n * INV_RATE;
        FriRoundTree* rt = new FriRoundTree();
        rt->domain = domain;
        rt->evaluated.resize(EXT_SIZE * domain);
        #pragma omp parallel for
        for (long k = 0; k < (long)EXT_SIZE; k++) expand_into_evaluate_ntt(&rt->evaluated[k * domain], &coeffs[k * n], n, 2);
        // 4 columns x domain rows viewed as (domain/16) rows x 64 columns
        rt->merkle = new MerkleTreeProver(rt->evaluated.data(), domain / FRI_FOLD, FRI_FOLD * EXT_SIZE);
        rt->merkle->commit(iop);
        cp.add("fri_root_" + std::to_string(rounds.size()), rt->merkle->root());
        Fp4 fm = iop.random_ext_elem();
        cp.add("fri_mix_" + std::to_string(rounds.size()), fm);
        size_t m = n / FRI_FOLD;
        std::vector<Fp> out(EXT_SIZE * m);
        #pragma omp parallel for schedule(static)
        for (long idx = 0; idx < (long)m; idx++) {
            Fp4 tot = Fp4::zero(), cur = Fp4::one();
            for (uint32_t i = 0; i < FRI_FOLD; i++) {
                size_t src = (size_t)bit_rev(i, 4) * m + idx;
                Fp4 e(coeffs[0 * n + src], coeffs[1 * n + src], coeffs[2 * n + src], coeffs[3 * n + src]);
                tot += cur * e;
                cur *= fm;
            }
            for (size_t k = 0; k < EXT_SIZE; k++) out[k * m + idx] = tot.c[k];
        }
        coeffs.swap(out);
        n = m;
        rounds.push_back(rt);
    }
    for (size_t k = 0; k < EXT_SIZE; k++) bit_reverse_inplace(&coeffs[k * n], n);
    iop.write_elems(coeffs.data(), coeffs.size());
    Digest fd = hash_elems(coeffs.data(), coeffs.size());
    iop.commit(fd);
    cp.add("fri_final_hash", fd);
    std::vector<uint32_t> positions;
    for (size_t q = 0; q < QUERIES; q++) {
        uint32_t rng = iop.random_bits(log2_exact(orig_domain));
        size_t pos = rng % orig_domain;
        positions.push_back((uint32_t)pos);
        inner(iop, pos);
        for (FriRoundTree* rt : rounds) {
            size_t rows = rt->domain / FRI_FOLD;
            size_t group = pos % rows;
            rt->merkle->prove(iop, group);
            pos = group;
        }
    }
    cp.add("query_positions", positions.data(), positions.size());
    for (FriRoundTree* rt : rounds) { delete rt->merkle; delete rt; }
}

// risc0-zkp `ProtocolInfo::encode`: one field element per byte of the 16-byte info string
static Digest protocol_info_digest(const char* info16) {
    Fp e[16];
    for (int i = 0; i < 16; i++) e[i] = Fp::from_u32((uint8_t)info16[i]);
    return hash_elems(e, 16);
}
// The three commits every transcript starts with (prover and verifier); returns the header digest H(globals ++ [po2]).
static Digest transcript_header(Poseidon2Rng& rng, const char* circuit_info, const Fp* globals, uint32_t po2) {
    rng.mix(protocol_info_digest("RISC0_STARK:v1__"));  // risc0-zkp PROOF_SYSTEM_INFO
    rng.mix(protocol_info_digest(circuit_info));
    Fp header[Circuit::N_GLOBAL + 1];
    for (size_t i = 0; i < Circuit::N_GLOBAL; i++) header[i] = globals[i];
    header[Circuit::N_GLOBAL] = Fp::raw(po2);
    const Digest d = hash_elems(header, Circuit::N_GLOBAL + 1);
    rng.mix(d);
    return d;
}

SegmentProof prove_segment(const Circuit& cir, unsigned po2, const Fp* globals, const Fp* code, const Fp* data,
                           uint64_t blind_seed, OracleTimes* times) {
    if (po2 < 12 || po2 > 24) throw std::runtime_error("prove: po2 out of range");
    const size_t N = (size_t)1 << po2, domain = N * INV_RATE;
    double t0 = now_s(), t_start = t0;
    SegmentProof out;
    Checkpoints& cp = out.cp;
    WriteIOP iop;

    // risc0-circuit-rv32im `SegmentProver::prove`: "At the start of the protocol, seed the Fiat-Shamir transcript with context
    // information about the proof system and circuit": commit(H(PROOF_SYSTEM_INFO.encode())), commit(H(CIRCUIT_INFO.encode()));
    // then "Concat globals and po2 into a vector": commit(H(header)), write header (po2 as a raw word).
    Digest gh = transcript_header(iop.rng, cir.circuit_info, globals, po2);
    iop.write_elems(globals, Circuit::N_GLOBAL);
    uint32_t po2w = po2;
    iop.write_u32s(&po2w, 1);
    cp.add("globals_hash", gh);

    PolyGroup code_g(make_coeffs(code, cir.w_code, N, true), cir.w_code, N);
    code_g.merkle->commit(iop);
    cp.add("code_root", code_g.merkle->root());
    PolyGroup data_g(make_coeffs(data, cir.w_data, N, true), cir.w_data, N);
    data_g.merkle->commit(iop);
    cp.add("data_root", data_g.merkle->root());
    if (times) { times->commit += now_s() - t0; } t0 = now_s();

    std::vector<Fp> mix(cir.n_mix());
    for (auto& m : mix) m = iop.random_elem();
    cp.add("accum_mix", reinterpret_cast<const uint32_t*>(mix.data()), mix.size());
    std::vector<Fp> accum((size_t)cir.w_accum * N);
    cir.step_accum(accum.data(), data, mix.data(), po2, blind_seed);
    if (times) { times->accum += now_s() - t0; } t0 = now_s();
    PolyGroup accum_g(make_coeffs(accum.data(), cir.w_accum, N, true), cir.w_accum, N);
    accum_g.merkle->commit(iop);
    cp.add("accum_root", accum_g.merkle->root());
    if (times) { times->commit += now_s() - t0; } t0 = now_s();

    // ---- finalize ----
    const PolyGroup* groups[NUM_GROUPS] = {&accum_g, &code_g, &data_g};
    Fp4 poly_mix = iop.random_ext_elem();
    cp.add("poly_mix", poly_mix);
    std::vector<Fp> check(EXT_SIZE * domain);
    {
        const Fp wd = rou_fwd(po2 + 2), three = Fp::from_u32(3);
        #pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)domain; i++) {
            auto get = [&](uint32_t g, uint32_t off, uint32_t back) -> Fp {
                size_t row = ((size_t)i + domain - INV_RATE * back) & (domain - 1);
                return groups[g]->evaluated[(size_t)off * domain + row];
            };
            Fp4 tot = cir.poly<Fp>(poly_mix, globals, mix.data(), get);
            Fp x = wd.pow((uint64_t)i);
            Fp y = (three * x).pow(N);
            Fp4 ret = tot * (y - fp_one()).inv();
            for (size_t k = 0; k < EXT_SIZE; k++) check[k * domain + i] = ret.c[k];
        }
    }
    // 4 polys of size 4N -> (bit-reversed order makes this free) 16 polys of size N; no zk_shift:
    // the evaluations were taken at y = w_4N^i of polynomials in y.
    PolyGroup check_g(make_coeffs(check.data(), EXT_SIZE, domain, false), CHECK_SIZE, N);
    check_g.merkle->commit(iop);
    cp.add("check_root", check_g.merkle->root());
    if (times) { times->check += now_s() - t0; } t0 = now_s();

    Fp4 z = iop.random_ext_elem();
    cp.add("z", z);
    const Fp back_one = rou_rev(po2);
    const size_t T = cir.taps.size();
    std::vector<Fp4> eval_u(T), coeff_u(T + CHECK_SIZE);
    #pragma omp parallel for schedule(dynamic)
    for (long t = 0; t < (long)T; t++) {
        const Tap& tp = cir.taps[t];
        Fp4 x = z * back_one.pow(tp.back);
        eval_u[t] = poly_eval_base(&groups[tp.group]->coeffs[(size_t)tp.offset * N], N, x);
    }
    for (const Reg& r : cir.regs) {
        Fp4 xs[4];
        for (uint32_t i = 0; i < r.size; i++) xs[i] = z * back_one.pow(cir.taps[r.tap_begin + i].back);
        poly_interpolate(&coeff_u[r.tap_begin], xs, &eval_u[r.tap_begin], r.size);
    }
    const Fp4 z4 = z.pow(EXT_SIZE);
    #pragma omp parallel for
    for (long c = 0; c < (long)CHECK_SIZE; c++) coeff_u[T + c] = poly_eval_base(&check_g.coeffs[(size_t)c * N], N, z4);
    iop.write_ext_elems(coeff_u.data(), coeff_u.size());
    Digest hash_u = hash_ext_elems(coeff_u.data(), coeff_u.size());
    iop.commit(hash_u);
    cp.add("hash_u", hash_u);

    Fp4 dmix = iop.random_ext_elem();
    cp.add("deep_mix", dmix);
    const size_t C = cir.combos.size();
    std::vector<Fp4> combos((C + 1) * N, Fp4::zero());
    {
        // per-register mix powers (registers in taps order: accum, code, data), then the 16 check polys
        std::vector<Fp4> reg_mix(cir.regs.size() + CHECK_SIZE);
        Fp4 cur = Fp4::one();
        for (auto& m : reg_mix) { m = cur; cur *= dmix; }
        #pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; i++) {
            for (size_t ri = 0; ri < cir.regs.size(); ri++) {
                const Reg& r = cir.regs[ri];
                combos[(size_t)r.combo * N + i] += reg_mix[ri] * groups[r.group]->coeffs[(size_t)r.offset * N + i];
            }
            for (size_t c = 0; c < CHECK_SIZE; c++) combos[C * N + i] += reg_mix[cir.regs.size() + c] * check_g.coeffs[c * N + i];
        }
        size_t pos = 0;
        for (size_t ri = 0; ri < cir.regs.size(); ri++) {
            const Reg& r = cir.regs[ri];
            for (uint32_t i = 0; i < r.size; i++) combos[(size_t)r.combo * N + i] -= reg_mix[ri] * coeff_u[pos + i];
            pos += r.size;
        }
        for (size_t c = 0; c < CHECK_SIZE; c++) combos[C * N] -= reg_mix[cir.regs.size() + c] * coeff_u[pos++];
    }
    for (size_t c = 0; c < C; c++)
        for (uint32_t back : cir.combos[c])
            if (poly_divide(&combos[c * N], N, z * back_one.pow(back)) != Fp4::zero()) throw std::runtime_error("prove: DEEP quotient has a remainder (trace violates the circuit?)");
    if (poly_divide(&combos[C * N], N, z4) != Fp4::zero()) throw std::runtime_error("prove: check quotient has a remainder");
    std::vector<Fp> final_coeffs(EXT_SIZE * N);
    #pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)N; i++) {
        Fp4 s = Fp4::zero();
        for (size_t c = 0; c <= C; c++) s += combos[c * N + i];
        for (size_t k = 0; k < EXT_SIZE; k++) final_coeffs[k * N + i] = s.c[k];
    }
    for (size_t k = 0; k < EXT_SIZE; k++) bit_reverse_inplace(&final_coeffs[k * N], N);
    cp.add("final_poly_hash", hash_elems(final_coeffs.data(), final_coeffs.size()));
    if (times) { times->deep += now_s() - t0; } t0 = now_s();

    fri_prove(iop, cp, std::move(final_coeffs), N, [&](WriteIOP& w, size_t idx) {
        for (const PolyGroup* g : groups) g->merkle->prove(w, idx);
        check_g.merkle->prove(w, idx);
    });
    if (times) { times->fri += now_s() - t0; times->total += now_s() - t_start; }
    out.seal = std::move(iop.proof);
    return out;
}

// ------------------------------------------------------------------------------------------------
// Verifier (mirror of risc0-zkp `verify::Verifier::verify` + `verify::fri::fri_verify`)
// ------------------------------------------------------------------------------------------------
static void fail(const char* why) { throw std::runtime_error(std::string("verify: ") + why); }

struct VerifyRound {
    size_t rows;  // domain / FRI_FOLD
    MerkleTreeVerifier merkle;
    Fp4 mix;
    VerifyRound(ReadIOP& iop, size_t in_domain) : rows(in_domain / FRI_FOLD), merkle(iop, in_domain / FRI_FOLD, FRI_FOLD * EXT_SIZE) { mix = iop.random_ext_elem(); }
    void verify_query(ReadIOP& iop, size_t& pos, Fp4& goal) const {
        size_t quot = pos / rows, group = pos % rows;
        std::vector<Fp> data = merkle.verify(iop, group);
        Fp4 ext[FRI_FOLD];
        for (size_t i = 0; i < FRI_FOLD; i++) ext[i] = Fp4(data[0 * FRI_FOLD + i], data[1 * FRI_FOLD + i], data[2 * FRI_FOLD + i], data[3 * FRI_FOLD + i]);
        if (ext[quot] != goal) fail("FRI query value does not match the previous layer");
        // fold: the 16 values are f(x0 * w_16^q); interpolate, undo x0^i, combine with mix^i
        interpolate_ntt(ext, FRI_FOLD);
        bit_reverse_inplace(ext, FRI_FOLD);
        unsigned root_po2 = log2_exact(FRI_FOLD * rows);
        Fp inv_wk = rou_rev(root_po2).pow(group);
        Fp4 tot = Fp4::zero(), mul_mix = Fp4::one();
        Fp mul = fp_one();
        for (size_t i = 0; i < FRI_FOLD; i++) { tot += ext[i] * mul * mul_mix; mul_mix *= mix; mul *= inv_wk; }
        goal = tot;
        pos = group;
    }
};

static void fri_verify(ReadIOP& iop, size_t degree, const std::function<Fp4(ReadIOP&, size_t)>& inner) {
    const size_t orig_domain = INV_RATE * degree;
    size_t domain = orig_domain;
    std::vector<VerifyRound> rounds;
    while (degree > FRI_MIN_DEGREE) {
        rounds.emplace_back(iop, domain);
        domain /= FRI_FOLD;
        degree /= FRI_FOLD;
    }
    std::vector<Fp> final_coeffs(EXT_SIZE * degree);
    iop.read_elems(final_coeffs.data(), final_coeffs.size());
    iop.commit(hash_elems(final_coeffs.data(), final_coeffs.size()));
    Fp gen = rou_fwd(log2_exact(domain));
    std::vector<Fp4> poly(degree);
    for (size_t i = 0; i < degree; i++) poly[i] = Fp4(final_coeffs[i], final_coeffs[degree + i], final_coeffs[2 * degree + i], final_coeffs[3 * degree + i]);
    for (size_t q = 0; q < QUERIES; q++) {
        uint32_t rng = iop.random_bits(log2_exact(orig_domain));
        size_t pos = rng % orig_domain;
        Fp4 goal = inner(iop, pos);
        for (const VerifyRound& r : rounds) r.verify_query(iop, pos, goal);
        Fp4 x(gen.pow(pos));
        if (poly_eval(poly.data(), degree, x) != goal) fail("FRI final polynomial does not match the folded query");
    }
}

void verify_segment(const Circuit& cir, const uint32_t* seal, size_t seal_words, const Digest& code_root, unsigned* po2_out) {
    ReadIOP iop(seal, seal_words);
    Fp globals[Circuit::N_GLOBAL];
    iop.read_elems(globals, Circuit::N_GLOBAL);
    uint32_t po2;
    iop.read_u32s(&po2, 1);
    if (po2 < 12 || po2 > 24) fail("po2 out of range");
    if (po2_out) *po2_out = po2;
    transcript_header(iop.rng, cir.circuit_info, globals, po2);
    const size_t N = (size_t)1 << po2, domain = N * INV_RATE;

    MerkleTreeVerifier code_v(iop, domain, cir.w_code);
    if (code_v.root() != code_root) fail("code root is not the control id for this po2");
    MerkleTreeVerifier data_v(iop, domain, cir.w_data);
    std::vector<Fp> mix(cir.n_mix());
    for (auto& m : mix) m = iop.random_elem();
    MerkleTreeVerifier accum_v(iop, domain, cir.w_accum);
    Fp4 poly_mix = iop.random_ext_elem();
    MerkleTreeVerifier check_v(iop, domain, CHECK_SIZE);
    Fp4 z = iop.random_ext_elem();
    const Fp back_one = rou_rev(po2);
    const size_t T = cir.taps.size();
    std::vector<Fp4> coeff_u(T + CHECK_SIZE);
    iop.read_ext_elems(coeff_u.data(), coeff_u.size());
    iop.commit(hash_ext_elems(coeff_u.data(), coeff_u.size()));

    // evaluations at z * w^-back from the per-register interpolants
    std::vector<Fp4> eval_u(T);
    std::vector<std::vector<uint32_t>> tap_of(NUM_GROUPS);
    for (uint32_t g = 0; g < NUM_GROUPS; g++) tap_of[g].resize(cir.group_width(g));
    for (const Reg& r : cir.regs) {
        tap_of[r.group][r.offset] = r.tap_begin;
        for (uint32_t i = 0; i < r.size; i++) {
            Fp4 x = z * back_one.pow(cir.taps[r.tap_begin + i].back);
            eval_u[r.tap_begin + i] = poly_eval(&coeff_u[r.tap_begin], r.size, x);
        }
    }
    std::vector<std::vector<uint32_t>> reg_size(NUM_GROUPS);
    for (uint32_t g = 0; g < NUM_GROUPS; g++) reg_size[g].assign(cir.group_width(g), 0);
    for (const Reg& r : cir.regs) reg_size[r.group][r.offset] = r.size;
    auto get = [&](uint32_t g, uint32_t off, uint32_t back) -> Fp4 {
        const uint32_t t0 = tap_of[g][off];
        for (uint32_t k = 0; k < reg_size[g][off]; k++) if (cir.taps[t0 + k].back == back) return eval_u[t0 + k];
        throw std::runtime_error("verify: constraint polynomial reads a tap that is not in the tap set");
    };
    Fp4 result = cir.poly<Fp4>(poly_mix, globals, mix.data(), get);
    // check(z) from the 16 check polys evaluated at z^4: check_k(y) = sum_ch y^rev2(ch) * P_{k,ch}(y^4)
    Fp4 check = Fp4::zero();
    for (uint32_t k = 0; k < EXT_SIZE; k++) {
        Fp4 basis = Fp4::zero(); basis.c[k] = fp_one();
        for (uint32_t i = 0; i < INV_RATE; i++) check += coeff_u[T + 4 * k + bit_rev(i, 2)] * z.pow(i) * basis;
    }
    check *= (z * Fp::from_u32(3)).pow(N) - Fp4::one();
    if (check != result) fail("constraint polynomial does not match the check polynomial at z");

    Fp4 dmix = iop.random_ext_elem();
    const size_t C = cir.combos.size();
    std::vector<Fp4> combo_u(cir.tot_combo_backs + 1, Fp4::zero());
    std::vector<Fp4> reg_mix(cir.regs.size() + CHECK_SIZE);
    {
        Fp4 cur = Fp4::one();
        for (auto& m : reg_mix) { m = cur; cur *= dmix; }
        for (size_t ri = 0; ri < cir.regs.size(); ri++) {
            const Reg& r = cir.regs[ri];
            for (uint32_t i = 0; i < r.size; i++) combo_u[cir.combo_begin[r.combo] + i] += reg_mix[ri] * coeff_u[r.tap_begin + i];
        }
        for (size_t c = 0; c < CHECK_SIZE; c++) combo_u[cir.tot_combo_backs] += reg_mix[cir.regs.size() + c] * coeff_u[T + c];
    }
    const Fp4 z4 = z.pow(EXT_SIZE);
    const Fp gen = rou_fwd(po2 + 2);
    const MerkleTreeVerifier* gv[NUM_GROUPS] = {&accum_v, &code_v, &data_v};
    fri_verify(iop, N, [&](ReadIOP& r, size_t idx) -> Fp4 {
        Fp4 x(gen.pow(idx));
        std::vector<Fp> rows[NUM_GROUPS];
        for (uint32_t g = 0; g < NUM_GROUPS; g++) rows[g] = gv[g]->verify(r, idx);
        std::vector<Fp> check_row = check_v.verify(r, idx);
        std::vector<Fp4> tot(C + 1, Fp4::zero());
        for (size_t ri = 0; ri < cir.regs.size(); ri++) {
            const Reg& rg = cir.regs[ri];
            tot[rg.combo] += reg_mix[ri] * rows[rg.group][rg.offset];
        }
        for (size_t c = 0; c < CHECK_SIZE; c++) tot[C] += reg_mix[cir.regs.size() + c] * check_row[c];
        Fp4 ret = Fp4::zero();
        for (size_t c = 0; c < C; c++) {
            Fp4 num = tot[c] - poly_eval(&combo_u[cir.combo_begin[c]], cir.combos[c].size(), x);
            Fp4 div = Fp4::one();
            for (uint32_t back : cir.combos[c]) div *= x - z * back_one.pow(back);
            ret += num * div.inv();
        }
        ret += (tot[C] - combo_u[cir.tot_combo_backs]) * (x - z4).inv();
        return ret;
    });
    iop.verify_complete();
}

Digest control_id(const Circuit& cir, unsigned po2) {
    const size_t N = (size_t)1 << po2;
    std::vector<Fp> code((size_t)cir.w_code * N);
    cir.gen_code(code.data(), po2);
    PolyGroup g(make_coeffs(code.data(), cir.w_code, N, true), cir.w_code, N);
    return g.merkle->root();
}

// Seal length in u32 words (SURVEY.md Appendix B model; asserted against the emitted seal in tests).
size_t seal_words_model(const Circuit& cir, unsigned po2) {
    const size_t W = cir.w_code + cir.w_data + cir.w_accum, T = cir.taps.size();
    auto path = [](size_t rows) { MerkleParams p(rows, 1); return 8 * (p.layers - p.top_layer); };
    auto tops = [](size_t rows) { MerkleParams p(rows, 1); return 8 * p.top_size; };
    size_t n = (size_t)1 << po2, domain = 4 * n;
    size_t words = Circuit::N_GLOBAL + 1 + 4 * tops(domain) + 4 * (T + CHECK_SIZE);
    size_t per_query = W + CHECK_SIZE + 4 * path(domain);
    while (n > FRI_MIN_DEGREE) {
        size_t rows = 4 * n / FRI_FOLD;
        words += tops(rows);
        per_query += 64 + path(rows);
        n /= FRI_FOLD;
    }
    words += 4 * n;
    return words + QUERIES * per_query;
}

}  // namespace orc
