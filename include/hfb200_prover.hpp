// hfb200_prover.hpp -- C++17 host side above the C ABI (include/hfb200.h): the reference's prover surface for the ONE hot path.
//
// The reference is compiled Rust and keeps this surface (/root/reference/host/src/main.rs):
//     let prover = default_prover();                                   // :420
//     let receipt = prover.prove(env, HYPERFRIDGE_ELF)?.receipt;       // :423   (ProveInfo { receipt, .. })
//     serde_json::to_string(&receipt)                                  // :250-252
//     receipt.journal.bytes / journal decode                           // :258-267
//     receipt.verify(HYPERFRIDGE_ID)                                   // :622-624, /root/reference/verifier/src/main.rs:118-126
// No Rust toolchain exists in this image, so the host mirror is written in C++ (header only, nothing but the C ABI underneath)
// with the same names, argument meaning and error behaviour: errors are exceptions carrying the library's message (anyhow::Error
// there), a dev-mode `Fake` receipt is refused by verify(), ProverOpts other than poseidon2 / composite are refused at construction.
// The executor (ELF -> Session -> Segments) is out of scope: a Session here is the list of segment traces
// the executor + witness generator would have produced, plus the journal the guest committed.  The Rust binding a maintainer adds is
// in hyperfridge-r0_b200/rust/ and INTEGRATION.md.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "hfb200.h"

namespace hfb200 {

struct Error : std::runtime_error { using std::runtime_error::runtime_error; };

// risc0_sys::ffi_wrap: NULL = ok, else a malloc'd message the caller frees
inline void ffi_wrap(const char* e) {
    if (!e) return;
    const std::string msg(e);
    hfb200_free_error(e);
    throw Error(msg);
}

using Digest = std::array<uint32_t, 8>;

// Subset of risc0_zkvm::ProverOpts that matters on this path (the host passes none: defaults only).
struct ProverOpts {
    std::string hashfn = "poseidon2";
    std::string receipt_kind = "composite";
    uint32_t max_segment_po2 = 20;
    hfb200_circuit_desc circuit{16, 192, 48, 0};
    std::vector<int> devices{0};
    int contexts_per_device = 2;
    bool reuse_control = false;  // opt-in: segments of equal po2 share the control group of the first one (true for rv32im)
    bool deterministic_blinding = false;  // tests / bench only: blinding from Segment::blind_seed alone (NOT zero-knowledge)
};

// Stand-in for HYPERFRIDGE_ID (/root/reference/host/src/main.rs:7): the executor is out of scope, so there is no ELF to digest.
inline Digest default_image_id() {
    static const char tag[] = "hyperfridge guest image id (stand-in: executor out of scope)";
    Digest d;
    ffi_wrap(hfb200_digest_bytes(reinterpret_cast<const uint8_t*>(tag), sizeof tag - 1, d.data()));
    return d;
}

// What upstream's `Segment` boils down to at the prover seam: po2 and the witness columns (+ blinding seed).  Host pointers,
// column-major u32[w][2^po2] Montgomery residues, owned by the caller for the duration of prove().
struct Segment {
    uint32_t index = 0, po2 = 0;
    const uint32_t* globals = nullptr;  // [32]
    const uint32_t* code = nullptr;
    const uint32_t* data = nullptr;
    uint64_t blind_seed = 0;
    // optional: words behind `code` / `data`; when non-zero, prove() checks them against (circuit, po2) before any pointer reaches the library
    size_t code_words = 0, data_words = 0;
};
struct Session {
    std::vector<Segment> segments;
    std::string journal;  // the String the guest committed (risc0 serde: u32-LE length, bytes, zero pad to 4)
    bool has_image_id = false;
    Digest image_id{};    // pre-state of segment 0; default_image_id() when has_image_id is false
};

struct Journal {
    std::vector<uint8_t> bytes;
    static Journal encode(const std::string& text) {
        Journal j;
        const uint32_t n = (uint32_t)text.size();
        for (int k = 0; k < 4; k++) j.bytes.push_back((uint8_t)(n >> (8 * k)));
        j.bytes.insert(j.bytes.end(), text.begin(), text.end());
        while (j.bytes.size() % 4) j.bytes.push_back(0);
        return j;
    }
    std::string decode() const {
        if (bytes.size() < 4) throw Error("journal too short");
        const uint32_t n = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | ((uint32_t)bytes[3] << 24);
        if (4 + (size_t)n > bytes.size()) throw Error("journal length prefix exceeds the payload");
        return std::string(bytes.begin() + 4, bytes.begin() + 4 + n);
    }
};

struct SegmentReceipt {
    std::vector<uint32_t> seal;
    uint32_t index = 0;
    std::string hashfn = "poseidon2";
};

namespace detail {
// Minimal reader for the serde-JSON shape of a receipt (numbers, strings, arrays, objects; no escapes beyond \" and \\).
struct Json {
    const char* p; const char* end;
    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) p++; }
    bool eat(char c) { ws(); if (p < end && *p == c) { p++; return true; } return false; }
    void need(char c) { if (!eat(c)) throw Error(std::string("receipt json: expected '") + c + "'"); }
    std::string str() {
        need('"');
        std::string s;
        while (p < end && *p != '"') { if (*p == '\\' && p + 1 < end) p++; s.push_back(*p++); }
        need('"');
        return s;
    }
    uint64_t num() {
        ws();
        if (p >= end || *p < '0' || *p > '9') throw Error("receipt json: expected a number");
        uint64_t v = 0;
        while (p < end && *p >= '0' && *p <= '9') { v = v * 10 + (uint64_t)(*p - '0'); if (v > 0xFFFFFFFFull) throw Error("receipt json: number out of range"); p++; }
        return v;
    }
    void skip() {  // any value
        ws();
        if (p >= end) throw Error("receipt json: truncated");
        if (*p == '"') { str(); return; }
        if (*p == '{' || *p == '[') {
            const char open = *p, close = open == '{' ? '}' : ']';
            p++;
            if (eat(close)) return;
            do { if (open == '{') { str(); need(':'); } skip(); } while (eat(','));
            need(close);
            return;
        }
        while (p < end && *p != ',' && *p != '}' && *p != ']') p++;  // number / true / false / null
    }
    template <typename F> void object(F&& field) {
        need('{');
        if (eat('}')) return;
        do { const std::string k = str(); need(':'); field(k); } while (eat(','));
        need('}');
    }
    template <typename F> void array(F&& item) {
        need('[');
        if (eat(']')) return;
        do { item(); } while (eat(','));
        need(']');
    }
};
}  // namespace detail

// risc0_zkvm::Receipt { inner: InnerReceipt::{Composite, Fake}, journal }
struct Receipt {
    bool fake = false;
    std::vector<SegmentReceipt> segments;  // inner = Composite
    Journal journal;

    // the claim a seal commits to (hfb200_claim_decode), in the shape the Python mirror writes; "null" if the seal is malformed
    static std::string claim_json(const std::vector<uint32_t>& seal) {
        hfb200_claim c;
        if (const char* e = hfb200_claim_decode(seal.data(), seal.size(), &c)) { hfb200_free_error(e); return "null"; }
        auto arr = [](const uint32_t* w) { std::string a = "["; for (int i = 0; i < 8; i++) { if (i) a += ','; a += std::to_string(w[i]); } return a + "]"; };
        return std::string("{\"pre\":") + arr(c.pre) + ",\"post\":" + arr(c.post) + ",\"exit_code\":\"" + (c.exit_code ? "SystemSplit" : "Halted") + "\",\"output\":" + arr(c.output) + "}";
    }
    std::string to_json() const {
        std::string s = "{\"inner\":";
        if (fake) s += "\"Fake\"";
        else {
            s += "{\"Composite\":{\"segments\":[";
            for (size_t i = 0; i < segments.size(); i++) {
                if (i) s += ',';
                s += "{\"seal\":[";
                for (size_t k = 0; k < segments[i].seal.size(); k++) { if (k) s += ','; s += std::to_string(segments[i].seal[k]); }
                s += "],\"index\":" + std::to_string(segments[i].index) + ",\"hashfn\":\"" + segments[i].hashfn + "\",\"verifier_parameters\":[0,0,0,0,0,0,0,0],\"claim\":" + claim_json(segments[i].seal) + "}";
            }
            s += "],\"assumption_receipts\":[],\"verifier_parameters\":[0,0,0,0,0,0,0,0]}}";
        }
        s += ",\"journal\":{\"bytes\":[";
        for (size_t k = 0; k < journal.bytes.size(); k++) { if (k) s += ','; s += std::to_string(journal.bytes[k]); }
        s += "]}}";
        return s;
    }
    // serde_json::from_slice::<Receipt> (/root/reference/verifier/src/main.rs:118-119)
    static Receipt from_json(const std::string& text) {
        Receipt r;
        detail::Json j{text.data(), text.data() + text.size()};
        bool have_inner = false, have_journal = false;
        j.object([&](const std::string& k) {
            if (k == "inner") {
                have_inner = true;
                j.ws();
                if (j.p < j.end && *j.p == '"') { if (j.str() != "Fake") throw Error("receipt json: unknown inner receipt"); r.fake = true; return; }
                j.object([&](const std::string& kind) {
                    if (kind != "Composite") throw Error("receipt json: only Composite / Fake receipts are on this path");
                    j.object([&](const std::string& f) {
                        if (f != "segments") { j.skip(); return; }
                        j.array([&] {
                            SegmentReceipt sr;
                            j.object([&](const std::string& g) {
                                if (g == "seal") j.array([&] { sr.seal.push_back((uint32_t)j.num()); });
                                else if (g == "index") sr.index = (uint32_t)j.num();
                                else if (g == "hashfn") sr.hashfn = j.str();
                                else j.skip();
                            });
                            r.segments.push_back(std::move(sr));
                        });
                    });
                });
            } else if (k == "journal") {
                have_journal = true;
                j.object([&](const std::string& f) {
                    if (f != "bytes") { j.skip(); return; }
                    j.array([&] { const uint64_t b = j.num(); if (b > 255) throw Error("receipt json: journal byte out of range"); r.journal.bytes.push_back((uint8_t)b); });
                });
            } else j.skip();
        });
        if (!have_inner || !have_journal) throw Error("receipt json: missing inner / journal");
        return r;
    }
    size_t seal_bytes() const { size_t n = 0; for (const auto& s : segments) n += 4 * s.seal.size(); return n; }

    // `receipt.verify(HYPERFRIDGE_ID)` (/root/reference/host/src/main.rs:622-624, /root/reference/verifier/src/main.rs:124-126, which
    // trusts receipt.journal after this call): (1) every segment seal against the control id of its po2 (upstream's per-po2
    // control-id table), indices 0..n-1, dev-mode receipts refused; (2) the claim chain (hfb200_verify_claims): segment 0 starts from
    // image_id, every segment continues from its predecessor's post-state, only the last one halts and its output is the digest of
    // journal.bytes -- a replaced journal, swapped / dropped segments or another image are rejected.
    void verify(const Digest& image_id, const std::map<uint32_t, Digest>& control_ids,
                const hfb200_circuit_desc& circuit = hfb200_circuit_desc{16, 192, 48, 0}) const {
        verify_seals(control_ids, circuit);
        std::vector<const uint32_t*> ptrs;
        std::vector<size_t> lens;
        for (const auto& s : segments) { ptrs.push_back(s.seal.data()); lens.push_back(s.seal.size()); }
        ffi_wrap(hfb200_verify_claims(ptrs.data(), lens.data(), ptrs.size(), image_id.data(), journal.bytes.data(), journal.bytes.size()));
    }
    // step (1) alone; NOT a substitute for verify(): it says nothing about the journal or the image id
    void verify_seals(const std::map<uint32_t, Digest>& control_ids, const hfb200_circuit_desc& circuit = hfb200_circuit_desc{16, 192, 48, 0}, unsigned threads = 0) const {
        if (fake) throw Error("verify: Fake receipt carries no seal (dev-mode receipts are refused)");
        if (segments.empty()) throw Error("verify: composite receipt without segments");
        std::vector<const uint32_t*> ptrs;
        std::vector<size_t> lens;
        std::vector<uint32_t> roots;
        for (size_t want = 0; want < segments.size(); want++) {
            const SegmentReceipt& s = segments[want];
            if (s.index != want) throw Error("verify: segment index " + std::to_string(s.index) + " at position " + std::to_string(want));
            if (s.hashfn != "poseidon2") throw Error("verify: hash suite " + s.hashfn + " is not on this path");
            if (s.seal.size() < 33) throw Error("verify: segment " + std::to_string(want) + ": seal truncated");
            const uint32_t po2 = s.seal[32];  // seal layout: 32 globals, po2, ...
            const auto it = control_ids.find(po2);
            if (it == control_ids.end()) throw Error("verify: segment " + std::to_string(want) + ": no control id for po2 " + std::to_string(po2));
            ptrs.push_back(s.seal.data()); lens.push_back(s.seal.size());
            roots.insert(roots.end(), it->second.begin(), it->second.end());
        }
        // the seals are independent: one library call fans them out over the host threads (0 = all of them)
        ffi_wrap(hfb200_verify_segments(&circuit, nullptr, ptrs.data(), lens.data(), ptrs.size(), roots.data(), nullptr, threads, nullptr));
    }
};

struct ProveInfo {
    Receipt receipt;
    std::vector<int> devices;      // which GPU proved each segment
    std::vector<float> segment_ms;
};

// `default_prover()` stand-in over hfb200_pool: segments are independent, whole segments go to whichever context is free.
class Prover {
  public:
    explicit Prover(const ProverOpts& o = ProverOpts()) : opts_(o) {
        if (o.hashfn != "poseidon2") throw Error("only the default poseidon2 hash suite is on the hot path (sha-256 suite: out of scope)");
        if (o.receipt_kind != "composite") throw Error("succinct / groth16 receipts need the recursion circuit: out of scope");
        if (o.devices.empty()) throw Error("ProverOpts: no devices");
        ffi_wrap(hfb200_pool_create(o.devices.data(), (int)o.devices.size(), o.contexts_per_device, o.max_segment_po2, &o.circuit, &pool_));
        if (o.deterministic_blinding) {
            const char* e = hfb200_pool_set_blinding(pool_, HFB200_BLIND_DETERMINISTIC);
            if (e) { hfb200_pool_destroy(pool_); pool_ = nullptr; ffi_wrap(e); }
        }
    }
    ~Prover() { hfb200_pool_destroy(pool_); }
    Prover(const Prover&) = delete;
    Prover& operator=(const Prover&) = delete;
    const ProverOpts& opts() const { return opts_; }

    // prover.prove(env, elf) -> ProveInfo
    // n_total_segments: segments of the whole session when this call proves only a share of it (multi-process operation); 0 = all
    ProveInfo prove(const Session& session, size_t seal_cap_words = (size_t)1 << 18, size_t n_total_segments = 0) {
        for (const Segment& s : session.segments) {
            if (s.po2 > opts_.max_segment_po2) throw Error("segment po2 " + std::to_string(s.po2) + " exceeds max_segment_po2 " + std::to_string(opts_.max_segment_po2));
            if (!s.globals || !s.data || (!s.code && !opts_.reuse_control)) throw Error("segment " + std::to_string(s.index) + ": NULL trace pointer");
            if ((s.code_words && s.code_words != ((size_t)opts_.circuit.w_code << s.po2)) || (s.data_words && s.data_words != ((size_t)opts_.circuit.w_data << s.po2)))
                throw Error("segment " + std::to_string(s.index) + ": trace shape does not match (circuit, po2)");
        }
        const size_t n = session.segments.size();
        std::vector<std::vector<uint32_t>> seals(n, std::vector<uint32_t>(seal_cap_words));
        std::vector<hfb200_segment_job> jobs(n);
        if (opts_.reuse_control) {
            std::map<uint32_t, const uint32_t*> first;
            for (const Segment& s : session.segments) {
                const auto it = first.find(s.po2);
                if (it == first.end()) first.emplace(s.po2, s.code);
                else if (it->second != s.code && std::memcmp(it->second, s.code, ((size_t)opts_.circuit.w_code << s.po2) * 4) != 0)
                    throw Error("reuse_control: segments of po2 " + std::to_string(s.po2) + " carry different control columns");
            }
            for (const auto& kv : first) ffi_wrap(hfb200_pool_load_control(pool_, kv.first, kv.second));
        }
        // what upstream's executor does for the prover: every segment's globals carry its claim (state chain from the image id,
        // SystemSplit for all but the last segment, the journal digest as the last one's output)
        const size_t n_total = n_total_segments ? n_total_segments : n;
        const Journal journal = Journal::encode(session.journal);
        Digest journal_digest;
        ffi_wrap(hfb200_digest_bytes(journal.bytes.data(), journal.bytes.size(), journal_digest.data()));
        std::map<uint32_t, uint32_t> po2_of;
        for (const Segment& s : session.segments) po2_of[s.index] = s.po2;
        std::vector<Digest> states(n_total + 1);
        states[0] = session.has_image_id ? session.image_id : default_image_id();
        for (size_t i = 0; i < n_total; i++) {
            const auto it = po2_of.find((uint32_t)i);
            const uint32_t p2 = it != po2_of.end() ? it->second : (n ? session.segments[0].po2 : 0);
            ffi_wrap(hfb200_claim_next_state(states[i].data(), (uint32_t)i, p2, states[i + 1].data()));
        }
        std::vector<std::array<uint32_t, HFB200_N_GLOBAL>> globals(n);
        for (size_t i = 0; i < n; i++) {
            const Segment& s = session.segments[i];
            if (s.index >= n_total) throw Error("segment index " + std::to_string(s.index) + " outside the session");
            std::memcpy(globals[i].data(), s.globals, sizeof globals[i]);
            hfb200_claim c{};
            const bool last = s.index + 1 == n_total;
            std::memcpy(c.pre, states[s.index].data(), 32); std::memcpy(c.post, states[s.index + 1].data(), 32);
            if (last) std::memcpy(c.output, journal_digest.data(), 32);
            c.exit_code = last ? HFB200_EXIT_HALTED : HFB200_EXIT_SYSTEM_SPLIT;
            ffi_wrap(hfb200_claim_encode(&c, globals[i].data()));
        }
        for (size_t i = 0; i < n; i++) {
            const Segment& s = session.segments[i];
            std::memset(&jobs[i], 0, sizeof jobs[i]);
            jobs[i].po2 = s.po2; jobs[i].globals = globals[i].data(); jobs[i].code = opts_.reuse_control ? nullptr : s.code; jobs[i].data = s.data;
            jobs[i].blind_seed = s.blind_seed; jobs[i].seal_out = seals[i].data(); jobs[i].seal_cap = seal_cap_words;
        }
        const char* e = hfb200_pool_prove(pool_, jobs.data(), n);
        std::string first_err;
        for (size_t i = 0; i < n; i++)
            if (jobs[i].error) { if (first_err.empty()) first_err = jobs[i].error; hfb200_free_error(jobs[i].error); }
        if (e) { if (first_err.empty()) first_err = e; hfb200_free_error(e); }
        if (!first_err.empty()) throw Error(first_err);
        ProveInfo info;
        for (size_t i = 0; i < n; i++) {
            seals[i].resize(jobs[i].seal_words);
            SegmentReceipt sr;
            sr.seal = std::move(seals[i]); sr.index = session.segments[i].index; sr.hashfn = opts_.hashfn;
            info.receipt.segments.push_back(std::move(sr));
            info.devices.push_back(jobs[i].device);
            info.segment_ms.push_back(jobs[i].ms);
        }
        info.receipt.journal = journal;
        return info;
    }

  private:
    ProverOpts opts_;
    hfb200_pool* pool_ = nullptr;
};

inline std::shared_ptr<Prover> default_prover(const ProverOpts& opts = ProverOpts()) { return std::make_shared<Prover>(opts); }

}  // namespace hfb200
