// C ABI of libhfb200.so (see include/hfb200.h for the boundary contract and reference citations).
#include "../../include/hfb200.h"
#include "prover.cuh"
#include "probe.cuh"
#include "verify.cuh"
#include <algorithm>
#include <atomic>
#include <map>
#include <memory>
#include <thread>

using namespace hf;

struct hfb200_ctx {
    Prover p;
    std::vector<uint32_t> seal;
#ifndef HFB200_EMU
    cudaEvent_t marks[4] = {nullptr, nullptr, nullptr, nullptr};
#endif
};

static const char* dup_err(const std::string& s) {
    char* m = (char*)std::malloc(s.size() + 1);
    if (!m) return "hfb200: out of memory while reporting an error";
    std::memcpy(m, s.c_str(), s.size() + 1);
    return m;
}
#define API_TRY try {
#define API_CATCH } catch (const std::exception& e) { return dup_err(e.what()); } catch (...) { return dup_err("hfb200: unknown error"); } return nullptr;

static const char* emit_seal(hfb200_ctx* ctx, uint32_t* seal_out, size_t seal_cap, size_t* seal_words) {
    if (seal_words) *seal_words = ctx->seal.size();
    if (seal_cap < ctx->seal.size() || !seal_out) return dup_err("seal buffer too small: need " + std::to_string(ctx->seal.size()) + " words");
    std::memcpy(seal_out, ctx->seal.data(), ctx->seal.size() * 4);
    return nullptr;
}

extern "C" {

const char* hfb200_version(void) {
#ifdef HFB200_EMU
    return "hfb200 0.1 (HOST EMULATOR - test build, not a product)";
#else
    return "hfb200 0.1 (sm_100a)";
#endif
}

void hfb200_free_error(const char* msg) { std::free((void*)msg); }

const char* hfb200_init(int device, uint32_t max_po2, const hfb200_circuit_desc* c, hfb200_ctx** out) {
    API_TRY
    if (!out) throw Err("hfb200_init: out is NULL");
    *out = nullptr;
    hfb200_circuit_desc d = c ? *c : hfb200_circuit_desc{16, 192, 48, 0};
    hfb200_ctx* ctx = new hfb200_ctx();
    try { ctx->p.init(device, max_po2, d.w_code, d.w_data, d.w_accum); } catch (...) { delete ctx; throw; }
    *out = ctx;
    API_CATCH
}
const char* hfb200_init_ir(int device, uint32_t max_po2, const hfb200_circuit_ir* c, hfb200_ctx** out) {
    API_TRY
    if (!out || !c) throw Err("hfb200_init_ir: NULL argument");
    *out = nullptr;
    static_assert(sizeof(hfb200_tap) == sizeof(IrTap) && sizeof(hfb200_poly_step) == sizeof(IrStep), "IR layouts must match");
    hfb200_ctx* ctx = new hfb200_ctx();
    try {
        ctx->p.init(device, max_po2, c->w_code, c->w_data, c->w_accum, reinterpret_cast<const IrTap*>(c->taps), c->n_taps,
                    reinterpret_cast<const IrStep*>(c->steps), c->n_steps, c->ret, c->n_mix);
    } catch (...) { delete ctx; throw; }
    *out = ctx;
    API_CATCH
}
const char* hfb200_ir_source(const hfb200_circuit_ir* c, char* out, size_t cap, size_t* need) {
    API_TRY
    if (!c) throw Err("hfb200_ir_source: NULL circuit");
    GenericCircuitHost g;
    g.init(nullptr, c->w_code, c->w_data, c->w_accum, c->n_mix, reinterpret_cast<const IrTap*>(c->taps), c->n_taps,
           reinterpret_cast<const IrStep*>(c->steps), c->n_steps, c->ret);
    const std::string src = jit_source(g);
    if (need) *need = src.size() + 1;
    if (out && cap) { const size_t n = std::min(cap - 1, src.size()); std::memcpy(out, src.data(), n); out[n] = 0; }
    API_CATCH
}
int hfb200_ir_jit_active(const hfb200_ctx* ctx, float* compile_ms) {
    if (!ctx) return 0;
    if (compile_ms) *compile_ms = ctx->p.jit.compile_ms;
    return ctx->p.jit.ready ? 1 : 0;
}
void hfb200_destroy(hfb200_ctx* ctx) {
    if (!ctx) return;
#ifndef HFB200_EMU
    for (auto& m : ctx->marks) if (m) { cudaEventDestroy(m); m = nullptr; }
#endif
    try { ctx->p.destroy(); } catch (...) {}
    delete ctx;
}

const char* hfb200_host_alloc(size_t bytes, void** out) {
    API_TRY
#ifndef HFB200_EMU
    CUDA_CHECK(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
#else
    *out = std::malloc(bytes);
#endif
    API_CATCH
}
void hfb200_host_free(void* p) {
#ifndef HFB200_EMU
    cudaFreeHost(p);
#else
    std::free(p);
#endif
}

const char* hfb200_prove_segment(hfb200_ctx* ctx, uint32_t po2, const uint32_t* globals, const uint32_t* code, const uint32_t* data,
                                 uint64_t blind_seed, uint32_t* seal_out, size_t seal_cap, size_t* seal_words) {
    API_TRY
    if (!ctx || !globals || !data) throw Err("hfb200_prove_segment: NULL argument");
    if (!code && !ctx->p.control_cached) throw Err("hfb200_prove_segment: code is NULL and no control group is loaded (hfb200_control_root)");
    ctx->p.begin(po2, globals, code, data, blind_seed);
    ctx->p.finish(nullptr, ctx->seal);
    if (const char* e = emit_seal(ctx, seal_out, seal_cap, seal_words)) return e;
    API_CATCH
}

const char* hfb200_segment_begin(hfb200_ctx* ctx, uint32_t po2, const uint32_t* globals, const uint32_t* code, const uint32_t* data,
                                 uint64_t blind_seed, uint32_t* mix_out, size_t mix_cap, size_t* mix_words) {
    API_TRY
    if (!ctx || !globals) throw Err("hfb200_segment_begin: NULL argument");
    ctx->p.begin(po2, globals, code, data, blind_seed);
    if (mix_words) *mix_words = ctx->p.mix.size();
    if (mix_out) {
        if (mix_cap < ctx->p.mix.size()) throw Err("mix buffer too small");
        std::memcpy(mix_out, ctx->p.mix.data(), ctx->p.mix.size() * 4);
    }
    API_CATCH
}
const char* hfb200_segment_finish(hfb200_ctx* ctx, const uint32_t* accum_or_null, uint32_t* seal_out, size_t seal_cap, size_t* seal_words) {
    API_TRY
    if (!ctx) throw Err("hfb200_segment_finish: NULL ctx");
    ctx->p.finish(accum_or_null, ctx->seal);
    if (const char* e = emit_seal(ctx, seal_out, seal_cap, seal_words)) return e;
    API_CATCH
}

const char* hfb200_witgen_synth(hfb200_ctx* ctx, uint32_t po2, uint64_t trace_seed, uint64_t blind_seed, uint32_t* globals_out) {
    API_TRY
    if (!ctx || !globals_out) throw Err("hfb200_witgen_synth: NULL argument");
    ctx->p.witgen(po2, trace_seed, blind_seed, globals_out);
    API_CATCH
}
const char* hfb200_prove_resident(hfb200_ctx* ctx, uint64_t blind_seed, uint32_t* seal_out, size_t seal_cap, size_t* seal_words) {
    API_TRY
    if (!ctx) throw Err("hfb200_prove_resident: NULL ctx");
    if (!ctx->p.have_trace) throw Err("hfb200_prove_resident: no resident trace (call hfb200_witgen_synth)");
    uint32_t g[N_GLOBAL];
    std::memcpy(g, ctx->p.globals, sizeof g);
    ctx->p.begin(ctx->p.po2, g, nullptr, nullptr, blind_seed);
    ctx->p.finish(nullptr, ctx->seal);
    if (const char* e = emit_seal(ctx, seal_out, seal_cap, seal_words)) return e;
    API_CATCH
}
const char* hfb200_read_group(hfb200_ctx* ctx, uint32_t group, uint32_t* out, size_t cap_words) {
    API_TRY
    if (!ctx || group > 2 || !ctx->p.tr[group]) throw Err("hfb200_read_group: bad argument");
    ctx->p.bind();
    const size_t words = (size_t)ctx->p.cir.group_width((int)group) << ctx->p.po2;
    if (cap_words < words) throw Err("hfb200_read_group: buffer too small");
    ctx->p.dev.d2h(out, ctx->p.tr[group], words * 4);
    ctx->p.dev.sync();
    API_CATCH
}

size_t hfb200_seal_words(const hfb200_ctx* ctx, uint32_t po2) { return ctx ? ctx->p.seal_words(po2) : 0; }

const char* hfb200_checkpoint(hfb200_ctx* ctx, const char* name, uint32_t* out, size_t cap, size_t* n_words) {
    API_TRY
    if (!ctx || !name) throw Err("hfb200_checkpoint: NULL argument");
    for (auto& kv : ctx->p.cps)
        if (kv.first == name) {
            if (n_words) *n_words = kv.second.size();
            if (out) { if (cap < kv.second.size()) throw Err("checkpoint buffer too small"); std::memcpy(out, kv.second.data(), kv.second.size() * 4); }
            return nullptr;
        }
    throw Err(std::string("no such checkpoint: ") + name);
    API_CATCH
}

const char* hfb200_last_stats(hfb200_ctx* ctx, hfb200_stats* out) {
    API_TRY
    if (!ctx || !out) throw Err("hfb200_last_stats: NULL argument");
    const Stats& s = ctx->p.stats;
    out->ms_total = s.ms_total; out->ms_device = s.ms_device; out->ms_h2d = s.ms_h2d; out->ms_ntt_main = s.ms_ntt_main; out->ms_hash_main = s.ms_hash_main;
    out->ms_accum = s.ms_accum; out->ms_check = s.ms_check; out->ms_deep = s.ms_deep; out->ms_fri = s.ms_fri;
    out->launches = s.launches; out->ntt_main_bytes = s.ntt_main_bytes;
    API_CATCH
}
uint64_t hfb200_total_launches(const hfb200_ctx* ctx) { return ctx ? ctx->p.dev.launches : 0; }

// ---- multi-GPU pool ---------------------------------------------------------------------------------
} // extern "C"
struct hfb200_pool {
    std::vector<std::unique_ptr<hfb200_ctx>> ctxs;
    std::vector<int> device_of;
    std::map<uint32_t, const uint32_t*> control;  // po2 -> caller-owned control columns (hfb200_pool_load_control)
};
extern "C" {

const char* hfb200_pool_create(const int* devices, int n_devices, int contexts_per_device, uint32_t max_po2,
                               const hfb200_circuit_desc* c, hfb200_pool** out) {
    API_TRY
    if (!out || !devices || n_devices <= 0 || contexts_per_device <= 0) throw Err("hfb200_pool_create: bad argument");
    *out = nullptr;
    hfb200_circuit_desc d = c ? *c : hfb200_circuit_desc{16, 192, 48, 0};
    std::unique_ptr<hfb200_pool> pool(new hfb200_pool());
    // contexts are created on their worker threads' devices up front (cudaSetDevice is per thread, init is serial here)
    for (int i = 0; i < n_devices; i++)
        for (int s_ = 0; s_ < contexts_per_device; s_++) {
            std::unique_ptr<hfb200_ctx> ctx(new hfb200_ctx());
            ctx->p.init(devices[i], max_po2, d.w_code, d.w_data, d.w_accum);
            pool->ctxs.push_back(std::move(ctx));
            pool->device_of.push_back(devices[i]);
        }
    *out = pool.release();
    API_CATCH
}

const char* hfb200_pool_prove(hfb200_pool* pool, hfb200_segment_job* jobs, size_t n_jobs) {
    API_TRY
    if (!pool || (!jobs && n_jobs)) throw Err("hfb200_pool_prove: bad argument");
    std::vector<size_t> order(n_jobs);
    for (size_t i = 0; i < n_jobs; i++) { order[i] = i; jobs[i].error = nullptr; jobs[i].seal_words = 0; jobs[i].device = -1; jobs[i].ms = 0; }
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return jobs[a].po2 > jobs[b].po2; });  // longest first
    std::atomic<size_t> next{0};
    auto worker = [&](size_t w) {
        hfb200_ctx* ctx = pool->ctxs[w].get();
#ifndef HFB200_EMU
        cudaSetDevice(pool->device_of[w]);
#endif
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= n_jobs) break;
            hfb200_segment_job& j = jobs[order[k]];
            const auto t0 = std::chrono::steady_clock::now();
            if (!j.code) {
                // shared control group: (re)commit it on this context when it is not the one resident for this po2
                const auto it = pool->control.find(j.po2);
                if (it == pool->control.end()) { j.error = dup_err("code is NULL and no control group was loaded for this po2 (hfb200_pool_load_control)"); continue; }
                if (!(ctx->p.control_cached && ctx->p.po2 == j.po2)) {
                    uint32_t root[8];
                    if ((j.error = hfb200_control_root(ctx, j.po2, it->second, root))) continue;
                }
            }
            j.error = hfb200_prove_segment(ctx, j.po2, j.globals, j.code, j.data, j.blind_seed, j.seal_out, j.seal_cap, &j.seal_words);
            j.device = pool->device_of[w];
            j.ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        }
    };
    std::vector<std::thread> th;
    for (size_t w = 0; w < pool->ctxs.size(); w++) th.emplace_back(worker, w);
    for (auto& t : th) t.join();
    for (size_t i = 0; i < n_jobs; i++)
        if (jobs[i].error) return dup_err(std::string("job ") + std::to_string(i) + ": " + jobs[i].error);
    API_CATCH
}

const char* hfb200_pool_load_control(hfb200_pool* pool, uint32_t po2, const uint32_t* code) {
    API_TRY
    if (!pool) throw Err("hfb200_pool_load_control: NULL pool");
    if (code) pool->control[po2] = code; else pool->control.erase(po2);
    API_CATCH
}

void hfb200_pool_destroy(hfb200_pool* pool) {
    if (!pool) return;
    for (size_t w = 0; w < pool->ctxs.size(); w++) {
#ifndef HFB200_EMU
        cudaSetDevice(pool->device_of[w]);
#endif
        try { pool->ctxs[w]->p.destroy(); } catch (...) {}
    }
    delete pool;
}

// ---- HAL-level operators ------------------------------------------------------------------------
struct DevBuf {
    Dev& d; uint32_t* p;
    DevBuf(Dev& dev, size_t words) : d(dev), p((uint32_t*)dev.alloc(words * 4)) {}
    ~DevBuf() { d.free(p); }
};

const char* hfb200_op_interpolate_ntt(hfb200_ctx* ctx, uint32_t* io, size_t count, size_t n, int zk_shift) {
    API_TRY
    ctx->p.bind();
    Dev& d = ctx->p.dev;
    const int lg = ilog2(n);
    if ((1ull << lg) != n || lg > 24 || count == 0) throw Err("op_interpolate_ntt: bad size");
    DevBuf b(d, count * n);
    d.h2d(b.p, io, count * n * 4);
    ctx->p.ntt.interpolate(b.p, n, b.p, n, (uint32_t)count, lg, zk_shift != 0);
    d.d2h(io, b.p, count * n * 4);
    d.sync();
    API_CATCH
}
const char* hfb200_op_expand_ntt(hfb200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, size_t n_in, uint32_t expand_bits) {
    API_TRY
    ctx->p.bind();
    Dev& d = ctx->p.dev;
    const int lg = ilog2(n_in);
    if ((1ull << lg) != n_in || lg + (int)expand_bits > 24 || (expand_bits != 0 && expand_bits != 2) || count == 0) throw Err("op_expand_ntt: bad size");
    DevBuf bi(d, count * n_in), bo(d, (count * n_in) << expand_bits);
    d.h2d(bi.p, in, count * n_in * 4);
    ctx->p.ntt.expand_evaluate(bi.p, n_in, bo.p, n_in << expand_bits, (uint32_t)count, lg, (int)expand_bits);
    d.d2h(out, bo.p, ((count * n_in) << expand_bits) * 4);
    d.sync();
    API_CATCH
}
const char* hfb200_op_lde(hfb200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, size_t n_in) {
    API_TRY
    ctx->p.bind();
    Dev& d = ctx->p.dev;
    const int lg = ilog2(n_in);
    if ((1ull << lg) != n_in || lg > 22 || count == 0) throw Err("op_lde: bad size");
    DevBuf bi(d, count * n_in), bs(d, count * n_in), bo(d, count * n_in * 4);
    d.h2d(bi.p, in, count * n_in * 4);
    ctx->p.ntt.lde(bi.p, n_in, bo.p, n_in * 4, bs.p, (uint32_t)count, lg);
    d.d2h(out, bo.p, count * n_in * 16);
    d.sync();
    API_CATCH
}
const char* hfb200_op_merkle(hfb200_ctx* ctx, const uint32_t* matrix, size_t rows, size_t cols, uint32_t* nodes_out) {
    API_TRY
    ctx->p.bind();
    Dev& d = ctx->p.dev;
    if ((1ull << ilog2(rows)) != rows || rows < 2 || cols == 0) throw Err("op_merkle: rows must be a power of two >= 2");
    DevBuf bm(d, rows * cols), bn(d, 2 * rows * 8);
    d.h2d(bm.p, matrix, rows * cols * 4);
    d.zero(bn.p, 8 * 4);
    ctx->p.merkle.build(bm.p, rows, (uint32_t)rows, (uint32_t)cols, bn.p);
    d.d2h(nodes_out, bn.p, 2 * rows * 8 * 4);
    d.sync();
    API_CATCH
}
const char* hfb200_op_poseidon2(hfb200_ctx* ctx, uint32_t* states, size_t n) {
    API_TRY
    ctx->p.bind();
    Dev& d = ctx->p.dev;
    DevBuf b(d, n * 24);
    d.h2d(b.p, states, n * 24 * 4);
    d.launch<PermuteKernel, 128, 1>((unsigned)((n + 127) / 128), 1, 128, 0, b.p, (uint32_t)n);
    d.d2h(states, b.p, n * 24 * 4);
    d.sync();
    API_CATCH
}
const char* hfb200_op_fri_fold(hfb200_ctx* ctx, uint32_t* out, const uint32_t* in, size_t n, const uint32_t* mix4) {
    API_TRY
    ctx->p.bind();
    Dev& d = ctx->p.dev;
    if (n < 16 || (n & 15)) throw Err("op_fri_fold: n must be a multiple of 16");
    DevBuf bi(d, 4 * n), bo(d, 4 * n / 16);
    d.h2d(bi.p, in, 4 * n * 4);
    FriFoldArgs fa{};
    fa.in = bi.p; fa.out = bo.p; fa.n = (uint32_t)n;
    const E4 fm = e4(mix4[0], mix4[1], mix4[2], mix4[3]);
    { E4 cur = e4_one(); for (int i = 0; i < 16; i++) { fa.mixpow[i] = cur; cur = e4_mul(cur, fm); } }
    d.launch<FriFoldKernel, 256, 1>((unsigned)((n / 16 + 255) / 256), 1, 256, 0, fa);
    d.d2h(out, bo.p, (4 * n / 16) * 4);
    d.sync();
    API_CATCH
}

const char* hfb200_bench_lde(hfb200_ctx* ctx, uint32_t po2, uint32_t count, uint32_t iters, float* ms_avg) {
    API_TRY
    Prover& p = ctx->p;
    p.bind();
    p.layout(po2);
    if (count == 0 || count > p.cir.cd.w_data) throw Err("bench_lde: count must be in [1, w_data]");
    if (!p.have_trace) throw Err("bench_lde: call hfb200_witgen_synth first");
    const size_t N = (size_t)1 << po2;
    Timer t; t.init(p.dev.stream);
    p.ntt.lde(p.tr[GROUP_DATA], N, p.ev[GROUP_DATA], 4 * N, p.scratch, count, (int)po2);  // warm-up
    t.start();
    for (uint32_t i = 0; i < iters; i++) p.ntt.lde(p.tr[GROUP_DATA], N, p.ev[GROUP_DATA], 4 * N, p.scratch, count, (int)po2);
    t.stop();
    *ms_avg = t.ms() / (iters ? iters : 1);
    t.destroy();
    API_CATCH
}
const char* hfb200_bench_merkle(hfb200_ctx* ctx, uint32_t po2, uint32_t count, uint32_t iters, float* ms_avg) {
    API_TRY
    Prover& p = ctx->p;
    p.bind();
    p.layout(po2);
    if (count == 0 || count > p.cir.cd.w_data) throw Err("bench_merkle: count must be in [1, w_data]");
    const size_t D = (size_t)4 << po2;
    Timer t; t.init(p.dev.stream);
    p.merkle.build(p.ev[GROUP_DATA], D, (uint32_t)D, count, p.nodes[GROUP_DATA]);
    t.start();
    for (uint32_t i = 0; i < iters; i++) p.merkle.build(p.ev[GROUP_DATA], D, (uint32_t)D, count, p.nodes[GROUP_DATA]);
    t.stop();
    *ms_avg = t.ms() / (iters ? iters : 1);
    t.destroy();
    API_CATCH
}

const char* hfb200_verify_segment(const hfb200_circuit_desc* c, const hfb200_circuit_ir* ir, const uint32_t* seal, size_t seal_words,
                                  const uint32_t* code_root, uint32_t* po2_out) {
    API_TRY
    if ((c == nullptr) == (ir == nullptr)) throw Err("hfb200_verify_segment: pass exactly one of circuit / ir");
    if (!seal || !code_root) throw Err("hfb200_verify_segment: NULL seal or code_root");
    VCircuit vc;
    if (c) vc.init_builtin(c->w_code, c->w_data, c->w_accum);
    else vc.init_ir(ir->w_code, ir->w_data, ir->w_accum, ir->n_mix, reinterpret_cast<const IrTap*>(ir->taps), ir->n_taps,
                    reinterpret_cast<const IrStep*>(ir->steps), ir->n_steps, ir->ret);
    verify_segment(vc, seal, seal_words, code_root, po2_out);
    API_CATCH
}
const char* hfb200_control_root(hfb200_ctx* ctx, uint32_t po2, const uint32_t* code, uint32_t* root_out) {
    API_TRY
    if (!ctx || !code || !root_out) throw Err("hfb200_control_root: NULL argument");
    Prover& p = ctx->p;
    if (p.begun) throw Err("hfb200_control_root: a segment is in flight (call hfb200_segment_finish first)");
    p.bind();
    p.layout(po2);
    const size_t N = (size_t)1 << po2, D = 4 * N;
    const uint32_t w = p.cir.group_width(GROUP_CODE);
    p.dev.h2d(p.tr[GROUP_CODE], code, (size_t)w * N * 4);
    p.ntt.lde(p.tr[GROUP_CODE], N, p.ev[GROUP_CODE], D, p.scratch, w, (int)po2);
    p.merkle.build(p.ev[GROUP_CODE], D, (uint32_t)D, w, p.nodes[GROUP_CODE]);
    // keep the committed control group for the segments that follow with code == NULL
    p.commit_tree(Tree{p.ev[GROUP_CODE], D, (uint32_t)D, w, p.nodes[GROUP_CODE]}, "code_root", &p.control_top);
    p.proof.clear(); p.cps.clear(); p.rng = HostRng();  // commit_tree wrote into the transcript of no segment
    std::memcpy(root_out, &p.control_top[8], 32);       // heap layout: node 1 is the root
    p.control_cached = true;
    p.have_trace = false;  // the resident trace (if any) lost its code group
    API_CATCH
}
const char* hfb200_bench_modmul(hfb200_ctx* ctx, int kind, uint32_t iters, double* products_per_s) {
    API_TRY
    if (!ctx || !products_per_s) throw Err("bench_modmul: NULL argument");
    if (kind < 0 || kind > 2) throw Err("bench_modmul: kind must be 0 (Montgomery), 1 (Shoup) or 2 (S-box chain)");
    Prover& p = ctx->p;
    p.bind();
    const unsigned blocks = (unsigned)p.dev.sm_count * 4, threads = 256;  // 8 warps per SM sub-partition
    uint32_t* out = (uint32_t*)p.dev.alloc((size_t)blocks * threads * 4);
    Timer t; t.init(p.dev.stream);
    p.dev.launch<ModmulProbeKernel, 256, 1>(blocks, 1, threads, 0, out, 12345u, iters, kind);  // warm-up (clocks)
    t.start();
    p.dev.launch<ModmulProbeKernel, 256, 1>(blocks, 1, threads, 0, out, 12345u, iters, kind);
    t.stop();
    const float ms = t.ms();
    t.destroy();
    p.dev.free(out);
    const double products = (double)blocks * threads * ModmulProbeKernel::ILP * (double)iters * (kind == 2 ? 4.0 : 1.0);
    *products_per_s = ms > 0 ? products / (ms * 1e-3) : 0.0;
    API_CATCH
}
const char* hfb200_mark(hfb200_ctx* ctx, int slot) {
    API_TRY
    if (!ctx || slot < 0 || slot > 3) throw Err("hfb200_mark: bad argument");
#ifndef HFB200_EMU
    ctx->p.bind();
    if (!ctx->marks[slot]) CUDA_CHECK(cudaEventCreate(&ctx->marks[slot]));
    CUDA_CHECK(cudaEventRecord(ctx->marks[slot], ctx->p.dev.stream));
#endif
    API_CATCH
}
const char* hfb200_mark_elapsed(hfb200_ctx* a, int slot_a, hfb200_ctx* b, int slot_b, float* ms) {
    API_TRY
    if (!a || !b || !ms || slot_a < 0 || slot_a > 3 || slot_b < 0 || slot_b > 3) throw Err("hfb200_mark_elapsed: bad argument");
    *ms = 0.f;
#ifndef HFB200_EMU
    if (!a->marks[slot_a] || !b->marks[slot_b]) throw Err("hfb200_mark_elapsed: mark not recorded");
    a->p.bind();
    CUDA_CHECK(cudaEventSynchronize(a->marks[slot_a]));
    CUDA_CHECK(cudaEventSynchronize(b->marks[slot_b]));
    CUDA_CHECK(cudaEventElapsedTime(ms, a->marks[slot_a], b->marks[slot_b]));
#endif
    API_CATCH
}

}  // extern "C"
