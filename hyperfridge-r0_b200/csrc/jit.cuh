// eval_check of a DATA-DEFINED circuit as a run-time specialised sm_100a kernel.
//
// Upstream turns the constraint IR (`risc0_zkp::adapter::PolyExtStepDef`) into generated C++/CUDA at crate build time
// (risc0-circuit-rv32im-sys `eval_check.cu`; /root/reference/Cargo.lock:3121-3132, not vendored).  Here the circuit
// arrives as data (`hfb200_init_ir`), so the same specialisation happens when the circuit is registered: the
// slot-allocated bytecode of `GenericCircuitHost` is printed as straight-line CUDA (one thread per LDE row, every value a
// register, constraint sums in the lazy 64-bit accumulators of field.cuh), compiled for sm_100a with NVRTC and loaded
// through the runtime's library API.  The interpreter kernel (`GenEvalCheckKernel`) stays as the path for builds or
// boxes without libnvrtc and as the cross-check in the tests (HFB200_IR_JIT=0 selects it).
#pragma once
#include <map>
#include <string>
#include <thread>
#include <atomic>
#include <mutex>
#include "circuit.cuh"

#ifndef HFB200_EMU
#include <dlfcn.h>
#include <nvrtc.h>
#endif

namespace hf {

struct JitEvalArgs {
    const uint32_t* ev[3];   // accum, code, data LDEs [w][domain]
    uint32_t* check;         // [4][domain]
    const E4* mixpow;        // poly_mix^k (device)
    const uint32_t* mix;     // accum mix (device)
    const uint32_t* globals; // device copy of the 32 globals
    uint32_t yinv[4];
    uint32_t po2;
};

static const char* const JIT_PRELUDE = R"JIT(
typedef unsigned int uint32_t;
typedef unsigned long long uint64_t;
#define FI __device__ __forceinline__
static constexpr uint32_t P = 2013265921u, P_INV = 0x88000001u, ONE = 268435454u, NBETA = 1073741848u;
FI uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
FI uint32_t fadd(uint32_t a, uint32_t b) { uint32_t x = a + b; return umin32(x, x - P); }
FI uint32_t fsub(uint32_t a, uint32_t b) { uint32_t x = a - b; return umin32(x, x + P); }
FI uint32_t fmul(uint32_t a, uint32_t b) {
    uint64_t o = (uint64_t)a * b; uint32_t m = (uint32_t)o * P_INV; uint32_t r = (uint32_t)(o >> 32) - __umulhi(m, P); return umin32(r, r + P);
}
struct __align__(16) E4 { uint32_t c[4]; };
FI E4 e4_scale(const E4& a, uint32_t s) { E4 r; r.c[0] = fmul(a.c[0], s); r.c[1] = fmul(a.c[1], s); r.c[2] = fmul(a.c[2], s); r.c[3] = fmul(a.c[3], s); return r; }
FI E4 e4_mul(const E4& a, const E4& b) {
    uint32_t a0 = a.c[0], a1 = a.c[1], a2 = a.c[2], a3 = a.c[3], b0 = b.c[0], b1 = b.c[1], b2 = b.c[2], b3 = b.c[3];
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
struct A64 { uint32_t lo, hi; };
FI void mac(A64& a, uint32_t x, uint32_t y) {
    const uint64_t t = (uint64_t)x * y + (((uint64_t)a.hi << 32) | a.lo);
    const uint32_t h = (uint32_t)(t >> 32); a.lo = (uint32_t)t; a.hi = umin32(h, h - P);
}
FI uint32_t redc(const A64& a) { const uint32_t m = a.lo * P_INV; const uint32_t r = a.hi - __umulhi(m, P); return umin32(r, r + P); }
struct E4A { A64 c[4]; };
FI E4A e4a_zero() { E4A a; for (int k = 0; k < 4; k++) { a.c[k].lo = 0; a.c[k].hi = 0; } return a; }
FI void e4a_mac(E4A& a, const E4& w, uint32_t t) { mac(a.c[0], w.c[0], t); mac(a.c[1], w.c[1], t); mac(a.c[2], w.c[2], t); mac(a.c[3], w.c[3], t); }
FI void e4a_add(E4A& a, const E4& t) { mac(a.c[0], t.c[0], ONE); mac(a.c[1], t.c[1], ONE); mac(a.c[2], t.c[2], ONE); mac(a.c[3], t.c[3], ONE); }
FI E4 e4a_redc(const E4A& a) { E4 r; r.c[0] = redc(a.c[0]); r.c[1] = redc(a.c[1]); r.c[2] = redc(a.c[2]); r.c[3] = redc(a.c[3]); return r; }
struct JitEvalArgs { const uint32_t* ev[3]; uint32_t* check; const E4* mixpow; const uint32_t* mix; const uint32_t* globals; uint32_t yinv[4]; uint32_t po2; };
)JIT";

#ifndef JIT_FENCE_EVERY
#define JIT_FENCE_EVERY 2u
#endif
// straight-line CUDA for the bytecode of chunk `k` of one circuit (kernel hfb200_eval_check_jit_<k>); chunk 0 stores its partial
// sum (times 1 / Z), the others add theirs
// The kernel is specialised for the segment size as well: with `domain` a literal, a tap read is `ev[<column * domain> + row_b]` -- a
// constant offset from four row indices.  With `domain` a run-time value ptxas treats every `column * domain` product as a common
// subexpression worth keeping (hundreds of 64-bit values per chunk): 3 KB of spills per thread and ~50 s of compile time per chunk.
static inline std::string jit_kernel_source(const GenericCircuitHost& g, size_t k, uint32_t po2) {
    const GenericCircuitHost::Chunk& c = g.chunks[k];
    const uint64_t domain = 4ull << po2;
    std::string s;
    // tuning knobs of the generated code (experiments: profiles/r2_ir_scale.txt)
    static const int minb = [] { const char* e = std::getenv("HFB200_IR_JIT_MINB"); const int v = e ? std::atoi(e) : 4; return v >= 1 && v <= 8 ? v : 4; }();
    static const uint32_t fence_every = [] { const char* e = std::getenv("HFB200_IR_JIT_FENCE"); const int v = e ? std::atoi(e) : (int)JIT_FENCE_EVERY; return (uint32_t)(v >= 0 ? v : 2); }();
    s += "extern \"C\" __global__ void __launch_bounds__(128, " + std::to_string(minb) + ") hfb200_eval_check_jit_" + std::to_string(k) + "(JitEvalArgs p) {\n";
    s += "  const uint64_t domain = " + std::to_string(domain) + "ull, dmask = domain - 1ull;\n";
    s += "  if (p.po2 != " + std::to_string(po2) + "u) return;\n";
    s += "  const uint32_t* __restrict__ ev0 = p.ev[0]; const uint32_t* __restrict__ ev1 = p.ev[1]; const uint32_t* __restrict__ ev2 = p.ev[2];\n";
    // grid-stride loop over the rows: the straight-line body is executed many times by every warp, so its instructions come from the
    // instruction cache instead of being streamed from L2 once per warp (what bounds a one-row-per-thread kernel of this size)
    s += "  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < domain; i += (uint64_t)gridDim.x * blockDim.x) {\n";
    for (uint32_t b : g.backs) s += "  const uint64_t r" + std::to_string(b) + " = (i + domain - " + std::to_string(4ull * b) + "ull) & dmask;\n";
    // q<b>: the row indices the loads actually use; made opaque again at every fence so that ptxas cannot keep the ADDRESS of a tap
    // read in a register pair until the same tap is read again hundreds of steps later (that alone was ~450 live values per chunk)
    std::string reopaque;
    for (uint32_t b : g.backs) { s += "  uint64_t q" + std::to_string(b) + " = r" + std::to_string(b) + ";\n"; reopaque += "  q" + std::to_string(b) + " = r" + std::to_string(b) + "; asm volatile(\"\" : \"+l\"(q" + std::to_string(b) + "));\n"; }
    for (uint32_t q = 0; q < c.n_fp_slots; q++) s += "  uint32_t f" + std::to_string(q) + " = 0;\n";
    for (uint32_t q = 0; q < c.n_mix_slots; q++) s += "  E4A m" + std::to_string(q) + " = e4a_zero();\n";
    auto F = [](uint32_t q) { return "f" + std::to_string(q); };
    auto M = [](uint32_t q) { return "m" + std::to_string(q); };
    uint32_t since_fence = 0;
    for (const BcIns& ins : c.prog) {
        const std::string d = std::to_string(ins.dst);
        // a compiler-level memory fence after every few constraints: without it ptxas hoists the (independent) loads of a whole
        // chunk -- hundreds of tap reads and mix powers -- to the top of the kernel and spills them (3 KB of stack per thread,
        // 50 s of compile time per chunk); latency is hidden by the other warps of the SM, not by this thread's own loads
        if (fence_every && (ins.op == BC_MEQZ || ins.op == BC_MCOND) && ++since_fence >= fence_every) { s += "  asm volatile(\"\" ::: \"memory\");\n" + reopaque; since_fence = 0; }
        switch (ins.op) {
            case BC_CONST: s += "  f" + d + " = " + std::to_string(ins.a) + "u;\n"; break;
            case BC_GET:
                s += "  f" + d + " = ev" + std::to_string(ins.a) + "[" + std::to_string((uint64_t)ins.b * domain) + "ull + q" + std::to_string(ins.c) + "];\n";
                break;
            case BC_GETG: s += "  f" + d + " = " + (ins.a == 0 ? "p.globals[" : "p.mix[") + std::to_string(ins.b) + "];\n"; break;
            case BC_ADD: s += "  f" + d + " = fadd(" + F(ins.a) + ", " + F(ins.b) + ");\n"; break;
            case BC_SUB: s += "  f" + d + " = fsub(" + F(ins.a) + ", " + F(ins.b) + ");\n"; break;
            case BC_MUL: s += "  f" + d + " = fmul(" + F(ins.a) + ", " + F(ins.b) + ");\n"; break;
            case BC_MTRUE: s += "  m" + d + " = e4a_zero();\n"; break;
            case BC_MEQZ:
                if (ins.dst != ins.a) s += "  m" + d + " = " + M(ins.a) + ";\n";
                s += "  e4a_mac(m" + d + ", p.mixpow[" + std::to_string(ins.c) + "], " + F(ins.b) + ");\n";
                break;
            default:  // BC_MCOND: tot = x.tot + cond * inner.tot * mix^k
                s += "  { const E4 t_ = e4_scale(e4_mul(e4a_redc(" + M(ins.b >> 16) + "), p.mixpow[" + std::to_string(ins.c) + "]), " + F(ins.b & 0xFFFFu) + "); E4A x_ = " + M(ins.a) +
                     "; e4a_add(x_, t_); m" + d + " = x_; }\n";
                break;
        }
    }
    s += "  const E4 tot = e4a_redc(m" + std::to_string(c.ret_slot) + ");\n  const uint32_t yi = p.yinv[i & 3];\n";
    if (k == 0) s += "  for (int k = 0; k < 4; k++) p.check[(uint64_t)k * domain + i] = fmul(tot.c[k], yi);\n  }\n}\n";
    else s += "  for (int k = 0; k < 4; k++) p.check[(uint64_t)k * domain + i] = fadd(p.check[(uint64_t)k * domain + i], fmul(tot.c[k], yi));\n  }\n}\n";
    return s;
}
// every chunk kernel in one translation unit (hfb200_ir_source; the library itself compiles one unit per chunk, in parallel)
static inline std::string jit_source(const GenericCircuitHost& g, uint32_t po2 = 20) {
    std::string s = JIT_PRELUDE;
    for (size_t k = 0; k < g.chunks.size(); k++) s += jit_kernel_source(g, k, po2);
    return s;
}

#ifndef HFB200_EMU
struct NvrtcApi {
    void* h = nullptr;
    nvrtcResult (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*DestroyProgram)(nvrtcProgram*) = nullptr;
    bool load() {
        if (h) return true;
        const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so"};
        for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_LOCAL); if (h) break; }
        if (!h) return false;
        bool ok = true;
        auto sym = [&](auto& fn, const char* name) { fn = reinterpret_cast<std::decay_t<decltype(fn)>>(dlsym(h, name)); ok = ok && fn != nullptr; };
        sym(CreateProgram, "nvrtcCreateProgram"); sym(CompileProgram, "nvrtcCompileProgram"); sym(GetCUBINSize, "nvrtcGetCUBINSize");
        sym(GetCUBIN, "nvrtcGetCUBIN"); sym(GetProgramLogSize, "nvrtcGetProgramLogSize"); sym(GetProgramLog, "nvrtcGetProgramLog");
        sym(DestroyProgram, "nvrtcDestroyProgram");
        if (!ok) { dlclose(h); h = nullptr; }
        return ok;
    }
};

struct JitEvalCheck {
    struct Compiled { std::vector<cudaLibrary_t> libs; std::vector<cudaKernel_t> kernels; };  // one kernel per chunk, launched in order
    std::map<uint32_t, Compiled> by_po2;  // specialised per segment size, compiled when a size is first proved (or at init for max_po2)
    const GenericCircuitHost* g = nullptr;
    bool ready = false;
    std::string note;   // why the JIT is not in use (empty when ready)
    float compile_ms = 0;  // the most recent specialisation

    static NvrtcApi& api() { static NvrtcApi a; return a; }
    void init(const GenericCircuitHost& gen, uint32_t first_po2) {
        if (const char* env = std::getenv("HFB200_IR_JIT")) if (std::atoi(env) == 0) { note = "disabled by HFB200_IR_JIT=0"; return; }
        static std::mutex api_mu;
        { std::lock_guard<std::mutex> lock(api_mu); if (!api().load()) { note = "libnvrtc.so.12 not found"; return; } }
        g = &gen;
        ready = true;
        ensure(first_po2);
    }
    void ensure(uint32_t po2) {
        if (!ready || by_po2.count(po2)) return;
        const auto t0 = std::chrono::steady_clock::now();
        const size_t K = g->chunks.size();
        std::vector<std::vector<char>> cubins(K);
        std::vector<std::string> errors(K);
        // one NVRTC program per chunk, compiled on up to 8 host threads (NVRTC is thread-safe across programs)
        std::atomic<size_t> next{0};
        auto work = [&] {
            NvrtcApi& a = api();
            for (;;) {
                const size_t k = next.fetch_add(1);
                if (k >= K) return;
                const std::string src = std::string(JIT_PRELUDE) + jit_kernel_source(*g, k, po2);
                nvrtcProgram prog = nullptr;
                if (a.CreateProgram(&prog, src.c_str(), "hfb200_eval_check_jit.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) { errors[k] = "nvrtcCreateProgram failed"; continue; }
                const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", "-lineinfo"};
                if (a.CompileProgram(prog, 3, opts) != NVRTC_SUCCESS) {
                    size_t n = 0; a.GetProgramLogSize(prog, &n);
                    std::string log(n, '\0'); if (n) a.GetProgramLog(prog, &log[0]);
                    a.DestroyProgram(&prog);
                    errors[k] = "NVRTC compilation of eval_check chunk " + std::to_string(k) + " failed: " + log.substr(0, 2000);
                    continue;
                }
                size_t nb = 0; a.GetCUBINSize(prog, &nb);
                cubins[k].resize(nb);
                a.GetCUBIN(prog, cubins[k].data());
                a.DestroyProgram(&prog);
            }
        };
        unsigned nthreads = std::thread::hardware_concurrency();
        if (nthreads == 0) nthreads = 1;
        nthreads = (unsigned)std::min<size_t>(std::min<unsigned>(nthreads, 8u), K);
        if (const char* env = std::getenv("HFB200_IR_JIT_THREADS")) { const int v = std::atoi(env); if (v >= 1) nthreads = (unsigned)v; }
        std::vector<std::thread> th;
        for (unsigned t = 1; t < nthreads; t++) th.emplace_back(work);
        work();
        for (auto& t : th) t.join();
        for (const std::string& e : errors) if (!e.empty()) throw Err("data-defined circuit: " + e);
        Compiled c;
        c.libs.resize(K, nullptr); c.kernels.resize(K, nullptr);
        for (size_t k = 0; k < K; k++) {
            CUDA_CHECK(cudaLibraryLoadData(&c.libs[k], cubins[k].data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
            const std::string name = "hfb200_eval_check_jit_" + std::to_string(k);
            CUDA_CHECK(cudaLibraryGetKernel(&c.kernels[k], c.libs[k], name.c_str()));
        }
        by_po2[po2] = std::move(c);
        compile_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    void destroy() { for (auto& kv : by_po2) for (auto& l : kv.second.libs) if (l) cudaLibraryUnload(l); by_po2.clear(); ready = false; }
    void launch(Dev& dev, const JitEvalArgs& a, uint64_t domain) {
        ensure(a.po2);
        JitEvalArgs args = a;
        void* params[] = {&args};
        static const unsigned per_sm = [] { const char* e = std::getenv("HFB200_IR_JIT_BLOCKS_PER_SM"); const int v = e ? std::atoi(e) : 8; return (unsigned)(v >= 1 ? v : 8); }();
        const unsigned block = 128;
        const unsigned grid = (unsigned)std::min<uint64_t>((domain + block - 1) / block, (uint64_t)dev.sm_count * per_sm);
        for (cudaKernel_t k : by_po2.at(a.po2).kernels) {
            CUDA_CHECK(cudaLaunchKernel((const void*)k, dim3(grid), dim3(block), params, 0, dev.stream));
            dev.launches++;
        }
    }
};
#else
struct JitEvalCheck {
    bool ready = false;
    std::string note = "host emulator build";
    float compile_ms = 0;
    void init(const GenericCircuitHost&, uint32_t) {}
    void destroy() {}
    void launch(Dev&, const JitEvalArgs&, uint64_t) {}
};
#endif

}  // namespace hf
