// Launch / memory plumbing shared by every kernel of libhfb200.
//
// Kernels are written once as `Body::run(const KCtx&, uint32_t* smem, args...)` functors:
//   * real build (nvcc, sm_100a): `kernel_entry<Body, ...>` is the __global__ wrapper, launched on the
//     context's stream; this is the only thing libhfb200.so contains.
//   * -DHFB200_EMU (tests/emu only, never part of libhfb200.so): the same Body is executed on the host,
//     block by block, so index arithmetic / twiddle schedules can be checked in the CPU-only test tier
//     before GPU minutes are spent.  Barrier kernels (kBarrier = true) are written as work-item loops
//     `for (i = cx.tid; i < n; i += cx.nt)` separated by cx.sync(), which the emulator runs with nt = 1.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <stdexcept>
#include <vector>
#include <atomic>
#include <mutex>
#include "field.cuh"

namespace hf {

struct KCtx {
    int tid, nt;              // thread index / threads per block (or per unit, see unit_ctx)
    unsigned bx, by, gx, gy;  // block index / grid size
    int bar_id = 0;           // 0: whole-CTA barrier; k > 0: named barrier k over `nt` threads (a "unit")
    HD void sync() const {
#ifdef __CUDA_ARCH__
        if (bar_id && nt == 32) __syncwarp();  // a one-warp unit needs no hardware barrier
        else if (bar_id) asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nt) : "memory");
        else __syncthreads();
#endif
    }
};

// A CTA may be split into independent "units" of UT threads (each with its own named barrier) that work on
// different columns but share per-CTA tables.  Device: unit = tid / UT, one pass.  Emulator (nt = 1): the units
// are simply visited one after the other.
HD int unit_first(const KCtx& cx, int UT) {
#ifdef __CUDA_ARCH__
    return cx.tid / UT;
#else
    (void)cx; (void)UT; return 0;
#endif
}
HD int unit_step(const KCtx& cx, int UT) {
#ifdef __CUDA_ARCH__
    return cx.nt / UT;
#else
    (void)cx; (void)UT; return 1;
#endif
}
HD KCtx unit_ctx(const KCtx& cx, int unit, int UT) {
#ifdef __CUDA_ARCH__
    KCtx u{cx.tid % UT, UT, cx.bx, cx.by, cx.gx, cx.gy, unit + 1};
#else
    (void)unit; (void)UT;
    KCtx u{0, 1, cx.bx, cx.by, cx.gx, cx.gy, 0};
#endif
    return u;
}

// ---- TMA bulk copies (cp.async.bulk, global -> shared, completion on an mbarrier) -------------------------------------
// Device only; the host emulator takes plain loops instead (callers switch on __CUDA_ARCH__).
#if defined(__CUDACC__) && !defined(HFB200_EMU)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// size and both addresses are multiples of 16 bytes
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#endif

struct Err : std::runtime_error { using std::runtime_error::runtime_error; };

#ifndef HFB200_EMU
// ------------------------------------------------------------------ real CUDA build
#define CUDA_CHECK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { char b_[512]; std::snprintf(b_, sizeof b_, "CUDA error %s at %s:%d: %s", cudaGetErrorName(e_), __FILE__, __LINE__, cudaGetErrorString(e_)); throw hf::Err(b_); } } while (0)

template <typename Body, int MAXT, int MINB, typename... A>
__global__ void __launch_bounds__(MAXT, MINB) kernel_entry(A... a) {
    extern __shared__ uint4 smem_raw_[];
    KCtx cx{(int)threadIdx.x, (int)blockDim.x, blockIdx.x, blockIdx.y, gridDim.x, gridDim.y};
    Body::run(cx, reinterpret_cast<uint32_t*>(smem_raw_), a...);
}

struct Dev {
    static constexpr int MAX_DEVICES = 64;
    cudaStream_t stream = nullptr;
    uint64_t launches = 0;
    int sm_count = 148;
    int device = 0;  // CUDA device ordinal of the owning context
    // CUDA-graph replay (prover.cuh, transcript mode 2): the host walks the same enqueue code again so that pinned staging areas
    // receive this segment's parameters, but every launch and copy is already a node of the instantiated graph and is skipped here
    bool replay = false;
    template <typename Body, int MAXT = 256, int MINB = 1, typename... A>
    void launch(unsigned gx, unsigned gy, int block, size_t smem, A... a) {
        if (gx == 0 || gy == 0) return;
        if (replay) { launches++; return; }
        auto k = kernel_entry<Body, MAXT, MINB, A...>;
        if (smem > 48 * 1024) {
            // The opt-in limit is a property of (kernel, device), shared by every context and host thread that uses
            // the device: it is only ever raised, under a lock, so a context asking for less cannot lower the limit
            // under another context's launch (seen as cudaErrorInvalidValue with three contexts in flight).
            static std::atomic<size_t> configured_bytes[MAX_DEVICES];
            static std::mutex mu;
            const int d = device >= 0 && device < MAX_DEVICES ? device : 0;
            if (smem > configured_bytes[d].load(std::memory_order_acquire)) {
                std::lock_guard<std::mutex> lock(mu);
                if (smem > configured_bytes[d].load(std::memory_order_relaxed)) {
                    CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    configured_bytes[d].store(smem, std::memory_order_release);
                }
            }
        }
        k<<<dim3(gx, gy), block, smem, stream>>>(a...);
        CUDA_CHECK(cudaGetLastError());
        launches++;
    }
    void* alloc(size_t bytes) { void* p; CUDA_CHECK(cudaMalloc(&p, bytes)); return p; }
    void free(void* p) { if (p) cudaFree(p); }
    void h2d(void* d, const void* h, size_t bytes) { if (!replay) CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, stream)); }
    void d2h(void* h, const void* d, size_t bytes) { if (!replay) CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, stream)); }
    void d2d(void* d, const void* s, size_t bytes) { if (!replay) CUDA_CHECK(cudaMemcpyAsync(d, s, bytes, cudaMemcpyDeviceToDevice, stream)); }
    void zero(void* d, size_t bytes) { if (!replay) CUDA_CHECK(cudaMemsetAsync(d, 0, bytes, stream)); }
    void sync() { CUDA_CHECK(cudaStreamSynchronize(stream)); }
};

struct Timer {
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaStream_t st = nullptr;
    void init(cudaStream_t s) { st = s; CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1)); }
    void destroy() { if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); e0 = e1 = nullptr; }
    void start() { CUDA_CHECK(cudaEventRecord(e0, st)); }
    void stop() { CUDA_CHECK(cudaEventRecord(e1, st)); }
    float ms() { CUDA_CHECK(cudaEventSynchronize(e1)); float m = 0; CUDA_CHECK(cudaEventElapsedTime(&m, e0, e1)); return m; }
};

#else
// ------------------------------------------------------------------ host emulator (tests/emu only)
#define CUDA_CHECK(x) do { } while (0)
struct Dev {
    void* stream = nullptr;
    uint64_t launches = 0;
    int sm_count = 148;
    bool replay = false;  // never set in the emulator (no CUDA graphs)
    template <typename Body, int MAXT = 256, int MINB = 1, typename... A>
    void launch(unsigned gx, unsigned gy, int block, size_t smem, A... a) {
        if (gx == 0 || gy == 0) return;
        launches++;
        #pragma omp parallel
        {
            std::vector<uint32_t> sm(smem / 4 + 8);
            #pragma omp for collapse(2) schedule(dynamic)
            for (long by = 0; by < (long)gy; by++)
                for (long bx = 0; bx < (long)gx; bx++) {
                    if (Body::kBarrier) {
                        Body::run(KCtx{0, 1, (unsigned)bx, (unsigned)by, gx, gy}, sm.data(), a...);
                    } else {
                        for (int t = 0; t < block; t++) Body::run(KCtx{t, block, (unsigned)bx, (unsigned)by, gx, gy}, sm.data(), a...);
                    }
                }
        }
    }
    void* alloc(size_t bytes) { void* p = std::malloc(bytes ? bytes : 1); if (!p) throw Err("emu: out of memory"); return p; }
    void free(void* p) { std::free(p); }
    void h2d(void* d, const void* h, size_t bytes) { std::memcpy(d, h, bytes); }
    void d2h(void* h, const void* d, size_t bytes) { std::memcpy(h, d, bytes); }
    void d2d(void* d, const void* s, size_t bytes) { std::memmove(d, s, bytes); }
    void zero(void* d, size_t bytes) { std::memset(d, 0, bytes); }
    void sync() {}
};
struct Timer {
    void init(void*) {}
    void destroy() {}
    void start() {}
    void stop() {}
    float ms() { return 0.f; }
};
#endif

}  // namespace hf
