// Strided NTT stage of the 2^10 x 2^a decomposition with TMA tensor-map I/O and a three-buffer pipeline (sm_100a only).
//
// The strided stage transforms 2^10 rows ("hi") that lie 2^a words apart, for every column offset "lo": a tile is
// 2^10 rows x 16 words (64-byte row segments) = 64 KB.  StridedKernel2 (ntt.cuh) moves such tiles with per-thread
// LDG/STG in two phases (load+round, round+store) inside one CTA per SM, so its memory phase and its arithmetic phase
// alternate instead of overlapping (ncu: DRAM 42 %, heavy-multiply pipe 56 %, issue 53 %).  Here the tile I/O belongs to
// the TMA engine:
//   * a 3-D tensor map {lo, hi, column} per buffer; one elected warp issues `cp.async.bulk.tensor.3d` loads (4 boxes of
//     256 rows x 64 B) that complete on an mbarrier, and `cp.async.bulk.tensor.3d` stores (bulk async-groups);
//   * three 64 KB shared-memory buffers per CTA: while the 512 threads run the two radix-32 register rounds on tile i
//     (in place, shared memory both ways), the loads of tile i+1 are in flight and tile i-1 is draining to HBM;
//   * the kernel never computes a global address: row pitch and column stride live in the tensor map, so one
//     instantiation per direction serves the trace-side (a = 10) and the LDE-side (a = 12) passes.
// Replaces the same risc0-zkp `Hal` ops as ntt.cuh (`batch_interpolate_ntt` / `batch_expand_into_evaluate_ntt` stages,
// SURVEY.md Appendix A.2).  The host emulator (tests/emu) cannot run TMA and keeps StridedKernel2 for these shapes.
#pragma once
#ifndef HFB200_EMU
#include <cuda.h>
#include <algorithm>
#include "ntt.cuh"

namespace hf {

static constexpr int TMA_ROWS_LOG = 10, TMA_SEG_LOG = 4;              // 2^10 rows x 2^4 words per tile
static constexpr uint32_t TMA_TILE_BYTES = (1u << (TMA_ROWS_LOG + TMA_SEG_LOG)) * 4u;  // 64 KB
static constexpr int TMA_NBUF = 3, TMA_THREADS = 512;
#ifndef TMA_DONE_MBAR
#define TMA_DONE_MBAR 1
#endif
static constexpr size_t TMA_SMEM = (size_t)TMA_NBUF * TMA_TILE_BYTES + (1u << TMA_ROWS_LOG) * 4 + 64;  // tiles + 512 (w, w') pairs + mbarriers

struct StrTmaArgs {
    uint32_t n_tiles;
    int a;  // log2 of the row pitch in words: tiles per column = 2^(a - 4)
    RootTables rt;
};

__device__ __forceinline__ void tma_load_3d(void* dst_smem, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst_smem)), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t c0, uint32_t c1, uint32_t c2, const void* src_smem) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];"
                 ::"l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(src_smem)) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// grid = one CTA per SM (persistent over tiles); block = 512 threads = one radix-32 item each per round
template <bool INV>
__global__ void __launch_bounds__(TMA_THREADS, 1) strided_tma_kernel(const __grid_constant__ CUtensorMap map_in, const __grid_constant__ CUtensorMap map_out, StrTmaArgs p) {
    extern __shared__ __align__(1024) uint8_t smem_tma_[];
    uint32_t* tw = reinterpret_cast<uint32_t*>(smem_tma_ + (size_t)TMA_NBUF * TMA_TILE_BYTES);
    uint64_t* full = reinterpret_cast<uint64_t*>(tw + (1u << TMA_ROWS_LOG));
    uint64_t* done = full + TMA_NBUF;
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const KCtx cx{tid, TMA_THREADS, blockIdx.x, 0u, gridDim.x, 1u};
    auto tile_buf = [&](uint32_t b) { return reinterpret_cast<uint32_t*>(smem_tma_ + (size_t)b * TMA_TILE_BYTES); };

    // w_{1024}^(+-i), i < 512, as Shoup (w, w') pairs in natural order (the rounds below use the strided-index layout)
    for (uint32_t i = (uint32_t)tid; i < (1u << (TMA_ROWS_LOG - 1)); i += TMA_THREADS) {
        const uint32_t w = INV ? tab_pow(p.rt.i_lo, p.rt.i_hi, i << (24 - TMA_ROWS_LOG)) : tab_pow(p.rt.f_lo, p.rt.f_hi, i << (24 - TMA_ROWS_LOG));
        tw[2 * i] = from_mont(w); tw[2 * i + 1] = shoup_quot_mont(w);
    }
    if (tid == 0) {
        for (int b = 0; b < TMA_NBUF; b++) { mbar_init(&full[b], 1); mbar_init(&done[b], TMA_THREADS / 32); }
        mbar_init_fence();
    }
    __syncthreads();

    const uint32_t tpc_log = (uint32_t)(p.a - TMA_SEG_LOG);
    // four boxes of 256 rows x 16 words, issued by lanes 0..3 of warp 0; lane 0 arms the barrier first
    auto issue_load = [&](uint32_t t, uint32_t b) {
        if (lane == 0) mbar_expect_tx(&full[b], TMA_TILE_BYTES);
        __syncwarp();
        if (lane < 4) tma_load_3d(tile_buf(b) + lane * (256u << TMA_SEG_LOG), &map_in, (t & ((1u << tpc_log) - 1u)) << TMA_SEG_LOG, (uint32_t)lane * 256u, t >> tpc_log, &full[b]);
    };
    if (warp == 0) {
        const uint32_t t0 = blockIdx.x, t1 = blockIdx.x + gridDim.x;
        if (t0 < p.n_tiles) issue_load(t0, 0);
        if (t1 < p.n_tiles) issue_load(t1, 1);
    }
    uint32_t it = 0;
    for (uint32_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x, it++) {
        const uint32_t b = it % TMA_NBUF;
        mbar_wait(&full[b], (it / TMA_NBUF) & 1u);
        const SmemIOT<31> S{tile_buf(b)};  // dense [row][16]: element (row, word) at row * 16 + word (pad shift 31 = no padding)
        if (INV) {
            round_t<5, true, TMA_ROWS_LOG, TMA_SEG_LOG, 5, false, false, true>(cx, tw, TMA_ROWS_LOG, TMA_SEG_LOG, 5, S, S);
            __syncthreads();
            round_t<5, true, TMA_ROWS_LOG, TMA_SEG_LOG, 0, false, false, true>(cx, tw, TMA_ROWS_LOG, TMA_SEG_LOG, 0, S, S);
        } else {
            round_t<5, false, TMA_ROWS_LOG, TMA_SEG_LOG, 0, false, false, true>(cx, tw, TMA_ROWS_LOG, TMA_SEG_LOG, 0, S, S);
            __syncthreads();
            round_t<5, false, TMA_ROWS_LOG, TMA_SEG_LOG, 5, false, false, true>(cx, tw, TMA_ROWS_LOG, TMA_SEG_LOG, 5, S, S);
        }
        fence_proxy_async_smem();  // this thread's shared-memory writes become visible to the TMA engine
#if TMA_DONE_MBAR
        // no second CTA-wide barrier: every warp announces "my part of tile it is in shared memory" on done[b] and moves on to
        // the next tile (already loaded); only warp 0 waits for all 16 warps before it hands the buffer to the TMA store
        __syncwarp();
        if (lane == 0) mbar_arrive(&done[b]);
        if (warp == 0) mbar_wait(&done[b], (it / TMA_NBUF) & 1u);
#else
        __syncthreads();
#endif
        if (warp == 0) {
            if (lane < 4) {
                tma_store_3d(&map_out, (t & ((1u << tpc_log) - 1u)) << TMA_SEG_LOG, (uint32_t)lane * 256u, t >> tpc_log, tile_buf(b) + lane * (256u << TMA_SEG_LOG));
                tma_commit();
            }
            const uint32_t t2 = t + 2u * gridDim.x;
            if (t2 < p.n_tiles) {
                // buffer (it + 2) % 3 held tile it - 1: its store (this lane's previous bulk group) must have read it out
                if (lane < 4) tma_wait_read<1>();
                __syncwarp();
                issue_load(t2, (it + 2u) % TMA_NBUF);
            }
        }
    }
    if (warp == 0 && lane < 4) tma_wait_all<0>();
}

// ---- host side: tensor maps + launch ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static inline EncodeTiledFn tma_encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// {lo (2^a words, contiguous), hi (2^10 rows, pitch 2^a words), column (pitch col_stride words)}; box = 16 x 256 x 1
static inline bool tma_make_map(CUtensorMap* m, const uint32_t* base, uint64_t col_stride, uint32_t ncols, int a) {
    EncodeTiledFn enc = tma_encode_fn();
    if (!enc) return false;
    if ((reinterpret_cast<uintptr_t>(base) & 15u) || ((col_stride * 4) & 15u) || a < TMA_SEG_LOG) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)1 << a, (cuuint64_t)1 << TMA_ROWS_LOG, ncols};
    const cuuint64_t strides[2] = {((cuuint64_t)4) << a, (cuuint64_t)col_stride * 4};
    const cuuint32_t box[3] = {1u << TMA_SEG_LOG, 256u, 1u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint32_t*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Returns false when the shape is not this kernel's (caller falls back to StridedKernel2).
static inline bool strided_tma(Dev* dev, const RootTables& rt, const uint32_t* in, uint64_t in_stride, uint32_t* out, uint64_t out_stride, uint32_t ncols, int a, int b, bool inv) {
    static const bool enabled = [] { const char* e = std::getenv("HFB200_NTT_TMA"); return !e || std::atoi(e) != 0; }();
    if (!enabled || b != TMA_ROWS_LOG || a < TMA_SEG_LOG || ncols == 0) return false;
    if (dev->replay) { dev->launches++; return true; }  // CUDA-graph replay: this launch is a node of the instantiated graph
    alignas(64) CUtensorMap mi, mo;
    if (!tma_make_map(&mi, in, in_stride, ncols, a) || !tma_make_map(&mo, out, out_stride, ncols, a)) return false;
    StrTmaArgs p{};
    const uint64_t tiles = (uint64_t)ncols << (a - TMA_SEG_LOG);
    if (tiles > 0xFFFFFFFFull) return false;
    p.n_tiles = (uint32_t)tiles; p.a = a; p.rt = rt;
    static std::atomic<bool> configured[Dev::MAX_DEVICES][2];
    const int d = dev->device >= 0 && dev->device < Dev::MAX_DEVICES ? dev->device : 0;
    if (!configured[d][inv ? 1 : 0].load(std::memory_order_acquire)) {
        if (inv) CUDA_CHECK(cudaFuncSetAttribute(strided_tma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
        else CUDA_CHECK(cudaFuncSetAttribute(strided_tma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TMA_SMEM));
        configured[d][inv ? 1 : 0].store(true, std::memory_order_release);
    }
    const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)dev->sm_count, tiles);
    if (inv) strided_tma_kernel<true><<<grid, TMA_THREADS, TMA_SMEM, dev->stream>>>(mi, mo, p);
    else strided_tma_kernel<false><<<grid, TMA_THREADS, TMA_SMEM, dev->stream>>>(mi, mo, p);
    CUDA_CHECK(cudaGetLastError());
    dev->launches++;
    return true;
}

}  // namespace hf
#endif  // !HFB200_EMU
