// Modular-multiply formulations for BabyBear on B200 (sm_100a): sustained throughput of each instruction sequence,
// timed over >= 10 ms per probe with CUDA events (clocks warm), 8 independent chains per thread, 8 warps per SMSP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mulprobe mulprobe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ILP 8
static constexpr uint32_t P = 2013265921u, PINV = 0x88000001u;

__device__ __forceinline__ uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }
// 1: WIDE, LO, HI, SUB, MIN (canonical)
__device__ __forceinline__ uint32_t mul_canon(uint32_t a, uint32_t b) {
    uint64_t o = (uint64_t)a * b; uint32_t m = (uint32_t)o * PINV; uint32_t r = (uint32_t)(o >> 32) - __umulhi(m, P); return umin32(r, r + P);
}
// 2: signed WIDE, LO, HI, SUB  -> (-p, p)
__device__ __forceinline__ int32_t mul_signed(int32_t a, int32_t b) {
    int64_t o = (int64_t)a * b; int32_t m = (int32_t)((uint32_t)o * PINV); return (int32_t)(o >> 32) - __mulhi(m, (int32_t)P);
}
// 3: WIDE, LO(negated inverse), WIDE with 64-bit accumulate -> hi word in (0, 2p)
__device__ __forceinline__ uint32_t mul_acc_lazy(uint32_t a, uint32_t b) {
    uint64_t o = (uint64_t)a * b; uint32_t m = (uint32_t)o * (0u - PINV);
    uint64_t r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(m), "r"(P), "l"(o));
    return (uint32_t)(r >> 32);
}
__device__ __forceinline__ uint32_t mul_acc_canon(uint32_t a, uint32_t b) { uint32_t r = mul_acc_lazy(a, b); return umin32(r, r - P); }
// 5: Shoup with canonical output
__device__ __forceinline__ uint32_t mul_shoup(uint32_t x, uint32_t d, uint32_t dp) { uint32_t q = __umulhi(x, dp); uint32_t r = x * d - q * P; return umin32(r, r - P); }

__device__ __forceinline__ uint32_t sbox_cur(uint32_t s, uint32_t rc) {  // library's signed chain
    const int32_t x = (int32_t)(s + rc - P);
    const int32_t x2 = mul_signed(x, x), x3 = mul_signed(x2, x), x4 = mul_signed(x2, x2);
    const uint32_t r = (uint32_t)mul_signed(x3, x4);
    return umin32(r, r + P);
}
__device__ __forceinline__ uint32_t sbox_old(uint32_t s, uint32_t rc) {  // unsigned chain with corrections
    uint32_t x = s + rc; x = umin32(x, x - P);
    uint32_t x2 = mul_canon(x, x), x3 = mul_canon(x2, x);
    uint64_t o = (uint64_t)x2 * x2; uint32_t m = (uint32_t)o * PINV; uint32_t x4 = (uint32_t)(o >> 32) - __umulhi(m, P) + P;
    return mul_canon(x3, x4);
}
__device__ __forceinline__ uint32_t sbox_acc(uint32_t s, uint32_t rc) {  // WIDE-accumulate chain
    uint32_t x = s + rc; x = umin32(x, x - P);
    const uint32_t x2 = mul_acc_canon(x, x);
    const uint32_t x3 = mul_acc_lazy(x2, x);
    const uint32_t x4 = mul_acc_canon(x2, x2);
    return mul_acc_canon(x3, x4);
}

template <int OP>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed, int iters) {
    uint32_t x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = (seed + threadIdx.x * 7 + i * 13) % P; y[i] = (seed * 3 + i + threadIdx.x * 11 + blockIdx.x) % P; }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 1) x[i] = mul_canon(x[i], y[i]);
            else if (OP == 2) x[i] = (uint32_t)mul_signed((int32_t)x[i], (int32_t)y[i]);
            else if (OP == 3) x[i] = mul_acc_lazy(x[i], y[i]);
            else if (OP == 4) x[i] = mul_acc_canon(x[i], y[i]);
            else if (OP == 5) x[i] = mul_shoup(x[i], y[i], y[(i + 1) % ILP]);
            else if (OP == 6) x[i] = sbox_cur(x[i], y[i]);
            else if (OP == 7) x[i] = sbox_old(x[i], y[i]);
            else if (OP == 8) x[i] = sbox_acc(x[i], y[i]);
            else if (OP == 9) { uint32_t s = x[i] + y[i]; x[i] = umin32(s, s - P); }
            else if (OP == 10) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(it));
            else if (OP == 11) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(it));
            else if (OP == 12) { uint64_t a = ((uint64_t)y[i] << 32) | x[i]; asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a) : "r"(x[i]), "r"(y[i])); x[i] = (uint32_t)a; y[i] = (uint32_t)(a >> 32); }
            else if (OP == 13) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i]));
            else if (OP == 14) asm volatile("min.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(y[i] ));
            else if (OP == 15) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(it)); asm volatile("min.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(it)); }
            else if (OP == 16) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(it)); asm volatile("min.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(it)); asm volatile("add.u32 %0, %0, %1;" : "+r"(y[i]) : "r"(it)); }
            else if (OP == 17) { asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(it)); asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(y[i]) : "r"(y[i]), "r"(it)); }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int OP>
void run(const char* name, int iters, int mults_per_op) {
    uint32_t* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    probe<OP><<<148 * 4, 256>>>(out, 12345, iters);  // warm-up (clocks)
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<148 * 4, 256>>>(out, 12345, iters);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = 148.0 * 4 * 256 * ILP * (double)iters;
    const double per_s = ops / (ms * 1e-3);
    // cycles per warp-op per SMSP at 1.965 GHz nominal
    const double cyc = 1.965e9 * 592.0 * 32.0 / per_s;
    printf("%-44s %8.3f ms  %8.1f Gop/s  %6.2f cycles/warp-op/SMSP @1.965GHz  (%d modmul per op -> %.1f Gmul/s)\n", name, ms, per_s * 1e-9, cyc, mults_per_op, per_s * mults_per_op * 1e-9);
    cudaFree(out);
}

int main() {
    const int N = 1 << 16;
    run<1>("1 canonical: WIDE LO HI SUB MIN", N, 1);
    run<2>("2 signed:    WIDE LO HI SUB", N, 1);
    run<3>("3 acc lazy:  WIDE LO WIDEacc", N, 1);
    run<4>("4 acc canon: WIDE LO WIDEacc MIN", N, 1);
    run<5>("5 shoup canon: HI LO LO MIN", N, 1);
    run<9>("9 fadd: IADD MIN", N, 0);
    run<6>("6 sbox signed chain (library)", N / 4, 4);
    run<7>("7 sbox unsigned chain (previous)", N / 4, 4);
    run<8>("8 sbox WIDE-accumulate chain", N / 4, 4);
    run<1>("1 again (clock drift check)", N, 1);
    run<10>("10 IMAD.LO", N, 0); run<11>("11 IMAD.HI", N, 0); run<12>("12 IMAD.WIDE acc64", N, 0); run<13>("13 IADD", N, 0); run<14>("14 MIN", N, 0);
    run<15>("15 IMAD.LO + MIN (independent)", N, 0); run<16>("16 IMAD.HI + MIN + IADD (independent)", N, 0); run<17>("17 IMAD.HI + IMAD.LO (independent)", N, 0);
    return 0;
}
