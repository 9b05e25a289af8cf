// Integer-pipe throughput probe for B200 (sm_100a): cycles per warp-instruction per SMSP for the ops that
// matter for BabyBear arithmetic.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define ILP 8

template <int OP>
__global__ void __launch_bounds__(256) probe(uint32_t* out, uint32_t seed, long long* cycles) {
    uint32_t x[ILP], y[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) { x[i] = seed + threadIdx.x * 7 + i * 13; y[i] = seed * 3 + i + threadIdx.x; }
    const uint32_t P = 2013265921u, PINV = 0x88000001u;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (OP == 0) x[i] = x[i] * y[i] + it;                                    // IMAD.LO
            else if (OP == 1) x[i] = __umulhi(x[i], y[i]) + it;                      // IMAD.HI
            else if (OP == 2) { uint64_t t = (uint64_t)x[i] * y[i] + it; x[i] = (uint32_t)t; y[i] ^= (uint32_t)(t >> 32); }  // IMAD.WIDE
            else if (OP == 3) x[i] = x[i] + y[i] + it;                               // IADD3
            else if (OP == 4) x[i] = min(x[i], y[i] - (uint32_t)it);                 // IMNMX + IADD
            else if (OP == 5) x[i] = x[i] + (y[i] << 4);                             // LEA
            else if (OP == 6) x[i] = (x[i] >> 5) ^ y[i];                             // SHF + LOP3
            else if (OP == 7) {                                                       // full Montgomery fmul
                uint64_t o = (uint64_t)x[i] * y[i]; uint32_t m = (uint32_t)o * PINV; uint32_t mp = __umulhi(m, P);
                uint32_t r = (uint32_t)(o >> 32) - mp; x[i] = min(r, r + P);
            } else if (OP == 8) {                                                     // fmul with m via 2 LEA
                uint64_t o = (uint64_t)x[i] * y[i]; uint32_t lo = (uint32_t)o; uint32_t t = lo + (lo << 4); uint32_t m = lo + (t << 27);
                uint32_t mp = __umulhi(m, P); uint32_t r = (uint32_t)(o >> 32) - mp; x[i] = min(r, r + P);
            } else if (OP == 9) {                                                     // Shoup constant multiply, lazy [0,2p)
                uint32_t q = __umulhi(x[i], y[i]); x[i] = x[i] * 123456789u - q * P;
            } else if (OP == 10) {                                                    // fadd
                uint32_t s = x[i] + y[i]; x[i] = min(s, s - P);
            } else if (OP == 11) {                                                    // mul.lo + mul.hi separately (no WIDE)
                uint32_t lo = x[i] * y[i], hi = __umulhi(x[i], y[i]); uint32_t m = lo * PINV; uint32_t mp = __umulhi(m, P);
                uint32_t r = hi - mp; x[i] = min(r, r + P);
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) acc ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_iter) {
    uint32_t* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 4 * 256 * 4 * 4); cudaMalloc(&cyc, 8);
    // 1 block of 256 threads per SM x 4 blocks -> 32 warps/SM = 8 warps per SMSP
    probe<OP><<<148 * 4, 256>>>(out, 12345, cyc);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<148 * 4, 256>>>(out, 12345, cyc);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // per SMSP: 8 warps, each ITERS*ILP groups
    double warp_groups_per_smsp = 8.0 * ITERS * ILP;
    printf("%-28s cycles(block0)=%lld  -> %.2f cycles per warp-op-group per SMSP (%d instr-ish each), %.3f ms\n", name, h, (double)h / warp_groups_per_smsp, ops_per_iter, ms);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("IMAD.LO", 1); run<1>("IMAD.HI", 1); run<2>("IMAD.WIDE", 1); run<3>("IADD3", 1); run<4>("IMNMX+IADD", 2);
    run<5>("LEA", 1); run<6>("SHF+LOP3", 2); run<7>("fmul montgomery", 6); run<8>("fmul mont, m by 2 LEA", 7);
    run<9>("shoup lazy", 3); run<10>("fadd", 3); run<11>("fmul lo+hi split", 7);
    return 0;
}
