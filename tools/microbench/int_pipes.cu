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
            } else if (OP == 12) {                                                    // signed Montgomery, no correction
                int64_t o = (int64_t)(int32_t)x[i] * (int32_t)y[i]; int32_t m = (int32_t)((uint32_t)o * PINV); int32_t mp = __mulhi(m, (int32_t)P);
                x[i] = (uint32_t)((int32_t)(o >> 32) - mp);
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

// FP64 modular multiply on exact-integer doubles: h = a*b (rounded), l = fma(a,b,-h) (exact low part),
// q = rint(h/p) by the magic-constant trick, r = fma(-q,p,h) + l in (-p, p)-ish.  6 FP64-pipe ops.
__device__ __forceinline__ double dmodmul(double a, double b) {
    const double Pd = 2013265921.0, PINVd = 1.0 / 2013265921.0, MAGIC = 6755399441055744.0;
    double h = a * b;
    double l = fma(a, b, -h);
    double q = fma(h, PINVd, MAGIC) - MAGIC;
    double r = fma(-q, Pd, h);
    return r + l;
}
// MODE 0: all warps integer fmul; 1: all warps fp64 modmul; 2: DFMA only; k >= 3: warps with (warp % (k-1)) == 0 run fp64, others integer
template <int MODE>
__global__ void __launch_bounds__(256) hybrid(uint32_t* out, uint32_t seed, int iters_int, int iters_f64) {
    const uint32_t P = 2013265921u, PINV = 0x88000001u;
    const int warp = threadIdx.x >> 5;
    bool f64 = MODE == 1 || MODE == 2 || (MODE >= 3 && (warp % (MODE - 1)) == 0);
    uint32_t acc = 0;
    if (!f64) {
        uint32_t x[ILP], y[ILP];
#pragma unroll
        for (int i = 0; i < ILP; i++) { x[i] = (seed + threadIdx.x * 7 + i * 13) % P; y[i] = (seed * 3 + i + threadIdx.x) % P; }
#pragma unroll 1
        for (int it = 0; it < iters_int; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                uint64_t o = (uint64_t)x[i] * y[i]; uint32_t m = (uint32_t)o * PINV; uint32_t mp = __umulhi(m, P);
                uint32_t r = (uint32_t)(o >> 32) - mp; x[i] = min(r, r + P);
            }
        }
#pragma unroll
        for (int i = 0; i < ILP; i++) acc ^= x[i];
    } else {
        double x[ILP], y[ILP];
#pragma unroll
        for (int i = 0; i < ILP; i++) { x[i] = (double)((seed + threadIdx.x * 7 + i * 13) % P); y[i] = (double)((seed * 3 + i + threadIdx.x) % P); }
#pragma unroll 1
        for (int it = 0; it < iters_f64; it++) {
#pragma unroll
            for (int i = 0; i < ILP; i++) {
                if (MODE == 2) x[i] = fma(x[i], y[i], 1.0);
                else x[i] = dmodmul(x[i], y[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < ILP; i++) acc ^= (uint32_t)(long long)x[i];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run_hybrid(const char* name, int iters_int, int iters_f64, double frac_f64) {
    uint32_t* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    hybrid<MODE><<<148 * 4, 256>>>(out, 12345, iters_int, iters_f64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    hybrid<MODE><<<148 * 4, 256>>>(out, 12345, iters_int, iters_f64);
    cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double threads = 148.0 * 4 * 256;
    double mults = threads * ILP * ((1.0 - frac_f64) * iters_int + frac_f64 * iters_f64);
    printf("%-44s %.3f ms  %.1f Gop/s (int iters %d, f64 iters %d, f64 warp share %.2f)\n", name, ms, mults / ms * 1e-6, iters_int, iters_f64, frac_f64);
    cudaFree(out);
}

// exactness check of dmodmul against the integer product
__global__ void dcheck(unsigned long long* bad, uint32_t seed) {
    const uint32_t P = 2013265921u;
    uint32_t a = (seed + 2654435761u * (blockIdx.x * blockDim.x + threadIdx.x)) % P, b = (seed * 7 + 40503u * threadIdx.x + 977u * blockIdx.x) % P;
    double x = (double)a;
    uint32_t xi = a;
    for (int it = 0; it < 64; it++) {
        x = dmodmul(x, (double)b);
        xi = (uint32_t)(((uint64_t)xi * b) % P);
        long long v = (long long)x; long long w = ((v % (long long)P) + P) % P;
        if ((uint32_t)w != xi || x != (double)v || x >= 2.2e9 || x <= -2.2e9) atomicAdd(bad, 1ull);
    }
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
    run<9>("shoup lazy", 3); run<10>("fadd", 3); run<11>("fmul lo+hi split", 7); run<12>("fmul signed, no correction", 4);
    unsigned long long* bad; cudaMalloc(&bad, 8); cudaMemset(bad, 0, 8);
    dcheck<<<1024, 256>>>(bad, 99991); unsigned long long hb; cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost);
    printf("dmodmul exactness: %llu mismatches of %d\n", hb, 1024 * 256 * 64);
    run_hybrid<0>("all warps integer fmul", 4096, 0, 0.0);
    run_hybrid<1>("all warps fp64 modmul", 0, 4096, 1.0);
    run_hybrid<2>("all warps DFMA", 0, 4096, 1.0);
    // hybrids: choose iteration counts so both kinds of warp finish at about the same time (tune from the two lines above)
    run_hybrid<3>("1 of 2 warps fp64, equal iters", 4096, 4096, 0.5);
    run_hybrid<3>("1 of 2 warps fp64, f64 iters/2", 4096, 2048, 0.5);
    run_hybrid<3>("1 of 2 warps fp64, f64 iters/4", 4096, 1024, 0.5);
    run_hybrid<5>("1 of 4 warps fp64, equal iters", 4096, 4096, 0.25);
    run_hybrid<5>("1 of 4 warps fp64, f64 iters/2", 4096, 2048, 0.25);
    run_hybrid<5>("1 of 4 warps fp64, f64 iters x2", 2048, 4096, 0.25);
    return 0;
}
