// Is a plain IMAD.WIDE (no carry flag) faster than the carry-chained IMAD.WIDE.X the Montgomery product uses?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// MODE 0: w[i] = x[i] * y + w[i]  (mad.wide, 64-bit accumulate, independent chains, multiplicand fixed per chain)
// MODE 1: t = x[i] * y (mul.wide); w[i] ^= t via two LOP3 (product only, no accumulate in the multiplier)
// MODE 2: mad.lo.cc / madc.hi.cc pairs chained through the carry flag (8 per group, like a mont_mul row)
// MODE 3: w[i] = x[i] * y + w[i] where y changes every iteration (y += 2)
template <int MODE> __global__ void __launch_bounds__(256) k(uint64_t* out, int iters, uint32_t seed) {
    uint64_t w[8]; uint32_t x[8], ca[8], cb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = threadIdx.x * 7 + i; x[i] = (threadIdx.x + 1) * 2654435761u + i * 40503u + seed; ca[i] = i; cb[i] = 3 * i; }
    uint32_t y = seed | 1;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(y));
            if (MODE == 3) y += 2;
        }
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) { uint64_t t; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"(x[i]), "r"(y)); w[i] ^= t; }
            y += 2;
        }
        if (MODE == 2) {
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(ca[0]), "+r"(cb[0]) : "r"(x[0]), "r"(y));
#pragma unroll
            for (int i = 1; i < 8; ++i)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(ca[i]), "+r"(cb[i]) : "r"(x[i]), "r"(y));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += w[i] + ca[i] + cb[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> static void run(const char* name, uint64_t* out, int sms) {
    const int blocks = sms * 8, threads = 256, iters = 8192;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(out, iters, 12345u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    printf("%-64s %8.3f ms   %6.1f products/clk/SM\n", name, best, (double)blocks * threads * iters * 8 / (best * 1e-3 * 1.965e9 * sms));
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint64_t* out; cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    run<0>("mad.wide, 64-bit accumulate, no carry flag, fixed operands", out, sms);
    run<3>("mad.wide, 64-bit accumulate, no carry flag, varying multiplier", out, sms);
    run<1>("mul.wide (product only) + 2 LOP3", out, sms);
    run<2>("mad.lo.cc/madc.hi.cc carry chain (IMAD.WIDE.X)", out, sms);
    return 0;
}
