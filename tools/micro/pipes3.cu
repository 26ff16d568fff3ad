// Which integer multiply forms share a pipe on sm_100a?  Independent per-register chains, 8 of each listed form.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// bit 0: mad.wide (64-bit accumulate)  bit 1: mad.lo  bit 2: mad.hi  bit 3: dfma  bit 4: lop3  bit 5: mad.lo.cc/madc.hi chain (WIDE.X)
template <int MASK> __global__ void __launch_bounds__(256) k(uint64_t* out, int iters, uint32_t seed) {
    uint64_t w[8]; uint32_t lo[8], hi[8], lg[8], ca[8], cb[8]; double d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { w[i] = threadIdx.x * 7 + i; lo[i] = threadIdx.x + i * 3; hi[i] = threadIdx.x * 5 + i; lg[i] = threadIdx.x ^ (i * 77); d[i] = 1.0 + threadIdx.x + i; ca[i] = i + threadIdx.x; cb[i] = 3 * i + threadIdx.x; }
    const uint32_t y = seed | 1;
    const double c1 = 1.0 + 1e-9 * seed, c2 = 1e-7 * seed;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MASK & 1) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[i]), "r"(y));
            if (MASK & 2) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(lo[i]) : "r"(y));
            if (MASK & 4) asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(hi[i]) : "r"(y));
            if (MASK & 8) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c1), "d"(c2));
            if (MASK & 16) asm volatile("lop3.b32 %0, %0, %1, %0, 0x96;" : "+r"(lg[i]) : "r"(y));
        }
        if (MASK & 32) {
            // one 8-long carry chain like a mont_mul row: (ca[i], cb[i]) += ca[i] * y with carries rippling through
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(ca[0]), "+r"(cb[0]) : "r"(lo[0]), "r"(y));
#pragma unroll
            for (int i = 1; i < 8; ++i)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(ca[i]), "+r"(cb[i]) : "r"(lo[i]), "r"(y));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += w[i] + lo[i] + hi[i] + lg[i] + (uint64_t)d[i] + ca[i] + cb[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MASK> static void run(const char* name, uint64_t* out, int sms) {
    const int blocks = sms * 8, threads = 256, iters = 8192;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MASK><<<blocks, threads>>>(out, iters, 12345u);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // each selected form contributes 8 ops per iteration per lane
    const double per_form = (double)blocks * threads * iters * 8 / (best * 1e-3 * 1.965e9 * sms);
    printf("%-40s %8.3f ms   %6.1f of each form /clk/SM\n", name, best, per_form);
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint64_t* out; cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    run<1>("mad.wide", out, sms);
    run<2>("mad.lo", out, sms);
    run<4>("mad.hi", out, sms);
    run<8>("dfma", out, sms);
    run<16>("lop3", out, sms);
    run<32>("carry chain (WIDE.X) 8 per iter", out, sms);
    run<1 | 2>("mad.wide + mad.lo", out, sms);
    run<1 | 4>("mad.wide + mad.hi", out, sms);
    run<2 | 4>("mad.lo + mad.hi", out, sms);
    run<1 | 2 | 4>("mad.wide + mad.lo + mad.hi", out, sms);
    run<1 | 8>("mad.wide + dfma", out, sms);
    run<2 | 8>("mad.lo + dfma", out, sms);
    run<4 | 8>("mad.hi + dfma", out, sms);
    run<1 | 16>("mad.wide + lop3", out, sms);
    run<32 | 2>("carry chain + mad.lo", out, sms);
    run<32 | 4>("carry chain + mad.hi", out, sms);
    run<32 | 2 | 4>("carry chain + mad.lo + mad.hi", out, sms);
    run<32 | 16>("carry chain + lop3", out, sms);
    return 0;
}
