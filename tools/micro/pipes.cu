// Which sm_100a pipes run side by side?  Streams of independent DFMA / IADD3 / IMAD.WIDE / LOP3 chains, alone and mixed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int ND, int NA, int NW, int NM> __global__ void __launch_bounds__(256) k(uint64_t* out, int iters, double seed) {
    double d[8]; uint32_t a[8]; uint64_t w[8]; uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = seed + i + threadIdx.x; a[i] = threadIdx.x * 3 + i; w[i] = threadIdx.x + 5 * i; m[i] = threadIdx.x ^ i; }
    const double c1 = seed * 0.5, c2 = seed * 0.25;
    uint32_t x = (uint32_t)seed | 1, y = (uint32_t)(seed * 3) | 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < ND; ++i) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c1), "d"(c2));
#pragma unroll
            for (int i = 0; i < NA; ++i) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(x));   // IADD3 (alu pipe)
#pragma unroll
            for (int i = 0; i < NW; ++i) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x), "r"(y));
#pragma unroll
            for (int i = 0; i < NM; ++i) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(m[i]) : "r"(x), "r"(y));  // IMAD 32-bit
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (uint64_t)d[i] + a[i] + w[i] + m[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ND, int NA, int NW, int NM> static void run(const char* name, uint64_t* out, int sms) {
    const int blocks = sms * 8, threads = 256, iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<ND, NA, NW, NM><<<blocks, threads>>>(out, iters, 3.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    const double per_thread_iter = 4.0;  // the r loop
    const double total_groups = (double)blocks * threads * iters * per_thread_iter;
    // clocks per SM per "group" (one of each listed op per lane): time * clk * sms / groups
    const double clk_per_group = best * 1e-3 * clk_khz * 1e3 * sms / total_groups;
    printf("%-28s %8.3f ms   lanes/clk/SM:", name, best);
    if (ND) printf("  dfma %.1f", ND / clk_per_group);
    if (NA) printf("  iadd %.1f", NA / clk_per_group);
    if (NW) printf("  imad.wide %.1f", NW / clk_per_group);
    if (NM) printf("  imad %.1f", NM / clk_per_group);
    printf("\n");
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint64_t* out; cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    printf("%s, %d SMs, %d kHz\n", prop.name, sms, prop.clockRate);
    run<8, 0, 0, 0>("dfma x8", out, sms);
    run<0, 8, 0, 0>("iadd x8", out, sms);
    run<0, 0, 8, 0>("imad.wide x8", out, sms);
    run<0, 0, 0, 8>("imad x8", out, sms);
    run<4, 4, 0, 0>("dfma x4 + iadd x4", out, sms);
    run<4, 0, 4, 0>("dfma x4 + imad.wide x4", out, sms);
    run<0, 4, 4, 0>("iadd x4 + imad.wide x4", out, sms);
    run<0, 8, 4, 0>("iadd x8 + imad.wide x4", out, sms);
    run<4, 4, 4, 0>("dfma x4 + iadd x4 + wide x4", out, sms);
    run<2, 0, 8, 0>("dfma x2 + imad.wide x8", out, sms);
    run<8, 0, 2, 0>("dfma x8 + imad.wide x2", out, sms);
    run<0, 4, 0, 4>("iadd x4 + imad x4", out, sms);
    run<4, 0, 0, 4>("dfma x4 + imad x4", out, sms);
    run<0, 0, 4, 4>("imad.wide x4 + imad x4", out, sms);
    return 0;
}
