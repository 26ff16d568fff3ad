// Register-operand pressure on sm_100a: do DFMA / IMAD.WIDE keep their rate when all operands are distinct registers?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// MODE 0: dfma d[i] = d[i]*c1 + c2 (shared operands)      MODE 1: dfma d[i] = d[i]*e[i] + f[i] (all distinct, e,f static)
// MODE 2: dfma d[i] = d[j]*d[k] + d[i] (distinct, varying)  MODE 3: mad.wide w[i] = lo(w[j]) * y + w[i]
// MODE 4: carry chain: mad.lo.cc/madc.hi.cc pairs (what mont_mul issues)   MODE 5: MODE 2 + MODE 4 interleaved
// MODE 6: dadd d[i] = d[j] + d[k]   MODE 7: MODE 4 + iadd3 stream  MODE 8: MODE 2 + iadd3 stream (lop3)
template <int MODE> __global__ void __launch_bounds__(256) k(uint64_t* out, int iters, double seed) {
    double d[8], e[8], f[8]; uint64_t w[8]; uint32_t a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { d[i] = seed + i + threadIdx.x; e[i] = 1.0 + 1e-9 * (i + threadIdx.x); f[i] = 1e-7 * i; w[i] = threadIdx.x + 5 * i; a[i] = threadIdx.x * 77 + i; b[i] = threadIdx.x + 13 * i; }
    const double c1 = seed * 0.5, c2 = seed * 0.25;
    uint32_t y = (uint32_t)(seed * 3) | 1;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(c1), "d"(c2));
            if (MODE == 1) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(e[i]), "d"(f[i]));
            if (MODE == 2 || MODE == 5 || MODE == 8) asm volatile("fma.rz.f64 %0, %1, %2, %0;" : "+d"(d[i]) : "d"(d[(i + 3) & 7]), "d"(d[(i + 5) & 7]));
            if (MODE == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((uint32_t)w[(i + 3) & 7]), "r"(y));
            if (MODE == 4 || MODE == 5 || MODE == 7) {
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(a[(i + 3) & 7]), "r"(y));
            }
            if (MODE == 6) asm volatile("add.rz.f64 %0, %1, %2;" : "=d"(d[i]) : "d"(d[(i + 3) & 7]), "d"(d[(i + 5) & 7]));
            if (MODE == 7 || MODE == 8) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(b[(i + 1) & 7]) : "r"(b[(i + 2) & 7]), "r"(y));
        }
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += (uint64_t)d[i] + w[i] + a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> static void run(const char* name, uint64_t* out, int sms, int ops_per_iter) {
    const int blocks = sms * 8, threads = 256, iters = 8192;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, threads>>>(out, iters, 3.0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double clk = 1.965e9;
    const double ops = (double)blocks * threads * iters * ops_per_iter;
    printf("%-44s %8.3f ms   %6.1f listed ops/clk/SM\n", name, best, ops / (best * 1e-3 * clk * sms));
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint64_t* out; cudaMalloc(&out, (size_t)sms * 8 * 256 * 8);
    run<0>("dfma shared operands (8)", out, sms, 8);
    run<1>("dfma distinct static operands (8)", out, sms, 8);
    run<2>("dfma distinct varying operands (8)", out, sms, 8);
    run<6>("dadd distinct (8)", out, sms, 8);
    run<3>("mad.wide varying (8)", out, sms, 8);
    run<4>("mad.lo.cc+madc.hi.cc pairs (8 pairs=16 ops)", out, sms, 16);
    run<5>("dfma varying (8) + carry pairs (16 ops)", out, sms, 24);
    run<7>("carry pairs (16 ops) + lop3 (8)", out, sms, 24);
    run<8>("dfma varying (8) + lop3 (8)", out, sms, 16);
    return 0;
}
