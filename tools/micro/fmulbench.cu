// Microbenchmark: Montgomery products per second through the integer pipe (fp.cuh), the FP64 pipe (fp_f64.cuh), and
// both at once (NI integer chains + NF FP64 chains per thread).  Also cross-checks the two multipliers on the device.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fmulbench fmulbench.cu && ./fmulbench
#include <cstdio>
#include <cuda_runtime.h>
#include "fp_f64.cuh"
using namespace pk;

template <class F, int NI, int NF> __global__ void __launch_bounds__(256) chains_kernel(F* out, int iters) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    F x[NI + NF + 1];
#pragma unroll
    for (int k = 0; k <= NI + NF; ++k) x[k] = F::from_u32(t * 7 + 3 + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NI; ++k) x[k] = x[k] * x[k + 1];
#pragma unroll
        for (int k = NI; k < NI + NF; ++k) x[k] = mul_f64(x[k], x[k + 1]);
    }
    F acc = x[0];
#pragma unroll
    for (int k = 1; k <= NI + NF; ++k) acc = acc + x[k];
    st_fp(out + t, acc);
}

template <class F> __global__ void check_kernel(unsigned long long* bad, int iters) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    F a = F::from_u32(t * 2654435761u + 1), b = F::from_u32(t ^ 0x9e3779b9u);
    unsigned long long n = 0;
    for (int i = 0; i < iters; ++i) {
        F r1 = a * b, r2 = mul_f64(a, b);
        if (r1 != r2) ++n;
        a = b; b = r1 + F::from_u32(i);
    }
    // edge values
    F z = F::zero(), m1 = z - F::one();
    if (mul_f64(m1, m1) != m1 * m1) ++n;
    if (mul_f64(z, m1) != z) ++n;
    F pm1; for (int k = 0; k < 8; ++k) pm1.v[k] = F::zero().v[k]; pm1 = z - F::from_u32(0).from_mont();  // 0
    if (n) atomicAdd(bad, n);
}

template <class F, int NI, int NF> static double run(F* out, int blocks, int threads, int iters) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        chains_kernel<F, NI, NF><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return (double)blocks * threads * iters * (NI + NF) / (best * 1e-3) / 1e9;
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    printf("%s, %d SMs\n", prop.name, sms);
    unsigned long long* bad; cudaMallocManaged(&bad, 8); *bad = 0;
    check_kernel<fr_t><<<sms * 4, 128>>>(bad, 2000);
    check_kernel<fq_t><<<sms * 4, 128>>>(bad, 2000);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cross-check f64 vs int multiplier: %llu mismatches (%s)\n", *bad, cudaGetErrorString(e));
    fq_t* out; cudaMalloc(&out, (size_t)sms * 16 * 256 * sizeof(fq_t));
    for (int threads : {128, 256}) {
        for (int bps : {2, 4, 8}) {
            const int blocks = sms * bps, iters = 1024;
            printf("threads %d blocks/SM %d  Gmul/s:", threads, bps);
            printf("  int4 %.1f", run<fq_t, 4, 0>(out, blocks, threads, iters));
            printf("  f64x2 %.1f", run<fq_t, 0, 2>(out, blocks, threads, iters));
            printf("  f64x4 %.1f", run<fq_t, 0, 4>(out, blocks, threads, iters));
            printf("  3i+1f %.1f", run<fq_t, 3, 1>(out, blocks, threads, iters));
            printf("  2i+1f %.1f", run<fq_t, 2, 1>(out, blocks, threads, iters));
            printf("  2i+2f %.1f", run<fq_t, 2, 2>(out, blocks, threads, iters));
            printf("  1i+1f %.1f", run<fq_t, 1, 1>(out, blocks, threads, iters));
            printf("  1i+2f %.1f\n", run<fq_t, 1, 2>(out, blocks, threads, iters));
            fflush(stdout);
        }
    }
    return *bad != 0;
}
