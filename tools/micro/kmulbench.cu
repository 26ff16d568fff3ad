// Integer Montgomery multiplier variants: CIOS (64+64 IMAD.WIDE) vs Karatsuba product + separate reduction (48+64), squaring (30+64).
#include <cstdio>
#include <cuda_runtime.h>
#include "fp_wide.cuh"
using namespace pk;
template <class F, int MODE> __device__ __forceinline__ F op(const F& a, const F& b) {
    F r;
    if (MODE == 0) limbs::mont_mul<FqParams>(r.v, a.v, b.v);
    if (MODE == 1) limbs::mont_mul_k<FqParams>(r.v, a.v, b.v);
    if (MODE == 2) limbs::mont_sqr_k<FqParams>(r.v, a.v);
    if (MODE == 3) limbs::mont_mul_sub_mul<FqParams>(r.v, a.v, b.v, b.v, a.v);  // counts as 2 products
    return r;
}
template <class F, int MODE, int NC> __global__ void __launch_bounds__(256) chains_kernel(F* out, int iters) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    F x[NC + 1];
#pragma unroll
    for (int k = 0; k <= NC; ++k) x[k] = F::from_u32(t * 7 + 3 + k);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < NC; ++k) x[k] = op<F, MODE>(x[k], x[k + 1]);
    }
    F acc = x[0];
#pragma unroll
    for (int k = 1; k <= NC; ++k) acc = acc + x[k];
    st_fp(out + t, acc);
}
template <class F> __global__ void check_kernel(unsigned long long* bad, int iters) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    F a = F::from_u32(t * 2654435761u + 1), b = F::from_u32(t ^ 0x9e3779b9u);
    unsigned long long n = 0;
    for (int i = 0; i < iters; ++i) {
        F r1 = a * b, r2, r3, r4;
        limbs::mont_mul<typename F::params>(r1.v, a.v, b.v);
        limbs::mont_mul_k<typename F::params>(r2.v, a.v, b.v);
        limbs::mont_sqr_k<typename F::params>(r3.v, a.v);
        limbs::mont_mul_sub_mul<typename F::params>(r4.v, a.v, b.v, b.v, r1.v);
        F sq; limbs::mont_mul<typename F::params>(sq.v, a.v, a.v);
        F br; limbs::mont_mul<typename F::params>(br.v, b.v, r1.v);
        if (r1 != r2) ++n;
        if (r3 != sq) ++n;
        if (r4 != r1 - br) ++n;
        a = b; b = r1 + F::from_u32(i);
    }
    if (n) atomicAdd(bad, n);
}
template <class F, int MODE, int NC> static double run(F* out, int blocks, int threads, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        chains_kernel<F, MODE, NC><<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    return (double)blocks * threads * iters * NC * (MODE == 3 ? 2 : 1) / (best * 1e-3) / 1e9;
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    unsigned long long* bad; cudaMallocManaged(&bad, 8); *bad = 0;
    check_kernel<fr_t><<<sms * 4, 128>>>(bad, 1000);
    check_kernel<fq_t><<<sms * 4, 128>>>(bad, 1000);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cross-check Karatsuba / squaring / lazy difference vs CIOS: %llu mismatches (%s)\n", *bad, cudaGetErrorString(e));
    fq_t* out; cudaMalloc(&out, (size_t)sms * 16 * 256 * sizeof(fq_t));
    for (int threads : {128, 256}) for (int bps : {2, 4, 8}) {
        const int blocks = sms * bps, iters = 1024;
        printf("threads %d blocks/SM %d  G products/s:", threads, bps);
        printf("  cios x4 %.1f", run<fq_t, 0, 4>(out, blocks, threads, iters));
        printf("  karatsuba x2 %.1f", run<fq_t, 1, 2>(out, blocks, threads, iters));
        printf("  karatsuba x4 %.1f", run<fq_t, 1, 4>(out, blocks, threads, iters));
        printf("  sqr x4 %.1f", run<fq_t, 2, 4>(out, blocks, threads, iters));
        printf("  a*b-c*d x2 %.1f\n", run<fq_t, 3, 2>(out, blocks, threads, iters));
        fflush(stdout);
    }
    return *bad != 0;
}
