// Micro-benchmark: where does the per-bin counting sort spend its time?  (tools only, not part of the library)
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <vector>
#define FS_THREADS 256
#define FS_CAP 7168
template <int MODE>
__global__ void __launch_bounds__(FS_THREADS) fine(const uint2* tmp, uint2* entries, const uint32_t* off, int fine_bits, uint32_t* sink) {
    extern __shared__ uint2 stage[];
    const uint32_t F = 1u << fine_bits, fmask = F - 1;
    uint32_t* hist = reinterpret_cast<uint32_t*>(stage + FS_CAP);
    uint32_t* part = hist + F;
    const uint32_t lo = off[blockIdx.x], hi = off[blockIdx.x + 1];
    if (lo == hi) return;
    const uint32_t size = hi - lo, staged = size < FS_CAP ? size : FS_CAP;
    for (uint32_t b = threadIdx.x; b < F; b += FS_THREADS) hist[b] = 0;
    {
        uint2 r[FS_CAP / FS_THREADS];
#pragma unroll
        for (int k = 0; k < FS_CAP / FS_THREADS; ++k) { const uint32_t i = threadIdx.x + k * FS_THREADS; if (i < staged) r[k] = tmp[lo + i]; }
#pragma unroll
        for (int k = 0; k < FS_CAP / FS_THREADS; ++k) { const uint32_t i = threadIdx.x + k * FS_THREADS; if (i < staged) stage[i] = r[k]; }
    }
    __syncthreads();
    if (MODE == 4) { if (stage[threadIdx.x].x == 0xdeadbeef) sink[0] = 1; return; }
    if (MODE != 3)
        for (uint32_t i = threadIdx.x; i < size; i += FS_THREADS) atomicAdd(&hist[stage[i].x & fmask], 1u);
    __syncthreads();
    const uint32_t per = (F + FS_THREADS - 1) / FS_THREADS;
    const uint32_t b0 = threadIdx.x * per < F ? threadIdx.x * per : F, b1 = (b0 + per < F) ? b0 + per : F;
    uint32_t sum = 0;
    for (uint32_t b = b0; b < b1; ++b) sum += hist[b];
    part[threadIdx.x] = sum;
    __syncthreads();
    for (unsigned d = 1; d < FS_THREADS; d <<= 1) {
        uint32_t t = threadIdx.x >= d ? part[threadIdx.x - d] : 0;
        __syncthreads();
        part[threadIdx.x] += t;
        __syncthreads();
    }
    uint32_t run = part[threadIdx.x] - sum;
    for (uint32_t b = b0; b < b1; ++b) { uint32_t x = hist[b]; hist[b] = run; run += x; }
    __syncthreads();
    if (MODE == 2) { if (hist[threadIdx.x & fmask] == 0xdeadbeef) sink[0] = 1; return; }
    uint32_t acc = 0;
    for (uint32_t i = threadIdx.x; i < size; i += FS_THREADS) {
        const uint2 ent = stage[i];
        const uint32_t pos = atomicAdd(&hist[ent.x & fmask], 1u);
        if (MODE == 1) acc += pos; else if (MODE == 5) entries[lo + i] = ent; else entries[lo + (pos % size)] = ent;
    }
    if (MODE == 1 && acc == 0xdeadbeef) sink[0] = 1;
}
template <int MODE> float run(const uint2* tmp, uint2* ent, const uint32_t* off, int nc, int fb, uint32_t* sink) {
    size_t smem = FS_CAP * sizeof(uint2) + ((size_t(1) << fb) + FS_THREADS) * 4;
    cudaFuncSetAttribute(fine<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    fine<MODE><<<nc, FS_THREADS, smem>>>(tmp, ent, off, fb, sink);
    cudaEventRecord(a);
    for (int i = 0; i < 5; ++i) fine<MODE><<<nc, FS_THREADS, smem>>>(tmp, ent, off, fb, sink);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("mode %d: %.3f ms  (%s)\n", MODE, ms / 5, cudaGetErrorString(cudaGetLastError()));
    return ms / 5;
}
int main() {
    const int nc = 8192, per = 6656, fb = 8;
    size_t n = (size_t)nc * per;
    std::vector<uint2> h(n); std::vector<uint32_t> off(nc + 1);
    uint64_t s = 88172645463325252ull;
    for (size_t i = 0; i < n; ++i) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = make_uint2(((uint32_t)(i / per) << fb) | (uint32_t)(s & 255), (uint32_t)(s >> 32)); }
    for (int i = 0; i <= nc; ++i) off[i] = (uint32_t)((size_t)i * per);
    uint2 *tmp, *ent; uint32_t *doff, *sink;
    cudaMalloc(&tmp, n * 8); cudaMalloc(&ent, n * 8); cudaMalloc(&doff, (nc + 1) * 4); cudaMalloc(&sink, 4);
    cudaMemcpy(tmp, h.data(), n * 8, cudaMemcpyHostToDevice); cudaMemcpy(doff, off.data(), (nc + 1) * 4, cudaMemcpyHostToDevice);
    printf("entries %zu (%.0f MB)\n", n, n * 8 / 1e6);
    run<4>(tmp, ent, doff, nc, fb, sink);  // staging only
    run<2>(tmp, ent, doff, nc, fb, sink);  // + count + scan
    run<3>(tmp, ent, doff, nc, fb, sink);  // no count pass, rank + scattered store
    run<1>(tmp, ent, doff, nc, fb, sink);  // full but no stores
    run<5>(tmp, ent, doff, nc, fb, sink);  // full, stores in input order (coalesced)
    run<0>(tmp, ent, doff, nc, fb, sink);  // full
    return 0;
}
