// G1-valued kernels around the SRS:
//   * srs_gen  — Crs::<Bn256, CrsForMonomialForm>::crs_42 (src/plonk.rs:41,47): out[i] = [tau^i] G.
//   * ec_intt  — Crs::<_, CrsForLagrangeForm>::from_powers (src/plonk.rs:179-185; `plonkit dump-lagrange`,
//                src/bin/main.rs:360-381): an inverse FFT over G1 POINTS turning [tau^j] G into [L_i(tau)] G.
//                Every butterfly is a 254-bit scalar multiplication plus a point add/sub.
// Both are setup-time tools, not part of the per-proof path.  First version: one kernel per radix-2 stage over an
// XYZZ array in global memory, plain double-and-add for the twiddle multiplication.
#include "msm.cuh"
#include "ntt.cuh"

namespace pk {

static inline dim3 grid1d(size_t n, int block) { return dim3((unsigned)((n + block - 1) / block)); }

// k * P for a canonical 254-bit scalar, left-to-right double-and-add
__device__ __forceinline__ g1_xyzz_t scalar_mul(const g1_xyzz_t& P, const fr_t& k_canonical) {
    g1_xyzz_t r = g1_xyzz_t::infinity();
    bool started = false;
    for (int i = 253; i >= 0; --i) {
        if (started) r = r.dbl();
        if ((k_canonical.v[i >> 5] >> (i & 31)) & 1) {
            r = started ? r.add(P) : P;
            started = true;
        }
    }
    return r;
}

__global__ void __launch_bounds__(128) srs_gen_kernel(g1_affine_t* out, fr_t tau, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t k = tau.pow_u64(i).from_mont();
    g1_affine_t g;
    g.x = fq_t::from_u32(1);
    g.y = fq_t::from_u32(2);
    g1_affine_t a = scalar_mul(g1_xyzz_t::from_affine(g), k).to_affine();
    if (!a.is_inf()) { a.x = a.x.from_mont(); a.y = a.y.from_mont(); }
    st_affine(out + i, a);
}

void srs_gen(pk_ctx* ctx, uint64_t n, uint64_t tau, uint64_t* out_xy) {
    PK_REQUIRE(n <= (uint64_t(1) << 26), PK_ERR_DEGREE_TOO_LARGE, "SRS larger than 2^26 (SETUP_MAX_POW2, src/plonk.rs:27)");
    DevBuf<g1_affine_t> d(n);
    fr_t t = fr_t::zero();
    t.v[0] = (uint32_t)tau;
    t.v[1] = (uint32_t)(tau >> 32);
    t = t.to_mont();
    srs_gen_kernel<<<grid1d(n, 128), 128, 0, ctx->stream>>>(d.p, t, n);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaMemcpyAsync(out_xy, d.p, n * sizeof(g1_affine_t), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
}

// a[brev(i)] = XYZZ(base_i)
__global__ void ec_load_bitrev_kernel(const g1_affine_t* bases, g1_xyzz_t* a, int log_n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >> log_n) return;
    size_t r = log_n ? (size_t)(__brev((unsigned)i) >> (32 - log_n)) : 0;
    st_xyzz(a + r, g1_xyzz_t::from_affine(ldg_affine(bases + i)));
}
// decimation-in-time stage s with inverse twiddles w^{-e} = -w^{n/2-e}
__global__ void __launch_bounds__(128) ec_stage_kernel(g1_xyzz_t* a, const fr_t* tw, int tw_shift, int log_n, int s) {
    size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t half_n = size_t(1) << (log_n - 1);
    if (u >= half_n) return;
    const size_t m = size_t(1) << s;
    const size_t lo = u & (m - 1);
    const size_t i0 = ((u >> s) << (s + 1)) | lo, i1 = i0 + m;
    const size_t e = lo << (log_n - 1 - s);
    g1_xyzz_t x = ld_xyzz(a + i0), y = ld_xyzz(a + i1);
    g1_xyzz_t t;
    if (e == 0) t = y;
    else t = scalar_mul(y, ldg_fp(tw + ((half_n - e) << tw_shift)).from_mont()).neg();
    st_xyzz(a + i0, x.add(t));
    st_xyzz(a + i1, x.add(t.neg()));
}
__global__ void __launch_bounds__(128) ec_finish_kernel(const g1_xyzz_t* a, g1_affine_t* out, fr_t ninv_canonical, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine_t p = scalar_mul(ld_xyzz(a + i), ninv_canonical).to_affine();
    if (!p.is_inf()) { p.x = p.x.from_mont(); p.y = p.y.from_mont(); }
    st_affine(out + i, p);
}

void ec_intt(pk_ctx* ctx, uint32_t log_n, uint64_t* out_xy) {
    PK_REQUIRE(log_n <= 26, PK_ERR_DEGREE_TOO_LARGE, "domain larger than 2^26");
    SrsTables* srs = ctx->srs;
    const size_t n = size_t(1) << log_n;
    PK_REQUIRE(srs && srs->n >= n, PK_ERR_DEGREE_TOO_LARGE, "SRS smaller than the requested Lagrange basis");
    cudaStream_t st = ctx->stream;
    DevBuf<g1_xyzz_t> a(n);
    DevBuf<g1_affine_t> out(n);
    ec_load_bitrev_kernel<<<grid1d(n, 256), 256, 0, st>>>(srs->table.p, a.p, (int)log_n);
    ctx->prof.kernel_launches++;
    if (log_n) {
        ensure_twiddles(ctx, (int)log_n);
        DomainCache* dc = ctx->domains;
        for (int s = 0; s < (int)log_n; ++s) {
            ec_stage_kernel<<<grid1d(n / 2, 128), 128, 0, st>>>(a.p, dc->tw.p, dc->tw_log - (int)log_n, (int)log_n, s);
            ctx->prof.kernel_launches++;
        }
    }
    fr_t ninv = fr_t::from_u32(2).inverse().pow_u64(log_n).from_mont();
    ec_finish_kernel<<<grid1d(n, 128), 128, 0, st>>>(a.p, out.p, ninv, n);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaMemcpyAsync(out_xy, out.p, n * sizeof(g1_affine_t), cudaMemcpyDeviceToHost, st));
    PK_CUDA(cudaStreamSynchronize(st));
}

}  // namespace pk
