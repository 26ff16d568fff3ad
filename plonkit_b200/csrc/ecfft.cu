// G1-valued kernels around the SRS:
//   * srs_gen  — Crs::<Bn256, CrsForMonomialForm>::crs_42 (src/plonk.rs:41,47): out[i] = [tau^i] G.
//   * ec_intt  — Crs::<_, CrsForLagrangeForm>::from_powers (src/plonk.rs:179-185; `plonkit dump-lagrange`,
//                src/bin/main.rs:360-381): an inverse FFT over G1 POINTS turning [tau^j] G into [L_i(tau)] G.
//                Every butterfly is a 254-bit scalar multiplication plus a point add/sub.
// Both are setup-time tools, not part of the per-proof path.  First version: one kernel per radix-2 stage over an
// XYZZ array in global memory; the twiddle multiplication is the GLV double-and-add of ecmul.cuh.
#include "ecmul.cuh"
#include "msm.cuh"
#include "ntt.cuh"

namespace pk {

static inline dim3 grid1d(size_t n, int block) { return dim3((unsigned)((n + block - 1) / block)); }

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
// decimation-in-time stage s with inverse twiddles w^{-e} = -w^{n/2-e}.  Bounded to 128 registers (4 resident blocks per
// SM); 96 and 168 registers were measured within 4 % of it (2^20: 513 / 507 / 526 ms).
__global__ void __launch_bounds__(128, 4) ec_stage_kernel(g1_xyzz_t* a, const fr_t* tw, int tw_shift, int log_n, int s) {
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

// ---------------------------------------------------------------- device-resident pieces of a four-step EC (i)NTT across GPUs
// (SURVEY.md §8e row "EC-iNTT": Crs::from_powers of a 2^25 key split over 8 GPUs, BASELINE configs[3]).  The caller
// (plonkit_b200/dist.py) owns the buffers and moves them between ranks; these run the local steps.
__global__ void ec_from_affine_kernel(const g1_affine_t* in, g1_xyzz_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_affine_t p = ldg_affine(in + i);
    if (!p.is_inf()) { p.x = p.x.to_mont(); p.y = p.y.to_mont(); }
    st_xyzz(out + i, g1_xyzz_t::from_affine(p));
}
__global__ void ec_rows_bitrev_kernel(const g1_xyzz_t* src, g1_xyzz_t* dst, int log_len) {  // blockIdx.y = row
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >> log_len) return;
    size_t r = log_len ? (size_t)(__brev((unsigned)i) >> (32 - log_len)) : 0;
    const size_t row = (size_t)blockIdx.y << log_len;
    st_xyzz(dst + row + r, ld_xyzz(src + row + i));
}
// one radix-2 decimation-in-time stage of `rows` independent transforms of length 2^log_len (blockIdx.y = row);
// forward twiddle w^e, inverse twiddle w^{-e} = -w^{len/2 - e}
__global__ void __launch_bounds__(128) ec_rows_stage_kernel(g1_xyzz_t* a, const fr_t* tw, int tw_shift, int log_len, int s, int inverse) {
    size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t half = size_t(1) << (log_len - 1);
    if (u >= half) return;
    a += (size_t)blockIdx.y << log_len;
    const size_t m = size_t(1) << s;
    const size_t lo = u & (m - 1);
    const size_t i0 = ((u >> s) << (s + 1)) | lo, i1 = i0 + m;
    const size_t e = lo << (log_len - 1 - s);
    g1_xyzz_t x = ld_xyzz(a + i0), y = ld_xyzz(a + i1);
    g1_xyzz_t t;
    if (e == 0) t = y;
    else if (inverse) t = scalar_mul(y, ldg_fp(tw + ((half - e) << tw_shift)).from_mont()).neg();
    else t = scalar_mul(y, ldg_fp(tw + (e << tw_shift)).from_mont());
    st_xyzz(a + i0, x.add(t));
    st_xyzz(a + i1, x.add(t.neg()));
}
// a[r][c] <- w^{+-(row0 + r) * c} * a[r][c]   (w: primitive 2^log_total-th root of unity)
__global__ void __launch_bounds__(128) ec_twiddle_rows_kernel(g1_xyzz_t* a, size_t rows, size_t cols, fr_t w, size_t row0, int log_total) {
    size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t r = blockIdx.y;
    if (c >= cols || r >= rows) return;
    const uint64_t mask = (uint64_t(1) << log_total) - 1;
    const uint64_t e = ((uint64_t)(row0 + r) * (uint64_t)c) & mask;
    if (e == 0) return;
    st_xyzz(a + r * cols + c, scalar_mul(ld_xyzz(a + r * cols + c), w.pow_u64(e).from_mont()));
}

void ec_dev_from_affine(pk_ctx* ctx, const g1_affine_t* in_canonical, g1_xyzz_t* out, size_t n) {
    ec_from_affine_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(in_canonical, out, n);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
}
void ec_dev_ntt_rows(pk_ctx* ctx, g1_xyzz_t* data, int log_len, size_t rows, bool inverse) {
    PK_REQUIRE(rows >= 1 && rows <= 65535, PK_ERR_INVALID, "row batch too large");
    if (log_len == 0) return;
    const size_t len = size_t(1) << log_len;
    ensure_twiddles(ctx, log_len);
    DomainCache* dc = ctx->domains;
    DevBuf<g1_xyzz_t> tmp(rows * len);
    ec_rows_bitrev_kernel<<<dim3((unsigned)((len + 255) / 256), (unsigned)rows), 256, 0, ctx->stream>>>(data, tmp.p, log_len);
    for (int s = 0; s < log_len; ++s)
        ec_rows_stage_kernel<<<dim3((unsigned)((len / 2 + 127) / 128), (unsigned)rows), 128, 0, ctx->stream>>>(
            tmp.p, dc->tw.p, dc->tw_log - log_len, log_len, s, inverse ? 1 : 0);
    ctx->prof.kernel_launches += 1 + log_len;
    PK_CUDA(cudaMemcpyAsync(data, tmp.p, rows * len * sizeof(g1_xyzz_t), cudaMemcpyDeviceToDevice, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaGetLastError());
}
void ec_dev_twiddle_rows(pk_ctx* ctx, g1_xyzz_t* a, size_t rows, size_t cols, int log_total, size_t row0, bool inverse) {
    PK_REQUIRE(rows >= 1 && rows <= 65535, PK_ERR_INVALID, "row batch too large");
    fr_t w = host_root_of_unity(log_total);
    if (inverse) w = w.inverse();
    ec_twiddle_rows_kernel<<<dim3((unsigned)((cols + 127) / 128), (unsigned)rows), 128, 0, ctx->stream>>>(a, rows, cols, w, row0, log_total);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
}
// out[i] = affine(2^-log_scale * in[i]), canonical limbs
void ec_dev_to_affine(pk_ctx* ctx, const g1_xyzz_t* in, g1_affine_t* out_canonical, size_t n, int log_scale) {
    fr_t scale = fr_t::from_u32(2).inverse().pow_u64(log_scale).from_mont();
    ec_finish_kernel<<<grid1d(n, 128), 128, 0, ctx->stream>>>(in, out_canonical, scale, n);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
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
