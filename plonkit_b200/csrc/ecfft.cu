// G1-valued kernels around the SRS:
//   * srs_gen  — Crs::<Bn256, CrsForMonomialForm>::crs_42 (src/plonk.rs:41,47): out[i] = [tau^i] G.
//   * ec_intt  — Crs::<_, CrsForLagrangeForm>::from_powers (src/plonk.rs:179-185; `plonkit dump-lagrange`,
//                src/bin/main.rs:360-381): an inverse FFT over G1 POINTS turning [tau^j] G into [L_i(tau)] G.
//                Every butterfly is a 254-bit scalar multiplication plus a point add/sub.
// Both are setup-time tools, not part of the per-proof path.  First version: one kernel per radix-2 stage over an
// XYZZ array in global memory; the twiddle multiplication is the GLV double-and-add of ecmul.cuh.
#include <cstdlib>
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
// decimation-in-time stage s with inverse twiddles w^{-e} = -w^{n/2-e}.
//
// The 1/n of the inverse transform rides on the twiddles instead of costing one more multiplication per point at the end:
// c F = c F_even + (c w^i) F_odd, so at every stage only the FIRST block (the all-even chain) uses twiddles scaled by
// c = 1/n, and the single leaf under it (a[0]) is scaled before stage 0: log n + 1 extra multiplications instead of n.
//
// Stages with fewer than 32 distinct twiddles run in twiddle-major thread order: a warp then holds ONE twiddle, and the
// e == 0 warps skip the multiplication instead of idling through their neighbours' (a point is 128 B, so the order of the
// butterflies within a stage does not change the memory transactions).
template <int MINB>
__global__ void __launch_bounds__(128, MINB) ec_stage_kernel(g1_xyzz_t* a, const fr_t* tw, int tw_shift, int log_n, int s, fr_t scale) {
    size_t u = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t half_n = size_t(1) << (log_n - 1);
    if (u >= half_n) return;
    const size_t m = size_t(1) << s;
    size_t lo, blk;
    if (s < 5) { lo = u >> (log_n - 1 - s); blk = u & ((size_t(1) << (log_n - 1 - s)) - 1); }
    else { lo = u & (m - 1); blk = u >> s; }
    const size_t i0 = (blk << (s + 1)) | lo, i1 = i0 + m;
    const size_t e = lo << (log_n - 1 - s);
    g1_xyzz_t x = ld_xyzz(a + i0), y = ld_xyzz(a + i1);
    fr_t k = scale;
    if (e) {
        k = ldg_fp(tw + ((half_n - e) << tw_shift));
        if (blk == 0) k = k * scale;
    }
    g1_xyzz_t t = (e || blk == 0) ? scalar_mul(y, k.from_mont()) : y;
    if (e) t = t.neg();
    st_xyzz(a + i0, x.add(t));
    st_xyzz(a + i1, x.add(t.neg()));
}
// out[i] = affine(scale * a[i]) (has_scale) or affine(a[i]), canonical limbs
__global__ void __launch_bounds__(128) ec_finish_kernel(const g1_xyzz_t* a, g1_affine_t* out, fr_t scale_canonical, int has_scale, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    g1_xyzz_t v = ld_xyzz(a + i);
    if (has_scale) v = scalar_mul(v, scale_canonical);
    g1_affine_t p = v.to_affine();
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
    size_t lo, blk;                                  // twiddle-major order in the first stages, as in ec_stage_kernel
    if (s < 5) { lo = u >> (log_len - 1 - s); blk = u & ((size_t(1) << (log_len - 1 - s)) - 1); }
    else { lo = u & (m - 1); blk = u >> s; }
    const size_t i0 = (blk << (s + 1)) | lo, i1 = i0 + m;
    const size_t e = lo << (log_len - 1 - s);
    g1_xyzz_t x = ld_xyzz(a + i0), y = ld_xyzz(a + i1);
    g1_xyzz_t t = y;
    if (e) {
        t = scalar_mul(y, ldg_fp(tw + ((inverse ? half - e : e) << tw_shift)).from_mont());
        if (inverse) t = t.neg();
    }
    st_xyzz(a + i0, x.add(t));
    st_xyzz(a + i1, x.add(t.neg()));
}
// a[r][c] <- scale * w^{+-(row0 + r) * c} * a[r][c]   (w: primitive 2^log_total-th root of unity; scale in Montgomery
// form, has_scale == 0 means 1: the 1/N of a four-step inverse transform costs nothing extra when it rides on this step)
__global__ void __launch_bounds__(128) ec_twiddle_rows_kernel(g1_xyzz_t* a, size_t rows, size_t cols, fr_t w, size_t row0, int log_total,
                                                              fr_t scale, int has_scale) {
    size_t c = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    size_t r = blockIdx.y;
    if (c >= cols || r >= rows) return;
    const uint64_t mask = (uint64_t(1) << log_total) - 1;
    const uint64_t e = ((uint64_t)(row0 + r) * (uint64_t)c) & mask;
    if (e == 0 && !has_scale) return;
    fr_t k = w.pow_u64(e);
    if (has_scale) k = k * scale;
    st_xyzz(a + r * cols + c, scalar_mul(ld_xyzz(a + r * cols + c), k.from_mont()));
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
// mode 0: forward twiddles; 1: inverse twiddles; 2: inverse twiddles times 2^-log_total (the scale of the whole transform)
void ec_dev_twiddle_rows(pk_ctx* ctx, g1_xyzz_t* a, size_t rows, size_t cols, int log_total, size_t row0, int mode) {
    PK_REQUIRE(rows >= 1 && rows <= 65535, PK_ERR_INVALID, "row batch too large");
    PK_REQUIRE(mode >= 0 && mode <= 2, PK_ERR_INVALID, "twiddle mode must be 0 (forward), 1 (inverse) or 2 (inverse, scaled)");
    fr_t w = host_root_of_unity(log_total);
    if (mode) w = w.inverse();
    const fr_t scale = mode == 2 ? fr_t::from_u32(2).inverse().pow_u64(log_total) : fr_t::one();
    ec_twiddle_rows_kernel<<<dim3((unsigned)((cols + 127) / 128), (unsigned)rows), 128, 0, ctx->stream>>>(a, rows, cols, w, row0, log_total,
                                                                                                       scale, mode == 2 ? 1 : 0);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
}
// out[i] = affine(2^-log_scale * in[i]), canonical limbs
void ec_dev_to_affine(pk_ctx* ctx, const g1_xyzz_t* in, g1_affine_t* out_canonical, size_t n, int log_scale) {
    fr_t scale = fr_t::from_u32(2).inverse().pow_u64(log_scale).from_mont();
    ec_finish_kernel<<<grid1d(n, 128), 128, 0, ctx->stream>>>(in, out_canonical, scale, log_scale ? 1 : 0, n);
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
        const fr_t ninv = fr_t::from_u32(2).inverse().pow_u64(log_n);      // rides on the twiddles of the all-even chain
        ec_twiddle_rows_kernel<<<1, 128, 0, st>>>(a.p, 1, 1, fr_t::one(), 0, (int)log_n, ninv, 1);   // the leaf a[0] *= 1/n
        ctx->prof.kernel_launches++;
        // resident blocks per SM the stage kernel is compiled for: 2 (242 registers, nothing spilled) measured 380 ms at 2^20,
        // 4 (128 registers, 1.9 KB spilled) 397 ms, 3 (168 registers) 415 ms; PK_EC_MINB selects the others for re-measurement
        static const int minb = getenv("PK_EC_MINB") ? atoi(getenv("PK_EC_MINB")) : 2;
        for (int s = 0; s < (int)log_n; ++s) {
            if (minb == 3) ec_stage_kernel<3><<<grid1d(n / 2, 128), 128, 0, st>>>(a.p, dc->tw.p, dc->tw_log - (int)log_n, (int)log_n, s, ninv);
            else if (minb == 2) ec_stage_kernel<2><<<grid1d(n / 2, 128), 128, 0, st>>>(a.p, dc->tw.p, dc->tw_log - (int)log_n, (int)log_n, s, ninv);
            else ec_stage_kernel<4><<<grid1d(n / 2, 128), 128, 0, st>>>(a.p, dc->tw.p, dc->tw_log - (int)log_n, (int)log_n, s, ninv);
            ctx->prof.kernel_launches++;
        }
    }
    ec_finish_kernel<<<grid1d(n, 128), 128, 0, st>>>(a.p, out.p, fr_t::one(), 0, n);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
    PK_CUDA(cudaMemcpyAsync(out_xy, out.p, n * sizeof(g1_affine_t), cudaMemcpyDeviceToHost, st));
    PK_CUDA(cudaStreamSynchronize(st));
}

}  // namespace pk
