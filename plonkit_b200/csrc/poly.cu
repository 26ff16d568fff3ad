// Polynomial glue kernels for the PLONK prover (see poly.cuh).  All of it is streaming 32-byte-element work:
// each kernel reads/writes whole arrays with 128-bit accesses; grids are sized from the element count.
#include "ntt.cuh"
#include "poly.cuh"

namespace pk {

static inline dim3 grid1d(size_t n, int block) { return dim3((unsigned)((n + block - 1) / block)); }
static const int DOT_BLOCKS = 592;  // 4 x 148 SMs


// ---------------------------------------------------------------- small helpers
__device__ __forceinline__ fr_t omega_pow(const fr_t* tw, int tw_shift, int log_n, size_t j) {
    const size_t half = size_t(1) << (log_n - 1);
    fr_t w = ldg_fp(tw + ((j & (half - 1)) << tw_shift));
    return (j & half) ? w.neg() : w;
}
__device__ __forceinline__ uint32_t brev_n(uint32_t x, int log_n) { return log_n ? __brev(x) >> (32 - log_n) : 0; }

template <bool MUL> __device__ __forceinline__ fr_t scan_op(const fr_t& a, const fr_t& b) { return MUL ? a * b : a + b; }
template <bool MUL> __device__ __forceinline__ fr_t scan_ident() { return MUL ? fr_t::one() : fr_t::zero(); }

// ---------------------------------------------------------------- fill / powers / lincomb
__global__ void fill_fr_kernel(fr_t* out, fr_t c, size_t n) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < n) st_fp(out + k, c);
}
void fr_fill(pk_ctx* ctx, fr_t* out, const fr_t& c, size_t n) {
    fill_fr_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(out, c, n);
    ctx->prof.kernel_launches++;
}

struct PowTable { fr_t p[29]; };  // base^(2^k)
__global__ void powers_kernel(fr_t* out, PowTable t, size_t first, size_t n) {  // out[i] = base^(first + i)
    size_t start = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 16;
    if (start >= n) return;
    fr_t x = fr_t::one();
    const size_t e = first + start;
#pragma unroll 1
    for (int k = 0; k < 29; ++k)
        if ((e >> k) & 1) x = x * t.p[k];
    const fr_t base = t.p[0];
    for (int r = 0; r < 16 && start + r < n; ++r) {
        st_fp(out + start + r, x);
        x = x * base;
    }
}
void poly_powers_from(pk_ctx* ctx, fr_t* out, const fr_t& base, size_t first, size_t n) {
    if (!n) return;
    PK_REQUIRE(first + n <= (size_t(1) << 29), PK_ERR_INVALID, "power table too long");
    PowTable t;
    t.p[0] = base;
    for (int k = 1; k < 29; ++k) t.p[k] = t.p[k - 1].sqr();
    powers_kernel<<<grid1d((n + 15) / 16, 128), 128, 0, ctx->stream>>>(out, t, first, n);
    ctx->prof.kernel_launches++;
}
void poly_powers(pk_ctx* ctx, fr_t* out, const fr_t& base, size_t n) { poly_powers_from(ctx, out, base, 0, n); }

struct LinArgs { const fr_t* in[16]; fr_t coef[16]; int n_terms; };
__global__ void lincomb_kernel(fr_t* out, LinArgs a, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t acc = ld_fp(a.in[0] + i) * a.coef[0];
    for (int k = 1; k < a.n_terms; ++k) acc = acc + ld_fp(a.in[k] + i) * a.coef[k];
    st_fp(out + i, acc);
}
void poly_lincomb(pk_ctx* ctx, fr_t* out, int nterms, const fr_t* const* in, const fr_t* coef, size_t n) {
    PK_REQUIRE(nterms >= 1 && nterms <= 16, PK_ERR_INVALID, "lincomb arity");
    LinArgs a;
    for (int k = 0; k < nterms; ++k) { a.in[k] = in[k]; a.coef[k] = coef[k]; }
    a.n_terms = nterms;
    lincomb_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(out, a, n);
    ctx->prof.kernel_launches++;
}

// ---------------------------------------------------------------- batched dot products (evaluate_at)
struct DotArgs { const fr_t* poly[16]; const fr_t* pow[16]; };
__device__ __forceinline__ void block_reduce_add(fr_t& v, fr_t* sh, fr_t* out) {
    const unsigned tid = threadIdx.x;
    sh[tid] = v;
    __syncthreads();
    for (unsigned d = blockDim.x >> 1; d > 0; d >>= 1) {
        if (tid < d) sh[tid] = sh[tid] + sh[tid + d];
        __syncthreads();
    }
    if (tid == 0) st_fp(out, sh[0]);
}
__global__ void __launch_bounds__(256) dot_kernel(DotArgs a, size_t n, fr_t* partial) {
    __shared__ fr_t sh[256];
    const fr_t* p = a.poly[blockIdx.y];
    const fr_t* w = a.pow[blockIdx.y];
    fr_t acc = fr_t::zero();
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        acc = acc + ld_fp(p + i) * ld_fp(w + i);
    block_reduce_add(acc, sh, partial + blockIdx.y * gridDim.x + blockIdx.x);
}
__global__ void __launch_bounds__(256) dot_final_kernel(const fr_t* partial, int nblk, fr_t* out) {
    __shared__ fr_t sh[256];
    fr_t acc = fr_t::zero();
    for (int i = threadIdx.x; i < nblk; i += blockDim.x) acc = acc + ld_fp(partial + blockIdx.x * nblk + i);
    block_reduce_add(acc, sh, out + blockIdx.x);
}
void poly_dot_batch_dev(pk_ctx* ctx, int npoly, const fr_t* const* polys, const fr_t* const* pows, size_t n, fr_t* out_dev) {
    PK_REQUIRE(npoly >= 1 && npoly <= 16, PK_ERR_INVALID, "dot batch arity");
    PolyScratch* sc = poly_scratch(ctx);
    sc->dot_partial.ensure(16 * DOT_BLOCKS);
    DotArgs a;
    for (int k = 0; k < npoly; ++k) { a.poly[k] = polys[k]; a.pow[k] = pows[k]; }
    int nblk = (int)((n + 255) / 256);
    if (nblk > DOT_BLOCKS) nblk = DOT_BLOCKS;
    if (nblk < 1) nblk = 1;
    dot_kernel<<<dim3(nblk, npoly), 256, 0, ctx->stream>>>(a, n, sc->dot_partial.p);
    dot_final_kernel<<<npoly, 256, 0, ctx->stream>>>(sc->dot_partial.p, nblk, out_dev);
    ctx->prof.kernel_launches += 2;
}
void poly_dot_batch(pk_ctx* ctx, int npoly, const fr_t* const* polys, const fr_t* const* pows, size_t n, fr_t* results_host) {
    PolyScratch* sc = poly_scratch(ctx);
    sc->dot_out.ensure(16);
    poly_dot_batch_dev(ctx, npoly, polys, pows, n, sc->dot_out.p);
    fr_t* host = reinterpret_cast<fr_t*>(ctx->pinned);
    PK_CUDA(cudaMemcpyAsync(host, sc->dot_out.p, npoly * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < npoly; ++k) results_host[k] = host[k];
}

// ---------------------------------------------------------------- scans (3 phases: tile reduce, spine, tile apply)
static const int SCAN_THREADS = 256, SCAN_ITEMS = 8, SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

template <bool MUL, bool REV> __global__ void __launch_bounds__(256) scan_reduce_kernel(const fr_t* in, fr_t* agg, size_t n) {
    __shared__ fr_t sh[SCAN_THREADS];
    const size_t base = blockIdx.x * (size_t)SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    fr_t acc = scan_ident<MUL>();
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        size_t i = base + r;
        if (i < n) acc = scan_op<MUL>(acc, ld_fp(in + (REV ? n - 1 - i : i)));
    }
    const unsigned tid = threadIdx.x;
    sh[tid] = acc;
    __syncthreads();
    for (unsigned d = SCAN_THREADS >> 1; d > 0; d >>= 1) {
        if (tid < d) sh[tid] = scan_op<MUL>(sh[tid], sh[tid + d]);
        __syncthreads();
    }
    if (tid == 0) st_fp(agg + blockIdx.x, sh[0]);
}
// in-place exclusive scan of m aggregates by one block of 1024 threads
template <bool MUL> __global__ void __launch_bounds__(1024) scan_spine_kernel(fr_t* agg, size_t m) {
    __shared__ fr_t sh[1024];
    const unsigned tid = threadIdx.x;
    const size_t per = (m + 1023) / 1024;
    size_t lo = tid * per, hi = lo + per;
    if (hi > m) hi = m;
    fr_t tot = scan_ident<MUL>();
    for (size_t i = lo; i < hi; ++i) tot = scan_op<MUL>(tot, ld_fp(agg + i));
    sh[tid] = tot;
    __syncthreads();
    for (unsigned d = 1; d < 1024; d <<= 1) {
        fr_t v = scan_ident<MUL>();
        if (tid >= d) v = sh[tid - d];
        __syncthreads();
        if (tid >= d) sh[tid] = scan_op<MUL>(v, sh[tid]);
        __syncthreads();
    }
    fr_t run = tid ? sh[tid - 1] : scan_ident<MUL>();
    for (size_t i = lo; i < hi; ++i) {
        fr_t x = ld_fp(agg + i);
        st_fp(agg + i, run);
        run = scan_op<MUL>(run, x);
    }
}
template <bool MUL, bool REV>
__global__ void __launch_bounds__(256) scan_apply_kernel(const fr_t* in, fr_t* out, const fr_t* agg_excl, size_t n) {
    __shared__ fr_t sh[SCAN_THREADS];
    const size_t base = blockIdx.x * (size_t)SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    fr_t x[SCAN_ITEMS];
    fr_t tot = scan_ident<MUL>();
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        size_t i = base + r;
        x[r] = (i < n) ? ld_fp(in + (REV ? n - 1 - i : i)) : scan_ident<MUL>();
        tot = scan_op<MUL>(tot, x[r]);
    }
    const unsigned tid = threadIdx.x;
    sh[tid] = tot;
    __syncthreads();
    for (unsigned d = 1; d < SCAN_THREADS; d <<= 1) {
        fr_t v = scan_ident<MUL>();
        if (tid >= d) v = sh[tid - d];
        __syncthreads();
        if (tid >= d) sh[tid] = scan_op<MUL>(v, sh[tid]);
        __syncthreads();
    }
    fr_t run = ld_fp(agg_excl + blockIdx.x);
    if (tid) run = scan_op<MUL>(run, sh[tid - 1]);
#pragma unroll
    for (int r = 0; r < SCAN_ITEMS; ++r) {
        size_t i = base + r;
        run = scan_op<MUL>(run, x[r]);
        if (i < n) st_fp(out + (REV ? n - 1 - i : i), run);
    }
}
template <bool MUL, bool REV> static void scan_impl(pk_ctx* ctx, const fr_t* in, fr_t* out, size_t n) {
    PolyScratch* sc = poly_scratch(ctx);
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    sc->scan_agg.ensure(tiles);
    scan_reduce_kernel<MUL, REV><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, sc->scan_agg.p, n);
    scan_spine_kernel<MUL><<<1, 1024, 0, ctx->stream>>>(sc->scan_agg.p, tiles);
    scan_apply_kernel<MUL, REV><<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, out, sc->scan_agg.p, n);
    ctx->prof.kernel_launches += 3;
}
void poly_scan(pk_ctx* ctx, bool mul, bool reverse, const fr_t* in, fr_t* out, size_t n) {
    if (!n) return;
    if (mul) { if (reverse) scan_impl<true, true>(ctx, in, out, n); else scan_impl<true, false>(ctx, in, out, n); }
    else { if (reverse) scan_impl<false, true>(ctx, in, out, n); else scan_impl<false, false>(ctx, in, out, n); }
    PK_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- division by (X - z)
// q_k = z^{-(k+1)} * sum_{j > k} p_j z^j   (k = 0..n-2), q_{n-1} = 0
__global__ void mul_pointwise_kernel(const fr_t* a, const fr_t* b, fr_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) st_fp(out + i, ld_fp(a + i) * ld_fp(b + i));
}
__global__ void divide_finish_kernel(const fr_t* suffix, const fr_t* zinvpow, fr_t* q, size_t n) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (k + 1 < n) st_fp(q + k, ld_fp(suffix + k + 1) * ld_fp(zinvpow + k + 1));
    else st_fp(q + k, fr_t::zero());
}
void fr_mul_pointwise(pk_ctx* ctx, const fr_t* a, const fr_t* b, fr_t* out, size_t n) {
    mul_pointwise_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(a, b, out, n);
    ctx->prof.kernel_launches++;
}
// chunk [lo, lo + len) of the same division on a rank of the sharded prover: suffix[i] = sum_{j >= i, j in chunk} p_j z^j,
// then q[k] = (suffix[k + 1] + carry) * z^-(k + 1) with carry = the sum over the chunks above (zi[i] = z^-(lo + i + 1))
void poly_divide_linear_chunk_scan(pk_ctx* ctx, const fr_t* p, const fr_t* zpow, fr_t* suffix, size_t len) {
    mul_pointwise_kernel<<<grid1d(len, 256), 256, 0, ctx->stream>>>(p, zpow, suffix, len);
    ctx->prof.kernel_launches++;
    poly_scan(ctx, false, true, suffix, suffix, len);
}
__global__ void divide_finish_chunk_kernel(const fr_t* suffix, const fr_t* zi, fr_t carry, fr_t* q, size_t len) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= len) return;
    fr_t sfx = i + 1 < len ? ld_fp(suffix + i + 1) + carry : carry;
    st_fp(q + i, sfx * ld_fp(zi + i));
}
void poly_divide_linear_chunk_finish(pk_ctx* ctx, const fr_t* suffix, const fr_t* zi, const fr_t& carry, fr_t* q, size_t len) {
    divide_finish_chunk_kernel<<<grid1d(len, 256), 256, 0, ctx->stream>>>(suffix, zi, carry, q, len);
    ctx->prof.kernel_launches++;
}
void poly_divide_linear(pk_ctx* ctx, const fr_t* p, const fr_t* zpow, const fr_t* zinvpow, fr_t* q, fr_t* tmp, size_t n) {
    mul_pointwise_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(p, zpow, tmp, n);
    ctx->prof.kernel_launches++;
    poly_scan(ctx, false, true, tmp, tmp, n);
    divide_finish_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(tmp, zinvpow, q, n);
    ctx->prof.kernel_launches++;
}

// ---------------------------------------------------------------- batch inversion (bellman: Polynomial<Values>::batch_inversion)
// No per-chunk Montgomery chains: with zeros mapped to one, out_i = (prod_{j<i} v_j) * (prod_{j>i} v_j) / prod_j v_j is
// two multiplicative scans and ONE field inversion on the host; zeros stay zero, as in bellman.
__global__ void zero_to_one_kernel(const fr_t* in, fr_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_t v = ld_fp(in + i);
    st_fp(out + i, v.is_zero() ? fr_t::one() : v);
}
__global__ void batch_inv_finish_kernel(const fr_t* in, const fr_t* pn, const fr_t* sd, fr_t tinv, fr_t* out, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (ld_fp(in + i).is_zero()) { st_fp(out + i, fr_t::zero()); return; }
    fr_t r = tinv;
    if (i) r = r * ld_fp(pn + i - 1);
    if (i + 1 < n) r = r * ld_fp(sd + i + 1);
    st_fp(out + i, r);
}
void poly_batch_inversion(pk_ctx* ctx, const fr_t* in, fr_t* out, fr_t* tmp_a, fr_t* tmp_b, size_t n) {
    if (!n) return;
    zero_to_one_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(in, tmp_a, n);
    ctx->prof.kernel_launches++;
    poly_scan(ctx, true, true, tmp_a, tmp_b, n);   // sd[i] = prod_{j >= i}
    poly_scan(ctx, true, false, tmp_a, tmp_a, n);  // pn[i] = prod_{j <= i}
    fr_t total;
    PK_CUDA(cudaMemcpyAsync(&total, tmp_a + n - 1, sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    batch_inv_finish_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(in, tmp_a, tmp_b, total.inverse(), out, n);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------- witness -> wire values
__global__ void wire_gather_kernel(const fr_t* vars, const uint32_t* idx, fr_t* nat, fr_t* br, int log_n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // over 4n
    const size_t n = size_t(1) << log_n;
    if (i >= 4 * n) return;
    const size_t c = i >> log_n, row = i & (n - 1);
    fr_t v = ld_fp(vars + idx[i]);
    st_fp(nat + i, v);
    st_fp(br + (c << log_n) + brev_n((uint32_t)row, log_n), v);
}
void wire_gather(pk_ctx* ctx, const fr_t* vars, const uint32_t* idx, fr_t* vals_nat, fr_t* vals_br, int log_n) {
    size_t n4 = size_t(4) << log_n;
    wire_gather_kernel<<<grid1d(n4, 256), 256, 0, ctx->stream>>>(vars, idx, vals_nat, vals_br, log_n);
    ctx->prof.kernel_launches++;
}
__global__ void pi_scatter_kernel(const fr_t* a_nat, fr_t* pi_br, uint32_t ni, int log_n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ni) st_fp(pi_br + brev_n(i, log_n), ld_fp(a_nat + i));
}
void pi_scatter(pk_ctx* ctx, const fr_t* vals_nat_a, fr_t* pi_br, uint32_t num_inputs, int log_n) {
    if (!num_inputs) return;
    pi_scatter_kernel<<<grid1d(num_inputs, 128), 128, 0, ctx->stream>>>(vals_nat_a, pi_br, num_inputs, log_n);
    ctx->prof.kernel_launches++;
}

// q_a a + q_b b + q_c c + q_d d + q_m a b + q_const + q_dnext d(next row) + PI = 0 on rows 0..n-2
__global__ void gate_check_kernel(const fr_t* w, const fr_t* sel, uint32_t ni, int log_n, uint32_t* flag) {
    size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n = size_t(1) << log_n;
    if (r + 1 >= n) return;
    fr_t a = ld_fp(w + r), b = ld_fp(w + n + r), c = ld_fp(w + 2 * n + r), d = ld_fp(w + 3 * n + r);
    fr_t dn = ld_fp(w + 3 * n + r + 1);
    fr_t acc = ld_fp(sel + r) * a + ld_fp(sel + n + r) * b + ld_fp(sel + 2 * n + r) * c + ld_fp(sel + 3 * n + r) * d +
               ld_fp(sel + 4 * n + r) * (a * b) + ld_fp(sel + 5 * n + r) + ld_fp(sel + 6 * n + r) * dn;
    if (r < ni) acc = acc + a;
    if (!acc.is_zero()) atomicOr(flag, 1u);
}
__global__ void gate_check_gated_kernel(const fr_t* w, const fr_t* sel, const uint8_t* gate_type, uint32_t ni, int log_n, uint32_t* flag) {
    size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n = size_t(1) << log_n;
    if (r + 1 >= n) return;
    fr_t a = ld_fp(w + r), b = ld_fp(w + n + r), c = ld_fp(w + 2 * n + r), d = ld_fp(w + 3 * n + r);
    bool ok;
    if (gate_type[r] == 1) {
        ok = (a * a == b) && (b * b == c) && (c * a == d);
    } else {
        fr_t dn = ld_fp(w + 3 * n + r + 1);
        fr_t acc = ld_fp(sel + r) * a + ld_fp(sel + n + r) * b + ld_fp(sel + 2 * n + r) * c + ld_fp(sel + 3 * n + r) * d +
                   ld_fp(sel + 4 * n + r) * (a * b) + ld_fp(sel + 5 * n + r) + ld_fp(sel + 6 * n + r) * dn;
        if (r < ni) acc = acc + a;
        ok = acc.is_zero();
    }
    if (!ok) atomicOr(flag, 1u);
}
bool gate_check_gated(pk_ctx* ctx, const fr_t* vals_nat, const fr_t* sel_vals, const uint8_t* gate_type, uint32_t num_inputs, int log_n) {
    PolyScratch* sc = poly_scratch(ctx);
    sc->flag.ensure(1);
    PK_CUDA(cudaMemsetAsync(sc->flag.p, 0, 4, ctx->stream));
    size_t n = size_t(1) << log_n;
    gate_check_gated_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(vals_nat, sel_vals, gate_type, num_inputs, log_n, sc->flag.p);
    ctx->prof.kernel_launches++;
    uint32_t* h = reinterpret_cast<uint32_t*>(ctx->pinned);
    PK_CUDA(cudaMemcpyAsync(h, sc->flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return *h == 0;
}
bool gate_check(pk_ctx* ctx, const fr_t* vals_nat, const fr_t* sel_vals, uint32_t num_inputs, int log_n) {
    PolyScratch* sc = poly_scratch(ctx);
    sc->flag.ensure(1);
    PK_CUDA(cudaMemsetAsync(sc->flag.p, 0, 4, ctx->stream));
    size_t n = size_t(1) << log_n;
    gate_check_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(vals_nat, sel_vals, num_inputs, log_n, sc->flag.p);
    ctx->prof.kernel_launches++;
    uint32_t* h = reinterpret_cast<uint32_t*>(ctx->pinned);
    PK_CUDA(cudaMemcpyAsync(h, sc->flag.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return *h == 0;
}

// ---------------------------------------------------------------- copy permutation
__device__ __forceinline__ fr_t times_k(const fr_t& x, int col) {  // k = (1, 5, 7, 10)
    if (col == 0) return x;
    fr_t x2 = x.dbl(), x4 = x2.dbl(), x5 = x4 + x;
    if (col == 1) return x5;
    if (col == 2) return x5 + x2;
    return x5.dbl();
}
__global__ void sigma_values_kernel(const uint32_t* target, fr_t* out, const fr_t* tw, int tw_shift, int log_n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n = size_t(1) << log_n;
    if (i >= 4 * n) return;
    uint32_t t = target[i];
    st_fp(out + i, times_k(omega_pow(tw, tw_shift, log_n, t & (n - 1)), (int)(t >> log_n)));
}
void sigma_values(pk_ctx* ctx, const uint32_t* sigma_target, fr_t* sigma_vals, int log_n) {
    ensure_twiddles(ctx, log_n);
    DomainCache* dc = ctx->domains;
    size_t n4 = size_t(4) << log_n;
    sigma_values_kernel<<<grid1d(n4, 256), 256, 0, ctx->stream>>>(sigma_target, sigma_vals, dc->tw.p, dc->tw_log - log_n, log_n);
    ctx->prof.kernel_launches++;
}
__global__ void perm_num_den_kernel(const fr_t* w, const fr_t* sig, fr_t beta, fr_t gamma, fr_t* num, fr_t* den, const fr_t* tw,
                                    int tw_shift, int log_n, size_t lo, size_t len) {  // rows [lo, lo + len) -> num/den[0..len)
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n = size_t(1) << log_n;
    if (i >= len) return;
    const size_t j = lo + i;
    fr_t bw = beta * omega_pow(tw, tw_shift, log_n, j);
    fr_t nn, dd;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        fr_t wv = ld_fp(w + c * n + j) + gamma;
        fr_t a = wv + times_k(bw, c);
        fr_t b = wv + beta * ld_fp(sig + c * n + j);
        nn = c ? nn * a : a;
        dd = c ? dd * b : b;
    }
    st_fp(num + i, nn);
    st_fp(den + i, dd);
}
void perm_num_den_range(pk_ctx* ctx, const fr_t* vals_nat, const fr_t* sigma_vals, const fr_t& beta, const fr_t& gamma, fr_t* num,
                        fr_t* den, int log_n, size_t lo, size_t len) {
    ensure_twiddles(ctx, log_n);
    DomainCache* dc = ctx->domains;
    perm_num_den_kernel<<<grid1d(len, 256), 256, 0, ctx->stream>>>(vals_nat, sigma_vals, beta, gamma, num, den, dc->tw.p,
                                                                   dc->tw_log - log_n, log_n, lo, len);
    ctx->prof.kernel_launches++;
}
void perm_num_den(pk_ctx* ctx, const fr_t* vals_nat, const fr_t* sigma_vals, const fr_t& beta, const fr_t& gamma, fr_t* num,
                  fr_t* den, int log_n) {
    perm_num_den_range(ctx, vals_nat, sigma_vals, beta, gamma, num, den, log_n, 0, size_t(1) << log_n);
}
// rows [lo, lo + len) of Z in natural order from the chunk-local scans: z[j] = pn[j - 1] * sd[j] * factor with pn[lo - 1] = 1
// (the products of the chunks below / above and 1 / prod den are folded into `factor`); z[0] = 1
__global__ void z_finish_chunk_kernel(const fr_t* pn, const fr_t* sd, fr_t factor, fr_t* z, size_t lo, size_t len) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= len) return;
    fr_t v = ld_fp(sd + i) * factor;
    if (i) v = v * ld_fp(pn + i - 1);
    st_fp(z + i, lo + i == 0 ? fr_t::one() : v);
}
void z_finish_chunk(pk_ctx* ctx, const fr_t* pn, const fr_t* sd, const fr_t& factor, fr_t* z, size_t lo, size_t len) {
    z_finish_chunk_kernel<<<grid1d(len, 256), 256, 0, ctx->stream>>>(pn, sd, factor, z, lo, len);
    ctx->prof.kernel_launches++;
}
__global__ void z_finish_kernel(const fr_t* pn, const fr_t* sd, fr_t tinv, fr_t* z_br, int log_n) {
    size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const size_t n = size_t(1) << log_n;
    if (j >= n) return;
    fr_t v = j ? ld_fp(pn + j - 1) * ld_fp(sd + j) * tinv : fr_t::one();
    st_fp(z_br + brev_n((uint32_t)j, log_n), v);
}
void z_finish(pk_ctx* ctx, const fr_t* pn, const fr_t* sd, const fr_t& tinv, fr_t* z_br, int log_n) {
    size_t n = size_t(1) << log_n;
    z_finish_kernel<<<grid1d(n, 256), 256, 0, ctx->stream>>>(pn, sd, tinv, z_br, log_n);
    ctx->prof.kernel_launches++;
}

// ---------------------------------------------------------------- quotient numerator / Z_H on the coset, slot layout
struct QuotKernelArgs {
    QuotientArgs a;
    fr_t alpha2, alpha3, alpha4, alpha5;
    fr_t bg[4];      // beta * g_s
    fr_t zhinv[4];   // 1 / (g_s^N - 1)
    const fr_t* tw;
    int tw_shift;    // tw_log - log_n
};
__global__ void __launch_bounds__(256) quotient_kernel(QuotKernelArgs q) {
    const int log_n = q.a.log_n;
    const size_t n = size_t(1) << log_n;
    const size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;  // position inside the arrays
    if (idx >= q.a.range_len) return;
    const size_t gidx = q.a.range_lo + idx;                            // position in the 4n slot layout
    const int s = (int)(gidx >> log_n);
    const uint32_t p = (uint32_t)(gidx & (n - 1));
    const uint32_t j = brev_n(p, log_n);
    // neighbour w*X: same slot, natural index j + 1 (gathered when the whole slot is resident)
    const size_t idn = q.a.z_next ? idx : ((size_t)s << log_n) + brev_n((j + 1) & (uint32_t)(n - 1), log_n) - q.a.range_lo;

    fr_t a = ld_fp(q.a.w[0] + idx), b = ld_fp(q.a.w[1] + idx), c = ld_fp(q.a.w[2] + idx), d = ld_fp(q.a.w[3] + idx);
    const fr_t dn = q.a.w3_next ? ld_fp(q.a.w3_next + idx) : ld_fp(q.a.w[3] + idn);
    fr_t gate = ld_fp(q.a.sel[0] + idx) * a + ld_fp(q.a.sel[1] + idx) * b + ld_fp(q.a.sel[2] + idx) * c + ld_fp(q.a.sel[3] + idx) * d +
                ld_fp(q.a.sel[4] + idx) * (a * b) + ld_fp(q.a.sel[5] + idx) + ld_fp(q.a.sel[6] + idx) * dn;
    if (q.a.num_direct_inputs < 0) {
        gate = gate + ld_fp(q.a.pi + idx);
    } else {
        for (int i = 0; i < q.a.num_direct_inputs; ++i) {
            const uint32_t pj = brev_n((j - (uint32_t)i) & (uint32_t)(n - 1), log_n);
            gate = gate + q.a.inputs[i] * ld_fp(q.a.l0 + ((size_t)s << log_n) + pj - q.a.range_lo);
        }
    }
    fr_t zv = ld_fp(q.a.z + idx), zn = q.a.z_next ? ld_fp(q.a.z_next + idx) : ld_fp(q.a.z + idn);
    fr_t bx = q.bg[s] * omega_pow(q.tw, q.tw_shift, log_n, j);
    fr_t ag = a + q.a.gamma, bgm = b + q.a.gamma, cg = c + q.a.gamma, dg = d + q.a.gamma;
    fr_t num = zv * (ag + bx) * (bgm + times_k(bx, 1)) * (cg + times_k(bx, 2)) * (dg + times_k(bx, 3));
    fr_t den = zn * (ag + q.a.beta * ld_fp(q.a.sig[0] + idx)) * (bgm + q.a.beta * ld_fp(q.a.sig[1] + idx)) *
               (cg + q.a.beta * ld_fp(q.a.sig[2] + idx)) * (dg + q.a.beta * ld_fp(q.a.sig[3] + idx));
    fr_t tot;
    if (q.a.gsel[0] == nullptr) {
        tot = gate + q.a.alpha * (num - den) + q.alpha2 * ld_fp(q.a.l0 + idx) * (zv - fr_t::one());
    } else {
        // `gate` = main-gate terms + PI; the public-input part stays outside the selector product
        const fr_t pi_part = q.a.num_direct_inputs < 0 ? ld_fp(q.a.pi + idx) : fr_t::zero();
        const fr_t main = gate - pi_part;
        const fr_t resc = q.a.alpha * (a * a - b) + q.alpha2 * (b * b - c) + q.alpha3 * (c * a - d);
        tot = ld_fp(q.a.gsel[0] + idx) * main + pi_part + ld_fp(q.a.gsel[1] + idx) * resc + q.alpha4 * (num - den) +
              q.alpha5 * ld_fp(q.a.l0 + idx) * (zv - fr_t::one());
    }
    st_fp(q.a.out + idx, tot * q.zhinv[s]);
}
void quotient_slots(pk_ctx* ctx, const QuotientArgs& a) {
    const int log_n = a.log_n;
    ensure_twiddles(ctx, log_n + 2);
    DomainCache* dc = ctx->domains;
    QuotKernelArgs q;
    q.a = a;
    if (q.a.range_len == 0) {
        q.a.range_lo = 0;
        q.a.range_len = size_t(4) << log_n;
    } else {
        PK_REQUIRE(a.w3_next && a.z_next && a.num_direct_inputs < 0, PK_ERR_INVALID,
                   "a partial quotient range needs the shifted LDEs and the public-input LDE");
    }
    q.alpha2 = a.alpha.sqr();
    q.alpha3 = q.alpha2 * a.alpha;
    q.alpha4 = q.alpha2.sqr();
    q.alpha5 = q.alpha4 * a.alpha;
    if (a.gsel[0]) PK_REQUIRE(a.num_direct_inputs < 0, PK_ERR_INVALID, "the two-gate quotient takes the public-input LDE");
    fr_t g7;
    for (int i = 0; i < 8; ++i) g7.v[i] = FrRoots::gen7(i);
    fr_t w4 = host_root_of_unity(log_n + 2);
    static const int brev2[4] = {0, 2, 1, 3};
    const uint64_t n = uint64_t(1) << log_n;
    for (int s = 0; s < 4; ++s) {
        fr_t gs = g7 * w4.pow_u64(brev2[s]);
        q.bg[s] = a.beta * gs;
        fr_t zh = gs.pow_u64(n) - fr_t::one();
        PK_REQUIRE(!zh.is_zero(), PK_ERR_DIVISION_BY_ZERO, "vanishing polynomial is zero on the coset");
        q.zhinv[s] = zh.inverse();
    }
    q.tw = dc->tw.p;
    q.tw_shift = dc->tw_log - log_n;
    quotient_kernel<<<grid1d(q.a.range_len, 256), 256, 0, ctx->stream>>>(q);
    ctx->prof.kernel_launches++;
    PK_CUDA(cudaGetLastError());
}

}  // namespace pk
