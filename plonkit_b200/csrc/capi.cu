// C ABI of the library (include/plonkit_b200.h): context management, marshalling between caller-owned host buffers
// and device memory, error translation.  No exception crosses the boundary.
#include <mutex>

#include "comm.cuh"
#include "msm.cuh"
#include "ntt.cuh"
#include "poly.cuh"

using namespace pk;

struct pk_comm_group { pk::LocalGroup g; explicit pk_comm_group(int w) : g(w) {} };

namespace pk {
void setup_create(pk_ctx* ctx, const pk_assembly* as, pk_setup** out);
void setup_commitments(pk_ctx* ctx, pk_setup* s, uint64_t out_xy[11][8]);
void setup_create_gated(pk_ctx* ctx, const pk_assembly_gated* as, pk_setup** out);
void setup_commitments_gated(pk_ctx* ctx, pk_setup* s, uint64_t out_xy[13][8]);
void witness_upload(pk_ctx* ctx, pk_setup* s, const uint64_t* var_values, uint64_t nvars);
void prove(pk_ctx* ctx, pk_setup* s, const uint64_t* var_values, uint64_t nvars, pk_proof* proof, uint64_t* inputs_out);
void setup_free(pk_setup* s);
void setup_use_lagrange(pk_ctx* ctx, pk_setup* s, bool on);
void ec_intt(pk_ctx* ctx, uint32_t log_n, uint64_t* out_xy);
void ec_dev_from_affine(pk_ctx* ctx, const g1_affine_t* in_canonical, g1_xyzz_t* out, size_t n);
void ec_dev_ntt_rows(pk_ctx* ctx, g1_xyzz_t* data, int log_len, size_t rows, bool inverse);
void ec_dev_twiddle_rows(pk_ctx* ctx, g1_xyzz_t* a, size_t rows, size_t cols, int log_total, size_t row0, int mode);
void ec_dev_to_affine(pk_ctx* ctx, const g1_xyzz_t* in, g1_affine_t* out_canonical, size_t n, int log_scale);
void srs_gen(pk_ctx* ctx, uint64_t n, uint64_t tau, uint64_t* out_xy);
void dist_setup_create(pk_ctx* ctx, const pk_assembly* as, pk_dist_setup** out);
void dist_setup_commitments(pk_ctx* ctx, pk_dist_setup* s, uint64_t out_xy[11][8]);
void dist_witness_upload(pk_ctx* ctx, pk_dist_setup* s, const uint64_t* var_values, uint64_t nvars);
void dist_prove(pk_ctx* ctx, pk_dist_setup* s, const uint64_t* var_values, uint64_t nvars, pk_proof* proof, uint64_t* inputs_out);
void dist_setup_free(pk_dist_setup* s);

struct CtxExtras { PolyScratch poly; };
static std::map<pk_ctx*, CtxExtras*> g_extras;
static std::mutex g_mu;
PolyScratch* poly_scratch(pk_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_extras.find(ctx);
    if (it == g_extras.end()) it = g_extras.emplace(ctx, new CtxExtras()).first;
    return &it->second->poly;
}

void profile_resolve(pk_ctx* ctx) {
    for (auto& pe : ctx->prof.pending) {
        cudaEventSynchronize(pe.b);
        float ms = 0;
        cudaEventElapsedTime(&ms, pe.a, pe.b);
        if (pe.kind == 0) { ctx->prof.msm_accum_ms += ms; ctx->prof.msm_accum_points += pe.units; }
        else { ctx->prof.ntt_ms += ms; ctx->prof.ntt_elements += pe.units; }
        cudaEventDestroy(pe.a);
        cudaEventDestroy(pe.b);
    }
    ctx->prof.pending.clear();
}
}  // namespace pk

#define PK_API_BEGIN(ctx)  \
    if (!(ctx)) return PK_ERR_INVALID; \
    try {                  \
        cudaSetDevice((ctx)->device);
#define PK_API_END(ctx)                                                   \
        return PK_OK;                                                     \
    } catch (const PkError& e) {                                          \
        (ctx)->last_error = e.what();                                     \
        cudaGetLastError();                                               \
        return e.code;                                                    \
    } catch (const std::exception& e) {                                   \
        (ctx)->last_error = e.what();                                     \
        return PK_ERR_INVALID;                                            \
    } catch (...) {                                                       \
        (ctx)->last_error = "unknown error";                              \
        return PK_ERR_INVALID;                                            \
    }

extern "C" {

int pk_create(int device, pk_ctx** out) {
    if (!out) return PK_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) return PK_ERR_CUDA;  // no CPU fallback exists
    if (device < 0 || device >= count) return PK_ERR_INVALID;
    pk_ctx* ctx = new pk_ctx();
    ctx->device = device;
    try {
        PK_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        PK_CUDA(cudaGetDeviceProperties(&prop, device));
        ctx->sm_count = prop.multiProcessorCount;
        PK_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        PK_CUDA(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
        PK_CUDA(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        PK_CUDA(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
        ctx->pinned_bytes = 1 << 16;
        PK_CUDA(cudaMallocHost(&ctx->pinned, ctx->pinned_bytes));
    } catch (const PkError&) {
        delete ctx;
        return PK_ERR_CUDA;
    }
    *out = ctx;
    return PK_OK;
}

void pk_destroy(pk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    profile_resolve(ctx);
    delete ctx->srs_lagrange;
    delete ctx->srs;
    delete ctx->domains;
    delete ctx->comm;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_extras.find(ctx);
        if (it != g_extras.end()) { delete it->second; g_extras.erase(it); }
    }
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* pk_last_error(const pk_ctx* ctx) { return ctx ? ctx->last_error.c_str() : "null context"; }

void pk_constants(uint64_t out[24]) {
    memset(out, 0, 24 * sizeof(uint64_t));
    uint32_t* o = reinterpret_cast<uint32_t*>(out);
    for (int i = 0; i < 8; ++i) {
        o[i] = FrParams::one(i); o[8 + i] = FrParams::r2(i);
        o[24 + i] = FqParams::one(i); o[32 + i] = FqParams::r2(i);
    }
    out[8] = FrParams::INV;
    out[20] = FqParams::INV;
}

int pk_srs_load_g1(pk_ctx* ctx, const uint64_t* bases_xy, uint64_t n, int window_bits) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(bases_xy != nullptr, PK_ERR_INVALID, "null bases");
    srs_load(ctx, bases_xy, n, window_bits);
    PK_API_END(ctx)
}

int pk_srs_load_g1_lagrange(pk_ctx* ctx, const uint64_t* bases_xy, uint64_t n, int window_bits) {
    PK_API_BEGIN(ctx)
    if (bases_xy == nullptr && n == 0) {  // unload
        PK_CUDA(cudaStreamSynchronize(ctx->stream));
        delete ctx->srs_lagrange;
        ctx->srs_lagrange = nullptr;
        return PK_OK;
    }
    PK_REQUIRE(bases_xy != nullptr, PK_ERR_INVALID, "null bases");
    PK_REQUIRE((n & (n - 1)) == 0, PK_ERR_INVALID, "a Lagrange-form key belongs to a power-of-two domain");
    srs_load(ctx, bases_xy, n, window_bits, true);
    PK_API_END(ctx)
}

int pk_srs_gen(pk_ctx* ctx, uint64_t n, uint64_t tau, uint64_t* out_xy) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(out_xy != nullptr && n >= 1, PK_ERR_INVALID, "bad argument");
    srs_gen(ctx, n, tau, out_xy);
    PK_API_END(ctx)
}

int pk_ntt(pk_ctx* ctx, uint64_t* fr, uint32_t log_n, int inverse, int coset, int fmt) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(fr != nullptr, PK_ERR_INVALID, "null data");
    PK_REQUIRE(log_n <= 28, PK_ERR_DEGREE_TOO_LARGE, "domain larger than 2^28");
    const size_t n = size_t(1) << log_n;
    cudaStream_t st = ctx->stream;
    DevBuf<fr_t> a(n), b(n), pw;
    PK_CUDA(cudaMemcpyAsync(a.p, fr, n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    if (fmt == PK_FMT_CANONICAL) fr_to_mont(ctx, a.p, n);
    fr_t g7, g7inv;
    for (int i = 0; i < 8; ++i) { g7.v[i] = FrRoots::gen7(i); g7inv.v[i] = FrRoots::gen7_inv(i); }
    if (!inverse) {
        if (coset) { pw.alloc(n); poly_powers(ctx, pw.p, g7, n); }
        ntt_forward_bitrev(ctx, a.p, b.p, (int)log_n, coset ? pw.p : nullptr);
        bitrev_permute(ctx, b.p, a.p, (int)log_n);
    } else {
        bitrev_permute(ctx, a.p, b.p, (int)log_n);
        ntt_inverse_from_bitrev(ctx, b.p, a.p, (int)log_n);
        if (coset) { pw.alloc(n); poly_powers(ctx, pw.p, g7inv, n); fr_mul_pointwise(ctx, a.p, pw.p, a.p, n); }
    }
    if (fmt == PK_FMT_CANONICAL) fr_from_mont(ctx, a.p, a.p, n);
    PK_CUDA(cudaMemcpyAsync(fr, a.p, n * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
    PK_CUDA(cudaStreamSynchronize(st));
    PK_CUDA(cudaGetLastError());
    PK_API_END(ctx)
}

// ---------------------------------------------------------------- polynomial primitives
static void upload_fr(pk_ctx* ctx, DevBuf<fr_t>& d, const uint64_t* host, size_t n) {
    d.alloc(n);
    PK_CUDA(cudaMemcpyAsync(d.p, host, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    fr_to_mont(ctx, d.p, n);
}
static void download_fr(pk_ctx* ctx, fr_t* dev, uint64_t* host, size_t n) {
    fr_from_mont(ctx, dev, dev, n);
    PK_CUDA(cudaMemcpyAsync(host, dev, n * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaGetLastError());
}

int pk_poly_evaluate_at(pk_ctx* ctx, const uint64_t* coeffs, uint64_t n, const uint64_t z[4], uint64_t out[4]) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(z != nullptr && out != nullptr && (coeffs != nullptr || n == 0), PK_ERR_INVALID, "null argument");
    if (n == 0) { memset(out, 0, 32); return PK_OK; }
    DevBuf<fr_t> c, pw(n);
    upload_fr(ctx, c, coeffs, n);
    poly_powers(ctx, pw.p, host_load_canonical<fr_t>(z), n);
    const fr_t* polys[1] = {c.p};
    const fr_t* pows[1] = {pw.p};
    fr_t r;
    poly_dot_batch(ctx, 1, polys, pows, n, &r);
    host_store_canonical(r, out);
    PK_API_END(ctx)
}

int pk_poly_divide_by_linear(pk_ctx* ctx, const uint64_t* coeffs, uint64_t n, const uint64_t z[4], uint64_t* quotient) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(z != nullptr && (n == 0 || (coeffs != nullptr && quotient != nullptr)), PK_ERR_INVALID, "null argument");
    if (n == 0) return PK_OK;
    const fr_t zf = host_load_canonical<fr_t>(z);
    if (zf.is_zero()) {  // division by X: a shift
        memmove(quotient, coeffs + 4, (n - 1) * 32);
        memset(quotient + 4 * (n - 1), 0, 32);
        return PK_OK;
    }
    DevBuf<fr_t> c, zp(n), zi(n), q(n), tmp(n);
    upload_fr(ctx, c, coeffs, n);
    poly_powers(ctx, zp.p, zf, n);
    poly_powers(ctx, zi.p, zf.inverse(), n);
    poly_divide_linear(ctx, c.p, zp.p, zi.p, q.p, tmp.p, n);
    download_fr(ctx, q.p, quotient, n);
    PK_API_END(ctx)
}

int pk_poly_shifted_grand_product(pk_ctx* ctx, const uint64_t* values, uint64_t n, uint64_t* out) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(n == 0 || (values != nullptr && out != nullptr), PK_ERR_INVALID, "null argument");
    if (n == 0) return PK_OK;
    DevBuf<fr_t> v, r(n);
    upload_fr(ctx, v, values, n);
    poly_scan(ctx, true, false, v.p, v.p, n);
    fr_fill(ctx, r.p, fr_t::one(), 1);
    if (n > 1) PK_CUDA(cudaMemcpyAsync(r.p + 1, v.p, (n - 1) * sizeof(fr_t), cudaMemcpyDeviceToDevice, ctx->stream));
    download_fr(ctx, r.p, out, n);
    PK_API_END(ctx)
}

int pk_poly_batch_inversion(pk_ctx* ctx, uint64_t* values, uint64_t n) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(n == 0 || values != nullptr, PK_ERR_INVALID, "null argument");
    if (n == 0) return PK_OK;
    DevBuf<fr_t> v, a(n), b(n);
    upload_fr(ctx, v, values, n);
    poly_batch_inversion(ctx, v.p, v.p, a.p, b.p, n);
    download_fr(ctx, v.p, values, n);
    PK_API_END(ctx)
}

// Polynomial<Fr, _>::{add_assign_scaled, mul_assign, scale, add_constant, distribute_powers}
int pk_poly_pointwise(pk_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, const uint64_t scalar[4], uint64_t n, uint64_t* out) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(op >= 0 && op <= 4, PK_ERR_INVALID, "unknown pointwise operation");
    PK_REQUIRE(n == 0 || (a != nullptr && out != nullptr), PK_ERR_INVALID, "null argument");
    PK_REQUIRE(op > 1 || n == 0 || b != nullptr, PK_ERR_ASSIGNMENT_MISSING, "second operand missing");
    PK_REQUIRE(op == 1 || scalar != nullptr, PK_ERR_INVALID, "scalar missing");
    if (n == 0) return PK_OK;
    DevBuf<fr_t> da, db, dc(n);
    upload_fr(ctx, da, a, n);
    const fr_t s = op == 1 ? fr_t::one() : host_load_canonical<fr_t>(scalar);
    if (op == 0) {            // a + s * b
        upload_fr(ctx, db, b, n);
        const fr_t* in[2] = {da.p, db.p};
        fr_t coef[2] = {fr_t::one(), s};
        poly_lincomb(ctx, dc.p, 2, in, coef, n);
    } else if (op == 1) {     // a * b
        upload_fr(ctx, db, b, n);
        fr_mul_pointwise(ctx, da.p, db.p, dc.p, n);
    } else if (op == 2) {     // s * a
        const fr_t* in[1] = {da.p};
        fr_t coef[1] = {s};
        poly_lincomb(ctx, dc.p, 1, in, coef, n);
    } else if (op == 3) {     // a + s
        db.alloc(n);
        fr_fill(ctx, db.p, s, n);
        const fr_t* in[2] = {da.p, db.p};
        fr_t coef[2] = {fr_t::one(), fr_t::one()};
        poly_lincomb(ctx, dc.p, 2, in, coef, n);
    } else {                  // a_i * s^i
        db.alloc(n);
        poly_powers(ctx, db.p, s, n);
        fr_mul_pointwise(ctx, da.p, db.p, dc.p, n);
    }
    download_fr(ctx, dc.p, out, n);
    PK_API_END(ctx)
}

int pk_lde4(pk_ctx* ctx, const uint64_t* coeffs, uint32_t log_n, uint64_t* out_4n, int bitreversed, int fmt) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(coeffs != nullptr && out_4n != nullptr, PK_ERR_INVALID, "null data");
    PK_REQUIRE(log_n + 2 <= 28, PK_ERR_DEGREE_TOO_LARGE, "4n domain larger than 2^28");
    const size_t n = size_t(1) << log_n;
    cudaStream_t st = ctx->stream;
    DevBuf<fr_t> a(n), b(4 * n), c;
    PK_CUDA(cudaMemcpyAsync(a.p, coeffs, n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
    if (fmt == PK_FMT_CANONICAL) fr_to_mont(ctx, a.p, n);
    lde4_slots(ctx, a.p, b.p, (int)log_n);
    fr_t* res = b.p;
    if (!bitreversed) { c.alloc(4 * n); bitrev_permute(ctx, b.p, c.p, (int)log_n + 2); res = c.p; }
    if (fmt == PK_FMT_CANONICAL) fr_from_mont(ctx, res, res, 4 * n);
    PK_CUDA(cudaMemcpyAsync(out_4n, res, 4 * n * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
    PK_CUDA(cudaStreamSynchronize(st));
    PK_CUDA(cudaGetLastError());
    PK_API_END(ctx)
}

int pk_msm_g1(pk_ctx* ctx, const uint64_t* scalars, uint64_t n, uint64_t base_offset, uint64_t out_xy[8], int* is_infinity, int fmt) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(out_xy != nullptr, PK_ERR_INVALID, "null output");
    PK_REQUIRE(n == 0 || scalars != nullptr, PK_ERR_ASSIGNMENT_MISSING, "null scalars");
    PK_REQUIRE(ctx->srs != nullptr, PK_ERR_DEGREE_TOO_LARGE, "no SRS loaded");
    PK_REQUIRE(base_offset + n <= ctx->srs->n, PK_ERR_DEGREE_TOO_LARGE, "MSM longer than the resident SRS");
    g1_affine_t r = g1_affine_t::infinity();
    if (n) {
        DevBuf<fr_t> s(n);
        PK_CUDA(cudaMemcpyAsync(s.p, scalars, n * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
        if (fmt == PK_FMT_CANONICAL) fr_to_mont(ctx, s.p, n);
        r = msm_run(ctx, s.p, n, base_offset);
    }
    affine_to_abi(r, out_xy);
    if (is_infinity) *is_infinity = r.is_inf() ? 1 : 0;
    PK_API_END(ctx)
}

int pk_g1_sum(const uint64_t* points_xy, uint64_t n, uint64_t out_xy[8]) {
    if (!out_xy || (n && !points_xy)) return PK_ERR_INVALID;
    g1_xyzz_t acc = g1_xyzz_t::infinity();
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t* s = points_xy + 8 * i;
        bool inf = true;
        for (int k = 0; k < 8; ++k) inf = inf && s[k] == 0;
        if (inf) continue;
        g1_affine_t p;
        p.x = host_load_canonical<fq_t>(s);
        p.y = host_load_canonical<fq_t>(s + 4);
        acc = acc.add_mixed(p);
    }
    affine_to_abi(acc.to_affine(), out_xy);
    return PK_OK;
}

// ---- device-pointer primitives (caller-owned device memory, e.g. torch tensors shared with NCCL collectives)
int pk_dev_fr_convert(pk_ctx* ctx, void* dev, uint64_t n, int to_mont) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(dev != nullptr || n == 0, PK_ERR_INVALID, "null device pointer");
    if (to_mont) fr_to_mont(ctx, static_cast<fr_t*>(dev), n);
    else fr_from_mont(ctx, static_cast<fr_t*>(dev), static_cast<fr_t*>(dev), n);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaGetLastError());
    PK_API_END(ctx)
}
int pk_dev_ntt_rows(pk_ctx* ctx, void* dev, uint32_t log_len, uint64_t rows, int inverse) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(dev != nullptr && rows >= 1, PK_ERR_INVALID, "bad argument");
    PK_REQUIRE(log_len <= 28, PK_ERR_DEGREE_TOO_LARGE, "domain larger than 2^28");
    DevBuf<fr_t> tmp((size_t)rows << log_len);
    ntt_rows_natural(ctx, static_cast<fr_t*>(dev), tmp.p, (int)log_len, rows, inverse != 0);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaGetLastError());
    PK_API_END(ctx)
}
int pk_dev_twiddle(pk_ctx* ctx, void* dev, uint64_t rows, uint64_t cols, uint32_t log_total, uint64_t row0, int inverse) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(dev != nullptr && rows >= 1 && cols >= 1, PK_ERR_INVALID, "bad argument");
    PK_REQUIRE(log_total <= 28, PK_ERR_DEGREE_TOO_LARGE, "domain larger than 2^28");
    twiddle_rows(ctx, static_cast<fr_t*>(dev), rows, cols, (int)log_total, row0, inverse != 0);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaGetLastError());
    PK_API_END(ctx)
}

int pk_dev_ec_from_affine(pk_ctx* ctx, const void* dev_affine, void* dev_xyzz, uint64_t n) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE((dev_affine != nullptr && dev_xyzz != nullptr) || n == 0, PK_ERR_INVALID, "null device pointer");
    if (n) ec_dev_from_affine(ctx, static_cast<const g1_affine_t*>(dev_affine), static_cast<g1_xyzz_t*>(dev_xyzz), n);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_API_END(ctx)
}
int pk_dev_ec_ntt_rows(pk_ctx* ctx, void* dev_xyzz, uint32_t log_len, uint64_t rows, int inverse) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(dev_xyzz != nullptr && rows >= 1, PK_ERR_INVALID, "bad argument");
    PK_REQUIRE(log_len <= 26, PK_ERR_DEGREE_TOO_LARGE, "domain larger than 2^26");
    ec_dev_ntt_rows(ctx, static_cast<g1_xyzz_t*>(dev_xyzz), (int)log_len, rows, inverse != 0);
    PK_API_END(ctx)
}
int pk_dev_ec_twiddle(pk_ctx* ctx, void* dev_xyzz, uint64_t rows, uint64_t cols, uint32_t log_total, uint64_t row0, int inverse) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(dev_xyzz != nullptr && rows >= 1 && cols >= 1, PK_ERR_INVALID, "bad argument");
    PK_REQUIRE(log_total <= 28, PK_ERR_DEGREE_TOO_LARGE, "domain larger than 2^28");
    ec_dev_twiddle_rows(ctx, static_cast<g1_xyzz_t*>(dev_xyzz), rows, cols, (int)log_total, row0, inverse);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_API_END(ctx)
}
int pk_dev_ec_to_affine(pk_ctx* ctx, const void* dev_xyzz, void* dev_affine, uint64_t n, uint32_t log_scale) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE((dev_affine != nullptr && dev_xyzz != nullptr) || n == 0, PK_ERR_INVALID, "null device pointer");
    if (n) ec_dev_to_affine(ctx, static_cast<const g1_xyzz_t*>(dev_xyzz), static_cast<g1_affine_t*>(dev_affine), n, (int)log_scale);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_API_END(ctx)
}

int pk_ec_intt_g1(pk_ctx* ctx, uint32_t log_n, uint64_t* out_xy) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(out_xy != nullptr, PK_ERR_INVALID, "null output");
    ec_intt(ctx, log_n, out_xy);
    PK_API_END(ctx)
}

int pk_setup_create(pk_ctx* ctx, const pk_assembly* assembly, pk_setup** out) {
    PK_API_BEGIN(ctx)
    setup_create(ctx, assembly, out);
    PK_API_END(ctx)
}
int pk_setup_create_gated(pk_ctx* ctx, const pk_assembly_gated* assembly, pk_setup** out) {
    PK_API_BEGIN(ctx)
    setup_create_gated(ctx, assembly, out);
    PK_API_END(ctx)
}
int pk_setup_commitments_gated(pk_ctx* ctx, pk_setup* setup, uint64_t out_xy[13][8]) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(setup != nullptr && out_xy != nullptr, PK_ERR_INVALID, "null argument");
    setup_commitments_gated(ctx, setup, out_xy);
    PK_API_END(ctx)
}
void pk_setup_destroy(pk_setup* setup) { setup_free(setup); }
int pk_setup_use_lagrange(pk_ctx* ctx, pk_setup* setup, int on) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(setup != nullptr, PK_ERR_INVALID, "null setup");
    setup_use_lagrange(ctx, setup, on != 0);
    PK_API_END(ctx)
}

int pk_setup_commitments(pk_ctx* ctx, pk_setup* setup, uint64_t out_xy[11][8]) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(setup != nullptr && out_xy != nullptr, PK_ERR_INVALID, "null argument");
    setup_commitments(ctx, setup, out_xy);
    PK_API_END(ctx)
}
int pk_witness_upload(pk_ctx* ctx, pk_setup* setup, const uint64_t* var_values, uint64_t nvars) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(setup != nullptr, PK_ERR_INVALID, "null setup");
    witness_upload(ctx, setup, var_values, nvars);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_API_END(ctx)
}
int pk_prove(pk_ctx* ctx, pk_setup* setup, const uint64_t* var_values, uint64_t nvars, pk_proof* proof, uint64_t* inputs_out) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(setup != nullptr, PK_ERR_INVALID, "null setup");
    prove(ctx, setup, var_values, nvars, proof, inputs_out);
    PK_API_END(ctx)
}

// ---------------------------------------------------------------- sharded prover (one proof over several GPUs)
// A rank that fails tells its peers, so that they fail in their next collective instead of waiting for it forever.
#define PK_DIST_API_END(ctx)                                              \
        return PK_OK;                                                     \
    } catch (const PkError& e) {                                          \
        (ctx)->last_error = e.what();                                     \
        cudaGetLastError();                                               \
        if ((ctx)->comm) (ctx)->comm->abort();                            \
        return e.code;                                                    \
    } catch (const std::exception& e) {                                   \
        (ctx)->last_error = e.what();                                     \
        if ((ctx)->comm) (ctx)->comm->abort();                            \
        return PK_ERR_INVALID;                                            \
    } catch (...) {                                                       \
        (ctx)->last_error = "unknown error";                              \
        if ((ctx)->comm) (ctx)->comm->abort();                            \
        return PK_ERR_INVALID;                                            \
    }

int pk_comm_group_create(int world, pk_comm_group** out) {
    if (!out || world < 1 || world > 8) return PK_ERR_INVALID;
    *out = new pk_comm_group(world);
    return PK_OK;
}
void pk_comm_group_destroy(pk_comm_group* g) { delete g; }
int pk_comm_attach_group(pk_ctx* ctx, pk_comm_group* group, int rank) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(group != nullptr, PK_ERR_INVALID, "null group");
    PK_REQUIRE(ctx->comm == nullptr, PK_ERR_INVALID, "context already belongs to a communicator");
    ctx->comm = make_local_comm(&group->g, rank, ctx->device);
    PK_API_END(ctx)
}
int pk_comm_nccl_unique_id(uint8_t out[128]) {
    if (!out) return PK_ERR_INVALID;
    try {
        nccl_unique_id(out);
    } catch (...) {
        return PK_ERR_CUDA;
    }
    return PK_OK;
}
int pk_comm_attach_nccl(pk_ctx* ctx, const uint8_t unique_id[128], int rank, int world) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(unique_id != nullptr, PK_ERR_INVALID, "null id");
    PK_REQUIRE(ctx->comm == nullptr, PK_ERR_INVALID, "context already belongs to a communicator");
    ctx->comm = make_nccl_comm(unique_id, rank, world);
    PK_API_END(ctx)
}
int pk_dist_setup_create(pk_ctx* ctx, const pk_assembly* assembly, pk_dist_setup** out) {
    PK_API_BEGIN(ctx)
    if (ctx->comm) ctx->comm->begin();
    dist_setup_create(ctx, assembly, out);
    PK_DIST_API_END(ctx)
}
void pk_dist_setup_destroy(pk_dist_setup* setup) { dist_setup_free(setup); }
int pk_dist_setup_commitments(pk_ctx* ctx, pk_dist_setup* setup, uint64_t out_xy[11][8]) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(setup != nullptr && out_xy != nullptr, PK_ERR_INVALID, "null argument");
    if (ctx->comm) ctx->comm->begin();
    dist_setup_commitments(ctx, setup, out_xy);
    PK_DIST_API_END(ctx)
}
int pk_dist_witness_upload(pk_ctx* ctx, pk_dist_setup* setup, const uint64_t* var_values, uint64_t nvars) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(setup != nullptr, PK_ERR_INVALID, "null setup");
    dist_witness_upload(ctx, setup, var_values, nvars);
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_API_END(ctx)
}
int pk_dist_prove(pk_ctx* ctx, pk_dist_setup* setup, const uint64_t* var_values, uint64_t nvars, pk_proof* proof, uint64_t* inputs_out) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(setup != nullptr, PK_ERR_INVALID, "null setup");
    if (ctx->comm) ctx->comm->begin();
    dist_prove(ctx, setup, var_values, nvars, proof, inputs_out);
    PK_DIST_API_END(ctx)
}

void pk_profile_enable(pk_ctx* ctx, int on) { if (ctx) ctx->prof.enabled = on != 0; }
void pk_profile_reset(pk_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    profile_resolve(ctx);
    bool en = ctx->prof.enabled;
    ctx->prof = Profile();
    ctx->prof.enabled = en;
}
void pk_profile_get(const pk_ctx* cctx, pk_profile* out) {
    if (!cctx || !out) return;
    pk_ctx* ctx = const_cast<pk_ctx*>(cctx);
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    profile_resolve(ctx);
    memset(out, 0, sizeof(*out));
    out->kernel_launches = ctx->prof.kernel_launches;
    out->msm_accum_launches = ctx->prof.msm_accum_launches;
    out->msm_accum_ms = ctx->prof.msm_accum_ms;
    out->msm_accum_points = ctx->prof.msm_accum_points;
    out->ntt_launches = ctx->prof.ntt_launches;
    out->ntt_ms = ctx->prof.ntt_ms;
    out->ntt_elements = ctx->prof.ntt_elements;
    for (int i = 0; i < 8; ++i) out->phase_ms[i] = ctx->prof.phase_ms[i];
    for (int i = 0; i < 3; ++i) { out->comm_ms[i] = ctx->prof.comm_ms[i]; out->comm_bytes[i] = ctx->prof.comm_bytes[i]; }
}

static std::map<pk_ctx*, std::pair<cudaEvent_t, cudaEvent_t>> g_timers;
int pk_timer_begin(pk_ctx* ctx) {
    PK_API_BEGIN(ctx)
    std::pair<cudaEvent_t, cudaEvent_t> t;
    {
        std::lock_guard<std::mutex> lk(g_mu);  // contexts are driven from different host threads
        auto& slot = g_timers[ctx];
        if (!slot.first) { PK_CUDA(cudaEventCreate(&slot.first)); PK_CUDA(cudaEventCreate(&slot.second)); }
        t = slot;
    }
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    PK_CUDA(cudaEventRecord(t.first, ctx->stream));
    PK_API_END(ctx)
}
int pk_timer_end(pk_ctx* ctx, double* ms) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(ms != nullptr, PK_ERR_INVALID, "null output");
    std::pair<cudaEvent_t, cudaEvent_t> t(nullptr, nullptr);
    {
        std::lock_guard<std::mutex> lk(g_mu);
        auto it = g_timers.find(ctx);
        if (it != g_timers.end()) t = it->second;
    }
    PK_REQUIRE(t.first, PK_ERR_INVALID, "pk_timer_begin was not called");
    PK_CUDA(cudaEventRecord(t.second, ctx->stream));
    PK_CUDA(cudaEventSynchronize(t.second));
    float f = 0;
    PK_CUDA(cudaEventElapsedTime(&f, t.first, t.second));
    *ms = f;
    PK_API_END(ctx)
}

// ---- device-resident micro-benchmarks
int pk_bench_ntt(pk_ctx* ctx, uint32_t log_n, int iters, double* ms_per_iter) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(ms_per_iter && iters > 0 && log_n >= 1 && log_n <= 28, PK_ERR_INVALID, "bad argument");
    const size_t n = size_t(1) << log_n;
    DevBuf<fr_t> a(n), b(n);
    poly_powers(ctx, a.p, fr_t::from_u32(0x706c6f6eu).sqr(), n);
    ntt_forward_bitrev(ctx, a.p, b.p, (int)log_n);  // warm-up (also builds twiddles)
    cudaEvent_t e0, e1;
    PK_CUDA(cudaEventCreate(&e0));
    PK_CUDA(cudaEventCreate(&e1));
    PK_CUDA(cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < iters; ++i) ntt_forward_bitrev(ctx, i & 1 ? b.p : a.p, i & 1 ? a.p : b.p, (int)log_n);
    PK_CUDA(cudaEventRecord(e1, ctx->stream));
    PK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_iter = ms / iters;
    PK_API_END(ctx)
}

int pk_bench_msm(pk_ctx* ctx, uint64_t n, int iters, double* ms_per_iter) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(ms_per_iter && iters > 0 && n >= 1, PK_ERR_INVALID, "bad argument");
    PK_REQUIRE(ctx->srs && ctx->srs->n >= n, PK_ERR_DEGREE_TOO_LARGE, "MSM longer than the resident SRS");
    DevBuf<fr_t> s(n);
    poly_powers(ctx, s.p, fr_t::from_u32(0x6b697431u).sqr().sqr(), n);  // pseudo-random looking scalars
    msm_run(ctx, s.p, n, 0);
    cudaEvent_t e0, e1;
    PK_CUDA(cudaEventCreate(&e0));
    PK_CUDA(cudaEventCreate(&e1));
    PK_CUDA(cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < iters; ++i) msm_run(ctx, s.p, n, 0);
    PK_CUDA(cudaEventRecord(e1, ctx->stream));
    PK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_iter = ms / iters;
    PK_API_END(ctx)
}

}  // extern "C"

// scalars for pk_bench_msm_pattern: 1 = the witness-like mix of SURVEY 8(d) (40 % zero, 10 % one, 50 % uniform), 2 = all ones
__global__ void msm_pattern_kernel(fr_t* s, size_t n, int pattern) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t h = (uint32_t)i * 2654435761u;
    h ^= h >> 15; h *= 2246822519u; h ^= h >> 13;
    const uint32_t r = h % 10;
    if (pattern == 2 || (pattern == 1 && r == 4)) st_fp(s + i, fr_t::one());
    else if (pattern == 1 && r < 4) st_fp(s + i, fr_t::zero());
}
extern "C" int pk_bench_msm_pattern(pk_ctx* ctx, uint64_t n, int pattern, int iters, double* ms_per_iter) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(ms_per_iter && iters > 0 && n >= 1 && pattern >= 0 && pattern <= 2, PK_ERR_INVALID, "bad argument");
    PK_REQUIRE(ctx->srs && ctx->srs->n >= n, PK_ERR_DEGREE_TOO_LARGE, "MSM longer than the resident SRS");
    DevBuf<fr_t> s(n);
    poly_powers(ctx, s.p, fr_t::from_u32(0x6b697431u).sqr().sqr(), n);
    if (pattern) msm_pattern_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(s.p, n, pattern);
    msm_run(ctx, s.p, n, 0);
    cudaEvent_t e0, e1;
    PK_CUDA(cudaEventCreate(&e0));
    PK_CUDA(cudaEventCreate(&e1));
    PK_CUDA(cudaEventRecord(e0, ctx->stream));
    for (int i = 0; i < iters; ++i) msm_run(ctx, s.p, n, 0);
    PK_CUDA(cudaEventRecord(e1, ctx->stream));
    PK_CUDA(cudaEventSynchronize(e1));
    float ms = 0;
    PK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_per_iter = ms / iters;
    PK_API_END(ctx)
}

// field-multiplier throughput microbenchmark: the integer roofline the NTT/MSM kernels live under
template <class F> __global__ void fieldmul_bench_kernel(F* out, int iters) {
    const unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    F a = F::from_u32(t + 3), b = F::from_u32(2 * t + 5), c = F::from_u32(7 * t + 11), d = F::from_u32(t ^ 0x5555);
    for (int i = 0; i < iters; ++i) {  // 4 independent chains
        a = a * b; b = b * c; c = c * d; d = d * a;
    }
    st_fp(out + t, a + b + c + d);
}
extern "C" int pk_bench_fieldmul(pk_ctx* ctx, int which, double* gmuls_per_s) {
    PK_API_BEGIN(ctx)
    PK_REQUIRE(gmuls_per_s != nullptr, PK_ERR_INVALID, "bad argument");
    const int blocks = ctx->sm_count * 8, threads = 256, iters = 2048;
    DevBuf<fr_t> out((size_t)blocks * threads);
    cudaEvent_t e0, e1;
    PK_CUDA(cudaEventCreate(&e0));
    PK_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        PK_CUDA(cudaEventRecord(e0, ctx->stream));
        if (which == 0) fieldmul_bench_kernel<fr_t><<<blocks, threads, 0, ctx->stream>>>(out.p, iters);
        else fieldmul_bench_kernel<fq_t><<<blocks, threads, 0, ctx->stream>>>(reinterpret_cast<fq_t*>(out.p), iters);
        PK_CUDA(cudaEventRecord(e1, ctx->stream));
        PK_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        PK_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    ctx->prof.kernel_launches += 4;
    *gmuls_per_s = (double)blocks * threads * iters * 4.0 / (best * 1e-3) / 1e9;
    PK_API_END(ctx)
}
