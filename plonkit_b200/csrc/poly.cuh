// Polynomial glue kernels between the NTTs and the MSMs (implemented in poly.cu): SURVEY.md §8 rows a12-a14
// (bellman's batch_inversion / calculate_shifted_grand_product, pointwise Polynomial ops, evaluate_at, divide_single).
#pragma once
#include "common.cuh"

namespace pk {

struct PolyScratch {
    DevBuf<fr_t> scan_agg;     // block aggregates of the scans
    DevBuf<fr_t> dot_partial;  // [16][DOT_BLOCKS]
    DevBuf<fr_t> dot_out;      // [16]
    DevBuf<uint32_t> flag;     // [1]
};
PolyScratch* poly_scratch(pk_ctx* ctx);

void fr_fill(pk_ctx* ctx, fr_t* out, const fr_t& c, size_t n);
void fr_mul_pointwise(pk_ctx* ctx, const fr_t* a, const fr_t* b, fr_t* out, size_t n);
// out[j] = base^j
void poly_powers(pk_ctx* ctx, fr_t* out, const fr_t& base, size_t n);
void poly_powers_from(pk_ctx* ctx, fr_t* out, const fr_t& base, size_t first, size_t n);  // out[i] = base^(first + i)
// out[i] = sum_k coef[k] * in[k][i]  (nterms <= 16); out may alias an input
void poly_lincomb(pk_ctx* ctx, fr_t* out, int nterms, const fr_t* const* in, const fr_t* coef, size_t n);
// results[k] = sum_i polys[k][i] * pows[k][i]   (npoly <= 16), results on the host (synchronises the stream)
void poly_dot_batch(pk_ctx* ctx, int npoly, const fr_t* const* polys, const fr_t* const* pows, size_t n, fr_t* results_host);
// same sums, left on the device (out_dev[k]), nothing synchronised: partial sums of a sharded evaluation
void poly_dot_batch_dev(pk_ctx* ctx, int npoly, const fr_t* const* polys, const fr_t* const* pows, size_t n, fr_t* out_dev);
// inclusive scan under * (mul = true) or + ; reverse = suffix scan.  in/out may alias.
void poly_scan(pk_ctx* ctx, bool mul, bool reverse, const fr_t* in, fr_t* out, size_t n);
// q(X) = (p(X) - p(z)) / (X - z) given zpow[j] = z^j and zinvpow[j] = z^-j ; tmp is an n-element scratch; q != p
void poly_divide_linear(pk_ctx* ctx, const fr_t* p, const fr_t* zpow, const fr_t* zinvpow, fr_t* q, fr_t* tmp, size_t n);

// the same division on one chunk of the coefficients (sharded prover; see poly.cu)
void poly_divide_linear_chunk_scan(pk_ctx* ctx, const fr_t* p, const fr_t* zpow, fr_t* suffix, size_t len);
void poly_divide_linear_chunk_finish(pk_ctx* ctx, const fr_t* suffix, const fr_t* zi, const fr_t& carry, fr_t* q, size_t len);
// out[i] = in[i]^-1, zeros stay zero (bellman batch_inversion); tmp_a, tmp_b: n-element scratch, out may alias in
void poly_batch_inversion(pk_ctx* ctx, const fr_t* in, fr_t* out, fr_t* tmp_a, fr_t* tmp_b, size_t n);

// ---- prover-specific kernels
// vals_nat[c][row] = vars[idx[c][row]] ; vals_br[c][brev(row)] = same
void wire_gather(pk_ctx* ctx, const fr_t* vars, const uint32_t* idx, fr_t* vals_nat, fr_t* vals_br, int log_n);
// pi_br[brev(i)] = wire a value of row i for i < num_inputs (pi_br must be zeroed)
void pi_scatter(pk_ctx* ctx, const fr_t* vals_nat_a, fr_t* pi_br, uint32_t num_inputs, int log_n);
// gate identity on rows 0..n-2 (values, natural order); returns true if every row vanishes (synchronises)
bool gate_check(pk_ctx* ctx, const fr_t* vals_nat, const fr_t* sel_vals, uint32_t num_inputs, int log_n);
// sigma_vals[c][row] = k_{c'} * w^{row'} for target = sigma_target[c*n + row] = c'*n + row'
void sigma_values(pk_ctx* ctx, const uint32_t* sigma_target, fr_t* sigma_vals, int log_n);
void perm_num_den(pk_ctx* ctx, const fr_t* vals_nat, const fr_t* sigma_vals, const fr_t& beta, const fr_t& gamma, fr_t* num,
                  fr_t* den, int log_n);
void perm_num_den_range(pk_ctx* ctx, const fr_t* vals_nat, const fr_t* sigma_vals, const fr_t& beta, const fr_t& gamma, fr_t* num,
                        fr_t* den, int log_n, size_t lo, size_t len);
void z_finish_chunk(pk_ctx* ctx, const fr_t* pn, const fr_t* sd, const fr_t& factor, fr_t* z, size_t lo, size_t len);
// z_br[brev(0)] = 1 ; z_br[brev(j)] = pn[j-1] * sd[j] * tinv
void z_finish(pk_ctx* ctx, const fr_t* pn, const fr_t* sd, const fr_t& tinv, fr_t* z_br, int log_n);

struct QuotientArgs {
    const fr_t* w[4];
    const fr_t* z;
    const fr_t* sel[7];
    const fr_t* sig[4];
    const fr_t* pi;        // LDE of the public-input polynomial; only read when num_direct_inputs < 0
    // PI(X) = sum_i in_i L_i(X) with L_i(X) = L_0(X w^-i): for a handful of inputs it is evaluated straight from
    // the resident L_0 table (index shift inside the slot) and needs no NTT at all
    int num_direct_inputs; // >= 0: use inputs[]; < 0: use the pi array
    fr_t inputs[8];
    const fr_t* l0;
    fr_t* out;
    fr_t beta, gamma, alpha;
    int log_n;
    // Sharded prover: the arrays above hold only the contiguous range [range_lo, range_lo + range_len) of the 4n slot
    // layout (element 0 of every array = position range_lo), and the values at w*X come from their own arrays (the LDEs
    // of d(wX), Z(wX)) because the neighbour position may live on another GPU.  range_len = 0 means the whole domain
    // with the in-slot neighbour gather (single-GPU prover).
    size_t range_lo = 0, range_len = 0;
    const fr_t* w3_next = nullptr;
    const fr_t* z_next = nullptr;
    // Two gate types (the recursive prover's shape): gsel[0] = s_main, gsel[1] = s_resc on the coset.  The numerator is
    // s_main * (main gate) + PI + s_resc * (alpha (a^2 - b) + alpha^2 (b^2 - c) + alpha^3 (c a - d)) + alpha^4 (copy
    // permutation) + alpha^5 L_0 (Z - 1).  gsel[0] == nullptr: the single-gate protocol (alpha, alpha^2).
    const fr_t* gsel[2] = {nullptr, nullptr};
};
// main-gate identity on rows of type 0, the Rescue x^5 relations on rows of type 1 (gate_type on the device)
bool gate_check_gated(pk_ctx* ctx, const fr_t* vals_nat, const fr_t* sel_vals, const uint8_t* gate_type, uint32_t num_inputs, int log_n);
void quotient_slots(pk_ctx* ctx, const QuotientArgs& a);

}  // namespace pk
