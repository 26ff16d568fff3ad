// NTT layer interface (implemented in ntt.cu).
#pragma once
#include "common.cuh"

namespace pk {

// Per-size coset tables for the quotient domain 7*H_4N of a size-N circuit (N = 2^log_n).
// "slot" s in 0..3 holds the coset 7*w4N^brev2(s) * H_N, each in bit-reversed order of j: together the 4N array is the
// 4N-point bit-reversed order of the natural coset 7*H_4N (what bellman's bitreversed LDE produces).
struct CosetTables {
    int log_n = 0;
    DevBuf<fr_t> scale4;     // [4][N]   (7 * w4N^brev2(s))^j            — pre-scale of the slot-s forward NTT
    DevBuf<fr_t> iscale4n;   // [4N]     7^{-k} / (4N)                    — post-scale of the size-4N inverse NTT
    DevBuf<fr_t> l0;         // [4][N]   L_0 on the coset, slot layout
};

struct DomainCache {
    int tw_log = 0;          // twiddle table covers the domain of size 2^tw_log: tw[e] = w^e, e = 0 .. 2^(tw_log-1)
    DevBuf<fr_t> tw;
    std::map<int, CosetTables*> coset;
    ~DomainCache() { for (auto& kv : coset) delete kv.second; }
};

fr_t host_root_of_unity(int log_n);                       // w_N, Montgomery form
void ensure_twiddles(pk_ctx* ctx, int log_n);             // table for domains up to 2^log_n
CosetTables* get_coset_tables(pk_ctx* ctx, int log_n);    // builds on first use (needs twiddles for log_n + 2)

// forward NTT, natural order in -> bit-reversed order out (decimation in frequency).  `batch` independent transforms,
// element strides between consecutive batches; pre (optional) multiplies element i of batch b by pre[b*pre_stride + i].
void ntt_forward_bitrev(pk_ctx* ctx, const fr_t* src, fr_t* dst, int log_n, const fr_t* pre = nullptr, int batch = 1,
                        size_t src_stride = 0, size_t dst_stride = 0, size_t pre_stride = 0);
// inverse NTT, bit-reversed order in -> natural order out (decimation in time), result multiplied by post[i] if given,
// else by 1/n.
void ntt_inverse_from_bitrev(pk_ctx* ctx, const fr_t* src, fr_t* dst, int log_n, const fr_t* post = nullptr);
void bitrev_permute(pk_ctx* ctx, const fr_t* src, fr_t* dst, int log_n, size_t rows = 1);  // out of place, src != dst
// rows independent (i)NTTs of length 2^log_len, natural order in and out, in place (tmp: same size scratch)
void ntt_rows_natural(pk_ctx* ctx, fr_t* data, fr_t* tmp, int log_len, size_t rows, bool inverse);
// a[r][c] *= w_N^{+-(row0 + r) * c}: the twiddle step between the two halves of a four-step NTT of size N = 2^log_total
void twiddle_rows(pk_ctx* ctx, fr_t* a, size_t rows, size_t cols, int log_total, size_t row0, bool inverse);
// coefficients (N, natural) -> evaluations on 7*H_4N in slot layout (4N)
void lde4_slots(pk_ctx* ctx, const fr_t* coeffs, fr_t* out4n, int log_n);
// evaluations on 7*H_4N in slot layout -> 4N coefficients (natural), in place allowed
void icoset4n_from_slots(pk_ctx* ctx, const fr_t* vals4n, fr_t* coeffs4n, int log_n);
// ---- pieces of the sharded prover (one proof over several GPUs; dist_prover.cu)
// block-local stages (bits 0 .. log_block-1) of a size-2^log_total inverse transform on one aligned block of the
// bit-reversed input; no scaling
void ntt_inverse_local_stages(pk_ctx* ctx, const fr_t* src, fr_t* dst, int log_block, int log_total);
// the upper log2(G) stages after the all-to-all, with the coset / size scaling fused (see ntt.cu)
void ntt_inverse_cross_stages(pk_ctx* ctx, const fr_t* in, fr_t* out, const fr_t* kscale, const fr_t* cscale, int G, int log_total,
                              size_t k0);
// the block-local stages with the all-to-all fused in: the last pass stores into the peers' receive buffers (see ntt.cu)
void ntt_inverse_local_stages_scatter(pk_ctx* ctx, const fr_t* src, fr_t* scratch, int log_block, int log_total, fr_t* const* peer,
                                      int world, int rank, int per_log);
// b[i] = cpow[i] * sum_u a[i + u nf] kappa^u, i < nf: restriction of a to the coset c H_nf (cpow[i] = c^i, kappa = c^nf)
void coset_fold(pk_ctx* ctx, const fr_t* a, const fr_t* cpow, const fr_t& kappa, int F, size_t nf, fr_t* b);
// out[i] = a[i] w_n^i: coefficients of a(w X)
void omega_scale(pk_ctx* ctx, const fr_t* a, fr_t* out, int log_n);
// canonical <-> Montgomery on device arrays
void fr_to_mont(pk_ctx* ctx, fr_t* data, size_t n);
void fr_from_mont(pk_ctx* ctx, const fr_t* src, fr_t* dst, size_t n);

}  // namespace pk
