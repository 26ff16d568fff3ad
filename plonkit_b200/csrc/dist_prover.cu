// ONE proof split across the GPUs of a node (BASELINE.json configs[2]: "MSM + NTT sharded across 8 x B200").
//
// Same prover as prover.cu (SetupForProver::prove, src/plonk.rs:132-176 -> bellman prove_by_steps), same bytes; the
// work is cut so that almost nothing crosses NVLink:
//
//   * commitments (11 MSMs, ~3/4 of a proof): sharded BY BASE CHUNK.  Rank r keeps the fixed-base window tables of
//     SRS[r n/G, (r+1) n/G) only and multiplies the matching coefficient chunk; the 128-byte XYZZ partial sums are
//     all-gathered and folded on the device ("all-reduce" in the EC group = all-gather + local fold: EC addition is not
//     an NCCL reduce-op).  Traffic per round: G x (<= 4) x 128 B.
//   * quotient (17 LDEs + the pointwise identity): sharded BY COSET.  The 4n-point domain 7 H_4n in the prover's slot
//     layout is a concatenation of 4 cosets of H_n; rank r owns the contiguous 1/G of it — whole cosets for G <= 4,
//     half a coset for G = 8 (the half is itself a coset of H_{n/2}: the coefficients are folded once, then an n/2-point
//     NTT runs).  Every LDE and the gate/permutation arithmetic touch only that range: zero bytes move.  Values at wX
//     (d(wX), Z(wX)) come from LDEs of the shifted polynomials, because the neighbour may sit on another rank.
//   * the size-4n inverse NTT of the quotient: block-local stages on each rank's range, ONE all-to-all (each rank sends
//     1/G of its range to every peer: 4n 32 B / G in total per rank), the last log2 G stages in registers with the
//     coset / size scaling fused, and an all-gather of the coefficients into natural order.
//   * grand product, evaluations, linearisation and opening polynomials: sharded by rows / coefficient chunk; what a
//     scan or a sum needs from the other chunks (prefix products, partial sums, division carries) travels as a few
//     32-byte scalars per rank, and the finished chunk of Z is all-gathered (n 32 B) for its inverse NTT.
//   * the four wire inverse NTTs are shared out by polynomial and all-gathered (4n 32 B); the witness gather and the
//     inverse NTT of Z are replicated.  Every rank holds the witness and derives the same transcript: no broadcast.
//
// Every rank returns the same proof.  The collectives come from comm.cuh (NCCL between processes, peer copies
// between the threads of one process).
#include "comm.cuh"
#include "keccak_host.hpp"
#include "msm.cuh"
#include "ntt.cuh"
#include "poly.cuh"

using namespace pk;

struct pk_dist_setup {
    pk_ctx* ctx = nullptr;
    int log_n = 0, G = 1, rank = 0, sub = 0;   // sub = log2 of the split of one coset (0 for G <= 4, 1 for G = 8)
    uint64_t n = 0, m = 0;                     // m = 4n / G: this rank's range of the quotient domain
    uint64_t cn = 0, clo = 0;                  // this rank's chunk of the bases / coefficients: [clo, clo + cn)
    uint32_t num_inputs = 0;
    uint64_t nvars = 0;
    bool have_witness = false;
    int parts = 1;                             // (sub-)cosets in the range: max(1, 4 / G)
    uint64_t nf = 0;                           // points per part: n >> sub
    // replicated setup data
    DevBuf<uint32_t> wire_idx;
    DevBuf<fr_t> sel_vals, sigma_vals, sel_coef, sigma_coef, vars;
    // this rank's range of the setup LDEs and tables
    DevBuf<fr_t> sel_lde, sigma_lde, l0_lde;   // [7][m], [4][m], [m]
    DevBuf<fr_t> cpow;                         // [parts][nf]: c^i of each part's coset shift c
    fr_t kappa[4];                             // c^nf per part
    DevBuf<fr_t> kscale;                       // [m / G]: 7^-(k0 + k') / 4n
    fr_t cscale[8];                            // 7^-(c m)
    // per-proof working set
    DevBuf<fr_t> w_nat, w_br, w_coef;          // [4][n]
    DevBuf<fr_t> w_lde;                        // [4][m]
    DevBuf<fr_t> z_coef, z_lde, znext_lde, dnext_lde, pi_coef, pi_lde;
    DevBuf<fr_t> t_part, a2a, t4;              // [m], [m], [4n]
    DevBuf<fr_t> tmp_a, tmp_b, tmp_c, fold;
    DevBuf<fr_t> zpow, zinvpow, zwpow, zwinvpow, r_coef;
    DevBuf<g1_xyzz_t> part_pts, all_pts;       // [16], [G][16]
    DevBuf<fr_t> xs, xr;                       // [16], [G][16]: partial scalars of a sharded scan / evaluation and their all-gather
    DevBuf<fr_t> zchunk;                       // [cn]
    DevBuf<fr_t> shifted;                      // [n]: coefficients of d(wX) / Z(wX) on their way into an LDE (side stream)
    // Fused all-to-all: when every rank can map its peers' receive buffers (NVLink peer access in one process, CUDA IPC
    // between processes), the last block-local pass of the quotient's inverse NTT stores straight into the peers' memory
    bool fused = false;
    fr_t* peer_a2a[8] = {};                    // every rank's a2a     [G][m/G]
};

namespace pk {

// host transcript helpers (same as prover.cu)
static fr_t d_fr_from_limbs32(const uint32_t v[8]) {
    fr_t x;
    for (int i = 0; i < 8; ++i) x.v[i] = v[i];
    return x.to_mont();
}
static void d_commit_fr(RollingKeccakTranscript& tr, const fr_t& x) {
    fr_t c = x.from_mont();
    tr.commit_limbs(c.v);
}
static void d_commit_g1(RollingKeccakTranscript& tr, const g1_affine_t& p) {
    if (p.is_inf()) {
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        tr.commit_limbs(z);
        tr.commit_limbs(z);
        return;
    }
    fq_t x = p.x.from_mont(), y = p.y.from_mont();
    tr.commit_limbs(x.v);
    tr.commit_limbs(y.v);
}
static fr_t d_challenge(RollingKeccakTranscript& tr) {
    uint32_t c[8];
    tr.challenge(c);
    return d_fr_from_limbs32(c);
}
static void d_fr_to_abi(const fr_t& x, uint64_t out[4]) {
    fr_t c = x.from_mont();
    memcpy(out, c.v, 32);
}

std::vector<uint32_t> build_sigma_targets(const uint32_t* wire_idx, uint64_t n, uint64_t nvars);  // prover.cu

// fold of G x nb partial sums: out[k] = sum_q all[q][k]
__global__ void g1_fold_kernel(const g1_xyzz_t* all, int G, int stride, int nb, g1_xyzz_t* out) {
    const int k = threadIdx.x;
    if (k >= nb) return;
    g1_xyzz_t acc = ld_xyzz(all + k);
    for (int q = 1; q < G; ++q) acc = acc.add(ld_xyzz(all + (size_t)q * stride + k));
    st_xyzz(out + k, acc);
}

// out[k] = sum_i polys[k][i] * SRS[i] over the whole key: local chunk MSM, all-gather of the partial sums, device fold
static void dist_commit(pk_dist_setup* s, const fr_t* const* polys, int nb, g1_affine_t* out) {
    pk_ctx* ctx = s->ctx;
    PK_REQUIRE(nb >= 1 && nb <= 16, PK_ERR_INVALID, "commit batch too large");
    const fr_t* chunk[16];
    for (int k = 0; k < nb; ++k) chunk[k] = polys[k] + s->clo;
    msm_run_batch_dev(ctx, chunk, nb, s->cn, 0, s->part_pts.p);
    ctx->comm->all_gather(s->part_pts.p, s->all_pts.p, 16 * sizeof(g1_xyzz_t), ctx->stream);
    g1_fold_kernel<<<1, 32, 0, ctx->stream>>>(s->all_pts.p, s->G, 16, nb, s->part_pts.p);
    ctx->prof.kernel_launches++;
    g1_xyzz_t* host = reinterpret_cast<g1_xyzz_t*>(ctx->pinned);
    PK_CUDA(cudaMemcpyAsync(host, s->part_pts.p, nb * sizeof(g1_xyzz_t), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < nb; ++k) out[k] = host[k].to_affine();
}

// CUDA-event bracket around one bulk collective (profiling only): comm_ms[slot] += its device time
struct CommTimer {
    pk_ctx* ctx;
    int slot;
    cudaEvent_t a = nullptr, b = nullptr;
    CommTimer(pk_ctx* c, int sl, uint64_t bytes) : ctx(c), slot(sl) {
        if (!ctx->prof.enabled) return;
        ctx->prof.comm_bytes[slot] += bytes;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        cudaEventRecord(a, ctx->stream);
    }
    ~CommTimer() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        cudaEventSynchronize(b);
        float ms = 0;
        cudaEventElapsedTime(&ms, a, b);
        ctx->prof.comm_ms[slot] += ms;
        cudaEventDestroy(a);
        cudaEventDestroy(b);
    }
};

// all-gather of the (<= 16) field elements in s->xs; host copy [G][16] returned (synchronises)
static std::vector<fr_t> gather_scalars(pk_dist_setup* s) {
    pk_ctx* ctx = s->ctx;
    ctx->comm->all_gather(s->xs.p, s->xr.p, 16 * sizeof(fr_t), ctx->stream);
    std::vector<fr_t> h((size_t)s->G * 16);
    PK_CUDA(cudaMemcpyAsync(h.data(), s->xr.p, h.size() * sizeof(fr_t), cudaMemcpyDeviceToHost, ctx->stream));
    PK_CUDA(cudaStreamSynchronize(ctx->stream));
    return h;
}
// results[k] = sum_i polys[k][i] pows[k][i] over the whole vectors: every rank sums its chunk, the partial sums are all-gathered
static void dist_dot_batch(pk_dist_setup* s, int npoly, const fr_t* const* polys, const fr_t* const* pows, fr_t* results) {
    const fr_t* pp[16];
    const fr_t* ww[16];
    for (int k = 0; k < npoly; ++k) { pp[k] = polys[k] + s->clo; ww[k] = pows[k] + s->clo; }
    poly_dot_batch_dev(s->ctx, npoly, pp, ww, s->cn, s->xs.p);
    std::vector<fr_t> h = gather_scalars(s);
    for (int k = 0; k < npoly; ++k) {
        fr_t acc = h[k];
        for (int q = 1; q < s->G; ++q) acc = acc + h[(size_t)q * 16 + k];
        results[k] = acc;
    }
}

// evaluations of the coefficient vector `coef` (n) on this rank's range of the quotient domain -> out (m)
static void range_lde(pk_dist_setup* s, const fr_t* coef, fr_t* out) {
    pk_ctx* ctx = s->ctx;
    const int F = 1 << s->sub;
    for (int p = 0; p < s->parts; ++p) {
        coset_fold(ctx, coef, s->cpow.p + (size_t)p * s->nf, s->kappa[p], F, s->nf, s->fold.p);
        ntt_forward_bitrev(ctx, s->fold.p, out + (size_t)p * s->nf, s->log_n - s->sub);
    }
}

static void dist_setup_poly(pk_dist_setup* s, const fr_t* vals_nat, fr_t* coef, fr_t* lde) {
    pk_ctx* ctx = s->ctx;
    bitrev_permute(ctx, vals_nat, s->tmp_a.p, s->log_n);
    ntt_inverse_from_bitrev(ctx, s->tmp_a.p, coef, s->log_n);
    range_lde(s, coef, lde);
}

__global__ void geom_table_kernel(fr_t* out, fr_t g, fr_t c0, size_t start, size_t n) {  // out[i] = c0 * g^(start + i)
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) st_fp(out + i, g.pow_u64(start + i) * c0);
}

void dist_setup_create(pk_ctx* ctx, const pk_assembly* as, pk_dist_setup** out) {
    PK_REQUIRE(as && out, PK_ERR_INVALID, "null argument");
    PK_REQUIRE(ctx->comm != nullptr, PK_ERR_INVALID, "context is not attached to a communicator (pk_comm_attach_*)");
    const int G = ctx->comm->world, rank = ctx->comm->rank;
    PK_REQUIRE(G == 1 || G == 2 || G == 4 || G == 8, PK_ERR_INVALID, "the sharded prover runs on 1, 2, 4 or 8 ranks");
    const uint64_t n = as->n;
    PK_REQUIRE(n >= 2 && (n & (n - 1)) == 0, PK_ERR_INVALID, "domain size must be a power of two >= 2");
    PK_REQUIRE(n >= (uint64_t)G * G, PK_ERR_INVALID, "circuit too small to shard over this many ranks");
    const int log_n = ilog2(n);
    PK_REQUIRE(log_n + 2 <= 28, PK_ERR_DEGREE_TOO_LARGE, "circuit larger than 2^26 gates (SETUP_MAX_POW2, src/plonk.rs:27)");
    PK_REQUIRE(ctx->srs && ctx->srs->n >= n / G, PK_ERR_DEGREE_TOO_LARGE, "this rank's SRS chunk is smaller than its share of the domain");
    PK_REQUIRE(as->num_inputs < n, PK_ERR_INVALID, "too many public inputs");
    PK_REQUIRE(as->nvars >= 1 && as->nvars < (uint64_t(1) << 32), PK_ERR_INVALID, "nvars out of range");
    std::vector<uint32_t> target = build_sigma_targets(as->wire_idx, n, as->nvars);

    pk_dist_setup* s = new pk_dist_setup();
    try {
        cudaStream_t st = ctx->stream;
        s->ctx = ctx; s->log_n = log_n; s->n = n; s->G = G; s->rank = rank;
        s->num_inputs = (uint32_t)as->num_inputs; s->nvars = as->nvars;
        s->m = 4 * n / G;
        s->cn = n / G; s->clo = (uint64_t)rank * s->cn;
        s->sub = G > 4 ? ilog2(G / 4) : 0;
        s->parts = G >= 4 ? 1 : 4 / G;
        s->nf = n >> s->sub;
        s->wire_idx.alloc(4 * n);
        s->sel_vals.alloc(7 * n); s->sigma_vals.alloc(4 * n); s->sel_coef.alloc(7 * n); s->sigma_coef.alloc(4 * n);
        s->vars.alloc(as->nvars);
        s->sel_lde.alloc(7 * s->m); s->sigma_lde.alloc(4 * s->m); s->l0_lde.alloc(s->m);
        s->cpow.alloc((size_t)s->parts * s->nf);
        s->kscale.alloc(s->m / G);
        s->w_nat.alloc(4 * n); s->w_br.alloc(4 * n); s->w_coef.alloc(4 * n); s->w_lde.alloc(4 * s->m);
        s->z_coef.alloc(n); s->z_lde.alloc(s->m); s->znext_lde.alloc(s->m); s->dnext_lde.alloc(s->m);
        s->pi_coef.alloc(n); s->pi_lde.alloc(s->m);
        s->t_part.alloc(s->m); s->a2a.alloc(s->m); s->t4.alloc(4 * n);
        s->tmp_a.alloc(n); s->tmp_b.alloc(n); s->tmp_c.alloc(n); s->fold.alloc(s->nf);
        s->zpow.alloc(n); s->zinvpow.alloc(n); s->zwpow.alloc(n); s->zwinvpow.alloc(n); s->r_coef.alloc(n);
        s->part_pts.alloc(16); s->all_pts.alloc((size_t)G * 16);
        s->xs.alloc(16); s->xr.alloc((size_t)G * 16); s->zchunk.alloc(s->cn); s->shifted.alloc(n);

        // coset shifts of this rank's parts.  Slot sl of the layout is the coset g_sl H_n with g_sl = 7 w_4n^brev2(sl);
        // part q of a slot split 2^sub ways holds the natural indices j = brev_sub(q) mod 2^sub, i.e. the coset
        // (g_sl w_n^brev_sub(q)) H_{n / 2^sub}.
        ensure_twiddles(ctx, log_n + 2);
        fr_t g7, g7inv;
        for (int i = 0; i < 8; ++i) { g7.v[i] = FrRoots::gen7(i); g7inv.v[i] = FrRoots::gen7_inv(i); }
        const fr_t w4 = host_root_of_unity(log_n + 2), wn = host_root_of_unity(log_n);
        static const int brev2[4] = {0, 2, 1, 3};
        const uint64_t lo4 = (uint64_t)rank * s->m;
        for (int p = 0; p < s->parts; ++p) {
            const uint64_t pos = lo4 + (uint64_t)p * s->nf;       // first position of the part in the slot layout
            const int sl = (int)(pos >> log_n);
            const uint64_t q = (pos & (n - 1)) / s->nf;           // which part of the slot
            uint64_t jl = 0;                                      // brev_sub(q)
            for (int b = 0; b < s->sub; ++b) jl |= ((q >> b) & 1) << (s->sub - 1 - b);
            const fr_t c = g7 * w4.pow_u64(brev2[sl]) * wn.pow_u64(jl);
            s->kappa[p] = c.pow_u64(s->nf);
            geom_table_kernel<<<(unsigned)((s->nf + 255) / 256), 256, 0, st>>>(s->cpow.p + (size_t)p * s->nf, c, fr_t::one(), 0, s->nf);
        }
        // scaling of the size-4n inverse: coefficient index c m + k0 + k' gets 7^-index / 4n
        const fr_t inv4n = fr_t::from_u32(2).inverse().pow_u64(log_n + 2);
        const uint64_t per = s->m / G, k0 = (uint64_t)rank * per;
        geom_table_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(s->kscale.p, g7inv, inv4n, k0, per);
        for (int c = 0; c < 8; ++c) s->cscale[c] = c < G ? g7inv.pow_u64((uint64_t)c * s->m) : fr_t::one();
        ctx->prof.kernel_launches += s->parts + 1;

        PK_CUDA(cudaMemcpyAsync(s->wire_idx.p, as->wire_idx, 4 * n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        PK_CUDA(cudaMemcpyAsync(s->sel_vals.p, as->selectors, 7 * n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
        fr_to_mont(ctx, s->sel_vals.p, 7 * n);
        DevBuf<uint32_t> d_target(4 * n);
        PK_CUDA(cudaMemcpyAsync(d_target.p, target.data(), 4 * n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        sigma_values(ctx, d_target.p, s->sigma_vals.p, log_n);
        for (int k = 0; k < 7; ++k) dist_setup_poly(s, s->sel_vals.p + k * n, s->sel_coef.p + k * n, s->sel_lde.p + k * s->m);
        for (int k = 0; k < 4; ++k) dist_setup_poly(s, s->sigma_vals.p + k * n, s->sigma_coef.p + k * n, s->sigma_lde.p + k * s->m);
        // L_0(X) = (1/n) sum_j X^j on the range
        fr_fill(ctx, s->tmp_b.p, fr_t::from_u32(2).inverse().pow_u64(log_n), n);
        range_lde(s, s->tmp_b.p, s->l0_lde.p);
        PK_CUDA(cudaStreamSynchronize(st));
        PK_CUDA(cudaGetLastError());
        // peer mappings for the fused exchanges (PK_DIST_FUSED=0 keeps the collectives)
        static const int want_fused = [] { const char* e = getenv("PK_DIST_FUSED"); return e ? atoi(e) : 1; }();
        if (want_fused && G > 1) {
            void* tmp[8];
            s->fused = ctx->comm->map_peers(s->a2a.p, s->m * sizeof(fr_t), tmp, st);
            for (int q = 0; q < G; ++q) s->peer_a2a[q] = static_cast<fr_t*>(tmp[q]);
        }
    } catch (...) {
        delete s;
        throw;
    }
    *out = s;
}

void dist_setup_commitments(pk_ctx* ctx, pk_dist_setup* s, uint64_t out_xy[11][8]) {
    (void)ctx;
    const fr_t* polys[11];
    for (int k = 0; k < 7; ++k) polys[k] = s->sel_coef.p + k * s->n;
    for (int k = 0; k < 4; ++k) polys[7 + k] = s->sigma_coef.p + k * s->n;
    g1_affine_t out[11];
    dist_commit(s, polys, 11, out);
    for (int k = 0; k < 11; ++k) affine_to_abi(out[k], out_xy[k]);
}

void dist_witness_upload(pk_ctx* ctx, pk_dist_setup* s, const uint64_t* var_values, uint64_t nvars) {
    PK_REQUIRE(var_values != nullptr, PK_ERR_ASSIGNMENT_MISSING, "witness is null");
    PK_REQUIRE(nvars == s->nvars, PK_ERR_ASSIGNMENT_MISSING, "witness length does not match the circuit");
    PK_CUDA(cudaMemcpyAsync(s->vars.p, var_values, nvars * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    fr_to_mont(ctx, s->vars.p, nvars);
    s->have_witness = true;
}

void dist_prove(pk_ctx* ctx, pk_dist_setup* s, const uint64_t* var_values, uint64_t nvars, pk_proof* proof, uint64_t* inputs_out) {
    PK_REQUIRE(proof != nullptr, PK_ERR_INVALID, "null proof");
    PK_REQUIRE(ctx->comm != nullptr && ctx->comm->world == s->G && ctx->comm->rank == s->rank, PK_ERR_INVALID,
               "setup belongs to another communicator");
    PK_REQUIRE(ctx->srs && ctx->srs->n >= s->cn, PK_ERR_DEGREE_TOO_LARGE, "this rank's SRS chunk is smaller than its share of the domain");
    const uint64_t n = s->n, m = s->m;
    const int log_n = s->log_n, G = s->G;
    const uint32_t ni = s->num_inputs;
    cudaStream_t st = ctx->stream;
    struct PhaseEvents {  // destroyed on every exit path
        cudaEvent_t e[8];
        PhaseEvents() { for (auto& x : e) cudaEventCreate(&x); }
        ~PhaseEvents() { for (auto& x : e) cudaEventDestroy(x); }
    } phase_events;
    cudaEvent_t* ev = phase_events.e;
    int evk = 0;
    auto mark = [&] { cudaEventRecord(ev[evk++], st); };
    mark();
    for (int i = 0; i < 3; ++i) { ctx->prof.comm_ms[i] = 0; ctx->prof.comm_bytes[i] = 0; }
    if (var_values) dist_witness_upload(ctx, s, var_values, nvars);
    PK_REQUIRE(s->have_witness, PK_ERR_ASSIGNMENT_MISSING, "no witness uploaded");
    mark();

    // fused exchanges store into the peers' buffers: nobody may still be reading them for the previous proof
    if (s->fused) ctx->comm->stream_barrier(st);

    // ---- witness -> wire values; is_satisfied_using_one_shot_check (src/plonk.rs:137).  Replicated: every rank reaches
    // the same verdict before the first collective.
    wire_gather(ctx, s->vars.p, s->wire_idx.p, s->w_nat.p, s->w_br.p, log_n);
    PK_REQUIRE(gate_check(ctx, s->w_nat.p, s->sel_vals.p, ni, log_n), PK_ERR_UNSATISFIED, "witness does not satisfy the circuit");
    std::vector<fr_t> inputs(ni);
    if (ni) {
        PK_CUDA(cudaMemcpyAsync(inputs.data(), s->w_nat.p, ni * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
        PK_CUDA(cudaStreamSynchronize(st));
    }
    RollingKeccakTranscript tr;
    for (uint32_t i = 0; i < ni; ++i) {
        d_commit_fr(tr, inputs[i]);
        if (inputs_out) d_fr_to_abi(inputs[i], inputs_out + 4 * i);
    }

    // ---- round 1
    g1_affine_t Cw[4];
    {
        // the four wire inverse NTTs are shared out by polynomial: rank r transforms the wire(s) that overlap block r of the
        // [4][n] coefficient array and the blocks are all-gathered in place (4n 32 B in total)
        const fr_t* polys[4];
        const uint64_t blk = 4 * n / G, b0 = (uint64_t)s->rank * blk;
        for (int c = (int)(b0 / n); c <= (int)((b0 + blk - 1) / n); ++c)
            ntt_inverse_from_bitrev(ctx, s->w_br.p + c * n, s->w_coef.p + c * n, log_n);
        {
            CommTimer t(ctx, 0, (uint64_t)(G - 1) * blk * sizeof(fr_t));
            ctx->comm->all_gather(s->w_coef.p + b0, s->w_coef.p, blk * sizeof(fr_t), st);
        }
        for (int c = 0; c < 4; ++c) polys[c] = s->w_coef.p + c * n;
        {
            // this rank's range of the LDEs that need only the wire coefficients and the public inputs — the four wires,
            // d(wX), PI — on the side stream, beside the commitment kernels
            SideStreamScope side(ctx);
            for (int c = 0; c < 4; ++c) range_lde(s, s->w_coef.p + c * n, s->w_lde.p + c * m);
            omega_scale(ctx, s->w_coef.p + 3 * n, s->shifted.p, log_n);
            range_lde(s, s->shifted.p, s->dnext_lde.p);
            if (ni <= 1) {
                // PI(X) = in_0 L_0(X) = (in_0 / n) sum_j X^j: no transform needed for the common single-input circuit
                const fr_t c0 = ni ? inputs[0] * fr_t::from_u32(2).inverse().pow_u64(log_n) : fr_t::zero();
                fr_fill(ctx, s->pi_coef.p, c0, n);
            } else {
                PK_CUDA(cudaMemsetAsync(s->pi_coef.p, 0, n * sizeof(fr_t), ctx->stream));
                pi_scatter(ctx, s->w_nat.p, s->pi_coef.p, ni, log_n);
                ntt_inverse_from_bitrev(ctx, s->pi_coef.p, s->pi_coef.p, log_n);
            }
            range_lde(s, s->pi_coef.p, s->pi_lde.p);
        }
        dist_commit(s, polys, 4, Cw);
        for (int c = 0; c < 4; ++c) d_commit_g1(tr, Cw[c]);
    }
    const fr_t beta = d_challenge(tr), gamma = d_challenge(tr);
    mark();

    // ---- round 2: the grand product, sharded by rows.  Every rank scans ITS chunk of the numerators / denominators; the
    // chunk totals are all-gathered, the prefix (suffix) products of the chunks below (above) and the one field inversion
    // happen on the host, and the finished chunk of Z is all-gathered (n 32 B in total) for the replicated inverse NTT.
    {
        const uint64_t cn = s->cn, clo = s->clo;
        perm_num_den_range(ctx, s->w_nat.p, s->sigma_vals.p, beta, gamma, s->tmp_a.p, s->tmp_b.p, log_n, clo, cn);
        poly_scan(ctx, true, false, s->tmp_a.p, s->tmp_a.p, cn);  // pn[i] = prod_{lo <= j <= lo + i} num_j
        poly_scan(ctx, true, true, s->tmp_b.p, s->tmp_b.p, cn);   // sd[i] = prod_{lo + i <= j < hi} den_j
        PK_CUDA(cudaMemcpyAsync(s->xs.p, s->tmp_a.p + cn - 1, sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
        PK_CUDA(cudaMemcpyAsync(s->xs.p + 1, s->tmp_b.p, sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
        std::vector<fr_t> h = gather_scalars(s);
        fr_t below = fr_t::one(), above = fr_t::one(), total_den = fr_t::one();
        for (int q = 0; q < G; ++q) {
            if (q < s->rank) below = below * h[(size_t)q * 16];
            if (q > s->rank) above = above * h[(size_t)q * 16 + 1];
            total_den = total_den * h[(size_t)q * 16 + 1];
        }
        PK_REQUIRE(!total_den.is_zero(), PK_ERR_DIVISION_BY_ZERO, "zero denominator in the permutation grand product");
        z_finish_chunk(ctx, s->tmp_a.p, s->tmp_b.p, below * above * total_den.inverse(), s->zchunk.p, clo, cn);
        {
            CommTimer t(ctx, 0, (uint64_t)(G - 1) * cn * sizeof(fr_t));
            ctx->comm->all_gather(s->zchunk.p, s->tmp_c.p, cn * sizeof(fr_t), st);
        }
        bitrev_permute(ctx, s->tmp_c.p, s->tmp_a.p, log_n);
        ntt_inverse_from_bitrev(ctx, s->tmp_a.p, s->z_coef.p, log_n);
    }
    g1_affine_t Cz;
    {
        {
            SideStreamScope side(ctx);  // Z and Z(wX) on this rank's range, beside the commitment of Z
            range_lde(s, s->z_coef.p, s->z_lde.p);
            omega_scale(ctx, s->z_coef.p, s->shifted.p, log_n);
            range_lde(s, s->shifted.p, s->znext_lde.p);
        }
        const fr_t* polys[1] = {s->z_coef.p};
        dist_commit(s, polys, 1, &Cz);
    }
    d_commit_g1(tr, Cz);
    const fr_t alpha = d_challenge(tr);
    mark();

    // ---- round 3: the quotient on this rank's range of the coset domain
    side_join(ctx);  // every LDE of this round was started beside the commitments of rounds 1 and 2
    QuotientArgs qa;
    qa.num_direct_inputs = -1;
    for (int c = 0; c < 4; ++c) { qa.w[c] = s->w_lde.p + c * m; qa.sig[c] = s->sigma_lde.p + c * m; }
    for (int k = 0; k < 7; ++k) qa.sel[k] = s->sel_lde.p + k * m;
    qa.z = s->z_lde.p; qa.pi = s->pi_lde.p; qa.l0 = s->l0_lde.p; qa.out = s->t_part.p;
    qa.beta = beta; qa.gamma = gamma; qa.alpha = alpha; qa.log_n = log_n;
    qa.range_lo = (size_t)s->rank * m; qa.range_len = m;
    qa.w3_next = s->dnext_lde.p; qa.z_next = s->znext_lde.p;
    quotient_slots(ctx, qa);
    // size-4n inverse: local stages, all-to-all, cross stages (+ scaling), all-gather of the coefficients
    const uint64_t per = m / G;
    if (s->fused) {
        // the all-to-all rides on the last local pass: every element is stored straight into the receive buffer of the one
        // rank that needs it (peer stores over NVLink, no amplification), then one stream barrier
        ntt_inverse_local_stages_scatter(ctx, s->t_part.p, s->t_part.p, ilog2(m), log_n + 2, s->peer_a2a, G, s->rank, ilog2(per));
        CommTimer t(ctx, 1, (uint64_t)(G - 1) * per * sizeof(fr_t));
        ctx->comm->stream_barrier(st);
    } else {
        ntt_inverse_local_stages(ctx, s->t_part.p, s->t_part.p, ilog2(m), log_n + 2);
        CommTimer t(ctx, 1, (uint64_t)(G - 1) * per * sizeof(fr_t));
        ctx->comm->all_to_all(s->t_part.p, s->a2a.p, per * sizeof(fr_t), st);
    }
    ntt_inverse_cross_stages(ctx, s->a2a.p, s->t_part.p, s->kscale.p, s->cscale, G, log_n + 2, (size_t)s->rank * per);
    {
        // the coefficients are needed by every rank: an all-gather (NCCL replicates in the switch; storing them G times
        // from the kernel was measured and is slower)
        const void* send[8];
        void* recv[8];
        for (int c = 0; c < G; ++c) { send[c] = s->t_part.p + (size_t)c * per; recv[c] = s->t4.p + (size_t)c * m; }
        CommTimer t(ctx, 2, (uint64_t)(G - 1) * G * per * sizeof(fr_t));
        ctx->comm->all_gather_multi(send, recv, G, per * sizeof(fr_t), st);
    }
    {
        fr_t top[3];
        PK_CUDA(cudaMemcpyAsync(top, s->t4.p + 4 * n - 3, 3 * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
        PK_CUDA(cudaStreamSynchronize(st));
        PK_REQUIRE(top[0].is_zero() && top[1].is_zero() && top[2].is_zero(), PK_ERR_UNSATISFIED, "quotient is not a polynomial");
    }
    g1_affine_t Ct[4];
    {
        const fr_t* polys[4] = {s->t4.p, s->t4.p + n, s->t4.p + 2 * n, s->t4.p + 3 * n};
        dist_commit(s, polys, 4, Ct);
        for (int i = 0; i < 4; ++i) d_commit_g1(tr, Ct[i]);
    }
    const fr_t zeta = d_challenge(tr);
    mark();

    // ---- round 4: evaluations, sharded by coefficient chunk (partial sums all-gathered: 13 x 32 B per rank)
    const uint64_t cn = s->cn, clo = s->clo;
    const fr_t omega = host_root_of_unity(log_n);
    const fr_t zeta_omega = zeta * omega;
    poly_powers_from(ctx, s->zpow.p + clo, zeta, clo, cn);
    poly_powers_from(ctx, s->zwpow.p + clo, zeta_omega, clo, cn);
    fr_t evv[13];
    {
        const fr_t* polys[13];
        const fr_t* pows[13];
        for (int c = 0; c < 4; ++c) polys[c] = s->w_coef.p + c * n;
        for (int c = 0; c < 3; ++c) polys[4 + c] = s->sigma_coef.p + c * n;
        for (int i = 0; i < 4; ++i) polys[7 + i] = s->t4.p + i * n;
        for (int k = 0; k < 11; ++k) pows[k] = s->zpow.p;
        polys[11] = s->w_coef.p + 3 * n; pows[11] = s->zwpow.p;
        polys[12] = s->z_coef.p; pows[12] = s->zwpow.p;
        dist_dot_batch(s, 13, polys, pows, evv);
    }
    const fr_t* wz = evv;
    const fr_t* sz = evv + 4;
    const fr_t dzw = evv[11], zzw = evv[12];
    const fr_t zeta_n = zeta.pow_u64(n);
    const fr_t zn2 = zeta_n.sqr(), zn3 = zn2 * zeta_n;
    const fr_t tz = evv[7] + zeta_n * evv[8] + zn2 * evv[9] + zn3 * evv[10];
    PK_REQUIRE(!(zeta - fr_t::one()).is_zero(), PK_ERR_DIVISION_BY_ZERO, "challenge z hit the domain");
    const fr_t n_fr = fr_t::from_u32(2).pow_u64(log_n);
    const fr_t l0z = (zeta_n - fr_t::one()) * (n_fr * (zeta - fr_t::one())).inverse();
    static const uint32_t KK[4] = {1, 5, 7, 10};
    fr_t zfac = alpha;
    for (int i = 0; i < 4; ++i) zfac = zfac * (wz[i] + beta * fr_t::from_u32(KK[i]) * zeta + gamma);
    zfac = zfac + alpha.sqr() * l0z;
    fr_t sfac = alpha * beta * zzw;
    for (int i = 0; i < 3; ++i) sfac = sfac * (wz[i] + beta * sz[i] + gamma);
    {
        const fr_t* in[9] = {s->sel_coef.p + 5 * n + clo, s->sel_coef.p + clo, s->sel_coef.p + n + clo, s->sel_coef.p + 2 * n + clo,
                             s->sel_coef.p + 3 * n + clo, s->sel_coef.p + 4 * n + clo, s->sel_coef.p + 6 * n + clo, s->z_coef.p + clo,
                             s->sigma_coef.p + 3 * n + clo};
        fr_t coef[9] = {fr_t::one(), wz[0], wz[1], wz[2], wz[3], wz[0] * wz[1], dzw, zfac, sfac.neg()};
        poly_lincomb(ctx, s->r_coef.p + clo, 9, in, coef, cn);
    }
    fr_t rz;
    {
        const fr_t* polys[1] = {s->r_coef.p};
        const fr_t* pows[1] = {s->zpow.p};
        dist_dot_batch(s, 1, polys, pows, &rz);
    }
    for (int c = 0; c < 4; ++c) d_commit_fr(tr, wz[c]);
    d_commit_fr(tr, dzw);
    for (int c = 0; c < 3; ++c) d_commit_fr(tr, sz[c]);
    d_commit_fr(tr, tz);
    d_commit_fr(tr, rz);
    d_commit_fr(tr, zzw);
    const fr_t v = d_challenge(tr);
    mark();

    // ---- round 5: opening polynomials, every rank its own chunk of the coefficients.  The synthetic division is a suffix
    // sum: the chunk's local sums plus a carry from the chunks above (their totals are all-gathered: 2 x 32 B per rank).
    fr_t vp[11];
    vp[0] = fr_t::one();
    for (int i = 1; i <= 10; ++i) vp[i] = vp[i - 1] * v;
    {
        const fr_t* in[12] = {s->t4.p + clo, s->t4.p + n + clo, s->t4.p + 2 * n + clo, s->t4.p + 3 * n + clo, s->r_coef.p + clo,
                              s->w_coef.p + clo, s->w_coef.p + n + clo, s->w_coef.p + 2 * n + clo, s->w_coef.p + 3 * n + clo,
                              s->sigma_coef.p + clo, s->sigma_coef.p + n + clo, s->sigma_coef.p + 2 * n + clo};
        fr_t coef[12] = {fr_t::one(), zeta_n, zn2, zn3, vp[1], vp[2], vp[3], vp[4], vp[5], vp[6], vp[7], vp[8]};
        poly_lincomb(ctx, s->tmp_a.p + clo, 12, in, coef, cn);
    }
    {
        const fr_t* in[2] = {s->z_coef.p + clo, s->w_coef.p + 3 * n + clo};
        fr_t coef[2] = {vp[9], vp[10]};
        poly_lincomb(ctx, s->tmp_b.p + clo, 2, in, coef, cn);
    }
    PK_REQUIRE(!zeta.is_zero(), PK_ERR_DIVISION_BY_ZERO, "challenge z is zero");
    poly_powers_from(ctx, s->zinvpow.p + clo, zeta.inverse(), clo + 1, cn);        // z^-(k + 1), k in the chunk
    poly_powers_from(ctx, s->zwinvpow.p + clo, zeta_omega.inverse(), clo + 1, cn);
    // suffix sums of agg(X) z^j (into r_coef's chunk... r is folded into agg already) and of agg2(X) (z w)^j
    poly_divide_linear_chunk_scan(ctx, s->tmp_a.p + clo, s->zpow.p + clo, s->tmp_c.p + clo, cn);
    poly_divide_linear_chunk_scan(ctx, s->tmp_b.p + clo, s->zwpow.p + clo, s->zchunk.p, cn);
    PK_CUDA(cudaMemcpyAsync(s->xs.p, s->tmp_c.p + clo, sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
    PK_CUDA(cudaMemcpyAsync(s->xs.p + 1, s->zchunk.p, sizeof(fr_t), cudaMemcpyDeviceToDevice, st));
    {
        std::vector<fr_t> h = gather_scalars(s);
        fr_t carry1 = fr_t::zero(), carry2 = fr_t::zero();
        for (int q = s->rank + 1; q < G; ++q) { carry1 = carry1 + h[(size_t)q * 16]; carry2 = carry2 + h[(size_t)q * 16 + 1]; }
        poly_divide_linear_chunk_finish(ctx, s->tmp_c.p + clo, s->zinvpow.p + clo, carry1, s->r_coef.p + clo, cn);
        poly_divide_linear_chunk_finish(ctx, s->zchunk.p, s->zwinvpow.p + clo, carry2, s->tmp_a.p + clo, cn);
    }
    g1_affine_t Wz[2];
    {
        const fr_t* polys[2] = {s->r_coef.p, s->tmp_a.p};
        dist_commit(s, polys, 2, Wz);
    }
    mark();
    cudaEventSynchronize(ev[evk - 1]);
    for (int i = 0; i + 1 < evk && i < 7; ++i) {
        float ms = 0;
        cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
        ctx->prof.phase_ms[i] = ms;
    }
    {
        float tot = 0;
        cudaEventElapsedTime(&tot, ev[0], ev[evk - 1]);
        ctx->prof.phase_ms[7] = tot;
    }

    memset(proof, 0, sizeof(*proof));
    proof->n = n - 1;
    proof->num_inputs = ni;
    for (int c = 0; c < 4; ++c) {
        affine_to_abi(Cw[c], proof->wire_commitments[c]);
        affine_to_abi(Ct[c], proof->quotient_poly_commitments[c]);
        d_fr_to_abi(wz[c], proof->wire_values_at_z[c]);
    }
    affine_to_abi(Cz, proof->grand_product_commitment);
    d_fr_to_abi(dzw, proof->wire_values_at_z_omega[0]);
    d_fr_to_abi(zzw, proof->grand_product_at_z_omega);
    d_fr_to_abi(tz, proof->quotient_polynomial_at_z);
    d_fr_to_abi(rz, proof->linearization_polynomial_at_z);
    for (int c = 0; c < 3; ++c) d_fr_to_abi(sz[c], proof->permutation_polynomials_at_z[c]);
    affine_to_abi(Wz[0], proof->opening_at_z_proof);
    affine_to_abi(Wz[1], proof->opening_at_z_omega_proof);
    d_fr_to_abi(beta, proof->challenges[0]);
    d_fr_to_abi(gamma, proof->challenges[1]);
    d_fr_to_abi(alpha, proof->challenges[2]);
    d_fr_to_abi(zeta, proof->challenges[3]);
    d_fr_to_abi(v, proof->challenges[4]);
}

void dist_setup_free(pk_dist_setup* s) {
    if (!s) return;
    if (s->ctx) {
        cudaSetDevice(s->ctx->device);
        cudaStreamSynchronize(s->ctx->stream);
        if (s->ctx->side) cudaStreamSynchronize(s->ctx->side);  // LDEs of an aborted proof may still be in flight there
    }
    delete s;
}

}  // namespace pk
