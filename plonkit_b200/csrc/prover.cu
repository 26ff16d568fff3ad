// Host orchestration of the PLONK prover on the device kernels.
//
// Restates bellman_ce's better_cs prover as plonkit drives it (SetupForProver::prepare_setup_for_prover,
// make_verification_key, prove: src/plonk.rs:97-176 -> setup / make_verification_key / prove_by_steps [ext],
// Cargo.lock:109-111).  The algebra is SURVEY.md App. A (cross-checked against contrib/template.sol:445-758);
// every proof element is a canonical affine point or field element, so equal mathematics gives equal bytes.
//
// Differences from the reference's structure, all invisible in the output:
//   * setup-dependent data (11 setup polynomials, their 4n coset evaluations, sigma values, L_0 on the coset,
//     twiddles, SRS window tables) stay resident on the device across proofs; the reference rebuilds them per call
//     (precomputations = None, src/plonk.rs:156);
//   * evaluations are kept in bit-reversed order between forward and inverse NTTs;
//   * the grand product is two multiplicative scans and ONE field inversion instead of a batch inversion + serial
//     product; division by (X - z) is an additive suffix scan instead of a serial recurrence.
// The Fiat-Shamir transcript (Keccak) and a few dozen scalar operations per round run on the host.
#include "keccak_host.hpp"
#include "msm.cuh"
#include "ntt.cuh"
#include "poly.cuh"

using namespace pk;

struct pk_setup {
    pk_ctx* ctx = nullptr;
    int log_n = 0;
    uint64_t n = 0;
    uint32_t num_inputs = 0;
    uint64_t nvars = 0;
    bool have_witness = false;
    bool use_lagrange = false;   // wire commitments from VALUES with the Lagrange-form key (bellman prove, src/plonk.rs:138-146)
    // two gate types (the recursive prover's shape, src/recursive/mod.rs:111-127): main gate + Rescue x^5 custom gate
    bool gated = false;
    DevBuf<uint8_t> gate_type;   // [n]
    DevBuf<fr_t> gsel_coef;      // [2][n]  s_main, s_resc (monomial)
    DevBuf<fr_t> gsel_lde;       // [2][4n]
    DevBuf<uint32_t> wire_idx;   // [4][n]
    DevBuf<fr_t> sel_vals;       // [7][n] natural order (gate check)
    DevBuf<fr_t> sigma_vals;     // [4][n] natural order (grand product)
    DevBuf<fr_t> sel_coef;       // [7][n] monomial
    DevBuf<fr_t> sigma_coef;     // [4][n]
    DevBuf<fr_t> sel_lde;        // [7][4n] slot layout
    DevBuf<fr_t> sigma_lde;      // [4][4n]
    DevBuf<fr_t> vars;           // witness, Montgomery
    // per-proof working set (allocated once)
    DevBuf<fr_t> w_nat, w_br, w_coef;   // [4][n]
    DevBuf<fr_t> w_lde;                 // [4][4n]
    DevBuf<fr_t> z_coef, z_lde, pi_coef, pi_lde, t4;
    DevBuf<fr_t> tmp_a, tmp_b, tmp_c;   // [n]
    DevBuf<fr_t> zpow, zinvpow, zwpow, zwinvpow, r_coef;
};

namespace pk {

static fr_t fr_from_limbs32(const uint32_t v[8]) {
    fr_t x;
    for (int i = 0; i < 8; ++i) x.v[i] = v[i];
    return x.to_mont();
}
static void tr_commit_fr(RollingKeccakTranscript& tr, const fr_t& x) {
    fr_t c = x.from_mont();
    tr.commit_limbs(c.v);
}
static void tr_commit_g1(RollingKeccakTranscript& tr, const g1_affine_t& p) {
    if (p.is_inf()) {  // infinity is absorbed as (0, 0)
        uint32_t z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        tr.commit_limbs(z);
        tr.commit_limbs(z);
        return;
    }
    fq_t x = p.x.from_mont(), y = p.y.from_mont();
    tr.commit_limbs(x.v);
    tr.commit_limbs(y.v);
}
static fr_t tr_challenge(RollingKeccakTranscript& tr) {
    uint32_t c[8];
    tr.challenge(c);
    return fr_from_limbs32(c);
}
static void fr_to_abi(const fr_t& x, uint64_t out[4]) {
    fr_t c = x.from_mont();
    memcpy(out, c.v, 32);
}

// sigma as target positions: cycles over the positions of each variable, rows scanned in order,
// columns a..d inside a row; dummy variable 0 stays the identity.
std::vector<uint32_t> build_sigma_targets(const uint32_t* wire_idx, uint64_t n, uint64_t nvars) {
    const uint32_t NONE = 0xffffffffu;
    std::vector<uint32_t> target(4 * n), first(nvars, NONE), prev(nvars, NONE);
    for (uint64_t i = 0; i < 4 * n; ++i) target[i] = (uint32_t)i;
    for (uint64_t r = 0; r < n; ++r)
        for (int c = 0; c < 4; ++c) {
            uint32_t var = wire_idx[(uint64_t)c * n + r];
            if (var == 0) continue;
            PK_REQUIRE(var < nvars, PK_ERR_ASSIGNMENT_MISSING, "wire_idx refers to a variable beyond nvars");
            uint32_t pos = (uint32_t)((uint64_t)c * n + r);
            if (prev[var] != NONE) target[prev[var]] = pos;
            else first[var] = pos;
            prev[var] = pos;
        }
    for (uint64_t v = 1; v < nvars; ++v)
        if (prev[v] != NONE) target[prev[v]] = first[v];
    return target;
}

static void setup_poly(pk_ctx* ctx, const fr_t* vals_nat, fr_t* coef, fr_t* lde, fr_t* tmp, int log_n) {
    bitrev_permute(ctx, vals_nat, tmp, log_n);
    ntt_inverse_from_bitrev(ctx, tmp, coef, log_n);
    lde4_slots(ctx, coef, lde, log_n);
}

static void setup_create_impl(pk_ctx* ctx, const pk_assembly* as, const uint8_t* gate_type, pk_setup** out);
void setup_create(pk_ctx* ctx, const pk_assembly* as, pk_setup** out) { setup_create_impl(ctx, as, nullptr, out); }
void setup_create_gated(pk_ctx* ctx, const pk_assembly_gated* as, pk_setup** out) {
    PK_REQUIRE(as && as->gate_type, PK_ERR_INVALID, "null gate types");
    setup_create_impl(ctx, &as->base, as->gate_type, out);
}
static void setup_create_impl(pk_ctx* ctx, const pk_assembly* as, const uint8_t* gate_type, pk_setup** out) {
    PK_REQUIRE(as && out, PK_ERR_INVALID, "null argument");
    const uint64_t n = as->n;
    PK_REQUIRE(n >= 2 && (n & (n - 1)) == 0, PK_ERR_INVALID, "domain size must be a power of two >= 2");
    const int log_n = ilog2(n);
    PK_REQUIRE(log_n + 2 <= 28, PK_ERR_DEGREE_TOO_LARGE, "circuit larger than 2^26 gates (SETUP_MAX_POW2, src/plonk.rs:27)");
    PK_REQUIRE(ctx->srs && ctx->srs->n >= n, PK_ERR_DEGREE_TOO_LARGE, "SRS smaller than the circuit domain");
    PK_REQUIRE(as->num_inputs < n, PK_ERR_INVALID, "too many public inputs");
    PK_REQUIRE(as->nvars >= 1 && as->nvars < (uint64_t(1) << 32), PK_ERR_INVALID, "nvars out of range");
    std::vector<uint32_t> target = build_sigma_targets(as->wire_idx, n, as->nvars);

    pk_setup* s = new pk_setup();
    try {
        s->ctx = ctx; s->log_n = log_n; s->n = n; s->num_inputs = (uint32_t)as->num_inputs; s->nvars = as->nvars;
        cudaStream_t st = ctx->stream;
        s->wire_idx.alloc(4 * n);
        s->sel_vals.alloc(7 * n); s->sigma_vals.alloc(4 * n);
        s->sel_coef.alloc(7 * n); s->sigma_coef.alloc(4 * n);
        s->sel_lde.alloc(28 * n); s->sigma_lde.alloc(16 * n);
        s->vars.alloc(as->nvars);
        s->w_nat.alloc(4 * n); s->w_br.alloc(4 * n); s->w_coef.alloc(4 * n); s->w_lde.alloc(16 * n);
        s->z_coef.alloc(n); s->z_lde.alloc(4 * n); s->pi_coef.alloc(n); s->pi_lde.alloc(4 * n); s->t4.alloc(4 * n);
        s->tmp_a.alloc(n); s->tmp_b.alloc(n); s->tmp_c.alloc(n);
        s->zpow.alloc(n); s->zinvpow.alloc(n); s->zwpow.alloc(n); s->zwinvpow.alloc(n); s->r_coef.alloc(n);

        PK_CUDA(cudaMemcpyAsync(s->wire_idx.p, as->wire_idx, 4 * n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        PK_CUDA(cudaMemcpyAsync(s->sel_vals.p, as->selectors, 7 * n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
        fr_to_mont(ctx, s->sel_vals.p, 7 * n);
        DevBuf<uint32_t> d_target(4 * n);
        PK_CUDA(cudaMemcpyAsync(d_target.p, target.data(), 4 * n * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        sigma_values(ctx, d_target.p, s->sigma_vals.p, log_n);
        for (int k = 0; k < 7; ++k)
            setup_poly(ctx, s->sel_vals.p + k * n, s->sel_coef.p + k * n, s->sel_lde.p + 4 * k * n, s->tmp_a.p, log_n);
        for (int k = 0; k < 4; ++k)
            setup_poly(ctx, s->sigma_vals.p + k * n, s->sigma_coef.p + k * n, s->sigma_lde.p + 4 * k * n, s->tmp_a.p, log_n);
        if (gate_type) {
            // gate selectors: s_main = 1 on main-gate rows (padding included), s_resc = 1 on Rescue x^5 rows
            s->gated = true;
            s->gate_type.alloc(n); s->gsel_coef.alloc(2 * n); s->gsel_lde.alloc(8 * n);
            std::vector<fr_t> vals(2 * n, fr_t::zero());
            for (uint64_t r = 0; r < n; ++r) {
                PK_REQUIRE(gate_type[r] <= 1, PK_ERR_INVALID, "unknown gate type");
                vals[(gate_type[r] == 1 ? n : 0) + r] = fr_t::one();
            }
            PK_CUDA(cudaMemcpyAsync(s->gate_type.p, gate_type, n, cudaMemcpyHostToDevice, st));
            PK_CUDA(cudaMemcpyAsync(s->tmp_b.p, vals.data(), n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
            setup_poly(ctx, s->tmp_b.p, s->gsel_coef.p, s->gsel_lde.p, s->tmp_a.p, log_n);
            PK_CUDA(cudaMemcpyAsync(s->tmp_b.p, vals.data() + n, n * sizeof(fr_t), cudaMemcpyHostToDevice, st));
            setup_poly(ctx, s->tmp_b.p, s->gsel_coef.p + n, s->gsel_lde.p + 4 * n, s->tmp_a.p, log_n);
            PK_CUDA(cudaStreamSynchronize(st));  // `vals` is pageable host memory: done with it before it goes away
        }
        PK_CUDA(cudaStreamSynchronize(st));
        PK_CUDA(cudaGetLastError());
    } catch (...) {
        delete s;
        throw;
    }
    *out = s;
}

void setup_commitments(pk_ctx* ctx, pk_setup* s, uint64_t out_xy[11][8]) {
    const fr_t* polys[11];
    for (int k = 0; k < 7; ++k) polys[k] = s->sel_coef.p + k * s->n;
    for (int k = 0; k < 4; ++k) polys[7 + k] = s->sigma_coef.p + k * s->n;
    g1_affine_t out[11];
    msm_run_batch(ctx, polys, 11, s->n, 0, out);
    for (int k = 0; k < 11; ++k) affine_to_abi(out[k], out_xy[k]);
}

void setup_commitments_gated(pk_ctx* ctx, pk_setup* s, uint64_t out_xy[13][8]) {
    PK_REQUIRE(s->gated, PK_ERR_INVALID, "setup has a single gate type (use pk_setup_commitments)");
    const fr_t* polys[13];
    for (int k = 0; k < 7; ++k) polys[k] = s->sel_coef.p + k * s->n;
    polys[7] = s->gsel_coef.p; polys[8] = s->gsel_coef.p + s->n;
    for (int k = 0; k < 4; ++k) polys[9 + k] = s->sigma_coef.p + k * s->n;
    g1_affine_t out[13];
    msm_run_batch(ctx, polys, 13, s->n, 0, out);
    for (int k = 0; k < 13; ++k) affine_to_abi(out[k], out_xy[k]);
}

void witness_upload(pk_ctx* ctx, pk_setup* s, const uint64_t* var_values, uint64_t nvars) {
    PK_REQUIRE(var_values != nullptr, PK_ERR_ASSIGNMENT_MISSING, "witness is null");
    PK_REQUIRE(nvars == s->nvars, PK_ERR_ASSIGNMENT_MISSING, "witness length does not match the circuit");
    PK_CUDA(cudaMemcpyAsync(s->vars.p, var_values, nvars * sizeof(fr_t), cudaMemcpyHostToDevice, ctx->stream));
    fr_to_mont(ctx, s->vars.p, nvars);
    s->have_witness = true;
}

struct PhaseClock {
    pk_ctx* ctx;
    cudaEvent_t ev[9];
    int k = 0;
    explicit PhaseClock(pk_ctx* c) : ctx(c) { for (auto& e : ev) cudaEventCreate(&e); }
    ~PhaseClock() { for (auto& e : ev) cudaEventDestroy(e); }
    void mark() { if (k < 9) cudaEventRecord(ev[k++], ctx->stream); }
    void finish() {
        cudaEventSynchronize(ev[k - 1]);
        for (int i = 0; i + 1 < k && i < 7; ++i) {
            float ms = 0;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            ctx->prof.phase_ms[i] = ms;
        }
        float tot = 0;
        cudaEventElapsedTime(&tot, ev[0], ev[k - 1]);
        ctx->prof.phase_ms[7] = tot;
    }
};

void prove(pk_ctx* ctx, pk_setup* s, const uint64_t* var_values, uint64_t nvars, pk_proof* proof, uint64_t* inputs_out) {
    PK_REQUIRE(proof != nullptr, PK_ERR_INVALID, "null proof");
    PK_REQUIRE(ctx->srs && ctx->srs->n >= s->n, PK_ERR_DEGREE_TOO_LARGE, "SRS smaller than the circuit domain");
    const uint64_t n = s->n;
    const int log_n = s->log_n;
    const uint32_t ni = s->num_inputs;
    cudaStream_t st = ctx->stream;
    PhaseClock clk(ctx);
    clk.mark();
    if (var_values) witness_upload(ctx, s, var_values, nvars);
    PK_REQUIRE(s->have_witness, PK_ERR_ASSIGNMENT_MISSING, "no witness uploaded");
    clk.mark();  // phase 0: h2d

    // ---- witness -> wire values; is_satisfied_using_one_shot_check (src/plonk.rs:137)
    wire_gather(ctx, s->vars.p, s->wire_idx.p, s->w_nat.p, s->w_br.p, log_n);
    const bool sat = s->gated ? gate_check_gated(ctx, s->w_nat.p, s->sel_vals.p, s->gate_type.p, ni, log_n)
                              : gate_check(ctx, s->w_nat.p, s->sel_vals.p, ni, log_n);
    PK_REQUIRE(sat, PK_ERR_UNSATISFIED, "witness does not satisfy the circuit");
    std::vector<fr_t> inputs(ni);
    if (ni) {
        PK_CUDA(cudaMemcpyAsync(inputs.data(), s->w_nat.p, ni * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
        PK_CUDA(cudaStreamSynchronize(st));
    }
    RollingKeccakTranscript tr;
    for (uint32_t i = 0; i < ni; ++i) {
        tr_commit_fr(tr, inputs[i]);
        if (inputs_out) fr_to_abi(inputs[i], inputs_out + 4 * i);
    }

    // ---- round 1: wire polynomials and commitments
    g1_affine_t Cw[4];
    {
        const fr_t* polys[4];
        for (int c = 0; c < 4; ++c) {
            ntt_inverse_from_bitrev(ctx, s->w_br.p + c * n, s->w_coef.p + c * n, log_n);
            polys[c] = s->w_coef.p + c * n;
        }
        {
            // the wire LDEs of round 3 need only the coefficients: they run on the side stream beside the commitment
            // kernels (whose sort / reduction phases leave the multiplier pipe idle)
            SideStreamScope side(ctx);
            for (int c = 0; c < 4; ++c) lde4_slots(ctx, s->w_coef.p + c * n, s->w_lde.p + 4 * c * n, log_n);
        }
        if (s->use_lagrange) {
            // commit_using_values: sum_i w(omega^i) [L_i(tau)] G over the witness values themselves (natural order)
            PK_REQUIRE(ctx->srs_lagrange && ctx->srs_lagrange->n == n, PK_ERR_DEGREE_TOO_LARGE,
                       "no Lagrange-form key of the circuit's domain size is loaded");
            const fr_t* vals[4];
            for (int c = 0; c < 4; ++c) vals[c] = s->w_nat.p + c * n;
            msm_run_batch(ctx, vals, 4, n, 0, Cw, ctx->srs_lagrange);
        } else {
            msm_run_batch(ctx, polys, 4, n, 0, Cw);  // the 4 wire commitments share one pass over the MSM kernels
        }
        for (int c = 0; c < 4; ++c) tr_commit_g1(tr, Cw[c]);
    }
    const fr_t beta = tr_challenge(tr), gamma = tr_challenge(tr);
    clk.mark();  // phase 1

    // ---- round 2: grand product Z
    perm_num_den(ctx, s->w_nat.p, s->sigma_vals.p, beta, gamma, s->tmp_a.p, s->tmp_b.p, log_n);
    poly_scan(ctx, true, false, s->tmp_a.p, s->tmp_a.p, n);  // pn[j] = prod_{i<=j} num_i
    poly_scan(ctx, true, true, s->tmp_b.p, s->tmp_b.p, n);   // sd[j] = prod_{i>=j} den_i
    fr_t total_den;
    PK_CUDA(cudaMemcpyAsync(&total_den, s->tmp_b.p, sizeof(fr_t), cudaMemcpyDeviceToHost, st));
    PK_CUDA(cudaStreamSynchronize(st));
    PK_REQUIRE(!total_den.is_zero(), PK_ERR_DIVISION_BY_ZERO, "zero denominator in the permutation grand product");
    z_finish(ctx, s->tmp_a.p, s->tmp_b.p, total_den.inverse(), s->tmp_c.p, log_n);
    ntt_inverse_from_bitrev(ctx, s->tmp_c.p, s->z_coef.p, log_n);
    {
        SideStreamScope side(ctx);
        lde4_slots(ctx, s->z_coef.p, s->z_lde.p, log_n);
    }
    const g1_affine_t Cz = msm_run(ctx, s->z_coef.p, n, 0);
    tr_commit_g1(tr, Cz);
    const fr_t alpha = tr_challenge(tr);
    clk.mark();  // phase 2

    // ---- round 3: quotient on the coset 7*H_4n (the LDEs of the wires and of Z were started beside rounds 1 and 2)
    side_join(ctx);
    CosetTables* ct = get_coset_tables(ctx, log_n);
    QuotientArgs qa;
    if (ni <= 8 && !s->gated) {
        // PI(X) = sum_i in_i L_0(X w^-i): read off the resident L_0 table inside the quotient kernel, no NTT
        qa.num_direct_inputs = (int)ni;
        for (uint32_t i = 0; i < ni; ++i) qa.inputs[i] = inputs[i];
    } else {
        qa.num_direct_inputs = -1;
        PK_CUDA(cudaMemsetAsync(s->tmp_c.p, 0, n * sizeof(fr_t), st));
        pi_scatter(ctx, s->w_nat.p, s->tmp_c.p, ni, log_n);
        ntt_inverse_from_bitrev(ctx, s->tmp_c.p, s->pi_coef.p, log_n);
        lde4_slots(ctx, s->pi_coef.p, s->pi_lde.p, log_n);
    }
    for (int c = 0; c < 4; ++c) { qa.w[c] = s->w_lde.p + 4 * c * n; qa.sig[c] = s->sigma_lde.p + 4 * c * n; }
    for (int k = 0; k < 7; ++k) qa.sel[k] = s->sel_lde.p + 4 * k * n;
    qa.z = s->z_lde.p; qa.pi = s->pi_lde.p; qa.l0 = ct->l0.p; qa.out = s->t4.p;
    qa.beta = beta; qa.gamma = gamma; qa.alpha = alpha; qa.log_n = log_n;
    if (s->gated) { qa.gsel[0] = s->gsel_lde.p; qa.gsel[1] = s->gsel_lde.p + 4 * n; }
    quotient_slots(ctx, qa);
    icoset4n_from_slots(ctx, s->t4.p, s->t4.p, log_n);
    {
        // deg t <= 4n - 5: the top coefficients must vanish, otherwise the numerator was not divisible by Z_H
        fr_t top[3];
        PK_CUDA(cudaMemcpyAsync(top, s->t4.p + 4 * n - 3, 3 * sizeof(fr_t), cudaMemcpyDeviceToHost, st));
        PK_CUDA(cudaStreamSynchronize(st));
        PK_REQUIRE(top[0].is_zero() && top[1].is_zero() && top[2].is_zero(), PK_ERR_UNSATISFIED, "quotient is not a polynomial");
    }
    g1_affine_t Ct[4];
    {
        const fr_t* polys[4] = {s->t4.p, s->t4.p + n, s->t4.p + 2 * n, s->t4.p + 3 * n};
        msm_run_batch(ctx, polys, 4, n, 0, Ct);
        for (int i = 0; i < 4; ++i) tr_commit_g1(tr, Ct[i]);
    }
    const fr_t zeta = tr_challenge(tr);
    clk.mark();  // phase 3

    // ---- round 4: evaluations and linearisation
    const fr_t omega = host_root_of_unity(log_n);
    const fr_t zeta_omega = zeta * omega;
    poly_powers(ctx, s->zpow.p, zeta, n);
    poly_powers(ctx, s->zwpow.p, zeta_omega, n);
    fr_t ev[15];
    {
        const fr_t* polys[15];
        const fr_t* pows[15];
        for (int c = 0; c < 4; ++c) polys[c] = s->w_coef.p + c * n;
        for (int c = 0; c < 3; ++c) polys[4 + c] = s->sigma_coef.p + c * n;
        for (int i = 0; i < 4; ++i) polys[7 + i] = s->t4.p + i * n;
        for (int k = 0; k < 11; ++k) pows[k] = s->zpow.p;
        polys[11] = s->w_coef.p + 3 * n; pows[11] = s->zwpow.p;
        polys[12] = s->z_coef.p; pows[12] = s->zwpow.p;
        if (s->gated) {
            polys[13] = s->gsel_coef.p; polys[14] = s->gsel_coef.p + n;
            pows[13] = pows[14] = s->zpow.p;
        }
        poly_dot_batch(ctx, s->gated ? 15 : 13, polys, pows, n, ev);
    }
    const fr_t smz = s->gated ? ev[13] : fr_t::one(), srz = s->gated ? ev[14] : fr_t::zero();
    // powers of alpha of the copy-permutation and L_0 terms: (alpha, alpha^2), or (alpha^4, alpha^5) behind the custom gate's three
    const fr_t a_perm = s->gated ? alpha.sqr().sqr() : alpha;
    const fr_t a_l0 = s->gated ? a_perm * alpha : alpha.sqr();
    const fr_t* wz = ev;          // a,b,c,d at zeta
    const fr_t* sz = ev + 4;      // sigma_0..2 at zeta
    const fr_t dzw = ev[11], zzw = ev[12];
    const fr_t zeta_n = zeta.pow_u64(n);
    const fr_t zn2 = zeta_n.sqr(), zn3 = zn2 * zeta_n;
    const fr_t tz = ev[7] + zeta_n * ev[8] + zn2 * ev[9] + zn3 * ev[10];
    PK_REQUIRE(!(zeta - fr_t::one()).is_zero(), PK_ERR_DIVISION_BY_ZERO, "challenge z hit the domain");
    const fr_t n_fr = fr_t::from_u32(2).pow_u64(log_n);
    const fr_t l0z = (zeta_n - fr_t::one()) * (n_fr * (zeta - fr_t::one())).inverse();
    static const uint32_t KK[4] = {1, 5, 7, 10};
    fr_t zfac = a_perm;
    for (int i = 0; i < 4; ++i) zfac = zfac * (wz[i] + beta * fr_t::from_u32(KK[i]) * zeta + gamma);
    zfac = zfac + a_l0 * l0z;
    fr_t sfac = a_perm * beta * zzw;
    for (int i = 0; i < 3; ++i) sfac = sfac * (wz[i] + beta * sz[i] + gamma);
    {
        // r(X) = s_main(z) [q_const + sum q_i w_i(z) + q_m a(z) b(z) + q_dnext d(z w)] + Z(X) zfac - sigma_3(X) sfac
        // (s_main(z) = 1 without gate selectors)
        const fr_t* in[9] = {s->sel_coef.p + 5 * n, s->sel_coef.p, s->sel_coef.p + n, s->sel_coef.p + 2 * n, s->sel_coef.p + 3 * n,
                             s->sel_coef.p + 4 * n, s->sel_coef.p + 6 * n, s->z_coef.p, s->sigma_coef.p + 3 * n};
        fr_t coef[9] = {smz, smz * wz[0], smz * wz[1], smz * wz[2], smz * wz[3], smz * wz[0] * wz[1], smz * dzw, zfac, sfac.neg()};
        poly_lincomb(ctx, s->r_coef.p, 9, in, coef, n);
    }
    fr_t rz;
    {
        const fr_t* polys[1] = {s->r_coef.p};
        const fr_t* pows[1] = {s->zpow.p};
        poly_dot_batch(ctx, 1, polys, pows, n, &rz);
    }
    for (int c = 0; c < 4; ++c) tr_commit_fr(tr, wz[c]);
    tr_commit_fr(tr, dzw);
    if (s->gated) { tr_commit_fr(tr, smz); tr_commit_fr(tr, srz); }
    for (int c = 0; c < 3; ++c) tr_commit_fr(tr, sz[c]);
    tr_commit_fr(tr, tz);
    tr_commit_fr(tr, rz);
    tr_commit_fr(tr, zzw);
    const fr_t v = tr_challenge(tr);
    clk.mark();  // phase 4

    // ---- round 5: opening proofs
    fr_t vp[13];
    vp[0] = fr_t::one();
    for (int i = 1; i <= 12; ++i) vp[i] = vp[i - 1] * v;
    if (!s->gated) {
        const fr_t* in[12] = {s->t4.p, s->t4.p + n, s->t4.p + 2 * n, s->t4.p + 3 * n, s->r_coef.p, s->w_coef.p, s->w_coef.p + n,
                              s->w_coef.p + 2 * n, s->w_coef.p + 3 * n, s->sigma_coef.p, s->sigma_coef.p + n, s->sigma_coef.p + 2 * n};
        fr_t coef[12] = {fr_t::one(), zeta_n, zn2, zn3, vp[1], vp[2], vp[3], vp[4], vp[5], vp[6], vp[7], vp[8]};
        poly_lincomb(ctx, s->tmp_a.p, 12, in, coef, n);
    } else {  // the two gate selectors are opened at z as well: v^6, v^7 between the wires and the sigmas
        const fr_t* in[14] = {s->t4.p, s->t4.p + n, s->t4.p + 2 * n, s->t4.p + 3 * n, s->r_coef.p, s->w_coef.p, s->w_coef.p + n,
                              s->w_coef.p + 2 * n, s->w_coef.p + 3 * n, s->gsel_coef.p, s->gsel_coef.p + n, s->sigma_coef.p,
                              s->sigma_coef.p + n, s->sigma_coef.p + 2 * n};
        fr_t coef[14] = {fr_t::one(), zeta_n, zn2, zn3, vp[1], vp[2], vp[3], vp[4], vp[5], vp[6], vp[7], vp[8], vp[9], vp[10]};
        poly_lincomb(ctx, s->tmp_a.p, 14, in, coef, n);
    }
    {
        const fr_t* in[2] = {s->z_coef.p, s->w_coef.p + 3 * n};
        fr_t coef[2] = {s->gated ? vp[11] : vp[9], s->gated ? vp[12] : vp[10]};
        poly_lincomb(ctx, s->tmp_b.p, 2, in, coef, n);
    }
    PK_REQUIRE(!zeta.is_zero(), PK_ERR_DIVISION_BY_ZERO, "challenge z is zero");
    poly_powers(ctx, s->zinvpow.p, zeta.inverse(), n);
    poly_powers(ctx, s->zwinvpow.p, zeta_omega.inverse(), n);
    // W_z = (agg(X) - agg(z)) / (X - z);  W_zw = (agg2(X) - agg2(z w)) / (X - z w)
    poly_divide_linear(ctx, s->tmp_a.p, s->zpow.p, s->zinvpow.p, s->r_coef.p, s->tmp_c.p, n);
    poly_divide_linear(ctx, s->tmp_b.p, s->zwpow.p, s->zwinvpow.p, s->tmp_a.p, s->tmp_c.p, n);
    g1_affine_t Wz[2];
    {
        const fr_t* polys[2] = {s->r_coef.p, s->tmp_a.p};
        msm_run_batch(ctx, polys, 2, n, 0, Wz);
    }
    const g1_affine_t W1 = Wz[0], W2 = Wz[1];
    clk.mark();  // phase 5
    clk.finish();

    // ---- Proof
    memset(proof, 0, sizeof(*proof));
    proof->n = n - 1;
    proof->num_inputs = ni;
    for (int c = 0; c < 4; ++c) {
        affine_to_abi(Cw[c], proof->wire_commitments[c]);
        affine_to_abi(Ct[c], proof->quotient_poly_commitments[c]);
        fr_to_abi(wz[c], proof->wire_values_at_z[c]);
    }
    affine_to_abi(Cz, proof->grand_product_commitment);
    fr_to_abi(dzw, proof->wire_values_at_z_omega[0]);
    fr_to_abi(zzw, proof->grand_product_at_z_omega);
    fr_to_abi(tz, proof->quotient_polynomial_at_z);
    fr_to_abi(rz, proof->linearization_polynomial_at_z);
    for (int c = 0; c < 3; ++c) fr_to_abi(sz[c], proof->permutation_polynomials_at_z[c]);
    affine_to_abi(W1, proof->opening_at_z_proof);
    affine_to_abi(W2, proof->opening_at_z_omega_proof);
    fr_to_abi(beta, proof->challenges[0]);
    fr_to_abi(gamma, proof->challenges[1]);
    fr_to_abi(alpha, proof->challenges[2]);
    fr_to_abi(zeta, proof->challenges[3]);
    fr_to_abi(v, proof->challenges[4]);
    if (s->gated) {
        proof->num_gate_selectors = 2;
        fr_to_abi(smz, proof->gate_selectors_at_z[0]);
        fr_to_abi(srz, proof->gate_selectors_at_z[1]);
    }
}

}  // namespace pk

namespace pk {
void setup_use_lagrange(pk_ctx* ctx, pk_setup* s, bool on) {
    if (on)
        PK_REQUIRE(ctx->srs_lagrange && ctx->srs_lagrange->n == s->n, PK_ERR_DEGREE_TOO_LARGE,
                   "no Lagrange-form key of the circuit's domain size is loaded (pk_srs_load_g1_lagrange)");
    s->use_lagrange = on;
}
void setup_free(pk_setup* s) {
    if (!s) return;
    if (s->ctx) {
        cudaSetDevice(s->ctx->device);
        cudaStreamSynchronize(s->ctx->stream);
        if (s->ctx->side) cudaStreamSynchronize(s->ctx->side);  // LDEs of an aborted proof may still be in flight there
    }
    delete s;
}
}  // namespace pk
