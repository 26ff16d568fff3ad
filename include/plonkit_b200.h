/* plonkit_b200 — C ABI of the B200-native prove path for fluidex/plonkit.
 *
 * The reference has no FFI: it links bellman_ce (Cargo.lock:109-111) statically.  This header is the boundary a
 * Rust shim would bind (INTEGRATION.md shows the `extern "C"` block); each entry point names the bellman
 * primitive / plonkit call site it stands in for.  All calls are blocking, return a status code (never throw
 * or abort across the boundary), take caller-owned host buffers, and keep device memory behind the opaque
 * context.  One proof at a time per context (like the reference's one Worker per call, src/plonk.rs:41,47,183).
 *
 * Data layout across the boundary
 *   Fr / Fq element : uint64_t[4], little-endian limbs.  `fmt` selects PK_FMT_CANONICAL (what bellman's
 *                     `PrimeFieldRepr` / file formats hold) or PK_FMT_MONTGOMERY (the in-memory `Fr`, R = 2^256).
 *   G1 affine point : uint64_t[8] = x || y, canonical limbs; (0, 0) = point at infinity.
 */
#ifndef PLONKIT_B200_H
#define PLONKIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pk_ctx pk_ctx;
typedef struct pk_setup pk_setup;
typedef struct pk_dist_setup pk_dist_setup;
typedef struct pk_comm_group pk_comm_group;

/* status codes; the shim maps them onto bellman's SynthesisError variants */
enum {
    PK_OK = 0,
    PK_ERR_ASSIGNMENT_MISSING = 1,   /* length mismatch / missing witness   (SynthesisError::AssignmentMissing)        */
    PK_ERR_DEGREE_TOO_LARGE = 2,     /* domain > 2^28 or SRS too small       (SynthesisError::PolynomialDegreeTooLarge) */
    PK_ERR_DIVISION_BY_ZERO = 3,     /*                                       (SynthesisError::DivisionByZero)           */
    PK_ERR_UNSATISFIED = 4,          /* gate identity fails on the witness   (SynthesisError::Unsatisfiable)            */
    PK_ERR_CUDA = 5,                 /* no device / CUDA runtime error — there is no CPU fallback                         */
    PK_ERR_INVALID = 6               /* bad argument                                                                       */
};

enum { PK_FMT_CANONICAL = 0, PK_FMT_MONTGOMERY = 1 };

/* ---- context ------------------------------------------------------------------------------------------- */
/* Creates a context on CUDA device `device`.  Stands in for Worker::new() (src/plonk.rs:41,47,183). */
int pk_create(int device, pk_ctx** out);
void pk_destroy(pk_ctx* ctx);
const char* pk_last_error(const pk_ctx* ctx);
/* Montgomery constants the device code was generated with: out[0..3]=R mod r, [4..7]=R^2 mod r, [8]=-r^-1 mod 2^32,
 * [12..15]=R mod q, [16..19]=R^2 mod q, [20]=-q^-1 mod 2^32.  No device needed. */
void pk_constants(uint64_t out[24]);

/* ---- SRS ----------------------------------------------------------------------------------------------- */
/* Makes the first n G1 bases of a Crs resident (Crs::read result, src/reader.rs:74-77) and builds the
 * fixed-base window tables the MSM uses.  window_bits = 0 picks it from n. */
int pk_srs_load_g1(pk_ctx* ctx, const uint64_t* bases_xy, uint64_t n, int window_bits);
/* The Lagrange-form key of a size-n domain (Crs<Bn256, CrsForLagrangeForm>, src/reader.rs:80-89; n a power of two):
 * kept resident beside the monomial one.  A setup switched to it with pk_setup_use_lagrange commits the four wire
 * polynomials from their VALUES (kate_commitment::commit_using_values, the `prove(..., key_lagrange_form)` branch of
 * src/plonk.rs:138-146) — same proof bytes, no dependence of the first round on the inverse NTTs.  bases_xy = NULL with
 * n = 0 unloads it. */
int pk_srs_load_g1_lagrange(pk_ctx* ctx, const uint64_t* bases_xy, uint64_t n, int window_bits);
/* Crs::<Bn256, CrsForMonomialForm>::crs_42(n) generalised to any tau (src/plonk.rs:41,47): out[i] = [tau^i] G. */
int pk_srs_gen(pk_ctx* ctx, uint64_t n, uint64_t tau, uint64_t* out_xy);

/* ---- bellman primitives (a8-a11 of SURVEY.md section 8) ----------------------------------------------- */
/* Polynomial::{fft, ifft, coset_fft, icoset_fft_for_generator}: radix-2 (i)NTT over Fr on the size-2^log_n
 * subgroup, natural order in and out, in place.  coset != 0 uses the coset 7*H (bellman's multiplicative generator). */
int pk_ntt(pk_ctx* ctx, uint64_t* fr, uint32_t log_n, int inverse, int coset, int fmt);
/* Polynomial::bitreversed_lde_using_bitreversed_ntt(factor = 4, coset_factor = 7): n coefficients -> 4n evaluations
 * on 7*H_4n.  bitreversed = 0: natural order; 1: the 4n-point bit-reversed order bellman's LDE returns. */
int pk_lde4(pk_ctx* ctx, const uint64_t* coeffs, uint32_t log_n, uint64_t* out_4n, int bitreversed, int fmt);
/* ---- polynomial primitives (a12, a14 of SURVEY.md section 8): canonical limbs in and out ------------------------------
 * Polynomial<Fr, Coefficients>::evaluate_at(worker, z) -> sum_i coeffs[i] z^i  (call sites behind src/plonk.rs:140,152). */
int pk_poly_evaluate_at(pk_ctx* ctx, const uint64_t* coeffs, uint64_t n, const uint64_t z[4], uint64_t out[4]);
/* kate_commitment::divide_single(poly, z): the n coefficients of (p(X) - p(z)) / (X - z), top coefficient 0. */
int pk_poly_divide_by_linear(pk_ctx* ctx, const uint64_t* coeffs, uint64_t n, const uint64_t z[4], uint64_t* quotient);
/* Polynomial<Fr, Values>::calculate_shifted_grand_product: out[0] = 1, out[i] = out[i-1] * values[i-1]. */
int pk_poly_shifted_grand_product(pk_ctx* ctx, const uint64_t* values, uint64_t n, uint64_t* out);
/* Polynomial<Fr, Values>::batch_inversion: values[i] <- values[i]^-1 in place, zeros stay zero. */
int pk_poly_batch_inversion(pk_ctx* ctx, uint64_t* values, uint64_t n);
/* Pointwise operations of bellman's Polynomial (row a13): op 0 add_assign_scaled (out = a + s b), 1 mul_assign (a .* b),
 * 2 scale (s a), 3 add_constant (a + s), 4 distribute_powers (out[i] = a[i] s^i).  b / scalar may be NULL where unused. */
int pk_poly_pointwise(pk_ctx* ctx, int op, const uint64_t* a, const uint64_t* b, const uint64_t scalar[4], uint64_t n, uint64_t* out);
/* multiexp::dense_multiexp / commit_using_monomials: sum_i scalars[i] * SRS[base_offset + i], normalised to affine. */
int pk_msm_g1(pk_ctx* ctx, const uint64_t* scalars, uint64_t n, uint64_t base_offset, uint64_t out_xy[8], int* is_infinity,
              int fmt);
/* Sum of n affine points (host arithmetic, no device needed).  This is the local fold of the multi-GPU MSM: every
 * rank computes the partial sum over its chunk of (scalar, base) pairs, the 64-byte partials are all-gathered, and
 * each rank adds them up — the "all-reduce" of SURVEY.md section 8(e) (EC addition is not an NCCL reduce-op). */
int pk_g1_sum(const uint64_t* points_xy, uint64_t n, uint64_t out_xy[8]);
/* Crs::<_, CrsForLagrangeForm>::from_powers (src/plonk.rs:179-185): EC inverse FFT of the first 2^log_n resident
 * monomial bases: out[i] = [L_i(tau)] G, natural order. */
int pk_ec_intt_g1(pk_ctx* ctx, uint32_t log_n, uint64_t* out_xy);

/* ---- device-pointer primitives for the multi-GPU path ------------------------------ */
/* These operate in place on CALLER-OWNED device memory (e.g. torch tensors that NCCL collectives also touch), hold
 * Montgomery-form elements, and return after the library's stream has drained.
 * pk_dev_fr_convert : canonical <-> Montgomery over n elements.
 * pk_dev_ntt_rows   : `rows` independent (i)NTTs of length 2^log_len over consecutive rows, natural order in and out.
 * pk_dev_twiddle    : a[r][c] *= w_N^(+-(row0 + r) * c), N = 2^log_total — the twiddle step between the local
 *                     column transforms and the all-to-all of a four-step distributed NTT. */
int pk_dev_fr_convert(pk_ctx* ctx, void* dev, uint64_t n, int to_mont);
int pk_dev_ntt_rows(pk_ctx* ctx, void* dev, uint32_t log_len, uint64_t rows, int inverse);
int pk_dev_twiddle(pk_ctx* ctx, void* dev, uint64_t rows, uint64_t cols, uint32_t log_total, uint64_t row0, int inverse);
/* The same local steps over G1 POINTS, for Crs::from_powers (src/plonk.rs:179-185) of a key too large / too slow for one
 * GPU: device arrays of 64-byte affine points (canonical limbs, (0,0) = infinity) and of 128-byte XYZZ accumulators.
 * ntt_rows is UNSCALED (no 1/len).  pk_dev_ec_twiddle: inverse = 0 forward twiddles, 1 inverse twiddles, 2 inverse twiddles
 * times 2^-log_total (the scale of the whole inverse transform, free on a step that multiplies every point anyway);
 * pk_dev_ec_to_affine multiplies by 2^-log_scale while normalising (log_scale = 0: normalise only). */
int pk_dev_ec_from_affine(pk_ctx* ctx, const void* dev_affine, void* dev_xyzz, uint64_t n);
int pk_dev_ec_ntt_rows(pk_ctx* ctx, void* dev_xyzz, uint32_t log_len, uint64_t rows, int inverse);
int pk_dev_ec_twiddle(pk_ctx* ctx, void* dev_xyzz, uint64_t rows, uint64_t cols, uint32_t log_total, uint64_t row0, int inverse);
int pk_dev_ec_to_affine(pk_ctx* ctx, const void* dev_xyzz, void* dev_affine, uint64_t n, uint32_t log_scale);

/* ---- prover -------------------------------------------------------------------------------------------- */
/* Gate tables of a width-4 circuit with d_next (PlonkCsWidth4WithNextStepParams).  n is the domain size (power of
 * two); rows 0..num_inputs-1 are the public-input gates; row n-1 is never a gate.  Variable 0 is the dummy
 * variable (value 0).  SURVEY.md App. A.2. */
typedef struct pk_assembly {
    uint64_t n;
    uint64_t num_inputs;
    uint64_t nvars;
    const uint32_t* wire_idx;   /* [4][n] variable id per (column a..d, row) */
    const uint64_t* selectors;  /* [7][n][4] canonical: q_a, q_b, q_c, q_d, q_m, q_const, q_dnext */
} pk_assembly;

/* Proof<Bn256, PlonkCsWidth4WithNextStepParams>, field order of contrib/template.sol:330-344; canonical limbs. */
typedef struct pk_proof {
    uint64_t n;          /* number of gates = domain size - 1 */
    uint64_t num_inputs;
    uint64_t wire_commitments[4][8];
    uint64_t grand_product_commitment[8];
    uint64_t quotient_poly_commitments[4][8];
    uint64_t wire_values_at_z[4][4];
    uint64_t wire_values_at_z_omega[1][4];
    uint64_t grand_product_at_z_omega[4];
    uint64_t quotient_polynomial_at_z[4];
    uint64_t linearization_polynomial_at_z[4];
    uint64_t permutation_polynomials_at_z[3][4];
    uint64_t opening_at_z_proof[8];
    uint64_t opening_at_z_omega_proof[8];
    uint64_t challenges[5][4]; /* beta, gamma, alpha, z, v — not part of proof.bin; exposed for known-answer tests */
    /* only for setups made with pk_setup_create_gated (two gate types): the gate selectors s_main, s_resc at z */
    uint64_t num_gate_selectors;
    uint64_t gate_selectors_at_z[2][4];
} pk_proof;

/* Gate tables of the recursive prover's shape (src/recursive/mod.rs:111-127: a ProvingAssembly over
 * Width4MainGateWithDNext + the Rescue x^5 custom gate, bellman better_better_cs): the same tables plus a gate type per
 * row — 0: main gate (selectors apply), 1: Rescue x^5 gate (a = x, b = x^2, c = x^4, d = x^5; selectors ignored).
 * BYTE PARITY UNPINNED: that prover and its proof layout are not in the reference tree; DESIGN.md section 9 states the
 * protocol this library runs for it. */
typedef struct pk_assembly_gated {
    pk_assembly base;
    const uint8_t* gate_type;   /* [n] */
} pk_assembly_gated;

/* SetupForProver::prepare_setup_for_prover (src/plonk.rs:97-119) after synthesis: uploads the gate tables, builds the
 * copy permutation, the 11 setup polynomials (iNTT) and — unlike the reference, which recomputes them in every
 * prove call (precomputations = None, src/plonk.rs:156) — keeps their 4n coset evaluations resident.
 * Requires pk_srs_load_g1 with at least n bases. */
int pk_setup_create(pk_ctx* ctx, const pk_assembly* assembly, pk_setup** out);
/* The setup of the recursive prover's proving call (create_recursive_circuit_setup, src/recursive/mod.rs:120-121): 13
 * setup polynomials (7 main-gate, the gate selectors s_main and s_resc, 4 sigma).  pk_prove on such a setup runs the
 * two-gate-type protocol (create_proof, :127) and fills gate_selectors_at_z; pk_setup_commitments_gated returns the 13
 * commitments in that order. */
int pk_setup_create_gated(pk_ctx* ctx, const pk_assembly_gated* assembly, pk_setup** out);
int pk_setup_commitments_gated(pk_ctx* ctx, pk_setup* setup, uint64_t out_xy[13][8]);
void pk_setup_destroy(pk_setup* setup);
/* on != 0: pk_prove commits the wires from values with the resident Lagrange-form key (which must have exactly the
 * circuit's domain size); PK_ERR_DEGREE_TOO_LARGE if none is loaded. */
int pk_setup_use_lagrange(pk_ctx* ctx, pk_setup* setup, int on);
/* SetupForProver::make_verification_key (src/plonk.rs:122-124): 11 commitments, order q_a,q_b,q_c,q_d,q_m,q_const,
 * q_dnext, sigma_0..sigma_3. */
int pk_setup_commitments(pk_ctx* ctx, pk_setup* setup, uint64_t out_xy[11][8]);
/* Copies the witness (variable values, canonical, [nvars][4]; entry 0 must be 0) to the device. */
int pk_witness_upload(pk_ctx* ctx, pk_setup* setup, const uint64_t* var_values, uint64_t nvars);
/* SetupForProver::prove(circuit, "keccak") with the monomial-form SRS (src/plonk.rs:132-176): input values of the
 * public inputs are written to inputs_out[num_inputs][4].  var_values may be NULL to reuse the uploaded witness. */
int pk_prove(pk_ctx* ctx, pk_setup* setup, const uint64_t* var_values, uint64_t nvars, pk_proof* proof, uint64_t* inputs_out);

/* ---- one proof sharded over the GPUs of a node (BASELINE.json configs[2]) ---------------- */
/* The reference runs one Worker (host threads) per call; here ONE SetupForProver::prove is cut across 1, 2, 4 or 8
 * ranks, one context per GPU: commitments by base chunk (every rank loads ITS chunk of the key with pk_srs_load_g1:
 * bases [rank n/world, (rank+1) n/world)), the quotient by coset, one all-to-all inside the size-4n inverse NTT.  The
 * calls below are COLLECTIVE: every rank of the communicator makes the same call with the same circuit / witness and
 * every rank receives the same proof (bytes identical to pk_prove's).
 *
 * Transports: pk_comm_attach_nccl — one process per GPU (torchrun); rank 0 makes the id with pk_comm_nccl_unique_id and
 * the launcher broadcasts it.  pk_comm_attach_group — the ranks are threads of one process (contexts on one or several
 * devices; direct peer copies).  A rank that fails aborts the communicator: its peers return an error from their next
 * collective instead of hanging. */
int pk_comm_group_create(int world, pk_comm_group** out);
void pk_comm_group_destroy(pk_comm_group* group);
int pk_comm_attach_group(pk_ctx* ctx, pk_comm_group* group, int rank);
int pk_comm_nccl_unique_id(uint8_t out[128]);
int pk_comm_attach_nccl(pk_ctx* ctx, const uint8_t unique_id[128], int rank, int world);
/* prepare_setup_for_prover / make_verification_key / prove, sharded (same arguments as the pk_setup_* / pk_prove calls) */
int pk_dist_setup_create(pk_ctx* ctx, const pk_assembly* assembly, pk_dist_setup** out);
void pk_dist_setup_destroy(pk_dist_setup* setup);
int pk_dist_setup_commitments(pk_ctx* ctx, pk_dist_setup* setup, uint64_t out_xy[11][8]);
int pk_dist_witness_upload(pk_ctx* ctx, pk_dist_setup* setup, const uint64_t* var_values, uint64_t nvars);
int pk_dist_prove(pk_ctx* ctx, pk_dist_setup* setup, const uint64_t* var_values, uint64_t nvars, pk_proof* proof,
                  uint64_t* inputs_out);

/* ---- measurement hooks --------------------------------------------------------------------------------- */
typedef struct pk_profile {
    uint64_t kernel_launches;       /* kernels launched by this library since the last reset */
    uint64_t msm_accum_launches;    /* launches of the bucket-accumulation kernel (the dominant kernel) */
    double msm_accum_ms;            /* their summed CUDA-event time (only while profiling is on) */
    uint64_t msm_accum_points;      /* (scalar, base) pairs those launches covered */
    uint64_t ntt_launches;
    double ntt_ms;
    uint64_t ntt_elements;          /* elements x passes */
    double phase_ms[8];             /* last pk_prove: round1..round5, setup-dependent, h2d, total (CUDA events) */
    /* last pk_dist_prove: the three bulk collectives (CUDA events on the library stream; while profiling is on) */
    double comm_ms[3];              /* [0] all-gather of wire / Z coefficients, [1] all-to-all of the quotient, [2] all-gather of its coefficients */
    uint64_t comm_bytes[3];         /* bytes this rank RECEIVES from its peers in each of them */
} pk_profile;
void pk_profile_enable(pk_ctx* ctx, int on);
void pk_profile_reset(pk_ctx* ctx);
void pk_profile_get(const pk_ctx* ctx, pk_profile* out);

/* CUDA-event stopwatch on the library's stream: pk_timer_begin records an event, pk_timer_end records a second one,
 * waits for it and returns the elapsed device time in milliseconds. */
int pk_timer_begin(pk_ctx* ctx);
int pk_timer_end(pk_ctx* ctx, double* ms);

/* device-resident micro-benchmarks (inputs generated on the device): return elapsed ms by CUDA events */
int pk_bench_ntt(pk_ctx* ctx, uint32_t log_n, int iters, double* ms_per_iter);
int pk_bench_msm(pk_ctx* ctx, uint64_t n, int iters, double* ms_per_iter);
int pk_bench_fieldmul(pk_ctx* ctx, int which /*0 Fr, 1 Fq*/, double* gmuls_per_s);
/* device-resident MSM timing with a scalar pattern: 0 pseudo-random, 1 witness-like (40 % zero, 10 % one, 50 % random:
 * what the Lagrange path feeds), 2 all ones */
int pk_bench_msm_pattern(pk_ctx* ctx, uint64_t n, int pattern, int iters, double* ms_per_iter);

#ifdef __cplusplus
}
#endif
#endif /* PLONKIT_B200_H */
