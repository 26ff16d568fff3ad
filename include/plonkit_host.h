/* plonkit_host.h — plain C++ host helpers of the prove path (plonkit_b200/libplonkit_host.so; no CUDA, loads anywhere).
 *
 * Witness assignment: the values of the variables that the R1CS -> width-4 transpilation introduces, computed from a
 * circom witness.  Replaces the value closures bellman evaluates when `SetupForProver::prove` re-synthesises the
 * transpiled circuit on every call (src/plonk.rs:132-176; witness values from src/circom_circuit.rs:75-133): the
 * transpilation runs once and leaves a straight-line program, replayed here per proof.
 * All field elements are BN254 Fr as 4 x uint64 little-endian limbs, canonical (non-Montgomery) unless said otherwise. */
#ifndef PLONKIT_HOST_H
#define PLONKIT_HOST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* values[num_direct + k] = consts[k] + sum_{j in [off[k], off[k+1])} term_coef[j] * values[term_var[j]],  k < num_new,
 * with term_var[j] < num_direct + k.  The caller fills values[0 .. num_direct); coefficients are canonical, or in
 * Montgomery form when coef_is_mont != 0 (ph_fr_to_mont).  threads > 1 splits the program where no later entry reads a
 * variable introduced before the cut.  Returns 0, or 1 + the index of the first malformed variable (an operand outside
 * the field, a term reading a variable that is not assigned yet). */
int64_t ph_assign_witness(uint64_t num_direct, uint64_t num_new, const uint64_t* off, const uint32_t* term_var,
                          const uint64_t* term_coef, int coef_is_mont, const uint64_t* consts, uint64_t* values, int threads);
/* out[i] = a[i] * 2^256 mod r */
void ph_fr_to_mont(const uint64_t* a, uint64_t* out, uint64_t n);
/* out[i] = a[i] * b[i] mod r (self-check hook for the arithmetic) */
void ph_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t n);


/* ---- circuit intake: the iden3 `.r1cs` parser (src/r1cs_file.rs:100-154, src/reader.rs:227-241) and the R1CS -> width-4
 * gate transpilation (src/transpile.rs:92-139 wraps bellman's adaptor; layout: DESIGN.md section 8, strict mode = the shapes
 * the reference's golden vectors pin).  plonkit_b200/circuit.py::_transpile_py and reader.py::_load_r1cs_from_bin_py state
 * the same in Python; tests hold the two equal. */
typedef struct ph_r1cs ph_r1cs;
typedef struct ph_gates ph_gates;
const char* ph_last_error(void);                       /* message of the last failed call on this thread */
int ph_r1cs_parse_bin(const uint8_t* buf, uint64_t len, ph_r1cs** out);            /* 0, or 1 with ph_last_error() */
/* lc_off: 3 * num_constraints + 1 offsets (A, B, C of every constraint) into lc_var / lc_coef (canonical) */
int ph_r1cs_from_csr(uint64_t num_inputs, uint64_t num_aux, uint64_t num_variables, uint64_t num_constraints, const uint64_t* lc_off,
                     const uint32_t* lc_var, const uint64_t* lc_coef, ph_r1cs** out);
void ph_r1cs_free(ph_r1cs* r);
/* out: num_inputs, num_aux, num_variables, num_constraints, num_terms, wire-map length */
void ph_r1cs_header(const ph_r1cs* r, uint64_t out[6]);
void ph_r1cs_export(const ph_r1cs* r, uint64_t* lc_off, uint32_t* lc_var, uint64_t* lc_coef, uint64_t* wire_map);
/* 0 on success; 1 / 2: a constraint shape strict mode refuses (A or B not a single variable / C side not pinned),
 * 3: the contradiction `constant = 0`.  detail = {code, constraint index, C-side variables, constant limbs 0..3} */
int ph_transpile(const ph_r1cs* r, int strict, ph_gates** out, uint64_t detail[7]);
void ph_gates_free(ph_gates* g);
/* out: rows, variables, direct variables (= the R1CS's), hints, constraints that produced gates, witness-program terms */
void ph_gates_header(const ph_gates* g, uint64_t out[6]);
/* wire_idx (4, n) uint32 and selectors (7, n, 4) uint64 zero-filled by the caller, n >= rows (the padded domain);
 * prog_off: variables - direct + 1 offsets; prog_const: one per introduced variable; stats: (constraint, gates) pairs */
void ph_gates_export(const ph_gates* g, uint64_t n, uint32_t* wire_idx, uint64_t* selectors, uint64_t* prog_off, uint32_t* prog_var,
                     uint64_t* prog_coef, uint64_t* prog_const, uint32_t* stat_constraint, uint32_t* stat_gates);

#ifdef __cplusplus
}
#endif
#endif
