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

#ifdef __cplusplus
}
#endif
#endif
