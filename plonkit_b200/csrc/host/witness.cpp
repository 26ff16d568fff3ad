// Host-side witness assignment for the prove path (plain C++, no CUDA): the values of the variables that the
// R1CS -> width-4 transpilation introduces (collapsed linear combinations, running sums) computed from a circom witness.
//
// In the reference this work happens inside `SetupForProver::prove` (src/plonk.rs:132-176): bellman re-synthesises the
// transpiled circuit on every call and evaluates each new variable's value closure against the witness
// (src/circom_circuit.rs:75-133 supplies the witness values).  Here the transpilation runs once, at
// `prepare_setup_for_prover`, and leaves a straight-line program
//       value[num_direct + k] = const_k + sum_j coef_{k,j} * value[var_{k,j}]        (var_{k,j} < num_direct + k)
// that this file evaluates per proof: 254-bit Montgomery arithmetic on 4 x 64-bit limbs.
#include <cstdint>
#include <cstring>
#include <exception>
#include <thread>
#include <vector>

#include "fr_host.hpp"

namespace {

using namespace phost;

// entries [lo, hi) of the program; returns 0 or 1 + the index of the first malformed variable
int64_t run_range(uint64_t lo, uint64_t hi, uint64_t num_direct, const uint64_t* off, const uint32_t* term_var, const uint64_t* term_coef,
                  int coef_is_mont, const uint64_t* consts, uint64_t* values) {
    const FrConstants& K = constants();
    for (uint64_t k = lo; k < hi; ++k) {
        Fr acc;
        memcpy(acc.v, consts + 4 * k, 32);
        if (!canonical(acc.v)) return (int64_t)(1 + num_direct + k);
        for (uint64_t j = off[k]; j < off[k + 1]; ++j) {
            const uint64_t var = term_var[j];
            if (var >= num_direct + k) return (int64_t)(1 + num_direct + k);
            Fr c, x;
            memcpy(c.v, term_coef + 4 * j, 32);
            memcpy(x.v, values + 4 * var, 32);
            if (!canonical(c.v)) return (int64_t)(1 + num_direct + k);
            if (!coef_is_mont) c = mont_mul(c, K.r2, K.n0);
            acc = add(acc, mont_mul(c, x, K.n0));  // (c R) x / R = c x
        }
        memcpy(values + 4 * (num_direct + k), acc.v, 32);
    }
    return 0;
}

}  // namespace

extern "C" {

// values: (num_direct + num_new) x 4 limbs, canonical; the first num_direct rows are filled by the caller (variable 0 and
// the witness values), rows num_direct.. are written here.  off: num_new + 1 offsets into term_var / term_coef
// (coefficients canonical, or in Montgomery form when coef_is_mont); consts: num_new x 4 limbs, canonical.
// threads > 1: the program is cut where no later entry reads a variable introduced before the cut (running sums keep a
// constraint's entries together) and the pieces run concurrently.  Returns 0, or 1 + the index of the first malformed
// variable (a term that reads a variable not yet assigned, or a non-canonical operand).
int64_t ph_assign_witness(uint64_t num_direct, uint64_t num_new, const uint64_t* off, const uint32_t* term_var,
                          const uint64_t* term_coef, int coef_is_mont, const uint64_t* consts, uint64_t* values, int threads) {
    constants();  // initialised before any thread starts
    for (uint64_t i = 0; i < num_direct; ++i)
        if (!canonical(values + 4 * i)) return (int64_t)(1 + i);
    if (threads < 1) threads = 1;
    if (threads == 1 || num_new < 4096)
        return run_range(0, num_new, num_direct, off, term_var, term_coef, coef_is_mont, consts, values);
    std::vector<uint64_t> cuts;
    std::vector<int64_t> rc;
    std::vector<std::thread> pool;
    try {
        // sufmin[k] = the oldest introduced variable (as a program index) read by any entry >= k
        std::vector<uint64_t> sufmin(num_new + 1, UINT64_MAX);
        for (uint64_t k = num_new; k-- > 0;) {
            uint64_t m = sufmin[k + 1];
            for (uint64_t j = off[k]; j < off[k + 1]; ++j)
                if (term_var[j] >= num_direct && term_var[j] - num_direct < m) m = term_var[j] - num_direct;
            sufmin[k] = m;
        }
        cuts.push_back(0);
        const uint64_t total = off[num_new], target = total / (uint64_t)threads + 1;
        for (uint64_t k = 1; k < num_new && (int)cuts.size() < threads; ++k)
            if (off[k] >= target * cuts.size() && sufmin[k] >= k) cuts.push_back(k);
        cuts.push_back(num_new);
        rc.assign(cuts.size() - 1, 0);
        pool.reserve(cuts.size());
    } catch (const std::exception&) {  // no memory for the plan of the split: nothing unwinds across the C boundary
        return run_range(0, num_new, num_direct, off, term_var, term_coef, coef_is_mont, consts, values);
    }
    const size_t pieces = cuts.size() - 1;
    for (size_t t = 1; t < pieces; ++t) {
        try {
            pool.emplace_back([&, t] { rc[t] = run_range(cuts[t], cuts[t + 1], num_direct, off, term_var, term_coef, coef_is_mont, consts, values); });
        } catch (const std::exception&) {  // no thread to be had: this piece runs here
            rc[t] = run_range(cuts[t], cuts[t + 1], num_direct, off, term_var, term_coef, coef_is_mont, consts, values);
        }
    }
    rc[0] = run_range(cuts[0], cuts[1], num_direct, off, term_var, term_coef, coef_is_mont, consts, values);
    for (auto& th : pool) th.join();
    for (size_t t = 0; t < pieces; ++t)
        if (rc[t]) return rc[t];
    return 0;
}

// out[i] = a[i] * 2^256 mod r: canonical -> Montgomery form (the coefficients of a program, once per circuit)
void ph_fr_to_mont(const uint64_t* a, uint64_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        Fr x;
        memcpy(x.v, a + 4 * i, 32);
        Fr r = to_mont(x);
        memcpy(out + 4 * i, r.v, 32);
    }
}

// out[i] = a[i] * b[i] mod r (canonical in, canonical out): self-check hook for the arithmetic above
void ph_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) {
        Fr x, y;
        memcpy(x.v, a + 4 * i, 32);
        memcpy(y.v, b + 4 * i, 32);
        Fr r = mul(x, y);
        memcpy(out + 4 * i, r.v, 32);
    }
}

}  // extern "C"
