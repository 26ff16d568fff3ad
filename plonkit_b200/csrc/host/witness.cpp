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
#include <thread>
#include <vector>

namespace {

typedef unsigned __int128 u128;

// BN254 scalar field modulus r, little-endian limbs
const uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};

struct Fr {
    uint64_t v[4];
};

uint64_t g_n0 = 0;  // -r^-1 mod 2^64
Fr g_r2;            // 2^512 mod r
bool g_ready = false;

inline bool geq_mod(const uint64_t* a) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] > MOD[i]) return true;
        if (a[i] < MOD[i]) return false;
    }
    return true;
}
inline void sub_mod_raw(uint64_t* a) {
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] - MOD[i] - borrow;
        a[i] = (uint64_t)t;
        borrow = (t >> 64) & 1;
    }
}
inline Fr add(const Fr& a, const Fr& b) {  // a, b < r < 2^254: no carry out
    Fr r;
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
    if (geq_mod(r.v)) sub_mod_raw(r.v);
    return r;
}
// Montgomery product a * b / 2^256 mod r (CIOS)
inline Fr mont_mul(const Fr& a, const Fr& b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
        u128 c = 0;
        for (int j = 0; j < 4; ++j) {
            c += (u128)a.v[j] * b.v[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        const uint64_t m = t[0] * g_n0;
        c = (u128)m * MOD[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; ++j) {
            c += (u128)m * MOD[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fr r;
    memcpy(r.v, t, 32);
    if (t[4] || geq_mod(r.v)) sub_mod_raw(r.v);
    return r;
}
void init() {
    if (g_ready) return;
    uint64_t inv = 1;  // Newton: inv = r^-1 mod 2^64
    for (int i = 0; i < 6; ++i) inv *= 2 - MOD[0] * inv;
    g_n0 = 0 - inv;
    Fr x = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; ++i) x = add(x, x);
    g_r2 = x;
    g_ready = true;
}
inline bool canonical(const uint64_t* a) { return !geq_mod(a); }

// entries [lo, hi) of the program; returns 0 or 1 + the index of the first malformed variable
int64_t run_range(uint64_t lo, uint64_t hi, uint64_t num_direct, const uint64_t* off, const uint32_t* term_var, const uint64_t* term_coef,
                  int coef_is_mont, const uint64_t* consts, uint64_t* values) {
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
            if (!coef_is_mont) c = mont_mul(c, g_r2);
            acc = add(acc, mont_mul(c, x));  // (c R) x / R = c x
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
    init();
    for (uint64_t i = 0; i < num_direct; ++i)
        if (!canonical(values + 4 * i)) return (int64_t)(1 + i);
    if (threads < 1) threads = 1;
    if (threads == 1 || num_new < 4096)
        return run_range(0, num_new, num_direct, off, term_var, term_coef, coef_is_mont, consts, values);
    // sufmin[k] = the oldest introduced variable (as a program index) read by any entry >= k
    std::vector<uint64_t> sufmin(num_new + 1, UINT64_MAX);
    for (uint64_t k = num_new; k-- > 0;) {
        uint64_t m = sufmin[k + 1];
        for (uint64_t j = off[k]; j < off[k + 1]; ++j)
            if (term_var[j] >= num_direct && term_var[j] - num_direct < m) m = term_var[j] - num_direct;
        sufmin[k] = m;
    }
    std::vector<uint64_t> cuts(1, 0);
    const uint64_t total = off[num_new], target = total / (uint64_t)threads + 1;
    for (uint64_t k = 1; k < num_new && (int)cuts.size() < threads; ++k)
        if (off[k] >= target * cuts.size() && sufmin[k] >= k) cuts.push_back(k);
    cuts.push_back(num_new);
    const size_t pieces = cuts.size() - 1;
    std::vector<int64_t> rc(pieces, 0);
    std::vector<std::thread> pool;
    for (size_t t = 1; t < pieces; ++t)
        pool.emplace_back([&, t] { rc[t] = run_range(cuts[t], cuts[t + 1], num_direct, off, term_var, term_coef, coef_is_mont, consts, values); });
    rc[0] = run_range(cuts[0], cuts[1], num_direct, off, term_var, term_coef, coef_is_mont, consts, values);
    for (auto& th : pool) th.join();
    for (size_t t = 0; t < pieces; ++t)
        if (rc[t]) return rc[t];
    return 0;
}

// out[i] = a[i] * 2^256 mod r: canonical -> Montgomery form (the coefficients of a program, once per circuit)
void ph_fr_to_mont(const uint64_t* a, uint64_t* out, uint64_t n) {
    init();
    for (uint64_t i = 0; i < n; ++i) {
        Fr x;
        memcpy(x.v, a + 4 * i, 32);
        Fr r = mont_mul(x, g_r2);
        memcpy(out + 4 * i, r.v, 32);
    }
}

// out[i] = a[i] * b[i] mod r (canonical in, canonical out): self-check hook for the arithmetic above
void ph_fr_mul(const uint64_t* a, const uint64_t* b, uint64_t* out, uint64_t n) {
    init();
    for (uint64_t i = 0; i < n; ++i) {
        Fr x, y;
        memcpy(x.v, a + 4 * i, 32);
        memcpy(y.v, b + 4 * i, 32);
        Fr r = mont_mul(mont_mul(x, g_r2), y);
        memcpy(out + 4 * i, r.v, 32);
    }
}

}  // extern "C"
