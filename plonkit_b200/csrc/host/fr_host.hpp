// BN254 scalar field on the host: 4 x 64-bit limbs, canonical values in and out (Montgomery form only inside mul).
// Shared by the plain C++ host helpers (witness assignment, R1CS parsing and transpilation); no CUDA.
#pragma once
#include <cstdint>
#include <cstring>

namespace phost {

typedef unsigned __int128 u128;

// modulus r, little-endian limbs
static const uint64_t MOD[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};

struct Fr {
    uint64_t v[4];
};

struct FrConstants {
    uint64_t n0;  // -r^-1 mod 2^64
    Fr r2;        // 2^512 mod r
};

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
inline bool canonical(const uint64_t* a) { return !geq_mod(a); }
inline bool is_zero(const Fr& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
inline bool equal(const Fr& a, const Fr& b) { return memcmp(a.v, b.v, 32) == 0; }
inline Fr zero() { Fr r = {{0, 0, 0, 0}}; return r; }
inline Fr from_u64(uint64_t x) { Fr r = {{x, 0, 0, 0}}; return r; }
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
inline Fr neg(const Fr& a) {
    if (is_zero(a)) return a;
    Fr r;
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)MOD[i] - a.v[i] - borrow;
        r.v[i] = (uint64_t)t;
        borrow = (t >> 64) & 1;
    }
    return r;
}
inline Fr sub(const Fr& a, const Fr& b) { return add(a, neg(b)); }
// Montgomery product a * b / 2^256 mod r (CIOS)
inline Fr mont_mul(const Fr& a, const Fr& b, uint64_t n0) {
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
        const uint64_t m = t[0] * n0;
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
inline const FrConstants& constants() {
    static const FrConstants k = [] {
        FrConstants c;
        uint64_t inv = 1;  // Newton: inv = r^-1 mod 2^64
        for (int i = 0; i < 6; ++i) inv *= 2 - MOD[0] * inv;
        c.n0 = 0 - inv;
        Fr x = {{1, 0, 0, 0}};
        for (int i = 0; i < 512; ++i) x = add(x, x);
        c.r2 = x;
        return c;
    }();
    return k;
}
inline Fr to_mont(const Fr& a) { const FrConstants& k = constants(); return mont_mul(a, k.r2, k.n0); }
// canonical * canonical -> canonical
inline Fr mul(const Fr& a, const Fr& b) { const FrConstants& k = constants(); return mont_mul(mont_mul(a, k.r2, k.n0), b, k.n0); }

}  // namespace phost
