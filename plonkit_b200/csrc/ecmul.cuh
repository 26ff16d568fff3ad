// Scalar multiplication k * P for the EC (i)NTT and the SRS generator, with the GLV
// endomorphism of BN254: k = k1 + k2 * LAMBDA (|k1|, |k2| < 2^128) and phi(P) = (BETA x, y) = LAMBDA * P, so
//     k * P = k1 * P + k2 * phi(P)
// runs over ~128 bits instead of 254.  bellman's `mul_assign` walks all 254 bits; the result is the same group element
// either way.
#pragma once
#include "ec.cuh"
#include "glv_constants.cuh"

namespace pk {
namespace glv {

// r[0..nr) = a[0..na) * b[0..nb)  (little-endian 32-bit limbs, nr >= na + nb is NOT required: excess limbs are dropped)
template <int NA, int NB, int NR> PK_HD void mul_limbs(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint64_t acc[NA + NB + 1];
    for (int i = 0; i < NA + NB + 1; ++i) acc[i] = 0;
    for (int i = 0; i < NA; ++i) {
        uint64_t carry = 0;
        for (int j = 0; j < NB; ++j) {
            uint64_t t = (uint64_t)a[i] * b[j] + (acc[i + j] & 0xffffffffu) + carry;
            acc[i + j] = t & 0xffffffffu;
            carry = t >> 32;
        }
        acc[i + NB] += carry;
    }
    for (int i = 0; i < NR; ++i) r[i] = i < NA + NB ? (uint32_t)acc[i] : 0;
}
// x[0..n) -= y[0..n)  (two's complement wrap)
template <int N> PK_HD void sub_limbs(uint32_t* x, const uint32_t* y) {
    uint64_t borrow = 0;
    for (int i = 0; i < N; ++i) {
        uint64_t t = (uint64_t)x[i] - y[i] - borrow;
        x[i] = (uint32_t)t;
        borrow = (t >> 32) & 1;
    }
}
template <int N> PK_HD bool abs_limbs(uint32_t* x) {  // returns true if x was negative (two's complement), x <- |x|
    if (!(x[N - 1] >> 31)) return false;
    uint64_t carry = 1;
    for (int i = 0; i < N; ++i) {
        uint64_t t = (uint64_t)(~x[i]) + carry;
        x[i] = (uint32_t)t;
        carry = t >> 32;
    }
    return true;
}

// canonical k < r  ->  |k1|, |k2| (5 limbs each, < 2^130) and their signs
PK_HD void decompose(const uint32_t* k, uint32_t* k1, bool& neg1, uint32_t* k2, bool& neg2) {
    uint32_t g1[3], g2[5], A[2], B[4], AB[4];
    for (int i = 0; i < 3; ++i) g1[i] = glvc::G1(i);
    for (int i = 0; i < 5; ++i) g2[i] = glvc::G2(i);
    for (int i = 0; i < 2; ++i) A[i] = glvc::A(i);
    for (int i = 0; i < 4; ++i) { B[i] = glvc::B(i); AB[i] = glvc::AB(i); }
    uint32_t t1[11], t2[13];
    mul_limbs<8, 3, 11>(t1, k, g1);
    mul_limbs<8, 5, 13>(t2, k, g2);
    uint32_t c1[3] = {t1[8], t1[9], t1[10]};                       // (k * G1) >> 256
    uint32_t c2[5] = {t2[8], t2[9], t2[10], t2[11], t2[12]};       // (k * G2) >> 256
    // k1 = k - c1 * A - c2 * (A + B)   over 10 limbs (two's complement)
    uint32_t x[10], p1[10], p2[10];
    for (int i = 0; i < 10; ++i) x[i] = i < 8 ? k[i] : 0;
    mul_limbs<3, 2, 10>(p1, c1, A);
    mul_limbs<5, 4, 10>(p2, c2, AB);
    sub_limbs<10>(x, p1);
    sub_limbs<10>(x, p2);
    neg1 = abs_limbs<10>(x);
    for (int i = 0; i < 5; ++i) k1[i] = x[i];
    // k2 = c1 * B - c2 * A
    uint32_t y[10], p3[10];
    mul_limbs<3, 4, 10>(y, c1, B);
    mul_limbs<5, 2, 10>(p3, c2, A);
    sub_limbs<10>(y, p3);
    neg2 = abs_limbs<10>(y);
    for (int i = 0; i < 5; ++i) k2[i] = y[i];
}

}  // namespace glv

// k * P for a canonical scalar k < r (any XYZZ point P, infinity included).
// GLV split, then a JOINT FIXED 2-BIT WINDOW over the two ~128-bit halves: the table holds a P1 + b P2 for a, b in 0..3
// (P1 = +-P, P2 = +-phi(P); 15 points, built with 2 doublings and 11 additions) and every window costs two doublings and
// ONE addition.  The schedule is the same for every lane, which is what matters on a SIMT machine: the bit-by-bit
// version below executes its addition in almost every iteration of a warp (any lane with a set bit pays for all).
// The window loop runs in plain Jacobian coordinates in lazy form (ec.cuh): 2 * 7 + 14 = 28 field products per window,
// none followed by a conditional subtraction, against the 2 * 9 + 14 = 32 of XYZZ (scalar_mul_xyzz below); 65 windows +
// the table + the conversions = ~2.0 K products.
PK_HD g1_xyzz_t scalar_mul(const g1_xyzz_t& P, const fr_t& k_canonical) {
    if (P.is_inf() || k_canonical.is_zero()) return g1_xyzz_t::infinity();
    uint32_t k1[5], k2[5];
    bool n1, n2;
    glv::decompose(k_canonical.v, k1, n1, k2, n2);
    g1_jacc_t T[16];                     // T[0] is never read: a window with both digits zero adds nothing
    fq_t beta;
    for (int i = 0; i < 8; ++i) beta.v[i] = glvc::BETA_MONT(i);
    const g1_jac_t J = g1_jac_t::from_xyzz(P);
    T[1] = J.cached();
    T[4] = T[1];
    T[4].X = T[1].X.lmul(beta);          // phi(P): x = X / Z^2 is scaled by beta
    if (n1) T[1] = T[1].neg();
    if (n2) T[4] = T[4].neg();
    T[2] = g1_jac_t::from_cached(T[1]).dbl().cached();
    T[3] = g1_jac_t::from_cached(T[2]).add(T[1]).cached();
    T[8] = g1_jac_t::from_cached(T[4]).dbl().cached();
    T[12] = g1_jac_t::from_cached(T[8]).add(T[4]).cached();
    for (int b = 1; b < 4; ++b)
        for (int a = 1; a < 4; ++a) T[a + 4 * b] = g1_jac_t::from_cached(T[a]).add(T[4 * b]).cached();
    g1_jac_t r = g1_jac_t::infinity();
    for (int w = 64; w >= 0; --w) {
        r = r.dbl().dbl();
        const int bit = 2 * w;
        const uint32_t a = (k1[bit >> 5] >> (bit & 31)) & 3u, b = (k2[bit >> 5] >> (bit & 31)) & 3u;
        const uint32_t idx = a | (b << 2);
        if (idx) r = r.add(T[idx]);
    }
    return r.to_xyzz();
}

// the same joint 2-bit window entirely in XYZZ coordinates (the first version; ~2.25 K products): kept as a cross-check
PK_HD g1_xyzz_t scalar_mul_xyzz(const g1_xyzz_t& P, const fr_t& k_canonical) {
    if (P.is_inf() || k_canonical.is_zero()) return g1_xyzz_t::infinity();
    uint32_t k1[5], k2[5];
    bool n1, n2;
    glv::decompose(k_canonical.v, k1, n1, k2, n2);
    g1_xyzz_t T[16];
    T[0] = g1_xyzz_t::infinity();
    T[1] = n1 ? P.neg() : P;
    fq_t beta;
    for (int i = 0; i < 8; ++i) beta.v[i] = glvc::BETA_MONT(i);
    T[4] = P;
    T[4].X = P.X * beta;                 // phi(P): x = X / ZZ is scaled by beta
    if (n2) T[4] = T[4].neg();
    T[2] = T[1].dbl();
    T[3] = T[2].add(T[1]);
    T[8] = T[4].dbl();
    T[12] = T[8].add(T[4]);
    for (int b = 1; b < 4; ++b)
        for (int a = 1; a < 4; ++a) T[a + 4 * b] = T[a].add(T[4 * b]);
    g1_xyzz_t r = g1_xyzz_t::infinity();
    for (int w = 64; w >= 0; --w) {
        r = r.dbl().dbl();
        const int bit = 2 * w;
        const uint32_t a = (k1[bit >> 5] >> (bit & 31)) & 3u, b = (k2[bit >> 5] >> (bit & 31)) & 3u;
        const uint32_t idx = a | (b << 2);
        if (idx) r = r.add(T[idx]);
    }
    return r;
}

// the same product by one interleaved double-and-add over the 130 bits of the two halves with the table
// {P1, P2, P1 + P2}: fewer field products on paper (~2.5 K), but data-dependent per lane; kept as a cross-check
PK_HD g1_xyzz_t scalar_mul_bitwise(const g1_xyzz_t& P, const fr_t& k_canonical) {
    if (P.is_inf() || k_canonical.is_zero()) return g1_xyzz_t::infinity();
    uint32_t k1[5], k2[5];
    bool n1, n2;
    glv::decompose(k_canonical.v, k1, n1, k2, n2);
    g1_xyzz_t T[3];
    T[0] = n1 ? P.neg() : P;
    fq_t beta;
    for (int i = 0; i < 8; ++i) beta.v[i] = glvc::BETA_MONT(i);
    T[1] = P;
    T[1].X = P.X * beta;
    if (n2) T[1] = T[1].neg();
    T[2] = T[0].add(T[1]);
    g1_xyzz_t r = g1_xyzz_t::infinity();
    bool started = false;
    for (int i = 129; i >= 0; --i) {
        if (started) r = r.dbl();
        const uint32_t sel = ((k1[i >> 5] >> (i & 31)) & 1u) | (((k2[i >> 5] >> (i & 31)) & 1u) << 1);
        if (sel) {
            r = started ? r.add(T[sel - 1]) : T[sel - 1];
            started = true;
        }
    }
    return r;
}

// plain left-to-right double-and-add over all 254 bits (the reference's schedule): kept as the cross-check
PK_HD g1_xyzz_t scalar_mul_plain(const g1_xyzz_t& P, const fr_t& k_canonical) {
    g1_xyzz_t r = g1_xyzz_t::infinity();
    bool started = false;
    for (int i = 253; i >= 0; --i) {
        if (started) r = r.dbl();
        if ((k_canonical.v[i >> 5] >> (i & 31)) & 1) {
            r = started ? r.add(P) : P;
            started = true;
        }
    }
    return r;
}

}  // namespace pk
