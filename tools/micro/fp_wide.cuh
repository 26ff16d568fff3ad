// EXPERIMENT, not linked into the library: Karatsuba product + separate Montgomery reduction, dedicated squaring and
// lazy reduction of a difference of products for the 8 x 32-bit-limb BN254 fields.
//
// Measured on B200 (kmulbench.cu, G products/s, chains of dependent products, 256 threads x 8 blocks/SM):
//     CIOS (fp.cuh, 128 IMAD.WIDE)          67.0
//     Karatsuba mul (48 + 64 IMAD.WIDE)     62.3 .. 65.2   <- slower: the ~150 extra IADD3/LOP3 are not free once the
//                                                              multiply pipe is saturated (ALU issues 64/clk/SM beside it)
//     squaring (3 x sqr4, 30 + 64)          74.4
//     a*b - c*d with one reduction          79.7 per product
// Used for sqr() and the Y-coordinate of the XYZZ additions, the bucket-accumulation kernel went from 102 to 116
// registers and its time per launch did not move (6.45 vs 6.42 ms), so the library keeps the plain CIOS product.
// tests/test_host.py checks these sequences against Python integers (host build), kmulbench.cu against CIOS on the GPU.
#pragma once
#include "../../plonkit_b200/csrc/fp.cuh"

namespace pk {
namespace limbs {
// ---------------------------------------------------------------- wide (unreduced) products: Karatsuba + separate reduction
// Measured on B200 (tools/micro/README.md): every integer multiply issues on ONE pipe (IMAD.WIDE = 2 units at 64
// units/clk/SM) while the ALU pipe (IADD3/LOP3, 128/clk/SM) idles under the multiplier.  So additions are traded for
// multiplies: one level of subtractive Karatsuba brings the 8x8-limb product from 64 to 48 IMAD.WIDE, a squaring to 30,
// and keeping product and Montgomery reduction apart lets sums of products share one reduction (lazy reduction).

// r[0..7] = a[0..3] * b[0..3].  Partial products at even limb positions accumulate in E (aligned pairs at limbs
// 0,2,4,6), those at odd positions in O (O[k] is limb k+1); every carry lands in a slot that holds at most a few units.
PK_HD void mul4(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t E[8], O[7];
    E[0] = mul_lo(a[0], b[0]); E[1] = mul_hi(a[0], b[0]);
    E[2] = mul_lo(a[2], b[0]); E[3] = mul_hi(a[2], b[0]);
    O[0] = mul_lo(a[1], b[0]); O[1] = mul_hi(a[1], b[0]);
    O[2] = mul_lo(a[3], b[0]); O[3] = mul_hi(a[3], b[0]);
    // b[1]
    O[0] = mad_lo_cc(a[0], b[1], O[0]); O[1] = madc_hi_cc(a[0], b[1], O[1]);
    O[2] = madc_lo_cc(a[2], b[1], O[2]); O[3] = madc_hi_cc(a[2], b[1], O[3]);
    O[4] = addc(0, 0);
    E[2] = mad_lo_cc(a[1], b[1], E[2]); E[3] = madc_hi_cc(a[1], b[1], E[3]);
    E[4] = madc_lo_cc(a[3], b[1], 0); E[5] = madc_hi(a[3], b[1], 0);
    // b[2]
    E[2] = mad_lo_cc(a[0], b[2], E[2]); E[3] = madc_hi_cc(a[0], b[2], E[3]);
    E[4] = madc_lo_cc(a[2], b[2], E[4]); E[5] = madc_hi_cc(a[2], b[2], E[5]);
    E[6] = addc(0, 0);
    O[2] = mad_lo_cc(a[1], b[2], O[2]); O[3] = madc_hi_cc(a[1], b[2], O[3]);
    O[4] = madc_lo_cc(a[3], b[2], O[4]); O[5] = madc_hi(a[3], b[2], 0);
    // b[3]
    O[2] = mad_lo_cc(a[0], b[3], O[2]); O[3] = madc_hi_cc(a[0], b[3], O[3]);
    O[4] = madc_lo_cc(a[2], b[3], O[4]); O[5] = madc_hi_cc(a[2], b[3], O[5]);
    O[6] = addc(0, 0);
    E[4] = mad_lo_cc(a[1], b[3], E[4]); E[5] = madc_hi_cc(a[1], b[3], E[5]);
    E[6] = madc_lo_cc(a[3], b[3], E[6]); E[7] = madc_hi(a[3], b[3], 0);
    r[0] = E[0];
    r[1] = add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 7; ++k) r[k] = addc_cc(E[k], O[k - 1]);
    r[7] = addc(E[7], O[6]);
}

// r[0..7] = a[0..3]^2: 4 diagonal + 6 off-diagonal products (doubled by a one-bit shift)
PK_HD void sqr4(uint32_t* r, const uint32_t* a) {
    uint32_t D[8], E[4], O[6], S[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { D[2 * i] = mul_lo(a[i], a[i]); D[2 * i + 1] = mul_hi(a[i], a[i]); }
    E[0] = mul_lo(a[0], a[2]); E[1] = mul_hi(a[0], a[2]);  // limbs 2,3
    E[2] = mul_lo(a[1], a[3]); E[3] = mul_hi(a[1], a[3]);  // limbs 4,5
    O[0] = mul_lo(a[0], a[1]); O[1] = mul_hi(a[0], a[1]);  // limbs 1,2
    O[2] = mul_lo(a[0], a[3]); O[3] = mul_hi(a[0], a[3]);  // limbs 3,4
    O[2] = mad_lo_cc(a[1], a[2], O[2]); O[3] = madc_hi_cc(a[1], a[2], O[3]);
    O[4] = madc_lo_cc(a[2], a[3], 0); O[5] = madc_hi(a[2], a[3], 0);  // limbs 5,6
    S[1] = O[0];
    S[2] = add_cc(O[1], E[0]);
    S[3] = addc_cc(O[2], E[1]);
    S[4] = addc_cc(O[3], E[2]);
    S[5] = addc_cc(O[4], E[3]);
    S[6] = addc_cc(O[5], 0);
    S[7] = addc(0, 0);
    // r = D + 2 * S
    r[0] = D[0];
    r[1] = add_cc(D[1], S[1] << 1);
#pragma unroll
    for (int k = 2; k < 7; ++k) r[k] = addc_cc(D[k], (S[k] << 1) | (S[k - 1] >> 31));
    r[7] = addc(D[7], (S[7] << 1) | (S[6] >> 31));
}

// d = |x - y| over 4 limbs; returns 0xffffffff if x < y, else 0
PK_HD uint32_t abs_diff4(uint32_t* d, const uint32_t* x, const uint32_t* y) {
    d[0] = sub_cc(x[0], y[0]);
    d[1] = subc_cc(x[1], y[1]);
    d[2] = subc_cc(x[2], y[2]);
    d[3] = subc_cc(x[3], y[3]);
    const uint32_t m = subc(0, 0);
    d[0] = sub_cc(d[0] ^ m, m);
    d[1] = subc_cc(d[1] ^ m, m);
    d[2] = subc_cc(d[2] ^ m, m);
    d[3] = subc(d[3] ^ m, m);
    return m;
}

// T[4..15] += mid[0..8] (the Karatsuba middle term at 2^128)
PK_HD void add_middle(uint32_t* T, const uint32_t* mid) {
    T[4] = add_cc(T[4], mid[0]);
#pragma unroll
    for (int k = 1; k < 9; ++k) T[4 + k] = addc_cc(T[4 + k], mid[k]);
    T[13] = addc_cc(T[13], 0);
    T[14] = addc_cc(T[14], 0);
    T[15] = addc(T[15], 0);
}

// T[0..15] = a[0..7] * b[0..7]:  a_lo b_hi + a_hi b_lo = z0 + z2 + (a_lo - a_hi)(b_hi - b_lo)
PK_HD void mul_wide(uint32_t* T, const uint32_t* a, const uint32_t* b) {
    uint32_t da[4], db[4], zm[8], mid[9];
    mul4(T, a, b);
    mul4(T + 8, a + 4, b + 4);
    const uint32_t sa = abs_diff4(da, a, a + 4);
    const uint32_t sb = abs_diff4(db, b + 4, b);
    mul4(zm, da, db);
    const uint32_t s = sa ^ sb;  // all ones: the cross term is negative
    mid[0] = add_cc(T[0], T[8]);
#pragma unroll
    for (int k = 1; k < 8; ++k) mid[k] = addc_cc(T[k], T[8 + k]);
    mid[8] = addc(0, 0);
    add_cc(s, 1);  // CF = (s != 0): the +1 of the two's complement
#pragma unroll
    for (int k = 0; k < 8; ++k) mid[k] = addc_cc(mid[k], zm[k] ^ s);
    mid[8] = addc(mid[8], s);
    add_middle(T, mid);
}

// T[0..15] = a[0..7]^2:  2 a_lo a_hi = a_lo^2 + a_hi^2 - (a_lo - a_hi)^2  (three 4-limb squarings, 30 IMAD.WIDE)
PK_HD void sqr_wide(uint32_t* T, const uint32_t* a) {
    uint32_t d[4], zm[8], mid[9];
    sqr4(T, a);
    sqr4(T + 8, a + 4);
    abs_diff4(d, a, a + 4);
    sqr4(zm, d);
    mid[0] = add_cc(T[0], T[8]);
#pragma unroll
    for (int k = 1; k < 8; ++k) mid[k] = addc_cc(T[k], T[8 + k]);
    mid[8] = addc(0, 0);
    mid[0] = sub_cc(mid[0], zm[0]);
#pragma unroll
    for (int k = 1; k < 8; ++k) mid[k] = subc_cc(mid[k], zm[k]);
    mid[8] = subc(mid[8], 0);
    add_middle(T, mid);
}

// r = T * 2^-256 mod p for a 16-limb T < 2 p^2 (one conditional subtraction suffices: r < 2 p^2 / 2^256 + p < 2p).
// Same two-array scheme as mont_mul (value = X + Y * 2^32), with the upper limbs of T fed in as the window slides.
template <class P> PK_HD void mont_reduce_wide(uint32_t* r, const uint32_t* T) {
    uint32_t X[8], Y[8], m;
#pragma unroll
    for (int k = 0; k < 8; ++k) X[k] = T[k];
    m = mul_lo(X[0], P::INV);
#pragma unroll
    for (int k = 0; k < 4; ++k) { Y[2 * k] = mul_lo(P::p(2 * k + 1), m); Y[2 * k + 1] = mul_hi(P::p(2 * k + 1), m); }
    mad_pairs_cc(X, P::p(0), P::p(2), P::p(4), P::p(6), m);
    Y[7] = addc(Y[7], 0);
#pragma unroll
    for (int i = 1; i < 8; ++i) {
        // slide by one limb: new X = Y + X[1] (limb 0), new Y = X[2..7] ++ T[7 + i]; the carry of the fold rides into Y
        uint32_t Z[8];
        Y[0] = add_cc(Y[0], X[1]);
        m = mul_lo(Y[0], P::INV);
        Z[0] = madc_lo_cc(P::p(1), m, X[2]);
        Z[1] = madc_hi_cc(P::p(1), m, X[3]);
        Z[2] = madc_lo_cc(P::p(3), m, X[4]);
        Z[3] = madc_hi_cc(P::p(3), m, X[5]);
        Z[4] = madc_lo_cc(P::p(5), m, X[6]);
        Z[5] = madc_hi_cc(P::p(5), m, X[7]);
        Z[6] = madc_lo_cc(P::p(7), m, T[7 + i]);
        Z[7] = madc_hi(P::p(7), m, 0);
        mad_pairs_cc(Y, P::p(0), P::p(2), P::p(4), P::p(6), m);
        Z[7] = addc(Z[7], 0);
#pragma unroll
        for (int k = 0; k < 8; ++k) { X[k] = Y[k]; Y[k] = Z[k]; }
    }
    r[0] = add_cc(Y[0], X[1]);
#pragma unroll
    for (int k = 1; k < 7; ++k) r[k] = addc_cc(Y[k], X[k + 1]);
    r[7] = addc(Y[7], T[15]);
    cond_sub_p<P>(r);
}

template <class P> PK_HD void mont_mul_k(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t T[16];
    mul_wide(T, a, b);
    mont_reduce_wide<P>(r, T);
}
template <class P> PK_HD void mont_sqr_k(uint32_t* r, const uint32_t* a) {
    uint32_t T[16];
    sqr_wide(T, a);
    mont_reduce_wide<P>(r, T);
}
// r = (a*b - c*d) * 2^-256 mod p with ONE reduction: a*b - c*d + p^2 lies in (0, 2 p^2)
template <class P> PK_HD void mont_mul_sub_mul(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d) {
    uint32_t T[16], U[16];
    mul_wide(T, a, b);
    mul_wide(U, c, d);
    T[0] = sub_cc(T[0], U[0]);
#pragma unroll
    for (int k = 1; k < 15; ++k) T[k] = subc_cc(T[k], U[k]);
    T[15] = subc(T[15], U[15]);
    T[0] = add_cc(T[0], P::p_sq(0));
#pragma unroll
    for (int k = 1; k < 15; ++k) T[k] = addc_cc(T[k], P::p_sq(k));
    T[15] = addc(T[15], P::p_sq(15));
    mont_reduce_wide<P>(r, T);
}

}  // namespace limbs
}  // namespace pk
