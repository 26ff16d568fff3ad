// Second Montgomery multiplier for BN254 on sm_100a: the product runs on the FP64 pipe (DFMA) instead of the
// 32-bit integer multiplier (IMAD.WIDE), with the SAME contract as limbs::mont_mul in fp.cuh:
//     r = a * b * 2^-256 mod p,   a, b, r fully reduced 8 x u32 little-endian limbs.
// EXPERIMENT, not linked into the library.  Idea: the kernels bound by the integer-multiply pipe (bucket accumulation,
// NTT butterflies) could issue some of their products through the FP64 pipe.  Measured on B200 (tools/micro/fmulbench.cu,
// pipes*.cu; numbers in DESIGN.md section 4.5): this routine alone reaches 49 G products/s against 67 G for the integer
// one, and running both in the same thread does not add up (IMAD.WIDE and DFMA contend), so the product path stays on
// IMAD.WIDE.  Kept, with its host test, as the evidence behind that decision.
//
// Method (exact integer arithmetic carried out in doubles; after Emmart, Zheng & Weems, "Faster modular
// exponentiation using double precision floating point arithmetic on the GPU", ARITH 2018):
//   * operands are split into 5 limbs of 51 bits, each an exactly representable double;
//   * a limb product P = x*y < 2^102 is split as  s' = fma_rz(x, y, s)  (s in [2^104, 2^105): ulp 2^52, so the
//     mantissa of s' accumulates floor-parts of the column's products — a whole column rides ONE chain) and
//     lo = fma_rz(x, y, (s - s') + 2^52) = 2^52 + (low 52 bits, exact): its raw bit pattern is added to a 64-bit
//     integer column sum (the exponent bits are a known constant, removed once per column);
//   * word-by-word Montgomery reduction with 51-bit words: q_i = t_i * (-p^-1) mod 2^51, products q_i * p_j go through
//     the same split; after 5 rounds the upper columns hold a*b*2^-255 mod p (< 2p), which is halved mod p
//     (2^-255 -> 2^-256) while it is repacked into 32-bit words, and conditionally reduced.
// Range proofs of the chains (all checked by tools/micro/f64_model.py over adversarial limb patterns):
//   a, b < p < 2^254  =>  top limbs < 2^50; a column of a*b holds at most 4 products < 2^102 or (column 4) 3 such
//   products and 2 < 2^101: sum < 2^104, so s never leaves [2^104, 2^105).  Column c of q*p: q_i <= 2^51 - 1 and
//   sum_j p_j < 2^53 for both BN254 moduli  =>  sum < 2^104 as well.
#pragma once
#include "../../plonkit_b200/csrc/fp.cuh"
#ifndef __CUDA_ARCH__
#include <cmath>
#include <cstring>
#endif

namespace pk {
namespace f64 {

constexpr int LB = 51;
constexpr uint64_t LMASK = (uint64_t(1) << LB) - 1;
constexpr uint64_t BITS_2P52 = 0x4330000000000000ull;   // bit pattern of 2^52
constexpr uint64_t BITS_2P104 = 0x4670000000000000ull;  // bit pattern of 2^104

// host: the caller runs under fesetround(FE_TOWARDZERO) (tests/host_check.cpp does); every DADD here is exact
PK_HD double fma_rz(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rz(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
PK_HD uint64_t dbits(double x) {
#ifdef __CUDA_ARCH__
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
PK_HD double bits_d(uint64_t u) {
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)u);
#else
    double x; memcpy(&x, &u, 8); return x;
#endif
}
// exact double of an integer v < 2^52
PK_HD double u52_to_double(uint64_t v) { return bits_d(BITS_2P52 | v) - 4503599627370496.0; }

// 8 x u32 -> 5 doubles (51-bit limbs)
PK_HD void split51(const uint32_t* a, double* A) {
    uint64_t w[5];
    w[0] = (uint64_t)a[0] | ((uint64_t)a[1] << 32);
    w[1] = (uint64_t)a[2] | ((uint64_t)a[3] << 32);
    w[2] = (uint64_t)a[4] | ((uint64_t)a[5] << 32);
    w[3] = (uint64_t)a[6] | ((uint64_t)a[7] << 32);
    A[0] = u52_to_double(w[0] & LMASK);
    A[1] = u52_to_double(((w[0] >> 51) | (w[1] << 13)) & LMASK);
    A[2] = u52_to_double(((w[1] >> 38) | (w[2] << 26)) & LMASK);
    A[3] = u52_to_double(((w[2] >> 25) | (w[3] << 39)) & LMASK);
    A[4] = u52_to_double(w[3] >> 12);
}

// one limb product into a column: hi part rides the chain s, anchored low part is added to the integer sum
PK_HD void mac(double x, double y, double& s, uint64_t& lo) {
    const double s2 = fma_rz(x, y, s);
    const double d = (s - s2) + 4503599627370496.0;
    lo += dbits(fma_rz(x, y, d));
    s = s2;
}

template <class P> PK_HD void mont_mul_split(uint32_t* r, const double* A, const double* B) {
    const double C0 = 20282409603651670423947251286016.0;  // 2^104
    uint64_t lo[10];
    double s[9];
#pragma unroll
    for (int c = 0; c < 10; ++c) lo[c] = 0;
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        s[c] = C0;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int j = c - i;
            if (j >= 0 && j <= 4) mac(A[i], B[j], s[c], lo[c]);
        }
    }
    // column totals so far: T_c = (lo[c] - n_c * BITS_2P52) + 2 * (dbits(s[c-1]) - BITS_2P104)
    double sq[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) sq[c] = C0;
    uint64_t carry = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        // n products of a*b in column i: i + 1; of q*p so far: i
        uint64_t t = lo[i] + carry - (uint64_t)(2 * i + 1) * BITS_2P52;
        if (i > 0) t += 2 * (dbits(s[i - 1]) - BITS_2P104) + 2 * (dbits(sq[i - 1]) - BITS_2P104);
        const uint64_t qi = (t * P::INV51) & LMASK;
        const double Q = u52_to_double(qi);
#pragma unroll
        for (int j = 0; j < 5; ++j) mac(Q, (double)P::p51(j), sq[i + j], lo[i + j]);
        // the q_i * p_0 low part just landed in lo[i]: the column is now 0 mod 2^51
        t = lo[i] + carry - (uint64_t)(2 * i + 2) * BITS_2P52;
        if (i > 0) t += 2 * (dbits(s[i - 1]) - BITS_2P104) + 2 * (dbits(sq[i - 1]) - BITS_2P104);
        carry = t >> LB;
    }
    uint64_t v[5];
#pragma unroll
    for (int c = 5; c < 10; ++c) {
        // a*b products in column c: 9 - c; q*p products: 9 - c
        uint64_t t = lo[c] + carry - (uint64_t)(2 * (9 - c)) * BITS_2P52;
        t += 2 * (dbits(s[c - 1]) - BITS_2P104) + 2 * (dbits(sq[c - 1]) - BITS_2P104);
        v[c - 5] = t;
        carry = 0;
        if (c == 5) {
            // v = a*b*2^-255 mod p up to one p: make it even by adding p when odd, the halving happens in the repack
            const uint64_t odd = 0 - (t & 1);
#pragma unroll
            for (int k = 0; k < 5; ++k) lo[5 + k] += P::p51(k) & odd;
            v[0] = t + (P::p51(0) & odd);
        }
    }
    // carry propagation over the 51-bit limbs
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k + 1] += v[k] >> LB;
        v[k] &= LMASK;
    }
    // repack (value >> 1) into 4 x u64
    uint64_t w0 = (v[0] >> 1) | (v[1] << 50);
    uint64_t w1 = (v[1] >> 14) | (v[2] << 37);
    uint64_t w2 = (v[2] >> 27) | (v[3] << 24);
    uint64_t w3 = (v[3] >> 40) | (v[4] << 11);
    r[0] = (uint32_t)w0; r[1] = (uint32_t)(w0 >> 32);
    r[2] = (uint32_t)w1; r[3] = (uint32_t)(w1 >> 32);
    r[4] = (uint32_t)w2; r[5] = (uint32_t)(w2 >> 32);
    r[6] = (uint32_t)w3; r[7] = (uint32_t)(w3 >> 32);
    limbs::cond_sub_p<P>(r);
}

template <class P> PK_HD void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    double A[5], B[5];
    split51(a, A);
    split51(b, B);
    mont_mul_split<P>(r, A, B);
}

}  // namespace f64

template <class P> PK_HD Fp<P> mul_f64(const Fp<P>& a, const Fp<P>& b) {
    Fp<P> r;
    f64::mont_mul<P>(r.v, a.v, b.v);
    return r;
}

}  // namespace pk
