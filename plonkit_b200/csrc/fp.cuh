// BN254 prime-field arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256), all in registers.
//
// Replaces, on the device, what ff_ce 0.12.0's derive-generated 4 x u64 Montgomery code does on the CPU for
// pairing_ce's bn256::{Fr,Fq} (Cargo.lock:594-596,1212-1214; reached from src/plonk.rs via bellman_ce).
// Values are kept fully reduced in [0, p) so every canonical output (proof.bin bytes) is unique.
//
// The multiplier is a CIOS Montgomery product organised for the 32x32->64 IMAD.WIDE datapath: partial products
// of even and odd limbs are accumulated into two separately aligned 8-limb accumulators so that every
// mad.lo.cc/madc.hi.cc pair lands on an aligned register pair and carries ride the CC flag (no extra adds).
//
// The same instruction sequence compiles for the host: the carry-flag primitives below have a PTX body under
// __CUDA_ARCH__ and an emulation (explicit carry variable) otherwise.  tests/test_host.py exercises the
// host build against Python integers, which validates the sequence itself; tests/test_gpu_prims.py then checks the
// device build against the same vectors.
#pragma once
#include <cstdint>

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif
#define PK_HD __host__ __device__ __forceinline__

#include "fp_constants.cuh"

namespace pk {

// ---------------------------------------------------------------- carry-flag primitives
namespace cc {
#ifdef __CUDA_ARCH__
PK_HD uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PK_HD uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PK_HD uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PK_HD uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PK_HD uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PK_HD uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PK_HD uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PK_HD uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
PK_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PK_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PK_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
PK_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
#else
// host emulation of the PTX condition-code semantics (CF = carry for add/mad, borrow for sub)
inline uint32_t& CF() { static thread_local uint32_t f = 0; return f; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; CF() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + CF(); CF() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + CF(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; CF() = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - CF(); CF() = (uint32_t)(t >> 32) & 1; return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - CF(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; CF() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + CF(); CF() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)mul_hi(a, b) + c + CF(); CF() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + CF(); }
#endif
}  // namespace cc

// ---------------------------------------------------------------- limb-level kernels (arrays live in registers)
namespace limbs {
using namespace cc;

// acc (aligned pairs at limbs 0,2,4,6) += {a[0],a[2],a[4],a[6]} * b ; leaves the carry-out in CF
PK_HD void mad_pairs_cc(uint32_t* acc, const uint32_t a0, const uint32_t a2, const uint32_t a4, const uint32_t a6, uint32_t b) {
    acc[0] = mad_lo_cc(a0, b, acc[0]);
    acc[1] = madc_hi_cc(a0, b, acc[1]);
    acc[2] = madc_lo_cc(a2, b, acc[2]);
    acc[3] = madc_hi_cc(a2, b, acc[3]);
    acc[4] = madc_lo_cc(a4, b, acc[4]);
    acc[5] = madc_hi_cc(a4, b, acc[5]);
    acc[6] = madc_lo_cc(a6, b, acc[6]);
    acc[7] = madc_hi_cc(a6, b, acc[7]);
}

// Montgomery step on T = X + Y*2^32 (X at limb 0, Y at limb 1): adds m*p with m = -T/p mod 2^32 so that X[0] = 0
template <class P> PK_HD void mont_reduce_row(uint32_t* X, uint32_t* Y) {
    uint32_t m = mul_lo(X[0], P::INV);
    mad_pairs_cc(Y, P::p(1), P::p(3), P::p(5), P::p(7), m);  // carry-out is provably zero (T < 2^288)
    mad_pairs_cc(X, P::p(0), P::p(2), P::p(4), P::p(6), m);
    Y[7] = addc(Y[7], 0);
}

// One CIOS row for multiplier limb b: on entry T = X + Y*2^32 with X[0] == 0; computes
// T' = (T/2^32 + a*b + m*p) with the roles of the two arrays swapped: on exit the new X lives in y[], the new Y in x[]
template <class P> PK_HD void mont_mul_row(uint32_t* x, uint32_t* y, const uint32_t* a, uint32_t b) {
    y[0] = add_cc(y[0], x[1]);  // fold the stray limb of the shifted accumulator, carry rides into the chain below
    x[0] = madc_lo_cc(a[1], b, x[2]);
    x[1] = madc_hi_cc(a[1], b, x[3]);
    x[2] = madc_lo_cc(a[3], b, x[4]);
    x[3] = madc_hi_cc(a[3], b, x[5]);
    x[4] = madc_lo_cc(a[5], b, x[6]);
    x[5] = madc_hi_cc(a[5], b, x[7]);
    x[6] = madc_lo_cc(a[7], b, 0);
    x[7] = madc_hi(a[7], b, 0);
    mad_pairs_cc(y, a[0], a[2], a[4], a[6], b);
    x[7] = addc(x[7], 0);
    mont_reduce_row<P>(y, x);
}

template <class P> PK_HD void cond_sub_p(uint32_t* r) {
    uint32_t t[8];
    t[0] = sub_cc(r[0], P::p(0));
#pragma unroll
    for (int i = 1; i < 8; ++i) t[i] = subc_cc(r[i], P::p(i));
    uint32_t borrow = subc(0, 0);  // 0xffffffff if r < p
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = borrow ? r[i] : t[i];
}

// REDUCE = false leaves the result in [0, 2p) ("lazy" form): for a, b < 2p the CIOS result is below
// (4p^2 + 2^256 p) / 2^256 < 1.76 p because 4p < 2^256, so chains of products and lazy add/sub never need the final
// conditional subtraction; only values that leave the kernel are normalised.
template <class P, bool REDUCE = true> PK_HD void mont_mul(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t ev[8], od[8];
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
        ev[j] = mul_lo(a[j], b[0]);
        ev[j + 1] = mul_hi(a[j], b[0]);
        od[j] = mul_lo(a[j + 1], b[0]);
        od[j + 1] = mul_hi(a[j + 1], b[0]);
    }
    mont_reduce_row<P>(ev, od);
    mont_mul_row<P>(ev, od, a, b[1]);
    mont_mul_row<P>(od, ev, a, b[2]);
    mont_mul_row<P>(ev, od, a, b[3]);
    mont_mul_row<P>(od, ev, a, b[4]);
    mont_mul_row<P>(ev, od, a, b[5]);
    mont_mul_row<P>(od, ev, a, b[6]);
    mont_mul_row<P>(ev, od, a, b[7]);
    // after 8 rows X is in od[] (od[0] == 0) and Y in ev[]:  result = Y + X/2^32  < 2p
    r[0] = add_cc(ev[0], od[1]);
#pragma unroll
    for (int k = 1; k < 7; ++k) r[k] = addc_cc(ev[k], od[k + 1]);
    r[7] = addc(ev[7], 0);
    if (REDUCE) cond_sub_p<P>(r);
}

// ---- lazy form: values in [0, 2p)
template <class P> PK_HD void cond_sub_2p(uint32_t* r) {
    uint32_t t[8];
    t[0] = sub_cc(r[0], P::p2(0));
#pragma unroll
    for (int i = 1; i < 8; ++i) t[i] = subc_cc(r[i], P::p2(i));
    uint32_t borrow = subc(0, 0);  // 0xffffffff if r < 2p
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = borrow ? r[i] : t[i];
}
template <class P> PK_HD void add_lazy(uint32_t* r, const uint32_t* a, const uint32_t* b) {  // a + b < 4p < 2^256
    r[0] = add_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r[i] = addc_cc(a[i], b[i]);
    r[7] = addc(a[7], b[7]);
    cond_sub_2p<P>(r);
}
template <class P> PK_HD void sub_lazy(uint32_t* r, const uint32_t* a, const uint32_t* b) {  // a - b (+ 2p on borrow)
    r[0] = sub_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) r[i] = subc_cc(a[i], b[i]);
    uint32_t borrow = subc(0, 0);
    r[0] = add_cc(r[0], P::p2(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; ++i) r[i] = addc_cc(r[i], P::p2(i) & borrow);
    r[7] = addc(r[7], P::p2(7) & borrow);
}

template <class P> PK_HD void add_mod(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    r[0] = add_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 7; ++i) r[i] = addc_cc(a[i], b[i]);
    r[7] = addc(a[7], b[7]);  // p < 2^254: no carry out of limb 7
    cond_sub_p<P>(r);
}

template <class P> PK_HD void sub_mod(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    r[0] = sub_cc(a[0], b[0]);
#pragma unroll
    for (int i = 1; i < 8; ++i) r[i] = subc_cc(a[i], b[i]);
    uint32_t borrow = subc(0, 0);
    r[0] = add_cc(r[0], P::p(0) & borrow);
#pragma unroll
    for (int i = 1; i < 7; ++i) r[i] = addc_cc(r[i], P::p(i) & borrow);
    r[7] = addc(r[7], P::p(7) & borrow);
}
}  // namespace limbs

// ---------------------------------------------------------------- field element
template <class P> struct alignas(16) Fp {
    typedef P params;
    uint32_t v[8];

    static PK_HD Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = 0;
        return r;
    }
    static PK_HD Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = P::one(i);
        return r;
    }
    PK_HD bool is_zero() const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= v[i];
        return o == 0;
    }
    PK_HD bool operator==(const Fp& b) const {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= v[i] ^ b.v[i];
        return o == 0;
    }
    PK_HD bool operator!=(const Fp& b) const { return !(*this == b); }
    PK_HD Fp operator*(const Fp& b) const { Fp r; limbs::mont_mul<P>(r.v, v, b.v); return r; }
    PK_HD Fp operator+(const Fp& b) const { Fp r; limbs::add_mod<P>(r.v, v, b.v); return r; }
    PK_HD Fp operator-(const Fp& b) const { Fp r; limbs::sub_mod<P>(r.v, v, b.v); return r; }
    PK_HD Fp sqr() const { Fp r; limbs::mont_mul<P>(r.v, v, v); return r; }
    // lazy form ([0, 2p), see limbs::mont_mul<P, false>): used inside the bucket-accumulation loop only
    PK_HD Fp lmul(const Fp& b) const { Fp r; limbs::mont_mul<P, false>(r.v, v, b.v); return r; }
    PK_HD Fp ladd(const Fp& b) const { Fp r; limbs::add_lazy<P>(r.v, v, b.v); return r; }
    PK_HD Fp lsub(const Fp& b) const { Fp r; limbs::sub_lazy<P>(r.v, v, b.v); return r; }
    PK_HD bool lis_zero() const {  // 0 or p
        uint32_t o = 0, q = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { o |= v[i]; q |= v[i] ^ P::p(i); }
        return o == 0 || q == 0;
    }
    PK_HD Fp lnormalize() const { Fp r = *this; limbs::cond_sub_p<P>(r.v); return r; }
    PK_HD Fp dbl() const { Fp r; limbs::add_mod<P>(r.v, v, v); return r; }
    PK_HD Fp neg() const { Fp z = zero(); Fp r; limbs::sub_mod<P>(r.v, z.v, v); return r; }

    // canonical (non-Montgomery) <-> Montgomery
    PK_HD Fp to_mont() const {
        Fp r2;
#pragma unroll
        for (int i = 0; i < 8; ++i) r2.v[i] = P::r2(i);
        return *this * r2;
    }
    PK_HD Fp from_mont() const {
        Fp o = zero();
        o.v[0] = 1;
        return *this * o;
    }
    static PK_HD Fp from_u32(uint32_t x) {
        Fp r = zero();
        r.v[0] = x;
        return r.to_mont();
    }
    // a^e for a 256-bit little-endian exponent
    PK_HD Fp pow(const uint32_t* e) const {
        Fp r = one();
        for (int i = 255; i >= 0; --i) {
            r = r.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) r = r * (*this);
        }
        return r;
    }
    PK_HD Fp pow_u64(uint64_t e) const {
        Fp r = one();
        for (int i = 63; i >= 0; --i) {
            r = r.sqr();
            if ((e >> i) & 1) r = r * (*this);
        }
        return r;
    }
    // Fermat inverse a^(p-2); inverse of 0 is 0
    PK_HD Fp inverse() const {
        Fp r = one();
        for (int i = 253; i >= 0; --i) {
            r = r.sqr();
            if ((P::p_minus_2(i >> 5) >> (i & 31)) & 1) r = r * (*this);
        }
        return r;
    }
};

typedef Fp<FrParams> fr_t;
typedef Fp<FqParams> fq_t;

// 128-bit vectorised global/shared accesses (two LDG.128 / STG.128 per element)
#ifdef __CUDACC__
template <class P> __device__ __forceinline__ Fp<P> ld_fp(const Fp<P>* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fp<P> r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class P> __device__ __forceinline__ Fp<P> ldg_fp(const Fp<P>* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fp<P> r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}
template <class P> __device__ __forceinline__ void st_fp(Fp<P>* p, const Fp<P>& x) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
#endif

}  // namespace pk
