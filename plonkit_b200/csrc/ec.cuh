// BN254 G1 (y^2 = x^3 + 3) point arithmetic for the MSM / EC-NTT kernels.
//
// Replaces pairing_ce's bn256::{G1Affine, G1} add/double/mixed-add (Cargo.lock:1212-1214), which bellman's
// dense_multiexp calls per (scalar, base) pair (call sites: commit_using_monomials under src/plonk.rs:140,152).
// The reference keeps Jacobian (X/Z^2, Y/Z^3); here accumulators use extended Jacobian "XYZZ" coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2), whose mixed addition needs 10 field multiplications instead of 11 and no
// field inversions.  Results are only ever exported as canonical affine points, so the representation is invisible.
//
// Affine infinity is encoded as (0, 0) (not on the curve); XYZZ infinity as ZZ = 0.
#pragma once
#include "fp.cuh"

namespace pk {

struct alignas(16) g1_affine_t {
    fq_t x, y;
    PK_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
    static PK_HD g1_affine_t infinity() { g1_affine_t p; p.x = fq_t::zero(); p.y = fq_t::zero(); return p; }
    PK_HD g1_affine_t neg() const { g1_affine_t p; p.x = x; p.y = y.neg(); return p; }
};

struct alignas(16) g1_xyzz_t {
    fq_t X, Y, ZZ, ZZZ;

    static PK_HD g1_xyzz_t infinity() {
        g1_xyzz_t p;
        p.X = fq_t::zero(); p.Y = fq_t::zero(); p.ZZ = fq_t::zero(); p.ZZZ = fq_t::zero();
        return p;
    }
    PK_HD bool is_inf() const { return ZZ.is_zero(); }
    static PK_HD g1_xyzz_t from_affine(const g1_affine_t& a) {
        if (a.is_inf()) return infinity();
        g1_xyzz_t p;
        p.X = a.x; p.Y = a.y; p.ZZ = fq_t::one(); p.ZZZ = fq_t::one();
        return p;
    }
    PK_HD g1_xyzz_t neg() const { g1_xyzz_t p = *this; p.Y = Y.neg(); return p; }

    // dbl-2008-s-1 (a = 0): 6M + 3S... expressed with the shared multiplier
    PK_HD g1_xyzz_t dbl() const {
        if (is_inf()) return *this;
        fq_t U = Y.dbl();
        fq_t V = U.sqr();
        fq_t W = U * V;
        fq_t S = X * V;
        fq_t XX = X.sqr();
        fq_t M = XX.dbl() + XX;
        g1_xyzz_t r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * Y;
        r.ZZ = V * ZZ;
        r.ZZZ = W * ZZZ;
        return r;
    }
    static PK_HD g1_xyzz_t dbl_affine(const g1_affine_t& a) {
        if (a.is_inf()) return infinity();
        fq_t U = a.y.dbl();
        fq_t V = U.sqr();
        fq_t W = U * V;
        fq_t S = a.x * V;
        fq_t XX = a.x.sqr();
        fq_t M = XX.dbl() + XX;
        g1_xyzz_t r;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * a.y;
        r.ZZ = V;
        r.ZZZ = W;
        return r;
    }
    // madd-2008-s: this + affine b, all special cases handled (b = inf, this = inf, b = +-this)
    PK_HD g1_xyzz_t add_mixed(const g1_affine_t& b) const {
        if (b.is_inf()) return *this;
        if (is_inf()) return from_affine(b);
        fq_t U2 = b.x * ZZ;
        fq_t S2 = b.y * ZZZ;
        fq_t Pd = U2 - X;
        fq_t Rd = S2 - Y;
        if (Pd.is_zero()) {
            if (Rd.is_zero()) return dbl_affine(b);
            return infinity();
        }
        fq_t PP = Pd.sqr();
        fq_t PPP = Pd * PP;
        fq_t Q = X * PP;
        g1_xyzz_t r;
        r.X = Rd.sqr() - PPP - Q.dbl();
        r.Y = Rd * (Q - r.X) - Y * PPP;
        r.ZZ = ZZ * PP;
        r.ZZZ = ZZZ * PPP;
        return r;
    }
    // add_mixed with the coordinates of *this and of the result in lazy form ([0, 2p), fp.cuh): no conditional
    // subtraction after any of the ten products.  b is canonical.  Used by the bucket-accumulation loop, which
    // normalises (lnorm) only when a run's sum leaves the registers.
    PK_HD g1_xyzz_t add_mixed_lazy(const g1_affine_t& b) const {
        if (b.is_inf()) return *this;
        if (ZZ.lis_zero()) return from_affine(b);
        fq_t U2 = b.x.lmul(ZZ);
        fq_t S2 = b.y.lmul(ZZZ);
        fq_t Pd = U2.lsub(X);
        fq_t Rd = S2.lsub(Y);
        if (Pd.lis_zero()) {
            if (Rd.lis_zero()) return dbl_affine(b);
            return infinity();
        }
        fq_t PP = Pd.lmul(Pd);
        fq_t PPP = Pd.lmul(PP);
        fq_t Q = X.lmul(PP);
        g1_xyzz_t r;
        r.X = Rd.lmul(Rd).lsub(PPP).lsub(Q.ladd(Q));
        r.Y = Rd.lmul(Q.lsub(r.X)).lsub(Y.lmul(PPP));
        r.ZZ = ZZ.lmul(PP);
        r.ZZZ = ZZZ.lmul(PPP);
        return r;
    }
    PK_HD g1_xyzz_t lnorm() const {
        g1_xyzz_t r;
        r.X = X.lnormalize(); r.Y = Y.lnormalize(); r.ZZ = ZZ.lnormalize(); r.ZZZ = ZZZ.lnormalize();
        return r;
    }
    // add-2008-s: general addition with all special cases
    PK_HD g1_xyzz_t add(const g1_xyzz_t& b) const {
        if (b.is_inf()) return *this;
        if (is_inf()) return b;
        fq_t U1 = X * b.ZZ;
        fq_t U2 = b.X * ZZ;
        fq_t S1 = Y * b.ZZZ;
        fq_t S2 = b.Y * ZZZ;
        fq_t Pd = U2 - U1;
        fq_t Rd = S2 - S1;
        if (Pd.is_zero()) {
            if (Rd.is_zero()) return dbl();
            return infinity();
        }
        fq_t PP = Pd.sqr();
        fq_t PPP = Pd * PP;
        fq_t Q = U1 * PP;
        g1_xyzz_t r;
        r.X = Rd.sqr() - PPP - Q.dbl();
        r.Y = Rd * (Q - r.X) - S1 * PPP;
        r.ZZ = ZZ * b.ZZ * PP;
        r.ZZZ = ZZZ * b.ZZZ * PPP;
        return r;
    }
    // one field inversion: x = X/ZZ, y = Y/ZZZ with 1/ZZ = ZZZ * I, 1/ZZZ = ZZ * I... where I = (ZZ*ZZZ)^-1
    PK_HD g1_affine_t to_affine() const {
        if (is_inf()) return g1_affine_t::infinity();
        fq_t I = (ZZ * ZZZ).inverse();
        g1_affine_t a;
        a.x = X * (ZZZ * I);
        a.y = Y * (ZZ * I);
        return a;
    }
    // k * this for a small non-negative k (bucket-reduction offsets): left-to-right double-and-add
    PK_HD g1_xyzz_t mul_small(uint32_t k) const {
        g1_xyzz_t r = infinity();
        for (int i = 31; i >= 0; --i) {
            r = r.dbl();
            if ((k >> i) & 1) r = r.add(*this);
        }
        return r;
    }
};

#ifdef __CUDACC__
__device__ __forceinline__ g1_affine_t ldg_affine(const g1_affine_t* p) {
    g1_affine_t a;
    a.x = ldg_fp(&p->x);
    a.y = ldg_fp(&p->y);
    return a;
}
__device__ __forceinline__ g1_xyzz_t ld_xyzz(const g1_xyzz_t* p) {
    g1_xyzz_t a;
    a.X = ld_fp(&p->X); a.Y = ld_fp(&p->Y); a.ZZ = ld_fp(&p->ZZ); a.ZZZ = ld_fp(&p->ZZZ);
    return a;
}
__device__ __forceinline__ void st_xyzz(g1_xyzz_t* p, const g1_xyzz_t& a) {
    st_fp(&p->X, a.X); st_fp(&p->Y, a.Y); st_fp(&p->ZZ, a.ZZ); st_fp(&p->ZZZ, a.ZZZ);
}
__device__ __forceinline__ void st_affine(g1_affine_t* p, const g1_affine_t& a) {
    st_fp(&p->x, a.x);
    st_fp(&p->y, a.y);
}
#endif

}  // namespace pk
