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

// Plain Jacobian coordinates (x = X/Z^2, y = Y/Z^3; infinity: Z = 0) for the window loop of the scalar multiplication
// (ecmul.cuh): with a = 0 a doubling is 7 field products against the 9 of XYZZ, and a window there is two doublings per
// addition.  g1_jacc_t is a table entry: a Jacobian point with Z^2 and Z^3 kept beside it, so that adding it costs 14.
// All coordinates of both types are in LAZY form ([0, 2p), fp.cuh): no conditional subtraction after any product; only
// to_xyzz() normalises.
struct g1_jacc_t {
    fq_t X, Y, Z, ZZ, ZZZ;
    PK_HD bool is_inf() const { return Z.lis_zero(); }
    PK_HD g1_jacc_t neg() const { g1_jacc_t r = *this; r.Y = fq_t::zero().lsub(Y); return r; }
};
struct g1_jac_t {
    fq_t X, Y, Z;

    static PK_HD g1_jac_t infinity() {
        g1_jac_t p;
        p.X = fq_t::zero(); p.Y = fq_t::zero(); p.Z = fq_t::zero();
        return p;
    }
    PK_HD bool is_inf() const { return Z.lis_zero(); }
    // an XYZZ point (canonical coordinates) as Jacobian with Z := ZZ: Z^2 = ZZ^2 and Z^3 = ZZ^3 = ZZZ^2, so X' = X ZZ, Y' = Y ZZZ
    static PK_HD g1_jac_t from_xyzz(const g1_xyzz_t& p) {
        if (p.is_inf()) return infinity();
        g1_jac_t r;
        r.X = p.X.lmul(p.ZZ); r.Y = p.Y.lmul(p.ZZZ); r.Z = p.ZZ;
        return r;
    }
    PK_HD g1_xyzz_t to_xyzz() const {
        if (is_inf()) return g1_xyzz_t::infinity();
        g1_xyzz_t r;
        const fq_t zz = Z.lmul(Z);
        r.X = X.lnormalize(); r.Y = Y.lnormalize(); r.ZZ = zz.lnormalize(); r.ZZZ = zz.lmul(Z).lnormalize();
        return r;
    }
    PK_HD g1_jacc_t cached() const {
        g1_jacc_t r;
        r.X = X; r.Y = Y; r.Z = Z; r.ZZ = Z.lmul(Z); r.ZZZ = r.ZZ.lmul(Z);
        return r;
    }
    static PK_HD g1_jac_t from_cached(const g1_jacc_t& c) { g1_jac_t r; r.X = c.X; r.Y = c.Y; r.Z = c.Z; return r; }
    // a = 0 doubling, 3M + 4S: A = X^2, B2 = 2 Y^2, C2 = B2^2 = 4 Y^4, D = 2 X B2 = 4 X Y^2, E = 3 A;
    // X3 = E^2 - 2 D, Y3 = E (D - X3) - 2 C2, Z3 = 2 Y Z
    PK_HD g1_jac_t dbl() const {
        if (is_inf()) return *this;
        g1_jac_t r;
        r.Z = Y.lmul(Z);
        r.Z = r.Z.ladd(r.Z);
        fq_t B2 = Y.lmul(Y);
        B2 = B2.ladd(B2);
        fq_t D = X.lmul(B2);
        D = D.ladd(D);
        fq_t C2 = B2.lmul(B2);
        C2 = C2.ladd(C2);
        fq_t E = X.lmul(X);
        E = E.ladd(E).ladd(E);
        r.X = E.lmul(E).lsub(D.ladd(D));
        r.Y = E.lmul(D.lsub(r.X)).lsub(C2);
        return r;
    }
    // add-2007-bl with the second operand's Z^2, Z^3 given: 11M + 3S, all special cases handled
    PK_HD g1_jac_t add(const g1_jacc_t& b) const {
        if (b.is_inf()) return *this;
        if (is_inf()) return from_cached(b);
        const fq_t Z1Z1 = Z.lmul(Z);
        const fq_t U1 = X.lmul(b.ZZ);
        const fq_t U2 = b.X.lmul(Z1Z1);
        const fq_t S1 = Y.lmul(b.ZZZ);
        const fq_t S2 = b.Y.lmul(Z.lmul(Z1Z1));
        const fq_t H = U2.lsub(U1);
        const fq_t R = S2.lsub(S1);
        if (H.lis_zero()) {
            if (R.lis_zero()) return dbl();
            return infinity();
        }
        const fq_t HH = H.lmul(H);
        const fq_t HHH = H.lmul(HH);
        const fq_t V = U1.lmul(HH);
        g1_jac_t r;
        r.X = R.lmul(R).lsub(HHH).lsub(V.ladd(V));
        r.Y = R.lmul(V.lsub(r.X)).lsub(S1.lmul(HHH));
        r.Z = Z.lmul(b.Z).lmul(H);
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
