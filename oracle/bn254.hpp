// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product; only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
//
// BN254 field and G1 arithmetic for the CPU restatement of plonkit's prove path.
// The arithmetic plonkit uses lives in third-party crates that are NOT under /root/reference:
//   ff_ce 0.12.0 (Cargo.lock:594-596)      -> 4 x u64 little-endian limbs, Montgomery form, R = 2^256
//   pairing_ce 0.24.2 (Cargo.lock:1212-1214) -> bn256::{Fr,Fq,G1Affine,G1 (Jacobian)}
// This file restates their published algorithms (Montgomery CIOS, Jacobian add/double/mixed-add for a=0 curves).
// Constants: SURVEY.md Appendix C (q, r from contrib/template.sol:7-8; curve y^2 = x^3 + 3, G = (1,2),
// contrib/template.sol:67-69).  All Montgomery constants are DERIVED from the modulus at start-up and are checked
// against Appendix C in tests/test_oracle.py.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>

namespace orc {

template <class T> using hvec = std::vector<T>;

typedef uint64_t u64;
typedef unsigned __int128 u128;

// ---------------------------------------------------------------- 256-bit helpers
static inline int cmp4(const u64* a, const u64* b) {
    for (int i = 3; i >= 0; --i) {
        if (a[i] < b[i]) return -1;
        if (a[i] > b[i]) return 1;
    }
    return 0;
}
static inline u64 add4(u64* r, const u64* a, const u64* b) {
    u128 c = 0;
    for (int i = 0; i < 4; ++i) { c += (u128)a[i] + b[i]; r[i] = (u64)c; c >>= 64; }
    return (u64)c;
}
static inline u64 sub4(u64* r, const u64* a, const u64* b) {
    u64 borrow = 0;
    for (int i = 0; i < 4; ++i) {
        u128 t = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)t;
        borrow = (u64)(t >> 64) & 1;
    }
    return borrow;
}

struct FrTag {};
struct FqTag {};

template <class Tag> struct Params {
    static u64 P[4];    // modulus
    static u64 R[4];    // 2^256 mod P   (Montgomery one)
    static u64 R2[4];   // 2^512 mod P
    static u64 INV;     // -P^{-1} mod 2^64
    static bool ready;
};
template <class Tag> u64 Params<Tag>::P[4];
template <class Tag> u64 Params<Tag>::R[4];
template <class Tag> u64 Params<Tag>::R2[4];
template <class Tag> u64 Params<Tag>::INV;
template <class Tag> bool Params<Tag>::ready = false;

template <class Tag> static void derive_params(const u64 p[4]) {
    typedef Params<Tag> PP;
    memcpy(PP::P, p, 32);
    // INV = -p^{-1} mod 2^64 by Newton iteration
    u64 inv = 1;
    for (int i = 0; i < 7; ++i) inv *= 2 - p[0] * inv;
    PP::INV = (u64)0 - inv;
    // R = 2^256 mod p: start from 1, double 256 times with reduction; R2: 512 times
    u64 x[4] = {1, 0, 0, 0};
    for (int i = 0; i < 512; ++i) {
        u64 carry = add4(x, x, x);
        if (carry || cmp4(x, p) >= 0) sub4(x, x, p);
        if (i == 255) memcpy(PP::R, x, 32);
    }
    memcpy(PP::R2, x, 32);
    PP::ready = true;
}

static inline void init_fields() {
    // r, q: contrib/template.sol:7-8
    static const u64 r[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static const u64 q[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    if (!Params<FrTag>::ready) derive_params<FrTag>(r);
    if (!Params<FqTag>::ready) derive_params<FqTag>(q);
}

// ---------------------------------------------------------------- prime field element (Montgomery form)
template <class Tag> struct Fp {
    typedef Params<Tag> PP;
    u64 v[4];

    static Fp zero() { Fp r; memset(r.v, 0, 32); return r; }
    static Fp one() { Fp r; memcpy(r.v, PP::R, 32); return r; }
    bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
    bool operator==(const Fp& o) const { return memcmp(v, o.v, 32) == 0; }
    bool operator!=(const Fp& o) const { return !(*this == o); }

    // Montgomery product (CIOS, 4 x 64-bit limbs), as in ff_ce's derive-generated mont_reduce
    static Fp mont_mul(const Fp& a, const Fp& b) {
        u64 t[6] = {0, 0, 0, 0, 0, 0};
        for (int i = 0; i < 4; ++i) {
            u128 c = 0;
            for (int j = 0; j < 4; ++j) {
                c += (u128)a.v[j] * b.v[i] + t[j];
                t[j] = (u64)c;
                c >>= 64;
            }
            c += t[4];
            t[4] = (u64)c;
            t[5] = (u64)(c >> 64);
            u64 m = t[0] * PP::INV;
            c = (u128)m * PP::P[0] + t[0];
            c >>= 64;
            for (int j = 1; j < 4; ++j) {
                c += (u128)m * PP::P[j] + t[j];
                t[j - 1] = (u64)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (u64)c;
            t[4] = t[5] + (u64)(c >> 64);
        }
        Fp r;
        memcpy(r.v, t, 32);
        if (t[4] || cmp4(r.v, PP::P) >= 0) sub4(r.v, r.v, PP::P);
        return r;
    }
    Fp operator*(const Fp& o) const { return mont_mul(*this, o); }
    Fp& operator*=(const Fp& o) { *this = mont_mul(*this, o); return *this; }
    Fp sqr() const { return mont_mul(*this, *this); }
    Fp operator+(const Fp& o) const {
        Fp r;
        u64 c = add4(r.v, v, o.v);
        if (c || cmp4(r.v, PP::P) >= 0) sub4(r.v, r.v, PP::P);
        return r;
    }
    Fp operator-(const Fp& o) const {
        Fp r;
        if (sub4(r.v, v, o.v)) add4(r.v, r.v, PP::P);
        return r;
    }
    Fp& operator+=(const Fp& o) { *this = *this + o; return *this; }
    Fp& operator-=(const Fp& o) { *this = *this - o; return *this; }
    Fp neg() const { return is_zero() ? *this : zero() - *this; }
    Fp dbl() const { return *this + *this; }

    // canonical (non-Montgomery) little-endian limbs <-> Montgomery
    static Fp from_canonical(const u64 c[4]) {
        Fp a, r2;
        memcpy(a.v, c, 32);
        memcpy(r2.v, PP::R2, 32);
        return mont_mul(a, r2);
    }
    void to_canonical(u64 out[4]) const {
        Fp one_raw;
        one_raw.v[0] = 1; one_raw.v[1] = one_raw.v[2] = one_raw.v[3] = 0;
        Fp r = mont_mul(*this, one_raw);
        memcpy(out, r.v, 32);
    }
    static Fp from_u64(u64 x) { u64 c[4] = {x, 0, 0, 0}; return from_canonical(c); }

    Fp pow(const u64* e, int nlimbs) const {
        Fp r = one();
        bool started = false;
        for (int i = nlimbs * 64 - 1; i >= 0; --i) {
            if (started) r = r.sqr();
            if ((e[i / 64] >> (i % 64)) & 1) { r = started ? r * (*this) : *this; started = true; }
        }
        return r;
    }
    Fp pow_u64(u64 e) const { return pow(&e, 1); }
    Fp inverse() const {  // Fermat: a^(p-2); 0 -> 0
        u64 e[4];
        u64 two[4] = {2, 0, 0, 0};
        sub4(e, PP::P, two);
        return pow(e, 4);
    }
};

typedef Fp<FrTag> Fr;
typedef Fp<FqTag> Fq;

// batch inversion (Montgomery's trick); zeros are left as zeros (bellman's batch_inversion skips them [ext])
template <class F> static void batch_inverse(F* a, size_t n) {
    hvec<F> pre(n);
    F acc = F::one();
    for (size_t i = 0; i < n; ++i) {
        pre[i] = acc;
        if (!a[i].is_zero()) acc *= a[i];
    }
    acc = acc.inverse();
    for (size_t i = n; i-- > 0;) {
        if (a[i].is_zero()) continue;
        F t = acc * pre[i];
        acc *= a[i];
        a[i] = t;
    }
}

// ---------------------------------------------------------------- G1: y^2 = x^3 + 3
struct G1Affine {
    Fq x, y;
    bool inf;
    static G1Affine infinity() { G1Affine p; p.x = Fq::zero(); p.y = Fq::zero(); p.inf = true; return p; }
    static G1Affine generator() { G1Affine p; p.x = Fq::from_u64(1); p.y = Fq::from_u64(2); p.inf = false; return p; }
    bool on_curve() const {
        if (inf) return true;
        return y.sqr() == x.sqr() * x + Fq::from_u64(3);
    }
};

// Jacobian coordinates (X/Z^2, Y/Z^3), the representation pairing_ce's G1 uses [ext]
struct G1 {
    Fq X, Y, Z;
    static G1 infinity() { G1 p; p.X = Fq::zero(); p.Y = Fq::one(); p.Z = Fq::zero(); return p; }
    static G1 from_affine(const G1Affine& a) {
        if (a.inf) return infinity();
        G1 p; p.X = a.x; p.Y = a.y; p.Z = Fq::one(); return p;
    }
    bool is_inf() const { return Z.is_zero(); }

    G1 dbl() const {  // dbl-2009-l
        if (is_inf()) return *this;
        Fq A = X.sqr(), B = Y.sqr(), C = B.sqr();
        Fq D = ((X + B).sqr() - A - C).dbl();
        Fq E = A.dbl() + A, F = E.sqr();
        G1 r;
        r.X = F - D.dbl();
        r.Y = E * (D - r.X) - C.dbl().dbl().dbl();
        r.Z = (Y * Z).dbl();
        return r;
    }
    G1 add(const G1& o) const {  // add-2007-bl with full special-case handling
        if (is_inf()) return o;
        if (o.is_inf()) return *this;
        Fq Z1Z1 = Z.sqr(), Z2Z2 = o.Z.sqr();
        Fq U1 = X * Z2Z2, U2 = o.X * Z1Z1;
        Fq S1 = Y * o.Z * Z2Z2, S2 = o.Y * Z * Z1Z1;
        if (U1 == U2) {
            if (S1 == S2) return dbl();
            return infinity();
        }
        Fq H = U2 - U1, I = H.dbl().sqr(), J = H * I;
        Fq rr = (S2 - S1).dbl(), V = U1 * I;
        G1 r;
        r.X = rr.sqr() - J - V.dbl();
        r.Y = rr * (V - r.X) - (S1 * J).dbl();
        r.Z = ((Z + o.Z).sqr() - Z1Z1 - Z2Z2) * H;
        return r;
    }
    G1 add_mixed(const G1Affine& o) const {  // madd-2007-bl
        if (o.inf) return *this;
        if (is_inf()) return from_affine(o);
        Fq Z1Z1 = Z.sqr();
        Fq U2 = o.x * Z1Z1, S2 = o.y * Z * Z1Z1;
        if (X == U2) {
            if (Y == S2) return dbl();
            return infinity();
        }
        Fq H = U2 - X, HH = H.sqr(), I = HH.dbl().dbl(), J = H * I;
        Fq rr = (S2 - Y).dbl(), V = X * I;
        G1 r;
        r.X = rr.sqr() - J - V.dbl();
        r.Y = rr * (V - r.X) - (Y * J).dbl();
        r.Z = (Z + H).sqr() - Z1Z1 - HH;
        return r;
    }
    G1 neg() const { G1 r = *this; r.Y = Y.neg(); return r; }
    G1Affine to_affine() const {
        if (is_inf()) return G1Affine::infinity();
        Fq zi = Z.inverse(), zi2 = zi.sqr();
        G1Affine a;
        a.x = X * zi2;
        a.y = Y * zi2 * zi;
        a.inf = false;
        return a;
    }
    // double-and-add scalar multiplication by a canonical 256-bit scalar (independent check path)
    G1 mul(const u64 k[4]) const {
        G1 r = infinity();
        for (int i = 255; i >= 0; --i) {
            r = r.dbl();
            if ((k[i / 64] >> (i % 64)) & 1) r = r.add(*this);
        }
        return r;
    }
};

}  // namespace orc
