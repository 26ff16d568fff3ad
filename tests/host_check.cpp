// Host build (g++) of the product's field / curve / transcript headers, for CPU-side checks of the exact
// instruction sequences the device code runs (the carry-flag primitives are emulated on the host, see fp.cuh).
#include "../plonkit_b200/csrc/ec.cuh"
#include "../plonkit_b200/csrc/ecmul.cuh"
#include "../tools/micro/fp_f64.cuh"
#include "../tools/micro/fp_wide.cuh"
#include <cfenv>
#include "../plonkit_b200/csrc/keccak_host.hpp"
using namespace pk;

template <class F> static F load_c(const uint64_t* s) { F x; memcpy(x.v, s, 32); return x.to_mont(); }
template <class F> static void store_c(const F& x, uint64_t* d) { F c = x.from_mont(); memcpy(d, c.v, 32); }
static g1_affine_t load_pt(const uint64_t* s) { g1_affine_t p; bool z = true; for (int i = 0; i < 8; ++i) if (s[i]) z = false;
    if (z) return g1_affine_t::infinity(); p.x = load_c<fq_t>(s); p.y = load_c<fq_t>(s + 4); return p; }
static void store_pt(const g1_affine_t& p, uint64_t* d) { if (p.is_inf()) { memset(d, 0, 64); return; } store_c(p.x, d); store_c(p.y, d + 4); }

extern "C" {
void hc_mont_mul(int which, const uint32_t* a, const uint32_t* b, uint32_t* o, int n) {
    for (int i = 0; i < n; ++i) { if (which == 0) limbs::mont_mul<FrParams>(o + 8 * i, a + 8 * i, b + 8 * i); else limbs::mont_mul<FqParams>(o + 8 * i, a + 8 * i, b + 8 * i); }
}
// the FP64-pipe multiplier (fp_f64.cuh): same contract as hc_mont_mul; the device uses DFMA.RZ, the host emulates it by
// running fma() under round-toward-zero
void hc_mont_mul_f64(int which, const uint32_t* a, const uint32_t* b, uint32_t* o, int n) {
    const int old = fegetround();
    fesetround(FE_TOWARDZERO);
    for (int i = 0; i < n; ++i) { if (which == 0) f64::mont_mul<FrParams>(o + 8 * i, a + 8 * i, b + 8 * i); else f64::mont_mul<FqParams>(o + 8 * i, a + 8 * i, b + 8 * i); }
    fesetround(old);
}
// Karatsuba + separate-reduction multiplier (fp.cuh): mode 0 = mont_mul_k(a,b), 1 = mont_sqr_k(a), 2 = mont_mul_sub_mul(a,b,c,d)
void hc_mont_wide(int which, int mode, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* o, int n) {
    for (int i = 0; i < n; ++i) {
        const uint32_t *A = a + 8 * i, *B = b + 8 * i, *C = c + 8 * i, *D = d + 8 * i;
        uint32_t* O = o + 8 * i;
        if (which == 0) {
            if (mode == 0) limbs::mont_mul_k<FrParams>(O, A, B); else if (mode == 1) limbs::mont_sqr_k<FrParams>(O, A); else limbs::mont_mul_sub_mul<FrParams>(O, A, B, C, D);
        } else {
            if (mode == 0) limbs::mont_mul_k<FqParams>(O, A, B); else if (mode == 1) limbs::mont_sqr_k<FqParams>(O, A); else limbs::mont_mul_sub_mul<FqParams>(O, A, B, C, D);
        }
    }
}
void hc_fr_ops(const uint64_t* a, const uint64_t* b, uint64_t* add, uint64_t* sub, uint64_t* mul, uint64_t* inv, int n) {
    for (int i = 0; i < n; ++i) {
        fr_t x = load_c<fr_t>(a + 4 * i), y = load_c<fr_t>(b + 4 * i);
        store_c(x + y, add + 4 * i); store_c(x - y, sub + 4 * i); store_c(x * y, mul + 4 * i); store_c(x.inverse(), inv + 4 * i);
    }
}
// out[0] = p + q (mixed), out[1] = p + q (full), out[2] = 2p, out[3] = k * p (mul_small),
// out[4] = ((p + q) + q) + p through add_mixed_lazy chained WITHOUT normalising in between
void hc_g1_ops(const uint64_t* p, const uint64_t* q, uint32_t k, uint64_t* out) {
    g1_affine_t P = load_pt(p), Q = load_pt(q);
    g1_xyzz_t X = g1_xyzz_t::from_affine(P), Y = g1_xyzz_t::from_affine(Q);
    // make the accumulator non-trivial (ZZ != 1) by going through 3P - 2P
    g1_xyzz_t X3 = X.dbl().add(X);
    g1_xyzz_t Xn = X3.add(X.dbl().neg());
    store_pt(Xn.add_mixed(Q).to_affine(), out);
    store_pt(Xn.add(Y.dbl().add(Y.neg())).to_affine(), out + 8);
    store_pt(Xn.dbl().to_affine(), out + 16);
    store_pt(Xn.mul_small(k).to_affine(), out + 24);
    // lazy chain: Xn + Q + Q + P  (second step hits the doubling-free general path with a non-trivial ZZ; coordinates stay
    // in [0, 2p) across the three additions and are normalised once)
    g1_xyzz_t L = Xn.add_mixed_lazy(Q).add_mixed_lazy(Q).add_mixed_lazy(P).lnorm();
    store_pt(L.to_affine(), out + 32);
}
// Jacobian window-loop arithmetic (ec.cuh g1_jac_t / g1_jacc_t) on accumulators with non-trivial Z:
// out[0] = p + q, out[1] = 2p, out[2] = (p + q) + q, each through XYZZ -> Jacobian -> XYZZ
void hc_jac_ops(const uint64_t* p, const uint64_t* q, uint64_t* out) {
    g1_xyzz_t X = g1_xyzz_t::from_affine(load_pt(p)), Y = g1_xyzz_t::from_affine(load_pt(q));
    g1_xyzz_t Xn = X.dbl().add(X).add(X.dbl().neg());   // == p with ZZ != 1 (infinity stays infinity)
    g1_xyzz_t Yn = Y.dbl().add(Y).add(Y.dbl().neg());
    g1_jac_t J = g1_jac_t::from_xyzz(Xn);
    g1_jacc_t C = g1_jac_t::from_xyzz(Yn).cached();
    store_pt(J.add(C).to_xyzz().to_affine(), out);
    store_pt(J.dbl().to_xyzz().to_affine(), out + 8);
    store_pt(J.add(C).add(C).to_xyzz().to_affine(), out + 16);
}
void hc_lazy_field(const uint32_t* a, const uint32_t* b, uint32_t* o, int n) {
    // o = normalise( lsub( lmul(lmul(a,b), ladd(a,b)), lmul(b,b) ) )  vs the same with canonical ops computed by the caller
    for (int i = 0; i < n; ++i) {
        fq_t x, y; memcpy(x.v, a + 8 * i, 32); memcpy(y.v, b + 8 * i, 32);
        fq_t r = x.lmul(y).lmul(x.ladd(y)).lsub(y.lmul(y)).lsub(x.ladd(x).ladd(y)).lnormalize();
        memcpy(o + 8 * i, r.v, 32);
    }
}
// GLV: out_k = (|k1| (5 limbs), sign1, |k2| (5 limbs), sign2); out_pt[0] = k * p (GLV, joint 2-bit window), out_pt[1] = k * p
// (plain 254-bit walk), out_pt[2] = k * p (GLV, bit by bit), out_pt[3] = k * p (GLV, joint window, all in XYZZ)
void hc_glv(const uint64_t* p, const uint32_t* k_canonical, uint32_t* out_k, uint64_t* out_pt) {
    bool n1, n2;
    glv::decompose(k_canonical, out_k, n1, out_k + 6, n2);
    out_k[5] = n1; out_k[11] = n2;
    fr_t k; memcpy(k.v, k_canonical, 32);
    g1_xyzz_t P = g1_xyzz_t::from_affine(load_pt(p));
    g1_xyzz_t P3 = P.dbl().add(P);                       // non-trivial ZZ
    g1_xyzz_t Q = P3.add(P.dbl().neg());                 // == P with ZZ != 1
    store_pt(scalar_mul(Q, k).to_affine(), out_pt);
    store_pt(scalar_mul_plain(Q, k).to_affine(), out_pt + 8);
    store_pt(scalar_mul_bitwise(Q, k).to_affine(), out_pt + 16);
    store_pt(scalar_mul_xyzz(Q, k).to_affine(), out_pt + 24);
}
void hc_keccak(const uint8_t* d, uint64_t n, uint8_t* out) { Keccak256 h; h.update(d, n); h.finish(out); }
// transcript: commit `n` 32-byte big-endian values, then draw `m` challenges (canonical LE limbs out)
void hc_transcript(const uint8_t* vals, int n, int m, uint32_t* out) {
    RollingKeccakTranscript t;
    for (int i = 0; i < n; ++i) t.commit_be(vals + 32 * i);
    for (int i = 0; i < m; ++i) t.challenge(out + 8 * i);
}
void hc_root(int log_n, uint64_t* out) {
    fr_t g; for (int i = 0; i < 8; ++i) g.v[i] = FrRoots::root_2_28(i);
    for (int i = 0; i < 28 - log_n; ++i) g = g.sqr();
    store_c(g, out);
}
}
