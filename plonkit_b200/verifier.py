"""`plonk::verify` (src/plonk.rs:189-210 -> bellman better_cs::verifier::verify) on the host, in Python integers.

Verification is host code in the reference too (no NTT, no MSM: ~30 scalar multiplications and one pairing check), so it
is not a GPU path; it is here so that the prove-path mirror is usable end to end (`plonkit verify`, src/bin/main.rs:427-438)
and so that the reference's fourth unit test (src/tests.rs:75-81: verify(vk.bin, proof.bin) == true) has its counterpart.

The verification identities and the transcript are the ones spelled out by the in-tree Solidity verifier,
contrib/template.sol (:267-307 transcript, :445-494 verify_at_z, :496-586 reconstruct_d, :588-689 verify_commitments,
:691-758 verify_initial); the final check is e(P_G, [1]G2) * e(P_X, [x]G2) == 1 with the two G2 elements of the
verification key.  The pairing is the optimal ate pairing over BN254 with the usual tower-free representation
Fq12 = Fq[w] / (w^12 - 18 w^6 + 82).
"""
from .bn254 import NON_RESIDUES, Q_MOD, R_MOD, limbs_to_ints, root_of_unity

# ---------------------------------------------------------------- Keccak-256 (original 0x01 padding; hashlib only has SHA-3)
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
       0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
       0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(x, n):
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def _keccak_f(a):
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i:off + 8 * i + 8], "little")
        a = _keccak_f(a)
    return b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))


class RollingKeccakTranscript:
    """contrib/template.sol:267-307"""

    def __init__(self):
        self.s0 = self.s1 = b"\x00" * 32
        self.ctr = 0

    def update_u256(self, v: int):
        body = self.s0 + self.s1 + int(v).to_bytes(32, "big")
        self.s0, self.s1 = keccak256(b"\x00\x00\x00\x00" + body), keccak256(b"\x00\x00\x00\x01" + body)

    def update_g1(self, p):
        x, y = (0, 0) if p is None else p          # infinity is absorbed as (0, 0)
        self.update_u256(x)
        self.update_u256(y)

    def challenge(self) -> int:
        h = keccak256(b"\x00\x00\x00\x02" + self.s0 + self.s1 + self.ctr.to_bytes(4, "big"))
        self.ctr += 1
        return int.from_bytes(h, "big") & ((1 << 253) - 1)


# ---------------------------------------------------------------- G1 (affine, Python ints; None = infinity)
def g1_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if (y1 + y2) % Q_MOD == 0:
            return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, Q_MOD) % Q_MOD
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q_MOD) % Q_MOD
    x3 = (lam * lam - x1 - x2) % Q_MOD
    return x3, (lam * (x1 - x3) - y1) % Q_MOD


def g1_neg(p):
    return None if p is None else (p[0], (-p[1]) % Q_MOD)


def g1_mul(p, k):
    k %= R_MOD
    acc = None
    while k:
        if k & 1:
            acc = g1_add(acc, p)
        p = g1_add(p, p)
        k >>= 1
    return acc


def g1_on_curve(p):
    return p is None or (p[1] * p[1] - p[0] ** 3 - 3) % Q_MOD == 0


# ---------------------------------------------------------------- extension fields as polynomials over Fq
class FQP:
    degree = 0
    mod = ()      # low-degree coefficients c_i of the modulus x^degree + sum c_i x^i

    def __init__(self, coeffs):
        self.c = [int(x) % Q_MOD for x in coeffs]

    @classmethod
    def one(cls):
        return cls([1] + [0] * (cls.degree - 1))

    @classmethod
    def zero(cls):
        return cls([0] * cls.degree)

    def __add__(self, o):
        return type(self)([a + b for a, b in zip(self.c, o.c)])

    def __sub__(self, o):
        return type(self)([a - b for a, b in zip(self.c, o.c)])

    def __neg__(self):
        return type(self)([-a for a in self.c])

    def __eq__(self, o):
        return self.c == o.c

    def scale(self, k):
        return type(self)([a * k for a in self.c])

    def __mul__(self, o):
        d = self.degree
        b = [0] * (2 * d - 1)
        for i, x in enumerate(self.c):
            if x:
                for j, y in enumerate(o.c):
                    b[i + j] += x * y
        for top in range(2 * d - 2, d - 1, -1):      # reduce by x^d = -sum mod_i x^i
            t = b[top]
            if t:
                for i, m in enumerate(self.mod):
                    if m:
                        b[top - d + i] -= t * m
        return type(self)(b[:d])

    def __pow__(self, e):
        r, base = type(self).one(), self
        while e:
            if e & 1:
                r = r * base
            base = base * base
            e >>= 1
        return r

    def inv(self):
        """extended Euclid over Fq[x]"""
        d = self.degree
        lm, hm = [1] + [0] * d, [0] * (d + 1)
        low, high = self.c + [0], list(self.mod) + [1]

        def deg(p):
            k = len(p) - 1
            while k and p[k] % Q_MOD == 0:
                k -= 1
            return k

        def rounded_div(a, b):
            dega, degb = deg(a), deg(b)
            temp, o = list(a), [0] * len(a)
            binv = pow(b[degb], -1, Q_MOD)
            for i in range(dega - degb, -1, -1):
                o[i] = (o[i] + temp[degb + i] * binv) % Q_MOD
                for c in range(degb + 1):
                    temp[c + i] = (temp[c + i] - o[i] * b[c]) % Q_MOD
            return o[:deg(o) + 1]
        while deg(low):
            r = rounded_div(high, low)
            r += [0] * (d + 1 - len(r))
            nm, new = list(hm), list(high)
            for i in range(d + 1):
                for j in range(d + 1 - i):
                    nm[i + j] -= lm[i] * r[j]
                    new[i + j] -= low[i] * r[j]
            nm, new = [x % Q_MOD for x in nm], [x % Q_MOD for x in new]
            lm, low, hm, high = nm, new, lm, low
        k = pow(low[0], -1, Q_MOD)
        return type(self)([x * k for x in lm[:d]])


class FQ2(FQP):
    degree = 2
    mod = (1, 0)                                    # u^2 + 1


class FQ12(FQP):
    degree = 12
    mod = (82, 0, 0, 0, 0, 0, -18, 0, 0, 0, 0, 0)   # w^12 - 18 w^6 + 82


# ---------------------------------------------------------------- curve arithmetic over any of the fields (affine, None = infinity)
def _double(p):
    x, y = p
    lam = (x * x).scale(3) * y.scale(2).inv()
    nx = lam * lam - x.scale(2)
    return nx, lam * (x - nx) - y


def _add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        return _double(p) if y1 == y2 else None
    lam = (y2 - y1) * (x2 - x1).inv()
    nx = lam * lam - x1 - x2
    return nx, lam * (x1 - nx) - y1


def _linefunc(p1, p2, t):
    """the line through p1 and p2 (tangent if equal), evaluated at t"""
    (x1, y1), (x2, y2), (xt, yt) = p1, p2, t
    if not x1 == x2:
        lam = (y2 - y1) * (x2 - x1).inv()
        return lam * (xt - x1) - (yt - y1)
    if y1 == y2:
        lam = (x1 * x1).scale(3) * y1.scale(2).inv()
        return lam * (xt - x1) - (yt - y1)
    return xt - x1


_W = FQ12([0, 1] + [0] * 10)
ATE_LOOP_COUNT = 29793968203157093288       # 6 u + 2, u = 4965661367192848881
LOG_ATE_LOOP_COUNT = 63
B2 = FQ2([3, 0]) * FQ2([9, 1]).inv()        # the twist y^2 = x^3 + 3 / (9 + u)


def _twist(q):
    """E'(Fq2) -> E(Fq12): the isomorphism Fq2 = Fq[u]/(u^2+1) -> Fq[w^6] (u = w^6 - 9), then x / w^2... as (x w^2, y w^3)"""
    x, y = q
    nx = FQ12([x.c[0] - 9 * x.c[1]] + [0] * 5 + [x.c[1]] + [0] * 5)
    ny = FQ12([y.c[0] - 9 * y.c[1]] + [0] * 5 + [y.c[1]] + [0] * 5)
    return nx * _W * _W, ny * _W * _W * _W


def g2_on_curve(q):
    return q is None or (q[1] * q[1] - q[0] * q[0] * q[0]) == B2


def miller_loop(q12, p12):
    if q12 is None or p12 is None:
        return FQ12.one()
    r, f = q12, FQ12.one()
    for i in range(LOG_ATE_LOOP_COUNT, -1, -1):
        f = f * f * _linefunc(r, r, p12)
        r = _double(r)
        if ATE_LOOP_COUNT & (1 << i):
            f = f * _linefunc(r, q12, p12)
            r = _add(r, q12)
    q1 = (q12[0] ** Q_MOD, q12[1] ** Q_MOD)
    nq2 = (q1[0] ** Q_MOD, -(q1[1] ** Q_MOD))
    f = f * _linefunc(r, q1, p12)
    r = _add(r, q1)
    f = f * _linefunc(r, nq2, p12)
    return f


def final_exponentiate(f):
    return f ** ((Q_MOD ** 12 - 1) // R_MOD)


def pairing_product_is_one(pairs) -> bool:
    """prod e(P_i, Q_i) == 1 for pairs (P in G1 as ints or None, Q in G2 as (FQ2, FQ2) or None)"""
    f = FQ12.one()
    for p, q in pairs:
        if p is None or q is None:
            continue
        p12 = (FQ12([p[0]] + [0] * 11), FQ12([p[1]] + [0] * 11))
        f = f * miller_loop(_twist(q), p12)
    return final_exponentiate(f) == FQ12.one()


def g2_from_bytes(buf: bytes):
    """128 B uncompressed: x.c1 | x.c0 | y.c1 | y.c0, big-endian"""
    if buf[0] & 0x40:
        return None
    v = [int.from_bytes(buf[32 * i:32 * i + 32], "big") for i in range(4)]
    return FQ2([v[1], v[0]]), FQ2([v[3], v[2]])


# ---------------------------------------------------------------- the verifier
def _pt(limbs):
    x, y = limbs_to_ints(limbs)
    return None if x == 0 and y == 0 else (x, y)


def verify(vk, proof) -> bool:
    """better_cs::verifier::verify::<_, _, RollingKeccakTranscript>(proof, vk, None)"""
    n = proof.n + 1
    if n & (n - 1) or proof.n != vk.n or proof.num_inputs != vk.num_inputs or len(proof.input_values) != proof.num_inputs:
        return False
    log_n = n.bit_length() - 1
    sel = [_pt(c) for c in vk.selector_commitments]          # q_a, q_b, q_c, q_d, q_m, q_const
    qdn = _pt(vk.next_step_selector_commitments[0])
    sig = [_pt(c) for c in vk.permutation_commitments]
    cw = [_pt(c) for c in proof.wire_commitments]
    cz = _pt(proof.grand_product_commitment)
    ct = [_pt(c) for c in proof.quotient_poly_commitments]
    w1, w2 = _pt(proof.opening_at_z_proof), _pt(proof.opening_at_z_omega_proof)
    if not all(g1_on_curve(p) for p in sel + [qdn] + sig + cw + [cz] + ct + [w1, w2]):
        return False
    g2 = [g2_from_bytes(vk.g2_raw[:128]), g2_from_bytes(vk.g2_raw[128:256])]
    if len(vk.g2_raw) != 256 or not all(g2_on_curve(q) for q in g2):
        return False
    wz, dzw = proof.wire_values_at_z, proof.wire_values_at_z_omega[0]
    zzw, tz, rz, sz = proof.grand_product_at_z_omega, proof.quotient_polynomial_at_z, proof.linearization_polynomial_at_z, proof.permutation_polynomials_at_z
    if any(v >= R_MOD for v in list(proof.input_values) + list(wz) + [dzw, zzw, tz, rz] + list(sz)):
        return False

    tr = RollingKeccakTranscript()
    for v in proof.input_values:
        tr.update_u256(v)
    for p in cw:
        tr.update_g1(p)
    beta, gamma = tr.challenge(), tr.challenge()
    tr.update_g1(cz)
    alpha = tr.challenge()
    for p in ct:
        tr.update_g1(p)
    zeta = tr.challenge()
    for v in list(wz) + [dzw] + list(sz) + [tz, rz, zzw]:
        tr.update_u256(v)
    v = tr.challenge()
    tr.update_g1(w1)
    tr.update_g1(w2)
    u = tr.challenge()

    r_ = R_MOD
    omega = root_of_unity(log_n)
    zeta_n = pow(zeta, n, r_)
    zh = (zeta_n - 1) % r_
    if zh == 0:
        return False
    n_inv = pow(n, -1, r_)

    def lagrange(i):
        wi = pow(omega, i, r_)
        return wi * zh % r_ * n_inv % r_ * pow(zeta - wi, -1, r_) % r_
    l0 = lagrange(0)
    k = NON_RESIDUES
    # verify_at_z
    rhs = rz
    for i, inp in enumerate(proof.input_values):
        rhs = (rhs + lagrange(i) * inp) % r_
    zpart = zzw
    for i in range(3):
        zpart = zpart * ((sz[i] * beta + gamma + wz[i]) % r_) % r_
    zpart = zpart * ((gamma + wz[3]) % r_) % r_
    rhs = (rhs - zpart * alpha - l0 * alpha * alpha) % r_
    if zh * tz % r_ != rhs:
        return False
    # reconstruct_d: the commitment of the linearisation polynomial, times v, plus the Z(X) opening term at z omega
    d = sel[5]
    for i in range(4):
        d = g1_add(d, g1_mul(sel[i], wz[i]))
    d = g1_add(d, g1_mul(sel[4], wz[0] * wz[1] % r_))
    d = g1_add(d, g1_mul(qdn, dzw))
    gpz = alpha
    for i in range(4):
        gpz = gpz * ((zeta * k[i] * beta + gamma + wz[i]) % r_) % r_
    gpz = (gpz + l0 * alpha * alpha) % r_
    lastp = beta * zzw % r_ * alpha % r_
    for i in range(3):
        lastp = lastp * ((beta * sz[i] + gamma + wz[i]) % r_) % r_
    d = g1_add(g1_add(d, g1_mul(cz, gpz)), g1_neg(g1_mul(sig[3], lastp)))
    d = g1_mul(d, v)
    d = g1_add(d, g1_mul(cz, pow(v, 9, r_) * u % r_))
    # verify_commitments
    agg = ct[0]
    zp = 1
    for i in range(1, 4):
        zp = zp * zeta_n % r_
        agg = g1_add(agg, g1_mul(ct[i], zp))
    agg = g1_add(agg, d)
    ac = v
    for i in range(4):
        ac = ac * v % r_
        agg = g1_add(agg, g1_mul(cw[i], ac))
    for i in range(3):
        ac = ac * v % r_
        agg = g1_add(agg, g1_mul(sig[i], ac))
    ac = ac * v % r_ * v % r_                                  # v^10
    agg = g1_add(agg, g1_mul(cw[3], ac * u % r_))
    val = (tz + v * rz) % r_
    a2 = v
    for i in range(4):
        a2 = a2 * v % r_
        val = (val + wz[i] * a2) % r_
    for i in range(3):
        a2 = a2 * v % r_
        val = (val + sz[i] * a2) % r_
    a2 = a2 * v % r_
    val = (val + zzw * a2 % r_ * u) % r_
    a2 = a2 * v % r_
    val = (val + dzw * a2 % r_ * u) % r_
    agg = g1_add(agg, g1_neg(g1_mul((1, 2), val)))
    pair_with_generator = g1_add(g1_add(agg, g1_mul(w1, zeta)), g1_mul(w2, zeta * omega % r_ * u % r_))
    pair_with_x = g1_neg(g1_add(w1, g1_mul(w2, u)))
    return pairing_product_is_one([(pair_with_generator, g2[0]), (pair_with_x, g2[1])])


def verify_gated(commitments, g2_raw, proof) -> bool:
    """Verifier of the two-gate-type prover (plonkit_b200.recursive; protocol of DESIGN.md section 9, byte layout unpinned).
    `commitments` (13, 8): q_a, q_b, q_c, q_d, q_m, q_const, q_dnext, s_main, s_resc, sigma_0..3.  Gate terms: the main gate
    enters the linearisation scaled by s_main(z); the Rescue x^5 gate (a^2 = b, b^2 = c, c a = d) enters the identity at z as
    s_resc(z) (alpha (a^2 - b) + alpha^2 (b^2 - c) + alpha^3 (c a - d)); the permutation argument uses alpha^4 and L_0 alpha^5."""
    n = proof.n + 1
    if n & (n - 1) or proof.gate_selectors_at_z is None or len(proof.input_values) != proof.num_inputs:
        return False
    log_n = n.bit_length() - 1
    vk = [_pt(c) for c in commitments]
    cw = [_pt(c) for c in proof.wire_commitments]
    cz = _pt(proof.grand_product_commitment)
    ct = [_pt(c) for c in proof.quotient_poly_commitments]
    w1, w2 = _pt(proof.opening_at_z_proof), _pt(proof.opening_at_z_omega_proof)
    if len(vk) != 13 or not all(g1_on_curve(p) for p in vk + cw + [cz] + ct + [w1, w2]):
        return False
    if len(g2_raw) != 256:
        return False
    g2 = [g2_from_bytes(g2_raw[:128]), g2_from_bytes(g2_raw[128:256])]
    if not all(g2_on_curve(q) for q in g2):
        return False
    wz, dzw = proof.wire_values_at_z, proof.wire_values_at_z_omega[0]
    zzw, tz, rz, sz = proof.grand_product_at_z_omega, proof.quotient_polynomial_at_z, proof.linearization_polynomial_at_z, proof.permutation_polynomials_at_z
    smz, srz = proof.gate_selectors_at_z
    if any(x >= R_MOD for x in list(proof.input_values) + list(wz) + [dzw, zzw, tz, rz, smz, srz] + list(sz)):
        return False

    tr = RollingKeccakTranscript()
    for x in proof.input_values:
        tr.update_u256(x)
    for p in cw:
        tr.update_g1(p)
    beta, gamma = tr.challenge(), tr.challenge()
    tr.update_g1(cz)
    alpha = tr.challenge()
    for p in ct:
        tr.update_g1(p)
    zeta = tr.challenge()
    for x in list(wz) + [dzw, smz, srz] + list(sz) + [tz, rz, zzw]:
        tr.update_u256(x)
    v = tr.challenge()
    tr.update_g1(w1)
    tr.update_g1(w2)
    u = tr.challenge()

    r_ = R_MOD
    al = [pow(alpha, i, r_) for i in range(6)]
    omega = root_of_unity(log_n)
    zeta_n = pow(zeta, n, r_)
    zh = (zeta_n - 1) % r_
    if zh == 0:
        return False
    n_inv = pow(n, -1, r_)

    def lagrange(i):
        wi = pow(omega, i, r_)
        return wi * zh % r_ * n_inv % r_ * pow(zeta - wi, -1, r_) % r_
    l0 = lagrange(0)
    k = NON_RESIDUES
    # the quotient identity at z
    rhs = rz
    for i, inp in enumerate(proof.input_values):
        rhs = (rhs + lagrange(i) * inp) % r_
    rhs = (rhs + srz * (al[1] * (wz[0] * wz[0] - wz[1]) + al[2] * (wz[1] * wz[1] - wz[2]) + al[3] * (wz[2] * wz[0] - wz[3]))) % r_
    zpart = zzw * al[4] % r_
    for i in range(3):
        zpart = zpart * ((sz[i] * beta + gamma + wz[i]) % r_) % r_
    zpart = zpart * ((gamma + wz[3]) % r_) % r_
    rhs = (rhs - zpart - l0 * al[5]) % r_
    if zh * tz % r_ != rhs:
        return False
    # the commitment of the linearisation polynomial
    main = vk[5]
    for i in range(4):
        main = g1_add(main, g1_mul(vk[i], wz[i]))
    main = g1_add(main, g1_mul(vk[4], wz[0] * wz[1] % r_))
    main = g1_add(main, g1_mul(vk[6], dzw))
    r_com = g1_mul(main, smz)
    gpz = al[4]
    for i in range(4):
        gpz = gpz * ((zeta * k[i] * beta + gamma + wz[i]) % r_) % r_
    gpz = (gpz + l0 * al[5]) % r_
    lastp = beta * zzw % r_ * al[4] % r_
    for i in range(3):
        lastp = lastp * ((beta * sz[i] + gamma + wz[i]) % r_) % r_
    r_com = g1_add(g1_add(r_com, g1_mul(cz, gpz)), g1_neg(g1_mul(vk[12], lastp)))
    # the two batched openings, at z (powers of v: t, r, a..d, s_main, s_resc, sigma_0..2) and at z omega (Z, d)
    vp = [pow(v, i, r_) for i in range(13)]
    f = ct[0]
    zp = 1
    for i in range(1, 4):
        zp = zp * zeta_n % r_
        f = g1_add(f, g1_mul(ct[i], zp))
    f = g1_add(f, g1_mul(r_com, v))
    for i in range(4):
        f = g1_add(f, g1_mul(cw[i], vp[2 + i]))
    f = g1_add(g1_add(f, g1_mul(vk[7], vp[6])), g1_mul(vk[8], vp[7]))
    for i in range(3):
        f = g1_add(f, g1_mul(vk[9 + i], vp[8 + i]))
    e = (tz + v * rz) % r_
    for i in range(4):
        e = (e + wz[i] * vp[2 + i]) % r_
    e = (e + smz * vp[6] + srz * vp[7]) % r_
    for i in range(3):
        e = (e + sz[i] * vp[8 + i]) % r_
    f2 = g1_add(g1_mul(cz, vp[11]), g1_mul(cw[3], vp[12]))
    e2 = (zzw * vp[11] + dzw * vp[12]) % r_
    # e(F - E G + z W1 + u (F2 - E2 G + z omega W2), G2) == e(W1 + u W2, [tau] G2)
    lhs = g1_add(g1_add(f, g1_neg(g1_mul((1, 2), e))), g1_mul(w1, zeta))
    lhs2 = g1_add(g1_add(f2, g1_neg(g1_mul((1, 2), e2))), g1_mul(w2, zeta * omega % r_))
    lhs = g1_add(lhs, g1_mul(lhs2, u))
    pair_with_x = g1_neg(g1_add(w1, g1_mul(w2, u)))
    return pairing_product_is_one([(lhs, g2[0]), (pair_with_x, g2[1])])
