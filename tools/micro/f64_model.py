import random
r = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
q = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47
L = 51
M = (1 << L) - 1
def rz53(x):
    if x == 0: return 0
    s = -1 if x < 0 else 1
    x = abs(x)
    n = x.bit_length()
    if n <= 53: return s * x
    sh = n - 53
    return s * ((x >> sh) << sh)
def exact53(x):
    assert rz53(x) == x, hex(x)
    return x
def fma_rz(a, b, c): return rz53(a * b + c)
def limbs(x): return [(x >> (L * i)) & M for i in range(5)]
C0 = 1 << 104
ANCH = 1 << 52
def mul(a, b, p):
    pinv = (-pow(p, -1, 1 << L)) % (1 << L)
    A, B, Pl = limbs(a), limbs(b), limbs(p)
    lo = [0] * 10  # integer sums of e_k
    hi = [0] * 10  # chain (s_n - C0) / 2^52
    for c in range(9):
        s = C0
        for i in range(5):
            j = c - i
            if j < 0 or j > 4: continue
            s2 = fma_rz(A[i], B[j], s)
            assert C0 <= s2 < 2 * C0
            d = exact53(s - s2 + ANCH)
            l = fma_rz(A[i], B[j], d)
            assert ANCH <= l < 2 * ANCH
            lo[c] += l - ANCH
            s = s2
        hi[c] = (s - C0) >> 52
    # reduction; separate chains per column for q*p
    sq = [C0] * 10
    carry = 0
    for i in range(5):
        t = lo[i] + carry + (2 * (hi[i - 1] + ((sq[i - 1] - C0) >> 52)) if i > 0 else 0)
        qi = ((t & M) * pinv) & M
        for j in range(5):
            c = i + j
            s2 = fma_rz(qi, Pl[j], sq[c])
            assert C0 <= s2 < 2 * C0, (i, j)
            d = exact53(sq[c] - s2 + ANCH)
            l = fma_rz(qi, Pl[j], d)
            assert ANCH <= l < 2 * ANCH
            lo[c] += l - ANCH
            sq[c] = s2
        t = lo[i] + carry + (2 * (hi[i - 1] + ((sq[i - 1] - C0) >> 52)) if i > 0 else 0)
        assert t & M == 0
        carry = t >> L
    res = 0
    for c in range(5, 10):
        t = lo[c] + carry + 2 * (hi[c - 1] + ((sq[c - 1] - C0) >> 52))
        assert t < 1 << 63
        res += (t & M) << (L * (c - 5))
        carry = t >> L
    assert carry == 0
    # res = a*b*2^-255 mod p (< 2p); halve
    assert res < 2 * p
    if res & 1: res += p
    res >>= 1
    if res >= p: res -= p
    return res
for p in (r, q):
    Rinv = pow(1 << 256, -1, p)
    tests = [(0, 0), (p - 1, p - 1), (1, p - 1), ((1 << 254) % p, p - 1)]
    for _ in range(20000): tests.append((random.randrange(p), random.randrange(p)))
    # adversarial: limbs all ones
    for _ in range(2000):
        a = sum(random.choice([0, M, M - 1, 1]) << (L * i) for i in range(5)) % p
        b = sum(random.choice([0, M, M - 1, 1]) << (L * i) for i in range(5)) % p
        tests.append((a, b))
    for a, b in tests:
        assert mul(a, b, p) == a * b * Rinv % p
    print("ok", hex(p)[:8], [hex(x) for x in limbs(p)], sum(limbs(p)) < (1 << 53))
