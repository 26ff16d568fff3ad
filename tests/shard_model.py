"""CPU model (Python integers) of the algebra behind the sharded prover's quotient step (plonkit_b200/csrc/dist_prover.cu):
which part of the 4n coset domain a rank owns, how it evaluates a polynomial there without seeing the other ranks
(coefficient fold + a smaller NTT), and how the size-4n inverse transform is split into block-local stages, one
all-to-all and log2(G) cross stages.  tests/test_shard_math.py checks it against the oracle's plain transforms, in one
process for 1/2/4/8 ranks and over gloo with 2 processes.  The CUDA implementation is checked separately (-m gpu)."""
from plonkit_b200.bn254 import R_MOD

GEN7 = 7
BREV2 = (0, 2, 1, 3)


def brev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def ntt_natural(a, w):
    """size-len(a) transform of a with the primitive root w, natural order in and out (O(n log n), recursive)"""
    n = len(a)
    if n == 1:
        return list(a)
    ev = ntt_natural(a[0::2], w * w % R_MOD)
    od = ntt_natural(a[1::2], w * w % R_MOD)
    out = [0] * n
    x = 1
    for k in range(n // 2):
        t = x * od[k] % R_MOD
        out[k] = (ev[k] + t) % R_MOD
        out[k + n // 2] = (ev[k] - t) % R_MOD
        x = x * w % R_MOD
    return out


class Plan:
    """what rank `rank` of `G` owns for a circuit domain of size n = 2^log_n (dist_setup_create)"""

    def __init__(self, log_n, G, rank, omega_n, omega_4n):
        self.log_n, self.G, self.rank = log_n, G, rank
        self.n = 1 << log_n
        self.m = 4 * self.n // G                       # range of the slot layout: [rank * m, (rank + 1) * m)
        self.sub = max(0, G.bit_length() - 1 - 2)      # log2 of the split of one coset
        self.parts = 1 if G >= 4 else 4 // G
        self.nf = self.n >> self.sub
        self.wn, self.w4 = omega_n, omega_4n
        self.shifts = []
        for p in range(self.parts):
            pos = rank * self.m + p * self.nf
            sl, q = pos >> log_n, (pos & (self.n - 1)) // self.nf
            jl = brev(q, self.sub)
            self.shifts.append(GEN7 * pow(self.w4, BREV2[sl], R_MOD) * pow(self.wn, jl, R_MOD) % R_MOD)

    def range_lde(self, coef):
        """evaluations of the polynomial `coef` (n coefficients) on this rank's range, in slot-layout order"""
        F = 1 << self.sub
        out = []
        w_nf = pow(self.wn, F, R_MOD)
        for c in self.shifts:
            kappa = pow(c, self.nf, R_MOD)
            b, ci = [], 1
            for i in range(self.nf):
                acc = 0
                for u in reversed(range(F)):
                    acc = (acc * kappa + coef[i + u * self.nf]) % R_MOD
                b.append(acc * ci % R_MOD)
                ci = ci * c % R_MOD
            nat = ntt_natural(b, w_nf)
            bits = self.log_n - self.sub
            out += [nat[brev(p, bits)] for p in range(self.nf)]
        return out


def slot_layout_of(natural_coset_values, log_n):
    """4n values on 7 H_4n in natural order (index J = 4 j + s') -> the prover's slot layout (= 4n-point bit reversal)"""
    n = 1 << log_n
    out = [0] * (4 * n)
    for s in range(4):
        for p in range(n):
            out[s * n + p] = natural_coset_values[4 * brev(p, log_n) + BREV2[s]]
    return out


def inverse_local_stages(block, log_total, w_total_inv):
    """the block-local decimation-in-time stages (bits 0 .. log2(len(block)) - 1) of a size-2^log_total inverse transform
    on one aligned block of the bit-reversed input"""
    a = list(block)
    m = len(a)
    s = 0
    while (1 << s) < m:
        half = 1 << s
        for i0 in range(m):
            if i0 & half:
                continue
            low = i0 & (half - 1)
            e = low << (log_total - 1 - s)
            t = a[i0 | half] * pow(w_total_inv, e, R_MOD) % R_MOD
            a[i0], a[i0 | half] = (a[i0] + t) % R_MOD, (a[i0] - t) % R_MOD
        s += 1
    return a


def inverse_cross_stages(recv, G, rank, log_total, w_total_inv):
    """recv[c][k'] = element rank * per + k' of rank c's block after its local stages.  Returns out[c][k'] = coefficient
    c * m + rank * per + k' of the inverse coset transform (scaled by 7^-index / 2^log_total)."""
    M = 1 << log_total
    m = M // G
    per = m // G
    log_m = m.bit_length() - 1
    inv_m = pow(M, R_MOD - 2, R_MOD)
    g7inv = pow(GEN7, R_MOD - 2, R_MOD)
    out = [[0] * per for _ in range(G)]
    for kk in range(per):
        v = [recv[c][kk] for c in range(G)]
        k = rank * per + kk
        u = 0
        while (1 << u) < G:
            s = log_m + u
            for c in range(G):
                if c & (1 << u):
                    continue
                low = k + (c & ((1 << u) - 1)) * m
                e = low << (log_total - 1 - s)
                t = v[c | (1 << u)] * pow(w_total_inv, e, R_MOD) % R_MOD
                v[c], v[c | (1 << u)] = (v[c] + t) % R_MOD, (v[c] - t) % R_MOD
            u += 1
        for c in range(G):
            idx = c * m + k
            out[c][kk] = v[c] * pow(g7inv, idx, R_MOD) % R_MOD * inv_m % R_MOD
    return out
