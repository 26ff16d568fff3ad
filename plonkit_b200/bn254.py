"""BN254 constants and host-side (Python int) helpers for marshalling across the C ABI.

Constants: contrib/template.sol:7-8 (q, r), :67-69 (G = (1, 2)); SURVEY.md App. C.
Field elements cross the C ABI as canonical little-endian u64[4]; affine points as u64[8] = x || y with
(0, 0) standing for the point at infinity (it is not on y^2 = x^3 + 3).
"""
import numpy as np

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001  # Fr
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47  # Fq
TWO_ADICITY = 28
MULT_GEN = 7
NON_RESIDUES = (1, 5, 7, 10)  # k_i of the copy-permutation argument (vk.bin bytes 752-847)


def root_of_unity(log_n: int) -> int:
    """omega_N = g^(2^(28-log_n)), g = 7^((r-1)/2^28)  """
    assert 0 <= log_n <= TWO_ADICITY
    g = pow(MULT_GEN, (R_MOD - 1) >> TWO_ADICITY, R_MOD)
    return pow(g, 1 << (TWO_ADICITY - log_n), R_MOD)


def ints_to_limbs(vals) -> np.ndarray:
    """iterable of ints (already reduced) -> (n, 4) uint64, canonical little-endian limbs"""
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def limbs_to_ints(a) -> list:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = a.tobytes()
    return [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(a.shape[0])]


def be_bytes_to_limbs(buf: bytes, n_words: int) -> np.ndarray:
    """n_words consecutive 32-byte big-endian integers -> (n_words, 4) uint64 LE limbs"""
    a = np.frombuffer(buf, dtype=np.uint8, count=32 * n_words).reshape(n_words, 32)
    return np.ascontiguousarray(a[:, ::-1]).view(np.uint64).reshape(n_words, 4).copy()


def limbs_to_be_bytes(a) -> bytes:
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = a.view(np.uint8).reshape(-1, 32)[:, ::-1]
    return np.ascontiguousarray(b).tobytes()
