"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes loader for oracle/liboracle.so (the CPU restatement of plonkit's prove path, see oracle.cpp).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product package (plonkit_b200/) never does.

Field elements cross this boundary as canonical (non-Montgomery) little-endian u64[4]; points as affine
(x, y) u64[8], (0, 0) = point at infinity.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

R_MOD = 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001
Q_MOD = 0x30644e72e131a029b85045b68181585d97816a916871ca8d3c208c16d87cfd47


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle.cpp", "bn254.hpp", "keccak.hpp")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class Assembly(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_uint64),
        ("num_inputs", ctypes.c_uint64),
        ("nvars", ctypes.c_uint64),
        ("wire_idx", ctypes.c_void_p),
        ("var_values", ctypes.c_void_p),
        ("selectors", ctypes.c_void_p),
    ]


class Assembly2(ctypes.Structure):
    _fields_ = [("base", Assembly), ("gate_type", ctypes.c_void_p)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = ctypes.CDLL(so)
        _LIB.orc_prove.restype = ctypes.c_int64
        _LIB.orc_on_curve.restype = ctypes.c_int
        _LIB.orc_verify_trapdoor.restype = ctypes.c_int
        _LIB.orc_prove2.restype = ctypes.c_int64
        _LIB.orc_verify_trapdoor2.restype = ctypes.c_int
        _LIB.orc_init()
    return _LIB


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ints_to_limbs(vals):
    """list of python ints -> (n, 4) uint64 canonical LE limbs"""
    buf = b"".join(int(v).to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy()


def limbs_to_ints(a):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = a.tobytes()
    return [int.from_bytes(b[32 * i:32 * i + 32], "little") for i in range(a.shape[0])]


def keccak256(data: bytes) -> bytes:
    out = ctypes.create_string_buffer(32)
    lib().orc_keccak256(data, ctypes.c_uint64(len(data)), out)
    return out.raw


def constants():
    out = np.zeros(24, dtype=np.uint64)
    lib().orc_constants(_p(out))
    vals = limbs_to_ints(out.reshape(6, 4))
    return {"fr": {"R": vals[0], "R2": vals[1], "INV": vals[2]}, "fq": {"R": vals[3], "R2": vals[4], "INV": vals[5]}}


def omega(log_n):
    out = np.zeros(4, dtype=np.uint64)
    lib().orc_omega(log_n, _p(out))
    return limbs_to_ints(out)[0]


def ntt(data, inverse=False, coset=False, threads=1):
    a = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4).copy()
    log_n = a.shape[0].bit_length() - 1
    assert 1 << log_n == a.shape[0]
    lib().orc_ntt(_p(a), log_n, int(inverse), int(coset), threads)
    return a


def naive_dft(data):
    a = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4)
    log_n = a.shape[0].bit_length() - 1
    out = np.zeros_like(a)
    lib().orc_naive_dft(_p(a), _p(out), log_n)
    return out


def lde4(coeffs, threads=1):
    a = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
    log_n = a.shape[0].bit_length() - 1
    out = np.zeros((4 * a.shape[0], 4), dtype=np.uint64)
    lib().orc_lde4(_p(a), log_n, _p(out), threads)
    return out


def poly_op(op, data, z=None):
    """bellman polynomial primitives restated serially (oracle.cpp orc_poly_op): op in evaluate_at | divide_by_linear |
    shifted_grand_product | batch_inversion"""
    code = {"evaluate_at": 0, "divide_by_linear": 1, "shifted_grand_product": 2, "batch_inversion": 3}[op]
    a = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4)
    zz = np.ascontiguousarray(z if z is not None else np.zeros(4), dtype=np.uint64).reshape(4)
    out = np.zeros(4, dtype=np.uint64) if code == 0 else np.zeros_like(a)
    lib().orc_poly_op(code, _p(a), ctypes.c_uint64(a.shape[0]), _p(zz), _p(out))
    return out


def msm(scalars, bases, threads=1):
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    assert s.shape[0] == b.shape[0]
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_msm(_p(s), _p(b), ctypes.c_uint64(s.shape[0]), _p(out), threads)
    return out


def msm_naive(scalars, bases):
    s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_msm_naive(_p(s), _p(b), ctypes.c_uint64(s.shape[0]), _p(out))
    return out


def on_curve(points):
    b = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 8)
    return bool(lib().orc_on_curve(_p(b), ctypes.c_uint64(b.shape[0])))


def g1_mul(point, k):
    p = np.ascontiguousarray(point, dtype=np.uint64).reshape(8)
    kk = ints_to_limbs([k % R_MOD])
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_mul(_p(p), _p(kk), _p(out))
    return out


def g1_mul_fixed(point, scalars, threads=1):
    """[k * P for k in scalars] -> (n, 8)"""
    p = np.ascontiguousarray(point, dtype=np.uint64).reshape(8)
    kk = ints_to_limbs([int(k) % R_MOD for k in scalars])
    out = np.zeros((kk.shape[0], 8), dtype=np.uint64)
    lib().orc_g1_mul_fixed(_p(p), _p(kk), ctypes.c_uint64(kk.shape[0]), _p(out), threads)
    return out


def g1_add(p, q):
    a = np.ascontiguousarray(p, dtype=np.uint64).reshape(8)
    b = np.ascontiguousarray(q, dtype=np.uint64).reshape(8)
    out = np.zeros(8, dtype=np.uint64)
    lib().orc_g1_add(_p(a), _p(b), _p(out))
    return out


def fr_mul(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(a)
    lib().orc_fr_mul(_p(a), _p(b), _p(out), ctypes.c_uint64(a.shape[0]))
    return out


def fq_mul(a, b):
    a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
    b = np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(a)
    lib().orc_fq_mul(_p(a), _p(b), _p(out), ctypes.c_uint64(a.shape[0]))
    return out


def srs_gen(n, tau=42, threads=1):
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().orc_srs_gen(ctypes.c_uint64(n), ctypes.c_uint64(tau), _p(out), threads)
    return out


def ec_intt(bases, threads=1):
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    log_n = b.shape[0].bit_length() - 1
    out = np.zeros_like(b)
    lib().orc_ec_intt(_p(b), log_n, _p(out), threads)
    return out


def _assembly(n, num_inputs, wire_idx, var_values, selectors):
    wire_idx = np.ascontiguousarray(wire_idx, dtype=np.uint32).reshape(4, n)
    selectors = np.ascontiguousarray(selectors, dtype=np.uint64).reshape(7, n, 4)
    keep = [wire_idx, selectors]
    a = Assembly()
    a.n = n
    a.num_inputs = num_inputs
    a.wire_idx = wire_idx.ctypes.data
    a.selectors = selectors.ctypes.data
    if var_values is not None:
        var_values = np.ascontiguousarray(var_values, dtype=np.uint64).reshape(-1, 4)
        a.nvars = var_values.shape[0]
        a.var_values = var_values.ctypes.data
        keep.append(var_values)
    else:
        a.nvars = int(wire_idx.max()) + 1
        a.var_values = None
    return a, keep


def setup_commitments(n, num_inputs, wire_idx, selectors, srs, nvars=None, threads=1, want_sigma=False):
    a, keep = _assembly(n, num_inputs, wire_idx, None, selectors)
    if nvars is not None:
        a.nvars = nvars
    srs = np.ascontiguousarray(srs, dtype=np.uint64).reshape(-1, 8)
    assert srs.shape[0] >= n
    out = np.zeros((11, 8), dtype=np.uint64)
    sig = np.zeros((4, n, 4), dtype=np.uint64) if want_sigma else None
    lib().orc_setup_commitments(ctypes.byref(a), _p(srs), _p(out), _p(sig) if want_sigma else None, threads)
    return (out, sig) if want_sigma else out


def prove(n, num_inputs, wire_idx, var_values, selectors, srs, threads=1, want_challenges=False):
    """-> proof.bin bytes"""
    a, keep = _assembly(n, num_inputs, wire_idx, var_values, selectors)
    srs = np.ascontiguousarray(srs, dtype=np.uint64).reshape(-1, 8)
    assert srs.shape[0] >= n
    buf = ctypes.create_string_buffer(16 + 32 * num_inputs + 1096 + 64)
    ch = np.zeros((5, 4), dtype=np.uint64)
    ln = lib().orc_prove(ctypes.byref(a), _p(srs), buf, _p(ch), threads)
    if ln < 0:
        raise RuntimeError({-1: "circuit not satisfied", -2: "quotient is not a polynomial"}.get(ln, "oracle error %d" % ln))
    proof = buf.raw[:ln]
    return (proof, limbs_to_ints(ch)) if want_challenges else proof


def verify_trapdoor(proof_bytes: bytes, vk_commitments, tau=42) -> bool:
    """contrib/template.sol verifier with the pairing check replaced by the G1 identity A + tau*B == 0 (known tau)."""
    vk = np.ascontiguousarray(vk_commitments, dtype=np.uint64).reshape(11, 8)
    rc = lib().orc_verify_trapdoor(proof_bytes, ctypes.c_uint64(len(proof_bytes)), _p(vk), ctypes.c_uint64(tau))
    if rc < 0:
        raise ValueError("malformed proof / vk")
    return rc == 1


# ---- prover with gate selectors and the Rescue x^5 custom gate (oracle.cpp orc_prove2; PARITY UNPINNED, see there)
def _assembly2(n, num_inputs, wire_idx, var_values, selectors, gate_type):
    a, keep = _assembly(n, num_inputs, wire_idx, var_values, selectors)
    gt = np.ascontiguousarray(gate_type, dtype=np.uint8).reshape(n)
    a2 = Assembly2()
    a2.base = a
    a2.gate_type = gt.ctypes.data
    return a2, keep + [gt]


def setup_commitments2(n, num_inputs, wire_idx, selectors, gate_type, srs, nvars=None, threads=1):
    a2, keep = _assembly2(n, num_inputs, wire_idx, None, selectors, gate_type)
    if nvars is not None:
        a2.base.nvars = nvars
    srs = np.ascontiguousarray(srs, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros((13, 8), dtype=np.uint64)
    lib().orc_setup_commitments2(ctypes.byref(a2), _p(srs), _p(out), threads)
    return out


def prove2(n, num_inputs, wire_idx, var_values, selectors, gate_type, srs, threads=1):
    a2, keep = _assembly2(n, num_inputs, wire_idx, var_values, selectors, gate_type)
    srs = np.ascontiguousarray(srs, dtype=np.uint64).reshape(-1, 8)
    assert srs.shape[0] >= n
    buf = ctypes.create_string_buffer(16 + 32 * num_inputs + 1096 + 72 + 64)
    ch = np.zeros((5, 4), dtype=np.uint64)
    ln = lib().orc_prove2(ctypes.byref(a2), _p(srs), buf, _p(ch), threads)
    if ln < 0:
        raise RuntimeError({-1: "circuit not satisfied", -2: "quotient is not a polynomial"}.get(ln, "oracle error %d" % ln))
    return buf.raw[:ln]


def verify_trapdoor2(proof_bytes: bytes, vk_commitments, tau=42) -> bool:
    vk = np.ascontiguousarray(vk_commitments, dtype=np.uint64).reshape(13, 8)
    rc = lib().orc_verify_trapdoor2(proof_bytes, ctypes.c_uint64(len(proof_bytes)), _p(vk), ctypes.c_uint64(tau))
    if rc < 0:
        raise ValueError("malformed proof / vk")
    return rc == 1


def last_timings():
    """(setup_s, prove_s, setup_lde_s) of the last prove() call: setup polynomials (reference: once, in
    prepare_setup_for_prover), the per-call work of SetupForProver::prove, and the share of the latter spent on the 11
    setup-polynomial LDEs that the reference recomputes on every call (precomputations = None, src/plonk.rs:156)"""
    out = (ctypes.c_double * 3)()
    lib().orc_last_timings(out)
    return out[0], out[1], out[2]
