"""ctypes binding of the C-ABI CUDA library (include/plonkit_b200.h).

There is no CPU fallback: if libplonkit_b200.so is missing, or no CUDA device can be opened, every compute call
raises.  Loading the library needs no GPU (tests check the exported symbols on CPU-only machines).
"""
import ctypes
import os
import weakref

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplonkit_b200.so")

PK_OK = 0
ERRORS = {
    1: "AssignmentMissing",
    2: "PolynomialDegreeTooLarge",
    3: "DivisionByZero",
    4: "Unsatisfiable",
    5: "CudaError",
    6: "InvalidArgument",
}
FMT_CANONICAL, FMT_MONTGOMERY = 0, 1

EXPORTS = [
    "pk_create", "pk_destroy", "pk_last_error", "pk_constants", "pk_srs_load_g1", "pk_srs_load_g1_lagrange", "pk_setup_use_lagrange",
    "pk_srs_gen", "pk_setup_create_gated", "pk_setup_commitments_gated", "pk_ntt", "pk_lde4",
    "pk_msm_g1", "pk_ec_intt_g1", "pk_setup_create", "pk_setup_destroy", "pk_setup_commitments", "pk_witness_upload",
    "pk_prove", "pk_profile_enable", "pk_profile_reset", "pk_profile_get", "pk_bench_ntt", "pk_bench_msm",
    "pk_bench_fieldmul", "pk_bench_msm_pattern", "pk_timer_begin", "pk_timer_end", "pk_g1_sum", "pk_dev_fr_convert", "pk_dev_ntt_rows",
    "pk_dev_twiddle", "pk_poly_evaluate_at", "pk_poly_divide_by_linear", "pk_poly_shifted_grand_product",
    "pk_poly_batch_inversion", "pk_poly_pointwise", "pk_dev_ec_from_affine", "pk_dev_ec_ntt_rows", "pk_dev_ec_twiddle", "pk_dev_ec_to_affine",
    "pk_comm_group_create", "pk_comm_group_destroy", "pk_comm_attach_group", "pk_comm_nccl_unique_id", "pk_comm_attach_nccl",
    "pk_dist_setup_create", "pk_dist_setup_destroy", "pk_dist_setup_commitments", "pk_dist_witness_upload", "pk_dist_prove",
]


class SynthesisError(RuntimeError):
    """Mirror of bellman's SynthesisError as surfaced by SetupForProver::prove (src/plonk.rs:136)."""

    def __init__(self, code, message):
        super().__init__("%s: %s" % (ERRORS.get(code, "Error %d" % code), message))
        self.code = code
        self.kind = ERRORS.get(code, "Unknown")


class LibraryMissing(RuntimeError):
    pass


class PkAssembly(ctypes.Structure):
    _fields_ = [("n", ctypes.c_uint64), ("num_inputs", ctypes.c_uint64), ("nvars", ctypes.c_uint64),
                ("wire_idx", ctypes.c_void_p), ("selectors", ctypes.c_void_p)]


class PkProof(ctypes.Structure):
    _fields_ = [
        ("n", ctypes.c_uint64), ("num_inputs", ctypes.c_uint64),
        ("wire_commitments", ctypes.c_uint64 * 32),
        ("grand_product_commitment", ctypes.c_uint64 * 8),
        ("quotient_poly_commitments", ctypes.c_uint64 * 32),
        ("wire_values_at_z", ctypes.c_uint64 * 16),
        ("wire_values_at_z_omega", ctypes.c_uint64 * 4),
        ("grand_product_at_z_omega", ctypes.c_uint64 * 4),
        ("quotient_polynomial_at_z", ctypes.c_uint64 * 4),
        ("linearization_polynomial_at_z", ctypes.c_uint64 * 4),
        ("permutation_polynomials_at_z", ctypes.c_uint64 * 12),
        ("opening_at_z_proof", ctypes.c_uint64 * 8),
        ("opening_at_z_omega_proof", ctypes.c_uint64 * 8),
        ("challenges", ctypes.c_uint64 * 20),
        ("num_gate_selectors", ctypes.c_uint64),
        ("gate_selectors_at_z", ctypes.c_uint64 * 8),
    ]


class PkAssemblyGated(ctypes.Structure):
    _fields_ = [("base", PkAssembly), ("gate_type", ctypes.c_void_p)]


class PkProfile(ctypes.Structure):
    _fields_ = [
        ("kernel_launches", ctypes.c_uint64), ("msm_accum_launches", ctypes.c_uint64), ("msm_accum_ms", ctypes.c_double),
        ("msm_accum_points", ctypes.c_uint64), ("ntt_launches", ctypes.c_uint64), ("ntt_ms", ctypes.c_double),
        ("ntt_elements", ctypes.c_uint64), ("phase_ms", ctypes.c_double * 8),
        ("comm_ms", ctypes.c_double * 3), ("comm_bytes", ctypes.c_uint64 * 3),
    ]


_lib = None


def load():
    """Loads the shared library (no GPU needed for that) and declares the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise LibraryMissing("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                             "(there is no CPU fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    vp, u64, i32, u32 = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint32
    lib.pk_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.pk_destroy.argtypes = [vp]
    lib.pk_destroy.restype = None
    lib.pk_last_error.argtypes = [vp]
    lib.pk_last_error.restype = ctypes.c_char_p
    lib.pk_constants.argtypes = [vp]
    lib.pk_constants.restype = None
    lib.pk_srs_load_g1.argtypes = [vp, vp, u64, i32]
    lib.pk_srs_load_g1_lagrange.argtypes = [vp, vp, u64, i32]
    lib.pk_setup_use_lagrange.argtypes = [vp, vp, i32]
    lib.pk_srs_gen.argtypes = [vp, u64, u64, vp]
    lib.pk_ntt.argtypes = [vp, vp, u32, i32, i32, i32]
    lib.pk_lde4.argtypes = [vp, vp, u32, vp, i32, i32]
    lib.pk_msm_g1.argtypes = [vp, vp, u64, u64, vp, ctypes.POINTER(i32), i32]
    lib.pk_ec_intt_g1.argtypes = [vp, u32, vp]
    lib.pk_setup_create.argtypes = [vp, ctypes.POINTER(PkAssembly), ctypes.POINTER(vp)]
    lib.pk_setup_destroy.argtypes = [vp]
    lib.pk_setup_destroy.restype = None
    lib.pk_setup_create_gated.argtypes = [vp, ctypes.POINTER(PkAssemblyGated), ctypes.POINTER(vp)]
    lib.pk_setup_commitments_gated.argtypes = [vp, vp, vp]
    lib.pk_setup_commitments.argtypes = [vp, vp, vp]
    lib.pk_witness_upload.argtypes = [vp, vp, vp, u64]
    lib.pk_prove.argtypes = [vp, vp, vp, u64, ctypes.POINTER(PkProof), vp]
    lib.pk_profile_enable.argtypes = [vp, i32]
    lib.pk_profile_enable.restype = None
    lib.pk_profile_reset.argtypes = [vp]
    lib.pk_profile_reset.restype = None
    lib.pk_profile_get.argtypes = [vp, ctypes.POINTER(PkProfile)]
    lib.pk_profile_get.restype = None
    lib.pk_bench_ntt.argtypes = [vp, u32, i32, ctypes.POINTER(ctypes.c_double)]
    lib.pk_bench_msm.argtypes = [vp, u64, i32, ctypes.POINTER(ctypes.c_double)]
    lib.pk_bench_fieldmul.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_double)]
    lib.pk_bench_msm_pattern.argtypes = [vp, u64, i32, i32, ctypes.POINTER(ctypes.c_double)]
    lib.pk_g1_sum.argtypes = [vp, u64, vp]
    lib.pk_dev_fr_convert.argtypes = [vp, vp, u64, i32]
    lib.pk_dev_ntt_rows.argtypes = [vp, vp, u32, u64, i32]
    lib.pk_dev_twiddle.argtypes = [vp, vp, u64, u64, u32, u64, i32]
    lib.pk_dev_ec_from_affine.argtypes = [vp, vp, vp, u64]
    lib.pk_dev_ec_ntt_rows.argtypes = [vp, vp, u32, u64, i32]
    lib.pk_dev_ec_twiddle.argtypes = [vp, vp, u64, u64, u32, u64, i32]
    lib.pk_dev_ec_to_affine.argtypes = [vp, vp, vp, u64, u32]
    lib.pk_poly_evaluate_at.argtypes = [vp, vp, u64, vp, vp]
    lib.pk_poly_divide_by_linear.argtypes = [vp, vp, u64, vp, vp]
    lib.pk_poly_shifted_grand_product.argtypes = [vp, vp, u64, vp]
    lib.pk_poly_batch_inversion.argtypes = [vp, vp, u64]
    lib.pk_poly_pointwise.argtypes = [vp, i32, vp, vp, vp, u64, vp]
    lib.pk_comm_group_create.argtypes = [i32, ctypes.POINTER(vp)]
    lib.pk_comm_group_destroy.argtypes = [vp]
    lib.pk_comm_group_destroy.restype = None
    lib.pk_comm_attach_group.argtypes = [vp, vp, i32]
    lib.pk_comm_nccl_unique_id.argtypes = [vp]
    lib.pk_comm_attach_nccl.argtypes = [vp, vp, i32, i32]
    lib.pk_dist_setup_create.argtypes = [vp, ctypes.POINTER(PkAssembly), ctypes.POINTER(vp)]
    lib.pk_dist_setup_destroy.argtypes = [vp]
    lib.pk_dist_setup_destroy.restype = None
    lib.pk_dist_setup_commitments.argtypes = [vp, vp, vp]
    lib.pk_dist_witness_upload.argtypes = [vp, vp, vp, u64]
    lib.pk_dist_prove.argtypes = [vp, vp, vp, u64, ctypes.POINTER(PkProof), vp]
    lib.pk_timer_begin.argtypes = [vp]
    lib.pk_timer_end.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    _lib = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Context:
    """One CUDA device context (pk_ctx).  Stands in for bellman's Worker (src/plonk.rs:41,47,183)."""

    def __init__(self, device=0):
        self._lib = load()
        h = ctypes.c_void_p()
        rc = self._lib.pk_create(device, ctypes.byref(h))
        if rc != PK_OK:
            raise SynthesisError(rc, "cannot open CUDA device %d (no GPU? there is no CPU fallback)" % device)
        self._h = h
        self.device = device
        self.srs_size = 0
        self.srs_tag = None  # which key is resident (plonk.SetupForProver._ensure_srs); None after a direct load
        self.lagrange_tag = None
        self._children = weakref.WeakSet()  # device-side objects that must be released before the context

    def close(self):
        if getattr(self, "_h", None):
            for child in list(self._children):
                child.close()
            self._lib.pk_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != PK_OK:
            raise SynthesisError(rc, self._lib.pk_last_error(self._h).decode())

    # ---- SRS
    def srs_load_g1(self, bases, window_bits=0, tag=None):
        b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        self.srs_tag = None
        self.lagrange_tag = None  # a Lagrange key sharing the old tables' working set goes with them
        self._check(self._lib.pk_srs_load_g1(self._h, _ptr(b), b.shape[0], window_bits))
        self.srs_size = b.shape[0]
        self.srs_tag = tag

    def srs_load_g1_lagrange(self, bases, window_bits=0, tag=None):
        """Lagrange-form key of a power-of-two domain, resident beside the monomial one; bases=None unloads it."""
        self.lagrange_tag = None
        if bases is None:
            self._check(self._lib.pk_srs_load_g1_lagrange(self._h, None, 0, 0))
            return
        b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        self._check(self._lib.pk_srs_load_g1_lagrange(self._h, _ptr(b), b.shape[0], window_bits))
        self.lagrange_tag = tag

    def srs_gen(self, n, tau=42):
        out = np.zeros((n, 8), dtype=np.uint64)
        self._check(self._lib.pk_srs_gen(self._h, n, tau, _ptr(out)))
        return out

    # ---- communicator of the sharded prover (this context = one rank)
    def attach_group(self, group, rank):
        self._check(self._lib.pk_comm_attach_group(self._h, group._h, rank))
        self.rank, self.world = rank, group.world

    def attach_nccl(self, unique_id: bytes, rank, world):
        buf = ctypes.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.pk_comm_attach_nccl(self._h, buf, rank, world))
        self.rank, self.world = rank, world

    # ---- primitives
    def ntt(self, data, inverse=False, coset=False, fmt=FMT_CANONICAL):
        a = np.ascontiguousarray(data, dtype=np.uint64).reshape(-1, 4).copy()
        log_n = a.shape[0].bit_length() - 1
        if 1 << log_n != a.shape[0]:
            raise SynthesisError(6, "length is not a power of two")
        self._check(self._lib.pk_ntt(self._h, _ptr(a), log_n, int(inverse), int(coset), fmt))
        return a

    def lde4(self, coeffs, bitreversed=False, fmt=FMT_CANONICAL):
        a = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        log_n = a.shape[0].bit_length() - 1
        if 1 << log_n != a.shape[0]:
            raise SynthesisError(6, "length is not a power of two")
        out = np.zeros((4 * a.shape[0], 4), dtype=np.uint64)
        self._check(self._lib.pk_lde4(self._h, _ptr(a), log_n, _ptr(out), int(bitreversed), fmt))
        return out

    # ---- polynomial primitives (bellman's evaluate_at / divide_single / calculate_shifted_grand_product / batch_inversion)
    def poly_evaluate_at(self, coeffs, z):
        c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        zz = np.ascontiguousarray(z, dtype=np.uint64).reshape(4)
        out = np.zeros(4, dtype=np.uint64)
        self._check(self._lib.pk_poly_evaluate_at(self._h, _ptr(c) if c.shape[0] else None, c.shape[0], _ptr(zz), _ptr(out)))
        return out

    def poly_divide_by_linear(self, coeffs, z):
        c = np.ascontiguousarray(coeffs, dtype=np.uint64).reshape(-1, 4)
        zz = np.ascontiguousarray(z, dtype=np.uint64).reshape(4)
        out = np.zeros_like(c)
        self._check(self._lib.pk_poly_divide_by_linear(self._h, _ptr(c) if c.shape[0] else None, c.shape[0], _ptr(zz),
                                                       _ptr(out) if c.shape[0] else None))
        return out

    def poly_shifted_grand_product(self, values):
        v = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros_like(v)
        self._check(self._lib.pk_poly_shifted_grand_product(self._h, _ptr(v) if v.shape[0] else None, v.shape[0],
                                                            _ptr(out) if v.shape[0] else None))
        return out

    def poly_batch_inversion(self, values):
        v = np.ascontiguousarray(values, dtype=np.uint64).reshape(-1, 4).copy()
        self._check(self._lib.pk_poly_batch_inversion(self._h, _ptr(v) if v.shape[0] else None, v.shape[0]))
        return v

    POINTWISE_OPS = {"add_assign_scaled": 0, "mul_assign": 1, "scale": 2, "add_constant": 3, "distribute_powers": 4}

    def poly_pointwise(self, op, a, b=None, scalar=None):
        """bellman's pointwise Polynomial operations (see pk_poly_pointwise); canonical limbs in and out"""
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(-1, 4)
        b = None if b is None else np.ascontiguousarray(b, dtype=np.uint64).reshape(-1, 4)
        sc = None if scalar is None else np.ascontiguousarray(scalar, dtype=np.uint64).reshape(4)
        out = np.zeros_like(a)
        n = a.shape[0]
        self._check(self._lib.pk_poly_pointwise(self._h, self.POINTWISE_OPS[op], _ptr(a) if n else None, _ptr(b) if (b is not None and n) else None,
                                                _ptr(sc), n, _ptr(out) if n else None))
        return out

    def msm_g1(self, scalars, base_offset=0, fmt=FMT_CANONICAL):
        s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros(8, dtype=np.uint64)
        inf = ctypes.c_int(0)
        self._check(self._lib.pk_msm_g1(self._h, _ptr(s) if s.shape[0] else None, s.shape[0], base_offset, _ptr(out),
                                        ctypes.byref(inf), fmt))
        return out

    def ec_intt_g1(self, log_n):
        out = np.zeros((1 << log_n, 8), dtype=np.uint64)
        self._check(self._lib.pk_ec_intt_g1(self._h, log_n, _ptr(out)))
        return out

    # ---- device-pointer primitives (caller-owned device memory; `ptr` is a raw CUDA device address)
    def dev_fr_convert(self, ptr, n, to_mont):
        self._check(self._lib.pk_dev_fr_convert(self._h, ctypes.c_void_p(ptr), n, int(to_mont)))

    def dev_ntt_rows(self, ptr, log_len, rows, inverse=False):
        self._check(self._lib.pk_dev_ntt_rows(self._h, ctypes.c_void_p(ptr), log_len, rows, int(inverse)))

    def dev_twiddle(self, ptr, rows, cols, log_total, row0, inverse=False):
        self._check(self._lib.pk_dev_twiddle(self._h, ctypes.c_void_p(ptr), rows, cols, log_total, row0, int(inverse)))

    def dev_ec_from_affine(self, ptr_affine, ptr_xyzz, n):
        self._check(self._lib.pk_dev_ec_from_affine(self._h, ctypes.c_void_p(ptr_affine), ctypes.c_void_p(ptr_xyzz), n))

    def dev_ec_ntt_rows(self, ptr, log_len, rows, inverse=False):
        self._check(self._lib.pk_dev_ec_ntt_rows(self._h, ctypes.c_void_p(ptr), log_len, rows, int(inverse)))

    def dev_ec_twiddle(self, ptr, rows, cols, log_total, row0, inverse=False):
        self._check(self._lib.pk_dev_ec_twiddle(self._h, ctypes.c_void_p(ptr), rows, cols, log_total, row0, int(inverse)))

    def dev_ec_to_affine(self, ptr_xyzz, ptr_affine, n, log_scale):
        self._check(self._lib.pk_dev_ec_to_affine(self._h, ctypes.c_void_p(ptr_xyzz), ctypes.c_void_p(ptr_affine), n, log_scale))

    # ---- profiling / micro-benchmarks
    def profile_enable(self, on=True):
        self._lib.pk_profile_enable(self._h, int(on))

    def profile_reset(self):
        self._lib.pk_profile_reset(self._h)

    def profile(self):
        p = PkProfile()
        self._lib.pk_profile_get(self._h, ctypes.byref(p))
        d = {k: getattr(p, k) for k, _ in PkProfile._fields_ if k not in ("phase_ms", "comm_ms", "comm_bytes")}
        d["phase_ms"] = list(p.phase_ms)
        d["comm_ms"], d["comm_bytes"] = list(p.comm_ms), list(p.comm_bytes)
        return d

    def timer_begin(self):
        self._check(self._lib.pk_timer_begin(self._h))

    def timer_end(self):
        ms = ctypes.c_double()
        self._check(self._lib.pk_timer_end(self._h, ctypes.byref(ms)))
        return ms.value

    def bench_ntt(self, log_n, iters=10):
        ms = ctypes.c_double()
        self._check(self._lib.pk_bench_ntt(self._h, log_n, iters, ctypes.byref(ms)))
        return ms.value

    def bench_msm(self, n, iters=3):
        ms = ctypes.c_double()
        self._check(self._lib.pk_bench_msm(self._h, n, iters, ctypes.byref(ms)))
        return ms.value

    def bench_msm_pattern(self, n, pattern, iters=3):
        ms = ctypes.c_double()
        self._check(self._lib.pk_bench_msm_pattern(self._h, n, pattern, iters, ctypes.byref(ms)))
        return ms.value

    def bench_fieldmul(self, which=0):
        g = ctypes.c_double()
        self._check(self._lib.pk_bench_fieldmul(self._h, which, ctypes.byref(g)))
        return g.value


class CommGroup:
    """In-process communicator (pk_comm_group): the ranks of a sharded prover as threads of this process."""

    def __init__(self, world):
        self._lib = load()
        h = ctypes.c_void_p()
        rc = self._lib.pk_comm_group_create(world, ctypes.byref(h))
        if rc != PK_OK:
            raise SynthesisError(rc, "pk_comm_group_create(%d)" % world)
        self._h, self.world = h, world

    def close(self):
        if getattr(self, "_h", None):
            self._lib.pk_comm_group_destroy(self._h)
            self._h = None


def nccl_unique_id() -> bytes:
    buf = ctypes.create_string_buffer(128)
    rc = load().pk_comm_nccl_unique_id(buf)
    if rc != PK_OK:
        raise SynthesisError(rc, "pk_comm_nccl_unique_id (is libnccl.so.2 loadable?)")
    return buf.raw


def g1_sum(points) -> np.ndarray:
    """Sum of affine points on the host (pk_g1_sum): the local fold of the sharded MSM.  Needs no GPU."""
    pts = np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 8)
    out = np.zeros(8, dtype=np.uint64)
    rc = load().pk_g1_sum(_ptr(pts) if pts.shape[0] else None, pts.shape[0], _ptr(out))
    if rc != PK_OK:
        raise SynthesisError(rc, "pk_g1_sum")
    return out


def constants():
    out = np.zeros(24, dtype=np.uint64)
    load().pk_constants(_ptr(out))
    return out
