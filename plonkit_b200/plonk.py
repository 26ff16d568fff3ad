"""Host-side mirror of the reference's prover façade, src/plonk.rs, over the CUDA library.

Same names, argument meaning and error behaviour as the Rust API (SURVEY.md §8b "Keep"):

    gen_key_monomial_form(power)                                   src/plonk.rs:30-48
    SetupForProver.prepare_setup_for_prover(circuit, key, lagrange) src/plonk.rs:97-119
    SetupForProver.make_verification_key()                          src/plonk.rs:122-124
    SetupForProver.validate_witness(circuit)                        src/plonk.rs:127-129
    SetupForProver.prove(circuit, transcript)                       src/plonk.rs:132-176
    SetupForProver.get_srs_lagrange_form_from_monomial_form()       src/plonk.rs:179-185
    analyse(circuit)                                                src/plonk.rs:72-95

`circuit` is a `CircomCircuit` (R1CS + witness, synthesised on the host exactly like the reference does) or an
already synthesised `Assembly` (gate tables + variable values).  All arithmetic runs in the CUDA library; without it
(or without a GPU) these calls raise — there is no CPU path.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import Context, SynthesisError
from .bn254 import limbs_to_ints
from .circuit import AUX_OFFSET, Assembly, CircomCircuit, analyse, is_satisfied, synthesize  # noqa: F401  (re-exported)
from .reader import CRS_42_G2, Crs, Proof, VerificationKey

SETUP_MIN_POW2 = 10  # src/plonk.rs:26
SETUP_MAX_POW2 = 26  # src/plonk.rs:27

_default_ctx = {}


def default_context(device=0) -> Context:
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]


def gen_key_monomial_form(power: int, ctx: Context = None) -> Crs:
    """Crs::crs_42(1 << power): [42^i] G plus the two G2 elements [1]G2, [42]G2 (src/plonk.rs:30-48)."""
    if not (SETUP_MIN_POW2 <= power <= SETUP_MAX_POW2):
        raise ValueError("setup power of two is not in the correct range")
    ctx = ctx or default_context()
    return Crs(ctx.srs_gen(1 << power, 42), CRS_42_G2, "monomial")


def _as_assembly(circuit) -> Assembly:
    if isinstance(circuit, Assembly):
        return circuit
    if isinstance(circuit, CircomCircuit):
        return synthesize(circuit)
    raise TypeError("circuit must be a CircomCircuit or an Assembly")


class _WitnessSource:
    """What a setup remembers of the circuit it was prepared from, so that `prove(circuit)` with another witness of the SAME
    R1CS assigns the variables through the circuit's WitnessPlan (native host code, milliseconds) instead of transpiling
    the R1CS again the way bellman's `prove` re-synthesises it (src/plonk.rs:132-176) — same values, same proof."""

    def __init__(self, circuit, asm):
        self.plan = getattr(asm, "plan", None) if isinstance(circuit, CircomCircuit) else None
        self.r1cs = circuit.r1cs if isinstance(circuit, CircomCircuit) else None
        self.strict = getattr(circuit, "strict", None)

    def values(self, circuit):
        """var_values ((nvars, 4) uint64) of `circuit`: an array as is, an Assembly's own, a CircomCircuit's via the plan"""
        if isinstance(circuit, np.ndarray):
            return circuit
        if (self.plan is not None and isinstance(circuit, CircomCircuit) and circuit.r1cs is self.r1cs
                and circuit.strict == self.strict):
            if circuit.witness is None:
                return None
            return self.plan.assign(circuit.witness, circuit.wire_mapping)
        return _as_assembly(circuit).var_values


def domain_log2(circuit) -> int:
    """log2 of the evaluation domain of a circuit (setup_polynomials.n.next_power_of_two(), src/plonk.rs:180)."""
    return _as_assembly(circuit).n.bit_length() - 1


class SetupForProver:
    """src/plonk.rs:50-55: setup polynomials + SRS, resident on the device."""

    def __init__(self):
        raise TypeError("use SetupForProver.prepare_setup_for_prover")

    @classmethod
    def prepare_setup_for_prover(cls, circuit, key_monomial_form: Crs, key_lagrange_form: Crs = None, ctx: Context = None):
        self = object.__new__(cls)
        asm = _as_assembly(circuit)
        size = asm.n.bit_length() - 1
        setup_power_of_two = max(size, SETUP_MIN_POW2)
        if not (SETUP_MIN_POW2 <= setup_power_of_two <= SETUP_MAX_POW2):
            raise ValueError("setup power of two is not in the correct range")
        if key_monomial_form.size < asm.n:
            raise SynthesisError(2, "monomial SRS holds %d bases, the circuit needs %d" % (key_monomial_form.size, asm.n))
        if key_lagrange_form is not None and key_lagrange_form.size != asm.n:
            # L_i(tau) depends on the domain: a Lagrange key serves exactly one domain size (dump-lagrange sizes it from the
            # circuit, src/bin/main.rs:360-381)
            raise SynthesisError(2, "Lagrange SRS holds %d bases, the circuit's domain has %d points" % (key_lagrange_form.size, asm.n))
        self.ctx = ctx or default_context()
        self.key_monomial_form = key_monomial_form
        self.key_lagrange_form = key_lagrange_form
        self.n = asm.n
        self.num_inputs = asm.num_inputs
        self.nvars = asm.nvars
        self._source = _WitnessSource(circuit, asm)
        self._loaded_key_id = None
        self._ensure_srs()
        wire_idx = np.ascontiguousarray(asm.wire_idx, dtype=np.uint32)
        selectors = np.ascontiguousarray(asm.selectors, dtype=np.uint64)
        a = _lib.PkAssembly(asm.n, asm.num_inputs, asm.nvars, wire_idx.ctypes.data, selectors.ctypes.data)
        h = ctypes.c_void_p()
        self.ctx._check(self.ctx._lib.pk_setup_create(self.ctx._h, ctypes.byref(a), ctypes.byref(h)))
        self._h = h
        self.ctx._children.add(self)
        return self

    def _ensure_srs(self):
        # the context keeps one SRS resident; reload only if another key was loaded in between.  The tag is a token
        # owned by the Crs object (never reused, unlike id()) and Context.srs_load_g1 clears it on every direct load.
        tag = (self.key_monomial_form.token, self.n)
        if self.ctx.srs_tag != tag:
            self.ctx.srs_load_g1(self.key_monomial_form.g1_bases[:self.n], tag=tag)
        if self.key_lagrange_form is not None:
            ltag = (self.key_lagrange_form.token, self.n)
            if self.ctx.lagrange_tag != ltag:
                self.ctx.srs_load_g1_lagrange(self.key_lagrange_form.g1_bases, tag=ltag)

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.ctx, "_h", None):  # the context owns the device; once it is gone so is this setup
                self.ctx._lib.pk_setup_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def make_verification_key(self) -> VerificationKey:
        self._ensure_srs()
        out = np.zeros((11, 8), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.pk_setup_commitments(self.ctx._h, self._h, out.ctypes.data))
        return VerificationKey(self.n - 1, self.num_inputs, out[:6].copy(), out[6:7].copy(), out[7:11].copy(), [5, 7, 10],
                               self.key_monomial_form.g2_raw)

    def validate_witness(self, circuit) -> None:
        """is_satisfied_using_one_shot_check on the host tables (src/plonk.rs:127-129); raises on failure."""
        if not is_satisfied(_as_assembly(circuit)):
            raise SynthesisError(4, "witness does not satisfy the circuit")

    def upload_witness(self, circuit_or_values):
        vals = self._source.values(circuit_or_values)
        if vals is None:
            raise SynthesisError(1, "circuit has no witness")
        vals = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
        self.ctx._check(self.ctx._lib.pk_witness_upload(self.ctx._h, self._h, vals.ctypes.data, vals.shape[0]))

    def prove(self, circuit=None, transcript: str = "keccak") -> Proof:
        """SetupForProver::prove.  circuit=None proves the witness uploaded with upload_witness (device-resident input).
        With a `key_lagrange_form` the wire commitments are taken from the witness values (bellman `prove`, src/plonk.rs:138-146;
        only 'keccak' exists on that branch, as in the reference) — the proof bytes are the same."""
        if self.key_lagrange_form is not None and transcript != "keccak":
            raise NotImplementedError("invalid transcript. use 'keccak'")  # src/plonk.rs:147-149: unimplemented!()
        if transcript == "rescue":
            raise NotImplementedError("the rescue transcript has no in-tree fixture (SURVEY.md §0 item 5); use 'keccak'")
        if transcript != "keccak":
            raise NotImplementedError("invalid transcript. use 'keccak' or 'rescue'")
        self._ensure_srs()
        vals, nvars = None, 0
        if circuit is not None:
            arr = self._source.values(circuit)
            if arr is None:
                raise SynthesisError(1, "circuit has no witness")
            vals = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
            nvars = vals.shape[0]
        pr = _lib.PkProof()
        inputs = np.zeros((max(self.num_inputs, 1), 4), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.pk_setup_use_lagrange(self.ctx._h, self._h, int(self.key_lagrange_form is not None)))
        self.ctx._check(self.ctx._lib.pk_prove(self.ctx._h, self._h, vals.ctypes.data if vals is not None else None, nvars,
                                               ctypes.byref(pr), inputs.ctypes.data))
        return _proof_from_struct(pr, inputs[:self.num_inputs])

    def get_srs_lagrange_form_from_monomial_form(self) -> Crs:
        self._ensure_srs()
        size = self.n  # setup_polynomials.n.next_power_of_two()
        pts = self.ctx.ec_intt_g1(size.bit_length() - 1)
        return Crs(pts, self.key_monomial_form.g2_raw, "lagrange")


class ShardedSetupForProver:
    """ONE rank of a `SetupForProver` whose proof is computed by `world` GPUs together (BASELINE.json
    configs[2]): same methods, same proof bytes; every call is collective (all ranks call it with the same arguments,
    all ranks get the same result).  `ctx` must already be attached to a communicator (Context.attach_nccl under
    torchrun, Context.attach_group for the ranks-as-threads form `ShardedProver` drives)."""

    def __init__(self):
        raise TypeError("use ShardedSetupForProver.prepare_setup_for_prover")

    @classmethod
    def prepare_setup_for_prover(cls, circuit, key_monomial_form: Crs, ctx: Context):
        self = object.__new__(cls)
        asm = _as_assembly(circuit)
        if key_monomial_form.size < asm.n:
            raise SynthesisError(2, "monomial SRS holds %d bases, the circuit needs %d" % (key_monomial_form.size, asm.n))
        self.ctx, self.key_monomial_form = ctx, key_monomial_form
        self.n, self.num_inputs, self.nvars = asm.n, asm.num_inputs, asm.nvars
        self._source = _WitnessSource(circuit, asm)
        world, rank = ctx.world, ctx.rank
        if asm.n % world:
            raise SynthesisError(6, "domain size is not divisible by the number of ranks")
        cn = asm.n // world
        ctx.srs_load_g1(key_monomial_form.g1_bases[rank * cn:(rank + 1) * cn])  # this rank's chunk of the key only
        wire_idx = np.ascontiguousarray(asm.wire_idx, dtype=np.uint32)
        selectors = np.ascontiguousarray(asm.selectors, dtype=np.uint64)
        a = _lib.PkAssembly(asm.n, asm.num_inputs, asm.nvars, wire_idx.ctypes.data, selectors.ctypes.data)
        h = ctypes.c_void_p()
        ctx._check(ctx._lib.pk_dist_setup_create(ctx._h, ctypes.byref(a), ctypes.byref(h)))
        self._h = h
        ctx._children.add(self)
        return self

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.ctx, "_h", None):
                self.ctx._lib.pk_dist_setup_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def make_verification_key(self) -> VerificationKey:
        out = np.zeros((11, 8), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.pk_dist_setup_commitments(self.ctx._h, self._h, out.ctypes.data))
        return VerificationKey(self.n - 1, self.num_inputs, out[:6].copy(), out[6:7].copy(), out[7:11].copy(), [5, 7, 10],
                               self.key_monomial_form.g2_raw)

    def upload_witness(self, values):
        vals = self._source.values(values)
        if vals is None:
            raise SynthesisError(1, "circuit has no witness")
        vals = np.ascontiguousarray(vals, dtype=np.uint64).reshape(-1, 4)
        self.ctx._check(self.ctx._lib.pk_dist_witness_upload(self.ctx._h, self._h, vals.ctypes.data, vals.shape[0]))

    def prove(self, circuit=None, transcript: str = "keccak") -> Proof:
        if transcript != "keccak":
            raise NotImplementedError("the sharded prover implements the 'keccak' transcript")
        vals, nvars = None, 0
        if circuit is not None:
            arr = self._source.values(circuit)
            if arr is None:
                raise SynthesisError(1, "circuit has no witness")
            vals = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1, 4)
            nvars = vals.shape[0]
        pr = _lib.PkProof()
        inputs = np.zeros((max(self.num_inputs, 1), 4), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.pk_dist_prove(self.ctx._h, self._h, vals.ctypes.data if vals is not None else None, nvars,
                                                    ctypes.byref(pr), inputs.ctypes.data))
        return _proof_from_struct(pr, inputs[:self.num_inputs])


class ShardedProver:
    """`world` ranks of a sharded prover as threads of THIS process, one context each, on `devices` (default: every rank
    on device 0 — how the single-GPU parity tests exercise world sizes 2, 4 and 8; pass distinct devices to spread one
    proof over the GPUs of a node from a single process, peer copies over NVLink).  `prove` returns rank 0's proof after
    checking that every rank produced the same bytes."""

    def __init__(self, circuit, key_monomial_form: Crs, world: int, devices=None):
        self.world = world
        devices = list(devices) if devices is not None else [0] * world
        self.group = _lib.CommGroup(world)
        self.ctxs = [Context(devices[r]) for r in range(world)]
        for r, c in enumerate(self.ctxs):
            c.attach_group(self.group, r)
        self.setups = self._collective(lambda r: ShardedSetupForProver.prepare_setup_for_prover(circuit, key_monomial_form, self.ctxs[r]))

    def _collective(self, fn):
        import threading
        out, errs = [None] * self.world, [None] * self.world

        def run(r):
            try:
                out[r] = fn(r)
            except Exception as ex:
                errs[r] = ex
        th = [threading.Thread(target=run, args=(r,)) for r in range(self.world)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        for e in errs:  # the first genuine failure, not the "a peer rank failed" echoes
            if e is not None and "peer rank" not in str(e):
                raise e
        for e in errs:
            if e is not None:
                raise e
        return out

    def make_verification_key(self) -> VerificationKey:
        return self._collective(lambda r: self.setups[r].make_verification_key())[0]

    def upload_witness(self, values):
        self._collective(lambda r: self.setups[r].upload_witness(values))

    def prove(self, circuit=None, transcript: str = "keccak") -> Proof:
        proofs = self._collective(lambda r: self.setups[r].prove(circuit, transcript))
        ref = proofs[0].to_bytes()
        if any(p.to_bytes() != ref for p in proofs[1:]):
            raise SynthesisError(6, "ranks of the sharded prover disagree on the proof")
        return proofs[0]

    def close(self):
        for s_ in getattr(self, "setups", []) or []:
            s_.close()
        for c in self.ctxs:
            c.close()
        self.group.close()
        self.setups, self.ctxs = [], []


class ProverPool:
    """Several `SetupForProver` instances of the SAME circuit on one GPU, each with its own library context (stream, SRS
    window tables, scratch) and its own host thread: proofs are independent objects, and while one sits in its
    latency-bound kernels (bucket sort, scans, transcript round trips) the integer-bound kernels of the others fill the
    SMs.  Measured on B200 at 2^20 gates: 21 proofs/s one at a time, 24 with three in flight (bench.py).  The reference
    has no counterpart (one `Worker` per call, src/plonk.rs:41,47,183); a proving service would run exactly this."""

    def __init__(self, circuit, key_monomial_form: Crs, inflight: int = 3, device: int = 0):
        if inflight < 1:
            raise ValueError("inflight must be >= 1")
        asm = _as_assembly(circuit)                      # one transpilation serves every prover of the pool
        source = _WitnessSource(circuit, asm)
        self.setups = []
        for _ in range(inflight):
            s_ = SetupForProver.prepare_setup_for_prover(asm, key_monomial_form, None, ctx=Context(device))
            s_._source = source
            self.setups.append(s_)

    def prove_all(self, witnesses, transcript: str = "keccak"):
        """Proves every witness (var_values arrays, Assemblies, or CircomCircuits holding other witnesses of the pool's
        R1CS — those are assigned by the worker threads through the circuit's WitnessPlan, native host code that runs
        beside the other provers' device work); returns the proofs in input order."""
        import queue
        import threading
        items = list(enumerate(witnesses))
        out, errs = [None] * len(items), []
        q = queue.Queue()
        for it in items:
            q.put(it)

        def worker(setup):
            while True:
                try:
                    i, w = q.get_nowait()
                except queue.Empty:
                    return
                try:
                    out[i] = setup.prove(w, transcript)
                except Exception as ex:  # re-raised on the caller's thread
                    errs.append(ex)
                    return
        threads = [threading.Thread(target=worker, args=(s_,)) for s_ in self.setups]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        return out

    def close(self):
        for s_ in self.setups:
            s_.close()
            s_.ctx.close()
        self.setups = []


def _arr(field, shape):
    return np.array(list(field), dtype=np.uint64).reshape(shape)


def _proof_from_struct(pr, inputs) -> Proof:
    p = Proof(
        n=pr.n, num_inputs=pr.num_inputs, input_values=limbs_to_ints(inputs) if len(inputs) else [],
        wire_commitments=_arr(pr.wire_commitments, (4, 8)),
        grand_product_commitment=_arr(pr.grand_product_commitment, (8,)),
        quotient_poly_commitments=_arr(pr.quotient_poly_commitments, (4, 8)),
        wire_values_at_z=limbs_to_ints(_arr(pr.wire_values_at_z, (4, 4))),
        wire_values_at_z_omega=limbs_to_ints(_arr(pr.wire_values_at_z_omega, (1, 4))),
        grand_product_at_z_omega=limbs_to_ints(_arr(pr.grand_product_at_z_omega, (1, 4)))[0],
        quotient_polynomial_at_z=limbs_to_ints(_arr(pr.quotient_polynomial_at_z, (1, 4)))[0],
        linearization_polynomial_at_z=limbs_to_ints(_arr(pr.linearization_polynomial_at_z, (1, 4)))[0],
        permutation_polynomials_at_z=limbs_to_ints(_arr(pr.permutation_polynomials_at_z, (3, 4))),
        opening_at_z_proof=_arr(pr.opening_at_z_proof, (8,)),
        opening_at_z_omega_proof=_arr(pr.opening_at_z_omega_proof, (8,)),
    )
    p.challenges = limbs_to_ints(_arr(pr.challenges, (5, 4)))
    return p


def verify(vk: VerificationKey, proof: Proof, transcript: str = "keccak") -> bool:
    """src/plonk.rs:189-210: bellman's better_cs verifier with the Keccak transcript — host arithmetic with the BN254 pairing
    (plonkit_b200/verifier.py), as in the reference; the `rescue` transcript's parameters are not in the reference tree."""
    if transcript == "rescue":
        raise NotImplementedError("the rescue transcript has no in-tree fixture (SURVEY.md section 0 item 5); use 'keccak'")
    if transcript != "keccak":
        raise NotImplementedError("invalid transcript. use 'keccak' or 'rescue'")
    from . import verifier
    return verifier.verify(vk, proof)
