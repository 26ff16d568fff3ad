"""CPU rehearsal of the device tests' HOST flows.

The reference-facing tests of tests/test_gpu_prove.py (golden proof.bin / vk.bin, Lagrange key, error behaviour, the CLI) are
run here against a stand-in for the CUDA library's context whose pk_* entry points are answered by the oracle.  Nothing is
measured or shipped this way — the product fails loudly without the CUDA library (tests/test_host.py) — the point is that
everything on the host side of the C ABI (file readers, compiled and Python transpilers, witness assignment, argument
marshalling, error mapping, file writers) is exercised by the very assertions the B200 run makes, on every CPU run.
"""
import ctypes
import io

import numpy as np
import pytest

import test_gpu_prove as device_tests
import test_gpu_shard as shard_tests
from plonkit_b200 import _lib, circuit, plonk, reader
from plonkit_b200.bn254 import ints_to_limbs


def _view(ptr, dtype, count):
    ct = {np.uint32: ctypes.c_uint32, np.uint64: ctypes.c_uint64}[dtype]
    return np.ctypeslib.as_array((ct * count).from_address(ptr))


class _Setup:
    pass


class OracleBackedContext:
    """Has the attributes and methods of _lib.Context that plonkit_b200.plonk and the CLI use; the oracle computes."""

    class _Lib:
        def __init__(self, ctx, orc):
            self.ctx, self.orc, self.setups, self.err = ctx, orc, {}, b""

        def pk_last_error(self, h):
            return self.err

        def pk_setup_create(self, h, a_ref, out_ref):
            a = a_ref._obj
            s = _Setup()
            s.n, s.num_inputs, s.nvars = int(a.n), int(a.num_inputs), int(a.nvars)
            s.wire_idx = _view(a.wire_idx, np.uint32, 4 * s.n).reshape(4, s.n).copy()
            s.selectors = _view(a.selectors, np.uint64, 28 * s.n).reshape(7, s.n, 4).copy()
            s.values, s.lagrange = None, 0
            if self.ctx.bases is None or self.ctx.bases.shape[0] < s.n:
                self.err = b"SRS smaller than the circuit"
                return 2
            handle = len(self.setups) + 1
            self.setups[handle] = s
            out_ref._obj.value = handle
            return 0

        def pk_setup_destroy(self, sh):
            self.setups.pop(sh.value if hasattr(sh, "value") else sh, None)
            return 0

        def _setup(self, sh):
            return self.setups[sh.value if hasattr(sh, "value") else sh]

        def pk_setup_use_lagrange(self, h, sh, flag):
            s = self._setup(sh)
            if flag and self.ctx.lagrange is None:
                self.err = b"no Lagrange key loaded"
                return 6
            s.lagrange = flag
            return 0

        def pk_setup_commitments(self, h, sh, out_ptr):
            s = self._setup(sh)
            com = self.orc.setup_commitments(s.n, s.num_inputs, s.wire_idx, s.selectors, self.ctx.bases[:s.n], nvars=s.nvars, threads=4)
            _view(out_ptr, np.uint64, 88)[:] = com.reshape(-1)
            return 0

        def pk_witness_upload(self, h, sh, ptr, nvars):
            s = self._setup(sh)
            if nvars != s.nvars:
                self.err = b"witness length"
                return 1
            s.values = _view(ptr, np.uint64, 4 * nvars).reshape(nvars, 4).copy()
            return 0

        def pk_prove(self, h, sh, ptr, nvars, pr_ref, inputs_ptr):
            s = self._setup(sh)
            if ptr:
                rc = self.pk_witness_upload(h, sh, ptr, nvars)
                if rc:
                    return rc
            if s.values is None:
                self.err = b"no witness"
                return 1
            try:
                raw, ch = self.orc.prove(s.n, s.num_inputs, s.wire_idx, s.values, s.selectors, self.ctx.bases[:s.n], threads=4,
                                         want_challenges=True)
            except RuntimeError as e:
                self.err = str(e).encode()
                return 4
            p = reader.Proof.read(io.BytesIO(raw))
            pr = pr_ref._obj
            pr.n, pr.num_inputs = p.n, p.num_inputs

            def put(field, arr):
                flat = np.ascontiguousarray(arr, dtype=np.uint64).reshape(-1)
                for i, x in enumerate(flat.tolist()):
                    field[i] = x
            put(pr.wire_commitments, p.wire_commitments)
            put(pr.grand_product_commitment, p.grand_product_commitment)
            put(pr.quotient_poly_commitments, p.quotient_poly_commitments)
            put(pr.wire_values_at_z, ints_to_limbs(p.wire_values_at_z))
            put(pr.wire_values_at_z_omega, ints_to_limbs(p.wire_values_at_z_omega))
            put(pr.grand_product_at_z_omega, ints_to_limbs([p.grand_product_at_z_omega]))
            put(pr.quotient_polynomial_at_z, ints_to_limbs([p.quotient_polynomial_at_z]))
            put(pr.linearization_polynomial_at_z, ints_to_limbs([p.linearization_polynomial_at_z]))
            put(pr.permutation_polynomials_at_z, ints_to_limbs(p.permutation_polynomials_at_z))
            put(pr.opening_at_z_proof, p.opening_at_z_proof)
            put(pr.opening_at_z_omega_proof, p.opening_at_z_omega_proof)
            put(pr.challenges, ints_to_limbs(ch))
            if p.num_inputs:
                _view(inputs_ptr, np.uint64, 4 * p.num_inputs)[:] = ints_to_limbs(p.input_values).reshape(-1)
            return 0

        # ---- the sharded prover's entry points: every rank holds its chunk of the key; the stand-in puts them together
        def pk_dist_setup_create(self, h, a_ref, out_ref):
            n, world = int(a_ref._obj.n), self.ctx.world
            if n % world or n < world * world:
                self.err = b"domain too small for this many ranks"
                return 6
            full, self.ctx.bases = self.ctx.bases, np.zeros((n, 8), dtype=np.uint64)     # size check of the 1-GPU path does not apply
            try:
                return self.pk_setup_create(h, a_ref, out_ref)
            finally:
                self.ctx.bases = full

        def _with_full_key(self, fn, *args):
            mine = self.ctx.bases
            self.ctx.bases = np.concatenate([m.bases for m in self.ctx.group.members])
            try:
                return fn(*args)
            finally:
                self.ctx.bases = mine

        def pk_dist_setup_commitments(self, h, sh, out_ptr):
            return self._with_full_key(self.pk_setup_commitments, h, sh, out_ptr)

        def pk_dist_prove(self, h, sh, ptr, nvars, pr_ref, inputs_ptr):
            return self._with_full_key(self.pk_prove, h, sh, ptr, nvars, pr_ref, inputs_ptr)

        def pk_dist_witness_upload(self, h, sh, ptr, nvars):
            return self.pk_witness_upload(h, sh, ptr, nvars)

        def pk_dist_setup_destroy(self, sh):
            return self.pk_setup_destroy(sh)

    def __init__(self, orc):
        self._lib, self._h, self._children = self._Lib(self, orc), 1, set()
        self.orc = orc
        self.srs_tag = self.lagrange_tag = None
        self.bases = self.lagrange = None
        self.rank, self.world, self.group = 0, 1, None

    def attach_group(self, group, rank):
        self.rank, self.world, self.group = rank, group.world, group
        group.members[rank] = self

    def _check(self, rc):
        if rc:
            raise _lib.SynthesisError(rc, self._lib.err.decode())

    def srs_load_g1(self, bases, window_bits=0, tag=None):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8).copy()
        self.srs_tag = tag

    def srs_load_g1_lagrange(self, bases, tag=None):
        self.lagrange = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8).copy()
        self.lagrange_tag = tag

    def srs_gen(self, n, tau=42):
        return self.orc.srs_gen(n, tau, threads=8)

    def ec_intt_g1(self, log_n):
        return self.orc.ec_intt(self.bases[:1 << log_n])

    def close(self):
        pass

    def profile(self):
        return {"kernel_launches": 1}


class _Group:
    def __init__(self, world):
        self.world, self.members = world, [None] * world

    def close(self):
        pass


@pytest.fixture()
def fake_ranks(orc, monkeypatch):
    """ShardedProver builds its own contexts and communicator: hand it stand-ins"""
    monkeypatch.setattr(plonk, "Context", lambda device=0: OracleBackedContext(orc))
    monkeypatch.setattr(_lib, "CommGroup", _Group)
    return orc


@pytest.fixture()
def fake(orc, monkeypatch):
    ctx = OracleBackedContext(orc)
    monkeypatch.setattr(plonk, "default_context", lambda device=0: ctx)
    return ctx


def test_stand_in_refuses_what_the_device_refuses(fake, simple_circuit, simple_key):
    setup = plonk.SetupForProver.prepare_setup_for_prover(simple_circuit, simple_key, None, ctx=fake)
    with pytest.raises(_lib.SynthesisError) as e:
        setup.prove(circuit.CircomCircuit(simple_circuit.r1cs, [1, 36, 3, 9]))
    assert e.value.code == 4


def test_rehearse_golden_proof_and_vk(fake, simple_circuit, simple_key):
    device_tests.test_prove_reproduces_reference_proof_bin(fake, simple_circuit, simple_key)
    device_tests.test_verification_key_reproduces_reference_vk_bin(fake, simple_circuit, simple_key)
    device_tests.test_lagrange_srs_matches_oracle(fake, fake.orc, simple_circuit, simple_key)


def test_rehearse_lagrange_key_fixture_and_synthetic_circuits(fake, simple_circuit, simple_key):
    device_tests.test_prove_with_lagrange_key_gives_the_same_bytes(fake, fake.orc, simple_circuit, simple_key)
    device_tests.test_poseidon_shaped_golden_fixture(fake, simple_key)
    device_tests.test_proofs_equal_oracle_on_synthetic_circuits(fake, fake.orc, "random", 6)     # incl. plonk.verify on the result
    device_tests.test_public_input_polynomial_paths(fake, fake.orc, 9)


def test_rehearse_error_behaviour(fake, simple_circuit, simple_key):
    device_tests.test_error_behaviour_mirrors_reference(fake, fake.orc, simple_circuit, simple_key)


def test_rehearse_cli_golden_files(fake, tmp_path):
    device_tests.test_cli_prove_and_export_vk_write_the_golden_files(tmp_path)
    (tmp_path / "k").mkdir()
    device_tests.test_cli_setup_writes_the_reference_key_file(tmp_path / "k")


def test_rehearse_cli_lagrange_key(fake, tmp_path):
    device_tests.test_cli_prove_with_lagrange_key(tmp_path)


def test_rehearse_cli_poseidon_shaped_r1cs(fake, tmp_path, simple_key):
    device_tests.test_cli_proves_a_poseidon_shaped_r1cs_with_wide_linear_combinations(tmp_path, fake.orc, simple_key)


def test_rehearse_both_transpilers(fake, tmp_path, simple_circuit, simple_key):
    """the same golden flows with the Python statements of the parser / transpiler / witness assignment (what runs when
    libplonkit_host.so is not built)"""
    circuit.NATIVE[0] = False
    try:
        device_tests.test_prove_reproduces_reference_proof_bin(fake, simple_circuit, simple_key)
        device_tests.test_cli_prove_and_export_vk_write_the_golden_files(tmp_path)
    finally:
        circuit.NATIVE[0] = True


def test_rehearse_sharded_prover_threads(fake_ranks, simple_circuit, simple_key):
    """the ranks-as-threads flows of tests/test_gpu_shard.py: every rank synthesises the SAME circuit object concurrently"""
    for _ in range(3):
        shard_tests.test_sharded_prover_reproduces_reference_proof_bin(simple_circuit, simple_key)
    shard_tests.test_sharded_proof_bytes_equal_oracle(fake_ranks, 4, "poseidon", 6)
    shard_tests.test_sharded_prover_errors_reach_every_rank_and_do_not_stick(fake_ranks)


def test_rehearse_smoke(fake_ranks, monkeypatch, capsys):
    """__graft_entry__.smoke()'s host flow (what the driver runs on the B200 before the bench)"""
    import __graft_entry__ as entry
    monkeypatch.setattr(_lib, "Context", lambda device=0: OracleBackedContext(fake_ranks))
    entry.smoke()
    assert "smoke ok" in capsys.readouterr().out
