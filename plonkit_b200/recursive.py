"""Host-side mirror of the kernel path behind the reference's proof aggregation, src/recursive/mod.rs.

`recursive::prove` (src/recursive/mod.rs:38-136) does two things: (1) it SYNTHESISES the aggregation circuit
(`RecursiveAggregationCircuitBn256`, :90-108, from crates outside the reference tree — host Rust, out of this repository's
scope), and (2) it proves it: `create_recursive_circuit_setup` (:120-121) and bellman's
better_better_cs `create_proof::<_, RollingKeccakTranscript>` (:127) over a `ProvingAssembly` with two gate types — the
width-4 main gate with d_next and the Rescue x^5 custom gate.  Part (2) is where the NTTs and MSMs are, and it is what
this module runs on the CUDA library, taking the synthesised assembly (gate tables + a gate type per row) as input.

BYTE PARITY UNPINNED: neither that prover nor its proof layout is in the reference tree and no fixture exists.  The
protocol implemented follows the round structure of SURVEY.md App. D;
proofs are checked by this repository's own restated verifier under the known trapdoor, not against reference bytes.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import Context, SynthesisError
from .bn254 import limbs_to_ints
from .circuit import Assembly
from .plonk import _proof_from_struct, default_context
from .reader import Crs, Proof

VK_TREE_DEPTH = 7  # src/recursive/mod.rs:35


class RecursiveVerificationKey:
    """The 13 setup commitments of the two-gate-type prover: 7 main-gate setup polynomials, the gate selectors s_main and
    s_resc, 4 permutation polynomials (export_vk, src/recursive/mod.rs:196-204, exports bellman's VerificationKey of the
    recursive circuit; its serialisation is unpinned)."""

    def __init__(self, n, num_inputs, commitments, g2_raw=b""):
        self.n, self.num_inputs, self.commitments, self.g2_raw = n, num_inputs, commitments, g2_raw


class RecursiveSetupForProver:
    """create_recursive_circuit_setup (src/recursive/mod.rs:120-121) + the proving call (:127) on a synthesised assembly."""

    def __init__(self, assembly: Assembly, gate_type, big_crs: Crs, ctx: Context = None):
        if big_crs.size < assembly.n:
            raise SynthesisError(2, "SRS holds %d bases, the circuit needs %d" % (big_crs.size, assembly.n))
        self.ctx = ctx or default_context()
        self.n, self.num_inputs, self.nvars = assembly.n, assembly.num_inputs, assembly.nvars
        self.ctx.srs_load_g1(big_crs.g1_bases[:assembly.n], tag=(big_crs.token, assembly.n))
        wire_idx = np.ascontiguousarray(assembly.wire_idx, dtype=np.uint32)
        selectors = np.ascontiguousarray(assembly.selectors, dtype=np.uint64)
        gt = np.ascontiguousarray(gate_type, dtype=np.uint8).reshape(assembly.n)
        a = _lib.PkAssemblyGated(_lib.PkAssembly(assembly.n, assembly.num_inputs, assembly.nvars, wire_idx.ctypes.data,
                                                 selectors.ctypes.data), gt.ctypes.data)
        h = ctypes.c_void_p()
        self.ctx._check(self.ctx._lib.pk_setup_create_gated(self.ctx._h, ctypes.byref(a), ctypes.byref(h)))
        self._h = h
        self._tag = (big_crs.token, assembly.n)
        self._bases = big_crs.g1_bases[:assembly.n]
        self._g2_raw = big_crs.g2_raw
        self.ctx._children.add(self)

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self.ctx, "_h", None):
                self.ctx._lib.pk_setup_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ensure_srs(self):
        if self.ctx.srs_tag != self._tag:
            self.ctx.srs_load_g1(self._bases, tag=self._tag)

    def export_vk(self) -> RecursiveVerificationKey:
        self._ensure_srs()
        out = np.zeros((13, 8), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.pk_setup_commitments_gated(self.ctx._h, self._h, out.ctypes.data))
        return RecursiveVerificationKey(self.n - 1, self.num_inputs, out, self._g2_raw)

    def create_proof(self, var_values) -> Proof:
        """better_better_cs create_proof::<_, RollingKeccakTranscript> (src/recursive/mod.rs:127) on the witness."""
        self._ensure_srs()
        vals = np.ascontiguousarray(var_values, dtype=np.uint64).reshape(-1, 4)
        pr = _lib.PkProof()
        inputs = np.zeros((max(self.num_inputs, 1), 4), dtype=np.uint64)
        self.ctx._check(self.ctx._lib.pk_setup_use_lagrange(self.ctx._h, self._h, 0))
        self.ctx._check(self.ctx._lib.pk_prove(self.ctx._h, self._h, vals.ctypes.data, vals.shape[0], ctypes.byref(pr), inputs.ctypes.data))
        proof = _proof_from_struct(pr, inputs[:self.num_inputs])
        proof.gate_selectors_at_z = limbs_to_ints(np.array(list(pr.gate_selectors_at_z), dtype=np.uint64).reshape(2, 4))
        return proof


def prove(big_crs, old_proofs, old_vk):
    """src/recursive/mod.rs:38-42.  The aggregation circuit (`RecursiveAggregationCircuitBn256`, rns / rescue parameters, the
    vk Merkle tree) comes from crates that are not in the reference tree and is synthesised by host Rust; hand the synthesised
    assembly to `RecursiveSetupForProver(...).create_proof(witness)` instead."""
    raise NotImplementedError("aggregation-circuit synthesis is host Rust outside this repository's scope; the proving call "
                              "behind it is recursive.RecursiveSetupForProver.create_proof")


def verify(vk: RecursiveVerificationKey, proof: Proof) -> bool:
    """The proof check of src/recursive/mod.rs:139-166 (better_better_cs verifier, RollingKeccakTranscript) for proofs of
    `RecursiveSetupForProver.create_proof`: host arithmetic with the BN254 pairing.  (The aggregation bookkeeping of that
    function - vks_tree root, the 4 aggregated limbs - belongs to the circuit synthesis and is out of scope.)"""
    from . import verifier
    if proof.n != vk.n or proof.num_inputs != vk.num_inputs:
        return False
    return verifier.verify_gated(vk.commitments, vk.g2_raw, proof)
