#!/usr/bin/env python3
"""Regenerates tests/golden/ (run in the build container, where /root/reference exists).

1. Copies the reference's OWN golden vectors for the prove path — data files its unit tests assert on
   (src/tests.rs:5-10,30-73): the `simple` circuit (R1CS json, witness), its vk.bin / proof.bin and the 2^10 SRS.
   These are fixtures, not sources.
2. Writes oracle-generated known-answer vectors used by the GPU parity tests at sizes where recomputing them
   on the GPU box would be wasteful (the oracle itself is pinned by (1), see tests/test_oracle.py).
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)


def main():
    dst = os.path.join(HERE, "simple")
    os.makedirs(dst, exist_ok=True)
    for rel in ("test/circuits/simple/circuit.r1cs.json", "test/circuits/simple/witness.json", "test/circuits/simple/vk.bin",
                "test/circuits/simple/proof.bin", "keys/setup/setup_2^10.key"):
        shutil.copy(os.path.join(REF, rel), os.path.join(dst, os.path.basename(rel)))
    import numpy as np
    from oracle import oracle as orc
    from plonkit_b200 import reader, synth
    key = reader.load_key_monomial_form(os.path.join(dst, "setup_2^10.key"))
    # poseidon-shaped circuit at N = 2^9 with the in-tree SRS: proof bytes + vk commitments from the oracle
    asm = synth.poseidon_chain_assembly(9)
    proof = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, key.g1_bases, threads=8)
    com = orc.setup_commitments(asm.n, asm.num_inputs, asm.wire_idx, asm.selectors, key.g1_bases, nvars=asm.nvars, threads=8)
    open(os.path.join(HERE, "poseidon9_proof.bin"), "wb").write(proof)
    np.save(os.path.join(HERE, "poseidon9_vk_commitments.npy"), com)
    # NTT / MSM known answers on seeded inputs
    x = synth.random_field_elements(1 << 10, seed=synth.SEED + 1)
    np.save(os.path.join(HERE, "ntt10_out.npy"), orc.ntt(x, threads=8))
    np.save(os.path.join(HERE, "msm10_out.npy"), orc.msm(x, key.g1_bases, threads=8))


if __name__ == "__main__":
    main()
