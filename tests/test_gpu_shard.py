"""GPU parity tests of the sharded single-proof prover (BASELINE.json configs[2]): ONE proof cut across
`world` ranks — commitments by base chunk, the quotient by coset, one all-to-all inside the size-4n inverse NTT — must give
the bytes of the single-GPU prover and of the oracle.  On a one-GPU box the ranks are threads of this process, each with
its own context on device 0 (pk_comm_attach_group); with >= 2 GPUs the same prover also runs one process per GPU over
NCCL (torchrun, tools/shard_check.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, SIMPLE, vk_commitments
from plonkit_b200 import _lib, plonk, synth
from plonkit_b200.bn254 import R_MOD, ints_to_limbs
from plonkit_b200.reader import Crs

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("kind,log_n", [("poseidon", 6), ("poseidon", 10), ("random", 9), ("poseidon", 13)])
def test_sharded_proof_bytes_equal_oracle(orc, world, kind, log_n):
    asm = synth.poseidon_chain_assembly(log_n) if kind == "poseidon" else synth.random_gate_assembly(log_n, seed=log_n, num_inputs=3)
    srs = orc.srs_gen(asm.n, 42, threads=8)
    sp = plonk.ShardedProver(asm, Crs(srs), world)
    try:
        ref = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=8)
        assert sp.prove(asm).to_bytes() == ref
        # resident witness, and the verification key through the sharded commitments
        sp.upload_witness(asm.var_values)
        assert sp.prove(None).to_bytes() == ref
        com = orc.setup_commitments(asm.n, asm.num_inputs, asm.wire_idx, asm.selectors, srs, nvars=asm.nvars, threads=8)
        assert (vk_commitments(sp.make_verification_key()) == com).all()
    finally:
        sp.close()


def test_sharded_prover_reproduces_reference_proof_bin(simple_circuit, simple_key):
    """src/tests.rs:48-73 through 2 ranks: the reference's golden proof.bin / vk.bin (n = 8 is too small for more)."""
    sp = plonk.ShardedProver(simple_circuit, simple_key, 2)
    try:
        assert sp.prove(simple_circuit).to_bytes() == open(os.path.join(SIMPLE, "proof.bin"), "rb").read()
        assert sp.make_verification_key().to_bytes() == open(os.path.join(SIMPLE, "vk.bin"), "rb").read()
    finally:
        sp.close()


def test_sharded_prover_errors_reach_every_rank_and_do_not_stick(orc):
    asm = synth.poseidon_chain_assembly(8)
    srs = orc.srs_gen(asm.n, 42, threads=4)
    sp = plonk.ShardedProver(asm, Crs(srs), 4)
    try:
        bad = asm.var_values.copy()
        bad[7] = ints_to_limbs([(int(bad[7][0]) + 1) % R_MOD])[0]
        with pytest.raises(_lib.SynthesisError) as e:
            sp.prove(bad)
        assert e.value.code == 4                     # Unsatisfiable, as SetupForProver.prove
        with pytest.raises(_lib.SynthesisError) as e:
            sp.prove(asm.var_values[:5])
        assert e.value.code == 1                     # AssignmentMissing
        ref = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=4)
        assert sp.prove(asm).to_bytes() == ref       # the communicator is usable again after a failed call
    finally:
        sp.close()
    with pytest.raises(_lib.SynthesisError):         # 8 ranks need n >= 64
        plonk.ShardedProver(synth.poseidon_chain_assembly(5), Crs(orc.srs_gen(32, 42)), 8)


def test_sharded_proof_2pow18_equals_single_gpu_prover(ctx, orc):
    """A size where every code path is in its large-input regime (multi-pass NTTs, default MSM window plan on 2^15-base
    chunks): 8 ranks == the single-GPU prover == the oracle."""
    log_n = 18
    asm = synth.poseidon_chain_assembly(log_n)
    srs = ctx.srs_gen(asm.n, 42)
    key = Crs(srs)
    single = plonk.SetupForProver.prepare_setup_for_prover(asm, key, None, ctx=ctx)
    want = single.prove(asm).to_bytes()
    single.close()
    assert want == orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=16)
    for world in (2, 8):
        sp = plonk.ShardedProver(asm, key, world)
        try:
            assert sp.prove(asm).to_bytes() == want
        finally:
            sp.close()


def test_sharded_prover_over_nccl_two_processes():
    """One process per GPU over NCCL (the transport bench.py's shard mode uses).  Needs two GPUs; the driver's one-GPU box
    skips it (the in-process transport above covers the prover itself there)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", os.path.join(ROOT, "tools", "shard_check.py"), "--log-n", "16"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sharded == single GPU: True" in r.stdout


def test_sharded_prover_threads_on_distinct_gpus(orc):
    """The in-process transport across REAL devices (peer copies and peer stores over NVLink): one rank per GPU as threads of
    this process.  Needs >= 2 GPUs; on the one-GPU box the same transport runs with every rank on device 0 (tests above)."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs 2 GPUs")
    world = 8 if ngpu >= 8 else (4 if ngpu >= 4 else 2)
    asm = synth.poseidon_chain_assembly(16)
    srs = orc.srs_gen(asm.n, 42, threads=16)
    want = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=16)
    sp = plonk.ShardedProver(asm, Crs(srs), world, devices=list(range(world)))
    try:
        assert sp.prove(asm).to_bytes() == want
        assert sp.prove(asm).to_bytes() == want      # second proof: the peers' receive buffers are reused
    finally:
        sp.close()
