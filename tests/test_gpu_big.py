"""Opt-in full-size parity runs (PK_BIG=1 python -m pytest tests/test_gpu_big.py -m gpu -s): sizes above what the
driver-run suite can afford, recorded once per round under profiles/ (SURVEY 8d cfg 3: "do once").

    2^22  random-gate circuit (BASELINE configs[2] shape): single-GPU proof bytes == oracle bytes; sharded (8 in-process ranks) == same
    2^24  (PK_BIG=24) single-GPU proof bytes == oracle bytes (BASELINE configs[2] size; the oracle needs ~5 min of 16 host threads
          and ~70 GB of host memory), == sharded proof (2 in-process ranks), trapdoor-verified
"""
import os
import time

import pytest

from conftest import vk_commitments
from plonkit_b200 import plonk, synth
from plonkit_b200.reader import Crs

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not os.environ.get("PK_BIG"), reason="opt-in: PK_BIG=1 (2^22) or PK_BIG=24")]


def _run(ctx, orc, log_n, with_oracle, worlds):
    t0 = time.time()
    asm = synth.random_gate_assembly_layered(log_n, seed=2024 + log_n)
    srs = ctx.srs_gen(asm.n, 42)
    key = Crs(srs)
    print("\n2^%d random-gate circuit + SRS: %.1f s" % (log_n, time.time() - t0), flush=True)
    setup = plonk.SetupForProver.prepare_setup_for_prover(asm, key, None, ctx=ctx)
    setup.upload_witness(asm.var_values)
    got = setup.prove(None).to_bytes()
    ctx.timer_begin()
    setup.prove(None)
    print("single-GPU prove: %.1f ms" % ctx.timer_end(), flush=True)
    com = vk_commitments(setup.make_verification_key())
    setup.close()
    assert orc.verify_trapdoor(got, com, 42)
    if with_oracle:
        t0 = time.time()
        ref = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=os.cpu_count() or 16)
        print("oracle prove: %.1f s; bytes equal: %s" % (time.time() - t0, ref == got), flush=True)
        assert ref == got
    for world in worlds:
        sp = plonk.ShardedProver(asm, key, world)
        try:
            assert sp.prove(asm).to_bytes() == got
            print("sharded over %d in-process ranks: bytes equal" % world, flush=True)
        finally:
            sp.close()


def test_2pow22_random_gate_circuit_bytes_equal_oracle(ctx, orc):
    _run(ctx, orc, 22, True, (8,))


@pytest.mark.skipif(os.environ.get("PK_BIG") != "24", reason="PK_BIG=24")
def test_2pow24_random_gate_circuit_single_equals_sharded_and_verifies(ctx, orc):
    _run(ctx, orc, 24, True, (2,))
