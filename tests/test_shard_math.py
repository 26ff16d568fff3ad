"""CPU tests of the sharded prover's algebra (tests/shard_model.py) against the oracle's plain transforms: ownership of the
coset domain, evaluation on a rank's range by fold + smaller NTT, and the split inverse transform — for 1, 2, 4 and 8
simulated ranks in one process, and over gloo with 2 processes (the exchange pattern of the real collectives)."""
import os

import numpy as np
import pytest

from conftest import ROOT
from plonkit_b200 import synth
from plonkit_b200.bn254 import R_MOD, ints_to_limbs, limbs_to_ints
from shard_model import Plan, inverse_cross_stages, inverse_local_stages, slot_layout_of


def _setup(orc, log_n, seed):
    n = 1 << log_n
    coef = limbs_to_ints(synth.random_field_elements(n, seed=seed))
    nat = limbs_to_ints(orc.lde4(ints_to_limbs(coef), threads=2))      # values on 7 H_4n, natural order
    return coef, slot_layout_of(nat, log_n), orc.omega(log_n), orc.omega(log_n + 2)


@pytest.mark.parametrize("G", [1, 2, 4, 8])
@pytest.mark.parametrize("log_n", [6, 8])
def test_rank_ranges_tile_the_coset_domain_and_match_the_plain_lde(orc, G, log_n):
    coef, slots, wn, w4 = _setup(orc, log_n, 40 + log_n)
    got = []
    for r in range(G):
        plan = Plan(log_n, G, r, wn, w4)
        part = plan.range_lde(coef)
        assert len(part) == plan.m
        got += part
    assert got == slots                                                # concatenated ranges == the whole slot layout


@pytest.mark.parametrize("G", [1, 2, 4, 8])
def test_split_inverse_transform_matches_the_plain_inverse_coset_ntt(orc, G):
    log_n = 6
    n, M = 1 << log_n, 4 << log_n
    vals_nat = limbs_to_ints(synth.random_field_elements(M, seed=9))   # values on the natural coset 7 H_4n
    want = limbs_to_ints(orc.ntt(ints_to_limbs(vals_nat), inverse=True, coset=True, threads=2))
    slots = slot_layout_of(vals_nat, log_n)
    w_inv = pow(orc.omega(log_n + 2), R_MOD - 2, R_MOD)
    m, per = M // G, M // G // G
    blocks = [inverse_local_stages(slots[r * m:(r + 1) * m], log_n + 2, w_inv) for r in range(G)]
    coeffs = [0] * M
    for r in range(G):
        recv = [blocks[c][r * per:(r + 1) * per] for c in range(G)]   # the all-to-all: block r of every rank's range
        out = inverse_cross_stages(recv, G, r, log_n + 2, w_inv)
        for c in range(G):                                             # the all-gather into natural order
            coeffs[c * m + r * per:c * m + (r + 1) * per] = out[c]
    assert coeffs == want


def _gloo_worker(rank, world, port, ret):
    import sys
    import torch
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, ROOT)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle import oracle as orc
        log_n = 6
        M = 4 << log_n
        m, per = M // world, M // world // world
        coef = limbs_to_ints(synth.random_field_elements(1 << log_n, seed=21))
        plan = Plan(log_n, world, rank, orc.omega(log_n), orc.omega(log_n + 2))
        mine = plan.range_lde(coef)                                    # this rank's range of the LDE, nobody else's data
        w_inv = pow(orc.omega(log_n + 2), R_MOD - 2, R_MOD)
        blk = inverse_local_stages(mine, log_n + 2, w_inv)
        send = torch.from_numpy(ints_to_limbs(blk).view(np.int64).reshape(world, per, 4).copy())
        recv = torch.empty_like(send)
        td.all_to_all_single(recv, send)                               # ONE all-to-all
        rc = [limbs_to_ints(recv[c].numpy().view(np.uint64)) for c in range(world)]
        out = inverse_cross_stages(rc, world, rank, log_n + 2, w_inv)
        pieces = torch.from_numpy(ints_to_limbs([x for c in range(world) for x in out[c]]).view(np.int64).copy())
        gathered = [torch.empty_like(pieces) for _ in range(world)]
        td.all_gather(gathered, pieces)                                # all-gather of the coefficients
        coeffs = [0] * M
        for r in range(world):
            vals = limbs_to_ints(gathered[r].numpy().view(np.uint64))
            for c in range(world):
                coeffs[c * m + r * per:c * m + (r + 1) * per] = vals[c * per:(c + 1) * per]
        # inverse of the LDE of a degree < n polynomial: the coefficients, then zeros
        ret[rank] = coeffs[:1 << log_n] == coef and not any(coeffs[1 << log_n:])
    finally:
        td.destroy_process_group()


def test_coset_sharded_lde_and_split_inverse_over_gloo_two_processes():
    import torch.multiprocessing as mp
    ret = mp.Manager().dict()
    mp.spawn(_gloo_worker, args=(2, 29641, ret), nprocs=2, join=True)
    assert dict(ret) == {0: True, 1: True}
