"""GPU tests of the multi-GPU building blocks on ONE device (world_size 1): the device-pointer NTT primitives behind the
four-step distributed NTT, and the sharded committer with the CUDA MSM.  tools/dist_check.py runs the same classes
under torchrun on 2+ GPUs."""
import numpy as np
import pytest

from plonkit_b200 import dist, synth

pytestmark = pytest.mark.gpu


def test_four_step_ntt_with_device_primitives(ctx, orc):
    import torch
    for log_n, log_n1 in ((8, 4), (13, 6), (16, 8)):
        x = synth.random_field_elements(1 << log_n, seed=600 + log_n)
        d = dist.DistributedNtt(log_n, 0, 1, dist.CudaNttOps(ctx, 0), log_n1=log_n1)
        local = torch.from_numpy(d.local_input(x).view(np.int64)).cuda()
        out = d.forward(local)
        assert (d.gather_natural(out) == orc.ntt(x, threads=8)).all(), log_n


def test_sharded_committer_single_rank(ctx, orc, simple_key):
    c = dist.ShardedCommitter(simple_key.g1_bases[:1000], 0, 1, ctx=ctx)
    s = synth.random_field_elements(1000, seed=3)
    assert (c.commit(s) == orc.msm(s, simple_key.g1_bases[:1000], threads=4)).all()


def test_four_step_ec_intt_with_device_primitives(ctx, orc, simple_key):
    """Crs::from_powers through the device-pointer EC primitives (pk_dev_ec_*) == the oracle's EC inverse NTT and ==
    the single-call pk_ec_intt_g1."""
    import torch
    for log_n, log_n1 in ((3, 1), (6, 3), (9, 5)):
        bases = simple_key.g1_bases[: 1 << log_n]
        d = dist.DistributedEcIntt(log_n, 0, 1, dist.CudaEcNttOps(ctx, 0), log_n1=log_n1)
        local = torch.from_numpy(d.local_input(bases).view(np.int64)).cuda()
        got = d.gather_natural(d.inverse(local))
        assert (got == orc.ec_intt(bases, threads=8)).all(), log_n
    ctx.srs_load_g1(simple_key.g1_bases[:512])
    assert (got == ctx.ec_intt_g1(9)).all()
