"""CPU tests (gloo, world_size 2 and 3) of the multi-GPU host logic: chunking, all-gather of partial sums and the
local fold through the product's own pk_g1_sum.  The device MSM of each chunk is replaced by the checker here; the
GPU test in tests/test_gpu_dist.py runs the same class with the CUDA MSM."""
import os

import numpy as np
import pytest

from conftest import ROOT, SIMPLE
from plonkit_b200 import _lib, dist, reader, synth


def test_chunk_bounds_partition_everything():
    for n in (1, 7, 8, 1000, 1 << 20):
        for world in (1, 2, 3, 8):
            edges = [dist.chunk_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1


def test_g1_sum_host_fold(orc, simple_key):
    g = simple_key.g1_bases
    assert (_lib.g1_sum(g[:5]) == orc.msm(np.tile(np.array([1, 0, 0, 0], dtype=np.uint64), (5, 1)), g[:5])).all()
    inf = np.zeros(8, dtype=np.uint64)
    assert (_lib.g1_sum(np.stack([inf, g[3], inf])) == g[3]).all()
    assert not _lib.g1_sum(np.zeros((0, 8), dtype=np.uint64)).any()
    neg = g[4].copy()
    from plonkit_b200.bn254 import Q_MOD, ints_to_limbs, limbs_to_ints
    neg[4:] = ints_to_limbs([Q_MOD - limbs_to_ints(g[4][4:])[0]])[0]
    assert not _lib.g1_sum(np.stack([g[4], neg])).any()
    assert (_lib.g1_sum(np.stack([g[4], g[4]])) == orc.g1_add(g[4], g[4])).all()


def _worker(rank, world, port, ret):
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, ROOT)
        from oracle import oracle as orc
        key = reader.load_key_monomial_form(os.path.join(SIMPLE, "setup_2^10.key"))
        n = 1001  # odd: uneven chunks
        bases = key.g1_bases[:n]
        committer = dist.ShardedCommitter(bases, rank, world, local_msm=lambda s, b: orc.msm(s, b, threads=2))
        s = synth.random_field_elements(n, seed=77)
        s[::9] = 0
        got = committer.commit(s)
        want = orc.msm(s, bases, threads=2)
        ret[rank] = bool((got == want).all())
        # wrong length -> AssignmentMissing on every rank, before any collective is entered
        try:
            committer.commit(s[:10])
            ret[rank] = False
        except _lib.SynthesisError as e:
            ret[rank] = ret[rank] and e.code == 1
    finally:
        td.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_commit_over_gloo(world):
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29610 + world
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert [ret[r] for r in range(world)] == [True] * world


class _HostNttOps:
    """Checker-backed stand-in for CudaNttOps (canonical limbs throughout, so enter/leave are no-ops)."""

    def __init__(self, orc):
        self.orc = orc

    def enter(self, t):
        pass

    def leave(self, t):
        pass

    def ntt_rows(self, t):
        a = t.numpy().view(np.uint64)
        for r in range(a.shape[0]):
            a[r] = self.orc.ntt(a[r])

    def twiddle(self, t, log_total, row0):
        from plonkit_b200.bn254 import R_MOD, ints_to_limbs, limbs_to_ints, root_of_unity
        a = t.numpy().view(np.uint64)
        w = root_of_unity(log_total)
        for r in range(a.shape[0]):
            base = pow(w, row0 + r, R_MOD)
            vals = limbs_to_ints(a[r])
            a[r] = ints_to_limbs([v * pow(base, c, R_MOD) % R_MOD for c, v in enumerate(vals)])


def _ntt_worker(rank, world, port, ret):
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, ROOT)
        from oracle import oracle as orc
        ok = True
        for log_n, log_n1 in ((6, 3), (7, 3), (8, 5)):
            x = synth.random_field_elements(1 << log_n, seed=500 + log_n)
            d = dist.DistributedNtt(log_n, rank, world, _HostNttOps(orc), log_n1=log_n1)
            out = d.forward(d.local_input(x))
            ok = ok and bool((d.gather_natural(out) == orc.ntt(x)).all())
        ret[rank] = ok
    finally:
        td.destroy_process_group()


def test_four_step_distributed_ntt_over_gloo():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ntt_worker, args=(2, 29633, ret), nprocs=2, join=True)
    assert [ret[r] for r in range(2)] == [True, True]


def test_four_step_single_rank_matches_plain_ntt(orc):
    x = synth.random_field_elements(1 << 6, seed=9)
    d = dist.DistributedNtt(6, 0, 1, _HostNttOps(orc))
    assert (d.gather_natural(d.forward(d.local_input(x))) == orc.ntt(x)).all()
    with pytest.raises(_lib.SynthesisError):
        dist.DistributedNtt(3, 0, 4, _HostNttOps(orc))  # factors 4 x 2 are not both divisible by 4 ranks


class _HostEcOps:
    """Checker-backed stand-in for CudaEcNttOps: points stay affine, the row transforms are the oracle's (scaled) EC
    inverse NTT, so `leave` has nothing left to scale."""

    def __init__(self, orc):
        self.orc = orc

    def enter(self, t):
        return t.clone()

    def ntt_rows(self, t):
        a = t.numpy().view(np.uint64)
        for r in range(a.shape[0]):
            a[r] = self.orc.ec_intt(a[r])

    def twiddle(self, t, log_total, row0):
        from plonkit_b200.bn254 import R_MOD, root_of_unity
        a = t.numpy().view(np.uint64)
        winv = pow(root_of_unity(log_total), -1, R_MOD)
        for r in range(a.shape[0]):
            for c in range(a.shape[1]):
                a[r, c] = self.orc.g1_mul(a[r, c], pow(winv, (row0 + r) * c, R_MOD))

    def leave(self, t, log_total):
        return t


def _ec_worker(rank, world, port, ret):
    import torch.distributed as td
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    td.init_process_group(backend="gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, ROOT)
        from oracle import oracle as orc
        key = reader.load_key_monomial_form(os.path.join(SIMPLE, "setup_2^10.key"))
        ok = True
        for log_n, log_n1 in ((4, 2), (5, 2)):
            bases = key.g1_bases[: 1 << log_n]
            d = dist.DistributedEcIntt(log_n, rank, world, _HostEcOps(orc), log_n1=log_n1)
            out = d.inverse(d.local_input(bases))
            ok = ok and bool((d.gather_natural(out) == orc.ec_intt(bases)).all())
        ret[rank] = ok
    finally:
        td.destroy_process_group()


def test_four_step_distributed_ec_intt_over_gloo():
    """Crs::from_powers split over 2 ranks: column transforms, inverse twiddles, one all-to-all, row transforms."""
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_ec_worker, args=(2, 29644, ret), nprocs=2, join=True)
    assert [ret[r] for r in range(2)] == [True, True]
