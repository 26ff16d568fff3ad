"""Multi-GPU sharding of the commitment step (SURVEY.md §8e, row "MSM").

One process per GPU (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).  Every rank keeps
the fixed-base window tables of ITS contiguous chunk of the SRS only, computes the partial multi-scalar
multiplication of that chunk on its GPU, and the 64-byte affine partial sums are all-gathered and folded locally
(`pk_g1_sum`): EC addition is not an NCCL reduce-op, so the "all-reduce" of the north star is all-gather + fold.
Traffic per commitment: world_size x 64 B.
"""
import numpy as np

from . import _lib


def chunk_bounds(n: int, world: int, rank: int):
    """Contiguous chunk [lo, hi) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ShardedCommitter:
    """commit(p) = sum_i p_i * SRS_i with the SRS sharded by base chunk across the ranks of `group`."""

    def __init__(self, bases, rank: int, world: int, ctx=None, group=None, local_msm=None, device=None):
        bases = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
        self.n = bases.shape[0]
        self.rank, self.world, self.group = rank, world, group
        self.lo, self.hi = chunk_bounds(self.n, world, rank)
        self.device = device
        if local_msm is not None:
            # test hook: a stand-in for the device MSM on the local chunk (the CPU tests pass the checker here)
            chunk = bases[self.lo:self.hi]
            self._local = lambda scalars: local_msm(scalars, chunk)
        else:
            if ctx is None:
                raise _lib.LibraryMissing("ShardedCommitter needs a CUDA context (there is no CPU fallback)")
            self.ctx = ctx
            if self.hi > self.lo:
                ctx.srs_load_g1(bases[self.lo:self.hi])
            self._local = lambda scalars: ctx.msm_g1(scalars) if self.hi > self.lo else np.zeros(8, dtype=np.uint64)

    def commit(self, scalars) -> np.ndarray:
        s = np.ascontiguousarray(scalars, dtype=np.uint64).reshape(-1, 4)
        if s.shape[0] != self.n:
            raise _lib.SynthesisError(1, "scalar vector length %d != SRS length %d" % (s.shape[0], self.n))
        partial = np.ascontiguousarray(self._local(s[self.lo:self.hi]), dtype=np.uint64).reshape(8)
        return _lib.g1_sum(self._all_gather(partial))

    def _all_gather(self, partial: np.ndarray) -> np.ndarray:
        if self.world == 1:
            return partial.reshape(1, 8)
        import torch
        import torch.distributed as td
        t = torch.from_numpy(partial.view(np.int64).copy())
        if self.device is not None:
            t = t.to(self.device)
        out = [torch.empty_like(t) for _ in range(self.world)]
        td.all_gather(out, t, group=self.group)
        return np.stack([o.cpu().numpy().view(np.uint64) for o in out])
