"""Multi-GPU building blocks driven from Python (torch.distributed; NCCL over NVLink on the GPU box, gloo in the CPU tests).

The PROVER's multi-GPU path is not here: one proof cut across the GPUs of a node is `plonk.ShardedSetupForProver` /
`plonk.ShardedProver` over the C ABI's `pk_dist_*` calls (csrc/dist_prover.cu), whose collectives the library issues itself.

What lives here is `dump-lagrange` across GPUs — `DistributedEcIntt` / `lagrange_key_distributed`: `Crs::from_powers`
(src/plonk.rs:179-185) as a four-step EC inverse NTT with one all-to-all of XYZZ points — and the two round-1 primitives it
grew out of, kept as tested building blocks: `ShardedCommitter` (commitment with the SRS sharded by base chunk, partial sums
all-gathered and folded with `pk_g1_sum`) and `DistributedNtt` (four-step field NTT with one all-to-all).
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


class CudaNttOps:
    """Local steps of the distributed NTT on caller-owned CUDA tensors through the library's device-pointer entry points
    (pk_dev_fr_convert / pk_dev_ntt_rows / pk_dev_twiddle).  Tensors are (rows, cols, 4) int64 views of u64 limbs."""

    def __init__(self, ctx, device):
        self.ctx, self.device = ctx, device

    def _sync(self):
        import torch
        torch.cuda.synchronize(self.device)  # the library runs on its own stream

    def enter(self, t):
        self._sync()
        self.ctx.dev_fr_convert(t.data_ptr(), t.shape[0] * t.shape[1], True)

    def leave(self, t):
        self._sync()
        self.ctx.dev_fr_convert(t.data_ptr(), t.shape[0] * t.shape[1], False)

    def ntt_rows(self, t):
        self._sync()
        self.ctx.dev_ntt_rows(t.data_ptr(), t.shape[1].bit_length() - 1, t.shape[0], False)

    def twiddle(self, t, log_total, row0):
        self._sync()
        self.ctx.dev_twiddle(t.data_ptr(), t.shape[0], t.shape[1], log_total, row0, False)


class DistributedNtt:
    """Four-step forward NTT of size N = N1 * N2 over `world` ranks with ONE all-to-all (SURVEY.md §8e, row "NTT").

    Input  (column blocks): rank r holds x[N2*n1 + r*C + j] for n1 < N1, j < C = N2/world, as an (N1, C, 4) array.
    Output (row blocks)   : rank q holds X[(q*K + i) + N1*k2] for i < K = N1/world, k2 < N2, as a (K, N2, 4) array.
    Steps: local column NTTs of size N1 -> twiddle w_N^(n2*k1) -> all-to-all transpose (N*32/world bytes leave each
    rank) -> local row NTTs of size N2.  `ops` supplies the local steps (CudaNttOps on the GPU; the CPU tests plug in
    the checker); torch moves the data (transposes, packing, all_to_all_single over NCCL or gloo).
    """

    def __init__(self, log_n, rank, world, ops, group=None, log_n1=None):
        self.log_n, self.rank, self.world, self.ops, self.group = log_n, rank, world, ops, group
        self.log_n1 = log_n1 if log_n1 is not None else (log_n + 1) // 2
        self.n1, self.n2 = 1 << self.log_n1, 1 << (log_n - self.log_n1)
        if self.n1 % world or self.n2 % world:
            raise _lib.SynthesisError(6, "both NTT factors must be divisible by the number of ranks")
        self.c, self.k = self.n2 // world, self.n1 // world

    def local_input(self, x_full):
        """The (N1, C, 4) block of this rank, cut from a full natural-order vector (helper for tests / bench)."""
        x = np.ascontiguousarray(x_full, dtype=np.uint64).reshape(self.n1, self.n2, 4)
        return np.ascontiguousarray(x[:, self.rank * self.c:(self.rank + 1) * self.c])

    def forward(self, local):
        import torch
        import torch.distributed as td
        t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local).view(np.int64))
        n2, c, k, w = self.n2, self.c, self.k, self.world
        cols = t.transpose(0, 1).contiguous()                  # (C, N1): this rank's columns as rows
        self.ops.enter(cols)
        self.ops.ntt_rows(cols)                                 # over n1 -> k1
        self.ops.twiddle(cols, self.log_n, self.rank * c)       # * w_N^(n2 * k1)
        send = cols.view(c, w, k, 4).permute(1, 0, 2, 3).contiguous()   # (world, C, K): block q goes to rank q
        recv = torch.empty_like(send)
        if w == 1:
            recv.copy_(send)
        else:
            td.all_to_all_single(recv, send, group=self.group)
        rows = recv.view(n2, k, 4).transpose(0, 1).contiguous()         # (K, N2): my k1 rows, all n2
        self.ops.ntt_rows(rows)                                 # over n2 -> k2
        self.ops.leave(rows)
        return rows

    def gather_natural(self, rows):
        """All ranks' outputs reassembled into the natural-order vector X (test helper; O(N) traffic)."""
        import torch
        import torch.distributed as td
        parts = [torch.empty_like(rows) for _ in range(self.world)]
        if self.world == 1:
            parts[0].copy_(rows)
        else:
            td.all_gather(parts, rows, group=self.group)
        full = torch.cat(parts, dim=0)                          # (N1, N2): [k1][k2]
        return full.transpose(0, 1).contiguous().cpu().numpy().view(np.uint64).reshape(-1, 4)  # index k1 + N1*k2


class CudaEcNttOps:
    """Local steps of the distributed EC inverse NTT on caller-owned CUDA tensors (pk_dev_ec_*).  Points travel as
    128-byte XYZZ accumulators ((..., 16) int64 views) between `enter` (affine -> XYZZ) and `leave` (normalise to affine);
    the 1/N of the inverse transform rides on the twiddle step, which multiplies every point anyway."""

    def __init__(self, ctx, device):
        self.ctx, self.device = ctx, device

    def _sync(self):
        import torch
        torch.cuda.synchronize(self.device)

    def enter(self, affine):
        import torch
        self._sync()
        n = affine.numel() // 8
        out = torch.empty(affine.shape[:-1] + (16,), dtype=torch.int64, device=affine.device)
        self.ctx.dev_ec_from_affine(affine.data_ptr(), out.data_ptr(), n)
        return out

    def ntt_rows(self, t):
        self._sync()
        self.ctx.dev_ec_ntt_rows(t.data_ptr(), t.shape[1].bit_length() - 1, t.shape[0], True)

    def twiddle(self, t, log_total, row0):
        self._sync()
        self.ctx.dev_ec_twiddle(t.data_ptr(), t.shape[0], t.shape[1], log_total, row0, 2)      # inverse twiddles times 1/N

    def leave(self, t, log_total):
        import torch
        self._sync()
        out = torch.empty(t.shape[:-1] + (8,), dtype=torch.int64, device=t.device)
        self.ctx.dev_ec_to_affine(t.data_ptr(), out.data_ptr(), t.numel() // 16, 0)
        return out


class DistributedEcIntt:
    """Crs::from_powers (src/plonk.rs:179-185) of a 2^log_n monomial key over `world` ranks: the four-step INVERSE NTT over
    G1 points with ONE all-to-all (SURVEY.md §8e row "EC-iNTT", BASELINE configs[3]).

    Input  (column blocks): rank r holds the monomial bases [tau^(N2*n1 + r*C + j)]G, n1 < N1, j < C = N2/world, as an
                            (N1, C, 8) array of affine points (canonical limbs).
    Output (row blocks)   : rank q holds [L_i(tau)]G for i = (q*K + a) + N1*k2, a < K = N1/world, k2 < N2, as (K, N2, 8).
    Every butterfly and every twiddle is a 254-bit scalar multiplication, so the work per rank is (N/world) * (log2(N)/2 + 1)
    of them; the all-to-all moves N * 128 B / world per rank once.
    """

    def __init__(self, log_n, rank, world, ops, group=None, log_n1=None):
        self.log_n, self.rank, self.world, self.ops, self.group = log_n, rank, world, ops, group
        self.log_n1 = log_n1 if log_n1 is not None else (log_n + 1) // 2
        self.n1, self.n2 = 1 << self.log_n1, 1 << (log_n - self.log_n1)
        if self.n1 % world or self.n2 % world:
            raise _lib.SynthesisError(6, "both NTT factors must be divisible by the number of ranks")
        self.c, self.k = self.n2 // world, self.n1 // world

    def local_input(self, bases_full):
        b = np.ascontiguousarray(bases_full, dtype=np.uint64).reshape(self.n1, self.n2, 8)
        return np.ascontiguousarray(b[:, self.rank * self.c:(self.rank + 1) * self.c])

    def inverse(self, local):
        import torch
        import torch.distributed as td
        t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local).view(np.int64))
        n2, c, k, w = self.n2, self.c, self.k, self.world
        pts = self.ops.enter(t.transpose(0, 1).contiguous())               # (C, N1, E): this rank's columns as rows
        e = pts.shape[-1]
        self.ops.ntt_rows(pts)                                             # over n1 -> k1 (inverse, unscaled)
        self.ops.twiddle(pts, self.log_n, self.rank * c)                   # * w_N^-(n2 * k1) / N
        send = pts.view(c, w, k, e).permute(1, 0, 2, 3).contiguous()       # (world, C, K): block q goes to rank q
        recv = torch.empty_like(send)
        if w == 1:
            recv.copy_(send)
        else:
            td.all_to_all_single(recv, send, group=self.group)
        rows = recv.view(n2, k, e).transpose(0, 1).contiguous()            # (K, N2): my k1 rows, all n2
        self.ops.ntt_rows(rows)                                            # over n2 -> k2
        return self.ops.leave(rows, self.log_n)                            # affine

    def gather_natural(self, rows):
        """All ranks' outputs reassembled into the natural-order Lagrange key (test helper; O(N) traffic)."""
        import torch
        import torch.distributed as td
        rows = rows.contiguous()
        parts = [torch.empty_like(rows) for _ in range(self.world)]
        if self.world == 1:
            parts[0].copy_(rows)
        else:
            td.all_gather(parts, rows, group=self.group)
        full = torch.cat(parts, dim=0)                                     # (N1, N2): [k1][k2]
        return full.transpose(0, 1).contiguous().cpu().numpy().view(np.uint64).reshape(-1, 8)  # index k1 + N1*k2


def lagrange_key_distributed(bases, ctx, rank, world, device, group=None):
    """`Crs::from_powers` (src/plonk.rs:179-185) of the first 2^k monomial bases across the ranks of `group`: every rank
    runs its share of the four-step EC inverse NTT on its GPU and all ranks end up with the full Lagrange key (natural
    order, (N, 8) canonical limbs) — what `plonkit dump-lagrange` writes.  One process per GPU under torchrun."""
    import torch
    b = np.ascontiguousarray(bases, dtype=np.uint64).reshape(-1, 8)
    log_n = b.shape[0].bit_length() - 1
    if 1 << log_n != b.shape[0]:
        raise _lib.SynthesisError(6, "the Lagrange key needs a power-of-two number of bases")
    d = DistributedEcIntt(log_n, rank, world, CudaEcNttOps(ctx, device), group=group)
    local = torch.from_numpy(d.local_input(b).view(np.int64)).cuda(device)
    return d.gather_natural(d.inverse(local))
