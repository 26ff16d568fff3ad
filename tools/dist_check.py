"""torchrun --nproc-per-node G tools/dist_check.py : multi-GPU checks + timings of the sharded MSM (all-gather + fold)
and of the four-step distributed NTT (one all-to-all) against the single-GPU results.  Prints one line per rank 0 item."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as td


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ["NCCL_DEBUG"] = "WARN"
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    from plonkit_b200 import _lib, dist, synth
    ctx = _lib.Context(local)
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    n = 1 << log_n
    # ---- sharded MSM: every rank generates the same SRS/scalars, keeps its chunk of the window tables
    log_e = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    srs = ctx.srs_gen(max(n, 1 << log_e), 42)
    s = synth.random_field_elements(n, seed=11)
    c = dist.ShardedCommitter(srs[:n], rank, world, ctx=ctx, device="cuda:%d" % local)
    got = c.commit(s)
    td.barrier(device_ids=[local])
    t0 = time.perf_counter()
    for _ in range(3):
        got = c.commit(s)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    if rank == 0:
        ctx.srs_load_g1(srs[:n])
        ref = ctx.msm_g1(s)
        print("sharded MSM 2^%d over %d GPUs: %.2f ms per commitment (incl. H2D of the scalar chunk), == single GPU: %s"
              % (log_n, world, dt * 1e3, bool((got == ref).all())), flush=True)
    td.barrier(device_ids=[local])
    # ---- four-step distributed NTT
    x = synth.random_field_elements(n, seed=12)
    d = dist.DistributedNtt(log_n, rank, world, dist.CudaNttOps(ctx, local))
    loc = torch.from_numpy(d.local_input(x).view(np.int64)).cuda()
    out = d.forward(loc)
    td.barrier(device_ids=[local])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        out = d.forward(loc)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    full = d.gather_natural(out)
    if rank == 0:
        ref = ctx.ntt(x)
        print("four-step NTT 2^%d over %d GPUs: %.2f ms (device-resident, one all-to-all), == single GPU: %s"
              % (log_n, world, dt * 1e3, bool((full == ref).all())), flush=True)
    td.barrier(device_ids=[local])
    # ---- four-step distributed EC inverse NTT (dump-lagrange / Crs::from_powers)
    key = srs[: 1 << log_e]
    de = dist.DistributedEcIntt(log_e, rank, world, dist.CudaEcNttOps(ctx, local))
    loc = torch.from_numpy(de.local_input(key).view(np.int64)).cuda()
    out = de.inverse(loc)
    td.barrier(device_ids=[local])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = de.inverse(loc)
    torch.cuda.synchronize()
    td.barrier(device_ids=[local])
    dt = time.perf_counter() - t0
    if log_e <= 22:
        full = de.gather_natural(out)
        if rank == 0:
            ctx.srs_load_g1(key)
            t0 = time.perf_counter()
            ref = ctx.ec_intt_g1(log_e)
            dt1 = time.perf_counter() - t0
            print("four-step EC-iNTT (dump-lagrange) 2^%d over %d GPUs: %.1f ms (device-resident, one all-to-all); one GPU: %.1f ms; "
                  "== single GPU: %s" % (log_e, world, dt * 1e3, dt1 * 1e3, bool((full == ref).all())), flush=True)
    else:
        # too large to redo on one GPU inside this check: every rank verifies 16 of its own outputs in the exponent,
        # out[i] == L_i(42) * G with L_i(t) = w^i (t^N - 1) / (N (t - w^i))  
        from oracle import oracle as orc
        from plonkit_b200.bn254 import R_MOD, root_of_unity
        N, w = 1 << log_e, root_of_unity(log_e)
        rows = out.cpu().numpy().view(np.uint64)            # (K, N2, 8)
        rng = np.random.default_rng(1000 + rank)
        g = np.zeros(8, dtype=np.uint64); g[0] = 1; g[4] = 2
        ok = True
        for _ in range(16):
            a, k2 = int(rng.integers(de.k)), int(rng.integers(de.n2))
            i = (rank * de.k + a) + de.n1 * k2
            wi = pow(w, i, R_MOD)
            li = wi * (pow(42, N, R_MOD) - 1) % R_MOD * pow(N * (42 - wi) % R_MOD, -1, R_MOD) % R_MOD
            ok = ok and bool((rows[a, k2] == orc.g1_mul(g, li)).all())
        flag = torch.tensor([1 if ok else 0], device="cuda:%d" % local)
        td.all_reduce(flag, op=td.ReduceOp.MIN)
        if rank == 0:
            print("four-step EC-iNTT (dump-lagrange) 2^%d over %d GPUs: %.1f ms (device-resident, one all-to-all); "
                  "16 outputs per rank == L_i(42) G: %s" % (log_e, world, dt * 1e3, bool(flag.item())), flush=True)
    td.destroy_process_group()


if __name__ == "__main__":
    main()
