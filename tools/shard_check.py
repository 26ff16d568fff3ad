#!/usr/bin/env python3
"""One proof sharded over the GPUs of a torchrun launch (one process per GPU, NCCL): checks that every rank's proof equals
the single-GPU prover's bytes and times both.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/shard_check.py --log-n 22
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as td  # noqa: E402

from plonkit_b200 import _lib, plonk, reader, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--log-n", type=int, default=16)
    ap.add_argument("--kind", default="poseidon", choices=["poseidon", "random"])
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--skip-single", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    td.init_process_group("nccl", device_id=torch.device("cuda", local))
    ids = [_lib.nccl_unique_id() if rank == 0 else None]
    td.broadcast_object_list(ids, src=0)
    ctx = _lib.Context(local)
    ctx.attach_nccl(ids[0], rank, world)
    t0 = time.time()
    if args.kind == "poseidon":
        asm = synth.poseidon_chain_assembly(args.log_n)
    else:  # BASELINE configs[2]: random wires
        asm = synth.random_gate_assembly_layered(args.log_n, seed=7)
    gen = _lib.Context(local)
    srs = gen.srs_gen(asm.n, 42)
    key = reader.Crs(srs)
    if rank == 0:
        print("circuit + SRS 2^%d: %.1f s" % (args.log_n, time.time() - t0), flush=True)
    want = None
    if not args.skip_single:
        single = plonk.SetupForProver.prepare_setup_for_prover(asm, key, None, ctx=gen)
        single.upload_witness(asm.var_values)
        want = single.prove(None).to_bytes()
        gen.timer_begin()
        for _ in range(args.iters):
            single.prove(None)
        ms_single = gen.timer_end() / args.iters
        single.close()
    gen.close()
    sp = plonk.ShardedSetupForProver.prepare_setup_for_prover(asm, key, ctx)
    sp.upload_witness(asm.var_values)
    got = sp.prove(None).to_bytes()
    td.barrier(device_ids=[local])
    torch.cuda.synchronize(local)
    ctx.timer_begin()
    for _ in range(args.iters):
        sp.prove(None)
    ms = torch.tensor([ctx.timer_end() / args.iters], dtype=torch.float64, device="cuda:%d" % local)
    td.all_reduce(ms, op=td.ReduceOp.MAX)
    phases = ctx.profile()["phase_ms"]
    ctx.profile_enable(True)     # one more proof with CUDA events around the bulk collectives
    sp.prove(None)
    prof = ctx.profile()
    ctx.profile_enable(False)
    ok = torch.tensor([int(want is None or got == want)], device="cuda:%d" % local)
    td.all_reduce(ok, op=td.ReduceOp.MIN)
    if rank == 0:
        print("2^%d gates on %d GPUs: sharded == single GPU: %s" % (args.log_n, world, bool(ok.item()) if want is not None else "skipped"))
        print("sharded proof %.2f ms (max over ranks)%s" % (ms.item(), "" if want is None else "; single GPU %.2f ms; speed-up %.2fx" % (ms_single, ms_single / ms.item())))
        print("phase ms (rank 0): h2d %.2f | wires %.2f | Z %.2f | quotient %.2f | evals %.2f | openings %.2f" % tuple(phases[:6]))
        names = ("all-gather wire+Z coefficients", "all-to-all quotient", "all-gather quotient coefficients")
        fused = os.environ.get("PK_DIST_FUSED", "1") != "0"
        print("bulk exchanges (CUDA events on the library stream; the all-to-all is %s):"
              % ("FUSED into the last block-local NTT pass as peer stores over NVLink - what is timed is the stream barrier after it"
                 if fused else "a collective after the NTT pass"))
        for k in range(3):
            ms_k, by = prof["comm_ms"][k], prof["comm_bytes"][k]
            rate = "" if (fused and k == 1) else "  -> %6.1f GB/s per rank" % (by / 1e6 / max(ms_k, 1e-9))
            print("  %-34s %8.3f ms  %9.1f MB received per rank%s" % (names[k], ms_k, by / 1e6, rate))
    sp.close()
    td.barrier(device_ids=[local])
    td.destroy_process_group()
    return 0 if ok.item() else 1


if __name__ == "__main__":
    sys.exit(main())
