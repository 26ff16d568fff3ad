#!/usr/bin/env python3
"""Benchmark of the prove path (BASELINE.json metric: proofs/sec on a 2^20-gate poseidon-shaped circuit).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port), host cores

A step = one SetupForProver::prove over one synthetic witness.  `value` is measured with the witness already
resident in HBM; `e2e` goes through the public API with the witness in pinned host memory (H2D copy, proof D2H
inside the timed region).  Timing is by CUDA events on the library's stream, max over ranks.  One JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "proofs/sec (2^20-gate poseidon)"
UNIT = "proofs/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 8 and f[0] == str(self.device):
                self.rows.append(f)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = sorted(float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit())
        reasons = []
        for name, col in (("hw_slowdown", 4), ("hw_thermal_slowdown", 5), ("sw_thermal_slowdown", 6), ("sw_power_cap", 7)):
            if any(r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        mx = float(self.rows[0][2]) if self.rows else None
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    td = None
    if world > 1:
        import torch
        import torch.distributed as td_
        torch.cuda.set_device(local)
        # NCCL / c10d write their banner and (with NCCL_DEBUG=INFO, which the driver may set to count ranks) their log lines to
        # stdout — at communicator creation and lazily later (channel setup, the sharded prover's own communicator).  stdout
        # must carry exactly one JSON line, so fd 1 points at stderr for the whole run and is restored just before that line
        # is printed (emit_json).
        sys.stdout.flush()
        global _SAVED_STDOUT
        _SAVED_STDOUT = os.dup(1)
        os.dup2(2, 1)
        td_.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
        td_.barrier(device_ids=[local])
        torch.cuda.synchronize(local)
        td = td_
    return world, rank, local, td


_SAVED_STDOUT = None


def emit_json(obj):
    """The one JSON line of the run, on the real stdout."""
    global _SAVED_STDOUT
    sys.stdout.flush()
    if _SAVED_STDOUT is not None:
        os.dup2(_SAVED_STDOUT, 1)
        os.close(_SAVED_STDOUT)
        _SAVED_STDOUT = None
    print(json.dumps(obj), flush=True)


def max_over_ranks(td, local, x):
    if td is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda:%d" % local)
    td.all_reduce(t, op=td.ReduceOp.MAX)
    return float(t.item())


def barrier(td, local):
    if td is not None:
        import torch
        td.barrier(device_ids=[local])
        torch.cuda.synchronize(local)


def effective_cores():
    """Host cores this process may really use: min(visible CPUs, affinity mask, cgroup CPU quota).  On the GPU boxes
    nproc reports 128 but the container's cgroup quota is 16 CPUs; oversubscribing it makes the CPU leg slower."""
    cores = os.cpu_count() or 1
    try:
        cores = min(cores, len(os.sched_getaffinity(0)))
    except Exception:
        pass
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            cores = max(1, min(cores, int(int(quota) / int(period))))
    except Exception:
        pass
    return cores


def run_cpu_prove(log_n_s, threads, repeats, warm, srs=None, budget_s=None):
    """Times the oracle's CPU prover (restated bellman algorithms) on a poseidon-shaped circuit of 2^log_n_s gates — the
    SAME circuit the GPU arm proves.  `srs`: an existing [42^i]G key to cut from (the timed quantity is the prove call,
    not key generation).  With `budget_s` the run stops early once the next proof would overshoot the budget: steps are
    kept in preference to warm-ups (at least one of each), and the caller prints the TRUE counts.
    Returns (list of (prove_s, setup_lde_s) per timed step, warm-ups actually run)."""
    from oracle import oracle as orc
    from plonkit_b200 import synth
    orc.build()
    asm = synth.poseidon_chain_assembly(log_n_s)
    srs = orc.srs_gen(asm.n, 42, threads=threads) if srs is None else np.ascontiguousarray(srs[: asm.n])
    t_begin = time.perf_counter()

    def one():
        orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=threads)
        _, prove_s, lde_s = orc.last_timings()
        return prove_s, lde_s
    first = one()
    warm_run = 1 if warm >= 1 else 0
    times = [] if warm_run else [first]
    per = time.perf_counter() - t_begin
    if budget_s is not None:
        afford = max(1, int((budget_s - per) / per))           # further proofs that fit
        steps_run = min(repeats - len(times), afford)
        warm_more = max(0, min(warm - warm_run, afford - steps_run))
    else:
        steps_run, warm_more = repeats - len(times), max(0, warm - warm_run)
    for _ in range(warm_more):
        one()
    warm_run += warm_more
    for _ in range(steps_run):
        times.append(one())
    return times, warm_run


def bench_reference(args):
    """The reference's CPU algorithm for the path (the oracle port: bellman's radix-2 FFT, dense Pippenger, the 5-round
    prover, on a persistent thread pool over every usable host core) on the GPU arm's exact workload: each step is ONE
    full SetupForProver::prove of the 2^log_n-gate circuit, nothing is scaled across sizes.  A 2^20 proof costs ~15 s of
    16 cores, so under a time budget fewer than --steps proofs may run; `steps` / `warmup` on the line are what ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = effective_cores()
    times, warm_run = run_cpu_prove(args.log_n, cores, args.steps, args.warmup, budget_s=args.ref_budget_s)
    steps_run = len(times)
    per_step = sum(t[0] for t in times) / steps_run
    per_step_cached = sum(t[0] - t[1] for t in times) / steps_run
    value = 1.0 / per_step
    sample = ("%d full SetupForProver::prove calls of the oracle port on the 2^%d-gate poseidon-shaped circuit of the GPU arm "
              "(no size scaling), %d threads; per call as the reference behaves: the 11 setup-polynomial LDEs are recomputed "
              "(precomputations = None, src/plonk.rs:156), setup polynomials themselves excluded" % (steps_run, args.log_n, cores))
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps_run,
        "warmup": warm_run, "steps_requested": args.steps, "warmup_requested": args.warmup,
        "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64 (4x64-bit Montgomery limbs)", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "cpu_baseline_setup_cached": {"value": 1.0 / per_step_cached, "unit": UNIT, "cores": cores, "kind": "port",
                                      "sample": "same calls with the time of the 11 setup-polynomial LDEs subtracted (what the "
                                                "CPU prover would cost if it cached them like the GPU arm does): "
                                                "speed-up = caching x kernels"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "restated CPU baseline (bellman's algorithms in C++ on a persistent thread pool; batch inversion and grand "
                "product chunk-parallel as in bellman; the Rust reference cannot be built in this image: no cargo/rustc)",
    }
    print(json.dumps(out), flush=True)
    return 0


def workload_config(args, world):
    return {
        "workload": "BASELINE configs[1]: 2^%d SRS ([42^i]G), poseidon-shaped circuit with 2^%d-1 gates (1 public input), "
                    "keccak transcript, single-B200 prove with NTT+MSM on device" % (args.log_n, args.log_n),
        "log_n": args.log_n,
        "parallelism": "1 GPU" if world == 1 else "%d independent provers (one per GPU, no data-path collective)" % world,
        "l2": "per-proof working set ~3 GB >> 126 MB L2: no flush needed between steps",
    }


def sharded_single_proof(args, td, rank, world, local, srs, asm, steps):
    """ONE proof of the same 2^log_n circuit computed by all `world` GPUs together (plonk.ShardedSetupForProver over NCCL:
    commitments by base chunk, quotient by coset, one all-to-all in the size-4n inverse NTT) — the strong-scaling latency
    line next to the replica throughput.  Returns (ms per proof as max over ranks, proof bytes, phase ms of this rank)."""
    import torch
    from plonkit_b200 import _lib, plonk, reader
    ids = [_lib.nccl_unique_id() if rank == 0 else None]
    td.broadcast_object_list(ids, src=0)
    c = _lib.Context(local)
    c.attach_nccl(ids[0], rank, world)
    sp = plonk.ShardedSetupForProver.prepare_setup_for_prover(asm, reader.Crs(srs), c)
    sp.upload_witness(asm.var_values)
    got = None
    for _ in range(3):
        got = sp.prove(None).to_bytes()
    barrier(td, local)
    c.timer_begin()
    for _ in range(steps):
        sp.prove(None)
    ms = max_over_ranks(td, local, c.timer_end() / steps)
    phases = c.profile()["phase_ms"]
    sp.close()
    c.close()
    return ms, got, phases


def bench_ours(args):
    world, rank, local, td = dist_setup(args.gpus)
    from plonkit_b200 import _lib, plonk, reader, synth
    import torch
    torch.cuda.set_device(local)
    n = 1 << args.log_n
    P = max(1, args.inflight)
    t0 = time.time()
    asm = synth.poseidon_chain_assembly(args.log_n, inputs=(3 + rank, 4, 5))
    # P independent provers per GPU (own library context = own stream, SRS window tables, setup, scratch): proofs are
    # independent objects, so while one proof sits in its latency-bound kernels (bucket sort, scans, transcript round
    # trips to the host) the integer-bound kernels of another fill the SMs.  Prover 0 also serves the per-kernel pass.
    ctxs, setups, srs = [], [], None
    for k in range(P):
        c = _lib.Context(local)
        if srs is None:
            srs = c.srs_gen(n, 42)
        ctxs.append(c)
        setups.append(plonk.SetupForProver.prepare_setup_for_prover(asm, reader.Crs(srs, b""), None, ctx=c))
    ctx, setup = ctxs[0], setups[0]
    prep_s = time.time() - t0
    # witness in pinned host memory (what a caller of the public API hands over)
    wit = torch.from_numpy(asm.var_values).pin_memory()
    wit_np = wit.numpy()
    h2d_bytes = int(wit_np.nbytes)
    d2h_bytes = 1144 - 16  # the proof: 9 G1 points + 16 field elements (+ challenges are not copied)

    ref_bytes = None
    for _ in range(args.warmup):
        for s_ in setups:
            ref_bytes = s_.prove(wit_np).to_bytes()

    def run_concurrent(work):
        """Runs work(k) for every prover on its own host thread; device time between two device-wide idle points."""
        torch.cuda.synchronize(local)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        errs = []

        def guarded(k):
            try:
                work(k)
            except Exception as ex:  # surface failures of worker threads
                errs.append(ex)
        th = [threading.Thread(target=guarded, args=(k,)) for k in range(P)]
        e0.record()
        for t in th:
            t.start()
        for t in th:
            t.join()
        torch.cuda.synchronize(local)  # every stream of every prover has drained
        e1.record()
        e1.synchronize()
        if errs:
            raise errs[0]
        return e0.elapsed_time(e1)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- timed region 1: witness resident in HBM; K steps, a step = one proof on each of the P provers
    for s_ in setups:
        s_.upload_witness(wit_np)
    for c in ctxs:
        c.profile_enable(False)
        c.profile_reset()
    last = [None] * P

    def resident(k):
        for _ in range(args.steps):
            last[k] = setups[k].prove(None)
    barrier(td, local)
    ms_dev = run_concurrent(resident)
    barrier(td, local)
    launches_timed = sum(c.profile()["kernel_launches"] for c in ctxs)
    ms_dev = max_over_ranks(td, local, ms_dev)
    # ---- single prover, one proof at a time (latency), and the per-kernel pass (neither is part of `value`): the
    # same K proofs with CUDA events around the dominant kernel
    ctx.timer_begin()
    for _ in range(args.steps):
        proof = setup.prove(None)
    ms_single = ctx.timer_end()
    ctx.profile_enable(True)
    ctx.profile_reset()
    ctx.timer_begin()
    for _ in range(args.steps):
        proof = setup.prove(None)
    ms_prof = ctx.timer_end()
    prof = ctx.profile()
    ctx.profile_enable(False)
    # ---- timed region 2: end to end through the public API, host buffers (H2D of the witness, proof D2H)
    pb = [None] * P

    def host_buffers(k):
        for _ in range(args.steps):
            pb[k] = setups[k].prove(wit_np).to_bytes()
    barrier(td, local)
    ms_e2e = run_concurrent(host_buffers)
    barrier(td, local)
    ms_e2e = max_over_ranks(td, local, ms_e2e)
    clocks = sampler.stop() if rank == 0 else None
    if ref_bytes is not None:
        assert all(b == ref_bytes for b in pb), "proof bytes changed between runs / provers"
    shard = None
    if world > 1 and not args.no_shard:
        # same circuit and witness on every rank (the replica legs above use a different public input per rank)
        asm0 = asm if rank == 0 else synth.poseidon_chain_assembly(args.log_n, inputs=(3, 4, 5))
        for s_ in setups[1:]:   # the extra replica provers are done: free their share of HBM
            s_.close()
        try:
            ms_shard, shard_bytes, shard_phases = sharded_single_proof(args, td, rank, world, local, srs, asm0, args.steps)
            shard = {"ms_per_proof": ms_shard, "proofs_per_s": 1e3 / ms_shard, "n_gpus": world, "transport": "nccl",
                     "bytes_equal_single_gpu_proof": bool(shard_bytes == ref_bytes) if rank == 0 else None,
                     "phase_ms_rank0": shard_phases[:6],
                     "what": "ONE proof of the same circuit computed by all %d GPUs together (strong scaling, latency): commitments "
                             "sharded by base chunk (all-gather of 128-byte partial sums + device fold), quotient sharded by coset, one "
                             "all-to-all inside the size-4n inverse NTT; not part of `value`" % world}
        except Exception as ex:  # a secondary figure must never cost the headline line
            shard = {"error": "%s: %s" % (type(ex).__name__, ex)}

    if rank != 0:
        return 0
    value = world * P * args.steps / (ms_dev * 1e-3)
    e2e = world * P * args.steps / (ms_e2e * 1e-3)
    peak, peak_src = peaks()
    launches = prof["msm_accum_launches"]
    accum_ms = prof["msm_accum_ms"] / max(launches, 1)
    # 32 B scalar + 64 B base per (scalar, base) pair; a launch covers a batch of up to 4 MSMs of N pairs
    alg_bytes = 96.0 * prof["msm_accum_points"] / max(launches, 1)
    achieved = alg_bytes / (accum_ms * 1e-3) / 1e9 if launches else 0.0
    traffic = None
    summ = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(summ):
        try:
            # ncu --set full capture of this kernel: DRAM bytes per (scalar, base) pair x pairs of an average launch
            traffic = json.load(open(summ)).get("msm_accum_dram_bytes_per_pair") * prof["msm_accum_points"] / max(launches, 1)
        except Exception:
            traffic = None
    cfg = workload_config(args, world)
    cfg["proofs_in_flight_per_gpu"] = P
    cfg["step"] = "one proof on each of the %d provers of a GPU (%d proofs per step per GPU)" % (P, P)
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 (8x32-bit Montgomery limbs, 254-bit modular integer arithmetic)", "data": "synthetic",
        "config": cfg,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * P, "d2h_bytes_per_step": d2h_bytes * P,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches_timed),
        "single_prover": {"ms_per_proof": ms_single / args.steps, "proofs_per_s": args.steps / (ms_single * 1e-3),
                          "note": "one proof at a time on one stream (latency); `value` keeps %d proofs in flight" % P},
        "roofline": {"bound": "hbm", "kernel": "msm_accum_kernel<true> (bucket accumulation; 4 batched launches cover the 11 MSMs of a proof)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "peak_source": peak_src, "alg_bytes_per_launch": alg_bytes, "ms_per_launch": accum_ms,
                     "share_of_step": prof["msm_accum_ms"] / ms_prof, "ms_per_step_profiled": ms_prof / args.steps,
                     "note": "integer-multiply-pipe bound (10 x 254-bit Montgomery products per pair and window; 89 % of the measured IMAD.WIDE issue ceiling, DESIGN.md 4.5); timed by CUDA events on the library stream in a separate single-prover pass; share_of_step is relative to that pass"},
        "ntt": {"ms_per_step": prof["ntt_ms"] / args.steps, "launches_per_step": prof["ntt_launches"] / args.steps},
        "phase_ms": prof["phase_ms"],
        "clocks": clocks,
        "prep_s": prep_s,
    }
    if shard is not None:
        if "ms_per_proof" in shard:
            shard["speedup_vs_one_gpu"] = (ms_single / args.steps) / shard["ms_per_proof"]
        out["sharded_single_proof"] = shard
    # BASELINE.json's secondary metrics on the same GPU: single-set MSM (uniform scalars) and NTT at the bench size
    try:
        ms_msm, ms_ntt = ctx.bench_msm(n, 3), ctx.bench_ntt(args.log_n, 5)
        out["micro"] = {"msm_mscalar_per_s": n / ms_msm / 1e3, "msm_ms": ms_msm, "ntt_gelem_per_s": n / ms_ntt / 1e6, "ntt_ms": ms_ntt,
                        "msm_hbm_frac_at_96B_per_pair": 96.0 * n / (ms_msm * 1e-3) / 1e9 / peak,
                        "ntt_hbm_frac_at_64B_per_elem": 64.0 * n / (ms_ntt * 1e-3) / 1e9 / peak, "log_n": args.log_n}
    except Exception as ex:  # secondary figures must never cost the headline line
        out["micro"] = {"error": str(ex)}
    if world == 1 and not args.no_cpu:
        cores = effective_cores()
        # bounded sample: ONE proof of the same 2^log_n circuit (~15 s of 16 cores at 2^20), no warm-up, no scaling
        t, _ = run_cpu_prove(args.log_n, cores, 1, 0, srs=srs)
        sample = ("one full prove of the oracle port on the same 2^%d-gate circuit, %d threads (setup polynomials excluded, "
                  "their 11 LDEs recomputed as the reference does)" % (args.log_n, cores))
        out["cpu_baseline"] = {"value": 1.0 / t[0][0], "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
        out["cpu_baseline_setup_cached"] = {"value": 1.0 / (t[0][0] - t[0][1]), "unit": UNIT, "cores": cores, "kind": "port",
                                            "sample": "same call minus the 11 setup-polynomial LDEs"}
    else:
        out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": effective_cores(), "kind": "port",
                               "sample": "not run (rank 0 at N = 1 only, and not with --no-cpu)"}
    emit_json(out)
    if td is not None:
        os.dup2(2, 1)  # late NCCL teardown chatter must not follow the JSON line
        td.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-n", type=int, default=20)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--inflight", type=int, default=3, help="independent provers (proofs in flight) per GPU")
    ap.add_argument("--no-shard", action="store_true", help="N > 1: skip the sharded single-proof (latency) leg")
    ap.add_argument("--ref-budget-s", type=float, default=360.0, help="--impl reference: wall-clock budget for the CPU proofs")
    args = ap.parse_args()
    if args.warmup < 1:
        args.warmup = 1
    if args.impl == "reference":
        return bench_reference(args)
    return bench_ours(args)


if __name__ == "__main__":
    sys.exit(main())
