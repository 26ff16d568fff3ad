"""How the oracle's CPU kernels scale with threads on this host (to interpret cpu_baseline numbers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import oracle as orc
from plonkit_b200 import synth

def main():
    print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
    try:
        print("cgroup cpu.max:", open("/sys/fs/cgroup/cpu.max").read().strip())
    except Exception as e:
        print("cgroup cpu.max: n/a", e)
    lg = 18
    s = synth.random_field_elements(1 << lg, seed=1)
    srs = orc.srs_gen(1 << lg, 42, threads=os.cpu_count())
    for th in (1, 8, 32, 64, 128):
        if th > 2 * (os.cpu_count() or 1):
            continue
        t = time.perf_counter(); orc.msm(s, srs, threads=th); a = time.perf_counter() - t
        t = time.perf_counter(); orc.ntt(s, threads=th); b = time.perf_counter() - t
        print("threads %3d  msm 2^%d %.3fs  ntt 2^%d %.3fs" % (th, lg, a, lg, b), flush=True)
    asm = synth.poseidon_chain_assembly(16)
    srs16 = srs[: 1 << 16]
    for th in (8, 32, 128):
        if th > 2 * (os.cpu_count() or 1):
            continue
        t = time.perf_counter(); orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs16, threads=th)
        print("threads %3d  prove 2^16 %.3fs  (setup_s, prove_s) = %s" % (th, time.perf_counter() - t, orc.last_timings()), flush=True)

if __name__ == "__main__":
    main()
