"""Size sweep on one GPU for the BASELINE.md table: NTT, MSM (uniform and witness-like scalars), EC-iNTT, and full
proofs at 2^22 / 2^24 checked with the trapdoor verifier (the oracle prover would need tens of minutes there)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    from plonkit_b200 import _lib, plonk, reader, synth
    from plonkit_b200.bn254 import ints_to_limbs
    from oracle import oracle as orc
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
    ctx = _lib.Context(0)
    print("fieldmul Gmul/s: Fr %.1f Fq %.1f" % (ctx.bench_fieldmul(0), ctx.bench_fieldmul(1)), flush=True)
    for lg in (10, 14, 16, 18, 20, 22, 24, 26):
        ms = ctx.bench_ntt(lg, 5)
        print("NTT 2^%d: %.3f ms  %.2f Gelem/s  (HBM-roofline frac %.3f at 64 B/elem)" % (lg, ms, (1 << lg) / ms / 1e6, 64.0 * (1 << lg) / (ms * 1e-3) / 6540.2e9), flush=True)
    max_lg = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    srs = ctx.srs_gen(1 << max_lg, 42)
    for lg in (16, 18, 20, 22, 24):
        if lg > max_lg:
            break
        ctx.srs_load_g1(srs[: 1 << lg])
        ms = ctx.bench_msm(1 << lg, 3)
        print("MSM 2^%d uniform: %.3f ms  %.1f Mscalar/s  (HBM-roofline frac %.4f at 96 B/pair)" % (lg, ms, (1 << lg) / ms / 1e3, 96.0 * (1 << lg) / (ms * 1e-3) / 6540.2e9), flush=True)
        if lg == 20:
            n = 1 << lg
            s = synth.random_field_elements(n, seed=3)
            s[: int(0.4 * n)] = 0
            s[int(0.4 * n): int(0.5 * n)] = ints_to_limbs([1])[0]
            ctx.msm_g1(s)
            t = time.perf_counter(); ctx.msm_g1(s); dt = time.perf_counter() - t
            print("MSM 2^20 witness-like (40%% zero, 10%% one), through host buffers: %.2f ms" % (dt * 1e3), flush=True)
    for lg in (10, 12, 14, 16, 18, 20):
        ctx.srs_load_g1(srs[: 1 << lg])
        ctx.ec_intt_g1(lg)  # first call at a size builds its twiddle table
        t = time.perf_counter(); ctx.ec_intt_g1(lg); dt = time.perf_counter() - t
        print("EC-iNTT (dump-lagrange) 2^%d: %.1f ms  (HBM-roofline frac %.6f at 128 B/point)" % (lg, dt * 1e3, 128.0 * (1 << lg) / dt / 6540.2e9), flush=True)
    # restated CPU baseline (oracle port, all usable host cores) for the same primitives, bounded sizes
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import effective_cores
    cores = effective_cores()
    x = synth.random_field_elements(1 << 20, seed=5)
    t = time.perf_counter(); orc.ntt(x, threads=cores); dt = time.perf_counter() - t
    print("CPU (oracle port, %d threads) NTT 2^20: %.1f ms  %.4f Gelem/s" % (cores, dt * 1e3, (1 << 20) / dt / 1e9), flush=True)
    t = time.perf_counter(); orc.lde4(x, threads=cores); dt = time.perf_counter() - t
    print("CPU (oracle port, %d threads) LDE4 2^20: %.1f ms" % (cores, dt * 1e3), flush=True)
    t = time.perf_counter(); orc.msm(x, srs[: 1 << 20], threads=cores); dt = time.perf_counter() - t
    print("CPU (oracle port, %d threads) MSM 2^20: %.1f ms  %.3f Mscalar/s" % (cores, dt * 1e3, (1 << 20) / dt / 1e6), flush=True)
    t = time.perf_counter(); orc.ec_intt(srs[: 1 << 12], threads=cores); dt = time.perf_counter() - t
    print("CPU (oracle port, %d threads) EC-iNTT 2^12: %.1f ms" % (cores, dt * 1e3), flush=True)
    for lg in (22, 24):
        if lg > max_lg:
            break
        t = time.perf_counter()
        asm = synth.poseidon_chain_assembly(lg)
        key = reader.Crs(srs[: 1 << lg], b"")
        setup = plonk.SetupForProver.prepare_setup_for_prover(asm, key, None, ctx=ctx)
        prep = time.perf_counter() - t
        setup.upload_witness(asm)
        setup.prove(None)
        ctx.timer_begin(); p = setup.prove(None); ms = ctx.timer_end()
        vk = setup.make_verification_key()
        com = np.concatenate([vk.selector_commitments, vk.next_step_selector_commitments, vk.permutation_commitments])
        ok = orc.verify_trapdoor(p.to_bytes(), com, 42)
        print("prove 2^%d: %.1f ms (prep %.1f s), trapdoor-verifies: %s, phases %s" % (lg, ms, prep, ok, [round(x, 1) for x in ctx.profile()["phase_ms"]]), flush=True)
        setup.close()


if __name__ == "__main__":
    main()
