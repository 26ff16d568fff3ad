"""Quick on-GPU sanity + micro-benchmarks (prints one line per item; never raises)."""
import sys, os, time, traceback
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

def item(name, fn):
    t = time.time()
    try:
        r = fn()
        print("[ok ] %-40s %s  (%.2fs)" % (name, r, time.time() - t), flush=True)
    except Exception as e:
        print("[ERR] %-40s %r" % (name, e), flush=True)
        traceback.print_exc()

def main():
    from plonkit_b200 import _lib, synth, plonk, reader
    from oracle import oracle as orc
    ctx = _lib.Context(0)
    item("fieldmul Fr Gmul/s", lambda: round(ctx.bench_fieldmul(0), 2))
    item("fieldmul Fq Gmul/s", lambda: round(ctx.bench_fieldmul(1), 2))
    x = synth.random_field_elements(1 << 10, seed=1)
    item("ntt 2^10 == oracle", lambda: bool((ctx.ntt(x) == orc.ntt(x)).all()))
    item("intt 2^10 == oracle", lambda: bool((ctx.ntt(x, inverse=True) == orc.ntt(x, inverse=True)).all()))
    key = reader.load_key_monomial_form(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "simple", "setup_2^10.key"))
    item("srs load 2^10", lambda: ctx.srs_load_g1(key.g1_bases))
    item("msm 2^10 == oracle", lambda: bool((ctx.msm_g1(x) == orc.msm(x, key.g1_bases, threads=8)).all()))
    for lg in (16, 20, 22, 24):
        item("bench ntt 2^%d ms" % lg, lambda lg=lg: round(ctx.bench_ntt(lg, 5), 3))
    for lg in (16, 20):
        def f(lg=lg):
            srs = ctx.srs_gen(1 << lg, 42)
            ctx.srs_load_g1(srs)
            return round(ctx.bench_msm(1 << lg, 3), 3)
        item("bench msm 2^%d ms" % lg, f)
    def prove20():
        asm = synth.poseidon_chain_assembly(16)
        srs = ctx.srs_gen(1 << 16, 42)
        setup = plonk.SetupForProver.prepare_setup_for_prover(asm, reader.Crs(srs, b""), None, ctx=ctx)
        p = setup.prove(asm)
        t = time.time(); p2 = setup.prove(None); dt = time.time() - t
        ref = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=8)
        return (p.to_bytes() == ref, p2.to_bytes() == ref, round(dt * 1e3, 2), ctx.profile()["phase_ms"])
    item("prove 2^16 == oracle, ms", prove20)

if __name__ == "__main__":
    main()
