"""Throughput with several independent provers in flight on ONE GPU (each its own library context, stream, SRS tables and
setup): while one proof sits in its latency-bound kernels (sort, scans, transcript round trips) another one's
integer-bound kernels fill the SMs.   python tools/concurrent_prove.py [log_n] [provers] [proofs per prover]"""
import os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    provers = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 6
    from plonkit_b200 import _lib, plonk, reader, synth
    ctxs, setups = [], []
    asm = synth.poseidon_chain_assembly(log_n)
    srs = None
    for k in range(provers):
        ctx = _lib.Context(0)
        if srs is None:
            srs = ctx.srs_gen(1 << log_n, 42)
        setup = plonk.SetupForProver.prepare_setup_for_prover(asm, reader.Crs(srs, b""), None, ctx=ctx)
        setup.upload_witness(asm)
        ref = setup.prove(None).to_bytes()
        setup.prove(None)
        ctxs.append(ctx); setups.append(setup)
    for active in range(1, provers + 1):
        outs = [None] * active
        def work(k):
            for _ in range(reps):
                outs[k] = setups[k].prove(None).to_bytes()
        th = [threading.Thread(target=work, args=(k,)) for k in range(active)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        assert all(o == ref for o in outs)
        print("%d prover(s) in flight: %.2f proofs/s (%.1f ms per proof per prover)" % (active, active * reps / dt, dt / reps * 1e3), flush=True)


if __name__ == "__main__":
    main()
