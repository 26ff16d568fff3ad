"""One-off parity check at BASELINE.json's full size: the device proof of the 2^log_n-gate poseidon-shaped circuit must be
byte-identical to the oracle's (the -m gpu suite stops at sizes the oracle finishes in seconds and checks 2^20 through the
trapdoor verifier; this tool spends the minute).   python tools/full_size_parity.py [log_n]"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    from bench import effective_cores
    from oracle import oracle as orc
    from plonkit_b200 import _lib, plonk, reader, synth
    cores = effective_cores()
    ctx = _lib.Context(0)
    asm = synth.poseidon_chain_assembly(log_n)
    srs = ctx.srs_gen(asm.n, 42)
    setup = plonk.SetupForProver.prepare_setup_for_prover(asm, reader.Crs(srs, b""), None, ctx=ctx)
    got = setup.prove(asm).to_bytes()
    t0 = time.perf_counter()
    ref = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=cores)
    dt = time.perf_counter() - t0
    print("2^%d gates: device proof sha256 %s" % (log_n, hashlib.sha256(got).hexdigest()))
    print("2^%d gates: oracle proof sha256 %s (%.1f s on %d host threads)" % (log_n, hashlib.sha256(ref).hexdigest(), dt, cores))
    print("byte-identical:", got == ref)
    return 0 if got == ref else 1


if __name__ == "__main__":
    sys.exit(main())
