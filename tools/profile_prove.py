"""One 2^log_n proof between cudaProfilerStart/Stop, for `ncu --profile-from-start off` captures (profiles/README.md)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    log_n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    from plonkit_b200 import _lib, plonk, reader, synth
    ctx = _lib.Context(0)
    asm = synth.poseidon_chain_assembly(log_n)
    srs = ctx.srs_gen(1 << log_n, 42)
    setup = plonk.SetupForProver.prepare_setup_for_prover(asm, reader.Crs(srs, b""), None, ctx=ctx)
    setup.upload_witness(asm)
    setup.prove(None)  # warm
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaProfilerStart()
    p = setup.prove(None)
    rt.cudaProfilerStop()
    print("proved", len(p.to_bytes()), ctx.profile()["phase_ms"])


if __name__ == "__main__":
    main()
