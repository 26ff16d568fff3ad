#!/usr/bin/env python3
"""Host-side intake of a large circuit, compiled (libplonkit_host.so) against the Python statements: `.r1cs` parse, R1CS -> width-4
transpilation, per-proof witness assignment.  No device needed.  usage: host_intake_time.py [hashes=450]  (450 chained
Poseidon(2)-shaped permutations = 1.04 M gates, a 2^20 domain)"""
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonkit_b200 import circuit, reader, synth  # noqa: E402


def timed(fn):
    t = time.perf_counter()
    out = fn()
    return out, time.perf_counter() - t


hashes = int(sys.argv[1]) if len(sys.argv) > 1 else 450
r1cs, wit = synth.poseidon_r1cs(hashes)
with tempfile.TemporaryDirectory() as d:
    rp, wp = os.path.join(d, "c.r1cs"), os.path.join(d, "w.wtns")
    synth.write_r1cs_bin(r1cs, rp)
    synth.write_wtns(wit, wp)
    print("circuit: %d constraints, .r1cs %d MB, host threads %d" % (len(r1cs.constraints), os.path.getsize(rp) >> 20, os.cpu_count()))
    res = {}
    for native in (True, False):
        circuit.NATIVE[0] = native
        r, t_read = timed(lambda: reader.load_r1cs(rp))
        asm, t_tr = timed(lambda: circuit.synthesize(circuit.CircomCircuit(r, None, None, circuit.AUX_OFFSET, False)))
        w = reader.load_witness_limbs(wp)
        asm.plan.assign(w, native=native)                       # first call converts the coefficients once
        _, t_as = timed(lambda: asm.plan.assign(w, native=native))
        _, t_as1 = timed(lambda: asm.plan.assign(w, native=native, threads=1)) if native else (None, t_as)
        res[native] = asm
        print("%-8s read %.2f s   transpile %.2f s   assign %.3f s%s   (%d gates, domain 2^%d, %d introduced variables, %d terms)" % (
            "compiled" if native else "python", t_read, t_tr, t_as, " (1 thread: %.3f s)" % t_as1 if native else "",
            asm.num_gates, asm.n.bit_length() - 1, asm.plan.num_new, len(asm.plan.term_var)))
    circuit.NATIVE[0] = True
    a, b = res[True], res[False]
    same = (a.wire_idx == b.wire_idx).all() and (a.selectors == b.selectors).all() and (a.plan.term_coef == b.plan.term_coef).all()
    print("tables and witness program identical:", bool(same))
