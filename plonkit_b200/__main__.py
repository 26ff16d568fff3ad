"""`python -m plonkit_b200 <subcommand>` — the prove-path subset of the plonkit CLI (src/bin/main.rs:20-256), same
subcommand names, flags, defaults and overwrite guards:

    analyse                  -c/--circuit  -o/--output analyse.json                         (main.rs:313-331)
    setup                    -p/--power  -m/--srs_monomial_form  --overwrite                (main.rs:334-343)
    dump-lagrange            -m  -l/--srs_lagrange_form  -c  --overwrite                    (main.rs:360-381)
                             (under torchrun with N processes: the EC inverse NTT is split over N GPUs)
    prove                    -m  [-l]  -c  -w witness.wtns  -p proof.bin  -j proof.json  -i public.json
                             -t keccak  --overwrite                                          (main.rs:384-424)
    export-verification-key  -m  -c  -v vk.bin  --overwrite                                 (main.rs:484-504)
    verify                   -p proof.bin  -v vk.bin  -t keccak   (host arithmetic, exit code 400) (main.rs:427-438)

Out of scope: generate-verifier (Solidity code generation) and the recursive-* subcommands.  `proof.json` / `public.json` follow contrib/template.sol:864-951 (33 words); their exact text encoding is
not pinned by any in-tree fixture.
"""
import argparse
import json
import os
import sys

from . import plonk, reader
from .circuit import AUX_OFFSET, CircomCircuit


def resolve_circuit_file(filename):  # main.rs:346-357
    if filename:
        return filename
    if os.path.exists("circuit.r1cs") or not os.path.exists("circuit.json"):
        return "circuit.r1cs"
    return "circuit.json"


def _guard(path, what, overwrite):
    if not overwrite and os.path.exists(path):
        raise SystemExit("duplicate %s file: %s" % (what, path))


def serialize_proof(proof):
    """bellman_vk_codegen::serialize_proof: (inputs, 33 proof words) as integers, order of template.sol:864-951"""
    from .bn254 import limbs_to_ints
    import numpy as np

    def pt(p):
        return limbs_to_ints(np.asarray(p, dtype=np.uint64).reshape(2, 4))
    words = []
    for c in proof.wire_commitments:
        words += pt(c)
    words += pt(proof.grand_product_commitment)
    for c in proof.quotient_poly_commitments:
        words += pt(c)
    words += list(proof.wire_values_at_z) + list(proof.wire_values_at_z_omega)
    words += [proof.grand_product_at_z_omega, proof.quotient_polynomial_at_z, proof.linearization_polynomial_at_z]
    words += list(proof.permutation_polynomials_at_z)
    words += pt(proof.opening_at_z_proof) + pt(proof.opening_at_z_omega_proof)
    assert len(words) == 33
    return list(proof.input_values), words


def cmd_analyse(o):
    c = CircomCircuit(reader.load_r1cs(resolve_circuit_file(o.circuit)), None, None, AUX_OFFSET, not o.allow_unpinned_transpilation)
    stats = plonk.analyse(c)
    with open(o.output, "w") as f:
        json.dump(stats, f, indent=2)
    stats.pop("constraint_stats", None)
    print("analyse result: %s" % json.dumps(stats, indent=2), file=sys.stderr)
    print("output to %s" % o.output, file=sys.stderr)


def cmd_setup(o):
    srs = plonk.gen_key_monomial_form(o.power)
    _guard(o.srs_monomial_form, "srs_monomial_form", o.overwrite)
    with open(o.srs_monomial_form, "wb") as f:
        srs.write(f)
    print("srs_monomial_form saved to %s" % o.srs_monomial_form, file=sys.stderr)


def cmd_dump_lagrange(o):
    c = CircomCircuit(reader.load_r1cs(resolve_circuit_file(o.circuit)), None, None, AUX_OFFSET, not o.allow_unpinned_transpilation)
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:
        # one process per GPU (torchrun): the EC inverse NTT is split four-step across the ranks, rank 0 writes the file
        import torch
        import torch.distributed as td
        from . import _lib, dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
        mono = reader.load_key_monomial_form(o.srs_monomial_form)
        n = 1 << max(plonk.domain_log2(c), 0)
        pts = dist.lagrange_key_distributed(mono.g1_bases[:n], _lib.Context(local), rank, world, local)
        td.barrier(device_ids=[local])
        td.destroy_process_group()
        if rank != 0:
            return
        key = reader.Crs(pts, mono.g2_raw, "lagrange")
    else:
        setup = plonk.SetupForProver.prepare_setup_for_prover(c, reader.load_key_monomial_form(o.srs_monomial_form), None)
        key = setup.get_srs_lagrange_form_from_monomial_form()
    _guard(o.srs_lagrange_form, "srs_lagrange_form", o.overwrite)
    with open(o.srs_lagrange_form, "wb") as f:
        key.write(f)
    print("srs_lagrange_form saved to %s" % o.srs_lagrange_form, file=sys.stderr)


def cmd_prove(o):
    c = CircomCircuit(reader.load_r1cs(resolve_circuit_file(o.circuit)), reader.load_witness_limbs(o.witness), None, AUX_OFFSET,
                      not o.allow_unpinned_transpilation)
    # the gate tables do not depend on the witness: transpile without it, then assign the witness through the circuit's plan
    shape = CircomCircuit(c.r1cs, None, None, AUX_OFFSET, c.strict)
    setup = plonk.SetupForProver.prepare_setup_for_prover(shape, reader.load_key_monomial_form(o.srs_monomial_form),
                                                          reader.maybe_load_key_lagrange_form(o.srs_lagrange_form))
    print("Proving...", file=sys.stderr)
    proof = setup.prove(c, o.transcript)
    _guard(o.proof, "proof", o.overwrite)
    with open(o.proof, "wb") as f:
        proof.write(f)
    print("Proof saved to %s" % o.proof, file=sys.stderr)
    inputs, words = serialize_proof(proof)
    _guard(o.proofjson, "proof json", o.overwrite)
    _guard(o.publicjson, "input json", o.overwrite)
    with open(o.proofjson, "w") as f:
        json.dump([hex(w) for w in words], f, indent=2)
    with open(o.publicjson, "w") as f:
        json.dump([hex(w) for w in inputs], f, indent=2)


def cmd_export_vk(o):
    c = CircomCircuit(reader.load_r1cs(resolve_circuit_file(o.circuit)), None, None, AUX_OFFSET, not o.allow_unpinned_transpilation)
    setup = plonk.SetupForProver.prepare_setup_for_prover(c, reader.load_key_monomial_form(o.srs_monomial_form), None)
    vk = setup.make_verification_key()
    _guard(o.vk, "vk", o.overwrite)
    with open(o.vk, "wb") as f:
        vk.write(f)
    print("Verification key saved to %s" % o.vk, file=sys.stderr)


def cmd_verify(o):  # main.rs:427-438
    vk = reader.load_verification_key(o.vk)
    proof = reader.load_proof(o.proof)
    if plonk.verify(vk, proof, o.transcript):
        print("Proof is valid.", file=sys.stderr)
    else:
        print("Proof is invalid!", file=sys.stderr)
        raise SystemExit(400)


def build_parser():
    ap = argparse.ArgumentParser(prog="plonkit", description="prove-path subset of the plonkit CLI on the CUDA library")
    sub = ap.add_subparsers(dest="command", required=True)
    p = sub.add_parser("analyse")
    p.add_argument("-c", "--circuit")
    p.add_argument("-o", "--output", default="analyse.json")
    p.set_defaults(fn=cmd_analyse)
    p = sub.add_parser("setup")
    p.add_argument("-p", "--power", type=int, required=True)
    p.add_argument("-m", "--srs_monomial_form", required=True)
    p.add_argument("--overwrite", action="store_true")
    p.set_defaults(fn=cmd_setup)
    p = sub.add_parser("dump-lagrange")
    p.add_argument("-m", "--srs_monomial_form", required=True)
    p.add_argument("-l", "--srs_lagrange_form", required=True)
    p.add_argument("-c", "--circuit")
    p.add_argument("--overwrite", action="store_true")
    p.set_defaults(fn=cmd_dump_lagrange)
    p = sub.add_parser("prove")
    p.add_argument("-m", "--srs_monomial_form", required=True)
    p.add_argument("-l", "--srs_lagrange_form")
    p.add_argument("-c", "--circuit")
    p.add_argument("-w", "--witness", default="witness.wtns")
    p.add_argument("-p", "--proof", default="proof.bin")
    p.add_argument("-j", "--proofjson", default="proof.json")
    p.add_argument("-i", "--publicjson", default="public.json")
    p.add_argument("-t", "--transcript", default="keccak")
    p.add_argument("--overwrite", action="store_true")
    p.set_defaults(fn=cmd_prove)
    p = sub.add_parser("export-verification-key")
    p.add_argument("-m", "--srs_monomial_form", required=True)
    p.add_argument("-c", "--circuit")
    p.add_argument("-v", "--vk", default="vk.bin")
    p.add_argument("--overwrite", action="store_true")
    p.set_defaults(fn=cmd_export_vk)
    p = sub.add_parser("verify")
    p.add_argument("-p", "--proof", default="proof.bin")
    p.add_argument("-v", "--vk", default="vk.bin")
    p.add_argument("-t", "--transcript", default="keccak")
    p.set_defaults(fn=cmd_verify)
    for name in ("analyse", "dump-lagrange", "prove", "export-verification-key"):
        # not a flag of the reference CLI: R1CS shapes whose gate layout no reference fixture pins (linear combinations
        # with more than two terms) are refused unless this is given; the proofs then verify but byte parity with the
        # reference is unpinned (circuit._transpile)
        sub.choices[name].add_argument("--allow-unpinned-transpilation", action="store_true")
    for name in ("generate-verifier", "generate-recursive-verifier", "export-recursive-verification-key",
                 "recursive-prove", "recursive-verify", "check-aggregation"):
        q = sub.add_parser(name)
        q.add_argument("rest", nargs=argparse.REMAINDER)
        q.set_defaults(fn=lambda o, n=name: (_ for _ in ()).throw(SystemExit(
            "%s is outside this repository's scope (prove path only; see DESIGN.md section 0)" % n)))
    return ap


def main(argv=None):
    o = build_parser().parse_args(argv)
    o.fn(o)
    return 0


if __name__ == "__main__":
    sys.exit(main())
