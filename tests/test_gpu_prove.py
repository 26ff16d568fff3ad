"""GPU parity tests of the prover path through the reference-shaped API (plonkit_b200.plonk) and the C ABI:
golden proof.bin / vk.bin of the reference, oracle equality on synthetic circuits, error behaviour, and
size-independent checks (trapdoor verification) at BASELINE.json's full size."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, SIMPLE, vk_commitments
from plonkit_b200 import _lib, circuit, plonk, synth
from plonkit_b200.bn254 import R_MOD, ints_to_limbs

pytestmark = pytest.mark.gpu


def test_prove_reproduces_reference_proof_bin(ctx, simple_circuit, simple_key):
    """src/tests.rs:48-73 (test_prove) through the CUDA path"""
    setup = plonk.SetupForProver.prepare_setup_for_prover(simple_circuit, simple_key, None, ctx=ctx)
    setup.validate_witness(simple_circuit)
    proof = setup.prove(simple_circuit, "keccak")
    assert proof.to_bytes() == open(os.path.join(SIMPLE, "proof.bin"), "rb").read()
    assert proof.challenges[0] == 0x0f72cf563829c88d02442b32aa5bc8b0aff226697faa846756e813710804a058  # beta
    # proving twice gives the same bytes (no blinding) and the resident-witness path agrees
    setup.upload_witness(simple_circuit)
    assert setup.prove(None).to_bytes() == proof.to_bytes()


def test_verification_key_reproduces_reference_vk_bin(ctx, simple_circuit, simple_key):
    """src/tests.rs:30-46 (test_export_verification_key) through the CUDA path"""
    c = circuit.CircomCircuit(simple_circuit.r1cs, None)
    setup = plonk.SetupForProver.prepare_setup_for_prover(c, simple_key, None, ctx=ctx)
    assert setup.make_verification_key().to_bytes() == open(os.path.join(SIMPLE, "vk.bin"), "rb").read()


def test_lagrange_srs_matches_oracle(ctx, orc, simple_circuit, simple_key):
    """src/tests.rs:66 calls get_srs_lagrange_form_from_monomial_form and discards it; here it is checked"""
    setup = plonk.SetupForProver.prepare_setup_for_prover(simple_circuit, simple_key, None, ctx=ctx)
    crs = setup.get_srs_lagrange_form_from_monomial_form()
    assert crs.form == "lagrange" and crs.size == 8
    assert (crs.g1_bases == orc.ec_intt(simple_key.g1_bases[:8])).all()


def test_prove_with_lagrange_key_gives_the_same_bytes(ctx, orc, simple_circuit, simple_key):
    """`plonkit prove -l <lagrange key>` (src/plonk.rs:138-146, CI: integration-test.yml:127-133): wire commitments from
    VALUES with the Lagrange-form key made by dump-lagrange == the monomial path == the reference's proof.bin; a key of the
    wrong domain size and a non-keccak transcript are refused as in the reference."""
    from plonkit_b200.reader import Crs
    base = plonk.SetupForProver.prepare_setup_for_prover(simple_circuit, simple_key, None, ctx=ctx)
    lag = base.get_srs_lagrange_form_from_monomial_form()
    setup = plonk.SetupForProver.prepare_setup_for_prover(simple_circuit, simple_key, lag, ctx=ctx)
    assert setup.prove(simple_circuit, "keccak").to_bytes() == open(os.path.join(SIMPLE, "proof.bin"), "rb").read()
    with pytest.raises(NotImplementedError):
        setup.prove(simple_circuit, "rescue")
    with pytest.raises(_lib.SynthesisError):
        plonk.SetupForProver.prepare_setup_for_prover(simple_circuit, simple_key, Crs(simple_key.g1_bases[:16], b"", "lagrange"), ctx=ctx)
    # a larger, witness-heavy circuit: values path == coefficient path == oracle; a setup WITHOUT the Lagrange key on the same
    # context afterwards still takes the monomial path
    asm = synth.random_gate_assembly(12, seed=77, num_inputs=2)
    srs = orc.srs_gen(asm.n, 42, threads=8)
    key = Crs(srs)
    mono = plonk.SetupForProver.prepare_setup_for_prover(asm, key, None, ctx=ctx)
    lag2 = mono.get_srs_lagrange_form_from_monomial_form()
    want = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=8)
    both = plonk.SetupForProver.prepare_setup_for_prover(asm, key, lag2, ctx=ctx)
    assert both.prove(asm).to_bytes() == want
    assert mono.prove(asm).to_bytes() == want
    for s_ in (base, setup, mono, both):
        s_.close()


def test_cli_prove_with_lagrange_key(tmp_path):
    """dump-lagrange then prove -l (test/test_poseidon_plonk.sh flow on the `simple` circuit): same proof.bin"""
    from plonkit_b200 import __main__ as cli
    key, circ = os.path.join(SIMPLE, "setup_2^10.key"), os.path.join(SIMPLE, "circuit.r1cs.json")
    lag = tmp_path / "lagrange.key"
    cli.main(["dump-lagrange", "-m", key, "-l", str(lag), "-c", circ])
    cli.main(["prove", "-m", key, "-l", str(lag), "-c", circ, "-w", os.path.join(SIMPLE, "witness.json"), "-p", str(tmp_path / "proof.bin"),
              "-j", str(tmp_path / "proof.json"), "-i", str(tmp_path / "public.json")])
    assert (tmp_path / "proof.bin").read_bytes() == open(os.path.join(SIMPLE, "proof.bin"), "rb").read()


def test_poseidon_shaped_golden_fixture(ctx, simple_key):
    asm = synth.poseidon_chain_assembly(9)
    setup = plonk.SetupForProver.prepare_setup_for_prover(asm, simple_key, None, ctx=ctx)
    assert setup.prove(asm).to_bytes() == open(os.path.join(GOLDEN, "poseidon9_proof.bin"), "rb").read()
    vk = setup.make_verification_key()
    assert (vk_commitments(vk) == np.load(os.path.join(GOLDEN, "poseidon9_vk_commitments.npy"))).all()


@pytest.mark.parametrize("kind,log_n", [("poseidon", 4), ("poseidon", 10), ("poseidon", 12), ("random", 6), ("random", 11),
                                        ("poseidon", 14), ("poseidon", 16)])
def test_proofs_equal_oracle_on_synthetic_circuits(ctx, orc, kind, log_n):
    asm = synth.poseidon_chain_assembly(log_n) if kind == "poseidon" else synth.random_gate_assembly(log_n, seed=log_n)
    srs = orc.srs_gen(asm.n, 42, threads=8)
    from plonkit_b200.reader import CRS_42_G2, Crs
    key = Crs(srs, CRS_42_G2, "monomial")
    setup = plonk.SetupForProver.prepare_setup_for_prover(asm, key, None, ctx=ctx)
    proof = setup.prove(asm)
    ref = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=8)
    assert proof.to_bytes() == ref
    com = orc.setup_commitments(asm.n, asm.num_inputs, asm.wire_idx, asm.selectors, srs, nvars=asm.nvars, threads=8)
    vk = setup.make_verification_key()
    assert (vk_commitments(vk) == com).all()
    assert orc.verify_trapdoor(proof.to_bytes(), com, 42)
    if log_n <= 12:
        assert plonk.verify(vk, proof, "keccak")   # the pairing-based verifier of the API (src/plonk.rs:189-210), no trapdoor


def test_prover_pool_keeps_order_and_bytes(orc, simple_key):
    """ProverPool: three provers in flight on one GPU give the same bytes as the oracle, in input order; a bad witness
    surfaces as the reference's error on the caller's thread."""
    asms = [synth.poseidon_chain_assembly(9, inputs=(3 + k, 4, 5)) for k in range(5)]
    pool = plonk.ProverPool(asms[0], simple_key, inflight=3)
    try:
        proofs = pool.prove_all([a.var_values for a in asms])
        for a, p in zip(asms, proofs):
            assert p.to_bytes() == orc.prove(a.n, a.num_inputs, a.wire_idx, a.var_values, a.selectors, simple_key.g1_bases[:a.n], threads=4)
        bad = asms[1].var_values.copy()
        bad[5] = ints_to_limbs([123456789])[0]
        with pytest.raises(_lib.SynthesisError) as e:
            pool.prove_all([asms[0].var_values, bad])
        assert e.value.code == 4
    finally:
        pool.close()


@pytest.mark.parametrize("num_inputs", [0, 3, 8, 9, 20])
def test_public_input_polynomial_paths(ctx, orc, num_inputs):
    """<= 8 inputs: PI(X) is read off the resident L_0 table; more: iNTT + LDE.  Both must give the oracle's bytes."""
    asm = synth.random_gate_assembly(8, seed=40 + num_inputs, num_inputs=num_inputs)
    assert asm.num_inputs == num_inputs and circuit.is_satisfied(asm)
    srs = orc.srs_gen(asm.n, 42, threads=4)
    from plonkit_b200.reader import Crs
    setup = plonk.SetupForProver.prepare_setup_for_prover(asm, Crs(srs, b""), None, ctx=ctx)
    proof = setup.prove(asm)
    assert proof.to_bytes() == orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=4)
    assert len(proof.input_values) == num_inputs


def test_error_behaviour_mirrors_reference(ctx, orc, simple_circuit, simple_key):
    setup = plonk.SetupForProver.prepare_setup_for_prover(simple_circuit, simple_key, None, ctx=ctx)
    # wrong witness: the reference panics in is_satisfied_using_one_shot_check ("must satisfy", src/plonk.rs:137)
    bad = circuit.CircomCircuit(simple_circuit.r1cs, [1, 36, 3, 9])
    with pytest.raises(_lib.SynthesisError) as e:
        setup.prove(bad)
    assert e.value.code == 4
    with pytest.raises(_lib.SynthesisError):
        setup.validate_witness(bad)
    # witness of the wrong length -> AssignmentMissing
    with pytest.raises(_lib.SynthesisError) as e:
        setup.prove(np.zeros((3, 4), dtype=np.uint64))
    assert e.value.code == 1
    # transcript names (src/plonk.rs:147-173)
    with pytest.raises(NotImplementedError):
        setup.prove(simple_circuit, "blake")
    # SRS smaller than the circuit
    from plonkit_b200.reader import Crs
    asm = synth.poseidon_chain_assembly(11)
    with pytest.raises(_lib.SynthesisError) as e:
        plonk.SetupForProver.prepare_setup_for_prover(asm, Crs(simple_key.g1_bases, b""), None, ctx=ctx)
    assert e.value.code == 2


@pytest.fixture(scope="module")
def srs20(ctx, orc):
    srs = ctx.srs_gen(1 << 20, 42)
    assert (srs[:1024] == orc.srs_gen(1024, 42, threads=8)).all()
    return srs


def test_full_size_2pow20_poseidon_proof_bytes_equal_oracle(ctx, orc, srs20):
    """BASELINE.json configs[1], the exact configuration bench.py times (2^20 - 1 gates, default MSM window plan c = 20 /
    13 windows, resident setup): proof bytes == the oracle's proof bytes (src/tests.rs:48-73 semantics, byte equality),
    verification-key commitments == oracle, the trapdoor verifier accepts, the resident-witness path gives the same
    bytes, and a tampered witness is rejected."""
    log_n = 20
    asm = synth.poseidon_chain_assembly(log_n)
    from plonkit_b200.reader import Crs
    setup = plonk.SetupForProver.prepare_setup_for_prover(asm, Crs(srs20), None, ctx=ctx)
    proof = setup.prove(asm)
    ref = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs20, threads=16)
    assert proof.to_bytes() == ref
    com = vk_commitments(setup.make_verification_key())
    assert (com == orc.setup_commitments(asm.n, asm.num_inputs, asm.wire_idx, asm.selectors, srs20, nvars=asm.nvars, threads=16)).all()
    assert orc.verify_trapdoor(proof.to_bytes(), com, 42)
    assert setup.prove(None).to_bytes() == ref
    assert proof.input_values == [3]
    vals = asm.var_values.copy()
    vals[1000] = ints_to_limbs([(int(vals[1000][0]) + 1) % R_MOD])[0]
    with pytest.raises(_lib.SynthesisError):
        setup.prove(vals)
    setup.close()


def test_full_size_2pow20_random_gate_circuit_bytes_equal_oracle(ctx, orc, srs20):
    """BASELINE.json configs[2] shape (SURVEY 8d cfg 3: half multiplication, half addition gates, operands re-used with
    p = 0.5 so the copy permutation is non-trivial) at 2^20 gates: proof bytes == oracle."""
    asm = synth.random_gate_assembly(20, seed=2020)
    assert asm.n == 1 << 20
    from plonkit_b200.reader import Crs
    setup = plonk.SetupForProver.prepare_setup_for_prover(asm, Crs(srs20), None, ctx=ctx)
    proof = setup.prove(asm)
    assert proof.to_bytes() == orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs20, threads=16)
    setup.close()


def test_cli_proves_a_poseidon_shaped_r1cs_with_wide_linear_combinations(tmp_path, orc, simple_key):
    """`plonkit prove` on a circomlib-Poseidon(2)-shaped .r1cs / .wtns pair (244 constraints, combinations up to 61 terms;
    test/circuits/poseidon's own R1CS cannot be produced here: no circom).  The transpilation of long combinations is BYTE
    PARITY UNPINNED (own layout), so the check is: the proof verifies against the exported verification key with the
    trapdoor verifier, and equals the oracle's proof of the same gate tables."""
    from plonkit_b200 import __main__ as cli
    from plonkit_b200 import reader
    r1cs, wit = synth.poseidon_r1cs()
    synth.write_r1cs_bin(r1cs, str(tmp_path / "circuit.r1cs"))
    synth.write_wtns(wit, str(tmp_path / "witness.wtns"))
    key = os.path.join(GOLDEN, "setup_2^12.key")
    if not os.path.exists(key):
        key = str(tmp_path / "setup.key")
        cli.main(["setup", "-p", "12", "-m", key])
    with pytest.raises(circuit.UnpinnedTranspilation):   # strict by default
        cli.main(["prove", "-m", key, "-c", str(tmp_path / "circuit.r1cs"), "-w", str(tmp_path / "witness.wtns"), "-p", str(tmp_path / "p0.bin")])
    common = ["-m", key, "-c", str(tmp_path / "circuit.r1cs"), "--allow-unpinned-transpilation"]
    cli.main(["export-verification-key"] + common + ["-v", str(tmp_path / "vk.bin")])
    cli.main(["prove"] + common + ["-w", str(tmp_path / "witness.wtns"), "-p", str(tmp_path / "proof.bin"),
                                  "-j", str(tmp_path / "proof.json"), "-i", str(tmp_path / "public.json")])
    proof = (tmp_path / "proof.bin").read_bytes()
    vk = reader.load_verification_key(str(tmp_path / "vk.bin"))
    assert orc.verify_trapdoor(proof, vk_commitments(vk), 42)
    asm = circuit.synthesize(circuit.CircomCircuit(r1cs, wit, strict=False))
    srs = reader.load_key_monomial_form(key).g1_bases
    assert proof == orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs[:asm.n], threads=8)
    assert reader.load_proof(str(tmp_path / "proof.bin")).input_values == [wit[1]]


def test_cli_setup_writes_the_reference_key_file(tmp_path):
    """`plonkit setup -p 10` (src/bin/main.rs:334-343 -> gen_key_monomial_form -> Crs::crs_42) == keys/setup/setup_2^10.key,
    G2 part included; and a verification key exported from that key carries the trailing 256 B of G2."""
    from plonkit_b200 import __main__ as cli
    from plonkit_b200 import reader
    out = tmp_path / "setup.key"
    cli.main(["setup", "-p", "10", "-m", str(out)])
    assert out.read_bytes() == open(os.path.join(SIMPLE, "setup_2^10.key"), "rb").read()
    vk = tmp_path / "vk.bin"
    cli.main(["export-verification-key", "-m", str(out), "-c", os.path.join(SIMPLE, "circuit.r1cs.json"), "-v", str(vk)])
    assert vk.read_bytes() == open(os.path.join(SIMPLE, "vk.bin"), "rb").read()
    assert reader.load_verification_key(str(vk)).g2_raw == reader.CRS_42_G2


def test_cli_prove_and_export_vk_write_the_golden_files(tmp_path):
    """`plonkit export-verification-key` / `plonkit prove` (src/bin/main.rs:384-424,484-504) through the CLI mirror"""
    from plonkit_b200 import __main__ as cli
    key = os.path.join(SIMPLE, "setup_2^10.key")
    circ = os.path.join(SIMPLE, "circuit.r1cs.json")
    vk, proof = tmp_path / "vk.bin", tmp_path / "proof.bin"
    cli.main(["export-verification-key", "-m", key, "-c", circ, "-v", str(vk)])
    assert vk.read_bytes() == open(os.path.join(SIMPLE, "vk.bin"), "rb").read()
    cli.main(["prove", "-m", key, "-c", circ, "-w", os.path.join(SIMPLE, "witness.json"), "-p", str(proof),
              "-j", str(tmp_path / "proof.json"), "-i", str(tmp_path / "public.json")])
    assert proof.read_bytes() == open(os.path.join(SIMPLE, "proof.bin"), "rb").read()
    import json
    assert json.load(open(tmp_path / "public.json")) == ["0x23"] and len(json.load(open(tmp_path / "proof.json"))) == 33
    with pytest.raises(SystemExit, match="duplicate proof file"):
        cli.main(["prove", "-m", key, "-c", circ, "-w", os.path.join(SIMPLE, "witness.json"), "-p", str(proof)])


@pytest.mark.parametrize("log_n", [6, 10, 13])
def test_two_gate_type_prover_matches_its_oracle_and_verifies(ctx, orc, log_n):
    """The recursive prover's proving call (src/recursive/mod.rs:120-127: main gate + Rescue x^5 custom gate behind gate
    selectors) on a synthetic circuit of that shape.  BYTE PARITY UNPINNED (no reference fixture, the prover's crates are
    not in the tree): the CUDA proof must equal the oracle's restatement of the same protocol byte for byte, the restated
    verifier (known trapdoor) must accept it against the exported 13-commitment key, and reject a tampered proof; a witness
    that breaks a custom-gate row is refused like an unsatisfied main-gate row."""
    from plonkit_b200 import recursive
    from plonkit_b200.reader import CRS_42_G2, Crs
    asm, gate_type = synth.rescue_chain_assembly(log_n)
    assert gate_type.sum() > 0 and asm.n == 1 << log_n
    srs = orc.srs_gen(asm.n, 42, threads=8)
    setup = recursive.RecursiveSetupForProver(asm, gate_type, Crs(srs, CRS_42_G2), ctx=ctx)
    proof_obj = setup.create_proof(asm.var_values)
    proof = proof_obj.to_bytes()
    assert proof == orc.prove2(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, gate_type, srs, threads=8)
    vk = setup.export_vk()
    assert (vk.commitments == orc.setup_commitments2(asm.n, asm.num_inputs, asm.wire_idx, asm.selectors, gate_type, srs,
                                                      nvars=asm.nvars, threads=8)).all()
    assert orc.verify_trapdoor2(proof, vk.commitments, 42)
    if log_n <= 10:
        assert recursive.verify(vk, proof_obj)          # the pairing check, no trapdoor
    bad = bytearray(proof)
    bad[-150] ^= 1          # inside s_resc(z)
    assert not orc.verify_trapdoor2(bytes(bad), vk.commitments, 42)
    vals = asm.var_values.copy()
    row = int(np.nonzero(gate_type)[0][0])
    vals[asm.wire_idx[1, row]] = ints_to_limbs([5])[0]      # x^2 wire of a custom-gate row
    with pytest.raises(_lib.SynthesisError) as e:
        setup.create_proof(vals)
    assert e.value.code == 4
    setup.close()
    with pytest.raises(NotImplementedError):
        recursive.prove(None, [], None)
