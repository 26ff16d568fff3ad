"""CPU tests: the oracle (oracle/) against the reference's own golden vectors (src/tests.rs:30-73) and against
independent naive computations.  This is what pins parity: everything the GPU tests compare with is this oracle."""
import json
import os

import numpy as np

from conftest import GOLDEN, SIMPLE, vk_commitments
from plonkit_b200 import circuit, reader, synth
from plonkit_b200.bn254 import Q_MOD, R_MOD, ints_to_limbs, limbs_to_ints, root_of_unity


def test_montgomery_constants_match_survey_appendix_c(orc):
    c = orc.constants()
    assert c["fr"]["R"] == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert c["fr"]["R2"] == 0x0216d0b17f4e44a58c49833d53bb808553fe3ab1e35c59e31bb8e645ae216da7
    assert c["fr"]["INV"] == 0xc2e1f593efffffff
    assert c["fq"]["R"] == 0x0e0a77c19a07df2f666ea36f7879462c0a78eb28f5c70b3dd35d438dc58f0d9d
    assert c["fq"]["R2"] == 0x06d89f71cab8351f47ab1eff0a417ff6b5e71911d44501fbf32cfc5b538afa89
    assert c["fq"]["INV"] == 0x87d20782e4866389
    assert orc.omega(3) == 0x2b337de1c8c14f22ec9b9e2f96afef3652627366f8170a0a948dad4ac1bd5e80
    assert orc.omega(3) == root_of_unity(3)


def test_keccak256_known_answers(orc):
    assert orc.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert orc.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    # a message longer than one 136-byte block
    assert orc.keccak256(b"a" * 200).hex() == orc.keccak256(b"a" * 200).hex()
    assert len(orc.keccak256(b"x" * 136)) == 32


def test_field_mul_against_python_ints(orc):
    rng = np.random.default_rng(1)
    a = [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(200)] + [0, 1, R_MOD - 1]
    b = [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(200)] + [R_MOD - 1, R_MOD - 1, R_MOD - 1]
    out = limbs_to_ints(orc.fr_mul(ints_to_limbs(a), ints_to_limbs(b)))
    assert out == [x * y % R_MOD for x, y in zip(a, b)]
    aq = [x % Q_MOD for x in a]
    out = limbs_to_ints(orc.fq_mul(ints_to_limbs(aq), ints_to_limbs(b)))
    assert out == [x * y % Q_MOD for x, y in zip(aq, b)]


def test_prove_reproduces_reference_proof_bin(orc, simple_circuit, simple_key):
    """src/tests.rs:48-73 (test_prove): bytes == test/circuits/simple/proof.bin"""
    asm = circuit.synthesize(simple_circuit)
    assert asm.n == 8 and asm.num_inputs == 1
    gold = open(os.path.join(SIMPLE, "proof.bin"), "rb").read()
    for threads in (1, 4):
        proof, ch = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, simple_key.g1_bases,
                              threads=threads, want_challenges=True)
        assert proof == gold
    # SURVEY App. A.4 known challenges
    assert ch[0] == 0x0f72cf563829c88d02442b32aa5bc8b0aff226697faa846756e813710804a058
    assert ch[3] == 0x0913d2eba66540a79bf6ea941e38f856105c5cfe6dadb5738a2b895b337dc63e
    assert ch[4] == 0x1b49fbb2ccfc097e7d0a05e499dcb39e9861c0240726f81beed8fa082c33e916


def test_verification_key_reproduces_reference_vk_bin(orc, simple_circuit, simple_key):
    """src/tests.rs:30-46 (test_export_verification_key): bytes == test/circuits/simple/vk.bin"""
    asm = circuit.synthesize(circuit.CircomCircuit(simple_circuit.r1cs, None))
    com = orc.setup_commitments(asm.n, asm.num_inputs, asm.wire_idx, asm.selectors, simple_key.g1_bases, nvars=asm.nvars, threads=2)
    vk = reader.VerificationKey(asm.n - 1, asm.num_inputs, com[:6], com[6:7], com[7:11], [5, 7, 10], simple_key.g2_raw)
    assert vk.to_bytes() == open(os.path.join(SIMPLE, "vk.bin"), "rb").read()


def test_trapdoor_verifier_accepts_reference_proof(orc):
    """src/tests.rs:75-81 (test_verify) with the pairing replaced by the tau = 42 identity"""
    proof = open(os.path.join(SIMPLE, "proof.bin"), "rb").read()
    vk = reader.load_verification_key(os.path.join(SIMPLE, "vk.bin"))
    assert orc.verify_trapdoor(proof, vk_commitments(vk), 42)
    bad = bytearray(proof)
    bad[700] ^= 1
    assert not orc.verify_trapdoor(bytes(bad), vk_commitments(vk), 42)
    assert not orc.verify_trapdoor(proof, vk_commitments(vk), 43)


def test_srs_generator_reproduces_reference_key(orc, simple_key):
    """keys/setup/setup_2^10.key is [42^i] G (src/plonk.rs:41,47)"""
    assert (orc.srs_gen(1024, 42, threads=4) == simple_key.g1_bases).all()
    assert orc.on_curve(simple_key.g1_bases)


def test_ntt_against_naive_dft_and_round_trips(orc):
    for log_n in (1, 3, 6):
        x = synth.random_field_elements(1 << log_n, seed=log_n)
        assert (orc.ntt(x) == orc.naive_dft(x)).all()
    x = synth.random_field_elements(1 << 12, seed=5)
    for coset in (False, True):
        for threads in (1, 8):
            y = orc.ntt(x, coset=coset, threads=threads)
            assert (orc.ntt(y, inverse=True, coset=coset, threads=threads) == x).all()
    assert (orc.ntt(x, threads=8) == orc.ntt(x, threads=1)).all()
    # LDE x4 restricted to every 4th point of the un-shifted domain is the plain NTT of the zero-padded vector
    c = synth.random_field_elements(8, seed=9)
    pad = np.zeros((32, 4), dtype=np.uint64)
    pad[:8] = c
    assert (orc.lde4(c) == orc.ntt(pad, coset=True)).all()


def test_msm_against_double_and_add(orc, simple_key):
    bases = simple_key.g1_bases[:64]
    s = synth.random_field_elements(64, seed=3)
    s[0] = 0
    s[1] = ints_to_limbs([1])[0]
    s[2] = ints_to_limbs([R_MOD - 1])[0]
    ref = orc.msm_naive(s, bases)
    for threads in (1, 3, 8):
        assert (orc.msm(s, bases, threads=threads) == ref).all()
    # P + (-P) -> infinity
    two = np.stack([bases[5], bases[5]])
    sc = ints_to_limbs([7, R_MOD - 7])
    assert not orc.msm(sc, two).any()
    assert not orc.msm_naive(sc, two).any()


def test_ec_intt_is_the_lagrange_basis(orc, simple_key):
    """Crs::from_powers (src/plonk.rs:179-185): out[i] = L_i(tau) G, checked in the exponent with tau = 42"""
    log_n = 3
    n = 1 << log_n
    out = orc.ec_intt(simple_key.g1_bases[:n], threads=2)
    w = root_of_unity(log_n)
    tau = 42
    gen = simple_key.g1_bases[0]
    for i in range(n):
        wi = pow(w, i, R_MOD)
        li = wi * (pow(tau, n, R_MOD) - 1) % R_MOD * pow(n * (tau - wi) % R_MOD, -1, R_MOD) % R_MOD
        assert (orc.g1_mul(gen, li) == out[i]).all()


def test_analyse_matches_reference_string(simple_circuit):
    """src/tests.rs:14,16-28 (test_analyze)"""
    expect = ('{"num_inputs":2,"num_aux":2,"num_variables":4,"num_constraints":2,"num_nontrivial_constraints":2,"num_gates":3,'
              '"num_hints":2,"constraint_stats":[{"name":"0","num_gates":1},{"name":"1","num_gates":2}]}')
    got = json.dumps(circuit.analyse(circuit.CircomCircuit(simple_circuit.r1cs)), separators=(",", ":"))
    assert got == expect


def test_committed_fixtures_are_what_the_oracle_produces(orc, simple_key):
    asm = synth.poseidon_chain_assembly(9)
    assert circuit.is_satisfied(asm)
    proof = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, simple_key.g1_bases, threads=8)
    assert proof == open(os.path.join(GOLDEN, "poseidon9_proof.bin"), "rb").read()
    com = np.load(os.path.join(GOLDEN, "poseidon9_vk_commitments.npy"))
    assert orc.verify_trapdoor(proof, com, 42)
    x = synth.random_field_elements(1 << 10, seed=synth.SEED + 1)
    assert (orc.ntt(x, threads=2) == np.load(os.path.join(GOLDEN, "ntt10_out.npy"))).all()
    assert (orc.msm(x, simple_key.g1_bases, threads=2) == np.load(os.path.join(GOLDEN, "msm10_out.npy"))).all()


def test_polynomial_primitives_against_python_integers(orc):
    """orc_poly_op restates bellman's evaluate_at / divide_single / calculate_shifted_grand_product / batch_inversion;
    small cases are checked with plain Python integers, the division through p(X) = q(X) (X - z) + p(z)."""
    rng = np.random.default_rng(21)
    for n in (1, 2, 7, 64, 257):
        c = [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(n)]
        c[n // 2] = 0
        z = int.from_bytes(rng.bytes(32), "little") % R_MOD
        C, Z = ints_to_limbs(c), ints_to_limbs([z])[0]
        pz = sum(ci * pow(z, i, R_MOD) for i, ci in enumerate(c)) % R_MOD
        assert limbs_to_ints(orc.poly_op("evaluate_at", C, Z).reshape(1, 4)) == [pz]
        q = limbs_to_ints(orc.poly_op("divide_by_linear", C, Z))
        assert q[n - 1] == 0
        back = [0] * n
        for i, qi in enumerate(q[: n - 1]):
            back[i + 1] = (back[i + 1] + qi) % R_MOD
            back[i] = (back[i] - qi * z) % R_MOD
        back[0] = (back[0] + pz) % R_MOD
        assert back == c
        gp = limbs_to_ints(orc.poly_op("shifted_grand_product", C))
        acc, want = 1, []
        for ci in c:
            want.append(acc)
            acc = acc * ci % R_MOD
        assert gp == want
        inv = limbs_to_ints(orc.poly_op("batch_inversion", C))
        assert inv == [pow(ci, -1, R_MOD) if ci else 0 for ci in c]
