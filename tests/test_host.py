"""CPU tests of the product's host side: file formats, circuit synthesis, the C-ABI library's exports, loud failure
without a GPU, and the host build of the device arithmetic headers (same carry-chain sequences as the kernels)."""
import ctypes
import io
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT, SIMPLE
from plonkit_b200 import _lib, circuit, plonk, reader, synth
from plonkit_b200.bn254 import Q_MOD, R_MOD, ints_to_limbs, limbs_to_ints, root_of_unity


# ---------------------------------------------------------------- C ABI surface
def test_library_loads_and_exports_every_declared_symbol():
    lib = _lib.load()  # needs no GPU
    hdr = open(os.path.join(ROOT, "include", "plonkit_b200.h")).read()
    declared = set(re.findall(r"\b(pk_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no prototypes found"
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None


def test_host_library_loads_and_exports_every_declared_symbol():
    """include/plonkit_host.h (plain C++ host helpers, no CUDA) against plonkit_b200/libplonkit_host.so, and its field product
    against Python integers"""
    lib = circuit.host_library()
    assert lib is not None, "libplonkit_host.so is not built (python __graft_entry__.py)"
    hdr = open(os.path.join(ROOT, "include", "plonkit_host.h")).read()
    declared = set(re.findall(r"\b(ph_[a-z0-9_]+)\s*\(", hdr))
    assert declared == {"ph_assign_witness", "ph_fr_to_mont", "ph_fr_mul", "ph_last_error", "ph_r1cs_parse_bin", "ph_r1cs_from_csr",
                        "ph_r1cs_free", "ph_r1cs_header", "ph_r1cs_export", "ph_transpile", "ph_gates_free", "ph_gates_header",
                        "ph_gates_export"}
    for name in declared:
        assert getattr(lib, name) is not None
    rng = np.random.default_rng(3)
    a = [0, 1, R_MOD - 1, R_MOD - 2, 1 << 253] + [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(500)]
    b = [R_MOD - 1] * 3 + [5, 1 << 253] + [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(500)]
    A, B = ints_to_limbs(a), ints_to_limbs(b)
    out = np.zeros_like(A)
    lib.ph_fr_mul(_p(A), _p(B), _p(out), ctypes.c_uint64(len(a)))
    assert limbs_to_ints(out) == [x * y % R_MOD for x, y in zip(a, b)]
    lib.ph_fr_to_mont(_p(A), _p(out), ctypes.c_uint64(len(a)))
    assert limbs_to_ints(out) == [x * (1 << 256) % R_MOD for x in a]


def test_constants_in_library_match_appendix_c():
    c = _lib.constants()
    vals = limbs_to_ints(c.reshape(6, 4))
    assert vals[0] == (1 << 256) % R_MOD and vals[1] == pow(1 << 256, 2, R_MOD)
    assert vals[3] == (1 << 256) % Q_MOD and vals[4] == pow(1 << 256, 2, Q_MOD)
    assert int(c[8]) == 0xefffffff and int(c[20]) == 0xe4866389


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="a GPU is present")
def test_no_gpu_means_loud_failure_not_a_fallback():
    with pytest.raises(_lib.SynthesisError) as e:
        _lib.Context(0)
    assert e.value.code == 5
    with pytest.raises(_lib.SynthesisError):
        plonk.gen_key_monomial_form(10)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "plonkit_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in src.lower() or f == "__init__.py", f


# ---------------------------------------------------------------- formats
def test_key_proof_vk_round_trip_reference_files(simple_key):
    raw = open(os.path.join(SIMPLE, "setup_2^10.key"), "rb").read()
    b = io.BytesIO()
    simple_key.write(b)
    assert b.getvalue() == raw
    assert simple_key.size == 1024 and len(simple_key.g2_raw) == 256
    praw = open(os.path.join(SIMPLE, "proof.bin"), "rb").read()
    p = reader.load_proof(os.path.join(SIMPLE, "proof.bin"))
    assert p.to_bytes() == praw and p.n == 7 and p.input_values == [35]
    assert not p.wire_commitments[3].any()  # d wire is identically zero -> point at infinity (0x40 encoding)
    vraw = open(os.path.join(SIMPLE, "vk.bin"), "rb").read()
    vk = reader.load_verification_key(os.path.join(SIMPLE, "vk.bin"))
    assert vk.to_bytes() == vraw and vk.non_residues == [5, 7, 10]
    # Crs::crs_42's G2 part is a constant ([1]G2, [42]G2): the embedded copy is the reference key's, and so is the vk's
    assert reader.CRS_42_G2 == raw[-256:] == vraw[-256:]
    # truncated files and foreign G2 counts are rejected, not silently accepted
    with pytest.raises(ValueError, match="truncated"):
        reader.Crs.read(io.BytesIO(raw[:-1]))
    with pytest.raises(ValueError, match="truncated"):
        reader.VerificationKey.read(io.BytesIO(vraw[:-256]))
    with pytest.raises(ValueError, match="two G2"):
        reader.Crs.read(io.BytesIO(raw[:-264] + struct.pack(">Q", 0)))
    # every Crs object has its own never-reused token (what SetupForProver uses to decide whether the resident SRS is current)
    assert reader.Crs(simple_key.g1_bases).token != reader.Crs(simple_key.g1_bases).token


def _wtns(vals, version=2, prime=None):
    prime = prime or bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")
    body = b"".join(v.to_bytes(32, "little") for v in vals)
    return (b"wtns" + struct.pack("<II", version, 2) + struct.pack("<IQ", 1, 40) + struct.pack("<I", 32) + prime +
            struct.pack("<I", len(vals)) + struct.pack("<IQ", 2, len(body)) + body)


def test_wtns_parser_and_its_error_paths():
    """src/reader.rs:124-175"""
    assert reader.load_witness_from_array(_wtns([1, 35, 3, 9])) == [1, 35, 3, 9]
    assert reader.load_witness_from_array(_wtns([])) == []
    for bad, msg in ((b"xtns" + _wtns([1])[4:], "invalid file header"), (_wtns([1], version=3), "unsupported file version"),
                     (_wtns([1], prime=b"\x02" + b"\x00" * 31), "invalid curve prime")):
        with pytest.raises(ValueError, match=msg):
            reader.load_witness_from_array(bad)
    with pytest.raises(ValueError):
        reader.load_witness_from_array(_wtns([R_MOD]))
    for cut in (3, 11, 20, 40, 70):
        with pytest.raises(ValueError):
            reader.load_witness_from_array(_wtns([1, 2, 3])[:cut])


def test_proof_list_and_witness_file_loaders(tmp_path):
    """src/reader.rs:27-47 (load_proofs_from_list), :101-116 (the per-format witness loaders)"""
    golden = os.path.join(SIMPLE, "proof.bin")
    (tmp_path / "list.txt").write_text(golden + "\n" + golden + "\n")
    proofs = reader.load_proofs_from_list(str(tmp_path / "list.txt"))
    assert len(proofs) == 2 and proofs[0].to_bytes() == proofs[1].to_bytes() == open(golden, "rb").read()
    (tmp_path / "empty.txt").write_text("")
    with pytest.raises(ValueError, match="no proof file found"):
        reader.load_proofs_from_list(str(tmp_path / "empty.txt"))
    other = reader.load_proof(golden)
    other.num_inputs, other.input_values = 2, other.input_values * 2
    (tmp_path / "other.bin").write_bytes(other.to_bytes())
    (tmp_path / "mixed.txt").write_text(golden + "\n" + str(tmp_path / "other.bin") + "\n")
    with pytest.raises(ValueError, match="num_inputs mismatch"):
        reader.load_proofs_from_list(str(tmp_path / "mixed.txt"))
    wj = os.path.join(SIMPLE, "witness.json")
    assert reader.load_witness_from_json_file(wj) == reader.load_witness_from_file(wj)
    (tmp_path / "w.wtns").write_bytes(_wtns([1, 35, 3, 9]))
    assert reader.load_witness_from_bin_file(str(tmp_path / "w.wtns")) == [1, 35, 3, 9]


def test_witness_as_limbs_equals_witness_as_integers(tmp_path):
    """reader.load_witness_limbs: the same values as load_witness_from_file, as the (len, 4) limb array the prover consumes
    (.wtns elements taken as they lie in the file), same field check; a CircomCircuit holding it behaves the same."""
    vals = [1, 35, 3, 9, R_MOD - 1, 1 << 200, (1 << 64) - 1, 1 << 64]
    path = tmp_path / "w.wtns"
    path.write_bytes(_wtns(vals))
    limbs = reader.load_witness_limbs(str(path))
    assert limbs.shape == (len(vals), 4) and limbs_to_ints(limbs) == vals == reader.load_witness_from_file(str(path))
    assert limbs_to_ints(reader.load_witness_limbs(os.path.join(SIMPLE, "witness.json"))) == \
        reader.load_witness_from_file(os.path.join(SIMPLE, "witness.json"))
    for bad in (R_MOD, R_MOD + 1, (1 << 256) - 1):
        path.write_bytes(_wtns([1, bad % (1 << 256)]))
        with pytest.raises(ValueError, match="not in the field"):
            reader.load_witness_limbs(str(path))
    path.write_bytes(_wtns([1, 2, 3])[:-5])
    with pytest.raises(ValueError, match="truncated"):
        reader.load_witness_limbs(str(path))
    r1cs = reader.load_r1cs(os.path.join(SIMPLE, "circuit.r1cs.json"))
    wj = reader.load_witness_from_file(os.path.join(SIMPLE, "witness.json"))
    a = circuit.CircomCircuit(r1cs, wj)
    b = circuit.CircomCircuit(r1cs, ints_to_limbs(wj))
    assert a.get_public_inputs() == b.get_public_inputs() and a.get_public_inputs_json() == b.get_public_inputs_json()
    assert (circuit.synthesize(a).var_values == circuit.synthesize(b).var_values).all()


def _r1cs_bin(n_wires, n_pub_out, n_pub_in, n_prv_in, constraints, field_size=32):
    prime = bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")
    header = struct.pack("<I", field_size) + prime + struct.pack("<IIIIQI", n_wires, n_pub_out, n_pub_in, n_prv_in, n_wires, len(constraints))
    def vec(lc):
        return struct.pack("<I", len(lc)) + b"".join(struct.pack("<I", w) + v.to_bytes(32, "little") for w, v in lc)
    cons = b"".join(vec(a) + vec(b) + vec(c) for a, b, c in constraints)
    wmap = b"".join(struct.pack("<Q", i) for i in range(n_wires))
    out = b"r1cs" + struct.pack("<II", 1, 3)
    for t, body in ((2, cons), (1, header), (3, wmap)):  # sections in arbitrary order, as the format allows
        out += struct.pack("<IQ", t, len(body)) + body
    return out


def test_r1cs_binary_parser_matches_json_loader(simple_circuit):
    """src/r1cs_file.rs:100-154, src/reader.rs:227-241"""
    cons = simple_circuit.r1cs.constraints
    r1cs, wmap = reader.load_r1cs_from_bin(_r1cs_bin(4, 0, 1, 1, cons))
    assert (r1cs.num_inputs, r1cs.num_aux, r1cs.num_variables) == (2, 2, 4)
    assert wmap == [0, 1, 2, 3]
    def norm(c):
        return [sorted(lc) for lc in c]
    assert [norm(c) for c in r1cs.constraints] == [norm(c) for c in cons]
    with pytest.raises(ValueError, match="Invalid magic number"):
        reader.load_r1cs_from_bin(b"xxxx" + b"\0" * 64)
    bad = bytearray(_r1cs_bin(4, 0, 1, 1, cons))
    with pytest.raises(ValueError):
        reader.load_r1cs_from_bin(bytes(bad[:8]) + struct.pack("<I", 3) + bytes(bad[12:]).replace(struct.pack("<IQ", 1, 64), struct.pack("<IQ", 1, 63), 1))


# ---------------------------------------------------------------- circuit synthesis
def test_simple_circuit_gate_table_is_the_pinned_one(simple_circuit):
    """SURVEY App. A.2 table, derived from the golden vk/proof"""
    asm = circuit.synthesize(simple_circuit)
    assert asm.wire_idx[:, :4].T.tolist() == [[1, 0, 0, 0], [2, 2, 3, 0], [1, 2, 4, 0], [3, 2, 4, 0]]
    vals = limbs_to_ints(asm.var_values)
    assert vals == [0, 35, 3, 9, R_MOD - 27]
    assert circuit.is_satisfied(asm)
    assert circuit.transpile_with_gates_count(simple_circuit) == (3, 2)
    assert simple_circuit.get_public_inputs() == [35]
    bad = circuit.CircomCircuit(simple_circuit.r1cs, [1, 36, 3, 9])
    assert not circuit.is_satisfied(circuit.synthesize(bad))


def test_unpinned_r1cs_shapes_are_refused():
    r = circuit.R1CS(2, 4, 6, [([(2, 1), (3, 1)], [(2, 1)], [(4, 1)])])  # A side with two variables
    with pytest.raises(circuit.UnpinnedTranspilation):
        circuit.synthesize(circuit.CircomCircuit(r, [1, 1, 1, 1, 2, 0]))
    r = circuit.R1CS(2, 4, 6, [([(2, 1)], [(3, 1)], [(1, 1), (4, 1), (5, 1)])])  # C side with three variables
    with pytest.raises(circuit.UnpinnedTranspilation):
        circuit.synthesize(circuit.CircomCircuit(r, [1, 1, 1, 1, 2, 0]))
    # 0 * LC = 0 is skipped, not transpiled (src/circom_circuit.rs:122-123)
    r = circuit.R1CS(2, 2, 4, [([], [(2, 1)], [])])
    assert circuit.synthesize(circuit.CircomCircuit(r, [1, 5, 6, 7])).num_gates == 1


def test_general_transpiler_handles_wide_linear_combinations():
    """strict = False (BYTE PARITY UNPINNED, circuit._transpile): linear combinations of any length are laid out as running
    sums through d / q_dnext = -1; the pinned shapes come out exactly as in strict mode; every gate table is satisfied by
    the extended witness and rejects a wrong one."""
    import random
    rnd = random.Random(5)
    # pinned shapes: identical tables in both modes
    strict = circuit.synthesize(circuit.CircomCircuit(reader.load_r1cs(os.path.join(SIMPLE, "circuit.r1cs.json")),
                                                      reader.load_witness_from_file(os.path.join(SIMPLE, "witness.json"))))
    general = circuit.synthesize(circuit.CircomCircuit(reader.load_r1cs(os.path.join(SIMPLE, "circuit.r1cs.json")),
                                                       reader.load_witness_from_file(os.path.join(SIMPLE, "witness.json")), strict=False))
    assert (strict.wire_idx == general.wire_idx).all() and (strict.selectors == general.selectors).all()
    # random constraints  (sum a_i w_i + ka) * (sum b_i w_i + kb) = sum c_i w_i + kc  with 0..12 terms a side
    nv = 40
    for trial in range(30):
        wit = [1] + [rnd.randrange(R_MOD) for _ in range(nv - 1)]
        cons = []
        for _ in range(6):
            def lc(k):
                return [(rnd.randrange(0, nv - 1), rnd.randrange(1, R_MOD)) for _ in range(k)]
            A, B = lc(rnd.randrange(0, 13)), lc(rnd.randrange(0, 13))
            ev = lambda L: sum(c * wit[v] for v, c in L) % R_MOD  # noqa: E731
            # C = fresh wire holding the product, plus noise terms that cancel, so the constraint holds
            wit.append(ev(A) * ev(B) % R_MOD)
            extra = lc(rnd.randrange(0, 10))
            C = [(len(wit) - 1, 1)] + extra + [(0, (-ev(extra)) % R_MOD)]
            cons.append((A, B, C))
        r = circuit.R1CS(num_inputs=3, num_aux=len(wit) - 3, num_variables=len(wit), constraints=cons)
        asm = circuit.synthesize(circuit.CircomCircuit(r, wit, strict=False))
        assert circuit.is_satisfied(asm), trial
        bad = list(wit)
        bad[-1] = (bad[-1] + 1) % R_MOD
        assert not circuit.is_satisfied(circuit.synthesize(circuit.CircomCircuit(r, bad, strict=False)))
        assert circuit.synthesize(circuit.CircomCircuit(r, None, strict=False)).var_values is None  # setup-only synthesis
    with pytest.raises(ValueError, match="contradiction"):
        circuit.synthesize(circuit.CircomCircuit(circuit.R1CS(2, 0, 2, [([(0, 2)], [(0, 3)], [(0, 5)])]), [1, 7], strict=False))


def test_poseidon_shaped_r1cs_round_trips_and_transpiles(tmp_path):
    """synth.poseidon_r1cs: circomlib Poseidon(2)-shaped R1CS (244 constraints, combinations up to 61 terms) through the
    iden3 .r1cs / .wtns writers and the reference-format parsers, refused in strict mode, satisfied in general mode."""
    r1cs, wit = synth.poseidon_r1cs()
    assert len(r1cs.constraints) == 244 and max(len(b) for _, b, _ in r1cs.constraints) > 50
    synth.write_r1cs_bin(r1cs, str(tmp_path / "c.r1cs"))
    synth.write_wtns(wit, str(tmp_path / "w.wtns"))
    r2, w2 = reader.load_r1cs(str(tmp_path / "c.r1cs")), reader.load_witness_from_file(str(tmp_path / "w.wtns"))
    assert r2.constraints == r1cs.constraints and w2 == wit and (r2.num_inputs, r2.num_variables) == (2, r1cs.num_variables)
    with pytest.raises(circuit.UnpinnedTranspilation):
        circuit.synthesize(circuit.CircomCircuit(r2, w2))
    asm = circuit.synthesize(circuit.CircomCircuit(r2, w2, strict=False))
    assert circuit.is_satisfied(asm) and asm.n == 4096 and asm.num_inputs == 1


def test_synthetic_circuits_are_satisfied_and_sized():
    for log_n in (6, 9, 11):
        a = synth.poseidon_chain_assembly(log_n)
        assert a.n == 1 << log_n and a.num_gates == a.n - 1 and a.num_inputs == 1
        assert int(a.wire_idx.max()) == a.nvars - 1
        assert circuit.is_satisfied(a)
    b = synth.random_gate_assembly(8)
    assert b.n == 256 and circuit.is_satisfied(b)
    x = synth.random_field_elements(100)
    assert all(v < R_MOD for v in limbs_to_ints(x)) and len(set(limbs_to_ints(x))) == 100


def test_gen_key_range_check_mirrors_reference():
    """src/plonk.rs:31-34"""
    for power in (9, 27):
        with pytest.raises(ValueError, match="setup power of two is not in the correct range"):
            plonk.gen_key_monomial_form(power)


# ---------------------------------------------------------------- host build of the device arithmetic
@pytest.fixture(scope="module")
def hc(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("hc") / "host_check.so")
    subprocess.check_call(["g++", "-O2", "-frounding-math", "-std=c++17", "-shared", "-fPIC", "-x", "c++", os.path.join(ROOT, "tests", "host_check.cpp"), "-o", so])
    return ctypes.CDLL(so)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def test_montgomery_multiplier_sequence(hc):
    rng = np.random.default_rng(7)
    for which, p in ((0, R_MOD), (1, Q_MOD)):
        edge = [0, 1, 2, p - 1, p - 2, 1 << 253, (1 << 256) % p, (1 << 32) - 1, (((1 << 32) - 1) << 224) % p]
        a = edge * len(edge) + [int.from_bytes(rng.bytes(32), "little") % p for _ in range(3000)]
        b = [e for e in edge for _ in edge] + [int.from_bytes(rng.bytes(32), "little") % p for _ in range(3000)]
        A, B = ints_to_limbs(a).view(np.uint32), ints_to_limbs(b).view(np.uint32)
        out = np.zeros_like(A)
        hc.hc_mont_mul(which, _p(A), _p(B), _p(out), len(a))
        rinv = pow(1 << 256, -1, p)
        assert limbs_to_ints(out.view(np.uint64)) == [x * y * rinv % p for x, y in zip(a, b)]


def test_karatsuba_multiplier_sequences(hc):
    """fp.cuh mont_mul_k / mont_sqr_k / mont_mul_sub_mul (Karatsuba product, separate Montgomery reduction, lazy
    reduction of a difference of products) against Python integers: edge values, halves that are equal / ordered
    either way (both signs of the Karatsuba cross term), all-ones limbs (every carry path), random inputs."""
    rng = np.random.default_rng(10)
    ones = (1 << 128) - 1
    for which, p in ((0, R_MOD), (1, Q_MOD)):
        edge = [0, 1, 2, p - 1, p - 2, 1 << 253, (1 << 256) % p, (1 << 128) - 1, 1 << 128, (ones << 128 | ones) % p,
                (5 << 128) | 5, (7 << 128) | 3, (3 << 128) | 7, ((1 << 125) << 128) | ones, (1 << 253) | ones]
        pat = [sum(int(rng.choice([0, 0xffffffff, 0xfffffffe, 1, 0x80000000])) << (32 * i) for i in range(8)) % p for _ in range(400)]
        rnd = lambda k: [int.from_bytes(rng.bytes(32), "little") % p for _ in range(k)]
        a = edge * len(edge) + pat + rnd(4000)
        b = [e for e in edge for _ in edge] + pat[::-1] + rnd(4000)
        c = b[::-1][:len(a)]
        d = a[::-1][:len(a)]
        A, B, C, D = (ints_to_limbs(x).view(np.uint32) for x in (a, b, c, d))
        rinv = pow(1 << 256, -1, p)
        out = np.zeros_like(A)
        hc.hc_mont_wide(which, 0, _p(A), _p(B), _p(C), _p(D), _p(out), len(a))
        assert limbs_to_ints(out.view(np.uint64)) == [x * y * rinv % p for x, y in zip(a, b)]
        hc.hc_mont_wide(which, 1, _p(A), _p(B), _p(C), _p(D), _p(out), len(a))
        assert limbs_to_ints(out.view(np.uint64)) == [x * x * rinv % p for x in a]
        hc.hc_mont_wide(which, 2, _p(A), _p(B), _p(C), _p(D), _p(out), len(a))
        assert limbs_to_ints(out.view(np.uint64)) == [(x * y - z * w) * rinv % p for x, y, z, w in zip(a, b, c, d)]


def test_fp64_pipe_multiplier_sequence(hc):
    """fp_f64.cuh: the DFMA-based Montgomery product must agree with Python integers (and so with the IMAD one) on edge
    values, limb patterns that maximise every column chain, and random inputs."""
    rng = np.random.default_rng(9)
    m51 = (1 << 51) - 1
    for which, p in ((0, R_MOD), (1, Q_MOD)):
        edge = [0, 1, 2, p - 1, p - 2, 1 << 253, (1 << 256) % p, (1 << 51) - 1, (1 << 51), ((1 << 255) - 1) % p]
        pat = [sum(int(rng.choice([0, m51, m51 - 1, 1])) << (51 * i) for i in range(5)) % p for _ in range(200)]
        a = edge * len(edge) + pat + [int.from_bytes(rng.bytes(32), "little") % p for _ in range(5000)]
        b = [e for e in edge for _ in edge] + pat[::-1] + [int.from_bytes(rng.bytes(32), "little") % p for _ in range(5000)]
        A, B = ints_to_limbs(a).view(np.uint32), ints_to_limbs(b).view(np.uint32)
        out = np.zeros_like(A)
        hc.hc_mont_mul_f64(which, _p(A), _p(B), _p(out), len(a))
        rinv = pow(1 << 256, -1, p)
        assert limbs_to_ints(out.view(np.uint64)) == [x * y * rinv % p for x, y in zip(a, b)]


def test_field_ops_and_root_of_unity(hc):
    rng = np.random.default_rng(8)
    a = [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(40)] + [0, R_MOD - 1, 1]
    b = [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(40)] + [R_MOD - 1, R_MOD - 1, 0]
    A, B = ints_to_limbs(a), ints_to_limbs(b)
    outs = [np.zeros_like(A) for _ in range(4)]
    hc.hc_fr_ops(_p(A), _p(B), *[_p(o) for o in outs], len(a))
    assert limbs_to_ints(outs[0]) == [(x + y) % R_MOD for x, y in zip(a, b)]
    assert limbs_to_ints(outs[1]) == [(x - y) % R_MOD for x, y in zip(a, b)]
    assert limbs_to_ints(outs[2]) == [x * y % R_MOD for x, y in zip(a, b)]
    assert limbs_to_ints(outs[3]) == [pow(x, -1, R_MOD) if x else 0 for x in a]
    for log_n in (1, 3, 20, 28):
        out = np.zeros(4, dtype=np.uint64)
        hc.hc_root(log_n, _p(out))
        assert limbs_to_ints(out)[0] == root_of_unity(log_n)


def test_xyzz_formulas_against_oracle_group_law(hc, orc, simple_key):
    g = simple_key.g1_bases
    inf = np.zeros(8, dtype=np.uint64)
    neg5 = g[5].copy()
    neg5[4:] = ints_to_limbs([Q_MOD - limbs_to_ints(g[5][4:])[0]])[0]
    cases = [(g[3], g[7]), (g[5], g[5]), (g[5], neg5), (g[5], inf), (inf, g[9]), (inf, inf)]
    for p, q in cases:
        out = np.zeros((5, 8), dtype=np.uint64)
        hc.hc_g1_ops(_p(np.ascontiguousarray(p)), _p(np.ascontiguousarray(q)), 11, _p(out))
        assert (out[0] == orc.g1_add(p, q)).all()
        assert (out[4] == orc.g1_add(orc.g1_add(orc.g1_add(p, q), q), p)).all()   # lazy-form chain, normalised once
        assert (out[1] == orc.g1_add(p, q)).all()
        assert (out[2] == orc.g1_add(p, p)).all()
        assert (out[3] == orc.g1_mul(p, 11)).all()


def test_jacobian_window_arithmetic_against_oracle_group_law(hc, orc, simple_key):
    """ec.cuh g1_jac_t (the window loop of ecmul.cuh's scalar multiplication): XYZZ -> Jacobian -> XYZZ conversions, the
    a = 0 doubling and the addition of a table entry, with the special cases (equal, opposite, either operand infinity)."""
    g = simple_key.g1_bases
    inf = np.zeros(8, dtype=np.uint64)
    neg5 = g[5].copy()
    neg5[4:] = ints_to_limbs([Q_MOD - limbs_to_ints(g[5][4:])[0]])[0]
    for p, q in [(g[3], g[7]), (g[5], g[5]), (g[5], neg5), (g[5], inf), (inf, g[9]), (inf, inf), (g[40], g[2])]:
        out = np.zeros((3, 8), dtype=np.uint64)
        hc.hc_jac_ops(_p(np.ascontiguousarray(p)), _p(np.ascontiguousarray(q)), _p(out))
        assert (out[0] == orc.g1_add(p, q)).all()
        assert (out[1] == orc.g1_add(p, p)).all()
        assert (out[2] == orc.g1_add(orc.g1_add(p, q), q)).all()


def test_lazy_form_field_chain(hc):
    """fp.cuh lazy form ([0, 2p)): products without the final conditional subtraction, add/sub modulo 2p; a chain of
    them normalised once must equal the canonical computation (Montgomery factors included)."""
    rng = np.random.default_rng(12)
    p = Q_MOD
    edge = [0, 1, p - 1, p - 2, (1 << 253), (1 << 254) % p]
    a = edge * len(edge) + [int.from_bytes(rng.bytes(32), "little") % p for _ in range(3000)]
    b = [e for e in edge for _ in edge] + [int.from_bytes(rng.bytes(32), "little") % p for _ in range(3000)]
    A, B = ints_to_limbs(a).view(np.uint32), ints_to_limbs(b).view(np.uint32)
    out = np.zeros_like(A)
    hc.hc_lazy_field(_p(A), _p(B), _p(out), len(a))
    ri = pow(1 << 256, -1, p)
    want = [((x * y * ri) * (x + y) * ri - y * y * ri - (2 * x + y)) % p for x, y in zip(a, b)]
    assert limbs_to_ints(out.view(np.uint64)) == want


def test_glv_decomposition_and_scalar_multiplication(hc, orc, simple_key):
    """ecmul.cuh: k = k1 + k2 * LAMBDA with short halves, and k * P through the endomorphism == the plain 254-bit walk ==
    the oracle's scalar multiplication (which the EC inverse NTT of Crs::from_powers is made of)."""
    lam = 0xb3c4d79d41a917585bfc41088d8daaa78b17ea66b99c90dd
    assert (lam * lam + lam + 1) % R_MOD == 0
    rng = np.random.default_rng(13)
    ks = [0, 1, 2, R_MOD - 1, R_MOD - 2, lam, lam + 1, (1 << 253), (1 << 127), (1 << 128) - 1] + \
         [int.from_bytes(rng.bytes(32), "little") % R_MOD for _ in range(60)]
    inf = np.zeros(8, dtype=np.uint64)
    for n, k in enumerate(ks):
        p = simple_key.g1_bases[3 + n % 50] if n != 5 else inf
        kk = ints_to_limbs([k]).view(np.uint32)
        out_k = np.zeros(12, dtype=np.uint32)
        out_pt = np.zeros((4, 8), dtype=np.uint64)
        hc.hc_glv(_p(np.ascontiguousarray(p)), _p(kk), _p(out_k), _p(out_pt))
        k1 = sum(int(out_k[i]) << (32 * i) for i in range(5)) * (-1 if out_k[5] else 1)
        k2 = sum(int(out_k[6 + i]) << (32 * i) for i in range(5)) * (-1 if out_k[11] else 1)
        assert (k1 + k2 * lam - k) % R_MOD == 0 and abs(k1) < (1 << 129) and abs(k2) < (1 << 129)
        want = orc.g1_mul(p, k)
        assert all((out_pt[j] == want).all() for j in range(4)), hex(k)


def test_host_keccak_and_transcript_match_oracle(hc, orc):
    for msg in (b"", b"abc", b"a" * 135, b"a" * 136, b"a" * 137, bytes(range(200)) * 3):
        out = ctypes.create_string_buffer(32)
        hc.hc_keccak(msg, len(msg), out)
        assert out.raw == orc.keccak256(msg)
    # transcript of the golden proof's first round gives the pinned beta
    p = reader.load_proof(os.path.join(SIMPLE, "proof.bin"))
    vals = reader.limbs_to_be_bytes(ints_to_limbs(p.input_values)) + reader.g1_to_bytes(p.wire_commitments)
    vals = bytearray(vals)
    vals[32 + 64 * 3] = 0  # the infinity flag byte is absorbed as zero
    ch = np.zeros((2, 8), dtype=np.uint32)
    hc.hc_transcript(bytes(vals), 1 + 8, 2, _p(ch))
    got = limbs_to_ints(ch.view(np.uint64))
    assert got[0] == 0x0f72cf563829c88d02442b32aa5bc8b0aff226697faa846756e813710804a058
    assert got[1] == 0x19f776d072bc5715a7fb2a727344f31eda1aa1b214c0b8b7f0bf9a8fed192264


# ---------------------------------------------------------------- CLI surface (src/bin/main.rs:20-256)
def test_cli_analyse_and_flag_surface(tmp_path):
    from plonkit_b200 import __main__ as cli
    out = tmp_path / "analyse.json"
    cli.main(["analyse", "-c", os.path.join(SIMPLE, "circuit.r1cs.json"), "-o", str(out)])
    got = __import__("json").load(open(out))
    assert got["num_gates"] == 3 and got["num_hints"] == 2 and got["constraint_stats"][1] == {"name": "1", "num_gates": 2}
    p = cli.build_parser()
    o = p.parse_args(["prove", "-m", "k.key", "-c", "c.r1cs"])
    assert (o.witness, o.proof, o.proofjson, o.publicjson, o.transcript, o.overwrite) == \
        ("witness.wtns", "proof.bin", "proof.json", "public.json", "keccak", False)
    o = p.parse_args(["export-verification-key", "-m", "k.key"])
    assert o.vk == "vk.bin" and o.circuit is None
    assert cli.resolve_circuit_file(None) in ("circuit.r1cs", "circuit.json") and cli.resolve_circuit_file("x.json") == "x.json"
    with pytest.raises(SystemExit):
        cli.main(["recursive-prove"])
    # overwrite guard (main.rs:336-339)
    existing = tmp_path / "exists.key"
    existing.write_bytes(b"x")
    with pytest.raises(SystemExit, match="duplicate srs_monomial_form file"):
        cli._guard(str(existing), "srs_monomial_form", False)
    cli._guard(str(existing), "srs_monomial_form", True)


def test_verify_accepts_the_reference_proof_with_the_real_pairing(tmp_path):
    """src/tests.rs:75-81 (test_verify): plonk::verify(vk.bin, proof.bin, "keccak") == true — host arithmetic with the BN254
    optimal-ate pairing and the two G2 elements of vk.bin, no trapdoor.  Tampered proofs / keys are rejected; the CLI mirrors
    `plonkit verify` (exit code 400 on an invalid proof, src/bin/main.rs:427-438)."""
    from plonkit_b200 import __main__ as cli
    from plonkit_b200 import verifier
    vk = reader.load_verification_key(os.path.join(SIMPLE, "vk.bin"))
    proof = reader.load_proof(os.path.join(SIMPLE, "proof.bin"))
    assert plonk.verify(vk, proof, "keccak") is True
    # pairing sanity on its own: e(42 G, [1]G2) e(-G, [42]G2) == 1, and not with 43
    g2 = [verifier.g2_from_bytes(vk.g2_raw[:128]), verifier.g2_from_bytes(vk.g2_raw[128:])]
    assert verifier.pairing_product_is_one([(verifier.g1_mul((1, 2), 42), g2[0]), (verifier.g1_neg((1, 2)), g2[1])])
    assert not verifier.pairing_product_is_one([(verifier.g1_mul((1, 2), 43), g2[0]), (verifier.g1_neg((1, 2)), g2[1])])
    bad = reader.load_proof(os.path.join(SIMPLE, "proof.bin"))
    bad.quotient_polynomial_at_z = (bad.quotient_polynomial_at_z + 1) % R_MOD
    assert plonk.verify(vk, bad) is False                      # fails the identity at z
    bad2 = reader.load_proof(os.path.join(SIMPLE, "proof.bin"))
    bad2.opening_at_z_proof = bad2.opening_at_z_omega_proof.copy()
    assert plonk.verify(vk, bad2) is False                     # fails the pairing check only
    vk2 = reader.load_verification_key(os.path.join(SIMPLE, "vk.bin"))
    vk2.g2_raw = vk2.g2_raw[:128] + vk2.g2_raw[:128]           # wrong [x]G2
    assert plonk.verify(vk2, proof) is False
    with pytest.raises(NotImplementedError):
        plonk.verify(vk, proof, "rescue")
    cli.main(["verify", "-p", os.path.join(SIMPLE, "proof.bin"), "-v", os.path.join(SIMPLE, "vk.bin")])
    pb = tmp_path / "bad.bin"
    pb.write_bytes(bad.to_bytes())
    with pytest.raises(SystemExit) as e:
        cli.main(["verify", "-p", str(pb), "-v", os.path.join(SIMPLE, "vk.bin")])
    assert e.value.code == 400


def test_verify_accepts_a_device_made_proof_fixture():
    """tests/golden/poseidon9_proof.bin (made by the CUDA prover, == the oracle's bytes) with a verification key assembled from
    the committed commitments and the constant G2 part of Crs::crs_42: accepted by the pairing-based verifier."""
    from conftest import GOLDEN
    com = np.load(os.path.join(GOLDEN, "poseidon9_vk_commitments.npy"))
    proof = reader.load_proof(os.path.join(GOLDEN, "poseidon9_proof.bin"))
    vk = reader.VerificationKey(proof.n, proof.num_inputs, com[:6], com[6:7], com[7:11], [5, 7, 10], reader.CRS_42_G2)
    assert plonk.verify(vk, proof) is True


def test_two_gate_type_verifier_accepts_the_oracle_proof_with_the_real_pairing(orc):
    """The pairing-based verifier of the two-gate-type prover (recursive.verify, the proof check of src/recursive/mod.rs:139-166;
    parity unpinned) accepts the oracle's proof of a 2^6 rescue-shaped circuit against the 13 setup commitments and the G2 part
    of Crs::crs_42, agrees with the oracle's trapdoor check, and rejects a tampered evaluation and a swapped opening."""
    from plonkit_b200 import recursive
    asm, gate_type = synth.rescue_chain_assembly(6)
    srs = orc.srs_gen(asm.n, 42, threads=2)
    raw = orc.prove2(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, gate_type, srs, threads=2)
    com = orc.setup_commitments2(asm.n, asm.num_inputs, asm.wire_idx, asm.selectors, gate_type, srs, nvars=asm.nvars, threads=2)
    proof = reader.Proof.read(io.BytesIO(raw), gated=True)
    assert proof.to_bytes() == raw
    vk = recursive.RecursiveVerificationKey(asm.n - 1, asm.num_inputs, com, reader.CRS_42_G2)
    assert orc.verify_trapdoor2(raw, com, 42)
    assert recursive.verify(vk, proof) is True
    bad = reader.Proof.read(io.BytesIO(raw), gated=True)
    bad.gate_selectors_at_z = [proof.gate_selectors_at_z[0], (proof.gate_selectors_at_z[1] + 1) % R_MOD]
    assert recursive.verify(vk, bad) is False
    bad = reader.Proof.read(io.BytesIO(raw), gated=True)
    bad.opening_at_z_proof = proof.opening_at_z_omega_proof
    assert recursive.verify(vk, bad) is False
    assert recursive.verify(recursive.RecursiveVerificationKey(asm.n - 1, asm.num_inputs, com, b""), proof) is False


def test_verifier_handles_a_selector_commitment_at_infinity(orc):
    """An addition-only circuit has q_m = 0, so its commitment in the verification key is the point at infinity ((0, 0) in
    vk.bin): the oracle's proof of it must be accepted by the pairing-based verifier, which skips that term, absorbs nothing
    for it in the transcript (vk commitments are not absorbed) and still rejects a wrong public input."""
    m1 = R_MOD - 1
    rows = [(1, 0, 0, 0, [m1, 0, 0, 0, 0, 0, 0]),            # public input x
            (1, 2, 3, 0, [1, 1, m1, 0, 0, 0, 0]),            # x + y - z = 0
            (3, 2, 4, 0, [1, 1, m1, 0, 0, 0, 0])]            # z + y - w = 0
    asm = circuit.assembly_from_rows(rows, [0, 3, 4, 7, 11], 1)
    assert asm.n == 4 and circuit.is_satisfied(asm)
    srs = orc.srs_gen(asm.n, 42, threads=1)
    raw = orc.prove(asm.n, asm.num_inputs, asm.wire_idx, asm.var_values, asm.selectors, srs, threads=1)
    com = orc.setup_commitments(asm.n, asm.num_inputs, asm.wire_idx, asm.selectors, srs, nvars=asm.nvars, threads=1)
    assert not com[4].any()                                   # [q_m] = infinity
    proof = reader.Proof.read(io.BytesIO(raw))
    vk = reader.VerificationKey(asm.n - 1, 1, com[:6], com[6:7], com[7:11], [5, 7, 10], reader.CRS_42_G2)
    assert plonk.verify(vk, proof) is True
    proof.input_values = [4]
    assert plonk.verify(vk, proof) is False


def test_witness_plan_replays_synthesis_natively_and_in_python():
    """circuit.WitnessPlan (the per-proof half of synthesis; bellman redoes all of it inside every prove, src/plonk.rs:132-176):
    var_values from a witness through the native host library (csrc/host/witness.cpp, 1 and 8 threads) and through Python
    integers == synthesize(circuit).var_values — for the reference's simple circuit (strict layout), for a Poseidon-shaped
    R1CS with ~60-term combinations (running sums), with a wire mapping, with the witness given as limbs, and for ANOTHER
    witness of the same R1CS assigned through the plan of the first."""
    assert circuit.host_library() is not None, "libplonkit_host.so is not built (python __graft_entry__.py)"
    simple = circuit.CircomCircuit(reader.load_r1cs(os.path.join(SIMPLE, "circuit.r1cs.json")),
                                   reader.load_witness_from_file(os.path.join(SIMPLE, "witness.json")))
    r1cs, wit = synth.poseidon_r1cs(20)
    pos = circuit.CircomCircuit(r1cs, wit, None, circuit.AUX_OFFSET, False)
    rng = np.random.default_rng(5)
    perm = [0] + [int(x) + 1 for x in rng.permutation(r1cs.num_variables - 1)]      # variable i reads witness[perm[i]]
    shuffled = [0] * len(wit)
    for i, j in enumerate(perm):
        shuffled[j] = wit[i]
    mapped = circuit.CircomCircuit(r1cs, shuffled, perm, circuit.AUX_OFFSET, False)
    for c in (simple, pos, mapped):
        asm = circuit.synthesize(c)
        assert asm.plan.nvars == asm.nvars
        for kw in ({"native": True, "threads": 1}, {"native": True, "threads": 8}, {"native": False}):
            assert (asm.plan.assign(c.witness, c.wire_mapping, **kw) == asm.var_values).all(), kw
        assert (asm.plan.assign(ints_to_limbs(c.witness), c.wire_mapping) == asm.var_values).all()
    # another witness of the same R1CS: the plan of the first circuit, no second transpilation
    plan = circuit.synthesize(pos).plan
    r2, wit2 = synth.poseidon_r1cs(20, inputs=(11, 12))
    assert len(r2.constraints) == len(r1cs.constraints) and wit2 != wit
    want = circuit.synthesize(circuit.CircomCircuit(r2, wit2, None, circuit.AUX_OFFSET, False)).var_values
    assert (plan.assign(wit2) == want).all() and (plan.assign(wit2, native=False) == want).all()
    # what SetupForProver keeps of its circuit: same R1CS object -> the plan; anything else -> a fresh synthesis
    src = plonk._WitnessSource(pos, circuit.synthesize(pos))
    other = circuit.CircomCircuit(r1cs, wit2, None, circuit.AUX_OFFSET, False)
    assert (src.values(other) == want).all()
    assert (src.values(circuit.CircomCircuit(r2, wit2, None, circuit.AUX_OFFSET, False)) == want).all()
    assert src.values(circuit.CircomCircuit(r1cs, None, None, circuit.AUX_OFFSET, False)) is None
    # the CLI's flow: set up from the witness-less circuit, then prove the one holding the witness as limbs
    shape = circuit.CircomCircuit(r1cs, None, None, circuit.AUX_OFFSET, False)
    asm0 = circuit.synthesize(shape)
    assert asm0.var_values is None and (asm0.selectors == circuit.synthesize(pos).selectors).all()
    held = circuit.CircomCircuit(r1cs, ints_to_limbs(wit2), None, circuit.AUX_OFFSET, False)
    assert (plonk._WitnessSource(shape, asm0).values(held) == want).all()
    # errors: a short witness, a value outside the field, a program that reads ahead
    with pytest.raises(ValueError, match="witness holds"):
        plan.assign(wit[:10])
    bad = ints_to_limbs(wit)
    bad[3] = ints_to_limbs([R_MOD])[0] + np.array([5, 0, 0, 0], dtype=np.uint64)
    with pytest.raises(ValueError, match="variable 3"):
        plan.assign(bad)
    with pytest.raises(ValueError, match="before it is assigned"):
        circuit.WitnessPlan(3, [(0, [(1, 1)]), (0, [(5, 1)])])


class _BoundaryRecorder:
    """Stands where the CUDA library's context would: records what the host layer hands across the C ABI (gate tables at
    pk_setup_create, variable values at pk_prove / pk_witness_upload) and returns an all-zero proof.  No compute."""

    class _Lib:
        def __init__(self):
            self.values, self.uploaded, self.lagrange = None, None, None

        def pk_setup_create(self, h, a_ref, out_ref):
            a = a_ref._obj
            n = int(a.n)
            self.meta = (n, int(a.num_inputs), int(a.nvars))
            self.wire_idx = np.ctypeslib.as_array((ctypes.c_uint32 * (4 * n)).from_address(a.wire_idx)).reshape(4, n).copy()
            self.selectors = np.ctypeslib.as_array((ctypes.c_uint64 * (28 * n)).from_address(a.selectors)).reshape(7, n, 4).copy()
            out_ref._obj.value = 77
            return 0

        def pk_setup_use_lagrange(self, h, sh, flag):
            self.lagrange = flag
            return 0

        def pk_prove(self, h, sh, vals_ptr, nvars, pr_ref, inputs_ptr):
            self.values = None if not vals_ptr else \
                np.ctypeslib.as_array((ctypes.c_uint64 * (4 * nvars)).from_address(vals_ptr)).reshape(nvars, 4).copy()
            pr_ref._obj.n, pr_ref._obj.num_inputs = self.meta[0] - 1, self.meta[1]
            return 0

        def pk_witness_upload(self, h, sh, vals_ptr, nvars):
            self.uploaded = np.ctypeslib.as_array((ctypes.c_uint64 * (4 * nvars)).from_address(vals_ptr)).reshape(nvars, 4).copy()
            return 0

        def pk_setup_destroy(self, sh):
            return 0

        # the sharded prover's entry points take the same arguments
        pk_dist_setup_create, pk_dist_prove, pk_dist_witness_upload, pk_dist_setup_destroy = \
            pk_setup_create, pk_prove, pk_witness_upload, pk_setup_destroy

    def __init__(self, rank=0, world=1):
        self._lib, self._h, self._children = self._Lib(), 1, set()
        self.srs_tag = self.lagrange_tag = None
        self.rank, self.world, self.loaded = rank, world, None

    def _check(self, rc):
        assert rc == 0

    def srs_load_g1(self, bases, tag=None):
        self.srs_tag, self.loaded = tag, bases

    def srs_load_g1_lagrange(self, bases, tag=None):
        self.lagrange_tag = tag


def test_host_layer_hands_the_synthesised_tables_and_values_across_the_c_abi(tmp_path, monkeypatch, simple_key):
    """Everything between the reference-facing calls and the C ABI, without a device: `SetupForProver` and the CLI's `prove`
    must pass pk_setup_create the gate tables of synthesize(circuit) and pk_prove the variable values of synthesize(circuit)
    — whether the witness comes with the circuit, as another witness of the same R1CS (WitnessPlan), as an Assembly, as an
    array, or from witness.json / .wtns files through the CLI."""
    from plonkit_b200 import __main__ as cli
    r1cs = reader.load_r1cs(os.path.join(SIMPLE, "circuit.r1cs.json"))
    wit = reader.load_witness_from_file(os.path.join(SIMPLE, "witness.json"))
    c = circuit.CircomCircuit(r1cs, wit)
    want = circuit.synthesize(c)
    rec = _BoundaryRecorder()
    setup = plonk.SetupForProver.prepare_setup_for_prover(c, simple_key, None, ctx=rec)
    assert rec._lib.meta == (want.n, want.num_inputs, want.nvars)
    assert (rec._lib.wire_idx == want.wire_idx).all() and (rec._lib.selectors == want.selectors).all()
    proof = setup.prove(c, "keccak")
    assert (rec._lib.values == want.var_values).all() and proof.n == want.n - 1 and rec._lib.lagrange == 0
    for form in (want, want.var_values, circuit.CircomCircuit(r1cs, ints_to_limbs(wit)),
                 circuit.CircomCircuit(reader.load_r1cs(os.path.join(SIMPLE, "circuit.r1cs.json")), wit)):
        rec._lib.values = None
        setup.prove(form)
        assert (rec._lib.values == want.var_values).all()
    setup.upload_witness(c)
    assert (rec._lib.uploaded == want.var_values).all()
    setup.prove(None)
    assert rec._lib.values is None                                  # the device-resident witness is used
    with pytest.raises(_lib.SynthesisError):
        setup.prove(circuit.CircomCircuit(r1cs, None))
    setup.close()
    # one rank of the sharded prover: its chunk of the key, the same tables, the same values
    rec1 = _BoundaryRecorder(rank=1, world=2)
    sh = plonk.ShardedSetupForProver.prepare_setup_for_prover(c, simple_key, rec1)
    assert (rec1.loaded == simple_key.g1_bases[want.n // 2:want.n]).all() and (rec1._lib.selectors == want.selectors).all()
    for form in (c, circuit.CircomCircuit(r1cs, ints_to_limbs(wit)), want, want.var_values):
        rec1._lib.values = None
        sh.prove(form)
        assert (rec1._lib.values == want.var_values).all()
    sh.upload_witness(want.var_values)
    assert (rec1._lib.uploaded == want.var_values).all()
    sh.close()
    # a pool: ONE transpilation, every prover assigns other witnesses of the R1CS through the shared plan
    made = []
    monkeypatch.setattr(plonk, "Context", lambda device=0: made.append(_BoundaryRecorder()) or made[-1])
    calls = []
    real_transpile = circuit._transpile
    monkeypatch.setattr(circuit, "_transpile", lambda *a, **k: calls.append(1) or real_transpile(*a, **k))
    pool = plonk.ProverPool(c, simple_key, inflight=2)
    pool.prove_all([circuit.CircomCircuit(r1cs, wit), circuit.CircomCircuit(r1cs, ints_to_limbs(wit)), want.var_values])
    assert len(made) == 2 and len(calls) == 1
    assert all(m._lib.values is None or (m._lib.values == want.var_values).all() for m in made)
    assert any(m._lib.values is not None for m in made)
    monkeypatch.setattr(circuit, "_transpile", real_transpile)
    # the CLI, golden simple circuit (strict mode) and a Poseidon-shaped .r1cs / .wtns pair (general mode)
    rec2 = _BoundaryRecorder()
    monkeypatch.setattr(plonk, "default_context", lambda device=0: rec2)
    key = os.path.join(SIMPLE, "setup_2^10.key")
    cli.main(["prove", "-m", key, "-c", os.path.join(SIMPLE, "circuit.r1cs.json"), "-w", os.path.join(SIMPLE, "witness.json"),
              "-p", str(tmp_path / "p.bin"), "-j", str(tmp_path / "p.json"), "-i", str(tmp_path / "i.json")])
    assert (rec2._lib.values == want.var_values).all() and (rec2._lib.selectors == want.selectors).all()
    pr1cs, pwit = synth.poseidon_r1cs(1)
    synth.write_r1cs_bin(pr1cs, str(tmp_path / "c.r1cs"))
    synth.write_wtns(pwit, str(tmp_path / "w.wtns"))
    pwant = circuit.synthesize(circuit.CircomCircuit(pr1cs, pwit, None, circuit.AUX_OFFSET, False))
    key = str(tmp_path / "big.key")                                  # the recorder never reads the bases: repeat the golden ones
    with open(key, "wb") as f:
        reader.Crs(np.tile(simple_key.g1_bases, (pwant.n // 1024, 1)), reader.CRS_42_G2).write(f)
    cli.main(["prove", "-m", key, "-c", str(tmp_path / "c.r1cs"), "-w", str(tmp_path / "w.wtns"), "--allow-unpinned-transpilation",
              "-p", str(tmp_path / "p2.bin"), "-j", str(tmp_path / "p2.json"), "-i", str(tmp_path / "i2.json")])
    assert rec2._lib.meta == (pwant.n, pwant.num_inputs, pwant.nvars)
    assert (rec2._lib.values == pwant.var_values).all() and (rec2._lib.wire_idx == pwant.wire_idx).all()
    assert (rec2._lib.selectors == pwant.selectors).all()


def _random_lc(rng, nv, kind):
    if kind == "empty":
        return []
    if kind == "const":
        return [(0, rng.choice([1, 5, R_MOD - 1, rng.randrange(R_MOD)]))]
    k = {"one": 1, "two": 2, "few": rng.randint(1, 4), "many": rng.randint(5, 14)}[kind]
    lc = [(rng.randrange(1, nv), rng.choice([1, 2, R_MOD - 1, rng.randrange(R_MOD), 0])) for _ in range(k)]
    if kind in ("few", "many"):
        if rng.random() < 0.3:
            lc.append((0, rng.randrange(R_MOD)))
        if rng.random() < 0.2:
            lc.append((lc[0][0], rng.randrange(R_MOD)))                # the same variable twice
        if rng.random() < 0.1:
            lc.append((lc[0][0], (R_MOD - lc[0][1]) % R_MOD))          # ... cancelling
    rng.shuffle(lc)
    return lc


def test_compiled_transpiler_equals_the_python_statement_on_random_circuits():
    """csrc/host/transpile.cpp against circuit._transpile_py (the readable statement of the same layout): gate tables, variable
    values, witness program, analyse() output, gate counts and the error raised (unpinned shape in strict mode, contradiction)
    are identical on random R1CS — empty and constant sides, duplicate and cancelling terms, zero coefficients, long
    combinations — in both modes; strict mode also on circuits made of pinned shapes only."""
    import random
    assert circuit.host_library() is not None, "libplonkit_host.so is not built (python __graft_entry__.py)"
    outcomes = {}
    for seed in range(500):
        for strict in (False, True):
            rng = random.Random(seed)
            nv, ni = rng.randint(4, 30), rng.randint(1, 3)
            pinned_only = strict and seed % 2 == 0
            cons = []
            for _ in range(rng.randint(1, 25)):
                if pinned_only:
                    def coef():
                        return rng.choice([1, 2, R_MOD - 1, rng.randrange(1, R_MOD)])
                    c_side = rng.choice(["one", "two", "two+const"])
                    cc = [(v, coef()) for v in rng.sample(range(1, nv), 1 if c_side == "one" else 2)]
                    if c_side == "two+const":
                        cc.append((0, rng.randrange(R_MOD)))
                    cons.append(([(rng.randrange(1, nv), coef())], [(rng.randrange(1, nv), coef())], cc))
                else:
                    kinds, w = ["empty", "const", "one", "two", "few", "many"], [8, 8, 40, 20, 25, 15]
                    cons.append(tuple(_random_lc(rng, nv, rng.choices(kinds, w)[0]) for _ in range(3)))
            r = circuit.R1CS(ni, nv - ni, nv, cons)
            c = circuit.CircomCircuit(r, [1] + [rng.randrange(R_MOD) for _ in range(nv - 1)], None, circuit.AUX_OFFSET, strict)
            res = []
            for native in (True, False):
                circuit.NATIVE[0] = native
                try:
                    a = circuit.synthesize(c)
                    res.append(("ok", a.n, a.num_inputs, a.nvars, a.num_gates, a.wire_idx.tobytes(), a.selectors.tobytes(),
                                a.var_values.tobytes(), a.plan.off.tobytes(), a.plan.term_var.tobytes(), a.plan.term_coef.tobytes(),
                                a.plan.consts.tobytes(), str(circuit.analyse(c)), circuit.transpile_with_gates_count(c)))
                except (circuit.UnpinnedTranspilation, ValueError) as e:
                    res.append((type(e).__name__, str(e)))
                finally:
                    circuit.NATIVE[0] = True
            assert res[0] == res[1], (seed, strict)
            outcomes[(strict, res[0][0])] = outcomes.get((strict, res[0][0]), 0) + 1
    assert outcomes.get((True, "ok"), 0) > 50 and outcomes.get((False, "ok"), 0) > 100      # both modes really transpile
    assert outcomes.get((True, "UnpinnedTranspilation"), 0) > 50 and outcomes.get((False, "ValueError"), 0) > 10


def test_compiled_r1cs_parser_equals_the_python_statement(tmp_path):
    """csrc/host/transpile.cpp's `.r1cs` parser against reader._load_r1cs_from_bin_py: same R1CS, same wire map, same refusals
    (magic, version, header size, field size, prime, coefficient outside the field, map size, wire 0, truncation)."""
    import random
    assert circuit.host_library() is not None

    def both(buf):
        out = []
        for native in (True, False):
            circuit.NATIVE[0] = native
            try:
                r, wmap = reader.load_r1cs_from_bin(buf)
                out.append(("ok", r.num_inputs, r.num_aux, r.num_variables, r.num_constraints, r.constraints, wmap))
            except ValueError as e:
                out.append(("ValueError", str(e)))
            finally:
                circuit.NATIVE[0] = True
        return out
    rng = random.Random(7)
    for case in range(40):
        nv = rng.randint(3, 40)
        cons = [tuple(_random_lc(rng, nv, rng.choice(["empty", "const", "one", "two", "few", "many"])) for _ in range(3))
                for _ in range(rng.randint(0, 30))]
        cons = [tuple([(v, c % R_MOD) for v, c in lc] for lc in sides) for sides in cons]
        buf = _r1cs_bin(nv, rng.randint(0, 1), 1, nv - 3, cons)
        a, b = both(buf)
        assert a == b and a[0] == "ok" and a[5] == [tuple(list(lc) for lc in sides) for sides in cons]
    r1cs, _ = synth.poseidon_r1cs(1)
    synth.write_r1cs_bin(r1cs, str(tmp_path / "p.r1cs"))
    good = open(tmp_path / "p.r1cs", "rb").read()
    a, b = both(good)
    assert a == b and a[0] == "ok" and a[4] == len(r1cs.constraints)
    small = _r1cs_bin(4, 0, 1, 1, [([(2, 1)], [(2, 1)], [(3, 1)])])
    broken = {
        "magic": b"x" + small[1:],
        "version": small[:4] + struct.pack("<I", 2) + small[8:],
        "header size": small.replace(struct.pack("<IQ", 1, 64), struct.pack("<IQ", 1, 63), 1),
        "field size": _r1cs_bin(4, 0, 1, 1, [], field_size=31),
        "prime": small.replace(bytes.fromhex("010000f093f5e143"), bytes.fromhex("020000f093f5e143"), 1),
        "coefficient": _r1cs_bin(4, 0, 1, 1, [([(2, 1)], [(2, 1)], [(3, 1)])]).replace((1).to_bytes(32, "little"), R_MOD.to_bytes(32, "little"), 1),
    }
    for name, buf in broken.items():
        a, b = both(buf)
        assert a == b and a[0] == "ValueError", (name, a, b)
    for cut in (6, 20, 60, 100, len(small) - 3):
        a, b = both(small[:cut])
        assert a[0] == b[0] == "ValueError", (cut, a, b)


def test_constraint_with_a_wire_beyond_the_circuit_is_refused():
    r = circuit.R1CS(2, 2, 4, [([(2, 1)], [(7, 1)], [(3, 1)])])
    for native in (True, False):
        circuit.NATIVE[0] = native
        try:
            with pytest.raises(ValueError, match="wire 7"):
                circuit.synthesize(circuit.CircomCircuit(r, None))
        finally:
            circuit.NATIVE[0] = True


def test_general_mode_layout_is_frozen():
    """This repository's own layout for unpinned constraint shapes (running sums through d / q_dnext) must not
    drift between rounds: digest of the gate tables and variable values of the Poseidon(2)-shaped circuit, both transpilers."""
    import hashlib
    r1cs, wit = synth.poseidon_r1cs(1)
    for native in (True, False):
        circuit.NATIVE[0] = native
        try:
            a = circuit.synthesize(circuit.CircomCircuit(r1cs, wit, None, circuit.AUX_OFFSET, False))
        finally:
            circuit.NATIVE[0] = True
        assert (a.n, a.num_gates, a.nvars) == (4096, 2309, 2311)
        assert hashlib.sha256(a.wire_idx.tobytes() + a.selectors.tobytes() + a.var_values.tobytes()).hexdigest() == \
            "9f5c195735571140fda1ccc06dc2840629fc2b106711c970fcdb7a0d55f7e056"


def test_concurrent_synthesis_and_assignment_on_shared_objects():
    """The ranks of a sharded prover (threads) transpile the same R1CS object at once, and the provers of a pool assign witnesses
    through one shared plan at once: lazily made state (the library's copy of the R1CS, the Montgomery coefficients of the
    plan) must be complete before another thread can see it."""
    import threading
    for trial in range(5):
        r1cs, wit = synth.poseidon_r1cs(6)
        want = None
        circuit.NATIVE[0] = False
        try:
            want = circuit.synthesize(circuit.CircomCircuit(r1cs, wit, None, circuit.AUX_OFFSET, False))
        finally:
            circuit.NATIVE[0] = True
        fresh = circuit.R1CS(r1cs.num_inputs, r1cs.num_aux, r1cs.num_variables, r1cs.constraints)
        out, errs = [None] * 8, []

        def run(i):
            try:
                a = circuit.synthesize(circuit.CircomCircuit(fresh, None, None, circuit.AUX_OFFSET, False))
                out[i] = (a, a.plan.assign(wit))
            except Exception as e:   # noqa: BLE001
                errs.append(e)
        th = [threading.Thread(target=run, args=(i,)) for i in range(8)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs
        for a, v in out:
            assert (a.selectors == want.selectors).all() and (a.wire_idx == want.wire_idx).all() and (v == want.var_values).all()
        shared = out[0][0].plan
        shared._coef_mont = None
        res = [None] * 8

        def assign(i):
            res[i] = shared.assign(wit, threads=2)
        th = [threading.Thread(target=assign, args=(i,)) for i in range(8)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert all((v == want.var_values).all() for v in res)


def test_compiled_r1cs_parser_survives_hostile_files():
    """Mutated `.r1cs` images (flipped bytes, truncations, huge counts, inserted junk) through the compiled parser and, when it
    accepts, the compiled transpiler: every call returns — a parsed circuit or a ValueError, never a crash or an allocation
    sized by a number from the file — and whatever both parsers accept, they read identically.  (The same mutations ran
    clean under AddressSanitizer / UBSan builds of libplonkit_host.so.)"""
    import random
    rng = random.Random(11)
    small = _r1cs_bin(6, 1, 1, 3, [([(2, 1), (0, 5)], [(3, 7)], [(4, 1), (5, R_MOD - 1)]), ([], [(1, 1)], [])])
    accepted = refused = 0
    for it in range(1500):
        b = bytearray(small)
        kind = rng.randrange(4)
        if kind == 0:
            for _ in range(rng.randint(1, 4)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        elif kind == 1:
            b = b[:rng.randrange(len(b))]
        elif kind == 2:
            pos = rng.randrange(0, len(b) - 8)
            b[pos:pos + 8] = struct.pack("<Q", rng.choice([0, 1, 2 ** 32 - 1, 2 ** 63, 2 ** 64 - 1, rng.randrange(2 ** 64)]))
        else:
            pos = rng.randrange(len(b))
            b[pos:pos] = bytes(rng.randrange(256) for _ in range(rng.randint(1, 40)))
        try:
            r, wmap = reader.load_r1cs_from_bin(bytes(b))
        except ValueError:
            refused += 1
            continue
        accepted += 1
        try:
            py, pymap = reader._load_r1cs_from_bin_py(bytes(b))
        except ValueError:
            py = None
        if py is not None:
            assert py == r and pymap == wmap
        try:
            circuit.synthesize(circuit.CircomCircuit(r, None, None, circuit.AUX_OFFSET, False))
        except (ValueError, circuit.UnpinnedTranspilation):
            pass
    assert accepted > 100 and refused > 100


def test_key_proof_and_vk_readers_refuse_malformed_files_cleanly():
    """Mutated golden setup_2^10.key / proof.bin / vk.bin (flipped bytes, truncations, huge counts): every read returns an object
    or raises ValueError — never another exception type, a partial object, or a read sized by a number from the file."""
    import random
    files = {"key": (os.path.join(SIMPLE, "setup_2^10.key"), lambda b: reader.Crs.read(io.BytesIO(b))),
             "proof": (os.path.join(SIMPLE, "proof.bin"), lambda b: reader.Proof.read(io.BytesIO(b))),
             "vk": (os.path.join(SIMPLE, "vk.bin"), lambda b: reader.VerificationKey.read(io.BytesIO(b)))}
    blobs = {k: open(p, "rb").read() for k, (p, _) in files.items()}
    rng = random.Random(3)
    refused = 0
    for it in range(900):
        name = rng.choice(sorted(files))
        b = bytearray(blobs[name])
        kind = rng.randrange(3)
        if kind == 0:
            for _ in range(rng.randint(1, 4)):
                b[rng.randrange(min(len(b), 400))] = rng.randrange(256)
        elif kind == 1:
            b = b[:rng.randrange(len(b))]
        else:
            pos = rng.randrange(0, min(len(b) - 8, 64))
            b[pos:pos + 8] = struct.pack(">Q", rng.choice([0, 1, 2 ** 32 - 1, 2 ** 63, 2 ** 64 - 1, rng.randrange(2 ** 40)]))
        try:
            files[name][1](bytes(b))
        except ValueError:
            refused += 1
    assert refused > 200
    with pytest.raises(ValueError, match="expected 4"):
        reader.Proof.read(io.BytesIO(blobs["proof"][:48] + struct.pack(">Q", 5) + blobs["proof"][56:]))
