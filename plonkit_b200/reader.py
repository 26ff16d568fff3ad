"""File formats on either side of the prove path (wire contract of the reference).

Mirrors src/reader.rs (loaders) and the bellman `Crs::{read,write}`, `Proof::{read,write}`,
`VerificationKey::{read,write}` encodings those loaders call: big-endian, canonical field elements,
G1 uncompressed = x || y with infinity = 0x40 followed by 63 zero bytes.  In memory everything is
little-endian u64 limbs (see bn254.py).
"""
import itertools
import json
import struct
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from .bn254 import R_MOD, be_bytes_to_limbs, ints_to_limbs, limbs_to_be_bytes, limbs_to_ints
from .circuit import R1CS

_PRIME_LE = bytes.fromhex("010000f093f5e1439170b97948e833285d588181b64550b829a031e1724e6430")  # reader.rs:155


# ---------------------------------------------------------------- G1 encodings
def g1_from_bytes(buf: bytes, n: int) -> np.ndarray:
    """n uncompressed points (64 B each) -> (n, 8) uint64 canonical LE, infinity -> (0, 0)"""
    raw = np.frombuffer(buf, dtype=np.uint8, count=64 * n).reshape(n, 64)
    flags = raw[:, 0] & 0xC0
    pts = be_bytes_to_limbs(raw.tobytes(), 2 * n).reshape(n, 8)
    inf = flags != 0
    if inf.any():
        if (flags[inf] != 0x40).any():
            raise ValueError("compressed G1 encodings are not expected in .key/.bin files")
        pts[inf] = 0
    return pts


def g1_to_bytes(pts) -> bytes:
    pts = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 8)
    out = bytearray(limbs_to_be_bytes(pts.reshape(-1, 4)))
    inf = ~pts.any(axis=1)
    for i in np.nonzero(inf)[0]:
        out[64 * i] = 0x40
    return bytes(out)


# ---------------------------------------------------------------- SRS (.key): Crs::read / Crs::write  (App. B.1)
# The two G2 elements of Crs::crs_42 (src/plonk.rs:41,47): [1]G2 and [42]G2, 128 B each (x.c1 | x.c0 | y.c1 | y.c0, big
# endian), exactly as keys/setup/setup_2^10.key carries them.  G2 arithmetic (verify side) is outside this repository,
# and the trapdoor of crs_42 is the constant 42, so the encodings are constants.
CRS_42_G2 = bytes.fromhex(
    "198e9393920d483a7260bfb731fb5d25f1aa493335a9e71297e485b7aef312c2"
    "1800deef121f1e76426a00665e5c4479674322d4f75edadd46debd5cd992f6ed"
    "090689d0585ff075ec9e99ad690c3395bc4b313370b38ef355acdadcd122975b"
    "12c85ea5db8c6deb4aab71808dcb408fe3d1e7690c43d37b4ce6cc0166fa7daa"
    "12740934ba9615b77b6a49b06fcce83ce90d67b1d0e2a530069e3a7306569a91"
    "116da8c89a0d090f3d8644ada33a5f1c8013ba7204aeca62d66d931b99afe6e7"
    "25222d9816e5f86b4a7dedd00d04acc5c979c18bd22b834ea8c6d07c0ba441db"
    "076441042e77b6309644b56251f059cf14befc72ac8a6157d30924e58dc4c172")


_crs_tokens = itertools.count(1)


def _read_exact(f, n: int, what: str) -> bytes:
    # n may come straight from the file: compare it with what is left before asking for that many bytes
    try:
        pos = f.tell()
        left = f.seek(0, 2) - pos
        f.seek(pos)
    except (OSError, AttributeError, ValueError):
        left = None
    if n < 0 or n > (1 << 48) or (left is not None and n > left):
        raise ValueError("truncated file: %s needs %d bytes%s" % (what, n, "" if left is None else ", %d are left" % left))
    b = f.read(n)
    if len(b) != n:
        raise ValueError("truncated file: %s needs %d bytes, got %d" % (what, n, len(b)))
    return b


@dataclass
class Crs:
    """bellman `Crs<Bn256, CrsForMonomialForm | CrsForLagrangeForm>`: g1 bases + the two G2 elements (kept raw)."""
    g1_bases: np.ndarray            # (n, 8) uint64 canonical LE affine
    g2_raw: bytes = b""             # n_g2 * 128 bytes exactly as in the file (verify-side only)
    form: str = "monomial"
    token: int = field(default_factory=lambda: next(_crs_tokens), compare=False, repr=False)  # identity of this key

    @property
    def size(self):
        return self.g1_bases.shape[0]

    @staticmethod
    def read(f) -> "Crs":
        n = struct.unpack(">Q", _read_exact(f, 8, "G1 count"))[0]
        g1 = g1_from_bytes(_read_exact(f, 64 * n, "G1 bases"), n)
        n2 = struct.unpack(">Q", _read_exact(f, 8, "G2 count"))[0]
        if n2 != 2:
            raise ValueError("a Crs file carries exactly two G2 elements, this one declares %d" % n2)
        g2 = _read_exact(f, 128 * n2, "G2 bases")
        return Crs(g1, g2)

    def write(self, f):
        f.write(struct.pack(">Q", self.size))
        f.write(g1_to_bytes(self.g1_bases))
        f.write(struct.pack(">Q", len(self.g2_raw) // 128))
        f.write(self.g2_raw)


def load_key_monomial_form(filename: str) -> Crs:  # src/reader.rs:74-77
    with open(filename, "rb") as f:
        return Crs.read(f)


def maybe_load_key_lagrange_form(option_filename: Optional[str]) -> Optional[Crs]:  # src/reader.rs:80-89
    if option_filename is None:
        return None
    with open(option_filename, "rb") as f:
        crs = Crs.read(f)
    crs.form = "lagrange"
    return crs


# ---------------------------------------------------------------- Proof (App. B.2)
@dataclass
class Proof:
    """bellman `Proof<Bn256, PlonkCsWidth4WithNextStepParams>`; field order = contrib/template.sol:330-344."""
    n: int
    num_inputs: int
    input_values: List[int]
    wire_commitments: np.ndarray               # (4, 8)
    grand_product_commitment: np.ndarray       # (8,)
    quotient_poly_commitments: np.ndarray      # (4, 8)
    wire_values_at_z: List[int]
    wire_values_at_z_omega: List[int]
    grand_product_at_z_omega: int
    quotient_polynomial_at_z: int
    linearization_polynomial_at_z: int
    permutation_polynomials_at_z: List[int]
    opening_at_z_proof: np.ndarray             # (8,)
    opening_at_z_omega_proof: np.ndarray       # (8,)
    # proofs of the two-gate-type prover (plonkit_b200.recursive; layout unpinned) also open the gate selectors at z
    gate_selectors_at_z: Optional[List[int]] = None

    def write(self, f):
        def frs(vals):
            return limbs_to_be_bytes(ints_to_limbs(vals))
        f.write(struct.pack(">QQ", self.n, self.num_inputs))
        f.write(frs(self.input_values))
        f.write(struct.pack(">Q", 4) + g1_to_bytes(self.wire_commitments))
        f.write(g1_to_bytes(self.grand_product_commitment))
        f.write(struct.pack(">Q", 4) + g1_to_bytes(self.quotient_poly_commitments))
        f.write(struct.pack(">Q", 4) + frs(self.wire_values_at_z))
        f.write(struct.pack(">Q", 1) + frs(self.wire_values_at_z_omega))
        f.write(frs([self.grand_product_at_z_omega, self.quotient_polynomial_at_z, self.linearization_polynomial_at_z]))
        f.write(struct.pack(">Q", 3) + frs(self.permutation_polynomials_at_z))
        if self.gate_selectors_at_z is not None:
            f.write(struct.pack(">Q", len(self.gate_selectors_at_z)) + frs(self.gate_selectors_at_z))
        f.write(g1_to_bytes(self.opening_at_z_proof))
        f.write(g1_to_bytes(self.opening_at_z_omega_proof))

    def to_bytes(self) -> bytes:
        import io
        b = io.BytesIO()
        self.write(b)
        return b.getvalue()

    @staticmethod
    def read(f, gated: bool = False) -> "Proof":
        """bellman `Proof::read`; a malformed file is a ValueError (short reads, counts other than the protocol's, a number of
        inputs no file could hold), never a partial object."""
        def u64():
            return struct.unpack(">Q", _read_exact(f, 8, "a count of the proof"))[0]

        def count(want, what):
            got = u64()
            if got != want:
                raise ValueError("proof file: %d %s, expected %d" % (got, what, want))

        def frs(k):
            return limbs_to_ints(be_bytes_to_limbs(_read_exact(f, 32 * k, "field elements of the proof"), k))

        def g1s(k):
            return g1_from_bytes(_read_exact(f, 64 * k, "points of the proof"), k)
        n, ni = u64(), u64()
        if ni > 1 << 28:
            raise ValueError("proof file: %d public inputs" % ni)
        inputs = frs(ni)
        count(4, "wire commitments")
        wc = g1s(4)
        gp = g1s(1)[0]
        count(4, "quotient commitments")
        qc = g1s(4)
        count(4, "wire values at z")
        wz = frs(4)
        count(1, "wire values at z omega")
        wzo = frs(1)
        gpz, tz, rz = frs(3)
        count(3, "permutation polynomials at z")
        pz = frs(3)
        gsel = None
        if gated:
            count(2, "gate selectors at z")
            gsel = frs(2)
        o1, o2 = g1s(1)[0], g1s(1)[0]
        return Proof(n, ni, inputs, wc, gp, qc, wz, wzo, gpz, tz, rz, pz, o1, o2, gsel)


def load_proof(filename: str) -> Proof:  # src/reader.rs:23-25
    with open(filename, "rb") as f:
        return Proof.read(f)


def load_proofs_from_list(list_file: str) -> List[Proof]:  # src/reader.rs:27-47
    """one proof file name per line (the input of `recursive-prove`); all proofs must have the same number of inputs"""
    with open(list_file) as f:
        names = [line.rstrip("\n") for line in f]
    proofs = [load_proof(name) for name in names]
    if not proofs:
        raise ValueError("no proof file found!")
    if any(p.num_inputs != proofs[0].num_inputs for p in proofs):
        raise ValueError("proofs num_inputs mismatch!")
    return proofs


# ---------------------------------------------------------------- VerificationKey (App. B.3)
@dataclass
class VerificationKey:
    n: int
    num_inputs: int
    selector_commitments: np.ndarray             # (6, 8): q_a,q_b,q_c,q_d,q_m,q_const
    next_step_selector_commitments: np.ndarray   # (1, 8): q_dnext
    permutation_commitments: np.ndarray          # (4, 8)
    non_residues: List[int] = field(default_factory=lambda: [5, 7, 10])
    g2_raw: bytes = b""                          # 2 x 128 B (G2 generator, [tau] G2) as in the SRS file

    def write(self, f):
        f.write(struct.pack(">QQ", self.n, self.num_inputs))
        f.write(struct.pack(">Q", 6) + g1_to_bytes(self.selector_commitments))
        f.write(struct.pack(">Q", 1) + g1_to_bytes(self.next_step_selector_commitments))
        f.write(struct.pack(">Q", 4) + g1_to_bytes(self.permutation_commitments))
        f.write(struct.pack(">Q", 3) + limbs_to_be_bytes(ints_to_limbs(self.non_residues)))
        f.write(self.g2_raw)

    def to_bytes(self) -> bytes:
        import io
        b = io.BytesIO()
        self.write(b)
        return b.getvalue()

    @staticmethod
    def read(f) -> "VerificationKey":
        def u64():
            return struct.unpack(">Q", _read_exact(f, 8, "a count of the verification key"))[0]

        def count(want, what):
            got = u64()
            if got != want:
                raise ValueError("verification key file: %d %s, expected %d" % (got, what, want))
        n, ni = u64(), u64()
        count(6, "selector commitments")
        sel = g1_from_bytes(_read_exact(f, 64 * 6, "selector commitments"), 6)
        count(1, "next-step selector commitments")
        nxt = g1_from_bytes(_read_exact(f, 64, "next-step selector commitment"), 1)
        count(4, "permutation commitments")
        perm = g1_from_bytes(_read_exact(f, 64 * 4, "permutation commitments"), 4)
        count(3, "non-residues")
        nr = limbs_to_ints(be_bytes_to_limbs(_read_exact(f, 96, "non-residues"), 3))
        g2 = _read_exact(f, 256, "G2 elements of the verification key")
        return VerificationKey(n, ni, sel, nxt, perm, nr, g2)


def load_verification_key(filename: str) -> VerificationKey:  # src/reader.rs:56-59
    with open(filename, "rb") as f:
        return VerificationKey.read(f)


# ---------------------------------------------------------------- witness (App. B.4)
def load_witness_from_file(filename: str) -> List[int]:  # src/reader.rs:92-98
    if filename.endswith("json"):
        with open(filename) as f:
            return [int(x) % R_MOD for x in json.load(f)]
    with open(filename, "rb") as f:
        return load_witness_from_array(f.read())


def load_witness_from_json_file(filename: str) -> List[int]:  # src/reader.rs:101-104
    with open(filename) as f:
        return [int(x) % R_MOD for x in json.load(f)]


def load_witness_from_bin_file(filename: str) -> List[int]:  # src/reader.rs:113-116
    with open(filename, "rb") as f:
        return load_witness_from_array(f.read())


def load_witness_limbs(filename: str) -> np.ndarray:
    """load_witness_from_file as a (len, 4) uint64 array of canonical little-endian limbs: for a `.wtns` file the 32-byte
    elements are taken as they lie in the file, no Python integer per value (what `WitnessPlan.assign` and the device consume)"""
    if filename.endswith("json"):
        return ints_to_limbs(load_witness_from_file(filename))
    with open(filename, "rb") as f:
        buffer = f.read()
    off, witness_len = _wtns_body(buffer)
    limbs = np.frombuffer(buffer, dtype="<u8", count=witness_len * 4, offset=off).reshape(witness_len, 4).astype(np.uint64)
    # value < r, limb by limb from the top
    r = ints_to_limbs([R_MOD])[0]
    less = np.zeros(witness_len, dtype=bool)
    equal = np.ones(witness_len, dtype=bool)
    for i in (3, 2, 1, 0):
        less |= equal & (limbs[:, i] < r[i])
        equal &= limbs[:, i] == r[i]
    if not less.all():
        raise ValueError("witness element is not in the field")
    return limbs


def load_witness_from_array(buffer: bytes) -> List[int]:  # src/reader.rs:119-175
    off, witness_len = _wtns_body(buffer)
    vals = [int.from_bytes(buffer[off + 32 * i:off + 32 * i + 32], "little") for i in range(witness_len)]
    if any(v >= R_MOD for v in vals):
        raise ValueError("witness element is not in the field")
    return vals


def _wtns_body(buffer: bytes):
    """header checks of src/reader.rs:119-175 -> (offset of the first element, number of elements)"""
    try:
        return _wtns_body_checked(buffer)
    except struct.error:
        raise ValueError("witness file is truncated") from None


def _wtns_body_checked(buffer: bytes):
    if buffer[:4] != b"wtns":
        raise ValueError("invalid file header")
    version, num_sections = struct.unpack_from("<II", buffer, 4)
    if version > 2:
        raise ValueError("unsupported file version")
    if num_sections != 2:
        raise ValueError("invalid num sections")
    off = 12
    sec_type, sec_size = struct.unpack_from("<IQ", buffer, off)
    off += 12
    if sec_type != 1:
        raise ValueError("invalid section type")
    if sec_size != 4 + 32 + 4:
        raise ValueError("invalid section len")
    field_size = struct.unpack_from("<I", buffer, off)[0]
    off += 4
    if field_size != 32:
        raise ValueError("invalid field byte size")
    if buffer[off:off + 32] != _PRIME_LE:
        raise ValueError("invalid curve prime")
    off += 32
    witness_len = struct.unpack_from("<I", buffer, off)[0]
    off += 4
    sec_type, sec_size = struct.unpack_from("<IQ", buffer, off)
    off += 12
    if sec_type != 2:
        raise ValueError("invalid section type")
    if sec_size != witness_len * field_size:
        raise ValueError("invalid witness section size %d" % sec_size)
    if len(buffer) < off + sec_size:
        raise ValueError("witness file is truncated")
    return off, witness_len


# ---------------------------------------------------------------- R1CS (App. B.5 / B.6)
def load_r1cs(filename: str) -> R1CS:  # src/reader.rs:178-185
    if filename.endswith("json"):
        with open(filename) as f:
            return load_r1cs_from_json(json.load(f))
    with open(filename, "rb") as f:
        r1cs, _wire_mapping = load_r1cs_from_bin(f.read())
    return r1cs


def load_r1cs_from_json(cj: dict) -> R1CS:  # src/reader.rs:194-218
    num_inputs = cj["nPubInputs"] + cj["nOutputs"] + 1
    num_aux = cj["nVars"] - num_inputs

    def conv(lc):
        # BTreeMap<String, String>: iteration order = lexicographic order of the index strings
        return [(int(k), int(v) % R_MOD) for k, v in sorted(lc.items())]
    constraints = [(conv(c[0]), conv(c[1]), conv(c[2])) for c in cj["constraints"]]
    return R1CS(num_inputs, num_aux, cj["nVars"], constraints)


def load_r1cs_from_bin(buf: bytes):  # src/r1cs_file.rs:100-154 + src/reader.rs:227-241
    """-> (R1CS, wire_mapping).  Parsed by the host library (csrc/host/transpile.cpp: the constraints stay arrays, no Python
    object per term) when it is built, else by `_load_r1cs_from_bin_py` — same checks, same result."""
    from . import circuit as _circuit
    lib = _circuit.host_library() if _circuit.NATIVE[0] else None
    if lib is None:
        return _load_r1cs_from_bin_py(buf)
    import ctypes
    h = ctypes.c_void_p()
    if lib.ph_r1cs_parse_bin(buf, ctypes.c_uint64(len(buf)), ctypes.byref(h)):
        raise ValueError(lib.ph_last_error().decode())
    handle = _circuit._NativeHandle(h, lib.ph_r1cs_free)
    hdr = (ctypes.c_uint64 * 6)()
    lib.ph_r1cs_header(handle.ptr, hdr)
    num_inputs, num_aux, num_variables, nc, nt, nmap = (int(x) for x in hdr)
    off = np.zeros(3 * nc + 1, dtype=np.uint64)
    var = np.zeros(nt, dtype=np.uint32)
    coef = np.zeros((nt, 4), dtype=np.uint64)
    wmap = np.zeros(nmap, dtype=np.uint64)
    vp = ctypes.c_void_p
    lib.ph_r1cs_export(handle.ptr, off.ctypes.data_as(vp), var.ctypes.data_as(vp), coef.ctypes.data_as(vp), wmap.ctypes.data_as(vp))
    r1cs = R1CS(num_inputs, num_aux, num_variables, None, csr=(off, var, coef))
    r1cs._native = handle
    return r1cs, wmap.tolist()


def _load_r1cs_from_bin_py(buf: bytes):
    try:
        return _load_r1cs_from_bin_checked(buf)
    except (struct.error, KeyError):
        raise ValueError("r1cs file is truncated or lacks a section") from None


def _load_r1cs_from_bin_checked(buf: bytes):
    if buf[:4] != b"r1cs":
        raise ValueError("Invalid magic number")
    version, num_sections = struct.unpack_from("<II", buf, 4)
    if version != 1:
        raise ValueError("Unsupported version")
    off = 12
    sections = {}
    for _ in range(num_sections):
        st, ss = struct.unpack_from("<IQ", buf, off)
        off += 12
        sections[st] = (off, ss)
        off += ss
    h_off, h_size = sections[1]
    field_size = struct.unpack_from("<I", buf, h_off)[0]
    if h_size != 32 + field_size:
        raise ValueError("Invalid header section size")
    if field_size != 32:
        raise ValueError("This parser only supports 32-byte fields")
    if buf[h_off + 4:h_off + 36] != _PRIME_LE:
        raise ValueError("This parser only supports bn256")
    n_wires, n_pub_out, n_pub_in, n_prv_in, n_labels, n_constraints = struct.unpack_from("<IIIIQI", buf, h_off + 36)
    c_off, _ = sections[2]
    constraints = []
    p = c_off

    def read_vec(p):
        k = struct.unpack_from("<I", buf, p)[0]
        p += 4
        out = []
        for _ in range(k):
            w = struct.unpack_from("<I", buf, p)[0]
            v = int.from_bytes(buf[p + 4:p + 36], "little")
            if v >= R_MOD:
                raise ValueError("coefficient is not in the field")
            out.append((w, v))
            p += 36
        return out, p
    for _ in range(n_constraints):
        a, p = read_vec(p)
        b, p = read_vec(p)
        c, p = read_vec(p)
        constraints.append((a, b, c))
    m_off, m_size = sections[3]
    if m_size != n_wires * 8:
        raise ValueError("Invalid map section size")
    wire_mapping = list(struct.unpack_from("<%dQ" % n_wires, buf, m_off))
    if wire_mapping and wire_mapping[0] != 0:
        raise ValueError("Wire 0 should always be mapped to 0")
    num_inputs = 1 + n_pub_in + n_pub_out
    if num_inputs > n_wires:
        raise ValueError("more public inputs and outputs than wires")
    return R1CS(num_inputs, n_wires - num_inputs, n_wires, constraints), wire_mapping
