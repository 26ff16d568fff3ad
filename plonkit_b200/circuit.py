"""Host-side circuit model: R1CS / CircomCircuit and their synthesis into width-4 PLONK gate tables.

Mirrors (names, argument meaning) the reference's host code for the step in front of the prove path:
  * `R1CS`, `CircomCircuit`, `get_public_inputs`      — src/circom_circuit.rs:32-72
  * variable allocation, aux_offset, skipped `0*LC=0` — src/circom_circuit.rs:75-133
  * `transpile_with_gates_count`, `ConstraintStat`    — src/transpile.rs:18-21,92-107,127-139
  * `analyse` / `AnalyseResult`                       — src/plonk.rs:57-95

The R1CS -> width-4 transpilation itself lives in bellman_ce's adaptor (not in the reference tree).  Only the
shapes pinned by the reference's golden vectors (src/tests.rs:14, test/circuits/simple/*) are accepted in
strict mode; see SURVEY.md App. A.2.  Everything here is serial host bookkeeping (as in the reference); the
resulting `Assembly` (gate tables) is what the CUDA prover consumes.
"""
import json
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .bn254 import R_MOD, ints_to_limbs, limbs_to_ints

AUX_OFFSET = 1  # src/plonk.rs:24
SELECTOR_NAMES = ("q_a", "q_b", "q_c", "q_d", "q_m", "q_const", "q_dnext")

LC = List[Tuple[int, int]]  # [(wire index, coefficient)]


class R1CS:  # src/circom_circuit.rs:32-38
    """constraints: [(A, B, C)], every side a list of (wire, coefficient).  A circuit parsed from a `.r1cs` file by the host
    library keeps them as arrays (`csr()`) and builds the Python list only if somebody asks for it."""

    def __init__(self, num_inputs: int, num_aux: int, num_variables: int, constraints=None, csr=None):
        self.num_inputs, self.num_aux, self.num_variables = num_inputs, num_aux, num_variables
        self._constraints = constraints if constraints is not None or csr is not None else []
        self._csr = csr          # (lc_off uint64[3 nc + 1], lc_var uint32[nt], lc_coef uint64[nt, 4] canonical)
        self._native = None      # the host library's copy (made on first use)

    @property
    def num_constraints(self) -> int:
        return len(self._constraints) if self._constraints is not None else (len(self._csr[0]) - 1) // 3

    @property
    def constraints(self):
        if self._constraints is None:
            off, var, coef = self._csr
            off, var, ints = off.tolist(), var.tolist(), limbs_to_ints(coef)
            self._constraints = [tuple(list(zip(var[off[3 * ci + k]:off[3 * ci + k + 1]], ints[off[3 * ci + k]:off[3 * ci + k + 1]]))
                                       for k in range(3)) for ci in range((len(off) - 1) // 3)]
        return self._constraints

    def csr(self):
        if self._csr is None:
            off, var, coef = [0], [], []
            for sides in self._constraints:
                for lc in sides:
                    for v, c in lc:
                        var.append(v)
                        coef.append(c % R_MOD)
                    off.append(len(var))
            self._csr = (np.asarray(off, dtype=np.uint64), np.asarray(var, dtype=np.uint32),
                         ints_to_limbs(coef) if coef else np.zeros((0, 4), dtype=np.uint64))
        return self._csr

    def __eq__(self, other):
        return isinstance(other, R1CS) and (self.num_inputs, self.num_aux, self.num_variables) == \
            (other.num_inputs, other.num_aux, other.num_variables) and self.constraints == other.constraints

    def __repr__(self):
        return "R1CS(num_inputs=%d, num_aux=%d, num_variables=%d, %d constraints)" % (
            self.num_inputs, self.num_aux, self.num_variables, self.num_constraints)


class _NativeHandle:
    """an object of the host library, freed with it"""

    def __init__(self, ptr, free):
        self.ptr, self._free = ptr, free

    def __del__(self):
        try:
            if self.ptr:
                self._free(self.ptr)
                self.ptr = None
        except Exception:
            pass


@dataclass
class CircomCircuit:  # src/circom_circuit.rs:40-47
    r1cs: R1CS
    witness: Optional[List[int]] = None
    wire_mapping: Optional[List[int]] = None
    aux_offset: int = AUX_OFFSET
    strict: bool = True   # False: also transpile constraint shapes no reference fixture pins (see _transpile)

    def _w(self, i):
        j = i if self.wire_mapping is None else self.wire_mapping[i]
        if isinstance(self.witness, np.ndarray):      # (len, 4) uint64 canonical limbs (reader.load_witness_limbs)
            return limbs_to_ints(self.witness[j:j + 1])[0]
        return self.witness[j]

    def witness_ints(self):
        """the witness as Python integers whatever form it is held in"""
        if isinstance(self.witness, np.ndarray):
            return limbs_to_ints(self.witness)
        return self.witness

    def get_public_inputs(self):  # src/circom_circuit.rs:50-59
        if self.witness is None:
            return None
        return [self._w(i) for i in range(1, self.r1cs.num_inputs)]

    def get_public_inputs_json(self) -> str:  # src/circom_circuit.rs:61-68
        inputs = self.get_public_inputs()
        if inputs is None:
            return "[]"
        return json.dumps([str(x) for x in inputs], indent=2)


@dataclass
class ConstraintStat:  # src/transpile.rs:11-16
    name: str
    num_gates: int


@dataclass
class Assembly:
    """Width-4 gate tables over the domain of size n (a power of two).

    Rows 0..num_inputs-1 are the public-input gates (a = input variable, q_a = -1), then the circuit's gates,
    then dummy padding; row n-1 is never a gate (bellman pads to n-1 gates).  Variable 0 is bellman's dummy
    variable (value 0).  SURVEY.md App. A.2.
    """
    n: int
    num_inputs: int
    wire_idx: np.ndarray                 # (4, n) uint32 variable ids
    selectors: np.ndarray                # (7, n, 4) uint64 canonical LE limbs, order SELECTOR_NAMES
    var_values: Optional[np.ndarray]     # (nvars, 4) uint64 canonical LE limbs, or None (setup only)
    nvars: int
    num_gates: int = 0                   # gates before padding (incl. input gates)
    plan: Optional["WitnessPlan"] = None  # how var_values follow from a circom witness (set by synthesize)

    def public_inputs(self):
        from .bn254 import limbs_to_ints
        if self.var_values is None:
            return None
        return limbs_to_ints(self.var_values[self.wire_idx[0, :self.num_inputs]])


class UnpinnedTranspilation(NotImplementedError):
    """Raised for R1CS constraint shapes whose width-4 layout no reference fixture pins."""


def _norm_lc(lc: LC):
    """-> (sorted [(var, coeff)], constant); wire 0 is the constant ONE (src/circom_circuit.rs:107-113)"""
    acc = {}
    const = 0
    for idx, coeff in lc:
        coeff %= R_MOD
        if idx == 0:
            const = (const + coeff) % R_MOD
        else:
            acc[idx] = (acc.get(idx, 0) + coeff) % R_MOD
    terms = sorted((i, c) for i, c in acc.items() if c != 0)
    return terms, const


@dataclass
class _Gates:
    rows: list = field(default_factory=list)   # (a, b, c, d, [7 selector ints])
    values: list = field(default_factory=list)  # variable values (None when no witness)
    program: list = field(default_factory=list)  # per new variable, in order: (const, [(var, coeff)]) — its value as a combination of earlier ones
    hints: int = 0
    stats: list = field(default_factory=list)
    # the compiled transpiler hands the same content over as arrays instead of rows / values / program / stats
    tables: Optional[tuple] = None     # (n, wire_idx (4, n), selectors (7, n, 4)) already padded to the domain
    var_values: Optional[np.ndarray] = None
    plan: Optional["WitnessPlan"] = None
    stat_arrays: Optional[tuple] = None
    native_rows: int = 0

    @property
    def n_rows(self) -> int:
        return self.native_rows if self.tables is not None else len(self.rows)

    @property
    def n_stats(self) -> int:
        return len(self.stat_arrays[0]) if self.stat_arrays is not None else len(self.stats)

    def stat_list(self):
        if self.stat_arrays is not None:
            return [ConstraintStat(str(int(c)), int(k)) for c, k in zip(*self.stat_arrays)]
        return self.stats


NATIVE = [True]   # tests set NATIVE[0] = False to run the Python statement of the transpiler / parser instead of the compiled one


def _transpile(circuit: CircomCircuit, strict: bool = True) -> _Gates:
    """R1CS constraints -> width-4 gates: the host library's compiled transpiler (csrc/host/transpile.cpp) when it is built,
    else `_transpile_py`, the readable statement of the same layout (tests hold the two equal)."""
    r = circuit.r1cs
    _, wires, _ = r.csr()
    if len(wires) and int(wires.max()) >= r.num_variables:
        raise ValueError("a constraint refers to wire %d but the circuit has %d variables" % (int(wires.max()), r.num_variables))
    lib = host_library() if NATIVE[0] else None
    if lib is None:
        return _transpile_py(circuit, strict)
    import ctypes
    native = r._native           # held locally: ranks of a sharded prover transpile the same R1CS from several threads
    if native is None:
        off, var, coef = r.csr()
        h = ctypes.c_void_p()
        vp = ctypes.c_void_p
        if lib.ph_r1cs_from_csr(ctypes.c_uint64(r.num_inputs), ctypes.c_uint64(r.num_aux), ctypes.c_uint64(r.num_variables),
                                ctypes.c_uint64(r.num_constraints), off.ctypes.data_as(vp), var.ctypes.data_as(vp),
                                coef.ctypes.data_as(vp), ctypes.byref(h)):
            raise ValueError(lib.ph_last_error().decode())
        native = r._native = _NativeHandle(h, lib.ph_r1cs_free)
    detail = (ctypes.c_uint64 * 7)()
    out = ctypes.c_void_p()
    rc = lib.ph_transpile(native.ptr, int(bool(strict)), ctypes.byref(out), detail)
    if rc == 1:
        raise UnpinnedTranspilation("constraint %d: A and B must be single-variable terms" % detail[1])
    if rc == 2:
        raise UnpinnedTranspilation("constraint %d: C side with %d variables is not pinned by any fixture" % (detail[1], detail[2]))
    if rc == 3:
        raise ValueError("constraint %d is the contradiction %d = 0" % (detail[1], sum(int(detail[3 + i]) << (64 * i) for i in range(4))))
    if rc:
        raise ValueError(lib.ph_last_error().decode() or "transpilation failed (%d)" % rc)
    gates = _NativeHandle(out, lib.ph_gates_free)
    hdr = (ctypes.c_uint64 * 6)()
    lib.ph_gates_header(gates.ptr, hdr)
    rows, nvars, ndirect, hints, nstats, nterms = (int(x) for x in hdr)
    n = 1
    while n < rows + 1:
        n *= 2
    wire_idx = np.zeros((4, n), dtype=np.uint32)
    selectors = np.zeros((7, n, 4), dtype=np.uint64)
    prog_off = np.zeros(nvars - ndirect + 1, dtype=np.uint64)
    prog_var = np.zeros(nterms, dtype=np.uint32)
    prog_coef = np.zeros((nterms, 4), dtype=np.uint64)
    prog_const = np.zeros((nvars - ndirect, 4), dtype=np.uint64)
    stat_c, stat_g = np.zeros(nstats, dtype=np.uint32), np.zeros(nstats, dtype=np.uint32)
    vp = ctypes.c_void_p
    lib.ph_gates_export(gates.ptr, ctypes.c_uint64(n), wire_idx.ctypes.data_as(vp), selectors.ctypes.data_as(vp),
                        prog_off.ctypes.data_as(vp), prog_var.ctypes.data_as(vp), prog_coef.ctypes.data_as(vp),
                        prog_const.ctypes.data_as(vp), stat_c.ctypes.data_as(vp), stat_g.ctypes.data_as(vp))
    g = _Gates()
    g.tables, g.native_rows, g.hints, g.stat_arrays = (n, wire_idx, selectors), rows, hints, (stat_c, stat_g)
    g.plan = WitnessPlan.from_arrays(ndirect, prog_off, prog_var, prog_coef, prog_const)
    if circuit.witness is not None:
        g.var_values = g.plan.assign(circuit.witness, circuit.wire_mapping)
    return g


def _transpile_py(circuit: CircomCircuit, strict: bool = True) -> _Gates:
    """R1CS constraints -> width-4 gates.

    strict = True accepts only the constraint shapes whose gate layout the reference's golden vectors pin
    (single-variable A and B, C = one variable or two variables + constant; src/tests.rs:14) and raises
    `UnpinnedTranspilation` otherwise.  strict = False also transpiles arbitrary linear combinations — BYTE PARITY
    UNPINNED: bellman's adaptor (src/transpile.rs:92-139 is only its wrapper) is not in the reference tree and no
    fixture shows its layout for long combinations, so this is this repository's own sound layout: a combination of up
    to three variables is collapsed into a fresh variable by one gate, a longer one by a running sum carried through the
    d wire with q_dnext = -1; pinned shapes come out exactly as in strict mode."""
    r = circuit.r1cs
    have_w = circuit.witness is not None
    g = _Gates()
    # variable id == witness index: Input(i) -> i ; Aux(j + aux_offset) -> num_inputs + j  (circom_circuit.rs:75-105)
    if have_w:
        wit = circuit.witness_ints()
        wm = circuit.wire_mapping
        g.values = [0] + [(wit[i] if wm is None else wit[wm[i]]) % R_MOD for i in range(1, r.num_variables)]
    else:
        g.values = [None] * r.num_variables
    # public-input gates first
    for i in range(1, r.num_inputs):
        g.rows.append((i, 0, 0, 0, [R_MOD - 1, 0, 0, 0, 0, 0, 0]))
    M1 = R_MOD - 1

    def new_var(terms, const):
        """a fresh variable whose value is sum(terms) + const: recorded in g.program (the per-proof witness assignment replays
        it, WitnessPlan) and evaluated right away when a witness is at hand"""
        g.program.append((const, list(terms)))
        g.values.append((sum(c * g.values[v] for v, c in terms) + const) % R_MOD if have_w else None)
        return len(g.values) - 1

    def chain(terms, const, out):
        """Gates enforcing sum(terms) + const - out = 0 (out = None: the sum itself is zero).  More than four summands
        run as a chain: every row adds up to three terms to the running sum in d and hands it to the next row's d."""
        items = list(terms) + ([(out, M1)] if out is not None else [])
        if len(items) <= 4:
            wires = [v for v, _ in items] + [0] * (4 - len(items))
            coefs = [c for _, c in items] + [0] * (4 - len(items))
            g.rows.append((wires[0], wires[1], wires[2], wires[3], [coefs[0], coefs[1], coefs[2], coefs[3], 0, const, 0]))
            return
        first, rest = items[:4], items[4:]
        acc = new_var(first, const)
        g.rows.append((first[0][0], first[1][0], first[2][0], first[3][0],
                       [first[0][1], first[1][1], first[2][1], first[3][1], 0, const, M1]))
        while rest:
            take, rest = rest[:3], rest[3:]
            wires = [v for v, _ in take] + [0] * (3 - len(take))
            coefs = [c for _, c in take] + [0] * (3 - len(take))
            if rest:
                nxt = new_var([(acc, 1)] + list(take), 0)
                g.rows.append((wires[0], wires[1], wires[2], acc, [coefs[0], coefs[1], coefs[2], 1, 0, 0, M1]))
                acc = nxt
            else:
                g.rows.append((wires[0], wires[1], wires[2], acc, [coefs[0], coefs[1], coefs[2], 1, 0, 0, 0]))

    def collapse(terms, const):
        """-> (variable, coefficient) with  LC == coefficient * variable; a fresh variable (and its gates) unless the
        combination already is a single variable"""
        if len(terms) == 1 and const == 0:
            return terms[0]
        t = new_var(terms, const)
        if len(terms) == 2:   # the pinned layout: (a = v1, b = v2, c = t), q_c = -1
            (v1, c1), (v2, c2) = terms
            g.rows.append((v1, v2, t, 0, [c1, c2, M1, 0, 0, const, 0]))
        else:
            chain(terms, const, t)
        return t, 1

    for ci, (A, B, C) in enumerate(r.constraints):
        if (len(A) == 0 or len(B) == 0) and len(C) == 0:  # 0 * LC = 0 is ignored (circom_circuit.rs:122-123)
            continue
        before = len(g.rows)
        (ta, ka), (tb, kb), (tc, kc) = _norm_lc(A), _norm_lc(B), _norm_lc(C)
        pinned = len(ta) == 1 and ka == 0 and len(tb) == 1 and kb == 0 and ((len(tc) == 1 and kc == 0) or len(tc) == 2)
        if not pinned and strict:
            if not (len(ta) == 1 and ka == 0 and len(tb) == 1 and kb == 0):
                raise UnpinnedTranspilation("constraint %d: A and B must be single-variable terms" % ci)
            raise UnpinnedTranspilation("constraint %d: C side with %d variables is not pinned by any fixture" % (ci, len(tc)))
        if not ta or not tb:
            # a constant factor: k * LC_other - LC_C = 0 is linear
            k, (to, ko) = (ka, (tb, kb)) if not ta else (kb, (ta, ka))
            lin = {}
            for v, c in to:
                lin[v] = (lin.get(v, 0) + k * c) % R_MOD
            for v, c in tc:
                lin[v] = (lin.get(v, 0) - c) % R_MOD
            terms = sorted((v, c) for v, c in lin.items() if c)
            const = (k * ko - kc) % R_MOD
            if not terms:
                if const:
                    raise ValueError("constraint %d is the contradiction %d = 0" % (ci, const))
            else:
                chain(terms, const, None)
        else:
            (x, alpha), (y, beta) = collapse(ta, ka), collapse(tb, kb)
            qm = alpha * beta % R_MOD
            if not tc:
                g.rows.append((x, y, 0, 0, [0, 0, 0, 0, qm, (R_MOD - kc) % R_MOD, 0]))
            else:
                z, gamma = collapse(tc, kc)
                g.rows.append((x, y, z, 0, [0, 0, (R_MOD - gamma) % R_MOD, 0, qm, 0, 0]))
        g.hints += 1
        g.stats.append(ConstraintStat(str(ci), len(g.rows) - before))
    return g


class WitnessPlan:
    """The per-proof half of synthesis: var_values of an `Assembly` from a circom witness, without transpiling again.

    bellman re-synthesises the transpiled circuit inside every `prove` (src/plonk.rs:132-176) and evaluates each variable's
    value there; the gate tables do not depend on the witness, so here `_transpile` runs once (prepare_setup_for_prover) and
    leaves (i) the direct variables — variable i is witness[wire_mapping[i]], variable 0 the dummy 0 — and (ii) a
    straight-line program for the variables it introduced: value = const + sum coeff * earlier value.  `assign` replays
    it in the native host library (csrc/host/witness.cpp) or, if that is not built, with Python integers — same values."""

    def __init__(self, num_direct: int, program):
        self.num_direct = num_direct
        self.num_new = len(program)
        off = np.zeros(self.num_new + 1, dtype=np.uint64)
        distinct, tv, tc, kc = {}, [], [], []      # few distinct coefficients: each is converted to limbs once
        for k, (const, terms) in enumerate(program):
            for v, c in terms:
                if not 0 <= v < num_direct + k:
                    raise ValueError("witness program entry %d reads variable %d before it is assigned" % (k, v))
                tv.append(v)
                tc.append(distinct.setdefault(c % R_MOD, len(distinct)))
            kc.append(distinct.setdefault(const % R_MOD, len(distinct)))
            off[k + 1] = len(tv)
        table = ints_to_limbs(list(distinct)) if distinct else np.zeros((0, 4), dtype=np.uint64)
        self.off = off
        self.term_var = np.asarray(tv, dtype=np.uint32)
        self.term_coef = np.ascontiguousarray(table[np.asarray(tc, dtype=np.int64)]) if tc else np.zeros((0, 4), dtype=np.uint64)
        self.consts = np.ascontiguousarray(table[np.asarray(kc, dtype=np.int64)]) if kc else np.zeros((0, 4), dtype=np.uint64)
        self._coef_mont = None      # the coefficients in Montgomery form, made once by the host library

    @classmethod
    def from_arrays(cls, num_direct, off, term_var, term_coef, consts):
        self = cls.__new__(cls)
        self.num_direct, self.num_new = int(num_direct), len(off) - 1
        self.off, self.term_var, self.term_coef, self.consts, self._coef_mont = off, term_var, term_coef, consts, None
        return self

    @property
    def nvars(self):
        return self.num_direct + self.num_new

    def assign(self, witness, wire_mapping=None, native: Optional[bool] = None, threads: Optional[int] = None) -> np.ndarray:
        """witness: list of integers or (len, 4) uint64 canonical limbs -> (nvars, 4) uint64 canonical limbs.
        native: None = the host library when it is built, True = require it, False = Python integers."""
        if isinstance(witness, np.ndarray):
            w = np.ascontiguousarray(witness, dtype=np.uint64).reshape(-1, 4)
        else:
            w = ints_to_limbs([x % R_MOD for x in witness])
        idx = np.arange(self.num_direct) if wire_mapping is None else np.asarray(wire_mapping[:self.num_direct], dtype=np.int64)
        if self.num_direct > len(idx) or (len(idx) and int(idx.max()) >= w.shape[0]):
            raise ValueError("witness holds %d values, the circuit has %d variables" % (w.shape[0], self.num_direct))
        values = np.zeros((self.nvars, 4), dtype=np.uint64)
        values[:self.num_direct] = w[idx]
        values[0] = 0                                   # variable 0 is bellman's dummy variable, not the constant ONE
        lib = host_library() if native in (None, True) else None
        if lib is None and native:
            raise RuntimeError("libplonkit_host.so is not built (python __graft_entry__.py)")
        if lib is not None:
            import ctypes
            import os
            vp = ctypes.c_void_p
            coef_mont = self._coef_mont
            if coef_mont is None:
                # filled before it is published: provers of a pool / ranks of a sharded prover assign concurrently
                coef_mont = np.zeros_like(self.term_coef)
                lib.ph_fr_to_mont(self.term_coef.ctypes.data_as(vp), coef_mont.ctypes.data_as(vp), ctypes.c_uint64(len(self.term_coef)))
                self._coef_mont = coef_mont
            if threads is None:
                threads = min(16, os.cpu_count() or 1)
            rc = lib.ph_assign_witness(ctypes.c_uint64(self.num_direct), ctypes.c_uint64(self.num_new), self.off.ctypes.data_as(vp),
                                       self.term_var.ctypes.data_as(vp), coef_mont.ctypes.data_as(vp), 1,
                                       self.consts.ctypes.data_as(vp), values.ctypes.data_as(vp), int(threads))
            if rc:
                raise ValueError("witness assignment failed at variable %d (value not in the field, or a malformed program)" % (rc - 1))
            return values
        vals = limbs_to_ints(values[:self.num_direct])
        if any(v >= R_MOD for v in vals):
            raise ValueError("witness value not in the field")
        coef, consts = limbs_to_ints(self.term_coef), limbs_to_ints(self.consts)
        for k in range(self.num_new):
            lo, hi = int(self.off[k]), int(self.off[k + 1])
            vals.append((consts[k] + sum(coef[j] * vals[self.term_var[j]] for j in range(lo, hi))) % R_MOD)
        if self.num_new:
            values[self.num_direct:] = ints_to_limbs(vals[self.num_direct:])
        return values


_HOST_LIB = [False]


def host_library():
    """plonkit_b200/libplonkit_host.so (plain C++ host helpers, built by __graft_entry__.build()), or None if not built"""
    if _HOST_LIB[0] is False:
        import ctypes
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libplonkit_host.so")
        lib = None
        if os.path.exists(path):
            lib = ctypes.CDLL(path)
            lib.ph_assign_witness.restype = ctypes.c_int64
            lib.ph_last_error.restype = ctypes.c_char_p
        _HOST_LIB[0] = lib
    return _HOST_LIB[0]


def transpile_with_gates_count(circuit: CircomCircuit, strict: Optional[bool] = None):
    """src/transpile.rs:127-139 -> (gates_count, hints_count).  Counts exclude the public-input gates."""
    g = _transpile(circuit, circuit.strict if strict is None else strict)
    n_in = circuit.r1cs.num_inputs - 1
    return g.n_rows - n_in, g.hints


def synthesize(circuit: CircomCircuit, strict: Optional[bool] = None) -> Assembly:
    g = _transpile(circuit, circuit.strict if strict is None else strict)
    if g.tables is not None:
        n, wire_idx, selectors = g.tables
        return Assembly(n=n, num_inputs=circuit.r1cs.num_inputs - 1, wire_idx=wire_idx, selectors=selectors,
                        var_values=g.var_values, nvars=g.plan.nvars, num_gates=g.n_rows, plan=g.plan)
    asm = assembly_from_rows(g.rows, g.values, circuit.r1cs.num_inputs - 1)
    asm.plan = WitnessPlan(circuit.r1cs.num_variables, g.program)
    return asm


def assembly_from_rows(rows, values, num_inputs) -> Assembly:
    n_gates = len(rows)
    n = 1
    while n < n_gates + 1:
        n *= 2
    wire_idx = np.zeros((4, n), dtype=np.uint32)
    # selector tables are sparse and draw on few distinct coefficients: convert each distinct value to limbs once
    distinct, nz_s, nz_r, nz_v = {}, [], [], []
    for r, (a, b, c, d, q) in enumerate(rows):
        wire_idx[0, r], wire_idx[1, r], wire_idx[2, r], wire_idx[3, r] = a, b, c, d
        for s in range(7):
            v = q[s]
            if v:
                nz_s.append(s)
                nz_r.append(r)
                nz_v.append(distinct.setdefault(v, len(distinct)))
    selectors = np.zeros((7, n, 4), dtype=np.uint64)
    if distinct:
        selectors[np.asarray(nz_s), np.asarray(nz_r)] = ints_to_limbs(list(distinct))[np.asarray(nz_v)]
    var_values = None
    if values and values[-1] is not None and all(v is not None for v in values):
        var_values = ints_to_limbs(values)
    return Assembly(n=n, num_inputs=num_inputs, wire_idx=wire_idx, selectors=selectors, var_values=var_values,
                    nvars=len(values), num_gates=n_gates)


def analyse(circuit: CircomCircuit, strict: Optional[bool] = None) -> dict:
    """src/plonk.rs:72-95 (field order as serialised by the reference; src/tests.rs:14)"""
    g = _transpile(circuit, circuit.strict if strict is None else strict)
    r = circuit.r1cs
    res = {
        "num_inputs": r.num_inputs,
        "num_aux": r.num_aux,
        "num_variables": r.num_variables,
        "num_constraints": r.num_constraints,
        "num_nontrivial_constraints": g.n_stats,
        "num_gates": g.n_rows - (r.num_inputs - 1),
        "num_hints": g.hints,
    }
    if g.n_stats:
        res["constraint_stats"] = [{"name": s.name, "num_gates": s.num_gates} for s in g.stat_list()]
    return res


def is_satisfied(asm: Assembly) -> bool:
    """is_satisfied_using_one_shot_check (src/plonk.rs:137) on the gate tables — host check, small circuits."""
    from .bn254 import limbs_to_ints
    if asm.var_values is None:
        return False
    vals = limbs_to_ints(asm.var_values)
    sel = [limbs_to_ints(asm.selectors[s]) for s in range(7)]
    w = [[vals[i] for i in asm.wire_idx[c]] for c in range(4)]
    for r in range(asm.n - 1):
        pi = w[0][r] if r < asm.num_inputs else 0
        acc = (sel[0][r] * w[0][r] + sel[1][r] * w[1][r] + sel[2][r] * w[2][r] + sel[3][r] * w[3][r]
               + sel[4][r] * w[0][r] * w[1][r] + sel[5][r] + sel[6][r] * w[3][r + 1] + pi)
        if acc % R_MOD:
            return False
    return True
