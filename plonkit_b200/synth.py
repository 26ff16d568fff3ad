"""Synthetic PLONK-native circuits of the sizes BASELINE.json names.

The reference's poseidon fixtures (r1cs / witness) do not exist in-tree and circom/snarkjs are absent, so the
2^20-gate "poseidon" workload is generated directly as width-4 gate tables with the shape of circomlib's
Poseidon(2) (t = 3, x^5 S-box = 3 multiplication gates, one addition gate per MDS row, 8 full + 57 partial rounds
= 438 gates per permutation), chained until the requested number of gates is reached.  Round constants and the
MDS matrix come from splitmix64 seeded with "plonkit" (circomlib's constants are not in the tree).  One public input.
"""
import numpy as np

from .bn254 import R_MOD, ints_to_limbs
from .circuit import Assembly

SEED = 0x706C6F6E6B6974  # "plonkit"
T, FULL_ROUNDS, PARTIAL_ROUNDS = 3, 8, 57


class SplitMix64:
    def __init__(self, seed):
        self.s = seed & 0xFFFFFFFFFFFFFFFF

    def next(self):
        self.s = (self.s + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        return z ^ (z >> 31)

    def field(self):
        v = 0
        for i in range(4):
            v |= self.next() << (64 * i)
        return v % R_MOD


def random_field_elements(n, seed=SEED) -> np.ndarray:
    """(n, 4) uint64 canonical limbs, vectorised splitmix64 (element i uses draws 4i..4i+3, reduced mod r)."""
    idx = np.arange(1, 4 * n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    limbs = z.reshape(n, 4).copy()
    # reduce mod r: the top limb is masked to 253 bits first (value < 2^253 < r), which keeps this vectorised
    limbs[:, 3] &= np.uint64((1 << 61) - 1)
    return limbs


def _poseidon_pattern():
    """One permutation as (local wire ids, selector rows, evaluation program).
    Local variable ids: -3,-2,-1 = incoming state; 0.. = variables allocated by this permutation."""
    rng = SplitMix64(SEED)
    rounds = FULL_ROUNDS + PARTIAL_ROUNDS
    rc = [[rng.field() for _ in range(T)] for _ in range(rounds)]
    mds = [[rng.field() for _ in range(T)] for _ in range(T)]
    gates = []  # (a, b, c, d, [q_a,q_b,q_c,q_d,q_m,q_const,q_dnext], kind)
    state = [-3, -2, -1]
    nxt = 0
    M1 = R_MOD - 1
    for r in range(rounds):
        full = r < FULL_ROUNDS // 2 or r >= FULL_ROUNDS // 2 + PARTIAL_ROUNDS
        for i in range(T if full else 1):
            x = state[i]
            x2, x4, x5 = nxt, nxt + 1, nxt + 2
            nxt += 3
            gates.append((x, x, x2, None, [0, 0, M1, 0, 1, 0, 0], "mul"))
            gates.append((x2, x2, x4, None, [0, 0, M1, 0, 1, 0, 0], "mul"))
            gates.append((x4, x, x5, None, [0, 0, M1, 0, 1, 0, 0], "mul"))
            state[i] = x5
        new_state = []
        for i in range(T):
            d = nxt
            nxt += 1
            gates.append((state[0], state[1], state[2], d, [mds[i][0], mds[i][1], mds[i][2], M1, 0, rc[r][i], 0], "add"))
            new_state.append(d)
        state = new_state
    return gates, nxt


def poseidon_chain_assembly(log_n: int, with_witness: bool = True, inputs=(3, 4, 5)) -> Assembly:
    """Poseidon-shaped circuit with exactly 2^log_n - 1 gates (domain size n = 2^log_n), 1 public input."""
    n = 1 << log_n
    n_gates = n - 1
    pattern, vars_per_perm = _poseidon_pattern()
    G = len(pattern)
    assert vars_per_perm == G
    body = n_gates - 1  # minus the public-input gate
    perms = (body + G - 1) // G

    # ---- tables, vectorised over permutations
    loc = np.zeros((4, G), dtype=np.int64)
    is_dummy = np.zeros((4, G), dtype=bool)
    sel_rows = [[0] * G for _ in range(7)]
    for g, (a, b, c, d, q, _) in enumerate(pattern):
        for col, v in enumerate((a, b, c, d)):
            if v is None:
                is_dummy[col, g] = True
            else:
                loc[col, g] = v
        for s in range(7):
            sel_rows[s][g] = q[s]
    sel_block = np.stack([ints_to_limbs(r) for r in sel_rows])  # (7, G, 4)
    base = 4 + np.arange(perms, dtype=np.int64) * G                # first variable id of permutation p
    idx_body = (base[None, :, None] + loc[:, None, :])             # (4, perms, G)
    idx_body = np.where(is_dummy[:, None, :], 0, idx_body).reshape(4, perms * G)[:, :body]
    wire_idx = np.zeros((4, n), dtype=np.uint32)
    wire_idx[0, 0] = 1  # public-input gate: a = variable 1
    wire_idx[:, 1:1 + body] = idx_body.astype(np.uint32)
    selectors = np.zeros((7, n, 4), dtype=np.uint64)
    selectors[0, 0] = ints_to_limbs([R_MOD - 1])[0]
    selectors[:, 1:1 + body] = np.broadcast_to(sel_block[:, None, :, :], (7, perms, G, 4)).reshape(7, perms * G, 4)[:, :body]
    # every gate allocates exactly one variable; gates past the truncation point (and their variables) do not exist
    nvars = 4 + body

    var_values = None
    if with_witness:
        vals = _evaluate_chain(pattern, perms, body, inputs)
        var_values = ints_to_limbs(vals[:nvars])
    return Assembly(n=n, num_inputs=1, wire_idx=wire_idx, selectors=selectors, var_values=var_values, nvars=nvars,
                    num_gates=n_gates)


def _evaluate_chain(pattern, perms, body, inputs):
    G = len(pattern)
    vals = [0] * (4 + perms * G)
    vals[1], vals[2], vals[3] = [v % R_MOD for v in inputs]
    p_mod = R_MOD
    prog = []
    for (a, b, c, d, q, kind) in pattern:
        if kind == "mul":
            prog.append((True, a, b, c, 0, 0, 0, 0))
        else:
            prog.append((False, a, b, c, d, q[0], q[1], (q[2], q[5])))
    done = 0
    for p in range(perms):
        b0 = 4 + p * G
        for (is_mul, a, b, c, d, qa, qb, rest) in prog:
            if done == body:
                return vals
            if is_mul:
                vals[b0 + c] = vals[b0 + a] * vals[b0 + b] % p_mod
            else:
                qc, k = rest
                vals[b0 + d] = (qa * vals[b0 + a] + qb * vals[b0 + b] + qc * vals[b0 + c] + k) % p_mod
            done += 1
    return vals


def random_gate_assembly(log_n: int, seed: int = SEED, reuse: float = 0.5, num_inputs: int = 1) -> Assembly:
    """BASELINE config 3 shape: 2^log_n - 1 gates, half multiplication gates c = a*b, half addition gates c = a + b + k;
    every operand re-uses an earlier variable with probability `reuse` (non-trivial copy permutation).
    `num_inputs` public inputs (variables 1..num_inputs, one input gate each)."""
    import random
    rnd = random.Random(seed)
    n = 1 << log_n
    n_gates = n - 1
    rows = []
    vals = [0] + [rnd.randrange(R_MOD) for _ in range(num_inputs)]
    for i in range(1, num_inputs + 1):
        rows.append((i, 0, 0, 0, [R_MOD - 1, 0, 0, 0, 0, 0, 0]))

    def operand():
        if len(vals) > 2 and rnd.random() < reuse:
            return rnd.randrange(1, len(vals))
        vals.append(rnd.randrange(R_MOD))
        return len(vals) - 1
    M1 = R_MOD - 1
    while len(rows) < n_gates:
        a, b = operand(), operand()
        if rnd.random() < 0.5:
            vals.append(vals[a] * vals[b] % R_MOD)
            rows.append((a, b, len(vals) - 1, 0, [0, 0, M1, 0, 1, 0, 0]))
        else:
            k = rnd.randrange(R_MOD)
            vals.append((vals[a] + vals[b] + k) % R_MOD)
            rows.append((a, b, len(vals) - 1, 0, [1, 1, M1, 0, 0, k, 0]))
    from .circuit import assembly_from_rows
    return assembly_from_rows(rows, vals, num_inputs)


def random_gate_assembly_layered(log_n: int, seed: int = SEED, layers: int = 16, reuse: float = 0.5) -> Assembly:
    """The BASELINE config 3 shape (half multiplication gates c = a*b, half addition gates c = a + b + k, operands re-used
    with probability `reuse`) at sizes where `random_gate_assembly`'s gate-by-gate Python loop is too slow (2^22 - 2^24):
    gates are generated in `layers` layers with numpy; an operand of a gate in layer l either re-uses a variable created
    in an earlier layer or is a fresh variable.  One public input."""
    from .bn254 import limbs_to_ints
    n = 1 << log_n
    n_gates = n - 1
    body = n_gates - 1
    rng = np.random.default_rng(seed & 0xFFFFFFFF)
    per = [body // layers + (1 if i < body % layers else 0) for i in range(layers)]
    wire_idx = np.zeros((4, n), dtype=np.uint32)
    selectors = np.zeros((7, n, 4), dtype=np.uint64)
    m1 = ints_to_limbs([R_MOD - 1])[0]
    one = ints_to_limbs([1])[0]
    wire_idx[0, 0] = 1
    selectors[0, 0] = m1
    values = [0, rng.integers(1, 1 << 62).item()]  # variable 0 = dummy, variable 1 = the public input
    row = 1
    for g in per:
        nv0 = len(values)                               # variables that exist before this layer
        fresh_a = rng.random(g) >= reuse if nv0 > 2 else np.ones(g, dtype=bool)
        fresh_b = rng.random(g) >= reuse if nv0 > 2 else np.ones(g, dtype=bool)
        na, nb = int(fresh_a.sum()), int(fresh_b.sum())
        new_vals = limbs_to_ints(random_field_elements(na + nb + g, seed=seed + row))  # fresh operands, then the constants k
        a_idx = rng.integers(1, nv0, size=g, dtype=np.int64)
        b_idx = rng.integers(1, nv0, size=g, dtype=np.int64)
        a_idx[fresh_a] = nv0 + np.arange(na)
        b_idx[fresh_b] = nv0 + na + np.arange(nb)
        values.extend(new_vals[:na + nb])
        ks = new_vals[na + nb:]
        is_mul = rng.random(g) < 0.5
        av = [values[i] for i in a_idx.tolist()]
        bv = [values[i] for i in b_idx.tolist()]
        cv = [(x * y % R_MOD) if mflag else ((x + y + k) % R_MOD) for x, y, k, mflag in zip(av, bv, ks, is_mul.tolist())]
        c_idx = len(values) + np.arange(g)
        values.extend(cv)
        sl = slice(row, row + g)
        wire_idx[0, sl], wire_idx[1, sl], wire_idx[2, sl] = a_idx, b_idx, c_idx
        selectors[2, sl] = m1                            # q_c = -1
        selectors[4, sl][is_mul] = one                   # q_m = 1
        add = ~is_mul
        selectors[0, sl][add] = one
        selectors[1, sl][add] = one
        kl = ints_to_limbs(ks)
        selectors[5, sl][add] = kl[add]
        row += g
    return Assembly(n=n, num_inputs=1, wire_idx=wire_idx, selectors=selectors, var_values=ints_to_limbs(values), nvars=len(values),
                    num_gates=n_gates)


# ---------------------------------------------------------------- an R1CS of circomlib's Poseidon(2) shape (wide LCs)
def poseidon_r1cs(hashes: int = 1, inputs=(3, 4)):
    """R1CS + witness shaped like circom's optimised output for circomlib 0.5 `Poseidon(2)` (test/circuits/poseidon):
    t = 3, 8 full + 57 partial rounds, x^5 S-box = 3 constraints (x2 = x*x, x4 = x2*x2, x5 = x4*x) whose factors are
    LINEAR COMBINATIONS: the MDS mix and the round constants are folded into the next S-box input, and in the partial
    rounds the two untouched state elements stay linear, so the combinations grow to ~60 terms.  243 constraints per
    hash; `hashes` permutations are chained (output 0 feeds input 0).  The real circuit's R1CS / witness cannot be made
    here (no circom / snarkjs) and circomlib's constants are not in the tree: constants come from splitmix64("plonkit").
    Returns (circuit.R1CS, witness list): wire 0 = ONE, wire 1 = the public output, then the two private inputs."""
    from .circuit import R1CS
    rng = SplitMix64(SEED ^ 0x5EED)
    rounds = FULL_ROUNDS + PARTIAL_ROUNDS
    rc = [[rng.field() for _ in range(T)] for _ in range(rounds)]
    mds = [[rng.field() for _ in range(T)] for _ in range(T)]
    witness = [1, 0, inputs[0] % R_MOD, inputs[1] % R_MOD]   # ONE, out (filled at the end), in0, in1
    constraints = []

    def lc_eval(lc):
        return sum(c * witness[v] for v, c in lc.items()) % R_MOD

    def lc_add_scaled(dst, src, k):
        for v, c in src.items():
            dst[v] = (dst.get(v, 0) + k * c) % R_MOD

    def new_wire(val):
        witness.append(val % R_MOD)
        return len(witness) - 1

    def mul(a_lc, b_lc):
        """new wire w with constraint a_lc * b_lc = w"""
        w = new_wire(lc_eval(a_lc) * lc_eval(b_lc))
        constraints.append((dict(a_lc), dict(b_lc), {w: 1}))
        return w

    # state as linear combinations over wires (wire 0 carries constants)
    state = [{0: 0}, {2: 1}, {3: 1}]
    for _ in range(hashes):
        for r in range(rounds):
            full = r < FULL_ROUNDS // 2 or r >= FULL_ROUNDS // 2 + PARTIAL_ROUNDS
            # add round constants
            for i in range(T):
                state[i] = dict(state[i])
                state[i][0] = (state[i].get(0, 0) + rc[r][i]) % R_MOD
            # S-box on the full state or on element 0
            for i in range(T if full else 1):
                x = state[i]
                x2 = mul(x, x)
                x4 = mul({x2: 1}, {x2: 1})
                x5 = mul({x4: 1}, x)
                state[i] = {x5: 1}
            # MDS mix, kept linear
            new_state = []
            for i in range(T):
                lc = {}
                for j in range(T):
                    lc_add_scaled(lc, state[j], mds[i][j])
                new_state.append({v: c for v, c in lc.items() if c})
            state = new_state
        # chain: the next permutation starts from (out0, in-lane 1, in-lane 2) of this one
    out_val = lc_eval(state[0])
    witness[1] = out_val
    # out === state[0]: 1 * LC = out
    constraints.append(({0: 1}, dict(state[0]), {1: 1}))

    def as_lc(d):
        return sorted((v, c) for v, c in d.items() if c)
    cons = [(as_lc(a), as_lc(b), as_lc(c)) for a, b, c in constraints]
    n_vars = len(witness)
    return R1CS(num_inputs=2, num_aux=n_vars - 2, num_variables=n_vars, constraints=cons), witness


def write_r1cs_bin(r1cs, path, n_pub_out=1, n_pub_in=0):
    """iden3 `.r1cs` (src/r1cs_file.rs:100-154): header, constraints, identity wire->label map."""
    import struct
    prime = R_MOD.to_bytes(32, "little")
    n_prv = r1cs.num_variables - 1 - n_pub_out - n_pub_in
    header = struct.pack("<I", 32) + prime + struct.pack("<IIIIQI", r1cs.num_variables, n_pub_out, n_pub_in, n_prv,
                                                         r1cs.num_variables, len(r1cs.constraints))

    def vec(lc):
        return struct.pack("<I", len(lc)) + b"".join(struct.pack("<I", v) + (c % R_MOD).to_bytes(32, "little") for v, c in lc)
    body = b"".join(vec(a) + vec(b) + vec(c) for a, b, c in r1cs.constraints)
    wmap = b"".join(struct.pack("<Q", i) for i in range(r1cs.num_variables))
    with open(path, "wb") as f:
        f.write(b"r1cs" + struct.pack("<II", 1, 3))
        for st, data in ((1, header), (2, body), (3, wmap)):
            f.write(struct.pack("<IQ", st, len(data)) + data)


def write_wtns(witness, path):
    """snarkjs `.wtns` (src/reader.rs:124-175)"""
    import struct
    body = b"".join((int(v) % R_MOD).to_bytes(32, "little") for v in witness)
    with open(path, "wb") as f:
        f.write(b"wtns" + struct.pack("<II", 2, 2) + struct.pack("<IQ", 1, 40) + struct.pack("<I", 32) + R_MOD.to_bytes(32, "little") +
                struct.pack("<I", len(witness)) + struct.pack("<IQ", 2, len(body)) + body)


# ---------------------------------------------------------------- circuits with the Rescue x^5 custom gate
def rescue_chain_assembly(log_n: int, inputs=(3, 4, 5)):
    """A circuit of the recursive prover's shape (src/recursive/mod.rs:111-127: ProvingAssembly over Width4MainGateWithDNext +
    the Rescue x^5 custom gate): the poseidon-shaped chain above with every S-box as ONE custom-gate row (a = x, b = x^2,
    c = x^4, d = x^5) instead of three multiplication gates; the MDS rows stay main-gate rows.  2^log_n - 1 gates, one public
    input.  Returns (Assembly, gate_type) with gate_type[row] = 1 on custom-gate rows."""
    n = 1 << log_n
    n_gates = n - 1
    rng = SplitMix64(SEED ^ 0x7E5C)
    rounds = FULL_ROUNDS + PARTIAL_ROUNDS
    rc = [[rng.field() for _ in range(T)] for _ in range(rounds)]
    mds = [[rng.field() for _ in range(T)] for _ in range(T)]
    M1 = R_MOD - 1
    vals = [0] + [v % R_MOD for v in inputs]
    rows = [(1, 0, 0, 0, [M1, 0, 0, 0, 0, 0, 0], 0)]
    state = [1, 2, 3]
    r = 0
    while len(rows) < n_gates:
        full = (r % rounds) < FULL_ROUNDS // 2 or (r % rounds) >= FULL_ROUNDS // 2 + PARTIAL_ROUNDS
        for i in range(T if full else 1):
            if len(rows) >= n_gates:
                break
            x = vals[state[i]]
            x2 = x * x % R_MOD
            x4 = x2 * x2 % R_MOD
            x5 = x4 * x % R_MOD
            base = len(vals)
            vals.extend([x2, x4, x5])
            rows.append((state[i], base, base + 1, base + 2, [0] * 7, 1))
            state[i] = base + 2
        new_state = []
        for i in range(T):
            if len(rows) >= n_gates:
                break
            k = rc[r % rounds][i]
            vals.append((mds[i][0] * vals[state[0]] + mds[i][1] * vals[state[1]] + mds[i][2] * vals[state[2]] + k) % R_MOD)
            rows.append((state[0], state[1], state[2], len(vals) - 1, [mds[i][0], mds[i][1], mds[i][2], M1, 0, k, 0], 0))
            new_state.append(len(vals) - 1)
        if len(new_state) == T:
            state = new_state
        r += 1
    from .circuit import assembly_from_rows
    asm = assembly_from_rows([row[:5] for row in rows], vals, 1)
    gate_type = np.zeros(n, dtype=np.uint8)
    gate_type[:len(rows)] = [row[5] for row in rows]
    return asm, gate_type
