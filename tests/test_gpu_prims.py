"""GPU parity tests of the bellman primitives behind the C ABI (pk_ntt, pk_lde4, pk_msm_g1, pk_srs_gen,
pk_ec_intt_g1) against the oracle on the same seeded inputs, plus the committed golden vectors.  Bit-exact."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from plonkit_b200 import _lib, synth
from plonkit_b200.bn254 import Q_MOD, R_MOD, ints_to_limbs, limbs_to_ints

pytestmark = pytest.mark.gpu


def bitrev(i, bits):
    return int(format(i, "0%db" % bits)[::-1], 2) if bits else 0


def test_field_multiplier_runs_and_reports_throughput(ctx):
    for which in (0, 1):
        g = ctx.bench_fieldmul(which)
        assert g > 1.0, "field multiplier throughput %.2f Gmul/s is implausibly low" % g


@pytest.mark.parametrize("log_n", [1, 2, 3, 5, 9, 10, 11, 13, 16])
def test_ntt_forward_inverse_and_coset_match_oracle(ctx, orc, log_n):
    x = synth.random_field_elements(1 << log_n, seed=100 + log_n)
    for inverse in (False, True):
        for coset in (False, True):
            got = ctx.ntt(x, inverse=inverse, coset=coset)
            assert (got == orc.ntt(x, inverse=inverse, coset=coset, threads=8)).all(), (log_n, inverse, coset)
    # Montgomery-form interface: same transform on the in-memory representation
    xm = ints_to_limbs([v * (1 << 256) % R_MOD for v in limbs_to_ints(x)])
    ym = ctx.ntt(xm, fmt=_lib.FMT_MONTGOMERY)
    assert limbs_to_ints(ym) == [v * (1 << 256) % R_MOD for v in limbs_to_ints(orc.ntt(x, threads=8))]


def test_ntt_golden_vector_and_edge_inputs(ctx, orc):
    x = synth.random_field_elements(1 << 10, seed=synth.SEED + 1)
    assert (ctx.ntt(x) == np.load(os.path.join(GOLDEN, "ntt10_out.npy"))).all()
    z = np.zeros((1 << 8, 4), dtype=np.uint64)
    assert not ctx.ntt(z).any()
    one = z.copy()
    one[0] = ints_to_limbs([1])[0]
    assert limbs_to_ints(ctx.ntt(one)) == [1] * 256                      # delta -> all ones
    top = np.tile(ints_to_limbs([R_MOD - 1]), (256, 1))
    assert (ctx.ntt(top) == orc.ntt(top)).all()                         # maximal elements
    with pytest.raises(_lib.SynthesisError):
        ctx.ntt(np.zeros((3, 4), dtype=np.uint64))                      # not a power of two


def test_ntt_and_lde4_at_bench_size_2pow20_match_oracle(ctx, orc):
    """BASELINE configs[1] size: forward / inverse / coset NTT and the x4 LDE of 2^20 elements == oracle, element for
    element; plus the round trip."""
    log_n = 20
    x = synth.random_field_elements(1 << log_n, seed=1)
    y = ctx.ntt(x)
    assert (y == orc.ntt(x, threads=16)).all()
    assert (ctx.ntt(y, inverse=True) == x).all()
    assert (ctx.ntt(x, inverse=True, coset=True) == orc.ntt(x, inverse=True, coset=True, threads=16)).all()
    ref = orc.lde4(x, threads=16)
    assert (ctx.lde4(x) == ref).all()
    br = ctx.lde4(x, bitreversed=True)          # the order the prover keeps (bellman's bitreversed LDE)
    idx = np.arange(4 << log_n, dtype=np.uint32)
    rev = np.zeros_like(idx)
    for b in range(log_n + 2):
        rev |= ((idx >> np.uint32(b)) & np.uint32(1)) << np.uint32(log_n + 1 - b)
    assert (br[rev] == ref).all()


@pytest.mark.parametrize("log_n", [1, 3, 8, 12])
def test_lde4_natural_and_bitreversed_orders(ctx, orc, log_n):
    c = synth.random_field_elements(1 << log_n, seed=200 + log_n)
    ref = orc.lde4(c, threads=8)
    assert (ctx.lde4(c) == ref).all()
    br = ctx.lde4(c, bitreversed=True)
    perm = np.array([bitrev(i, log_n + 2) for i in range(4 << log_n)])
    assert (br[perm] == ref).all()


def _load(ctx, bases, window_bits=0):
    ctx.srs_load_g1(bases, window_bits)


def test_msm_matches_oracle_on_random_and_golden_inputs(ctx, orc, simple_key):
    _load(ctx, simple_key.g1_bases)
    x = synth.random_field_elements(1 << 10, seed=synth.SEED + 1)
    assert (ctx.msm_g1(x) == np.load(os.path.join(GOLDEN, "msm10_out.npy"))).all()
    for n in (1, 2, 31, 32, 33, 257, 1000):
        s = synth.random_field_elements(n, seed=300 + n)
        assert (ctx.msm_g1(s) == orc.msm(s, simple_key.g1_bases[:n], threads=8)).all(), n
    # base_offset
    s = synth.random_field_elements(100, seed=7)
    assert (ctx.msm_g1(s, base_offset=37) == orc.msm(s, simple_key.g1_bases[37:137], threads=4)).all()
    # Montgomery-form scalars
    sm = ints_to_limbs([v * (1 << 256) % R_MOD for v in limbs_to_ints(s)])
    assert (ctx.msm_g1(sm, fmt=_lib.FMT_MONTGOMERY) == orc.msm(s, simple_key.g1_bases[:100], threads=4)).all()


@pytest.mark.parametrize("window_bits", [0, 4, 9, 13, 16])
def test_msm_edge_cases(ctx, orc, simple_key, window_bits):
    bases = simple_key.g1_bases[:256].copy()
    bases[10] = bases[11]                        # repeated base
    neg = bases[20].copy()
    neg[4:] = ints_to_limbs([Q_MOD - limbs_to_ints(bases[20][4:])[0]])[0]
    bases[21] = neg                              # P and -P
    bases[30] = 0                                # a base at infinity
    _load(ctx, bases, window_bits)
    n = 256
    cases = {}
    cases["zeros"] = np.zeros((n, 4), dtype=np.uint64)
    cases["ones"] = np.tile(ints_to_limbs([1]), (n, 1))
    cases["max"] = np.tile(ints_to_limbs([R_MOD - 1]), (n, 1))
    w = synth.random_field_elements(n, seed=11)
    w[::3] = 0
    w[1::5] = ints_to_limbs([1])[0]
    w[2::7] = ints_to_limbs([2])[0]
    cases["witness-like"] = w
    c = np.zeros((n, 4), dtype=np.uint64)
    c[20] = ints_to_limbs([12345])[0]
    c[21] = ints_to_limbs([12345])[0]            # k*P + k*(-P) = infinity
    cases["cancel"] = c
    d = np.zeros((n, 4), dtype=np.uint64)
    d[10] = ints_to_limbs([5])[0]
    d[11] = ints_to_limbs([5])[0]                # P + P through the same bucket (doubling path)
    cases["double"] = d
    cases["pow2"] = ints_to_limbs([1 << (i % 254) for i in range(n)])
    cases["half-window"] = ints_to_limbs([(1 << 15) + (1 << 31) * (i & 1) for i in range(n)])
    for name, s in cases.items():
        got = ctx.msm_g1(s)
        assert (got == orc.msm(s, bases, threads=4)).all(), (name, window_bits)
    assert not ctx.msm_g1(cases["cancel"]).any()                     # infinity -> (0, 0)
    assert not ctx.msm_g1(np.zeros((0, 4), dtype=np.uint64)).any()   # empty input
    with pytest.raises(_lib.SynthesisError) as e:
        ctx.msm_g1(np.zeros((257, 4), dtype=np.uint64))              # longer than the SRS
    assert e.value.code == 2


def test_msm_2pow16_uniform_and_skewed(ctx, orc):
    n = 1 << 16
    bases = orc.srs_gen(n, 42, threads=8)
    _load(ctx, bases)
    s = synth.random_field_elements(n, seed=5)
    assert (ctx.msm_g1(s) == orc.msm(s, bases, threads=8)).all()
    s[: n // 2] = ints_to_limbs([1])[0]          # one giant bucket
    s[n // 2: n // 2 + n // 4] = 0
    assert (ctx.msm_g1(s) == orc.msm(s, bases, threads=8)).all()
    # many repeated bases and P / -P pairs inside the same buckets: P + P and P + (-P) inside the accumulation
    b2 = bases.copy()
    b2[1::2] = b2[0::2]
    neg = limbs_to_ints(b2[2::4, 4:].reshape(-1, 4))
    b2[2::4, 4:] = ints_to_limbs([Q_MOD - v for v in neg]).reshape(-1, 4)
    _load(ctx, b2)
    s2 = synth.random_field_elements(n, seed=6)
    s2[1::2] = s2[0::2]
    assert (ctx.msm_g1(s2) == orc.msm(s2, b2, threads=8)).all()


@pytest.fixture(scope="module")
def srs20(ctx, orc):
    """[42^i]G, i < 2^20, made on the device and spot-checked against the oracle: the first 2^10 points against the
    oracle's generator, 64 random ones against a fixed-base multiplication G * (42^i mod r)."""
    n = 1 << 20
    srs = ctx.srs_gen(n, 42)
    assert (srs[:1024] == orc.srs_gen(1024, 42, threads=8)).all()
    g = ints_to_limbs([1, 2]).reshape(8)
    rng = np.random.default_rng(7)
    for i in rng.integers(1024, n, size=64):
        assert (srs[i] == orc.g1_mul(g, pow(42, int(i), R_MOD))).all(), i
    return srs


@pytest.mark.parametrize("log_n", [18, 20])
def test_msm_at_bench_size_default_window_matches_oracle(ctx, orc, srs20, log_n):
    """The MSM configuration bench.py times: N = 2^20 bases, default window plan (c = 20, 13 windows), and 2^18
    (c = 18); uniform scalars and the witness-like mix of SURVEY 8(d) (40 % zero, 10 % one, 50 % uniform)."""
    n = 1 << log_n
    _load(ctx, srs20[:n])
    s = synth.random_field_elements(n, seed=50 + log_n)
    assert (ctx.msm_g1(s) == orc.msm(s, srs20[:n], threads=16)).all()
    w = s.copy()
    sel = np.random.default_rng(log_n).random(n)
    w[sel < 0.4] = 0
    w[(sel >= 0.4) & (sel < 0.5)] = ints_to_limbs([1])[0]
    assert (ctx.msm_g1(w) == orc.msm(w, srs20[:n], threads=16)).all()


@pytest.mark.parametrize("window_bits", [21, 22])
def test_msm_wide_windows_with_a_three_set_tail(ctx, orc, simple_key, window_bits):
    """Explicit window_bits >= 21 with an 11-commitment batch (4 + 4 + 3 sets, what make_verification_key issues): the
    3-set tail has 3 * 2^(c-1) buckets, which is not a power of two."""
    from plonkit_b200 import plonk, reader
    asm = synth.poseidon_chain_assembly(10)
    key = reader.Crs(simple_key.g1_bases)
    ctx.srs_load_g1(key.g1_bases, window_bits, tag=(key.token, asm.n))
    setup = plonk.SetupForProver.prepare_setup_for_prover(asm, key, None, ctx=ctx)
    com = orc.setup_commitments(asm.n, asm.num_inputs, asm.wire_idx, asm.selectors, key.g1_bases, nvars=asm.nvars, threads=8)
    from conftest import vk_commitments
    assert (vk_commitments(setup.make_verification_key()) == com).all()
    setup.close()


def test_srs_generator_matches_reference_key(ctx, simple_key):
    """Crs::crs_42 (src/plonk.rs:41,47) == keys/setup/setup_2^10.key"""
    assert (ctx.srs_gen(1024, 42) == simple_key.g1_bases).all()


def test_ec_intt_matches_oracle(ctx, orc, simple_key):
    """Crs::from_powers (src/plonk.rs:179-185)"""
    _load(ctx, simple_key.g1_bases)
    for log_n in (0, 1, 4, 7):
        got = ctx.ec_intt_g1(log_n)
        assert (got == orc.ec_intt(simple_key.g1_bases[: 1 << log_n], threads=8)).all(), log_n


def test_ec_intt_2pow12_full_compare(ctx, orc, srs20):
    """Crs::from_powers at 2^12: every output point == the oracle's EC inverse FFT (the oracle needs ~2 s here)."""
    _load(ctx, srs20[: 1 << 12])
    assert (ctx.ec_intt_g1(12) == orc.ec_intt(srs20[: 1 << 12], threads=16)).all()


@pytest.mark.parametrize("log_n,samples", [(16, 1 << 16), (20, 1 << 10)])
def test_ec_intt_outputs_are_the_lagrange_basis_in_the_exponent(ctx, orc, srs20, log_n, samples):
    """BASELINE configs[3] check (SURVEY 8d cfg 4): out[i] == L_i(42) * G with L_i(tau) = w^i (tau^N - 1) / (N (tau - w^i))
    computed in Fr with Python integers and one fixed-base multiplication by the oracle.  2^16: EVERY index (full
    compare; the oracle's own EC inverse FFT would need a minute here, the closed form is exact and independent of it);
    2^20: 2^10 random indices of the full-size run."""
    n = 1 << log_n
    _load(ctx, srs20[:n])
    out = ctx.ec_intt_g1(log_n)
    g = ints_to_limbs([1, 2]).reshape(8)
    tau = 42
    w = orc.omega(log_n)
    zh = (pow(tau, n, R_MOD) - 1) % R_MOD
    ninv = pow(n, R_MOD - 2, R_MOD)
    idx = np.arange(n) if samples >= n else np.unique(np.random.default_rng(log_n).integers(0, n, size=samples))
    # batch-invert (tau - w^i) with Montgomery's trick in Python integers
    wi = [pow(w, int(i), R_MOD) for i in idx] if samples < n else None
    if wi is None:
        wi, x = [], 1
        for _ in range(n):
            wi.append(x)
            x = x * w % R_MOD
    den = [(tau - v) % R_MOD for v in wi]
    pre, acc = [], 1
    for d in den:
        pre.append(acc)
        acc = acc * d % R_MOD
    inv = pow(acc, R_MOD - 2, R_MOD)
    scal = [0] * len(den)
    for k in range(len(den) - 1, -1, -1):
        scal[k] = wi[k] * zh % R_MOD * ninv % R_MOD * (inv * pre[k] % R_MOD) % R_MOD
        inv = inv * den[k] % R_MOD
    exp = orc.g1_mul_fixed(g, scal, threads=16)
    assert (out[idx] == exp).all()


@pytest.mark.parametrize("n", [1, 2, 255, 4096, 100003, 1 << 20])
def test_polynomial_primitives_match_oracle(ctx, orc, n):
    """pk_poly_* (bellman's evaluate_at, divide_single, calculate_shifted_grand_product, batch_inversion) == oracle;
    at 2^20 the serial checker still finishes in seconds."""
    c = synth.random_field_elements(n, seed=900 + n % 1000)
    c[n // 3] = 0                                     # a zero inside the grand product / inversion
    z = synth.random_field_elements(1, seed=77)[0]
    assert (ctx.poly_evaluate_at(c, z) == orc.poly_op("evaluate_at", c, z)).all()
    assert (ctx.poly_divide_by_linear(c, z) == orc.poly_op("divide_by_linear", c, z)).all()
    zero = np.zeros(4, dtype=np.uint64)
    assert (ctx.poly_divide_by_linear(c, zero) == orc.poly_op("divide_by_linear", c, zero)).all()
    assert (ctx.poly_shifted_grand_product(c) == orc.poly_op("shifted_grand_product", c)).all()
    assert (ctx.poly_batch_inversion(c) == orc.poly_op("batch_inversion", c)).all()
    nz = c.copy()
    nz[n // 3] = ints_to_limbs([5])[0]
    inv = ctx.poly_batch_inversion(nz)
    if n <= 4096:                                     # x * x^-1 == 1 checked with Python integers
        assert all(a * b % R_MOD == 1 for a, b in zip(limbs_to_ints(nz), limbs_to_ints(inv)))


@pytest.mark.parametrize("n", [1, 7, 4096, 100003])
def test_pointwise_polynomial_operations_match_python_integers(ctx, n):
    """pk_poly_pointwise (bellman Polynomial::add_assign_scaled / mul_assign / scale / add_constant / distribute_powers, SURVEY
    8 row a13) against Python integers; the same kernels (lincomb, pointwise product, power table) carry the prover's
    linearisation and opening aggregation."""
    a = synth.random_field_elements(n, seed=700 + n % 100)
    b = synth.random_field_elements(n, seed=800 + n % 100)
    a[n // 2] = 0
    s = synth.random_field_elements(1, seed=9)[0]
    ai, bi, si = limbs_to_ints(a), limbs_to_ints(b), limbs_to_ints(s.reshape(1, 4))[0]
    want = {
        "add_assign_scaled": [(x + si * y) % R_MOD for x, y in zip(ai, bi)],
        "mul_assign": [x * y % R_MOD for x, y in zip(ai, bi)],
        "scale": [si * x % R_MOD for x in ai],
        "add_constant": [(x + si) % R_MOD for x in ai],
    }
    p, dp = 1, []
    for x in ai:
        dp.append(x * p % R_MOD)
        p = p * si % R_MOD
    want["distribute_powers"] = dp
    for op, ref in want.items():
        got = ctx.poly_pointwise(op, a, b if op in ("add_assign_scaled", "mul_assign") else None, None if op == "mul_assign" else s)
        assert limbs_to_ints(got) == ref, op
    assert ctx.poly_pointwise("scale", np.zeros((0, 4), dtype=np.uint64), None, s).shape == (0, 4)
    with pytest.raises(_lib.SynthesisError):
        ctx.poly_pointwise("add_assign_scaled", a, None, s)          # second operand missing


def test_polynomial_primitives_empty_input(ctx):
    e = np.zeros((0, 4), dtype=np.uint64)
    z = ints_to_limbs([3])[0]
    assert not ctx.poly_evaluate_at(e, z).any()
    assert ctx.poly_divide_by_linear(e, z).shape == (0, 4)
    assert ctx.poly_shifted_grand_product(e).shape == (0, 4) and ctx.poly_batch_inversion(e).shape == (0, 4)
