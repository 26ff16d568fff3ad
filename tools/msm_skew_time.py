#!/usr/bin/env python3
"""MSM 2^20 through the C ABI (host buffers) with uniform scalars and with the witness-like mix of SURVEY 8(d) (40 % zero,
10 % one, 50 % uniform: what the Lagrange path feeds): wall time per call, results checked against each other's structure."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonkit_b200 import _lib, synth  # noqa: E402
from plonkit_b200.bn254 import ints_to_limbs  # noqa: E402

ctx = _lib.Context(0)
n = 1 << 20
ctx.srs_load_g1(ctx.srs_gen(n, 42))
s = synth.random_field_elements(n, seed=50)
w = s.copy()
sel = np.random.default_rng(20).random(n)
w[sel < 0.4] = 0
w[(sel >= 0.4) & (sel < 0.5)] = ints_to_limbs([1])[0]
ones = np.tile(ints_to_limbs([1]), (n, 1))
for name, x in (("uniform", s), ("witness-like", w), ("all ones", ones)):
    ctx.msm_g1(x)
    t = time.perf_counter()
    for _ in range(5):
        ctx.msm_g1(x)
    print("MSM 2^20 %-13s %.2f ms per call (incl. 32 MB H2D)" % (name, (time.perf_counter() - t) / 5 * 1e3), flush=True)
