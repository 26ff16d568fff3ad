#!/usr/bin/env python3
"""Device-resident MSM 2^20 timing with uniform scalars, the witness-like mix of SURVEY 8(d) (40 % zero, 10 % one, 50 % uniform:
what the Lagrange path feeds) and the all-ones worst case (every entry in ONE bucket)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonkit_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
n = 1 << 20
ctx.srs_load_g1(ctx.srs_gen(n, 42))
for name, pat in (("uniform", 0), ("witness-like", 1), ("all ones", 2)):
    ctx.bench_msm_pattern(n, pat, 1)
    print("MSM 2^20 %-13s %.3f ms" % (name, min(ctx.bench_msm_pattern(n, pat, 5) for _ in range(2))), flush=True)
