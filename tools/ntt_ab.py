#!/usr/bin/env python3
"""A/B of the two NTT pass kernels (plain 128-bit loads vs TMA bulk staging): run once per setting of PK_NTT_TMA.
    PK_NTT_TMA=0 python tools/ntt_ab.py ; PK_NTT_TMA=1 python tools/ntt_ab.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonkit_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
for lg in (16, 18, 20, 22, 24):
    ctx.bench_ntt(lg, 3)
    best = min(ctx.bench_ntt(lg, 20) for _ in range(3))
    print("PK_NTT_TMA=%s  NTT 2^%d: %.4f ms  (%.2f Gelem/s)" % (os.environ.get("PK_NTT_TMA", "0"), lg, best, (1 << lg) / best / 1e6), flush=True)
