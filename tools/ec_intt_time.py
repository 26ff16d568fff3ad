#!/usr/bin/env python3
"""Times pk_ec_intt_g1 (Crs::from_powers, dump-lagrange) at a few sizes: wall time of the call (device work + D2H of the key)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from plonkit_b200 import _lib  # noqa: E402

ctx = _lib.Context(0)
top = int(sys.argv[1]) if len(sys.argv) > 1 else 20
srs = ctx.srs_gen(1 << top, 42)
ctx.srs_load_g1(srs)
for lg in range(12, top + 1, 2):
    ctx.ec_intt_g1(lg)
    t = time.perf_counter()
    ctx.ec_intt_g1(lg)
    print("EC-iNTT 2^%d: %.1f ms" % (lg, (time.perf_counter() - t) * 1e3), flush=True)
