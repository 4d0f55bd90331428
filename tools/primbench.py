#!/usr/bin/env python
"""Memory-operand primitive microbenchmarks (k_prim_bench) at L = 17: fraction of the IMAD.WIDE
peak reached by mul-only, line_mul, sqr2 and dbl_line sequences at several occupancies."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bgn_b200 import Engine, bench_imad_peak, workmodel

g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb512.json")))
e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
ms, ipt = bench_imad_peak(0, 4096, 148 * 8, 256)
peak = 148 * 8 * 256 * ipt / (ms * 1e-3)
MM = {1: 1, 10: 1, 11: 5, 12: 2, 13: 12, 20: 1, 22: 2}
for base in (30, 40, 50, 60, 70):
    MM.update({base: 5, base + 2: 2, base + 3: 12, base + 4: 13})
MM[80] = MM[81] = MM[82] = MM[83] = 5  # lazy line_mul: counted as 5 modmuls = one line_mul, so frac compares line_mul rates
MODES = [int(x) for x in sys.argv[1].split(',')] if len(sys.argv) > 1 else [1, 20, 11, 70, 50, 60, 40, 73, 53, 63, 43, 72, 42, 74, 44]
for mode in MODES:
    for threads in (128, 256):
        if mode == 1 and threads > 128:
            continue
        iters = 400
        ms = e.bench_mulmod(mode, iters, 148, threads)
        mm = 148 * threads * iters * MM[mode] / (ms * 1e-3)
        print("mode %2d threads %3d  %.2f Gmodmul/s  frac %.3f" % (mode, threads, mm / 1e9, mm * 595 / peak))
