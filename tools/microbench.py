#!/usr/bin/env python
"""Integer-pipe microbenchmarks on cuda:0 (K1 evidence): peak IMAD.WIDE.U32 rate and the
register-resident Montgomery product at several occupancies / ILP.  Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bgn_b200 import Engine, bench_imad_peak, workmodel  # noqa: E402

SM = 148


def main():
    res = {"imad_peak": [], "mulmod": []}
    for blocks_per_sm, threads in ((1, 128), (2, 128), (4, 128), (4, 256), (8, 256)):
        ms, ipt = bench_imad_peak(0, 4096, SM * blocks_per_sm, threads)
        res["imad_peak"].append({"blocks_per_sm": blocks_per_sm, "threads": threads,
                                 "Tinstr_per_s": SM * blocks_per_sm * threads * ipt / (ms * 1e-3) / 1e12})
    peak = max(r["Tinstr_per_s"] for r in res["imad_peak"])
    for kb in (128, 512, 1024):
        with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb%d.json" % kb)) as f:
            g = json.load(f)
        e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
        L = e.limbs
        for ilp in (1, 2):
            for blocks_per_sm, threads in ((1, 32), (1, 64), (1, 128), (2, 128), (3, 128), (4, 128)):
                iters = 2000 if L <= 17 else 500
                ms = e.bench_mulmod(ilp, iters, SM * blocks_per_sm, threads)
                mm = SM * blocks_per_sm * threads * ilp * iters / (ms * 1e-3)
                res["mulmod"].append({"L": L, "ilp": ilp, "blocks_per_sm": blocks_per_sm, "threads": threads,
                                      "Gmodmul_per_s": mm / 1e9,
                                      "frac_of_imad_peak": mm * workmodel.products_per_modmul(L) / 1e12 / peak})
        e.close()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
