#!/usr/bin/env python
"""Which instruction classes share the integer-multiply pipe?  

Runs bgn_bench_issue_mix on cuda:0 for every mix and reports, per class, instructions / clk / SM at the
measured SM clock, plus the ratio T(mix) / (T(part A) + T(part B)) and T(mix) / max(T(A), T(B)): a mix
whose time is the SUM of its parts runs on one pipe; one whose time is the MAX co-issues.  One
IMAD.WIDE.U32 = one 32x32->64 product; one (IMAD.LO, IMAD.HI) pair = one product; FFMA / DFMA are listed
in instructions (a 32x32 product rebuilt from 24-bit / 53-bit mantissa pieces needs several).
Prints one JSON object."""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bgn_b200.engine import bench_issue_mix  # noqa: E402

SM = 148
NAMES = {0: "IMAD.WIDE", 1: "IMAD.LO+IMAD.HI", 2: "IMAD.WIDE + (LO,HI) 1:1", 3: "FFMA", 4: "IMAD.WIDE + FFMA 1:1",
         5: "DFMA", 6: "IMAD.WIDE + DFMA 1:1", 7: "IMAD.WIDE + (LO,HI) 2:1", 8: "IMAD.WIDE + DFMA 2:1",
         9: "IMAD.WIDE + DFMA 1:2", 10: "IMAD.LO", 11: "IMAD.HI", 12: "IMAD.WIDE + IMAD.LO 1:1",
         13: "IMAD.WIDE + IMAD.HI 1:1"}
PARTS = {1: (10, 11), 2: (0, 1), 4: (0, 3), 6: (0, 5), 12: (0, 10), 13: (0, 11)}


def sm_clock_mhz():
    try:
        out = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm", "--format=csv,noheader,nounits", "-i", "0"])
        return float(out.decode().split()[0])
    except Exception:
        return 1965.0


def main():
    blocks, threads, iters = SM * 8, 256, 2048
    res = {"config": {"blocks": blocks, "threads": threads, "iters": iters}, "mixes": {}}
    times = {}
    for mix in range(14):
        bench_issue_mix(0, mix, 64, blocks, threads)
        best = None
        for _ in range(3):
            ms, per = bench_issue_mix(0, mix, iters, blocks, threads)
            if best is None or ms < best[0]:
                best = (ms, per)
        ms, per = best
        times[mix] = ms
        mhz = sm_clock_mhz()
        clk = ms * 1e-3 * mhz * 1e6
        nthr = blocks * threads
        entry = {"name": NAMES[mix], "ms": ms, "sm_mhz_after": mhz}
        for cls, cnt in zip(("imad_wide", "imad_lo", "imad_hi", "ffma", "dfma"), per):
            if cnt:
                entry[cls + "_per_clk_per_sm"] = cnt * nthr / clk / SM
        prods = per[0] + min(per[1], per[2])  # a (LO, HI) pair is one 32x32->64 product
        if prods:
            entry["products_32x32_per_clk_per_sm"] = prods * nthr / clk / SM
        res["mixes"][str(mix)] = entry
    for mix, (a, b) in PARTS.items():
        res["mixes"][str(mix)]["time_over_sum_of_parts"] = times[mix] / (times[a] + times[b])
        res["mixes"][str(mix)]["time_over_max_of_parts"] = times[mix] / max(times[a], times[b])
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
