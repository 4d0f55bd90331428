#!/usr/bin/env python
"""k_miller_split with and without the parabola steps (option split_para) against the warps each thread role
occupies per SM: 11 x 11 slots, keyBits = 512.  One JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bgn_b200 import Engine

D = 11

def main():
    with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb512.json")) as f:
        g = json.load(f)
    eng = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=0)
    EB, SB = eng.elem_bytes, eng.scalar_bytes
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    nmax = 2516
    xs = torch.randint(-1, 2, (nmax * D,), generator=gen, device=dev, dtype=torch.int64)
    r = torch.randint(0, 256, (nmax * D, SB), generator=gen, device=dev, dtype=torch.uint8)
    r[:, 0] &= 0x3F
    a = eng.encrypt_batch(xs, r.reshape(-1))
    b = eng.encrypt_batch(xs.flip(0), r.flip(0).reshape(-1))
    eng.timing_enable(True)
    eng.set_option("miller_split", 1)
    rows = []
    for cnt in (148, 296, 444, 592, 740, 888, 1036, 1184, 1332, 1480, 1776, 2072, 2368, 2516):
        row = {"count": cnt, "teams_per_sm": (cnt + 147) // 148, "warps_per_role": (((cnt + 147) // 148) * D + 31) // 32}
        ref = None
        for para in (0, 1):
            eng.set_option("split_para", para)
            out = torch.empty(cnt * 2 * D * EB, dtype=torch.uint8, device=dev)
            best = None
            for _ in range(3):
                eng.timing_reset()
                eng.multpoly_batch(a[: cnt * D * EB], D, b[: cnt * D * EB], D, cnt, out=out)
                k = eng.timing_get("k_miller_split")[0]
                best = k if best is None else min(best, k)
            row["para%d_ms" % para] = best
            if ref is None:
                ref = out.clone()
            else:
                row["bytes_equal"] = bool((out == ref).all().item())
        rows.append(row)
        print(json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps({"rows": rows}, indent=1))
    eng.close()

main()
