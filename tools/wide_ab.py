#!/usr/bin/env python
"""A/B of the team kernel k_miller (12 shared-memory slots per thread: 23 teams of 11 = 8 warps per SM) against
k_miller_wide (10 slots, evaluation points read from the batch arrays: up to 29 teams = 10 warps per SM) on the
headline batch (2^14 MultPoly products, 11 x 11 slots, keyBits = 512) for several teams-per-block settings.
Prints one JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bgn_b200 import Engine, bench_imad_peak, workmodel  # noqa: E402

D = 11


def main():
    with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb512.json")) as f:
        g = json.load(f)
    p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    eng = Engine(p, n, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=0)
    EB, SB = eng.elem_bytes, eng.scalar_bytes
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    ms, ipt = bench_imad_peak(0, 4096, 148 * 8, 256)
    peak = 148 * 8 * 256 * ipt / (ms * 1e-3)
    cnt = 1 << 14
    xs = torch.randint(-1, 2, (cnt * D,), generator=gen, device=dev, dtype=torch.int64)
    r = torch.randint(0, 256, (cnt * D, SB), generator=gen, device=dev, dtype=torch.uint8)
    r[:, 0] &= 0x3F
    a = eng.encrypt_batch(xs, r.reshape(-1))
    b = eng.encrypt_batch(xs.flip(0), r.flip(0).reshape(-1))
    eng.timing_enable(True)
    prod = workmodel.miller_unit_products(p, n, l, D, D)
    rows, ref = [], None
    for n_units in (cnt, 4144, 3404):
        for tpb in (0, 23, 26, 27, 28, 29):
            eng.set_option("miller_wide", tpb)
            eng.set_option("miller_split", 0)
            out = torch.empty(n_units * 2 * D * EB, dtype=torch.uint8, device=dev)
            eng.multpoly_batch(a[: n_units * D * EB], D, b[: n_units * D * EB], D, n_units, out=out)
            best = None
            for _ in range(2):
                eng.timing_reset()
                eng.multpoly_batch(a[: n_units * D * EB], D, b[: n_units * D * EB], D, n_units, out=out)
                k = eng.timing_get("k_miller")[0]
                best = k if best is None else min(best, k)
            if tpb == 0:
                ref = out.clone()
            row = {"units": n_units, "teams_per_block": tpb or "k_miller (23)", "ms": best, "imad_frac": n_units * prod / (best * 1e-3) / peak,
                   "bytes_equal": bool((out == ref).all().item())}
            rows.append(row)
            print(json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps({"imad_wide_peak_T": peak / 1e12, "products_per_unit": prod, "rows": rows}, indent=1))
    eng.close()


if __name__ == "__main__":
    main()
