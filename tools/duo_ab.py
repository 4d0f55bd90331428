#!/usr/bin/env python
"""A/B of the knobs of k_pair_duo (pairwarp.cuh) on cuda:0 at keyBits=512: loop shape of the products
(0 = unrolled, 1 / 2 / 4 = 2U rows per iteration), warp pairs per block, and whether the pairs of a block
run in lockstep (one block-wide barrier per step) or each on its own named barrier.  Kernel ms, best of 3.
Prints one JSON object.   usage: tools/duo_ab.py"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bgn_b200 import Engine, bench_imad_peak, workmodel  # noqa: E402


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
    nmax = 37888
    xs = torch.randint(-1000, 1000, (nmax,), generator=gen, device=dev, dtype=torch.int64)
    r = torch.randint(0, 256, (nmax, SB), generator=gen, device=dev, dtype=torch.uint8)
    r[:, 0] &= 0x3F
    a = eng.encrypt_batch(xs, r.reshape(-1))
    b = eng.encrypt_batch(xs.flip(0), r.flip(0).reshape(-1))
    eng.timing_enable(True)
    prod = workmodel.pair_duo_products(p, n, l)
    eng.set_option("pair_duo", 0)
    ref = {}
    rows = []
    for cnt in (1024, 4096, 16384, 37888):
        ref[cnt] = eng.pair_batch(a[: cnt * EB], b[: cnt * EB])
    eng.set_option("pair_duo", 1)
    for loop in (0, 4, 2, 1):
        for pairs, bb in ((1, 0), (2, 1), (4, 1), (4, 0)):
            eng.set_option("pair_duo_loop", loop)
            eng.set_option("pair_duo_pairs", pairs)
            eng.set_option("pair_duo_blockbar", bb)
            row = {"loop": loop, "pairs_per_block_max": pairs, "block_barrier": bb}
            for cnt in (1024, 4096, 16384, 37888):
                out = torch.empty(cnt * EB, dtype=torch.uint8, device=dev)
                eng.pair_batch(a[: cnt * EB], b[: cnt * EB], out=out)
                best = None
                for _ in range(3):
                    eng.timing_reset()
                    eng.pair_batch(a[: cnt * EB], b[: cnt * EB], out=out)
                    k = eng.timing_get("k_pair_duo")[0]
                    best = k if best is None else min(best, k)
                row[str(cnt)] = {"ms": best, "imad_frac": cnt * prod / (best * 1e-3) / peak,
                                 "ok": bool((out == ref[cnt]).all().item())}
            rows.append(row)
            print(json.dumps(row), file=sys.stderr, flush=True)
    print(json.dumps({"imad_wide_peak_T": peak / 1e12, "rows": rows}, indent=1))
    eng.close()


if __name__ == "__main__":
    main()
