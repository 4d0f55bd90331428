#!/usr/bin/env python
"""A/B of the thread mappings of a pairing on cuda:0 (north_star: "thread-per-element versus
warp-cooperative mapping is chosen by measurement"), over batch sizes from one pairing to 2^17:

  --which fixed     e(., P) (makeL2 / MakePolyL2 / level-1 decrypt): one thread per pairing with its state in
                    shared memory (k_miller_fixed) against a LANE PAIR per pairing, registers only
                    (k_miller_fixed_pair, pairlane.cuh)
  --which general   e(a, b) (Mult): one thread per pairing (k_miller, team of 1) against TWO WARPS per 32
                    pairings (k_pair_duo, pairwarp.cuh)

For each: kernel ms (CUDA events on the library's stream, best of `reps`), pairings/s, and the EXECUTED
32x32->64 products over the IMAD.WIDE peak measured in the same run.  The crossovers decide the batch-size
rules of api.cu (run_miller_fixed, pair_common).  Prints one JSON object.
usage: tools/mapping_ab.py [--which fixed|general] [--key-bits 512]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bgn_b200 import Engine, bench_imad_peak, workmodel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--key-bits", type=int, default=512)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--max-log2", type=int, default=17)
    ap.add_argument("--which", default="fixed", choices=["fixed", "general"])
    args = ap.parse_args()
    with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb%d.json" % args.key_bits)) as f:
        g = json.load(f)
    p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    eng = Engine(p, n, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=0)
    EB, SB = eng.elem_bytes, eng.scalar_bytes
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    ms, ipt = bench_imad_peak(0, 4096, 148 * 8, 256)
    peak = 148 * 8 * 256 * ipt / (ms * 1e-3)
    nmax = 1 << args.max_log2
    xs = torch.randint(-1000, 1000, (nmax,), generator=gen, device=dev, dtype=torch.int64)
    r = torch.randint(0, 256, (nmax, SB), generator=gen, device=dev, dtype=torch.uint8)
    r[:, 0] &= 0x3F
    cts = eng.encrypt_batch(xs, r.reshape(-1))
    eng.timing_enable(True)
    fixed = args.which == "fixed"
    if fixed:
        prods = {0: workmodel.miller_fixed_products(p, n, l), 1: workmodel.miller_fixed_pair_products(p, n, l)}
        option, other, prefix = "fixed_pair", "lane_pair", "k_miller_fixed"
        cts2 = None
    else:
        prods = {0: workmodel.miller_unit_products(p, n, l, 1, 1), 1: workmodel.pair_duo_products(p, n, l)}
        option, other, prefix = "pair_duo", "two_warps", None
        cts2 = eng.encrypt_batch(xs.flip(0), r.flip(0).reshape(-1))
    res = {"which": args.which, "key_bits": args.key_bits, "limbs": eng.limbs, "imad_wide_peak_T": peak / 1e12,
           "products_per_pairing": {"one_thread": prods[0], other: prods[1]}, "sizes": []}
    sizes = [1, 32, 256, 1 << 10, 1 << 12, 1 << 13, 1 << 14, 1 << 15, 37888, 1 << 16, 1 << 17]
    for cnt in [s for s in sizes if s <= nmax]:
        row = {"count": cnt}
        outs = {}
        for mode, name in ((0, "one_thread"), (1, other)):
            eng.set_option(option, mode)
            buf = cts[: cnt * EB]
            out = torch.empty(cnt * EB, dtype=torch.uint8, device=dev)

            def call():
                if fixed:
                    eng.make_l2_batch(buf, out=out)
                else:
                    eng.pair_batch(buf, cts2[: cnt * EB], out=out)

            call()
            best = None
            for _ in range(args.reps):
                eng.timing_reset()
                call()
                k = eng.timing_get(prefix or ("k_pair_duo" if mode else "k_miller"))[0]
                best = k if best is None else min(best, k)
            outs[mode] = out
            row[name] = {"kernel_ms": best, "pairings_per_s": cnt / (best * 1e-3),
                         "imad_frac": cnt * prods[mode] / (best * 1e-3) / peak}
        row["bytes_equal"] = bool((outs[0] == outs[1]).all().item())
        row["speedup"] = row["one_thread"]["kernel_ms"] / row[other]["kernel_ms"]
        res["sizes"].append(row)
    print(json.dumps(res, indent=1))
    eng.close()


if __name__ == "__main__":
    main()
