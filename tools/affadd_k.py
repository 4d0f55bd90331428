#!/usr/bin/env python
"""How far does the affine addition with a shared inversion (k_g1_affadd) get towards the IMAD.WIDE peak when a
thread shares its inversion between many additions?  Device-resident handles, 2^18 .. 2^22 additions, elements
per inversion forced through BGN_NORM_PER_THREAD / BGN_NORM_THREADS (read at context creation).  One JSON object."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    per, threads = int(sys.argv[1]), int(sys.argv[2])
    os.environ["BGN_NORM_PER_THREAD"] = str(per)
    os.environ["BGN_NORM_THREADS"] = str(threads)
    from bgn_b200 import Engine, bench_imad_peak, workmodel
    with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb512.json")) as f:
        g = json.load(f)
    p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    eng = Engine(p, n, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=0)
    EB, SB, L = eng.elem_bytes, eng.scalar_bytes, eng.limbs
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    ms, ipt = bench_imad_peak(0, 4096, 148 * 8, 256)
    peak = 148 * 8 * 256 * ipt / (ms * 1e-3)
    rows = []
    eng.timing_enable(True)
    for lg in (18, 20, 21, 22):
        cnt = 1 << lg
        xs = torch.randint(-1, 2, (2 * cnt,), generator=gen, device=dev, dtype=torch.int64)
        r = torch.randint(0, 256, (2 * cnt, SB), generator=gen, device=dev, dtype=torch.uint8)
        r[:, 0] &= 0x3F
        c = eng.encrypt_batch(xs, r.reshape(-1))
        hA = eng.import_batch(1, c[: cnt * EB])
        hB = eng.import_batch(1, c[cnt * EB:])
        del c, r
        hR = eng.g1_add_h(hA, hB)
        best = None
        for _ in range(3):
            eng.timing_reset()
            eng.g1_add_h(hA, hB, out=hR)
            k = eng.timing_get("k_g1_affadd")[0]
            best = k if best is None else min(best, k)
        rows.append({"additions": cnt, "per_thread_min": per, "threads_aim": threads, "k_g1_affadd_ms": best,
                     "additions_per_s": cnt / (best * 1e-3),
                     "imad_frac": cnt * workmodel.affadd_products(L) / (best * 1e-3) / peak})
        print(json.dumps(rows[-1]), file=sys.stderr, flush=True)
        for h in (hA, hB, hR):
            h.free()
    print(json.dumps({"imad_wide_peak_T": peak / 1e12, "rows": rows}, indent=1))
    eng.close()


if __name__ == "__main__":
    main()
