#!/usr/bin/env python
"""Encrypt (BASELINE config 2: 720 896 coefficient encryptions, keyBits = 512) against the window width of Q's
fixed-base table -- 16 / 18 / 20 / 22 / 24 bits -- and the form of its points: twisted Edwards (the default:
8 products per table point) or Weierstrass (complete mixed Jacobian addition: 8 products + 3 squarings).  Table
size and build time, kernel time, IMAD fraction, and the same at keyBits = 1024 for 16 / 18 / 20.  One JSON object."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bgn_b200 import Engine, bench_imad_peak, workmodel  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    ms, ipt = bench_imad_peak(0, 4096, 148 * 8, 256)
    peak = 148 * 8 * 256 * ipt / (ms * 1e-3)
    res = {"imad_wide_peak_T": peak / 1e12, "rows": []}
    for kb, widths, cnt in ((512, (16, 18, 20, 22, 24), 720896), (1024, (16, 18, 20), 1 << 18)):
        with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb%d.json" % kb)) as f:
            g = json.load(f)
        p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
        eng = Engine(p, n, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=0)
        EB, SB, L = eng.elem_bytes, eng.scalar_bytes, eng.limbs
        gen = torch.Generator(device=dev)
        gen.manual_seed(5)
        xs = torch.randint(-1, 2, (cnt,), generator=gen, device=dev, dtype=torch.int64)
        r = torch.randint(0, 256, (cnt, SB), generator=gen, device=dev, dtype=torch.uint8)
        r[:, 0] &= 0x3F
        r = r.reshape(-1)
        out = torch.empty(cnt * EB, dtype=torch.uint8, device=dev)
        eng.timing_enable(True)
        ref = None
        for edw, bits in [(e, b) for e in (1, 0) for b in widths]:
            eng.set_option("enc_edwards", edw)
            eng.set_option("enc_window", bits)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            eng.encrypt_batch(xs[:32], r[: 32 * SB])  # builds the table
            torch.cuda.synchronize()
            build_s = time.perf_counter() - t0
            eng.encrypt_batch(xs, r, out=out)
            best = None
            for _ in range(3):
                eng.timing_reset()
                eng.encrypt_batch(xs, r, out=out)
                t, k = eng.timing_last_call(), eng.timing_get("k_encrypt")[0]
                if best is None or t < best[0]:
                    best = (t, k)
            if ref is None:
                ref = out.clone()
            row = {"key_bits": kb, "edwards": bool(edw), "window_bits": bits,
                   "table_GB": workmodel.enc_table_bytes(SB, bits, L, bool(edw)) / 1e9,
                   "table_build_s": build_s, "encryptions": cnt, "call_ms": best[0], "k_encrypt_ms": best[1],
                   "per_s": cnt / (best[0] * 1e-3),
                   "imad_frac": cnt * workmodel.encrypt_products(n, SB, bits, L, edwards=bool(edw)) / (best[1] * 1e-3) / peak,
                   "bytes_equal_first_row": bool((out == ref).all().item())}
            res["rows"].append(row)
            print(json.dumps(row), file=sys.stderr, flush=True)
        eng.close()
        del xs, r, out, ref
        torch.cuda.empty_cache()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
