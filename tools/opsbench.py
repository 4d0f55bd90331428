#!/usr/bin/env python
"""Per-operation throughput at keyBits=512 on cuda:0 with inputs resident in HBM: the BASELINE.json
configs other than the headline one (config 2: Encrypt + EAdd of 2^16 x 11 coefficients, config 4:
Decrypt of 2^14 level-2 ciphertexts over T = 2^20), MultConst, and each one's dominant kernel as a
fraction of the IMAD.WIDE peak (executed modmuls from bgn_b200.workmodel).  Prints one JSON object.
usage: tools/opsbench.py [--plaintexts 65536] [--decrypts 16384]"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from bgn_b200 import Engine, bench_imad_peak, workmodel  # noqa: E402

D = 11
T = 1 << 20


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--plaintexts", type=int, default=1 << 16)
    ap.add_argument("--decrypts", type=int, default=1 << 14)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--key-bits", type=int, default=512)
    ap.add_argument("--emults", type=int, default=0, help="also time MultPoly of this many pairs")
    ap.add_argument("--d", type=int, default=D, help="coefficient slots per polynomial for --emults")
    ap.add_argument("--enc-window", type=int, default=0, help="fixed-base window of Q: 0 (the default choice), 8, 16 .. 24 bits")
    args = ap.parse_args()
    with open(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "kb%d.json" % args.key_bits)) as f:
        g = json.load(f)
    p, n, l, q1 = int(g["p"], 16), int(g["n"], 16), g["l"], int(g["q1"], 16)
    eng = Engine(p, n, l, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), device=0)
    eng.set_option("enc_window", args.enc_window)
    if args.enc_window == 0:
        args.enc_window = workmodel.enc_window_auto(eng.scalar_bytes, eng.limbs)
    L, EB, SB = eng.limbs, eng.elem_bytes, eng.scalar_bytes
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    ms, ipt = bench_imad_peak(0, 4096, 148 * 8, 256)
    peak = 148 * 8 * 256 * ipt / (ms * 1e-3)
    ppm = workmodel.products_per_modmul(L)
    eng.timing_enable(True)
    res = {"key_bits": args.key_bits, "limbs": L, "imad_wide_peak_T": peak / 1e12, "ops": {}}

    def timed(fn, reps=args.reps):
        fn()  # warm-up
        best, kern = None, None
        for _ in range(reps):
            eng.timing_reset()
            fn()
            t = eng.timing_last_call()
            if best is None or t < best:
                best = t
                kern = {k: eng.timing_get(k)[0] for k in
                        ("k_encrypt", "k_normalize", "k_g1_add", "k_g1_mulvar", "k_gt_pow", "k_bsgs_lookup", "k_miller", "k_miller_fixed", "k_pair_duo", "k_miller_split",
                         "k_dec_lucas", "k_gt_blind", "k_g1_affadd", "k_g1_polyconv", "k_gt_polyconv", "k_gt_mul",
                         "k_g1_from_bytes", "k_g1_to_bytes", "k_fp2_from_bytes", "k_fp2_to_bytes")}
        kern["k_miller"] -= kern["k_miller_fixed"] + kern["k_miller_split"]  # timing_get matches by prefix
        return best, {k: v for k, v in kern.items() if v > 1e-9}

    def entry(name, units, unit_name, ms_call, kern, modmuls_per_unit=None, dominant=None):
        e = {"units": units, "unit": unit_name, "ms": ms_call, "per_s": units / (ms_call * 1e-3), "kernel_ms": kern}
        if modmuls_per_unit and dominant and dominant in kern:
            e["dominant_kernel"] = dominant
            e["modmuls_per_unit"] = modmuls_per_unit
            e["imad_frac"] = units * modmuls_per_unit * ppm / (kern[dominant] * 1e-3) / peak
        res["ops"][name] = e

    # ---- config 2: Encrypt of 2^16 plaintexts x 11 balanced base-3 digits, r < n
    cnt = args.plaintexts * D
    digits = torch.randint(-1, 2, (cnt,), generator=gen, device=dev, dtype=torch.int64)
    r = torch.randint(0, 256, (cnt, SB), generator=gen, device=dev, dtype=torch.uint8)
    r[:, 0] &= 0x3F
    r = r.reshape(-1)
    out = torch.empty(cnt * EB, dtype=torch.uint8, device=dev)
    t, k = timed(lambda: eng.encrypt_batch(digits, r, out=out))
    entry("encrypt", cnt, "coefficient encryptions", t, k, workmodel.encrypt_products(n, SB, args.enc_window, L) / ppm, "k_encrypt")
    res["ops"]["encrypt"]["window_bits"] = args.enc_window
    res["ops"]["encrypt"]["plaintexts_per_s"] = args.plaintexts / (t * 1e-3)

    if args.emults:
        d, m = args.d, args.emults
        c1, c2 = out[: m * d * EB], out[m * d * EB: 2 * m * d * EB]
        o_em = torch.empty(m * 2 * d * EB, dtype=torch.uint8, device=dev)
        t, k = timed(lambda: eng.multpoly_batch(c1, d, c2, d, m, out=o_em))
        entry("emult_d%d" % d, m, "MultPoly products (%d pairings each)" % (d * d), t, k,
              workmodel.miller_unit_products(p, n, l, d, d) / ppm, "k_miller")
        res["ops"]["emult_d%d" % d]["pairings_per_s"] = m * d * d / (t * 1e-3)

    # ---- config 2: EAdd = pairwise AddPoly of the two halves
    half = cnt // 2
    a, b = out[: half * EB], out[half * EB:]
    o2 = torch.empty(half * EB, dtype=torch.uint8, device=dev)
    t, k = timed(lambda: eng.g1_add_batch(a, b, out=o2))
    entry("eadd_l1", half, "coefficient additions", t, k, 11, "k_g1_add")

    # ---- MultConst on level 1 by 16-bit constants
    m = min(half, 1 << 16)
    kk = torch.randint(0, 256, (m, 2), generator=gen, device=dev, dtype=torch.uint8).reshape(-1)
    o3 = torch.empty(m * EB, dtype=torch.uint8, device=dev)
    t, k = timed(lambda: eng.g1_mulconst_batch(out[: m * EB], kk, 2, out=o3))
    entry("multconst_l1_16bit", m, "scalar multiplications", t, k)

    # ---- config 4: Decrypt of level-2 ciphertexts e(E(a), E(b)), |a b| < T, half negative, 1 % zeros
    nd = args.decrypts
    av = torch.randint(1, 1 << 10, (nd,), generator=gen, device=dev, dtype=torch.int64)
    bv = torch.randint(-(1 << 10) + 1, 1 << 10, (nd,), generator=gen, device=dev, dtype=torch.int64)
    bv[:: 100] = 0
    rr = torch.randint(0, 256, (nd, SB), generator=gen, device=dev, dtype=torch.uint8)
    rr[:, 0] &= 0x3F
    ca = eng.encrypt_batch(av, rr.reshape(-1))
    cb = eng.encrypt_batch(bv, rr.flip(0).reshape(-1))
    t, k = timed(lambda: eng.pair_batch(ca, cb))
    if "k_pair_duo" in k:
        entry("pair_single", nd, "pairings (two warps per 32 pairings)", t, k, workmodel.pair_duo_products(p, n, l) / ppm, "k_pair_duo")
    else:
        entry("pair_single", nd, "pairings (unshared, one team of 1)", t, k,
              workmodel.miller_unit_products(p, n, l, 1, 1) / ppm, "k_miller")
    l2 = eng.pair_batch(ca, cb)
    # the same kernels on a batch that gives every scheduler two warps (2^17 pairings)
    rep = (1 << 17) // nd
    if rep > 1:
        ca_big, cb_big = ca.repeat(rep), cb.repeat(rep)
        t, k = timed(lambda: eng.pair_batch(ca_big, cb_big), reps=1)
        entry("pair_single_2e17", nd * rep, "pairings (unshared)", t, k,
              workmodel.miller_unit_products(p, n, l, 1, 1) / ppm, "k_miller")
        t, k = timed(lambda: eng.make_l2_batch(ca_big), reps=1)
        entry("make_l2_2e17", nd * rep, "pairings with P (line table)", t, k,
              workmodel.miller_fixed_products(p, n, l) / ppm, "k_miller_fixed")
        del ca_big, cb_big
    eng.set_secret(q1, T)
    vals = {}

    def dec():
        vals["v"], vals["s"] = eng.decrypt_batch(l2, True)

    t, k = timed(dec)
    if "k_dec_lucas" in k:
        entry("decrypt_l2", nd, "decryptions (T = 2^20)", t, k, workmodel.dec_lucas_modmuls(q1), "k_dec_lucas")
    else:
        entry("decrypt_l2", nd, "decryptions (T = 2^20)", t, k, workmodel.gt_pow_modmuls(q1), "k_gt_pow")
    ok = bool((vals["v"] == av * bv).all().item()) and not bool(vals["s"].any().item())
    res["ops"]["decrypt_l2"]["plaintexts_match"] = ok
    t, k = timed(lambda: eng.gt_pow_secret_batch(l2))
    entry("gt_pow_q1", nd, "GT exponentiations by q1", t, k, workmodel.gt_pow_modmuls(q1), "k_gt_pow")
    dl1 = eng.encrypt_batch(torch.randint(-1000, 1000, (nd,), generator=gen, device=dev, dtype=torch.int64),
                            rr.reshape(-1))
    t, k = timed(lambda: eng.decrypt_batch(dl1, False))
    entry("decrypt_l1", nd, "decryptions of level-1 ciphertexts", t, k, workmodel.miller_fixed_pair_products(p, n, l) / ppm,
          "k_miller_fixed")
    # ---- larger decrypt batch (the kernel's rate once every scheduler holds two warps)
    big = 1 << 18
    l2big = l2.repeat(big // nd) if big > nd else l2
    t, k = timed(lambda: eng.decrypt_batch(l2big, True))
    dom = "k_dec_lucas" if "k_dec_lucas" in k else "k_gt_pow"
    entry("decrypt_l2_2e18", l2big.numel() // EB, "decryptions (T = 2^20)", t, k,
          workmodel.dec_lucas_modmuls(q1) if dom == "k_dec_lucas" else workmodel.gt_pow_modmuls(q1), dom)

    # ---- non-deterministic mode: re-randomisation of level-1 / level-2 coefficients
    nb = max(nd, min(1 << 18, cnt) // nd * nd)
    rb = torch.randint(0, 256, (nb, SB), generator=gen, device=dev, dtype=torch.uint8)
    rb[:, 0] &= 0x3F
    rb = rb.reshape(-1)
    ob = torch.empty(nb * EB, dtype=torch.uint8, device=dev)
    src1 = out[: nb * EB] if nb <= cnt else out.repeat((nb + cnt - 1) // cnt)[: nb * EB]
    t, k = timed(lambda: eng.g1_blind_batch(src1, rb, out=ob))
    entry("blind_l1", nb, "level-1 re-randomisations (+ r*Q)", t, k,
          (workmodel.encrypt_products(n, SB, args.enc_window, L, 0.0) + workmodel.madd_products(L)) / ppm, "k_encrypt")
    l2b = l2.repeat(nb // nd)
    t, k = timed(lambda: eng.gt_blind_batch(l2b, rb, out=ob))
    entry("blind_l2", nb, "level-2 re-randomisations (* e(Q,Q)^r)", t, k, 3 * SB * 255.0 / 256.0, "k_gt_blind")

    # ---- polynomial helpers on 2^16 polynomials of 11 slots
    npoly = min(args.plaintexts, 1 << 16)
    polys = out[: npoly * D * EB]
    digits3 = [2, 1, 0, 2, 1]
    t, k = timed(lambda: eng.multconstpoly_batch(polys, D, False, digits3, False, npoly))
    entry("multconstpoly_l1", npoly, "MultConstPoly (11 slots x 5 digits)", t, k)
    t, k = timed(lambda: eng.evalpoly_batch(polys, D, False, 3, npoly))
    entry("evalpoly_l1", npoly, "EvalPoly (11 slots)", t, k)
    nl2 = 1 << 12
    t, k = timed(lambda: eng.make_poly_l2_batch(polys[: nl2 * D * EB], D, nl2))
    entry("make_poly_l2", nl2, "MakePolyL2 (11 pairings with P each)", t, k,
          D * workmodel.miller_fixed_products(p, n, l) / ppm, "k_miller_fixed")
    print(json.dumps(res, indent=1))
    eng.close()


if __name__ == "__main__":
    main()
