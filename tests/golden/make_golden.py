#!/usr/bin/env python
"""Generates tests/golden/kb*.json from the CPU oracle (oracle/bgn_oracle.py).

The reference (sachaservan/bgn) holds no golden vectors and cannot be built in
this image (no Go, no libpbc: SURVEY.md 8(c)), so these vectors are DERIVED FROM
THE RESTATED SPEC, not PBC-produced ("parity unpinned").  They pin (i) the
oracle against accidental change and (ii) the CUDA path against the oracle.

    python tests/golden/make_golden.py            # rewrites the fixtures

Deterministic: fixed seeds per key size.  Byte strings are PBC element_to_bytes
layout (x||y / re||im, big-endian, width ceil(bits(p)/8)), hex-encoded.
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import bgn_oracle as O  # noqa: E402

SEEDS = {64: 0xB6A00064, 128: 0xB6A00001, 256: 0xB6A00256, 512: 0xB6A00002, 1024: 0xB6A00003}
MSG_SPACE = {64: 1021, 128: 1021, 256: 10000, 512: 1 << 20, 1024: 1021}


def hx(b: bytes) -> str:
    return b.hex()


def make(kb: int, small: bool) -> dict:
    pk, sk = O.keygen(kb, MSG_SPACE[kb], seed=SEEDS[kb])
    par, p, n = pk.params, pk.params.p, pk.params.n
    rng = random.Random(SEEDS[kb] ^ 0x5EED)
    g1b = lambda P: hx(O.g1_to_bytes(P, par))  # noqa: E731
    gtb = lambda a: hx(O.gt_to_bytes(a, par))  # noqa: E731
    q2 = n // sk.key
    out = {
        "note": "derived from oracle/bgn_oracle.py (restated PBC type-a1 definitions); not PBC-produced",
        "key_bits": kb, "p": hex(p), "n": hex(n), "l": par.l, "q1": hex(sk.key), "q2": hex(q2),
        "P": g1b(pk.P), "Q": g1b(pk.Q), "msg_space": pk.msg_space, "coord_bytes": par.coord_bytes,
        "pbc_params": par.pbc_string(),
    }
    # ---- Encrypt (bgn.go:340-353), incl. x = 0, r = 0 (-> O), negative digits (poly.go:17-21)
    xs = [0, 1, -1, 2, 0, 1021, -77, 5, (1 << 40) + 3]
    rs = [0, 0, rng.randrange(n), rng.randrange(n), rng.randrange(n), n - 1, 1, rng.randrange(n), rng.randrange(n)]
    if small:
        xs, rs = xs[:6], rs[:6]
    enc = []
    for x, r in zip(xs, rs):
        c = O.encrypt_with_randomness(pk, abs(x), r).C
        if x < 0:
            c = O.g1_neg(c, p)
        enc.append(c)
    out["encrypt"] = {"x": xs, "r": [hex(r) for r in rs], "out": [g1b(c) for c in enc]}
    # ---- G1 add / sub / neg incl. O operands, doubling and inverse pairs (bgn.go:482, 419, 436-439)
    A = [enc[1], enc[2], None, enc[3], enc[3], enc[3], None]
    Bv = [enc[3], enc[4], enc[2], None, enc[3], O.g1_neg(enc[3], p), None]
    out["g1_add"] = {"a": [g1b(a) for a in A], "b": [g1b(b) for b in Bv],
                     "out": [g1b(O.g1_add(a, b, p)) for a, b in zip(A, Bv)]}
    out["g1_sub"] = {"a": [g1b(a) for a in A], "b": [g1b(b) for b in Bv],
                     "out": [g1b(O.g1_add(a, O.g1_neg(b, p), p)) for a, b in zip(A, Bv)]}
    out["g1_neg"] = {"a": [g1b(a) for a in A], "out": [g1b(O.g1_neg(a, p)) for a in A]}
    # ---- MultConst on L1 (bgn.go:258)
    ks = [0, 1, 2, 3, 1000003, rng.randrange(n)]
    pts = [enc[2], enc[2], enc[3], None, enc[4], enc[3]]
    if small:
        ks, pts = ks[:5], pts[:5]
    kw = (n.bit_length() + 7) // 8
    out["g1_mulconst"] = {"a": [g1b(a) for a in pts], "k": [hex(k) for k in ks], "kbytes": kw,
                          "out": [g1b(O.g1_mul(k, a, p)) for k, a in zip(ks, pts)]}
    # ---- Mult = pairing (bgn.go:294-314), makeL2 (bgn.go:316-321)
    pa = [pk.P, enc[2], enc[3], None, enc[4]]
    pb = [pk.P, enc[3], enc[4], enc[2], None]
    if small:
        pa, pb = pa[:3] + [None], pb[:3] + [enc[2]]
    pr = [O.pairing(a, b, par) for a, b in zip(pa, pb)]
    out["pair"] = {"a": [g1b(a) for a in pa], "b": [g1b(b) for b in pb], "out": [gtb(e) for e in pr]}
    ml = [enc[2], None] if small else [enc[2], enc[4], None]
    out["make_l2"] = {"a": [g1b(a) for a in ml], "out": [gtb(O.pairing(a, pk.P, par)) for a in ml]}
    # ---- GT mul / div / inv / pow (bgn.go:460, 397, 277)
    ga, gb = [pr[0], pr[1], pr[2], O.GT_ONE], [pr[1], pr[2], pr[2], pr[0]]
    out["gt_mul"] = {"a": [gtb(a) for a in ga], "b": [gtb(b) for b in gb],
                     "out": [gtb(O.fp2_mul(a, b, p)) for a, b in zip(ga, gb)]}
    out["gt_div"] = {"a": [gtb(a) for a in ga], "b": [gtb(b) for b in gb],
                     "out": [gtb(O.fp2_mul(a, O.fp2_inv(b, p), p)) for a, b in zip(ga, gb)]}
    out["gt_inv"] = {"a": [gtb(a) for a in ga], "out": [gtb(O.fp2_inv(a, p)) for a in ga]}
    gk = [0, 1, 2, 77777, rng.randrange(n)]
    gp = [pr[0], pr[1], pr[2], pr[0], pr[1]]
    out["gt_pow"] = {"a": [gtb(a) for a in gp], "k": [hex(k) for k in gk], "kbytes": kw,
                     "out": [gtb(O.fp2_pow(a, k, p)) for a, k in zip(gp, gk)]}
    # ---- MultPoly (poly.go:123-156): d1=2/3, d2=3/4 with an O coefficient
    d1, d2 = (2, 3) if small else (3, 4)
    c1 = [O.encrypt_with_randomness(pk, x, rng.randrange(n)).C for x in [1, 0, 2][:d1]]
    c2 = [O.encrypt_with_randomness(pk, x, rng.randrange(n)).C for x in [2, 1, 1, 0][:d2]]
    c2[1] = None  # deterministic encryption of 0 is O
    ct1 = O.PolyCiphertext([O.Ciphertext(c, False) for c in c1], d1, 0, False)
    ct2 = O.PolyCiphertext([O.Ciphertext(c, False) for c in c2], d2, 0, False)
    mp = O.mult_poly(pk, ct1, ct2)
    out["multpoly"] = {"d1": d1, "d2": d2, "c1": [g1b(c) for c in c1], "c2": [g1b(c) for c in c2],
                       "out": [gtb(c.C) for c in mp.coefficients]}
    # ---- L2 sum (AddPoly fold, poly.go:191-204 -> bgn.go:460): nterms x ncoeff
    pool = [e for e in pr if e != O.GT_ONE] + [c.C for c in mp.coefficients]
    nterms, ncoeff = 5, 3
    terms = [pool[(t * ncoeff + c) % len(pool)] for t in range(nterms) for c in range(ncoeff)]
    red = []
    for c in range(ncoeff):
        acc = O.GT_ONE
        for t in range(nterms):
            acc = O.fp2_mul(acc, terms[t * ncoeff + c], p)
        red.append(acc)
    out["l2_sum"] = {"nterms": nterms, "ncoeff": ncoeff, "in": [gtb(a) for a in terms], "out": [gtb(a) for a in red]}
    # ---- Decrypt (bgn.go:218-250, gsbs.go:54-106): GT^q1 and plaintext recovery
    T = pk.msg_space
    bound = O._bsgs_bound(T)
    ms = [0, 1, -1, 2, T - 1, -(T - 1), bound * bound + bound + 2, bound * bound + bound + 3, -5, 17]
    if small:
        ms = ms[:8]
    ePP = O.pairing(pk.P, pk.P, par)
    l2 = []
    for i, m in enumerate(ms):
        # e(P,P)^m * e(Q,Q)^s: an L2 ciphertext of m with a blinding component
        s = 0 if i % 2 == 0 else rng.randrange(n)
        blind = O.fp2_pow(O.pairing(pk.Q, pk.Q, par), s, p) if s else O.GT_ONE
        l2.append(O.fp2_mul(O.fp2_pow(ePP, m % n, p), blind, p))
    O.setup_decryption(pk, sk)
    vals, status = [], []
    for c in l2:
        try:
            vals.append(O.decrypt(pk, sk, O.Ciphertext(c, True)))
            status.append(0)
        except O.DLError:
            vals.append(0)
            status.append(1)
    out["decrypt_l2"] = {"in": [gtb(c) for c in l2], "m": ms, "out": vals, "status": status,
                         "csk": [gtb(O.fp2_pow(c, sk.key, p)) for c in l2]}
    l1m = [0, 1, -1, 3, 1000, -1000][: (4 if small else 6)]
    l1 = []
    for i, m in enumerate(l1m):
        c = O.encrypt_with_randomness(pk, abs(m), rng.randrange(n) if i % 2 else 0).C
        l1.append(O.g1_neg(c, p) if m < 0 else c)
    vals = [O.decrypt(pk, sk, O.Ciphertext(c, False)) for c in l1]
    out["decrypt_l1"] = {"in": [g1b(c) for c in l1], "m": l1m, "out": vals, "status": [0] * len(vals)}
    # ---- sections added later draw from their own generator so the ones above never change
    rng2 = random.Random(SEEDS[kb] ^ 0xF1F2)
    # non-deterministic mode, injected r (bgn.go:264-268, 488-495 / 283-287, 306-310, 469-474):
    # + r*Q on level 1 (incl. O as the value and r = 0), * e(Q,Q)^r on level 2
    nk = 3 if small else 5
    ba = ([enc[1], None, enc[3], enc[2], enc[4]])[:nk]
    br = ([rng2.randrange(n), rng2.randrange(n), 0, n - 1, rng2.randrange(n)])[:nk]
    out["g1_blind"] = {"a": [g1b(a) for a in ba], "r": [hex(r) for r in br],
                       "out": [g1b(O.g1_add(a, O.g1_mul(r, pk.Q, p), p)) for a, r in zip(ba, br)]}
    eQQ = O.pairing(pk.Q, pk.Q, par)
    bg = ([pr[0], O.GT_ONE, pr[1], pr[2], pr[1]])[:nk]
    out["gt_blind"] = {"a": [gtb(a) for a in bg], "r": [hex(r) for r in br],
                       "out": [gtb(O.fp2_mul(a, O.fp2_pow(eQQ, r, p), p)) for a, r in zip(bg, br)]}
    # MultConstPoly / EvalPoly / MakePolyL2 (poly.go:58-120, 159-163) on two polynomials of 3 slots
    pcount, pd = 2, 3
    polys = []
    for u in range(pcount):
        coeffs = [[1, -1, 0], [0, 1, 1]][u]
        cs = []
        for i, x in enumerate(coeffs):
            r = 0 if (u == 1 and i == 0) else rng2.randrange(n)  # one O coefficient
            c = O.encrypt_with_randomness(pk, abs(x), r).C
            cs.append(O.Ciphertext(O.g1_neg(c, p) if x < 0 else c, False))
        polys.append(O.PolyCiphertext(cs, pd, 0, False))
    polys2 = [O.make_poly_l2(pk, ct) for ct in polys]  # pd + 1 slots
    out["make_poly_l2"] = {"d": pd, "count": pcount, "in": [g1b(c.C) for ct in polys for c in ct.coefficients],
                           "out": [gtb(c.C) for ct in polys2 for c in ct.coefficients]}
    for name, cts, dd in (("l1", polys, pd), ("l2", polys2, pd + 1)):
        enc_el = (lambda c: gtb(c.C)) if name == "l2" else (lambda c: g1b(c.C))
        cases = []
        for constant in ((4.12, -2.0) if not small else (5.0,)):
            digits = pk.new_unbalanced_plaintext(abs(constant)).coefficients
            res = [O.mult_const_poly(pk, ct, constant) for ct in cts]
            cases.append({"constant": constant, "digits": digits, "negate": constant < 0,
                          "out": [enc_el(c) for ct in res for c in ct.coefficients]})
        out["multconstpoly_" + name] = {"d": dd, "count": pcount, "in": [enc_el(c) for ct in cts for c in ct.coefficients],
                                        "cases": cases}
        out["evalpoly_" + name] = {"d": dd, "count": pcount, "base": pk.poly_base,
                                   "in": [enc_el(c) for ct in cts for c in ct.coefficients],
                                   "out": [enc_el(O.eval_poly(pk, ct)) for ct in cts]}
    if not small:
        out["nondet_poly"] = nondet_poly_section(pk, kb)
    return out


def nondet_poly_section(pk, kb: int) -> dict:
    """Round 2: the polynomial operations on a Deterministic=false key.  Every newCryptoRandom(pk.N) draw
    of the reference's control flow (poly.go:45-55, 97-118, 140-152, 191-204 -> bgn.go:260-269, 279-288,
    302-311, 421-432, 466-474, 488-495) is taken from the recorded stream `draws`, in sequential program
    order; the fixture records the stream, the inputs and each operation's output bytes and draw count."""
    par, p, n = pk.params, pk.params.p, pk.params.n
    nd = O.PublicKey(par, pk.P, pk.Q, pk.msg_space, deterministic=False, poly_base=pk.poly_base,
                     fp_scale_base=pk.fp_scale_base, fp_precision=pk.fp_precision)
    rng = random.Random(SEEDS[kb] ^ 0xD3A5)
    g1b = lambda P: hx(O.g1_to_bytes(P, par))  # noqa: E731

    def enc(coeffs):
        cs = []
        for x in coeffs:
            c = O.encrypt_with_randomness(nd, abs(x), rng.randrange(n)).C
            cs.append(O.Ciphertext(O.g1_neg(c, p) if x < 0 else c, False))
        return O.PolyCiphertext(cs, len(coeffs), 0, False)

    a, b = enc([1, -1, 1]), enc([0, 1])
    draws = [rng.randrange(n) for _ in range(64)]
    ops = {}

    def run(name, fn):
        it = iter(draws)
        used = [0]

        def src():
            used[0] += 1
            return next(it)

        nd.rand_source = src
        res = fn()
        nd.rand_source = None
        if isinstance(res, O.Ciphertext):
            ops[name] = {"draws_used": used[0], "L2": res.L2, "out": [hx(O.ct_bytes(nd, res))]}
        else:
            ops[name] = {"draws_used": used[0], "L2": res.L2, "degree": res.degree, "scale_factor": res.scale_factor,
                         "out": [hx(O.ct_bytes(nd, c)) for c in res.coefficients]}
        return res

    m = run("mult_poly", lambda: O.mult_poly(nd, a, b))
    run("add_poly", lambda: O.add_poly(nd, a, b))
    run("sub_poly", lambda: O.sub_poly(nd, a, b))
    run("neg_poly", lambda: O.neg_poly(nd, a))
    run("neg_poly_l2", lambda: O.neg_poly(nd, m))
    run("mult_const_poly_neg2", lambda: O.mult_const_poly(nd, a, -2.0))
    run("make_poly_l2", lambda: O.make_poly_l2(nd, b))
    run("add_poly_mixed_levels", lambda: O.add_poly(nd, m, b))
    run("eval_poly", lambda: O.eval_poly(nd, a))
    return {"a": [g1b(c.C) for c in a.coefficients], "b": [g1b(c.C) for c in b.coefficients],
            "draws": [hex(x) for x in draws], "ops": ops}


if __name__ == "__main__":
    only_new = "--only-new" in sys.argv  # add the sections an existing fixture lacks, keep the rest byte for byte
    for kb in (64, 128, 256, 512, 1024):
        path = os.path.join(HERE, "kb%d.json" % kb)
        if only_new and os.path.exists(path):
            with open(path) as f:
                d = json.load(f)
            if "nondet_poly" not in d and kb != 1024:
                pk, _ = O.keygen(kb, MSG_SPACE[kb], seed=SEEDS[kb])
                d["nondet_poly"] = nondet_poly_section(pk, kb)
        else:
            d = make(kb, small=(kb == 1024))
        with open(path, "w") as f:
            json.dump(d, f, indent=1)
        print("wrote", path, os.path.getsize(path), "bytes")
