"""Shared body of the non-deterministic-mode parity tests (CPU: oracle-backed stand-in engine; GPU: the
CUDA engine).  Every polynomial operation of the mirror is run on a Deterministic=False key whose
`rand_source` replays a seeded stream, and compared byte for byte with the oracle's literal transcription
of poly.go fed by the SAME stream -- so the mirror's one-blind-per-slot algebra (bgn_b200/bgn.py) is checked
against the reference's draw-per-operation control flow (poly.go:45-55, 97-118, 140-152, 191-204 ->
bgn.go:260-269, 279-288, 302-311, 421-432, 466-474, 488-495), including that both consume the same number
of draws."""
import random

import numpy as np

from oracle import bgn_oracle as O

from bgn_b200.bgn import PolyCiphertextBatch, PublicKey


class Stream:
    """newCryptoRandom stand-in: a seeded PRNG that counts its draws"""

    def __init__(self, seed, n):
        self.rng, self.n, self.count = random.Random(seed), n, 0

    def __call__(self):
        self.count += 1
        return self.rng.randrange(self.n)


def nondet_pair(pk0: PublicKey, opk: O.PublicKey):
    pk = PublicKey(pk0.engine.p, pk0.N, pk0.engine.l, pk0.P, pk0.Q, pk0.MsgSpace, Deterministic=False,
                   engine=pk0.engine)
    pk._secret_set, pk._secret_key = pk0._secret_set, pk0._secret_key
    ond = O.PublicKey(opk.params, opk.P, opk.Q, opk.msg_space, deterministic=False)
    ond.table_g1, ond.table_gt, ond.tables_computed = opk.table_g1, opk.table_gt, opk.tables_computed
    return pk, ond


def run_poly_cases(pk0: PublicKey, opk: O.PublicKey, sk=None, osk=None, values=(9.13, 4.0, 1 / 3)):
    pk, ond = nondet_pair(pk0, opk)
    seed = [100]

    def both(fm, fo):
        """run the mirror op and the oracle op on identical streams -> (mirror result, oracle result)"""
        seed[0] += 1
        sm, so = Stream(seed[0], pk.N), Stream(seed[0], pk.N)
        pk.rand_source, ond.rand_source = sm, so
        rm, ro = fm(), fo()
        assert sm.count == so.count, "mirror drew %d scalars, the reference's control flow draws %d" % (sm.count, so.count)
        return rm, ro

    def same(ct, oct):
        assert (ct.Degree, ct.ScaleFactor, ct.L2) == (oct.degree, oct.scale_factor, oct.L2)
        assert ct.CoeffBytes() == O.poly_ct_bytes(ond, oct)

    def enc(v, negate_some=False):
        pt, opt = pk.NewPolyPlaintext(v), ond.new_poly_plaintext(v)
        if negate_some:  # exercise the Sub(encryptZero(), Encrypt(|c|)) branch of EncryptPoly with its own draw
            pt.Coefficients = [-c if i % 2 else c for i, c in enumerate(pt.Coefficients)]
            opt.coefficients = [-c if i % 2 else c for i, c in enumerate(opt.coefficients)]
        c, oc = both(lambda: pk.EncryptPoly(pt), lambda: O.encrypt_poly(ond, opt))
        same(c, oc)
        return c, oc

    a, oa = enc(values[0], True)
    b, ob = enc(values[1])
    c, oc = enc(values[2])

    same(*both(lambda: pk.NegPoly(a), lambda: O.neg_poly(ond, oa)))
    m, om = both(lambda: pk.MultPoly(a, b), lambda: O.mult_poly(ond, oa, ob))
    same(m, om)
    # the unused top slot of MultPoly stays the GT identity (poly.go:130-137): never blinded
    eb = pk.elem_bytes
    assert m.Coefficients[-1].C == b"\x00" * (eb // 2 - 1) + b"\x01" + b"\x00" * (eb // 2)
    same(*both(lambda: pk.NegPoly(m), lambda: O.neg_poly(ond, om)))
    l2b, ol2b = both(lambda: pk.MakePolyL2(b), lambda: O.make_poly_l2(ond, ob))
    same(l2b, ol2b)
    for const in (2.0, -2.0, 4.12):
        same(*both(lambda: pk.MultConstPoly(a, const), lambda: O.mult_const_poly(ond, oa, const)))
    same(*both(lambda: pk.MultConstPoly(l2b, -5.0), lambda: O.mult_const_poly(ond, ol2b, -5.0)))
    same(*both(lambda: pk.AddPoly(a, b), lambda: O.add_poly(ond, oa, ob)))      # scale alignment + tail pass-through
    same(*both(lambda: pk.AddPoly(b, c), lambda: O.add_poly(ond, ob, oc)))
    same(*both(lambda: pk.SubPoly(a, c), lambda: O.sub_poly(ond, oa, oc)))
    same(*both(lambda: pk.AddPoly(m, c), lambda: O.add_poly(ond, om, oc)))      # level promotion through MakePolyL2
    same(*both(lambda: pk.SubPoly(m, l2b), lambda: O.sub_poly(ond, om, ol2b)))
    for poly, opoly in ((b, ob), (l2b, ol2b)):
        e, oe = both(lambda: pk.EvalPoly(poly), lambda: O.eval_poly(ond, opoly))
        assert e.C == O.ct_bytes(ond, oe) and e.L2 == oe.L2

    # ---- batch entry points against the scalar mirror on identical streams
    def batch_of(polys):
        return PolyCiphertextBatch(np.frombuffer(b"".join(p.CoeffBytes() for p in polys), dtype=np.uint8),
                                   len(polys), polys[0].Degree, polys[0].ScaleFactor, polys[0].L2)

    def cat(polys):
        return b"".join(p.CoeffBytes() for p in polys)

    def raw(x):
        return bytes(x.tobytes()) if hasattr(x, "tobytes") else bytes(x.cpu().numpy().tobytes())

    def both_mirror(fm, fo):  # both sides are the mirror: batch form vs scalar form
        seed[0] += 1
        s1 = Stream(seed[0], pk.N)
        pk.rand_source = s1
        rm = fm()
        s2 = Stream(seed[0], pk.N)
        pk.rand_source = s2
        ro = fo()
        assert s1.count == s2.count, "batch form drew %d scalars, the scalar form %d" % (s1.count, s2.count)
        return rm, ro

    def enc_padded(v, pad_to):
        pt = pk.NewPolyPlaintext(v)
        pp = type(pt)(pt.Coefficients[: pt.Degree] + [0] * (pad_to - pt.Degree), pad_to, pt.ScaleFactor, pt.params)
        pk.rand_source = Stream(1000 + pad_to + int(v * 9), pk.N)
        return pk.EncryptPoly(pp)

    ints = [enc_padded(v, 4) for v in (5.0, 7.0)]
    thirds = [enc_padded(v, 3) for v in (1 / 3, 2 / 3)]
    bi, bt = batch_of(ints), batch_of(thirds)
    r1, r2 = both_mirror(lambda: pk.MultPolyBatch(bi, bt), lambda: [pk.MultPoly(x, y) for x, y in zip(ints, thirds)])
    assert raw(r1.data) == cat(r2) and (r1.Degree, r1.ScaleFactor, r1.L2) == (r2[0].Degree, r2[0].ScaleFactor, True)
    prods = r2
    r1, r2 = both_mirror(lambda: pk.AddPolyBatch(bi, bt), lambda: [pk.AddPoly(x, y) for x, y in zip(ints, thirds)])
    assert raw(r1.data) == cat(r2) and r1.Degree == r2[0].Degree
    bp = batch_of(prods)
    r1, r2 = both_mirror(lambda: pk.AddPolyBatch(bp, bi), lambda: [pk.AddPoly(x, y) for x, y in zip(prods, ints)])
    assert raw(r1.data) == cat(r2) and r1.L2
    r1, r2 = both_mirror(lambda: pk.MultConstPolyBatch(bi, -2.0), lambda: [pk.MultConstPoly(x, -2.0) for x in ints])
    assert raw(r1.data) == cat(r2)
    r1, r2 = both_mirror(lambda: pk.MakePolyL2Batch(bi), lambda: [pk.MakePolyL2(x) for x in ints])
    assert raw(r1.data) == cat(r2)
    r1, r2 = both_mirror(lambda: pk.NegPolyBatch(bp), lambda: [pk.NegPoly(x) for x in prods])
    assert raw(r1.data) == cat(r2)
    r1, r2 = both_mirror(lambda: pk.EvalPolyBatch(bi), lambda: [pk.EvalPoly(x) for x in ints])
    assert raw(r1) == b"".join(x.C for x in r2)

    # ---- two calls on a non-deterministic key differ (fresh CSPRNG draws), and still decrypt alike
    pk.rand_source = None
    for f in (lambda: pk.MultPoly(a, b), lambda: pk.AddPoly(b, c), lambda: pk.NegPoly(b),
              lambda: pk.MultConstPoly(b, 2.0), lambda: pk.SubPoly(b, c), lambda: pk.MakePolyL2(b)):
        x, y = f(), f()
        assert x.CoeffBytes() != y.CoeffBytes()
        if sk is not None:
            assert sk.DecryptPoly(x, pk).Coefficients == sk.DecryptPoly(y, pk).Coefficients
    x, y = pk.MultPolyBatch(bi, bt), pk.MultPolyBatch(bi, bt)
    assert raw(x.data) != raw(y.data)


def run_golden_section(pk0: PublicKey, g: dict):
    """The mirror on a Deterministic=false key, replaying the fixture's recorded randomness stream, against the
    committed bytes of tests/golden/kb*.json["nondet_poly"] (made by the oracle's literal poly.go control
    flow, tests/golden/make_golden.py) -- including how many draws every operation consumes."""
    from bgn_b200.bgn import Ciphertext, PolyCiphertext
    v = g["nondet_poly"]
    pk = PublicKey(pk0.engine.p, pk0.N, pk0.engine.l, pk0.P, pk0.Q, pk0.MsgSpace, Deterministic=False, engine=pk0.engine)
    draws = [int(x, 16) for x in v["draws"]]

    def poly(hexes, l2=False):
        return PolyCiphertext([Ciphertext(bytes.fromhex(h), l2) for h in hexes], len(hexes), 0, l2)

    a, b = poly(v["a"]), poly(v["b"])
    res = {}

    def check(name, fn):
        used = [0]
        it = iter(draws)

        def src():
            used[0] += 1
            return next(it)

        pk.rand_source = src
        out = fn()
        exp = v["ops"][name]
        assert used[0] == exp["draws_used"], (name, used[0], exp["draws_used"])
        if isinstance(out, Ciphertext):
            assert (out.C.hex(), out.L2) == (exp["out"][0], exp["L2"]), name
        else:
            assert (out.Degree, out.ScaleFactor, out.L2) == (exp["degree"], exp["scale_factor"], exp["L2"]), name
            assert [c.C.hex() for c in out.Coefficients] == exp["out"], name
        res[name] = out
        return out

    m = check("mult_poly", lambda: pk.MultPoly(a, b))
    check("add_poly", lambda: pk.AddPoly(a, b))
    check("sub_poly", lambda: pk.SubPoly(a, b))
    check("neg_poly", lambda: pk.NegPoly(a))
    check("neg_poly_l2", lambda: pk.NegPoly(m))
    check("mult_const_poly_neg2", lambda: pk.MultConstPoly(a, -2.0))
    check("make_poly_l2", lambda: pk.MakePolyL2(b))
    check("add_poly_mixed_levels", lambda: pk.AddPoly(m, b))
    check("eval_poly", lambda: pk.EvalPoly(a))
