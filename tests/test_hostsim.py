"""The DEVICE programs (bgn_b200/csrc/*.cuh) executed on the CPU -- carry flag, atomics and
thread grid emulated (tests/hostsim) -- against the golden vectors.  This checks the kernels'
per-thread logic without a GPU; the GPU tests repeat the same vectors through the C-ABI.
The simulation is test infrastructure: it is never linked into libbgn_b200.so."""
import json

import pytest

from conftest import load_golden
from oracle import bgn_oracle as O
import sim

SIM_KB = (64, 128, 512)  # limb counts 3, 5, 17 are instantiated in hostsim.cpp
_cache = {}


def setup(kb):
    if kb not in _cache:
        g = load_golden(kb)
        par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
        P = O.g1_from_bytes(bytes.fromhex(g["P"]), par)
        Q = O.g1_from_bytes(bytes.fromhex(g["Q"]), par)
        S = sim.Sim(par, P, Q, int(g["q1"], 16))
        _cache[kb] = (g, par, S, {})
    g, par, S, tabs = _cache[kb]
    S.activate()
    return g, par, S, tabs


@pytest.fixture(autouse=True)
def no_range_violations():
    """Every simulated kernel run is also a worst-case range proof (arith.cuh, bgnsim): no product
    operand may reach R - p and no relaxed difference may go negative, for ANY input values."""
    yield
    import ctypes
    out = (ctypes.c_double * 4)()
    sim.lib().hs_range_report(out, 0)
    assert int(out[3]) == 0, "range violations reported by the simulated device code"


def g1s(par, hexes):
    return [O.g1_from_bytes(bytes.fromhex(h), par) for h in hexes]


def gts(par, hexes):
    return [O.gt_from_bytes(bytes.fromhex(h), par) for h in hexes]


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_encrypt(kb):
    g, par, S, tabs = setup(kb)
    if "P" not in tabs:
        tabs["P"] = S.build_table(S.P, 8)
        tabs["Q"] = S.build_table(S.Q, S.nbytes)
    v = g["encrypt"]
    k = len(v["x"]) if kb < 512 else 5
    got = S.encrypt(v["x"][:k], [int(r, 16) for r in v["r"][:k]], tabs["P"], tabs["Q"])
    assert got == g1s(par, v["out"][:k])
    if kb == 64:  # the 16-bit window table the GPU path uses (2^16 - 1 points per window)
        if "Q16" not in tabs:
            tabs["Q16"] = S.build_table16(tabs["Q"], S.nbytes)
        assert S.encrypt(v["x"], [int(r, 16) for r in v["r"]], tabs["P"], tabs["Q16"], wbitsQ=16) == g1s(par, v["out"])
        small = [int(r, 16) >> 8 for r in v["r"]]  # zero top byte: a window with an empty high half

        def enc(x, r):
            c = O.g1_add(O.g1_mul(abs(x), S.P, par.p), O.g1_mul(r, S.Q, par.p), par.p)
            return O.g1_neg(c, par.p) if x < 0 else c

        assert S.encrypt(v["x"], small, tabs["P"], tabs["Q16"], wbitsQ=16) == [enc(x, r) for x, r in zip(v["x"], small)]
        # windows that are not whole bytes (api.cu: ensure_tabQw builds 18 = 2 x 9, 20 = 2 x 10, 22 = 2 x 11 this
        # way): 10 = 2 x 5 and 9 = 3 x 3 bits from a temporary narrow table, digits straddling byte boundaries
        for nsub, hb in ((2, 5), (3, 3)):
            wbits = nsub * hb
            nwin = (8 * S.nbytes + wbits - 1) // wbits
            narrow = S.build_table(S.Q, nwin * nsub, hb)
            wide = S.build_table_wide(narrow, nwin * nsub, nwin, nsub, hb)
            assert S.encrypt(v["x"], [int(r, 16) for r in v["r"]], tabs["P"], wide, wbitsQ=wbits) == g1s(par, v["out"])
            assert S.encrypt(v["x"], small, tabs["P"], wide, wbitsQ=wbits) == [enc(x, r) for x, r in zip(v["x"], small)]
    got = S.encrypt(v["x"][:k], None, tabs["P"], tabs["Q"])  # EncryptDeterministic
    exp = [O.g1_mul(x, S.P, par.p) for x in v["x"][:k]]
    assert got == exp
    # the same sums with the tables in twisted Edwards form (curve.cuh: Ed; what the GPU path uses by default)
    if "Pe" not in tabs:
        tabs["Pe"] = S.table_to_edwards(tabs["P"])
        tabs["Qe"] = S.table_to_edwards(tabs["Q"])
    S.range_report()
    got = S.encrypt(v["x"][:k], [int(r, 16) for r in v["r"][:k]], tabs["Pe"], tabs["Qe"], edw=True)
    assert got == g1s(par, v["out"][:k])
    assert S.encrypt(v["x"][:k], None, tabs["Pe"], tabs["Qe"], edw=True) == exp
    _, _, unknown, violations = S.range_report()
    assert unknown == 0 and violations == 0
    if kb == 64:
        # degenerate sums: r = 0 and x = 0 (the identity), r*Q = -(x*P)-like cancellations are covered by the
        # re-randomisation below: base + r*Q with base = -(r*Q) gives O, base = r*Q doubles
        rq = [O.g1_mul(int(r, 16), S.Q, par.p) for r in v["r"][:4]]
        rs = [int(r, 16) for r in v["r"][:4]]
        base = [O.g1_neg(rq[0], par.p), rq[1], None, rq[3]]
        exp_b = [O.g1_add(b, q, par.p) if b is not None else q for b, q in zip(base, rq)]
        assert S.encrypt(None, rs, tabs["Pe"], tabs["Qe"], base=base, edw=True) == exp_b
        assert S.encrypt([0, 0], [0, 0], tabs["Pe"], tabs["Qe"], edw=True) == [None, None]
        assert S.encrypt([5, -5], [0, 0], tabs["Pe"], tabs["Qe"], edw=True) == [O.g1_mul(5, S.P, par.p), O.g1_neg(O.g1_mul(5, S.P, par.p), par.p)]
        q16e = S.table_to_edwards(tabs["Q16"])
        assert S.encrypt(v["x"], [int(r, 16) for r in v["r"]], tabs["Pe"], q16e, wbitsQ=16, edw=True) == g1s(par, v["out"])
        assert S.encrypt(v["x"], small, tabs["Pe"], q16e, wbitsQ=16, edw=True) == [enc(x, r) for x, r in zip(v["x"], small)]
        # an all-zero entry is O (an unused slot), not an error; a finite point without an Edwards image --
        # (-1, sqrt(-2)) has order 4 -- raises the flag (api.cu then keeps the Weierstrass tables)
        import numpy as np
        t2 = np.zeros(2 * 2 * S.L, dtype=np.uint32)
        t2[2 * S.L:] = tabs["P"][: 2 * S.L]
        assert S.table_to_edwards(t2) is not None
        assert par.p % 8 == 3
        y4 = pow(par.p - 2, (par.p + 1) // 4, par.p)
        assert y4 * y4 % par.p == par.p - 2
        t2[: S.L] = S.soa([par.p - 1]).ravel()
        t2[S.L: 2 * S.L] = S.soa([y4]).ravel()
        assert S.table_to_edwards(t2) is None


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_g1_ops(kb):
    g, par, S, _ = setup(kb)
    v = g["g1_add"]
    assert S.g1_add(g1s(par, v["a"]), g1s(par, v["b"])) == g1s(par, v["out"])
    v = g["g1_sub"]
    assert S.g1_add(g1s(par, v["a"]), g1s(par, v["b"]), subtract=True) == g1s(par, v["out"])
    v = g["g1_neg"]
    assert S.g1_add([None], g1s(par, v["a"]), subtract=True, bcast1=True) == g1s(par, v["out"])
    v = g["g1_mulconst"]
    k = len(v["k"]) if kb < 512 else 5
    got = S.g1_mulvar(g1s(par, v["a"][:k]), [int(x, 16) for x in v["k"][:k]], v["kbytes"])
    assert got == g1s(par, v["out"][:k])


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_pairing(kb):
    g, par, S, _ = setup(kb)
    v = g["pair"]
    assert S.pair(g1s(par, v["a"]), g1s(par, v["b"])) == gts(par, v["out"])
    v = g["make_l2"]
    a = g1s(par, v["a"])
    assert S.miller(a, 1, [S.P], 1, len(a), 1, e_bcast=True) == gts(par, v["out"])


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_pair_duo(kb):
    """k_pair_duo (pairwarp.cuh: point arithmetic on one warp, line evaluation + accumulator on another,
    double-buffered hand-over) against the golden Mult vectors (incl. O operands) and the one-thread
    kernel; the X warp is run one step ahead, so a clobbered buffer would show.  Also the range proof
    of its routines, its variant with the plain (non-lazy) F_p^2 product the 1024-bit field uses, and
    its work model: 9 / 11 products per doubling / addition on the X side, 4 + 2 / 3 products plus the
    lazily reduced F_p^2 product on the F side."""
    import ctypes as C
    g, par, S, _ = setup(kb)
    v = g["pair"]
    A, Bp = g1s(par, v["a"]), g1s(par, v["b"])
    wide = (C.c_uint64 * 8)()
    sim.lib().hs_wide_count(wide, 1)
    sim.lib().hs_mul_count(1)
    S.range_report()
    assert S.pair_duo(A, Bp) == gts(par, v["out"])
    hi, headroom, unknown, viol = S.range_report()
    assert viol == 0 and unknown == 0 and hi <= 40.0, (hi, unknown, viol)
    sim.lib().hs_wide_count(wide, 1)
    nmul = sim.lib().hs_mul_count(1)
    from bgn_b200 import workmodel
    live = sum(1 for x, y in zip(A, Bp) if x is not None and y is not None)
    mm, mw, rd, nsq = workmodel.pair_duo_counts(par.p, par.n, par.l, loop=0)  # the simulator's build: unrolled products
    assert (nmul, wide[0], wide[1] - wide[4], wide[4]) == (live * mm, live * mw, live * rd, live * nsq), (
        nmul, list(wide), live, mm, mw, rd, nsq)
    if kb < 512:
        assert S.pair_duo(Bp, A, np_=5) == S.pair(Bp, A)
        try:
            sim.use_variant(None, ("-DBGN_LINE_LAZY=0",), "_nolazy")
            S.activate()
            assert S.pair_duo(A, Bp) == gts(par, v["out"])
            _, _, unknown, viol = S.range_report()
            assert unknown == 0 and viol == 0
        finally:
            sim.use_variant(None)


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_multpoly(kb):
    g, par, S, _ = setup(kb)
    v = g["multpoly"]
    c1, c2 = g1s(par, v["c1"]), g1s(par, v["c2"])
    assert S.multpoly(c1, v["d1"], c2, v["d2"], 1) == gts(par, v["out"])
    assert S.multpoly(c2, v["d2"], c1, v["d1"], 1) == gts(par, v["out"])
    if kb < 512:  # three units, ragged last block
        assert S.multpoly(c1 * 3, v["d1"], c2 * 3, v["d2"], 3) == gts(par, v["out"]) * 3
        # the point of order 2, (0, 0), as a coefficient on the evaluation side (only a handle can carry it: the
        # byte format reads it as O): its pairings are 1, as the oracle's Miller loop finds -- every line value
        # at it lies in F_p -- and the evaluation points must not be normalised by 1 / y = 1 / 0
        d1, d2 = v["d1"], v["d2"]
        a1, a2 = list(c1), list(c2)
        (a2 if d1 <= d2 else a1)[0] = (0, 0)  # the polynomial with fewer slots takes the Miller side
        exp = []
        for j in range(d1 + d2):
            acc = (1, 0)
            for i in range(d1):
                k = j - i
                if 0 <= k < d2 and a1[i] is not None and a2[k] is not None:
                    mp, ep = (a1[i], a2[k]) if d1 <= d2 else (a2[k], a1[i])
                    acc = O.fp2_mul(acc, O.pairing(mp, ep, par), par.p)
            exp.append(acc)
        assert S.multpoly(a1, d1, a2, d2, 1) == exp
        assert S.multpoly_split(a1, d1, a2, d2, 1) == exp


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_multpoly_wide(kb):
    """k_miller_wide (MillerTeam<L, true>: evaluation points read from the batch arrays, 10 shared-memory
    slots per thread) against the golden MultPoly vectors, several units per block, and the broadcast
    evaluation side of makeL2."""
    g, par, S, _ = setup(kb)
    v = g["multpoly"]
    c1, c2 = g1s(par, v["c1"]), g1s(par, v["c2"])
    S.range_report()
    assert S.multpoly(c1, v["d1"], c2, v["d2"], 1, wide=True) == gts(par, v["out"])
    _, _, unknown, viol = S.range_report()
    assert viol == 0 and unknown == 0
    if kb < 512:
        assert S.multpoly(c2 * 3, v["d2"], c1 * 3, v["d1"], 3, wide=True) == gts(par, v["out"]) * 3
        m = g["make_l2"]
        a = g1s(par, m["a"])
        assert S.miller(a, 1, [S.P], 1, len(a), 1, e_bcast=True, wide=True) == gts(par, m["out"])


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_multpoly_split(kb):
    """k_miller_split (teamsplit.cuh: two threads per output-slot pair, partial accumulators multiplied
    before the final exponentiation, shared memory laid out per column) against the golden MultPoly
    vectors and the team kernel on other shapes (d1 != d2, d1 = 2, several units per block, O
    coefficients); the run is the range proof of the split schedule."""
    g, par, S, _ = setup(kb)
    v = g["multpoly"]
    c1, c2 = g1s(par, v["c1"]), g1s(par, v["c2"])
    S.range_report()
    assert S.multpoly_split(c1, v["d1"], c2, v["d2"], 1) == gts(par, v["out"])
    assert S.multpoly_split(c1, v["d1"], c2, v["d2"], 1, para=0) == gts(par, v["out"])  # separate doubling / addition steps
    hi, headroom, unknown, viol = S.range_report()
    assert viol == 0 and unknown == 0 and hi <= 40.0, (hi, unknown, viol)
    if kb < 512:
        assert S.multpoly_split(c2, v["d2"], c1, v["d1"], 1) == S.multpoly(c2, v["d2"], c1, v["d1"], 1)
        assert S.multpoly_split(c1 * 3, v["d1"], c2 * 3, v["d2"], 3) == gts(par, v["out"]) * 3
        assert S.multpoly_split(c1[:2] * 3, 2, c2 * 3, v["d2"], 3, teams_per_block=1) == S.multpoly(c1[:2] * 3, 2, c2 * 3, v["d2"], 3)
        sq = c2[: v["d2"]]
        assert S.multpoly_split(sq, v["d2"], sq, v["d2"], 1) == S.multpoly(sq, v["d2"], sq, v["d2"], 1)


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_gt_ops(kb):
    g, par, S, _ = setup(kb)
    v = g["gt_mul"]
    assert S.gt_mul(gts(par, v["a"]), gts(par, v["b"])) == gts(par, v["out"])
    v = g["gt_div"]
    assert S.gt_mul(gts(par, v["a"]), gts(par, v["b"]), conj_b=True) == gts(par, v["out"])
    v = g["gt_inv"]
    assert S.gt_pow(gts(par, v["a"]), 2) == gts(par, v["out"])
    v = g["gt_pow"]
    assert S.gt_pow(gts(par, v["a"]), 0, [int(k, 16) for k in v["k"]], v["kbytes"]) == gts(par, v["out"])
    v = g["l2_sum"]
    assert S.gt_reduce(gts(par, v["in"]), v["nterms"], v["ncoeff"], 1) == gts(par, v["out"])
    two = S.gt_reduce(gts(par, v["in"]), v["nterms"], v["ncoeff"], 2)  # two partial products per slot
    merged = S.gt_reduce(two, 2, v["ncoeff"], 1)
    assert merged == gts(par, v["out"])


@pytest.mark.parametrize("kb", (64, 128))
def test_sim_decrypt(kb):
    g, par, S, _ = setup(kb)
    v = g["decrypt_l2"]
    csk = S.gt_pow(gts(par, v["in"]), 1)  # C^q1
    assert csk == gts(par, v["csk"])
    assert S.gt_pow(gts(par, v["in"]), 1, pair=True) == csk  # the lane-pair kernel
    gsk = O.fp2_pow(O.pairing(S.P, S.P, par), int(g["q1"], 16), par.p)
    for S_steps in (None, 7):  # the reference-sized table and a deliberately tiny one
        S.bsgs_setup(gsk, g["msg_space"], S_steps)
        out, status = S.bsgs_lookup(csk)
        assert [int(s) for s in status] == v["status"]
        assert [int(x) for x in out] == v["out"]


@pytest.mark.parametrize("kb", (64, 128))
def test_sim_decrypt_lucas(kb):
    """k_dec_lucas (Lucas ladder on the trace, a lane pair per ciphertext, search by real part)
    gives the plaintexts and statuses of the reference's decrypt, incl. negatives, zero, the
    out-of-bounds values, and status 1 for a non-unitary input."""
    g, par, S, _ = setup(kb)
    v = g["decrypt_l2"]
    gsk = O.fp2_pow(O.pairing(S.P, S.P, par), int(g["q1"], 16), par.p)
    S.bsgs_setup(gsk, g["msg_space"], None)
    cts = gts(par, v["in"])
    out, status = S.dec_lucas(cts)
    assert [int(s) for s in status] == v["status"]
    assert [int(x) for x in out] == v["out"]
    # every |m| up to a few dozen, both signs, blinded by a random e(Q,Q) power
    import random
    rng = random.Random(kb)
    ePP, eQQ = O.pairing(S.P, S.P, par), O.pairing(S.Q, S.Q, par)
    ms = list(range(-20, 21)) + [S.mmax, -S.mmax, S.mmax + 1]
    cts = [O.fp2_mul(O.fp2_pow(ePP, m % par.n, par.p), O.fp2_pow(eQQ, rng.randrange(par.n), par.p), par.p) for m in ms]
    out, status = S.dec_lucas(cts)
    assert [int(x) for x in out] == ms[:-1] + [0]
    assert [int(s) for s in status] == [0] * (len(ms) - 1) + [1]
    bad = (cts[3][0], (cts[3][1] + 1) % par.p)  # not of norm 1
    assert S.dec_lucas([bad]) == ([0], [1])


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_bytes(kb):
    g, par, S, _ = setup(kb)
    v = g["g1_add"]
    raw = b"".join(bytes.fromhex(h) for h in v["a"])
    pts = S.g1_from_bytes(raw, len(v["a"]))
    assert pts == g1s(par, v["a"])
    assert S.g1_to_bytes(pts) == raw
    bad = bytearray(raw[: 2 * S.B])
    bad[-1] ^= 1
    assert S.g1_from_bytes(bytes(bad), 1) == [None]  # off-curve -> O
    e = gts(par, g["pair"]["out"])
    assert S.gt_roundtrip_bytes(e) == b"".join(bytes.fromhex(h) for h in g["pair"]["out"])


@pytest.mark.parametrize("kb", (64, 512))
def test_work_model_matches_executed_products(kb):
    """bgn_b200.workmodel (bench.py's roofline numerator) == Montgomery products the simulated
    kernel really executes."""
    from bgn_b200 import workmodel
    g, par, S, _ = setup(kb)
    v = g["multpoly"]
    c1 = [O.g1_mul(3 + i, S.P, par.p) for i in range(v["d1"])]
    c2 = [O.g1_mul(5 + i, S.Q, par.p) for i in range(v["d2"])]
    lib = sim.lib()
    ctypes = __import__("ctypes")
    lib.hs_mul_count.restype = ctypes.c_uint64
    lib.hs_mul_count(1)
    wide = (ctypes.c_uint64 * 8)()
    lib.hs_wide_count(wide, 1)
    S.miller(c1, v["d1"], c2, v["d2"], 1, v["d1"] + v["d2"], teams_per_block=1)
    fused = lib.hs_mul_count(1)  # multiply-and-reduce products (2L^2 + L each)
    lib.hs_wide_count(wide, 1)   # double-width multiplications (L^2) and separate reductions (L^2 + L)
    L = S.L
    executed = (fused * (2 * L * L + L) + wide[0] * L * L + wide[1] * (L * L + L) + wide[3] * (3 * L * L + L)
                + wide[4] * (L * (L + 1) // 2) + wide[5] * (4 * L * L + L))
    assert executed == workmodel.miller_unit_products(par.p, par.n, par.l, v["d1"], v["d2"])
    naf = workmodel.naf_digits(par.n)
    adds = sum(1 for i in range(1, len(naf) - 1) if naf[i] != 0) if workmodel.parabola_on(L, dE=v["d2"]) else 0
    assert wide[3] == (workmodel.miller_unit_lines(par.n, v["d1"], v["d2"], L) if workmodel.EVAL_NORM else 0) + adds * v["d1"] * workmodel.DADD_DOT2
    assert wide[5] == workmodel.miller_unit_parabolas(par.n, v["d1"], v["d2"], L)
    assert wide[4] == workmodel.miller_unit_squarings(par.p, par.n, par.l, v["d1"], v["d2"]) + adds * v["d1"] * workmodel.DADD_SQRS
    assert (wide[0] > 0) == workmodel.line_lazy(L)
    assert workmodel.pick_limbs(par.p) == S.L


@pytest.mark.parametrize("kb", SIM_KB)
def test_range_tracker_proves_miller_ranges(kb):
    """The fused Miller routines (fused.cuh) skip the per-operation reduction; the simulation
    propagates the worst-case bound of every value (with the smallest headroom the limb-count rule
    allows, R/p = 256): no untracked operand, no violation, all bounds far below the headroom."""
    g, par, S, _ = setup(kb)
    S.range_report()
    v = g["multpoly"]
    c1, c2 = g1s(par, v["c1"]), g1s(par, v["c2"])
    assert S.multpoly(c1, v["d1"], c2, v["d2"], 1) == gts(par, v["out"])
    worst, headroom, unknown, violations = S.range_report()
    assert headroom == 256.0
    assert unknown == 0 and violations == 0
    assert 8.0 < worst < 64.0  # (with the parabola step the team kernel no longer forms the chord slope 2 (S2 - Y + 16p))
    assert sim.lib().hs_selftest_violation() == 1  # and the checker does fire


def test_sim_fused_routines_with_dedicated_squaring():
    """The option BGN_FUSED_SQR (dedicated squarings inside dbl_line / madd_line / fe_prepare; measured
    neutral in k_miller and not shipped) keeps bytes, ranges and its work model."""
    import ctypes as C
    from bgn_b200 import workmodel
    try:
        sim.use_variant(None, ("-DBGN_FUSED_SQR=1",), "_fsqr")
        workmodel.FUSED_SQR = True
        for kb in (64, 512):
            g, par, S, _ = setup(kb)
            v = g["multpoly"]
            c1, c2 = g1s(par, v["c1"]), g1s(par, v["c2"])
            wide = (C.c_uint64 * 8)()
            sim.lib().hs_wide_count(wide, 1)
            assert S.multpoly(c1, v["d1"], c2, v["d2"], 1) == gts(par, v["out"])
            sim.lib().hs_wide_count(wide, 1)
            live1 = sum(1 for x in c1 if x is not None)
            if live1 == len(c1) and all(x is not None for x in c2):
                assert wide[4] == workmodel.miller_unit_squarings(par.p, par.n, par.l, min(v["d1"], v["d2"]), max(v["d1"], v["d2"]))
            assert wide[4] > 0
            _, _, unknown, violations = S.range_report()
            assert unknown == 0 and violations == 0
    finally:
        workmodel.FUSED_SQR = False
        sim.use_variant(None)


@pytest.mark.parametrize("loop", (1, 2, 4))
def test_sim_miller_loop_variants(loop):
    """Fp::mul_loop<U> (the row loop the 1024-bit field ships with) gives the same bytes."""
    try:
        sim.use_variant(loop)
        for kb in (64, 128):
            g, par, S, _ = setup(kb)
            v = g["multpoly"]
            c1, c2 = g1s(par, v["c1"]), g1s(par, v["c2"])
            assert S.multpoly(c1, v["d1"], c2, v["d2"], 1) == gts(par, v["out"])
            v = g["pair"]
            assert S.pair(g1s(par, v["a"]), g1s(par, v["b"])) == gts(par, v["out"])
            _, _, unknown, violations = S.range_report()
            assert unknown == 0 and violations == 0
    finally:
        sim.use_variant(None)


@pytest.mark.parametrize("flags,tag", [(("-DBGN_MILLER_GP=1", "-DBGN_MILLER_NT=32", "-DBGN_LINE_LAZY=0"), "_gp"),
                                       (("-DBGN_MILLER_GP=1", "-DBGN_MILLER_NT=32"), "_gplazy"),
                                       (("-DBGN_LINE_KARATSUBA=2",), "_kara"),
                                       (("-DBGN_EVAL_NORM=0",), "_plainlines"),
                                       (("-DBGN_PARABOLA=0",), "_noparabola")])
def test_sim_miller_layout_and_multiplier_variants(flags, tag):
    """Build variants of the Miller kernel: the thread-interleaved layout with the private slots in
    global memory (what the 1024-bit field ships with), the Karatsuba multiplier (measured, not
    shipped), the plain line form (evaluation points not normalised: the A/B switch of DESIGN 2.4) and the
    separate doubling / addition steps (no parabola: the A/B switch of DESIGN 2.5)
    produce the same bytes and stay within the proven ranges."""
    try:
        sim.use_variant(None, flags, tag)
        for kb in (64, 512):
            g, par, S, _ = setup(kb)
            v = g["multpoly"]
            c1, c2 = g1s(par, v["c1"]), g1s(par, v["c2"])
            assert S.multpoly(c1, v["d1"], c2, v["d2"], 1) == gts(par, v["out"])
            if kb == 64:
                assert S.multpoly(c1 * 3, v["d1"], c2 * 3, v["d2"], 3) == gts(par, v["out"]) * 3
            v = g["pair"]
            assert S.pair(g1s(par, v["a"]), g1s(par, v["b"])) == gts(par, v["out"])
            _, _, unknown, violations = S.range_report()
            assert unknown == 0 and violations == 0
    finally:
        sim.use_variant(None)


# ---------------------------------------------------------------- non-deterministic mode + poly helpers (8(f1), 8(f2))
@pytest.mark.parametrize("kb", (64, 128))
def test_sim_blind(kb):
    """bgn_g1_blind_batch / bgn_gt_blind_batch device programs against the oracle's
    non-deterministic Add / Mult with injected r (bgn.go:466-474, 488-495)."""
    import random
    g, par, S, tabs = setup(kb)
    if "P" not in tabs:
        tabs["P"] = S.build_table(S.P, 8)
        tabs["Q"] = S.build_table(S.Q, S.nbytes)
    rng = random.Random(kb)
    pk = O.PublicKey(par, S.P, S.Q, g["msg_space"], deterministic=False)
    a = [O.encrypt_with_randomness(pk, x, rng.randrange(par.n)) for x in (0, 1, 5)]
    b = [O.encrypt_with_randomness(pk, x, rng.randrange(par.n)) for x in (2, 3, 4)]
    rs = [rng.randrange(par.n) for _ in a] + [0]
    pdet = O.PublicKey(par, S.P, S.Q, g["msg_space"])
    sums = [O.add(pdet, x, y).C for x, y in zip(a, b)] + [None]  # O as a base point
    exp = [O.add(pk, x, y, r).C for x, y, r in zip(a, b, rs)] + [None]
    assert S.encrypt(None, rs, tabs["P"], tabs["Q"], base=sums) == exp
    # level 2
    if "E" not in tabs:
        tabs["E"] = S.gt_table(O.pairing(S.Q, S.Q, par), S.nbytes)
    prods = [O.mult(pdet, x, y).C for x, y in zip(a, b)]
    expm = [O.mult(pk, x, y, r).C for x, y, r in zip(a, b, rs)]
    assert S.gt_blind(prods, rs[:3], tabs["E"]) == expm
    assert S.gt_blind(prods[:1], [0], tabs["E"]) == prods[:1]


@pytest.mark.parametrize("kb", (64, 128))
@pytest.mark.parametrize("l2", (False, True))
def test_sim_polyconv(kb, l2):
    """MultConstPoly / EvalPoly device programs against the oracle (poly.go:58-120)."""
    import random
    g, par, S, _ = setup(kb)
    rng = random.Random(kb + 5)
    pk = O.PublicKey(par, S.P, S.Q, g["msg_space"])
    count = 3
    polys = []
    for u in range(count):
        pt = pk.new_poly_plaintext([9.13, 4.0, 0.0][u])
        ct = O.encrypt_poly(pk, pt, [rng.randrange(par.n) if u else 0 for _ in range(pt.degree)])
        polys.append(O.make_poly_l2(pk, ct) if l2 else ct)
    d = max(p_.degree for p_ in polys)
    ident = O.make_l2(pk, O.encrypt_zero(pk)) if l2 else O.encrypt_zero(pk)
    for p_ in polys:  # pad to a common slot count with the identity, as a batch caller does
        while p_.degree < d:
            p_.coefficients.append(ident)
            p_.degree += 1
    flat = [c.C for p_ in polys for c in p_.coefficients]
    for constant in (4.12, -2.0):
        digits = pk.new_unbalanced_plaintext(abs(constant)).coefficients
        nd = len(digits)
        got = S.polyconv(flat, d, l2, digits, 0, d + nd, constant < 0, count)
        exp = [c.C for p_ in polys for c in O.mult_const_poly(pk, p_, constant).coefficients]
        assert got == exp
    w = [pk.poly_base ** (d - 1 - k) for k in range(d)]
    got = S.polyconv(flat, d, l2, w, d - 1, 1, False, count)
    assert got == [O.eval_poly(pk, p_).C for p_ in polys]


def test_sim_padded_gt_bytes():
    g, par, S, _ = setup(64)
    vals = gts(par, g["gt_mul"]["a"])[:4]
    out = S.gt_to_bytes_padded(vals, 2, 1)
    one = O.gt_to_bytes((1, 0), par)
    exp = b"".join(O.gt_to_bytes(v, par) for v in vals[:2]) + one + b"".join(O.gt_to_bytes(v, par) for v in vals[2:]) + one
    assert out == exp


@pytest.mark.parametrize("kb", (64, 128))
def test_sim_new_sections_golden(kb):
    """the committed golden vectors of the 8(f) sections through the simulated device programs"""
    g, par, S, tabs = setup(kb)
    if "P" not in tabs:
        tabs["P"] = S.build_table(S.P, 8)
        tabs["Q"] = S.build_table(S.Q, S.nbytes)
    v = g["g1_blind"]
    assert S.encrypt(None, [int(r, 16) for r in v["r"]], tabs["P"], tabs["Q"], base=g1s(par, v["a"])) == g1s(par, v["out"])
    if "E" not in tabs:
        tabs["E"] = S.gt_table(O.pairing(S.Q, S.Q, par), S.nbytes)
    v = g["gt_blind"]
    assert S.gt_blind(gts(par, v["a"]), [int(r, 16) for r in v["r"]], tabs["E"]) == gts(par, v["out"])
    for lvl, conv in (("l1", g1s), ("l2", gts)):
        v = g["multconstpoly_" + lvl]
        for case in v["cases"]:
            got = S.polyconv(conv(par, v["in"]), v["d"], lvl == "l2", case["digits"], 0, v["d"] + len(case["digits"]),
                             case["negate"], v["count"])
            assert got == conv(par, case["out"])
        v = g["evalpoly_" + lvl]
        w = [v["base"] ** (v["d"] - 1 - k) for k in range(v["d"])]
        assert S.polyconv(conv(par, v["in"]), v["d"], lvl == "l2", w, v["d"] - 1, 1, False, v["count"]) == conv(par, v["out"])


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_pair_fixed(kb):
    """k_miller_fixed (replay of the recorded line table of P) against the golden makeL2 vectors
    and the oracle's pairing for arbitrary evaluation points, incl. O."""
    g, par, S, tabs = setup(kb)
    if "linesP" not in tabs:
        tabs["linesP"] = S.record_lines(S.P)
    v = g["make_l2"]
    assert S.pair_fixed(tabs["linesP"], g1s(par, v["a"])) == gts(par, v["out"])
    if kb < 512:
        pts = g1s(par, g["g1_add"]["out"])
        assert S.pair_fixed(tabs["linesP"], pts, nt=3) == [O.pairing(pt, S.P, par) for pt in pts]
        # a point of even order has a line whose third coefficient vanishes: no normalised table (api.cu then
        # keeps the general kernel)
        assert S.record_lines((0, 0)) is None


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_pair_fixed_lane_pair(kb):
    """k_miller_fixed_pair (pairlane.cuh: one pairing on a pair of lanes, squaring / line evaluation /
    dot-product halves swapped by shuffles) against the golden makeL2 vectors, the one-thread kernel
    and the oracle, incl. O; the run is also the range proof of its relaxed arithmetic and pins its
    work model: per point 1.5 Montgomery products and 1 dot product per lane and step."""
    import ctypes as C
    g, par, S, tabs = setup(kb)
    if "linesP" not in tabs:
        tabs["linesP"] = S.record_lines(S.P)
    v = g["make_l2"]
    pts = g1s(par, v["a"])
    wide = (C.c_uint64 * 8)()
    sim.lib().hs_wide_count(wide, 1)
    sim.lib().hs_mul_count(1)
    S.range_report()
    assert S.pair_fixed_pair(tabs["linesP"], pts) == gts(par, v["out"])
    hi, headroom, unknown, viol = S.range_report()
    assert viol == 0 and hi <= 16.0, (hi, viol)
    sim.lib().hs_wide_count(wide, 1)
    nmul = sim.lib().hs_mul_count(1)
    from bgn_b200 import workmodel
    finite = sum(1 for pt in pts if pt is not None)
    if finite == len(pts):
        exp_dot, exp_mul = workmodel.miller_fixed_pair_counts(par.p, par.n, par.l)
        assert (wide[3], nmul, wide[4]) == (exp_dot * len(pts), exp_mul * len(pts), 0), (list(wide), nmul)
    if kb < 512:
        pts = g1s(par, g["g1_add"]["out"])
        assert S.pair_fixed_pair(tabs["linesP"], pts) == [O.pairing(pt, S.P, par) for pt in pts]
        assert S.pair_fixed_pair(tabs["linesP"], pts) == S.pair_fixed(tabs["linesP"], pts, nt=3)


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_dedicated_squaring(kb):
    """Fp::sqr (L (L + 1) / 2 products + one Montgomery reduction: cross products against the doubled
    operand, squares leading the even chains) equals Fp::mul(a, a) LIMB FOR LIMB -- same Montgomery
    quotient, same representative -- on random values, on limbs with their top bits set (the bit the
    doubling carries from limb i into limb i + 1 must not enter row i), and on relaxed operands up to
    64 p."""
    import ctypes as C
    import random
    import numpy as np
    g, par, S, _ = setup(kb)
    rng = random.Random(kb + 5)
    p, L = par.p, S.L
    vals = [0, 1, p - 1, p - 2, (1 << (32 * L - 9)) - 1]
    vals += [rng.randrange(p) for _ in range(200 if kb < 512 else 40)]
    vals += [sum((0x80000000 | rng.getrandbits(31)) << (32 * j) for j in range(L)) % p for _ in range(50 if kb < 512 else 10)]
    vals += [sum(0xFFFFFFFF << (32 * j) for j in range(L)) % p, sum(0x80000000 << (32 * j) for j in range(L)) % p]
    a = S.soa(vals, mont=False)
    relaxed = [v + k * p for v, k in zip(vals[:40], [rng.randrange(64) for _ in range(40)])]
    b = np.zeros((len(relaxed), L), dtype=np.uint32)
    for e, v in enumerate(relaxed):
        for j in range(L):
            b[e, j] = (v >> (32 * j)) & 0xFFFFFFFF
    sim.lib().hs_track_array(sim.P32(b), C.c_size_t(len(relaxed)), L, C.c_double(64.0))
    Rinv = pow(1 << (32 * L), -1, p)
    for arr, src in ((a, vals), (b, relaxed)):
        r1, r2 = np.zeros_like(arr), np.zeros_like(arr)
        assert sim.lib().hs_fp_sqr(L, sim.P32(r1), sim.P32(r2), sim.P32(arr), C.c_size_t(len(src))) == 0
        assert (r1 == r2).all(), "dedicated squaring and product differ"
        got = [sum(int(r1[e, j]) << (32 * j) for j in range(L)) for e in range(len(src))]
        assert all(gv % p == v * v * Rinv % p for gv, v in zip(got, src))
    _, _, _, viol = S.range_report()
    assert viol == 0


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_inv_gcd(kb):
    """F<L>::inv_gcd (constant-time binary GCD on plain ALU instructions) equals the Fermat inverse
    and the integer inverse, incl. 0 -> 0, 1, p - 1 and a value in [p, 2p)."""
    import ctypes as C
    import random
    import numpy as np
    g, par, S, _ = setup(kb)
    rng = random.Random(kb + 77)
    p = par.p
    vals = [0, 1, 2, p - 1, p - 2, (p + 1) // 2] + [rng.randrange(p) for _ in range(12 if kb < 512 else 4)]
    for v in vals:
        a = S.soa([v])
        r1, r2 = np.zeros_like(a), np.zeros_like(a)
        assert sim.lib().hs_fp_inv_gcd(S.L, sim.P32(r1), sim.P32(a)) == 0
        assert sim.lib().hs_fp_inv(S.L, sim.P32(r2), sim.P32(a)) == 0
        exp = pow(v, -1, p) if v else 0
        assert S.unsoa(r1, 1) == [exp] and S.unsoa(r2, 1) == [exp]
    # lazy-form input: v + p represents v
    a = S.soa([5])
    raw = sum(int(a[0, j]) << (32 * j) for j in range(S.L)) + p
    for j in range(S.L):
        a[0, j] = (raw >> (32 * j)) & 0xFFFFFFFF
    sim.lib().hs_track_array(sim.P32(a), C.c_size_t(1), S.L, C.c_double(2.0))
    r1 = np.zeros_like(a)
    assert sim.lib().hs_fp_inv_gcd(S.L, sim.P32(r1), sim.P32(a)) == 0
    assert S.unsoa(r1, 1) == [pow(5, -1, p)]


def test_sim_evalpoly_wide_weights():
    """EvalPoly of a 50-slot polynomial: base^(d-1) = 3^49 needs the weights' upper 64 bits."""
    import random
    g, par, S, _ = setup(64)
    rng = random.Random(50)
    pk = O.PublicKey(par, S.P, S.Q, g["msg_space"])
    d = 50
    cs = [O.encrypt_with_randomness(pk, rng.randrange(3), rng.randrange(par.n)) for _ in range(d)]
    ct = O.PolyCiphertext(cs, d, 0, False)
    w = [3 ** (d - 1 - k) for k in range(d)]
    assert w[0] >> 64
    assert S.polyconv([c.C for c in cs], d, False, w, d - 1, 1, False, 1) == [O.eval_poly(pk, ct).C]


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_g1_affadd(kb):
    """k_g1_affadd against the golden Add / Sub / Neg vectors (O operands, doubling, inverse pairs) for
    several thread counts (elements per shared inversion), and a 2-torsion corner: (0,0) + (0,0) = O."""
    g, par, S, _ = setup(kb)
    for G in (1, 3, 50):
        v = g["g1_add"]
        assert S.g1_affadd(g1s(par, v["a"]), g1s(par, v["b"]), G=G) == g1s(par, v["out"])
        v = g["g1_sub"]
        assert S.g1_affadd(g1s(par, v["a"]), g1s(par, v["b"]), subtract=True, G=G) == g1s(par, v["out"])
    v = g["g1_neg"]
    assert S.g1_affadd([None], g1s(par, v["a"]), subtract=True, bcast1=True) == g1s(par, v["out"])
    if kb == 64:
        pt = g1s(par, g["g1_add"]["a"])[0]
        assert S.g1_affadd([(0, 0), pt], [(0, 0), (0, 0)]) == [None, O.g1_add(pt, (0, 0), par.p)]


@pytest.mark.parametrize("kb", SIM_KB)
def test_sim_inv_safegcd(kb):
    """The batched-division-step inversion (Fp::inv_safegcd) behind its verify-and-fall-back wrapper:
    correct on random and structured operands, and the fast path itself succeeds (no fall-backs)."""
    import ctypes as C
    import random
    import numpy as np
    g, par, S, _ = setup(kb)
    rng = random.Random(kb + 99)
    p = par.p
    bits = p.bit_length()
    vals = [0, 1, 2, 3, p - 1, p - 2, (p + 1) // 2, p >> 1, 1 << (bits - 2), (1 << (bits - 2)) + 1, (1 << 30) - 1, 1 << 30,
            (1 << 60) + 1]
    vals += [rng.randrange(p) for _ in range(400 if kb < 512 else 60)]
    vals += [rng.getrandbits(rng.randrange(1, bits)) % p for _ in range(200 if kb < 512 else 30)]
    a = S.soa(vals)
    r = np.zeros_like(a)
    fb = C.c_uint64()
    assert sim.lib().hs_fp_inv_safe(S.L, sim.P32(r), sim.P32(a), C.c_size_t(len(vals)), C.byref(fb)) == 0
    assert S.unsoa(r, len(vals)) == [pow(v, -1, p) if v else 0 for v in vals]
    assert fb.value == 0, "the fast inversion fell back %d times" % fb.value
