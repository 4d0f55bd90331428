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
    got = S.encrypt(v["x"][:k], None, tabs["P"], tabs["Q"])  # EncryptDeterministic
    exp = [O.g1_mul(x, S.P, par.p) for x in v["x"][:k]]
    assert got == exp


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
def test_sim_multpoly(kb):
    g, par, S, _ = setup(kb)
    v = g["multpoly"]
    c1, c2 = g1s(par, v["c1"]), g1s(par, v["c2"])
    assert S.multpoly(c1, v["d1"], c2, v["d2"], 1) == gts(par, v["out"])
    assert S.multpoly(c2, v["d2"], c1, v["d1"], 1) == gts(par, v["out"])
    if kb < 512:  # three units, ragged last block
        assert S.multpoly(c1 * 3, v["d1"], c2 * 3, v["d2"], 3) == gts(par, v["out"]) * 3


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
    gsk = O.fp2_pow(O.pairing(S.P, S.P, par), int(g["q1"], 16), par.p)
    for S_steps in (None, 7):  # the reference-sized table and a deliberately tiny one
        S.bsgs_setup(gsk, g["msg_space"], S_steps)
        out, status = S.bsgs_lookup(csk)
        assert [int(s) for s in status] == v["status"]
        assert [int(x) for x in out] == v["out"]


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
    wide = (ctypes.c_uint64 * 3)()
    lib.hs_wide_count(wide, 1)
    S.miller(c1, v["d1"], c2, v["d2"], 1, v["d1"] + v["d2"], teams_per_block=1)
    fused = lib.hs_mul_count(1)  # multiply-and-reduce products (2L^2 + L each)
    lib.hs_wide_count(wide, 1)   # double-width multiplications (L^2) and separate reductions (L^2 + L)
    L = S.L
    executed = fused * (2 * L * L + L) + wide[0] * L * L + wide[1] * (L * L + L)
    assert executed == workmodel.miller_unit_products(par.p, par.n, par.l, v["d1"], v["d2"])
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
    assert 16.0 < worst < 64.0  # the chord slope 2 (S2 - Y + 16p) is the largest value
    assert sim.lib().hs_selftest_violation() == 1  # and the checker does fire


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
                                       (("-DBGN_LINE_KARATSUBA=2",), "_kara")])
def test_sim_miller_layout_and_multiplier_variants(flags, tag):
    """Build variants of the Miller kernel: the thread-interleaved layout with the private slots in
    global memory (what the 1024-bit field ships with) and the Karatsuba multiplier (measured, not
    shipped) produce the same bytes and stay within the proven ranges."""
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
