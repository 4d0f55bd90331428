"""GPU parity: the CUDA path, called through the C-ABI (ctypes -> libbgn_b200.so),
against (i) the committed golden vectors and (ii) the oracle on seeded inputs.
Bit-exact: every comparison is on serialised bytes / integers."""
import random

import numpy as np
import pytest

from conftest import load_golden, unhex

pytestmark = pytest.mark.gpu

_engines = {}


def engine_for(g):
    from bgn_b200 import Engine
    kb = g["key_bits"]
    if kb not in _engines:
        e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
        e.set_secret(int(g["q1"], 16), g["msg_space"])
        _engines[kb] = e
    return _engines[kb]


def buf(xs):
    return np.frombuffer(unhex(xs), dtype=np.uint8)


def scal(e, hexes, width):
    return e.scalars_be([int(h, 16) for h in hexes], width)


def test_ctx_info(golden):
    e = engine_for(golden)
    assert e.coord_bytes == golden["coord_bytes"]
    assert e.scalar_bytes == (int(golden["n"], 16).bit_length() + 7) // 8
    assert 32 * e.limbs >= int(golden["p"], 16).bit_length() + 3


def test_encrypt(golden):
    e, v = engine_for(golden), golden["encrypt"]
    out = e.encrypt_batch(np.array(v["x"], dtype=np.int64), scal(e, v["r"], e.scalar_bytes))
    assert out.tobytes() == unhex(v["out"])


def test_encrypt_deterministic(golden):
    """r_be = NULL is EncryptDeterministic (bgn.go:325-331) == randomness 0."""
    e = engine_for(golden)
    xs = np.array([0, 1, 2, -1, 1021], dtype=np.int64)
    a = e.encrypt_batch(xs, None)
    b = e.encrypt_batch(xs, e.scalars_be([0] * len(xs)))
    assert a.tobytes() == b.tobytes()
    assert a[: e.elem_bytes].tobytes() == bytes(e.elem_bytes)  # x = 0 -> O -> all-zero bytes


@pytest.mark.parametrize("op", ["g1_add", "g1_sub"])
def test_g1_binop(golden, op):
    e, v = engine_for(golden), golden[op]
    out = getattr(e, op + "_batch")(buf(v["a"]), buf(v["b"]))
    assert out.tobytes() == unhex(v["out"])


def test_g1_neg(golden):
    e, v = engine_for(golden), golden["g1_neg"]
    assert e.g1_neg_batch(buf(v["a"])).tobytes() == unhex(v["out"])


def test_g1_mulconst(golden):
    e, v = engine_for(golden), golden["g1_mulconst"]
    out = e.g1_mulconst_batch(buf(v["a"]), scal(e, v["k"], v["kbytes"]), v["kbytes"])
    assert out.tobytes() == unhex(v["out"])


def test_pair(golden):
    e, v = engine_for(golden), golden["pair"]
    assert e.pair_batch(buf(v["a"]), buf(v["b"])).tobytes() == unhex(v["out"])
    # symmetric pairing
    assert e.pair_batch(buf(v["b"]), buf(v["a"])).tobytes() == unhex(v["out"])


def test_make_l2(golden):
    e, v = engine_for(golden), golden["make_l2"]
    assert e.make_l2_batch(buf(v["a"])).tobytes() == unhex(v["out"])


@pytest.mark.parametrize("op", ["gt_mul", "gt_div"])
def test_gt_binop(golden, op):
    e, v = engine_for(golden), golden[op]
    assert getattr(e, op + "_batch")(buf(v["a"]), buf(v["b"])).tobytes() == unhex(v["out"])


def test_gt_inv(golden):
    e, v = engine_for(golden), golden["gt_inv"]
    assert e.gt_inv_batch(buf(v["a"])).tobytes() == unhex(v["out"])


def test_gt_pow(golden):
    e, v = engine_for(golden), golden["gt_pow"]
    out = e.gt_pow_batch(buf(v["a"]), scal(e, v["k"], v["kbytes"]), v["kbytes"])
    assert out.tobytes() == unhex(v["out"])


def test_multpoly(golden):
    e, v = engine_for(golden), golden["multpoly"]
    out = e.multpoly_batch(buf(v["c1"]), v["d1"], buf(v["c2"]), v["d2"], 1)
    assert out.tobytes() == unhex(v["out"])
    # swapped operands: the product is commutative
    out = e.multpoly_batch(buf(v["c2"]), v["d2"], buf(v["c1"]), v["d1"], 1)
    assert out.tobytes() == unhex(v["out"])


def test_multpoly_batch_of_copies(golden):
    """ragged block occupancy: 37 units of the same product must all equal the golden one."""
    e, v = engine_for(golden), golden["multpoly"]
    cnt = 37
    out = e.multpoly_batch(np.tile(buf(v["c1"]), cnt), v["d1"], np.tile(buf(v["c2"]), cnt), v["d2"], cnt)
    assert out.tobytes() == unhex(v["out"]) * cnt


def test_l2_sum(golden):
    e, v = engine_for(golden), golden["l2_sum"]
    assert e.l2_sum_reduce(buf(v["in"]), v["nterms"], v["ncoeff"]).tobytes() == unhex(v["out"])


def test_l2_sum_large_tree(golden):
    """multi-pass tree: 1000 copies of term set -> product = golden^(1000) per coefficient."""
    e, v = engine_for(golden), golden["l2_sum"]
    reps = 200
    got = e.l2_sum_reduce(np.tile(buf(v["in"]), reps), v["nterms"] * reps, v["ncoeff"])
    k = e.scalars_be([reps] * v["ncoeff"], 2)
    exp = e.gt_pow_batch(buf(v["out"]), k, 2)
    assert got.tobytes() == exp.tobytes()


def test_gt_pow_secret(golden):
    e, v = engine_for(golden), golden["decrypt_l2"]
    assert e.gt_pow_secret_batch(buf(v["in"])).tobytes() == unhex(v["csk"])


def test_decrypt_l2(golden):
    e, v = engine_for(golden), golden["decrypt_l2"]
    vals, st = e.decrypt_batch(buf(v["in"]), True)
    assert list(st) == v["status"]
    assert [int(x) for x in vals] == v["out"]


def test_decrypt_l1(golden):
    e, v = engine_for(golden), golden["decrypt_l1"]
    vals, st = e.decrypt_batch(buf(v["in"]), False)
    assert list(st) == v["status"]
    assert [int(x) for x in vals] == v["out"]


def test_decrypt_l1_both_routes():
    """Level-1 Decrypt as one pairing e(C, q1 P) with the line table of q1*P (the default) and as e(C, P) followed by
    the exponentiation give the same plaintexts and statuses: golden vectors, a random batch with negative values,
    zeros, O and out-of-range plaintexts, byte form and handle form."""
    from bgn_b200 import Engine
    g = load_golden(128)
    v = g["decrypt_l1"]
    rng = np.random.default_rng(5)
    T = g["msg_space"]
    x = rng.integers(-T, T + 1, 600)
    x[:4] = [0, T, -T, 4 * T * T]  # the last one is out of range: status 1
    res = []
    for route in (1, 0):
        e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
        e.set_option("dec_pair_q1", route)
        e.set_secret(int(g["q1"], 16), T)
        vals, st = e.decrypt_batch(buf(v["in"]), False)
        assert list(st) == v["status"] and [int(t) for t in vals] == v["out"], route
        r = rng.integers(0, 256, (len(x), e.scalar_bytes), dtype=np.uint8) if route else r
        r[:, 0] &= 0x3F
        ct = e.encrypt_batch(x, r.reshape(-1))
        ct[5 * e.elem_bytes: 6 * e.elem_bytes] = 0  # O decrypts to 0
        vals, st = e.decrypt_batch(ct, False)
        h = e.import_batch(1, ct)
        vh, sh = e.decrypt_h(h)
        assert (np.asarray(vh) == np.asarray(vals)).all() and (np.asarray(sh) == np.asarray(st)).all()
        h.free()
        res.append((np.asarray(vals).copy(), np.asarray(st).copy()))
        ok = np.asarray(st) == 0
        exp = x.copy()
        exp[5] = 0
        assert (np.asarray(vals)[ok] == exp[ok]).all() and not ok[3] and ok[:3].all() and ok[4:].all(), route
        e.close()
    assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all()


# ---------------------------------------------------------------- non-deterministic mode
def test_g1_blind(golden):
    e, v = engine_for(golden), golden["g1_blind"]
    assert e.g1_blind_batch(buf(v["a"]), scal(e, v["r"], e.scalar_bytes)).tobytes() == unhex(v["out"])


def test_gt_blind(golden):
    e, v = engine_for(golden), golden["gt_blind"]
    assert e.gt_blind_batch(buf(v["a"]), scal(e, v["r"], e.scalar_bytes)).tobytes() == unhex(v["out"])


# ---------------------------------------------------------------- polynomial helpers
@pytest.mark.parametrize("lvl", ["l1", "l2"])
def test_multconstpoly(golden, lvl):
    e, v = engine_for(golden), golden["multconstpoly_" + lvl]
    for case in v["cases"]:
        out = e.multconstpoly_batch(buf(v["in"]), v["d"], lvl == "l2", case["digits"], case["negate"], v["count"])
        assert out.tobytes() == unhex(case["out"]), case["constant"]


@pytest.mark.parametrize("lvl", ["l1", "l2"])
def test_evalpoly(golden, lvl):
    e, v = engine_for(golden), golden["evalpoly_" + lvl]
    out = e.evalpoly_batch(buf(v["in"]), v["d"], lvl == "l2", v["base"], v["count"])
    assert out.tobytes() == unhex(v["out"])


def test_make_poly_l2(golden):
    e, v = engine_for(golden), golden["make_poly_l2"]
    assert e.make_poly_l2_batch(buf(v["in"]), v["d"], v["count"]).tobytes() == unhex(v["out"])


@pytest.mark.parametrize("kb", [128, 512, 1024])
def test_fixed_pairing_lane_pair_equals_one_thread(kb):
    """e(., P) on a pair of lanes per point (k_miller_fixed_pair, pairlane.cuh) gives the bytes of the
    one-thread kernel, of the general Miller kernel and of the oracle: odd counts (a warp with a
    half-used pair and idle pairs), O inputs, level-1 decryption through it."""
    from bgn_b200 import Engine
    from oracle import bgn_oracle as O
    g = load_golden(kb)
    e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    e.set_secret(int(g["q1"], 16), g["msg_space"])
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    P = O.g1_from_bytes(bytes.fromhex(g["P"]), par)
    rng = random.Random(kb)
    count = 77 if kb < 1024 else 19
    xs = np.array([rng.randrange(-40, 40) for _ in range(count)], dtype=np.int64)
    xs[3] = 0
    cts = e.encrypt_batch(xs, e.scalars_be([rng.randrange(par.n) for _ in range(count)]))
    cts[5 * e.elem_bytes: 6 * e.elem_bytes] = 0  # O
    outs = {}
    for mode in (1, 0):
        e.set_option("fixed_pair", mode)
        outs[mode] = e.make_l2_batch(cts).tobytes()
        vals, st = e.decrypt_batch(cts, False)
        assert not st.any() and vals.tolist() == [0 if i == 5 else int(x) for i, x in enumerate(xs)]
    e.set_option("fixed_lines", 0)
    general = e.make_l2_batch(cts).tobytes()
    assert outs[1] == outs[0] == general
    eb = e.elem_bytes
    for i in (0, 3, 5, count - 1):
        pt = O.g1_from_bytes(cts[i * eb:(i + 1) * eb].tobytes(), par)
        assert outs[1][i * eb:(i + 1) * eb] == O.gt_to_bytes(O.pairing(pt, P, par), par)
    v = g["make_poly_l2"]
    e.set_option("fixed_lines", 1)
    e.set_option("fixed_pair", 1)
    assert e.make_poly_l2_batch(buf(v["in"]), v["d"], v["count"]).tobytes() == unhex(v["out"])
    e.close()


@pytest.mark.parametrize("kb", [128, 512, 1024])
def test_general_pairing_two_warps_equals_one_thread(kb):
    """e(a, b) with a pairing split over two warps (k_pair_duo, pairwarp.cuh) gives the bytes of the
    one-thread kernel and of the oracle: counts that leave a warp pair partly empty and span several
    blocks, O on either side, a point paired with itself and with its negative."""
    from bgn_b200 import Engine
    from oracle import bgn_oracle as O
    g = load_golden(kb)
    e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    rng = random.Random(kb + 1)
    count = 75 if kb < 1024 else 37
    eb = e.elem_bytes
    xs = np.array([rng.randrange(-40, 40) for _ in range(count)], dtype=np.int64)
    a = e.encrypt_batch(xs, e.scalars_be([rng.randrange(par.n) for _ in range(count)]))
    b = e.encrypt_batch(xs[::-1].copy(), e.scalars_be([rng.randrange(par.n) for _ in range(count)]))
    a[2 * eb: 3 * eb] = 0                      # O on the Miller side
    b[4 * eb: 5 * eb] = 0                      # O on the evaluation side
    b[6 * eb: 7 * eb] = a[6 * eb: 7 * eb]      # e(A, A)
    b[8 * eb: 9 * eb] = e.g1_neg_batch(a[8 * eb: 9 * eb])  # e(A, -A)
    outs = {}
    for mode in (1, 0):
        e.set_option("pair_duo", mode)
        outs[mode] = e.pair_batch(a, b).tobytes()
    assert outs[1] == outs[0]
    for i in (0, 2, 4, 6, 8, 33, count - 1):
        pa = O.g1_from_bytes(a[i * eb:(i + 1) * eb].tobytes(), par)
        pb = O.g1_from_bytes(b[i * eb:(i + 1) * eb].tobytes(), par)
        assert outs[1][i * eb:(i + 1) * eb] == O.gt_to_bytes(O.pairing(pa, pb, par), par)
    e.set_option("pair_duo", 1)
    v = g["pair"]
    assert e.pair_batch(buf(v["a"]), buf(v["b"])).tobytes() == unhex(v["out"])
    assert e.pair_batch(buf(v["b"]), buf(v["a"])).tobytes() == unhex(v["out"])
    e.close()


@pytest.mark.parametrize("kb", [128, 512])
def test_device_resident_handles_chain(kb):
    """Encrypt -> EAdd -> EMult (MultPoly) -> L2 sum -> Decrypt on device-resident handles (bgn_buf,
    `*_h` entry points) gives the bytes of the byte-format entry points at every stage, decrypts to the
    plaintext result, and launches no (de)serialisation kernel between the two edges."""
    from bgn_b200 import Engine
    g = load_golden(kb)
    e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    e.set_secret(int(g["q1"], 16), g["msg_space"])
    rng = random.Random(kb + 5)
    n = int(g["n"], 16)
    count, d = 9, 3
    xa = np.array([rng.randrange(-1, 2) for _ in range(count * d)], dtype=np.int64)
    xb = np.array([rng.randrange(-1, 2) for _ in range(count * d)], dtype=np.int64)
    xc = np.array([rng.randrange(-1, 2) for _ in range(count * d)], dtype=np.int64)
    ra, rb, rc = (e.scalars_be([rng.randrange(n) for _ in range(count * d)]) for _ in range(3))
    # byte path
    A, B_, C_ = e.encrypt_batch(xa, ra), e.encrypt_batch(xb, rb), e.encrypt_batch(xc, rc)
    S = e.g1_add_batch(A, B_)
    Dif = e.g1_sub_batch(A, B_)
    Pr = e.multpoly_batch(S, d, C_, d, count)
    Tot = e.l2_sum_reduce(Pr, count, 2 * d)
    vals, st = e.decrypt_batch(Tot, True)
    # handle path
    e.timing_enable(True)
    e.timing_reset()
    hA, hB, hC = e.encrypt_h(xa, ra), e.encrypt_h(xb, rb), e.encrypt_h(xc, rc)
    hS = e.g1_add_h(hA, hB)
    hD = e.g1_add_h(hA, hB, subtract=True)
    hP = e.multpoly_h(hS, d, hC, d, count)
    hT = e.l2_sum_reduce_h(hP, count, 2 * d)
    hv, hs = e.decrypt_h(hT)
    for k in ("k_g1_from_bytes", "k_g1_to_bytes", "k_fp2_from_bytes", "k_fp2_to_bytes"):
        assert e.timing_get(k)[1] == 0, k + " ran inside the handle chain"
    assert (len(hS), hS.kind, len(hP), hP.kind, len(hT)) == (count * d, 1, count * 2 * d, 2, 2 * d)
    assert hA.to_bytes().tobytes() == A.tobytes() and hS.to_bytes().tobytes() == S.tobytes()
    assert hD.to_bytes().tobytes() == Dif.tobytes()
    assert hP.to_bytes().tobytes() == Pr.tobytes() and hT.to_bytes().tobytes() == Tot.tobytes()
    assert hv.tolist() == vals.tolist() and not hs.any() and not st.any()
    plain = np.sum([np.convolve((xa + xb)[u * d:(u + 1) * d], xc[u * d:(u + 1) * d]) for u in range(count)], axis=0)
    assert hv.tolist() == plain.tolist() + [0]
    # import / export round trip, level-1 decrypt and plain pairings on handles, output reuse
    imp = e.import_batch(1, S)
    assert imp.to_bytes().tobytes() == S.tobytes()
    v1, s1 = e.decrypt_h(imp)
    assert not s1.any() and v1.tolist() == (xa + xb).tolist()
    hM = e.pair_h(hA, hB)
    assert hM.to_bytes().tobytes() == e.pair_batch(A, B_).tobytes()
    assert e.pair_h(hA).to_bytes().tobytes() == e.make_l2_batch(A).tobytes()
    assert e.gt_mul_h(hM, hM).to_bytes().tobytes() == e.gt_mul_batch(hM.to_bytes(), hM.to_bytes()).tobytes()
    one = b"\x00" * (e.coord_bytes - 1) + b"\x01" + b"\x00" * e.coord_bytes
    assert e.gt_mul_h(hM, hM, divide=True).to_bytes().tobytes() == one * (count * d)
    reuse = e.g1_add_h(hA, hB, out=hD)
    assert reuse is hD and hD.to_bytes().tobytes() == S.tobytes()
    with pytest.raises(Exception):
        e.g1_add_h(hA, hM)  # wrong group
    with pytest.raises(Exception):
        e.g1_add_h(hA, hB, out=hA)  # output aliases an operand
    for h in (hA, hB, hC, hS, hD, hP, hT, imp, hM):
        h.free()
    e.close()


@pytest.mark.parametrize("kb,count,d1,d2", [(512, 40, 11, 11), (512, 33, 3, 7), (256, 50, 5, 4), (128, 300, 2, 2),
                                            (128, 20, 6, 11), (64, 1500, 3, 3)])
def test_multpoly_split_team_kernel_equals_team_kernel(kb, count, d1, d2):
    """MultPoly on the split team kernel (k_miller_split, teamsplit.cuh: two threads per output-slot pair)
    gives the bytes of the team kernel and of the automatic policy, with O coefficients sprinkled in;
    one product is also checked against the oracle."""
    from bgn_b200 import Engine
    from oracle import bgn_oracle as O
    g = load_golden(kb)
    e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    rng = random.Random(kb + count)
    eb = e.elem_bytes
    x1 = np.array([rng.randrange(-1, 2) for _ in range(count * d1)], dtype=np.int64)
    x2 = np.array([rng.randrange(-1, 2) for _ in range(count * d2)], dtype=np.int64)
    c1 = e.encrypt_batch(x1, e.scalars_be([rng.randrange(par.n) for _ in range(count * d1)]))
    c2 = e.encrypt_batch(x2, e.scalars_be([rng.randrange(par.n) for _ in range(count * d2)]))
    c1[eb: 2 * eb] = 0
    c2[(d2 + 1) * eb: (d2 + 2) * eb] = 0
    outs = {}
    e.timing_enable(True)
    for mode in (0, 1, -1):
        e.set_option("miller_split", mode)
        e.timing_reset()
        outs[mode] = e.multpoly_batch(c1, d1, c2, d2, count).tobytes()
        if mode == 1:
            assert e.timing_get("k_miller_split")[1] >= 1, "the split kernel did not run"
    assert outs[1] == outs[0] == outs[-1]
    pk = O.PublicKey(par, O.g1_from_bytes(bytes.fromhex(g["P"]), par), O.g1_from_bytes(bytes.fromhex(g["Q"]), par), 0)
    u = count - 1
    A = [O.Ciphertext(O.g1_from_bytes(c1[(u * d1 + i) * eb:(u * d1 + i + 1) * eb].tobytes(), par), False) for i in range(d1)]
    Bq = [O.Ciphertext(O.g1_from_bytes(c2[(u * d2 + i) * eb:(u * d2 + i + 1) * eb].tobytes(), par), False) for i in range(d2)]
    exp = O.mult_poly(pk, O.PolyCiphertext(A, d1, 0, False), O.PolyCiphertext(Bq, d2, 0, False))
    assert outs[1][u * (d1 + d2) * eb:] == O.poly_ct_bytes(pk, exp)
    e.close()


@pytest.mark.parametrize("kb,count,d1,d2,tpb", [(512, 70, 11, 11, 28), (256, 90, 5, 4, 7), (128, 64, 1, 1, 20), (64, 500, 3, 3, 50)])
def test_multpoly_wide_team_kernel_equals_team_kernel(kb, count, d1, d2, tpb):
    """MultPoly on the wide team kernel (k_miller_wide: evaluation points read from the batch arrays, 10
    shared-memory slots per thread, up to 10 warps per block) gives the team kernel's bytes, with O
    coefficients on both sides."""
    from bgn_b200 import Engine
    g = load_golden(kb)
    e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    rng = random.Random(kb + count + 1)
    n = int(g["n"], 16)
    eb = e.elem_bytes
    x1 = np.array([rng.randrange(-1, 2) for _ in range(count * d1)], dtype=np.int64)
    x2 = np.array([rng.randrange(-1, 2) for _ in range(count * d2)], dtype=np.int64)
    c1 = e.encrypt_batch(x1, e.scalars_be([rng.randrange(n) for _ in range(count * d1)]))
    c2 = e.encrypt_batch(x2, e.scalars_be([rng.randrange(n) for _ in range(count * d2)]))
    c1[0:eb] = 0
    c2[(count * d2 - 1) * eb:] = 0
    e.timing_enable(True)
    e.set_option("miller_split", 0)
    e.set_option("pair_duo", 0)
    ref = e.multpoly_batch(c1, d1, c2, d2, count).tobytes()
    e.set_option("miller_wide", tpb)
    e.timing_reset()
    got = e.multpoly_batch(c1, d1, c2, d2, count).tobytes()
    assert e.timing_get("k_miller_wide")[1] == 1, "the wide kernel did not run"
    assert got == ref
    e.close()


@pytest.mark.parametrize("kb", [128, 512])
def test_group_of_contexts_equals_single_context(kb):
    """bgn_group (several GPUs behind one handle, one host thread per device inside each call): every GPU of the
    box when there are several, and in any case three contexts taking turns on device 0 -- an uneven split
    with an empty shard (2 units over 3 members).  Sharded Encrypt / EAdd / MultPoly / Decrypt give the
    single-context bytes; the inner product equals MultPoly + L2 sum on one context and decrypts to the
    plaintext result."""
    import torch
    from bgn_b200 import Engine, EngineGroup
    g = load_golden(kb)
    key = (int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    one = Engine(*key)
    one.set_secret(int(g["q1"], 16), g["msg_space"])
    rng = random.Random(kb + 9)
    n = key[1]
    layouts = [[0, 0, 0]]
    if torch.cuda.device_count() >= 2:
        layouts.append(list(range(torch.cuda.device_count())))
    for devs in layouts:
        grp = EngineGroup(*key, devices=devs)
        grp.set_secret(int(g["q1"], 16), g["msg_space"])
        for count, d in ((2, 3), (29, 3)):
            xa = np.array([rng.randrange(-1, 2) for _ in range(count * d)], dtype=np.int64)
            xb = np.array([rng.randrange(-1, 2) for _ in range(count * d)], dtype=np.int64)
            ra = one.scalars_be([rng.randrange(n) for _ in range(count * d)])
            rb = one.scalars_be([rng.randrange(n) for _ in range(count * d)])
            A, Bq = grp.encrypt_batch(xa, ra), grp.encrypt_batch(xb, rb)
            assert A.tobytes() == one.encrypt_batch(xa, ra).tobytes() and Bq.tobytes() == one.encrypt_batch(xb, rb).tobytes()
            assert grp.g1_add_batch(A, Bq).tobytes() == one.g1_add_batch(A, Bq).tobytes()
            prod = grp.multpoly_batch(A, d, Bq, d, count)
            assert prod.tobytes() == one.multpoly_batch(A, d, Bq, d, count).tobytes()
            total = grp.inner_product(A, d, Bq, d, count)
            assert total.tobytes() == one.l2_sum_reduce(prod, count, 2 * d).tobytes()
            vals, st = grp.decrypt_batch(total, True)
            plain = np.sum([np.convolve(xa[u * d:(u + 1) * d], xb[u * d:(u + 1) * d]) for u in range(count)], axis=0)
            assert not st.any() and vals.tolist() == plain.tolist() + [0]
            v1, s1 = grp.decrypt_batch(A, False)
            assert not s1.any() and v1.tolist() == xa.tolist()
        grp.close()
    one.close()
    with pytest.raises(Exception):
        EngineGroup(*key, devices=[0, 99])


def test_empty_batches(golden):
    e = engine_for(golden)
    z = np.zeros(0, dtype=np.uint8)
    assert e.g1_add_batch(z, z).size == 0
    assert e.pair_batch(z, z).size == 0
    assert e.encrypt_batch(np.zeros(0, dtype=np.int64), None).size == 0
    vals, st = e.decrypt_batch(z, True)
    assert vals.size == 0 and st.size == 0


def test_decrypt_before_setup():
    from bgn_b200 import BgnError, Engine
    g = load_golden(64)
    e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    with pytest.raises(BgnError) as ei:
        e.decrypt_batch(buf(g["decrypt_l2"]["in"]), True)
    assert ei.value.status == -3 and "DL tables not computed" in str(ei.value)
    e.close()


def test_bad_params_rejected():
    from bgn_b200 import BgnError, Engine
    g = load_golden(64)
    with pytest.raises(BgnError):  # p + 1 != l * n
        Engine(int(g["p"], 16), int(g["n"], 16), g["l"] + 4, bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    bad = bytearray(bytes.fromhex(g["P"]))
    bad[-1] ^= 1
    with pytest.raises(BgnError):  # generator not on the curve
        Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes(bad), bytes.fromhex(g["Q"]))


def test_off_curve_input_becomes_O(golden):
    """curve_from_bytes: a pair not on the curve is O."""
    e, v = engine_for(golden), golden["g1_add"]
    a = bytearray(bytes.fromhex(v["a"][0]))
    a[-1] ^= 1
    out = e.g1_add_batch(np.frombuffer(bytes(a), dtype=np.uint8), buf([v["b"][0]]))
    assert out.tobytes() == bytes.fromhex(v["b"][0])


def test_generators_of_even_order_keep_the_complete_addition():
    """A context whose Q is not of odd order (Q + the point of order 2: on the curve, so it is accepted) must not
    use the Edwards form, whose unified addition is only free of exceptional cases on odd-order points: Encrypt
    and the re-randomisation still equal the oracle's x*P + r*Q computed with the complete affine law."""
    from bgn_b200 import Engine
    from oracle import bgn_oracle as O
    g = load_golden(128)
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    P = O.g1_from_bytes(bytes.fromhex(g["P"]), par)
    Q = O.g1_from_bytes(bytes.fromhex(g["Q"]), par)
    Q2 = O.g1_add(Q, (0, 0), par.p)
    assert Q2 is not None and O.g1_mul(par.n, Q2, par.p) == (0, 0)
    e = Engine(par.p, par.n, par.l, bytes.fromhex(g["P"]), O.g1_to_bytes(Q2, par))
    rng = random.Random(77)
    xs = [rng.randrange(-2, 3) for _ in range(12)]
    rs = [rng.randrange(par.n) for _ in range(12)]
    rs[0], rs[1], rs[2] = 0, 1, 2
    exp = b""
    for x, r in zip(xs, rs):
        c = O.g1_add(O.g1_mul(abs(x), P, par.p), O.g1_mul(r, Q2, par.p), par.p)
        exp += O.g1_to_bytes(O.g1_neg(c, par.p) if x < 0 else c, par)
    for bits in (0, 16, 8):
        e.set_option("enc_window", bits)
        assert e.encrypt_batch(np.array(xs, dtype=np.int64), e.scalars_be(rs, e.scalar_bytes)).tobytes() == exp, bits
    e.close()


# ---------------------------------------------------------------- seeded random parity vs the oracle
@pytest.mark.parametrize("kb", [128, 512])
def test_random_vs_oracle(kb):
    from oracle import bgn_oracle as O
    g = load_golden(kb)
    e = engine_for(g)
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    P = O.g1_from_bytes(bytes.fromhex(g["P"]), par)
    Q = O.g1_from_bytes(bytes.fromhex(g["Q"]), par)
    pk = O.PublicKey(par, P, Q, g["msg_space"])
    rng = random.Random(kb)
    cnt = 24 if kb == 128 else 6
    xs = [rng.randrange(-3, 4) for _ in range(cnt)]
    rs = [rng.randrange(par.n) for _ in range(cnt)]
    exp = []
    for x, r in zip(xs, rs):
        c = O.encrypt_with_randomness(pk, abs(x), r).C
        exp.append(O.g1_neg(c, par.p) if x < 0 else c)
    got = e.encrypt_batch(np.array(xs, dtype=np.int64), e.scalars_be(rs))
    assert got.tobytes() == b"".join(O.g1_to_bytes(c, par) for c in exp)
    # pairings of consecutive ciphertexts
    a, b = got[: (cnt - 1) * e.elem_bytes], got[e.elem_bytes:]
    pe = [O.pairing(exp[i], exp[i + 1], par) for i in range(cnt - 1)]
    assert e.pair_batch(a.copy(), b.copy()).tobytes() == b"".join(O.gt_to_bytes(v, par) for v in pe)


# ---------------------------------------------------------------- size-independent properties at scale
def test_properties_512_scale():
    """keyBits=512, a few thousand elements: encrypt -> EMult -> decrypt recovers the product
    polynomial; bilinearity e(aP, bP) = e(P,P)^(ab); commutativity of the batch product."""
    g = load_golden(512)
    e = engine_for(g)
    n = int(g["n"], 16)
    rng = random.Random(7)
    count, d = 256, 4
    c1 = np.array([[rng.randrange(-1, 2) for _ in range(d)] for _ in range(count)], dtype=np.int64)
    c2 = np.array([[rng.randrange(-1, 2) for _ in range(d)] for _ in range(count)], dtype=np.int64)
    r1 = e.scalars_be([rng.randrange(n) for _ in range(count * d)])
    r2 = e.scalars_be([rng.randrange(n) for _ in range(count * d)])
    E1 = e.encrypt_batch(c1.reshape(-1), r1)
    E2 = e.encrypt_batch(c2.reshape(-1), r2)
    prod = e.multpoly_batch(E1, d, E2, d, count)
    assert prod.tobytes() == e.multpoly_batch(E2, d, E1, d, count).tobytes()
    vals, st = e.decrypt_batch(prod, True)
    assert not st.any()
    vals = vals.reshape(count, 2 * d)
    for u in range(count):
        exp = np.convolve(c1[u], c2[u]).tolist() + [0]
        assert vals[u].tolist() == exp
    # L1 decrypt of the inputs
    v1, s1 = e.decrypt_batch(E1, False)
    assert not s1.any() and v1.tolist() == c1.reshape(-1).tolist()
    # inner product: sum over units, slot-wise
    total = e.l2_sum_reduce(prod, count, 2 * d)
    tv, ts = e.decrypt_batch(total, True)
    assert not ts.any()
    assert tv.tolist() == sum(np.convolve(c1[u], c2[u]) for u in range(count)).tolist() + [0]


def test_device_pointer_io():
    """device-resident buffers (CUDA tensors) give the same bytes as host buffers."""
    import torch
    g = load_golden(128)
    e, v = engine_for(g), g["pair"]
    a = torch.from_numpy(buf(v["a"]).copy()).cuda()
    b = torch.from_numpy(buf(v["b"]).copy()).cuda()
    out = e.pair_batch(a, b)
    assert out.is_cuda and out.cpu().numpy().tobytes() == unhex(v["out"])


def test_mirror_api_truth_table():
    """cmd/main.go:74-107: Add / Mult / Neg over {0, 1, -1} through the Go-named host mirror."""
    from bgn_b200 import PublicKey, SecretKey
    g = load_golden(128)
    pk = PublicKey.FromPBCParams(g["pbc_params"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), g["msg_space"])
    sk = SecretKey(int(g["q1"], 16))
    pk.SetupDecryption(sk)
    cts = {m: pk.Encrypt(m) for m in (0, 1, -1)}
    for a in (0, 1, -1):
        assert sk.Decrypt(pk.Neg(cts[a]), pk) == -a
        for b in (0, 1, -1):
            assert sk.Decrypt(pk.Add(cts[a], cts[b]), pk) == a + b
            assert sk.Decrypt(pk.Sub(cts[a], cts[b]), pk) == a - b
            assert sk.Decrypt(pk.Mult(cts[a], cts[b]), pk) == a * b
            assert sk.Decrypt(pk.Add(pk.Mult(cts[a], cts[b]), cts[a]), pk) == a * b + a
    assert sk.Decrypt(pk.MultConst(cts[1], 7), pk) == 7
    assert sk.Decrypt(pk.MultConst(pk.makeL2(cts[-1]), 9), pk) == -9


def test_mirror_api_poly():
    """poly_test.go:92-189 with the reference's constants (POLYBASE=3, FPSCALEBASE=3, FPPREC=0.0001)."""
    from bgn_b200 import PublicKey, SecretKey
    g = load_golden(128)
    pk = PublicKey.FromPBCParams(g["pbc_params"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), g["msg_space"])
    sk = SecretKey(int(g["q1"], 16))
    pk.SetupDecryption(sk)
    f1 = lambda x: "%.1f" % x  # noqa: E731
    c = pk.EncryptPoly(pk.NewPolyPlaintext(9.123))
    assert f1(sk.DecryptPoly(c, pk).PolyEval()) == "9.1"
    a, b = pk.EncryptPoly(pk.NewPolyPlaintext(0.1)), pk.EncryptPoly(pk.NewPolyPlaintext(4.2))
    assert f1(sk.DecryptPoly(pk.AddPoly(a, b), pk).PolyEval()) == "4.3"
    a, b = pk.EncryptPoly(pk.NewPolyPlaintext(50.1)), pk.EncryptPoly(pk.NewPolyPlaintext(41.2))
    s = pk.AddPoly(pk.MakePolyL2(a), pk.MakePolyL2(b))
    assert s.L2 and f1(sk.DecryptPoly(s, pk).PolyEval()) == "91.3"
    a = pk.EncryptPoly(pk.NewPolyPlaintext(9.13))
    assert f1(sk.DecryptPoly(pk.MultConstPoly(a, 4.12), pk).PolyEval()) == f1(9.13 * 4.12)
    assert f1(sk.DecryptPoly(pk.MultConstPoly(pk.MakePolyL2(a), 4.12), pk).PolyEval()) == f1(9.13 * 4.12)
    a, b = pk.EncryptPoly(pk.NewPolyPlaintext(1.1)), pk.EncryptPoly(pk.NewPolyPlaintext(40.2))
    assert f1(sk.DecryptPoly(pk.MultPoly(a, b), pk).PolyEval()) == f1(1.1 * 40.2)
    assert f1(sk.DecryptPoly(pk.SubPoly(b, a), pk).PolyEval()) == f1(40.2 - 1.1)


def test_mirror_api_nondeterministic():
    """Deterministic=false: every homomorphic op re-randomises (bgn.go:260-269, 279-288, 302-311,
    466-474, 488-495).  With the randomness injected the bytes must equal the oracle's; with fresh
    randomness the plaintext algebra must still hold and ciphertexts must differ between calls."""
    from bgn_b200 import PublicKey, SecretKey
    from oracle import bgn_oracle as O
    g = load_golden(128)
    pk = PublicKey.FromPBCParams(g["pbc_params"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), g["msg_space"],
                                 Deterministic=False)
    sk = SecretKey(int(g["q1"], 16))
    pk.SetupDecryption(sk)
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    opk = O.PublicKey(par, O.g1_from_bytes(bytes.fromhex(g["P"]), par), O.g1_from_bytes(bytes.fromhex(g["Q"]), par),
                      g["msg_space"], deterministic=False)
    rng = random.Random(3)
    r = lambda: rng.randrange(par.n)  # noqa: E731
    r1, r2, r3, r4, r5, r6 = (r() for _ in range(6))
    a, b = pk.EncryptWithRandomness(3, r1), pk.EncryptWithRandomness(2, r2)
    oa, ob = O.encrypt_with_randomness(opk, 3, r1), O.encrypt_with_randomness(opk, 2, r2)
    assert a.C == O.ct_bytes(opk, oa) and b.C == O.ct_bytes(opk, ob)
    assert pk.Add(a, b, r=r3).C == O.ct_bytes(opk, O.add(opk, oa, ob, r3))
    assert pk.Sub(a, b, r=r4).C == O.ct_bytes(opk, O.sub(opk, oa, ob, r4))
    m, om = pk.Mult(a, b, r=r5), O.mult(opk, oa, ob, r5)
    assert m.L2 and m.C == O.ct_bytes(opk, om)
    assert pk.Add(m, m, r=r6).C == O.ct_bytes(opk, O.add(opk, om, om, r6))
    assert pk.MultConst(a, 5, r=r3).C == O.ct_bytes(opk, O.mult_const(opk, oa, 5, r3))
    assert pk.MultConst(m, 5, r=r4).C == O.ct_bytes(opk, O.mult_const(opk, om, 5, r4))
    # fresh randomness: same plaintexts, different ciphertexts
    s1, s2 = pk.Add(a, b), pk.Add(a, b)
    assert s1.C != s2.C and sk.Decrypt(s1, pk) == sk.Decrypt(s2, pk) == 5
    p1, p2 = pk.Mult(a, b), pk.Mult(a, b)
    assert p1.C != p2.C and sk.Decrypt(p1, pk) == sk.Decrypt(p2, pk) == 6
    assert sk.Decrypt(pk.Sub(p1, pk.makeL2(a)), pk) == 3


@pytest.mark.parametrize("kb", [64, 128])
def test_mirror_api_nondeterministic_poly_ops(kb):
    """Deterministic=false at the polynomial level: MultPoly / AddPoly / SubPoly / NegPoly / MultConstPoly /
    MakePolyL2 / EvalPoly / EncryptPoly and their *Batch forms on the CUDA engine, driven by a replayed
    randomness stream, byte for byte against the oracle's literal poly.go control flow on the same stream
    (tests/nondet_cases.py); two un-injected calls must differ and decrypt alike."""
    from bgn_b200 import PublicKey, SecretKey
    from nondet_cases import run_poly_cases
    from oracle import bgn_oracle as O
    g = load_golden(kb)
    pk = PublicKey.FromPBCParams(g["pbc_params"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), g["msg_space"])
    sk = SecretKey(int(g["q1"], 16))
    pk.SetupDecryption(sk)
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    opk = O.PublicKey(par, O.g1_from_bytes(bytes.fromhex(g["P"]), par), O.g1_from_bytes(bytes.fromhex(g["Q"]), par),
                      g["msg_space"])
    run_poly_cases(pk, opk, sk)
    pk.engine.close()


@pytest.mark.parametrize("kb", [64, 128, 256, 512])
def test_mirror_api_nondeterministic_poly_golden(kb):
    """the committed non-deterministic polynomial fixtures (tests/golden/kb*.json["nondet_poly"]: oracle's
    literal poly.go control flow on a recorded randomness stream) through the mirror on the CUDA engine"""
    from bgn_b200 import PublicKey
    from nondet_cases import run_golden_section
    g = load_golden(kb)
    pk = PublicKey.FromPBCParams(g["pbc_params"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), g["msg_space"])
    run_golden_section(pk, g)
    pk.engine.close()


def test_mirror_wire_roundtrip():
    """bgn_test.go:37-85 (TestMarshalUnmarshal*): ciphertext -> gob envelope -> ciphertext keeps the
    element, level, Degree and ScaleFactor, on both levels; malformed element bytes follow
    Element.SetBytes (off-curve -> O)."""
    from bgn_b200 import PublicKey, SecretKey
    g = load_golden(128)
    pk = PublicKey.FromPBCParams(g["pbc_params"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), g["msg_space"])
    sk = SecretKey(int(g["q1"], 16))
    pk.SetupDecryption(sk)
    c = pk.Encrypt(9)
    c2 = pk.NewCiphertextFromBytes(c.Bytes())
    assert (c2.C, c2.L2) == (c.C, False) and sk.Decrypt(c2, pk) == 9
    m = pk.Mult(c, c)
    m2 = pk.NewCiphertextFromBytes(m.Bytes())
    assert (m2.C, m2.L2) == (m.C, True) and sk.Decrypt(m2, pk) == 81
    pc = pk.EncryptPoly(pk.NewPolyPlaintext(9.123))
    for poly in (pc, pk.MakePolyL2(pc)):
        back = pk.NewPolyCiphertextFromBytes(poly.Bytes())
        assert back.CoeffBytes() == poly.CoeffBytes()
        assert (back.Degree, back.ScaleFactor, back.L2) == (poly.Degree, poly.ScaleFactor, poly.L2)
    bad = bytearray(c.C)
    bad[-1] ^= 1
    from bgn_b200 import gobwire
    o = pk.NewCiphertextFromBytes(gobwire.encode_ciphertext(bytes(bad), False))
    assert o.C == bytes(pk.elem_bytes)
    with pytest.raises(ValueError):
        pk.NewCiphertextFromBytes(b"")
    # the public key itself (bgn.go:597-666): marshal, load onto the GPU again, same ciphertext bytes
    pk2 = PublicKey.UnmarshalBinary(pk.MarshalBinary())
    assert (pk2.N, pk2.MsgSpace, pk2.Deterministic, pk2.PairingParams) == (pk.N, pk.MsgSpace, True, pk.PairingParams)
    assert pk2.EncryptWithRandomness(5, 12345).C == pk.EncryptWithRandomness(5, 12345).C
    pk2.engine.close()


def test_batch_poly_helpers_match_mirror():
    """MultConstPolyBatch / EvalPolyBatch / MakePolyL2Batch (one kernel launch per batch) give the
    same bytes as the per-polynomial mirror methods composed from the scalar C-ABI primitives."""
    from bgn_b200 import PublicKey, SecretKey
    from bgn_b200.bgn import PolyCiphertextBatch
    g = load_golden(128)
    pk = PublicKey.FromPBCParams(g["pbc_params"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), g["msg_space"])
    sk = SecretKey(int(g["q1"], 16))
    pk.SetupDecryption(sk)
    rng = random.Random(11)
    vals = [9.13, 4.0, 17.5, 0.0]
    pts = [pk.NewPolyPlaintext(v) for v in vals]
    d = max(p.Degree for p in pts)
    coeffs = np.array([p.Coefficients[: p.Degree] + [0] * (d - p.Degree) for p in pts], dtype=np.int64)
    rs = [[rng.randrange(pk.N) for _ in range(d)] for _ in pts]
    batch = pk.EncryptPolyBatch(coeffs, pk.engine.scalars_be([r for row in rs for r in row]), ScaleFactor=0)
    singles = []
    for p, row in zip(pts, rs):
        pp = type(p)(p.Coefficients[: p.Degree] + [0] * (d - p.Degree), d, p.ScaleFactor, p.params)
        singles.append(pk.EncryptPoly(pp, rs=row))
    assert batch.data.tobytes() == b"".join(s.CoeffBytes() for s in singles)
    for lvl2 in (False, True):
        bb = pk.MakePolyL2Batch(batch) if lvl2 else batch
        ss = [pk.MakePolyL2(s) for s in singles] if lvl2 else singles
        assert bb.data.tobytes() == b"".join(s.CoeffBytes() for s in ss)
        for constant in (4.12, -3.0):
            got = pk.MultConstPolyBatch(bb, constant)
            exp = [pk.MultConstPoly(s, constant) for s in ss]
            assert got.Degree == exp[0].Degree and got.data.tobytes() == b"".join(x.CoeffBytes() for x in exp)
        ev = pk.EvalPolyBatch(bb)
        assert ev.tobytes() == b"".join(pk.EvalPoly(s).C for s in ss)
    # decrypt the evaluated polynomials: EvalPoly recovers the encoded integer
    evals, st = pk.engine.decrypt_batch(pk.EvalPolyBatch(batch), False)
    for v, ok, val in zip(evals, st, vals):
        if val == int(val):  # integer plaintexts are below MsgSpace
            assert not ok and int(v) == int(val)


# ---------------------------------------------------------------- mid-size parity vs the C port of the oracle
@pytest.mark.parametrize("kb,count,d1,d2", [(512, 48, 11, 11), (512, 64, 3, 7), (1024, 6, 8, 8), (256, 200, 5, 4),
                                            (256, 9, 1, 7), (256, 7, 40, 5), (128, 5, 32, 32), (64, 3, 128, 128),
                                            (512, 300, 1, 1), (512, 24, 13, 13)])
def test_multpoly_vs_cpu_ref(kb, count, d1, d2):
    """Encrypt + EMult at BASELINE.json's shapes (d = 11 at 512 bit, d = 8 at 1024 bit) and at the shape of the
    reference's own BenchmarkMultPoly (plaintext 100.1 -> 13 slots -> 169 pairings, poly_test.go:55-66) on sizes
    the multi-threaded C oracle finishes in seconds; includes zero digits with r = 0 (O coefficients)."""
    import os
    from oracle.cpu_ref import CpuRef
    g = load_golden(kb)
    e = engine_for(g)
    R = CpuRef(int(g["p"], 16), int(g["n"], 16), g["l"], threads=min(16, os.cpu_count() or 1))
    rng = np.random.default_rng(kb + count)
    P, Q = bytes.fromhex(g["P"]), bytes.fromhex(g["Q"])

    def batch(d):
        x = rng.integers(-1, 2, count * d)
        r = rng.integers(0, 256, (count * d, e.scalar_bytes), dtype=np.uint8)
        r[:, 0] &= 0x3F
        r[::7] = 0  # every 7th coefficient deterministic: zero digits there are the point at infinity
        got = e.encrypt_batch(x, r.reshape(-1))
        exp = R.encrypt_batch(P, Q, x, r.reshape(-1))
        assert got.tobytes() == exp.tobytes()
        return got

    c1, c2 = batch(d1), batch(d2)
    assert e.multpoly_batch(c1, d1, c2, d2, count).tobytes() == R.multpoly_batch(c1, d1, c2, d2, count).tobytes()


def test_full_size_properties_emult():
    """BASELINE.json config 3 at FULL size (2^14 pairs, d = 11): the oracle cannot follow, so check
    size-independent properties: commutativity, decrypt == plaintext convolution on a sample of the
    batch, and the L2 sum of the whole batch decrypting to the plaintext inner product."""
    import torch
    g = load_golden(512)
    e = engine_for(g)
    count, d = 1 << 14, 11
    dev = "cuda:0"
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)

    def batch():
        x = torch.randint(-1, 2, (count * d,), generator=gen, device=dev, dtype=torch.int64)
        r = torch.randint(0, 256, (count * d, e.scalar_bytes), generator=gen, device=dev, dtype=torch.uint8)
        r[:, 0] &= 0x3F
        return x, e.encrypt_batch(x, r.reshape(-1))

    x1, c1 = batch()
    x2, c2 = batch()
    prod = e.multpoly_batch(c1, d, c2, d, count)
    assert torch.equal(prod, e.multpoly_batch(c2, d, c1, d, count))
    EB = e.elem_bytes
    sample = [0, 1, 4097, count - 1]
    sel = torch.cat([prod[u * 2 * d * EB:(u + 1) * 2 * d * EB] for u in sample])
    vals, st = e.decrypt_batch(sel, True)
    assert not st.any().item()
    vals = vals.cpu().numpy().reshape(len(sample), 2 * d)
    a1 = x1.cpu().numpy().reshape(count, d)
    a2 = x2.cpu().numpy().reshape(count, d)
    for row, u in zip(vals, sample):
        assert row.tolist() == np.convolve(a1[u], a2[u]).tolist() + [0]
    # inner product over the whole batch: per-slot sums are bounded by 11 * 2^14 < T = 2^20
    total = e.l2_sum_reduce(prod, count, 2 * d)
    tv, ts = e.decrypt_batch(total, True)
    assert not ts.any().item()
    exp = np.zeros(2 * d, dtype=np.int64)
    for k in range(d):
        for j in range(d):
            exp[k + j] += int((a1[:, k] * a2[:, j]).sum())
    assert tv.cpu().numpy().tolist() == exp.tolist()


def test_full_size_properties_encrypt_eadd():
    """BASELINE.json config 2 at FULL size: 2^16 plaintexts x 11 balanced base-3 digits = 720 896
    coefficient encryptions with r < n, then 2^15 pairwise AddPoly.  Size-independent properties, all
    bit-exact: (i) E(x1, r1) + E(x2, r2) == E(x1 + x2, r1 + r2 mod n) computed by a second Encrypt;
    (ii) Sub undoes Add; (iii) re-randomising with r and then with n - r is the identity;
    (iv) a sample of the sums decrypts to the digit sums."""
    import torch
    g = load_golden(512)
    e = engine_for(g)
    n = int(g["n"], 16)
    count, d = 1 << 16, 11
    dev = "cuda:0"
    gen = torch.Generator(device=dev)
    gen.manual_seed(2)
    SB, EB = e.scalar_bytes, e.elem_bytes
    half = count * d // 2
    x = torch.randint(-1, 2, (count * d,), generator=gen, device=dev, dtype=torch.int64)
    # randomness below 2^(8 SB - 2) <= n / 2: r1 + r2 needs no reduction mod n, so it can be formed bytewise
    r = torch.randint(0, 256, (count * d, SB), generator=gen, device=dev, dtype=torch.uint8)
    r[:, 0] &= 0x1F
    c = e.encrypt_batch(x, r.reshape(-1))
    a, b = c[: half * EB], c[half * EB:]
    s = e.g1_add_batch(a, b)
    # (i) homomorphism against a direct encryption of the sums (non-negative digits only: the sign
    # convention -(|x| P + r Q) makes E(x1) + E(x2) = E(x1 + x2, r1 + r2) hold for x1, x2 >= 0)
    xa, xb = x[:half].abs(), x[half:].abs()
    ra = r[:half].to(torch.int32)
    rb = r[half:].to(torch.int32)
    acc = ra + rb
    for col in range(SB - 1, 0, -1):  # propagate byte carries, most significant byte is column 0
        carry = acc[:, col] >> 8
        acc[:, col] &= 0xFF
        acc[:, col - 1] += carry
    assert int(acc[:, 0].max().item()) < 256
    pos = e.encrypt_batch(xa, r[:half].reshape(-1))
    pos2 = e.encrypt_batch(xb, r[half:].reshape(-1))
    direct = e.encrypt_batch(xa + xb, acc.to(torch.uint8).reshape(-1))
    assert torch.equal(e.g1_add_batch(pos, pos2), direct)
    # (ii) Sub undoes Add
    assert torch.equal(e.g1_sub_batch(s, b), a)
    # (iii) blind by r, then by n - r
    m = 1 << 15
    rr = r[:m]
    nr = e.scalars_be([n - int.from_bytes(bytes(row), "big") for row in rr.cpu().numpy()])
    bl = e.g1_blind_batch(s[: m * EB], rr.reshape(-1))
    assert not torch.equal(bl, s[: m * EB])
    assert torch.equal(e.g1_blind_batch(bl, torch.from_numpy(nr).to(dev)), s[: m * EB])
    # (iv) a sample decrypts to the plaintext sums
    idx = [0, 1, 2, 12345, half - 1]
    sel = torch.cat([s[i * EB:(i + 1) * EB] for i in idx])
    vals, st = e.decrypt_batch(sel, False)
    assert not st.any().item()
    assert vals.cpu().tolist() == [int(x[i].item() + x[half + i].item()) for i in idx]


def test_full_size_decrypt_l2():
    """BASELINE.json config 4 at FULL size: 2^14 level-2 ciphertexts e(E(a), E(b)) with a b uniform in
    (-T, T), T = 2^20, half of them negative and 1 % exact zeros; every plaintext is recovered, through
    the Lucas path and (same inputs) through the generic exponentiation + table search."""
    import os
    import torch
    from bgn_b200 import Engine
    g = load_golden(512)
    e = engine_for(g)
    count = 1 << 14
    dev = "cuda:0"
    gen = torch.Generator(device=dev)
    gen.manual_seed(4)
    av = torch.randint(1, 1 << 10, (count,), generator=gen, device=dev, dtype=torch.int64)
    bv = torch.randint(-(1 << 10) + 1, 1 << 10, (count,), generator=gen, device=dev, dtype=torch.int64)
    bv[::100] = 0
    r = torch.randint(0, 256, (2 * count, e.scalar_bytes), generator=gen, device=dev, dtype=torch.uint8)
    r[:, 0] &= 0x3F
    ca = e.encrypt_batch(av, r[:count].reshape(-1))
    cb = e.encrypt_batch(bv, r[count:].reshape(-1))
    l2 = e.pair_batch(ca, cb)
    vals, st = e.decrypt_batch(l2, True)
    assert not st.any().item() and torch.equal(vals, av * bv)
    assert int((av * bv < 0).sum().item()) > count // 3
    os.environ["BGN_DEC_LUCAS"] = "0"
    try:
        e2 = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    finally:
        del os.environ["BGN_DEC_LUCAS"]
    e2.set_secret(int(g["q1"], 16), g["msg_space"])
    v2, s2 = e2.decrypt_batch(l2, True)
    assert torch.equal(v2, vals) and not s2.any().item()
    e2.close()


def test_encrypt_windows():
    """The HBM-scale fixed-base windows of Q (enc_window = 16 .. 24 bits; 18, 20 and 22 are built from a temporary
    narrow table of 9-, 10- and 11-bit windows, 24 from the key's 8-bit table: 4 GB at keyBits=128) give the golden
    bytes, and on a random batch the same bytes as the 8-bit table built with the context, incl. scalars with
    zero top bytes (windows that reach past the scalar's top bit)."""
    from bgn_b200 import BgnError, Engine
    g = load_golden(128)
    ew = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
    with pytest.raises(BgnError):
        ew.set_option("enc_window", 12)
    with pytest.raises(BgnError):
        ew.set_option("enc_window", 19)
    with pytest.raises(BgnError):
        ew.set_option("no_such_knob", 1)
    rng = np.random.default_rng(24)
    cnt = 4096
    x = rng.integers(-1, 2, cnt)
    r = rng.integers(0, 256, (cnt, ew.scalar_bytes), dtype=np.uint8)
    r[:, 0] &= 0x3F
    r[::5, :4] = 0
    r[::7] = 0
    r[3::11] = 255  # r = 2^(8 bytes) - 1 >= n: every window at its last entry, the top window past the bits of n
    ew.set_option("enc_window", 8)  # no wide table: the 8-bit windows built with the context
    exp = ew.encrypt_batch(x, r.reshape(-1)).tobytes()
    # tables in twisted Edwards form (the default; curve.cuh: Ed) and in Weierstrass form
    for edw, widths in ((1, (0, 16, 18, 20, 22, 24)), (0, (0, 16, 20, 24))):  # 0: the automatic choice (20 bits here)
        ew.set_option("enc_edwards", edw)
        for bits in widths:
            ew.set_option("enc_window", bits)
            v = g["encrypt"]
            out = ew.encrypt_batch(np.array(v["x"], dtype=np.int64), scal(ew, v["r"], ew.scalar_bytes))
            assert out.tobytes() == unhex(v["out"]), (edw, bits)
            v = g["g1_blind"]
            assert ew.g1_blind_batch(buf(v["a"]), scal(ew, v["r"], ew.scalar_bytes)).tobytes() == unhex(v["out"]), (edw, bits)
            assert ew.encrypt_batch(x, r.reshape(-1)).tobytes() == exp, (edw, bits)
    ew.set_option("enc_edwards", 1)
    # plaintext-only sums (EncryptDeterministic: P's Edwards table alone), incl. x = 0 and large |x|
    xd = np.array([0, 1, -1, 2, 255, 256, -65537, (1 << 62) + 12345, -(1 << 62)], dtype=np.int64)
    got = ew.encrypt_batch(xd, None).tobytes()
    ew.set_option("enc_edwards", 0)
    assert ew.encrypt_batch(xd, None).tobytes() == got
    ew.set_option("enc_edwards", 1)
    ew.set_option("enc_table_max_mb", 1)  # the automatic choice under a table bound: back to 16 bits
    ew.set_option("enc_window", 0)
    assert ew.encrypt_batch(x, r.reshape(-1)).tobytes() == exp
    ew.close()


def test_two_contexts_two_threads_same_device():
    """Contexts of different keys share one device's __constant__ key material: calls are serialised
    per device and the constants follow the active context.  Two threads hammer two keys at once and
    every result must still equal the golden bytes."""
    import threading
    errs = []

    def work(kb):
        try:
            g = load_golden(kb)
            e, v, w = engine_for(g), g["pair"], g["encrypt"]
            for _ in range(12):
                assert e.pair_batch(buf(v["a"]), buf(v["b"])).tobytes() == unhex(v["out"])
                out = e.encrypt_batch(np.array(w["x"], dtype=np.int64), scal(e, w["r"], e.scalar_bytes))
                assert out.tobytes() == unhex(w["out"])
        except Exception as ex:  # noqa: BLE001
            errs.append((kb, repr(ex)))

    for kb in (64, 128):
        engine_for(load_golden(kb))  # create outside the threads (the engine cache is not thread-safe)
    ts = [threading.Thread(target=work, args=(kb,)) for kb in (64, 128)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs


def test_addpoly_batch_mixed_shapes():
    """AddPolyBatch with mixed levels, slot counts and scale factors equals AddPoly per polynomial
    (poly.go:171-226: MakePolyL2 promotion, alignPolyCiphertexts, tail pass-through)."""
    from bgn_b200 import PublicKey, SecretKey
    from bgn_b200.bgn import PolyCiphertextBatch
    g = load_golden(128)
    pk = PublicKey.FromPBCParams(g["pbc_params"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]), g["msg_space"])
    sk = SecretKey(int(g["q1"], 16))
    pk.SetupDecryption(sk)
    rng = random.Random(5)

    def make(values, pad_to):
        polys = []
        for v in values:
            pt = pk.NewPolyPlaintext(v)
            co = pt.Coefficients[: pt.Degree] + [0] * (pad_to - pt.Degree)
            pp = type(pt)(co, pad_to, pt.ScaleFactor, pt.params)
            polys.append(pk.EncryptPoly(pp, rs=[rng.randrange(pk.N) for _ in range(pad_to)]))
        sf = polys[0].ScaleFactor
        assert all(p.ScaleFactor == sf for p in polys)
        data = np.frombuffer(b"".join(p.CoeffBytes() for p in polys), dtype=np.uint8).copy()
        return polys, PolyCiphertextBatch(data, len(polys), pad_to, sf, False)

    ints, bi = make([5.0, 12.0, 40.0], 5)          # ScaleFactor 0, 5 slots
    thirds, bt = make([1 / 3, 2 / 3, 4 / 3], 3)    # ScaleFactor 1, 3 slots
    assert bt.ScaleFactor == 1
    cases = [(ints, bi, thirds, bt), (thirds, bt, ints, bi)]
    l2i = [pk.MakePolyL2(p) for p in ints]
    bl2 = pk.MakePolyL2Batch(bi)
    cases.append((l2i, bl2, thirds, bt))  # level 2 + level 1, different scale factors and slot counts
    for pa, ba, pb, bb in cases:
        got = pk.AddPolyBatch(ba, bb)
        exp = [pk.AddPoly(x, y) for x, y in zip(pa, pb)]
        assert (got.Degree, got.ScaleFactor, got.L2) == (exp[0].Degree, exp[0].ScaleFactor, exp[0].L2)
        assert bytes(got.data.tobytes()) == b"".join(e.CoeffBytes() for e in exp)
    vals = [sk.DecryptPoly(pk.AddPoly(x, y), pk).PolyEval() for x, y in zip(ints, thirds)]
    assert ["%.2f" % v for v in vals] == ["5.33", "12.67", "41.33"]


@pytest.mark.parametrize("steps", [7, 33, 1026])
def test_decrypt_with_giant_steps(steps):
    """A baby-step table smaller than the message space (the reference's own sizing is
    ceil(sqrt(T)) + 2 = 1026 entries at T = 2^20, gsbs.go:41-51): the search walks giant steps with
    the full element (k_gt_pow + k_bsgs_lookup) and must return what the one-probe Lucas path does."""
    from bgn_b200 import Engine
    for kb in (128, 512):
        g = load_golden(kb)
        e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
        e.set_secret(int(g["q1"], 16), g["msg_space"], baby_steps=steps)
        for lvl, key in ((True, "decrypt_l2"), (False, "decrypt_l1")):
            v = g[key]
            vals, st = e.decrypt_batch(buf(v["in"]), lvl)
            assert list(st) == v["status"] and [int(x) for x in vals] == v["out"]
        e.close()


@pytest.mark.parametrize("key_bits", [64, 128])
def test_keygen_on_gpu(key_bits):
    """NewKeyGen (bgn.go:65-138): fresh keys whose generators have the right orders (checked with the
    oracle) and which encrypt / multiply / decrypt correctly -- bgn_test.go's flow on a key made here."""
    from bgn_b200 import NewKeyGen
    from oracle import bgn_oracle as O
    rng = random.Random(key_bits)
    pk, sk = NewKeyGen(key_bits, 1021, 3, 3, 0.0001, True, rng=rng)
    par = O.a1_from_string(pk.PairingParams)
    P, Q = O.g1_from_bytes(pk.P, par), O.g1_from_bytes(pk.Q, par)
    assert pk.N.bit_length() == key_bits and O.g1_mul(pk.N, P, par.p) is None and O.g1_mul(sk.Key, P, par.p) is not None
    assert O.g1_mul(sk.Key, Q, par.p) is None
    pk.SetupDecryption(sk)
    a, b = pk.Encrypt(12), pk.Encrypt(-5)
    assert sk.Decrypt(pk.Add(a, b), pk) == 7 and sk.Decrypt(pk.Mult(a, b), pk) == -60
    c = pk.EncryptPoly(pk.NewPolyPlaintext(9.123))
    assert "%.1f" % sk.DecryptPoly(c, pk).PolyEval() == "9.1"
    pk.engine.close()


def test_gadgets_on_gpu():
    """gadgets_test.go:8-108 on a key generated here (keyBits=128), values and randomness below N."""
    from bgn_b200 import NewDecryptionProof, NewKeyGen, gadgets
    rng = random.Random(31)
    pk, sk = NewKeyGen(128, 1021, rng=rng)
    N = pk.N
    r, v, r2, v2 = (rng.randrange(N) for _ in range(4))
    ct, ct2 = pk.EncryptWithRandomness(v, r), pk.EncryptWithRandomness(v2, r2)
    assert pk.CheckDecryptionProof(ct, NewDecryptionProof(v, r))
    assert not pk.CheckDecryptionProof(ct, NewDecryptionProof(v, r2))
    assert not pk.CheckDecryptionProof(ct, NewDecryptionProof(r2, r))
    assert pk.CheckDecryptionProof(pk.Add(ct, ct2), NewDecryptionProof(v + v2, r + r2))
    good = pk.NewProofOfPlaintextKnowledge(sk, v, r)
    assert pk.CheckProofOfPlaintextKnoewledge(ct, good)
    assert not pk.CheckProofOfPlaintextKnoewledge(ct, pk.NewProofOfPlaintextKnowledge(sk, v, r2))
    assert not pk.CheckProofOfPlaintextKnoewledge(ct, pk.NewProofOfPlaintextKnowledge(sk, r2, r))
    many = [pk.NewProofOfPlaintextKnowledge(sk, v + i, r + i) for i in range(16)]
    cts = [p.Ct for p in many]
    assert gadgets.check_proofs_of_plaintext_knowledge(pk, cts, many) == [True] * 16
    cts[3], cts[4] = cts[4], cts[3]
    assert gadgets.check_proofs_of_plaintext_knowledge(pk, cts, many) == [True] * 3 + [False] * 2 + [True] * 11
    pk.engine.close()


def test_context_lifecycle_releases_device_memory():
    """create -> use every lazily built table (16-bit windows of Q, e(Q,Q) table, line table, baby steps)
    -> destroy, 12 times: the free device memory must come back (no leak in bgn_ctx_destroy)."""
    import torch
    from bgn_b200 import Engine
    g = load_golden(128)
    torch.cuda.synchronize()
    free0 = torch.cuda.mem_get_info()[0]
    v = g["encrypt"]
    for i in range(12):
        e = Engine(int(g["p"], 16), int(g["n"], 16), g["l"], bytes.fromhex(g["P"]), bytes.fromhex(g["Q"]))
        e.set_secret(int(g["q1"], 16), g["msg_space"])
        c = e.encrypt_batch(np.array(v["x"], dtype=np.int64), scal(e, v["r"], e.scalar_bytes))
        assert c.tobytes() == unhex(v["out"])
        l2 = e.make_l2_batch(c)
        e.gt_blind_batch(l2, scal(e, v["r"], e.scalar_bytes))
        e.decrypt_batch(c, False)
        e.close()
    torch.cuda.synchronize()
    free1 = torch.cuda.mem_get_info()[0]
    assert free0 - free1 < 64 << 20, "device memory not released: %d MiB" % ((free0 - free1) >> 20)
