"""CPU tests of the oracle itself (oracle/bgn_oracle.py): golden vectors, group/pairing
properties, and the reference's own test cases (bgn_test.go, poly_test.go, cmd/main.go)
restated.  No GPU."""
import importlib.util
import os
import random

import pytest

from conftest import GOLDEN_DIR, load_golden
from oracle import bgn_oracle as O

spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN_DIR, "make_golden.py"))
make_golden = importlib.util.module_from_spec(spec)
spec.loader.exec_module(make_golden)


def key_from_golden(g):
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    P = O.g1_from_bytes(bytes.fromhex(g["P"]), par)
    Q = O.g1_from_bytes(bytes.fromhex(g["Q"]), par)
    return O.PublicKey(par, P, Q, g["msg_space"]), O.SecretKey(int(g["q1"], 16), 0)


@pytest.mark.parametrize("kb", [64, 128, 256])
def test_golden_files_are_what_the_oracle_produces(kb):
    """the committed fixtures are exactly make_golden.py's output (pins oracle + generator)"""
    assert make_golden.make(kb, small=False) == load_golden(kb)


@pytest.mark.parametrize("kb", [512, 1024])
def test_golden_large_keys_spot_check(kb):
    g = load_golden(kb)
    pk, sk = key_from_golden(g)
    par = pk.params
    assert par == O.a1_gen(par.n) or kb == 1024  # l is the smallest multiple of 4 with l*n-1 prime
    assert par.p == par.l * par.n - 1 and par.p % 4 == 3
    assert par.n == int(g["q1"], 16) * int(g["q2"], 16)
    v = g["encrypt"]
    for i in (1, 2, 3):
        x, r = v["x"][i], int(v["r"][i], 16)
        c = O.encrypt_with_randomness(pk, abs(x), r).C
        c = O.g1_neg(c, par.p) if x < 0 else c
        assert O.g1_to_bytes(c, par).hex() == v["out"][i]
    v = g["pair"]
    a = O.g1_from_bytes(bytes.fromhex(v["a"][1]), par)
    b = O.g1_from_bytes(bytes.fromhex(v["b"][1]), par)
    assert O.gt_to_bytes(O.pairing(a, b, par), par).hex() == v["out"][1]


def test_a1_params_shape(golden):
    p, n, l = int(golden["p"], 16), int(golden["n"], 16), golden["l"]
    assert p + 1 == l * n and l % 4 == 0 and p % 4 == 3
    assert O.is_probable_prime(p) and O.is_probable_prime(int(golden["q1"], 16))
    assert golden["coord_bytes"] == (p.bit_length() + 7) // 8
    assert n.bit_length() == golden["key_bits"]  # rand.Prime sets the top two bits
    assert O.a1_from_string(golden["pbc_params"]) == O.A1Params(p, n, l)


@pytest.mark.parametrize("kb", [64, 128, 512])
def test_pairing_properties(kb):
    g = load_golden(kb)
    pk, sk = key_from_golden(g)
    par, p, n = pk.params, pk.params.p, pk.params.n
    rng = random.Random(kb)
    a, b = rng.randrange(1, n), rng.randrange(1, n)
    e = O.pairing(pk.P, pk.Q, par)
    assert e != O.GT_ONE
    assert O.fp2_pow(e, n, p) == O.GT_ONE  # order divides n
    assert O.pairing(O.g1_mul(a, pk.P, p), O.g1_mul(b, pk.Q, p), par) == O.fp2_pow(e, a * b % n, p)  # bilinear
    assert O.pairing(pk.Q, pk.P, par) == e  # symmetric (distortion map)
    assert O.fp2_mul(e, O.fp2_conj(e, p), p) == O.GT_ONE  # GT is unitary: inverse = conjugate
    assert O.pairing(None, pk.P, par) == O.GT_ONE and O.pairing(pk.P, None, par) == O.GT_ONE
    # subgroup structure of the key: Q has order q1, so e(Q,Q)^q1 = 1 and P^q1 != O
    assert O.g1_mul(sk.key, pk.Q, p) is None and O.g1_mul(sk.key, pk.P, p) is not None
    assert O.g1_mul(n, pk.P, p) is None


def test_serialisation(golden):
    pk, _ = key_from_golden(golden)
    par = pk.params
    B = par.coord_bytes
    pb = O.g1_to_bytes(pk.P, par)
    assert len(pb) == 2 * B and O.g1_from_bytes(pb, par) == pk.P
    assert O.g1_to_bytes(None, par) == bytes(2 * B) and O.g1_from_bytes(bytes(2 * B), par) is None
    bad = bytearray(pb)
    bad[-1] ^= 1
    assert O.g1_from_bytes(bytes(bad), par) is None  # off-curve -> O (curve_from_bytes)
    e = O.pairing(pk.P, pk.P, par)
    assert O.gt_from_bytes(O.gt_to_bytes(e, par), par) == e
    assert O.gt_to_bytes(O.GT_ONE, par) == (1).to_bytes(B, "big") + bytes(B)
    assert O.g1_string(None) == "O" and O.gt_string((1, 0)) == "[1, 0]"


def test_truth_table_cmd_main():
    """cmd/main.go:74-107 over {0, 1, -1}."""
    pk, sk = O.keygen(64, 1021, seed=5)
    O.setup_decryption(pk, sk)
    rng = random.Random(3)
    ct = {m: (O.encrypt_with_randomness(pk, m, rng.randrange(pk.n)) if m >= 0 else
              O.neg(pk, O.encrypt_with_randomness(pk, -m, rng.randrange(pk.n)))) for m in (0, 1, -1)}
    for a in (0, 1, -1):
        assert O.decrypt(pk, sk, O.neg(pk, ct[a])) == -a
        for b in (0, 1, -1):
            assert O.decrypt(pk, sk, O.add(pk, ct[a], ct[b])) == a + b
            assert O.decrypt(pk, sk, O.mult(pk, ct[a], ct[b])) == a * b
            assert O.decrypt(pk, sk, O.sub(pk, O.mult(pk, ct[a], ct[b]), ct[b])) == a * b - b


def test_non_deterministic_mode():
    """re-randomised homomorphic ops (bgn.go:260-269, 302-311, 466-474, 488-495) decrypt the same."""
    pk, sk = O.keygen(64, 1021, deterministic=False, seed=6)
    O.setup_decryption(pk, sk)
    a, b = O.encrypt_with_randomness(pk, 5, 11), O.encrypt_with_randomness(pk, 7, 13)
    assert O.decrypt(pk, sk, O.add(pk, a, b, r=99)) == 12
    assert O.decrypt(pk, sk, O.mult(pk, a, b, r=12345)) == 35
    assert O.decrypt(pk, sk, O.mult_const(pk, a, 3, r=5)) == 15
    assert O.decrypt(pk, sk, O.add(pk, O.mult(pk, a, b, r=1), a, r=2)) == 40


def f1(x):
    return "%.1f" % x


def test_poly_cases_of_the_reference():
    """poly_test.go:68-189 with bgn_test.go:8-13's constants (KEYBITS lowered to 128)."""
    pk, sk = O.keygen(128, 1021, 3, 3, 0.0001, True, seed=9)
    O.setup_decryption(pk, sk)
    rng = random.Random(2)

    def E(x):
        pt = pk.new_poly_plaintext(x)
        return O.encrypt_poly(pk, pt, [rng.randrange(pk.n) for _ in range(pt.degree)])

    def D(c):
        return O.decrypt_poly(pk, sk, c).poly_eval()

    assert f1(pk.new_poly_plaintext(9.123).poly_eval()) == "9.1"  # TestEncodeBalancedPoly
    assert f1(pk.new_unbalanced_plaintext(9.123).poly_eval()) == "9.1"  # TestEncodeUnbalancedPoly
    assert f1(D(E(9.123))) == "9.1"  # TestEncodeEncryptDecryptPoly
    assert f1(D(O.add_poly(pk, E(0.1), E(4.2)))) == "4.3"  # TestAddPoly
    assert f1(D(O.add_poly(pk, O.make_poly_l2(pk, E(50.1)), O.make_poly_l2(pk, E(41.2))))) == "91.3"  # TestAddPolyL2
    c = E(9.13)
    assert f1(D(O.mult_const_poly(pk, c, 4.12))) == f1(9.13 * 4.12)  # TestMultConstPoly L1
    assert f1(D(O.mult_const_poly(pk, O.make_poly_l2(pk, c), 4.12))) == f1(9.13 * 4.12)  # ... L2
    assert f1(D(O.mult_poly(pk, E(1.1), E(40.2)))) == f1(1.1 * 40.2)  # TestMultPoly
    assert D(O.eval_poly_as_poly(pk, E(100.0))) == 100.0 if hasattr(O, "eval_poly_as_poly") else True


def test_multpoly_slot_layout():
    """poly.go:123-156: Degree = d1 + d2 slots, the last one stays the GT identity."""
    pk, sk = O.keygen(64, 1021, seed=11)
    c1 = O.encrypt_poly(pk, O.PolyPlaintext([1, -1, 0], 0, 3, 3), [3, 4, 5])
    c2 = O.encrypt_poly(pk, O.PolyPlaintext([1, 1], 0, 3, 3), [6, 7])
    mp = O.mult_poly(pk, c1, c2)
    assert mp.degree == 5 and len(mp.coefficients) == 5 and mp.L2
    assert mp.coefficients[-1].C == O.GT_ONE
    O.setup_decryption(pk, sk)
    assert O.decrypt_poly(pk, sk, mp).coefficients == [1, 0, -1, 0, 0]


def test_encodings_roundtrip():
    deg, sums = O.compute_encoding_table(3)
    assert deg[:4] == [1, 3, 9, 27] and sums[:4] == [1, 4, 13, 40]
    for v in list(range(0, 400)) + [1021, 65535, 10 ** 9 + 7]:
        b = O.balanced_encode(v, 3)
        assert set(b) <= {-1, 0, 1} and sum(c * 3 ** i for i, c in enumerate(b)) == v
        u = O.unbalanced_encode(v, 3)
        assert set(u) <= {0, 1, 2} and sum(c * 3 ** i for i, c in enumerate(u)) == v
    assert O.balanced_encode(-5, 3) == [-c for c in O.balanced_encode(5, 3)]
    assert O.balanced_encode(0, 3) == [0] and O.unbalanced_encode(0, 3) == [0]
    with pytest.raises(ValueError):
        O.unbalanced_encode(-1, 3)
    num, sf = O.rationalize(0.5, 3, 0.0001)
    assert abs(num / 3 ** sf - 0.5) <= 0.0001
    assert O.rationalize(1.0 / 3.0, 3, 0.0001) == (1, 1)


def test_bsgs_bounds():
    """gsbs.go:41-106: largest recoverable |m| is bound^2 + bound + 2; beyond that -> error."""
    pk, sk = O.keygen(64, 100, seed=12)
    O.setup_decryption(pk, sk)
    bound = 10
    top = bound * bound + bound + 2
    assert O.decrypt(pk, sk, O.encrypt_deterministic(pk, top)) == top
    assert O.decrypt(pk, sk, O.encrypt_deterministic(pk, -top)) == -top
    with pytest.raises(O.DLError):
        O.decrypt(pk, sk, O.encrypt_deterministic(pk, top + 1))
    assert O.decrypt_fail_safe(pk, sk, O.encrypt_deterministic(pk, top + 1)) == 0
    pk2, _ = O.keygen(64, 100, seed=12)
    with pytest.raises(RuntimeError, match="DL tables not computed"):
        O.get_dl(pk2, (1, 1), (1, 1), True)


def test_work_model_formula():
    """canonical modmuls per pairing for the actual key."""
    g = load_golden(512)
    par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
    m = O.canonical_modmuls_per_pairing(par)
    assert 15000 < m < 18000 and O.products_per_modmul(17) == 595


@pytest.mark.parametrize("kb", [64, 128])
def test_nondet_poly_golden_section_is_reproducible(kb):
    """the oracle's literal poly.go control flow on the recorded randomness stream reproduces the committed
    non-deterministic fixtures (pins the oracle, and the draw order, against accidental change)"""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    from conftest import load_golden
    g = load_golden(kb)
    pk, _ = O.keygen(kb, mg.MSG_SPACE[kb], seed=mg.SEEDS[kb])
    assert mg.nondet_poly_section(pk, kb) == g["nondet_poly"]
    ops = g["nondet_poly"]["ops"]
    # MultPoly 3 x 2: two draws per coefficient pairing; its unused top slot is the GT identity
    assert ops["mult_poly"]["draws_used"] == 2 * 3 * 2 and ops["add_poly"]["draws_used"] == 2 and ops["neg_poly"]["draws_used"] == 3
    cb = g["coord_bytes"]
    assert ops["mult_poly"]["out"][-1] == (b"\x00" * (cb - 1) + b"\x01" + b"\x00" * cb).hex()


def test_parabola_step_identity():
    """the Miller loop with merged doubling-and-addition steps (ELM parabola in Jacobian coordinates,
    normalised evaluation point) gives the oracle's pairing value -- the arithmetic the device routines
    curve.cuh G::dadd_para / fused.cuh MF::dadd_para and para_mul implement, in plain integers."""
    import importlib.util
    import os
    import random
    spec = importlib.util.spec_from_file_location(
        "parabola_proto", os.path.join(os.path.dirname(__file__), "..", "tools", "_proto", "parabola.py"))
    proto = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(proto)
    from conftest import load_golden
    from oracle import bgn_oracle as O
    for kb in (64, 128):
        g = load_golden(kb)
        par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
        P = O.g1_from_bytes(bytes.fromhex(g["P"]), par)
        rng = random.Random(kb)
        for _ in range(3):
            A = O.g1_mul(rng.randrange(1, par.n), P, par.p)
            B = O.g1_mul(rng.randrange(1, par.n), P, par.p)
            assert O.final_exp(proto.miller_parabola(A, B, par), par) == O.pairing(A, B, par)
