"""The host mirror of the Go API (bgn_b200/bgn.py) on the CPU: its engine is replaced by a stand-in that
answers every C-ABI call from the oracle (tests/fake_engine.py), so what is tested here is the HOST logic --
level promotion, scale alignment, tail pass-through, slot counts, wire formats, error behaviour -- against
the oracle's own implementation of the same reference functions (poly.go, bgn.go)."""
import random

import numpy as np
import pytest

from conftest import load_golden
from fake_engine import FakeEngine
from oracle import bgn_oracle as O

from bgn_b200.bgn import DLError, PolyCiphertextBatch, PublicKey, SecretKey


@pytest.fixture(scope="module")
def keys():
    g = load_golden(64)
    p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    P, Q = bytes.fromhex(g["P"]), bytes.fromhex(g["Q"])
    pk = PublicKey(p, n, l, P, Q, g["msg_space"], engine=FakeEngine(p, n, l, P, Q))
    sk = SecretKey(int(g["q1"], 16))
    pk.SetupDecryption(sk)
    par = O.A1Params(p, n, l)
    opk = O.PublicKey(par, O.g1_from_bytes(P, par), O.g1_from_bytes(Q, par), g["msg_space"])
    osk = O.SecretKey(sk.Key, 0)
    O.setup_decryption(opk, osk)
    return pk, sk, opk, osk


def test_truth_table(keys):
    """cmd/main.go:74-107."""
    pk, sk, _, _ = keys
    cts = {m: pk.Encrypt(m) for m in (0, 1, -1)}
    for a in (0, 1, -1):
        assert sk.Decrypt(pk.Neg(cts[a]), pk) == -a
        for b in (0, 1, -1):
            assert sk.Decrypt(pk.Add(cts[a], cts[b]), pk) == a + b
            assert sk.Decrypt(pk.Sub(cts[a], cts[b]), pk) == a - b
            assert sk.Decrypt(pk.Mult(cts[a], cts[b]), pk) == a * b
            assert sk.Decrypt(pk.Add(pk.Mult(cts[a], cts[b]), cts[a]), pk) == a * b + a


def test_decrypt_errors(keys):
    pk, sk, _, _ = keys
    big = pk.EncryptDeterministic(10 ** 6)  # beyond bound^2 + bound + 2 for MsgSpace = 1021
    with pytest.raises(DLError):
        sk.Decrypt(big, pk)
    assert sk.DecryptFailSafe(big, pk) == 0
    g = load_golden(64)
    p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    P, Q = bytes.fromhex(g["P"]), bytes.fromhex(g["Q"])
    fresh = PublicKey(p, n, l, P, Q, g["msg_space"], engine=FakeEngine(p, n, l, P, Q))
    with pytest.raises(RuntimeError, match="DL tables not computed"):  # gsbs.go:56-58
        sk.Decrypt(fresh.Encrypt(1), fresh)


def test_poly_ops_match_oracle_bytes(keys):
    """EncryptPoly / AddPoly / SubPoly / MultPoly / MultConstPoly / MakePolyL2 / EvalPoly give the
    oracle's bytes and metadata for the same injected randomness."""
    pk, sk, opk, osk = keys
    rng = random.Random(1)

    def both(v):
        pt, opt = pk.NewPolyPlaintext(v), opk.new_poly_plaintext(v)
        assert pt.Coefficients[: pt.Degree] == opt.coefficients and pt.ScaleFactor == opt.scale_factor
        rs = [rng.randrange(pk.N) for _ in range(pt.Degree)]
        return pk.EncryptPoly(pt, rs=rs), O.encrypt_poly(opk, opt, rs)

    def same(ct, oct):
        assert (ct.Degree, ct.ScaleFactor, ct.L2) == (oct.degree, oct.scale_factor, oct.L2)
        assert ct.CoeffBytes() == O.poly_ct_bytes(opk, oct)

    a, oa = both(9.13)
    b, ob = both(4.0)
    c, oc = both(1 / 3)
    same(a, oa)
    same(pk.AddPoly(a, b), O.add_poly(opk, oa, ob))          # different scale factors: alignment
    same(pk.AddPoly(b, c), O.add_poly(opk, ob, oc))
    same(pk.SubPoly(b, c), O.sub_poly(opk, ob, oc))
    same(pk.MultPoly(b, c), O.mult_poly(opk, ob, oc))
    same(pk.MakePolyL2(b), O.make_poly_l2(opk, ob))
    same(pk.AddPoly(pk.MakePolyL2(b), c), O.add_poly(opk, O.make_poly_l2(opk, ob), oc))  # level promotion
    for const in (4.12, -2.0, 0.5):
        same(pk.MultConstPoly(b, const), O.mult_const_poly(opk, ob, const))
        same(pk.MultConstPoly(pk.MakePolyL2(b), const), O.mult_const_poly(opk, O.make_poly_l2(opk, ob), const))
    assert pk.EvalPoly(b).C == O.ct_bytes(opk, O.eval_poly(opk, ob))
    assert "%.1f" % sk.DecryptPoly(pk.MultPoly(b, c), pk).PolyEval() == "1.3"


def test_batch_entry_points_equal_per_polynomial_calls(keys):
    pk, sk, _, _ = keys
    rng = random.Random(2)

    def make(values, pad_to):
        polys = []
        for v in values:
            pt = pk.NewPolyPlaintext(v)
            pp = type(pt)(pt.Coefficients[: pt.Degree] + [0] * (pad_to - pt.Degree), pad_to, pt.ScaleFactor, pt.params)
            polys.append(pk.EncryptPoly(pp, rs=[rng.randrange(pk.N) for _ in range(pad_to)]))
        data = np.frombuffer(b"".join(x.CoeffBytes() for x in polys), dtype=np.uint8).copy()
        return polys, PolyCiphertextBatch(data, len(polys), pad_to, polys[0].ScaleFactor, False)

    ints, bi = make([5.0, 7.0], 4)
    thirds, bt = make([1 / 3, 2 / 3], 3)
    for pa, ba, pb, bb in ((ints, bi, thirds, bt), (thirds, bt, ints, bi),
                           ([pk.MakePolyL2(x) for x in ints], pk.MakePolyL2Batch(bi), thirds, bt)):
        got = pk.AddPolyBatch(ba, bb)
        exp = [pk.AddPoly(x, y) for x, y in zip(pa, pb)]
        assert (got.Degree, got.ScaleFactor, got.L2) == (exp[0].Degree, exp[0].ScaleFactor, exp[0].L2)
        assert bytes(np.asarray(got.data).tobytes()) == b"".join(e.CoeffBytes() for e in exp)
    prod = pk.MultPolyBatch(bi, bt)
    assert bytes(prod.data.tobytes()) == b"".join(pk.MultPoly(x, y).CoeffBytes() for x, y in zip(ints, thirds))
    total = pk.InnerProduct(bi, bt)
    assert "%.2f" % sk.DecryptPoly(total, pk).PolyEval() == "%.2f" % (5 / 3 + 14 / 3)
    mc = pk.MultConstPolyBatch(bi, -2.0)
    assert bytes(mc.data.tobytes()) == b"".join(pk.MultConstPoly(x, -2.0).CoeffBytes() for x in ints)
    assert pk.EvalPolyBatch(bi).tobytes() == b"".join(pk.EvalPoly(x).C for x in ints)
    vals, st = sk.DecryptPolyBatch(bi, pk)
    assert not st.any() and [int(sum(c * 3 ** i for i, c in enumerate(row))) for row in vals] == [5, 7]


def test_nondeterministic_mode_matches_oracle(keys):
    pk0, sk, opk, _ = keys
    pk = PublicKey(pk0.engine.p, pk0.N, pk0.engine.l, pk0.P, pk0.Q, pk0.MsgSpace, Deterministic=False,
                   engine=pk0.engine)
    pk._secret_set, pk._secret_key = True, sk.Key
    ond = O.PublicKey(opk.params, opk.P, opk.Q, opk.msg_space, deterministic=False)
    rng = random.Random(3)
    r = [rng.randrange(pk.N) for _ in range(8)]
    a, b = pk.EncryptWithRandomness(3, r[0]), pk.EncryptWithRandomness(2, r[1])
    oa, ob = O.encrypt_with_randomness(ond, 3, r[0]), O.encrypt_with_randomness(ond, 2, r[1])
    assert pk.Add(a, b, r=r[2]).C == O.ct_bytes(ond, O.add(ond, oa, ob, r[2]))
    assert pk.Sub(a, b, r=r[3]).C == O.ct_bytes(ond, O.sub(ond, oa, ob, r[3]))
    m, om = pk.Mult(a, b, r=r[4]), O.mult(ond, oa, ob, r[4])
    assert m.C == O.ct_bytes(ond, om)
    assert pk.MultConst(m, 5, r=r[5]).C == O.ct_bytes(ond, O.mult_const(ond, om, 5, r[5]))
    assert pk.Add(m, a, r=r[6]).C == O.ct_bytes(ond, O.add(ond, om, oa, r[6]))  # mixed levels
    s1, s2 = pk.Add(a, b), pk.Add(a, b)
    assert s1.C != s2.C and sk.Decrypt(s1, pk) == sk.Decrypt(s2, pk) == 5


def test_nondeterministic_poly_ops_match_oracle(keys):
    """every polynomial operation on a Deterministic=False key, fed by a replayed randomness stream,
    against the oracle's literal poly.go control flow on the same stream (tests/nondet_cases.py)"""
    from nondet_cases import run_poly_cases
    pk, sk, opk, osk = keys
    run_poly_cases(pk, opk, sk, osk)


@pytest.mark.parametrize("kb", [64, 128])
def test_nondeterministic_poly_ops_match_golden(kb):
    """the committed non-deterministic fixtures (tests/golden/kb*.json["nondet_poly"]) through the mirror on
    the oracle-backed stand-in engine"""
    from nondet_cases import run_golden_section
    g = load_golden(kb)
    p, n, l = int(g["p"], 16), int(g["n"], 16), g["l"]
    P, Q = bytes.fromhex(g["P"]), bytes.fromhex(g["Q"])
    pk = PublicKey(p, n, l, P, Q, g["msg_space"], engine=FakeEngine(p, n, l, P, Q))
    run_golden_section(pk, g)


def test_secret_key_mismatch_is_an_error(keys):
    pk, sk, _, _ = keys
    with pytest.raises(ValueError, match="not the one installed"):
        SecretKey(sk.Key + 2).Decrypt(pk.Encrypt(1), pk)


def test_big_and_negative_scalars(keys):
    """EncryptDeterministic takes any *big.Int (bgn.go:325-331); MultConst by a negative constant is
    |k| * (-C)."""
    pk, sk, opk, _ = keys
    x = (1 << 70) + 12345
    assert pk.EncryptDeterministic(x).C == O.ct_bytes(opk, O.encrypt_deterministic(opk, x % pk.N))
    c = pk.EncryptDeterministic(7)
    assert sk.Decrypt(pk.MultConst(c, -3), pk) == -21
    assert sk.Decrypt(pk.MultConst(pk.makeL2(c), -3), pk) == -21


def test_wire_roundtrip(keys):
    pk, sk, _, _ = keys
    c = pk.Encrypt(9)
    back = pk.NewCiphertextFromBytes(c.Bytes())
    assert (back.C, back.L2) == (c.C, False)
    pc = pk.MakePolyL2(pk.EncryptPoly(pk.NewPolyPlaintext(9.123)))
    pb = pk.NewPolyCiphertextFromBytes(pc.Bytes())
    assert pb.CoeffBytes() == pc.CoeffBytes() and (pb.Degree, pb.ScaleFactor, pb.L2) == (pc.Degree, pc.ScaleFactor, True)
    assert c.String().startswith("[") and pk.encryptZero().String() == "O\n"
    with pytest.raises(ValueError):
        pk.NewPolyCiphertextFromBytes(b"")


def test_keygen_host_logic():
    """NewKeyGen (bgn.go:65-138) with the group operations answered by the oracle-backed stand-in:
    type-A1 parameters, generator orders, key sizes, the reference's panics, and a round trip."""
    from bgn_b200 import NewKeyGen
    from bgn_b200.keygen import a1_params, is_probable_prime, rand_prime
    rng = random.Random(11)
    assert [m for m in range(2, 60) if is_probable_prime(m, rng)] == [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43,
                                                                     47, 53, 59]
    q = rand_prime(rng, 24)
    assert q.bit_length() == 24 and q >> 22 == 3 and is_probable_prime(q, rng)
    pk, sk = NewKeyGen(48, 1021, 3, 3, 0.0001, True, rng=rng, engine_factory=FakeEngine)
    par = O.a1_from_string(pk.PairingParams)
    assert pk.N.bit_length() == 48 and pk.N % sk.Key == 0 and par.n == pk.N
    assert (par.p, par.l) == a1_params(pk.N, rng) == (O.a1_gen(pk.N).p, O.a1_gen(pk.N).l)
    P, Q = O.g1_from_bytes(pk.P, par), O.g1_from_bytes(pk.Q, par)
    q1, q2 = sk.Key, pk.N // sk.Key
    assert O.g1_mul(pk.N, P, par.p) is None and O.g1_mul(q1, P, par.p) is not None and O.g1_mul(q2, P, par.p) is not None
    assert O.g1_mul(q1, Q, par.p) is None and Q is not None  # Q generates the subgroup of order q1
    pk.SetupDecryption(sk)
    assert sk.Decrypt(pk.Mult(pk.Encrypt(-7), pk.Encrypt(6)), pk) == -42
    with pytest.raises(ValueError, match="divisible by 2"):
        NewKeyGen(33, 10, rng=rng, engine_factory=FakeEngine)
    with pytest.raises(ValueError, match=">= 16"):
        NewKeyGen(8, 10, rng=rng, engine_factory=FakeEngine)
    with pytest.raises(ValueError, match="Message space"):
        NewKeyGen(32, 1 << 20, rng=rng, engine_factory=FakeEngine)


def test_gadgets(keys):
    """gadgets_test.go:8-108: decryption proofs (valid, aggregate, bad) and proofs of plaintext
    knowledge (valid, bad) with values and randomness below N."""
    from bgn_b200 import NewDecryptionProof
    from bgn_b200 import gadgets
    pk, _, _, _ = keys
    g = load_golden(64)
    sk = SecretKey(int(g["q1"], 16), R=0)
    rng = random.Random(8)
    N = pk.N
    r, v, r2, v2 = (rng.randrange(N) for _ in range(4))
    ct = pk.EncryptWithRandomness(v, r)
    assert pk.CheckDecryptionProof(ct, NewDecryptionProof(v, r))
    assert not pk.CheckDecryptionProof(ct, NewDecryptionProof(v, r2))
    assert not pk.CheckDecryptionProof(ct, NewDecryptionProof(r2, r))
    ct2 = pk.EncryptWithRandomness(v2, r2)
    assert pk.CheckDecryptionProof(pk.Add(ct, ct2), NewDecryptionProof(v + v2, r + r2))  # aggregate
    assert gadgets.check_decryption_proofs(pk, [ct, ct2, ct], [NewDecryptionProof(v, r), NewDecryptionProof(v2, r2),
                                                                 NewDecryptionProof(v2, r)]) == [True, True, False]


def test_proof_of_plaintext_knowledge_on_generated_key():
    """needs SecretKey.R (Q = P^(R q2)), so the key comes from NewKeyGen"""
    from bgn_b200 import NewKeyGen
    rng = random.Random(21)
    pk, sk = NewKeyGen(48, 1021, rng=rng, engine_factory=FakeEngine)
    N = pk.N
    r, v, r2 = (rng.randrange(N) for _ in range(3))
    ct = pk.EncryptWithRandomness(v, r)
    assert pk.CheckProofOfPlaintextKnoewledge(ct, pk.NewProofOfPlaintextKnowledge(sk, v, r))
    assert not pk.CheckProofOfPlaintextKnoewledge(ct, pk.NewProofOfPlaintextKnowledge(sk, v, r2))
    assert not pk.CheckProofOfPlaintextKnoewledge(ct, pk.NewProofOfPlaintextKnowledge(sk, r2, r))
