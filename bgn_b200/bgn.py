"""Host-side mirror of the reference's Go API (bgn.go, poly.go, ciphertext.go).

Same names, argument meaning and error behaviour as the reference so that the
parity tests read like bgn_test.go / poly_test.go; every group operation is one
call into the CUDA library through the C-ABI (bgn_b200.engine.Engine).  The
single-element methods (Encrypt, Add, Mult, ...) are batches of one; the
`*Batch` methods are the new entry points north_star asks for and are what a
production caller uses.

What differs from the reference, on purpose:
  * elements are held as PBC `element_to_bytes` byte strings, not *pbc.Element;
  * randomness is injected (`r=`) or drawn from `secrets`; the reference draws
    from crypto/rand inside the call (bgn.go:567-574);
  * O (point at infinity) is all-zero bytes (SURVEY.md 8(a) note);
  * decryption tables are per key, not process globals (gsbs.go:12-15).
"""
from __future__ import annotations

import math
import secrets
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import gadgets, gobwire
from .engine import Engine
from .plaintext import (EncodingTable, NewPolyPlaintext, NewUnbalancedPlaintext, PolyEncodingParams, PolyPlaintext)


def _element_string(raw: bytes, L2: bool) -> str:
    """pbc Element.String() in base 10: "[x, y]" for G1 (or "O"), "[re, im]" for GT
    (ciphertext.go:60-72 prints these)."""
    h = len(raw) // 2
    a, b = int.from_bytes(raw[:h], "big"), int.from_bytes(raw[h:], "big")
    if not L2 and a == 0 and b == 0:
        return "O"
    return "[%d, %d]" % (a, b)


class DLError(Exception):
    """errors.New("cannot find discrete log; out of bounds") (gsbs.go:105)."""


@dataclass
class Ciphertext:  # ciphertext.go:12-15
    C: bytes  # Element.Bytes(): x||y (G1) or re||im (GT)
    L2: bool

    def Copy(self) -> "Ciphertext":
        return Ciphertext(self.C, self.L2)

    def Bytes(self) -> bytes:
        """ciphertext.go:76-91: the gob envelope of ciphertextWrapper{CBytes, L2}."""
        return gobwire.encode_ciphertext(self.C, self.L2)

    def String(self) -> str:
        return _element_string(self.C, self.L2) + "\n"  # ciphertext.go:60-62


@dataclass
class PolyCiphertext:  # ciphertext.go:26-31
    Coefficients: List[Ciphertext]
    Degree: int  # number of coefficient slots (poly.go:13)
    ScaleFactor: int
    L2: bool

    def Copy(self) -> "PolyCiphertext":
        return PolyCiphertext(self.Coefficients, self.Degree, self.ScaleFactor, self.L2)

    def String(self) -> str:
        return "".join(_element_string(c.C, c.L2) + "\n" for c in self.Coefficients)  # ciphertext.go:64-72

    def CoeffBytes(self) -> bytes:
        """the coefficients' element bytes back to back: the layout of a PolyCiphertextBatch row"""
        return b"".join(c.C for c in self.Coefficients)

    def Bytes(self) -> bytes:
        """ciphertext.go:93-116: the gob envelope of polyCiphertextWrapper."""
        return gobwire.encode_poly_ciphertext([c.C for c in self.Coefficients], self.Degree, self.ScaleFactor, self.L2)


@dataclass
class PolyCiphertextBatch:
    """`count` polynomial ciphertexts of equal slot count in one contiguous buffer
    (count x Degree x elem_bytes); `data` is a numpy array or a CUDA torch tensor."""
    data: object
    count: int
    Degree: int
    ScaleFactor: int
    L2: bool


@dataclass
class SecretKey:  # bgn.go:58-62
    Key: int  # q1
    R: int = 0
    PolyBase: int = 3

    def Decrypt(self, ct: Ciphertext, pk: "PublicKey") -> int:
        """bgn.go:205-207; raises DLError where the reference returns an error."""
        vals, st = pk._need_secret(self).decrypt_batch(np.frombuffer(ct.C, dtype=np.uint8), ct.L2)
        if st[0]:
            raise DLError("cannot find discrete log; out of bounds")
        return int(vals[0])

    def DecryptFailSafe(self, ct: Ciphertext, pk: "PublicKey") -> int:
        """bgn.go:210-216: 0 on failure."""
        try:
            return self.Decrypt(ct, pk)
        except DLError:
            return 0

    def DecryptPoly(self, ct: PolyCiphertext, pk: "PublicKey") -> PolyPlaintext:
        """poly.go:32-42 (the reference drops decrypt errors; failed slots are 0 here)."""
        buf = np.frombuffer(b"".join(c.C for c in ct.Coefficients[: ct.Degree]), dtype=np.uint8)
        vals, _ = pk._need_secret(self).decrypt_batch(buf, ct.L2)
        return PolyPlaintext([int(v) for v in vals], ct.Degree, ct.ScaleFactor, pk.PolyEncodingParams)

    def DecryptPolyBatch(self, cts: PolyCiphertextBatch, pk: "PublicKey"):
        """-> (int64 [count, Degree], status uint8 [count, Degree])."""
        vals, st = pk._need_secret(self).decrypt_batch(cts.data, cts.L2)
        return vals.reshape(cts.count, cts.Degree), st.reshape(cts.count, cts.Degree)


class PublicKey:
    """bgn.go:28-41.  Construct from imported key material (KeyGen itself is a
    one-off host computation and out of the accelerated path)."""

    def __init__(self, p: int, n: int, l: int, P: bytes, Q: bytes, MsgSpace: int, Deterministic: bool = True,
                 polyBase: int = 3, fpScaleBase: int = 3, fpPrecision: float = 0.0001, device: int = 0,
                 engine=None):
        """`engine`: an already constructed engine for this key (anything with Engine's methods; the CPU
        test-suite passes a stand-in so the host logic of this file can be exercised without a GPU)."""
        self.N = n
        self.P, self.Q = bytes(P), bytes(Q)
        self.MsgSpace = MsgSpace
        self.PairingParams = "type a1\np %d\nn %d\nl %d\n" % (p, n, l)
        self.Deterministic = Deterministic
        self.PolyEncodingParams = PolyEncodingParams(polyBase, fpScaleBase, fpPrecision)
        self._table = EncodingTable(polyBase)  # computeEncodingTable, bgn.go:135
        self.engine = engine if engine is not None else Engine(p, n, l, P, Q, device)
        self._secret_set = False

    @classmethod
    def FromPBCParams(cls, params: str, P: bytes, Q: bytes, MsgSpace: int, **kw) -> "PublicKey":
        """pbc.NewPairingFromString + SetBytes (bgn.go:640-653)."""
        vals = dict(line.split(None, 1) for line in params.strip().splitlines())
        if vals.get("type") != "a1":
            raise ValueError("only PBC type a1 parameters are supported")
        return cls(int(vals["p"]), int(vals["n"]), int(vals["l"]), P, Q, MsgSpace, **kw)

    def MarshalBinary(self) -> bytes:
        """bgn.go:597-624: the gob envelope of publicKeyWrapper.  G1 (an element used only as a factory,
        bgn.go:29) is written as the generator P."""
        pe = self.PolyEncodingParams
        return gobwire.encode_public_key(self.P, self.P, self.Q, self.N, self.MsgSpace, self.PairingParams,
                                         self.Deterministic, pe.PolyBase, pe.FPScaleBase, pe.FPPrecision)

    @classmethod
    def UnmarshalBinary(cls, data: bytes, device: int = 0) -> "PublicKey":
        """bgn.go:628-666: a public key marshalled by the reference (or by MarshalBinary) onto a GPU."""
        w = gobwire.decode_public_key(data)
        return cls.FromPBCParams(w["PairingParams"], w["P"], w["Q"], w["MsgSpace"], Deterministic=w["Deterministic"],
                                 polyBase=w["PolyBase"], fpScaleBase=w["FPScaleBase"], fpPrecision=w["FPPrecision"],
                                 device=device)

    # ---------------------------------------------------------------- setup
    def SetupDecryption(self, sk: SecretKey):
        """bgn.go:195-201 + PrecomputeTables (gsbs.go:41-51): tables live on the device."""
        self.engine.set_secret(sk.Key, self.MsgSpace)
        self._secret_set = True

    def _need_secret(self, sk: SecretKey) -> Engine:
        if not self._secret_set:
            raise RuntimeError("DL tables not computed!")  # gsbs.go:56-58
        return self.engine

    # ---------------------------------------------------------------- helpers
    @property
    def elem_bytes(self) -> int:
        return self.engine.elem_bytes

    def _np(self, *cts: Ciphertext) -> np.ndarray:
        return np.frombuffer(b"".join(c.C for c in cts), dtype=np.uint8)

    def _rand(self, r: Optional[int]) -> int:
        return secrets.randbelow(self.N) if r is None else r

    def _blind_g1(self, elem: np.ndarray, r: Optional[int]) -> np.ndarray:
        """+ r*Q (bgn.go:264-268, 491-495): bgn_g1_blind_batch, fixed-base windows of Q."""
        return self.engine.g1_blind_batch(elem, self.engine.scalars_be([self._rand(r) % self.N]))

    def _blind_gt(self, elem: np.ndarray, r: Optional[int]) -> np.ndarray:
        """* e(Q,Q)^r (bgn.go:283-287, 306-310, 469-474): bgn_gt_blind_batch; e(Q,Q) is a fixed-base
        table built once per key (the reference recomputes the pairing per call)."""
        return self.engine.gt_blind_batch(elem, self.engine.scalars_be([self._rand(r) % self.N]))

    def _rand_be(self, count: int, rs=None):
        """count scalars below N as the C-ABI's big-endian buffer (newCryptoRandom, bgn.go:567-574)"""
        if rs is None:
            rs = [secrets.randbelow(self.N) for _ in range(count)]
        if hasattr(rs, "dtype") or type(rs).__module__.startswith("torch"):
            return rs  # already a packed buffer
        assert len(rs) == count
        return self.engine.scalars_be([int(r) % self.N for r in rs])

    # ---------------------------------------------------------------- wire formats
    def _canonical(self, raw: bytes, L2: bool) -> bytes:
        """Element.SetBytes semantics (coordinates reduced mod p; a G1 pair off the curve becomes O)
        obtained from the device's own parser: x -> O + x  /  1 * x."""
        eb = self.elem_bytes
        if len(raw) % eb:
            raise ValueError("element bytes have the wrong length for this key")
        buf = np.frombuffer(raw, dtype=np.uint8)
        if L2:
            one = (b"\x00" * (eb // 2 - 1) + b"\x01" + b"\x00" * (eb // 2)) * (len(raw) // eb)
            return self.engine.gt_mul_batch(np.frombuffer(one, dtype=np.uint8), buf).tobytes()
        return self.engine.g1_add_batch(np.zeros(len(raw), dtype=np.uint8), buf).tobytes()

    def NewCiphertextFromBytes(self, data: bytes) -> Ciphertext:
        """bgn.go:505-528 -- without the pairing e(Q,Q) the reference evaluates per call (bgn.go:517)."""
        if len(data) == 0:
            raise ValueError("no data provided")
        C, L2 = gobwire.decode_ciphertext(data)
        return Ciphertext(self._canonical(C, L2), L2)

    def NewPolyCiphertextFromBytes(self, data: bytes) -> PolyCiphertext:
        """bgn.go:530-555; one device call for all coefficients."""
        if len(data) == 0:
            raise ValueError("no data provided")
        coeffs, degree, sf, L2 = gobwire.decode_poly_ciphertext(data)
        raw = self._canonical(b"".join(coeffs), L2) if coeffs else b""
        eb = self.elem_bytes
        return PolyCiphertext([Ciphertext(raw[i * eb:(i + 1) * eb], L2) for i in range(len(coeffs))], degree, sf, L2)

    # ---------------------------------------------------------------- plaintexts
    def NewPolyPlaintext(self, m: float) -> PolyPlaintext:
        return NewPolyPlaintext(m, self.PolyEncodingParams, self._table)

    def NewUnbalancedPlaintext(self, m: float) -> PolyPlaintext:
        return NewUnbalancedPlaintext(m, self.PolyEncodingParams, self._table)

    # ---------------------------------------------------------------- scalar scheme
    def _encrypt_big(self, xs: Sequence[int], rs: Sequence[int]) -> bytes:
        """P^x * Q^r for arbitrary-size x (the reference takes a *big.Int, gadgets_test.go encrypts
        values below N): x*P as a variable-scalar multiplication of P, then + r*Q through the
        fixed-base windows of Q.  x and r are reduced mod N, the order of P (Q's order divides it)."""
        eng, nb = self.engine, self.engine.scalar_bytes
        P = np.frombuffer(self.P * len(xs), dtype=np.uint8)
        xp = eng.g1_mulconst_batch(P, eng.scalars_be([x % self.N for x in xs], nb), nb)
        out = eng.g1_blind_batch(xp, eng.scalars_be([r % self.N for r in rs], nb))
        return bytes(np.asarray(out).tobytes())

    def EncryptWithRandomness(self, x: int, r: int) -> Ciphertext:
        """bgn.go:340-353: C = P^x * Q^r."""
        if abs(x) >= 1 << 63:
            return Ciphertext(self._encrypt_big([x], [r]), False)
        out = self.engine.encrypt_batch(np.array([x], dtype=np.int64), self.engine.scalars_be([r % self.N]))
        return Ciphertext(out.tobytes(), False)

    # ---------------------------------------------------------------- zero-knowledge gadgets (gadgets.go)
    def NewProofOfPlaintextKnowledge(self, sk: SecretKey, v: int, z: int, nonce1: Optional[int] = None):
        return gadgets.new_proof_of_plaintext_knowledge(self, sk, v, z, nonce1)

    def CheckDecryptionProof(self, ct: Ciphertext, proof) -> bool:
        return gadgets.check_decryption_proofs(self, [ct], [proof])[0]

    def CheckProofOfPlaintextKnoewledge(self, ct: Ciphertext, proof) -> bool:  # the reference's spelling
        return gadgets.check_proofs_of_plaintext_knowledge(self, [ct], [proof])[0]

    CheckProofOfPlaintextKnowledge = CheckProofOfPlaintextKnoewledge

    def Encrypt(self, x: int, r: Optional[int] = None) -> Ciphertext:
        """bgn.go:334-337."""
        return self.EncryptWithRandomness(x, self._rand(r))

    def EncryptDeterministic(self, x: int) -> Ciphertext:
        """bgn.go:325-331."""
        out = self.engine.encrypt_batch(np.array([x], dtype=np.int64), None)
        return Ciphertext(out.tobytes(), False)

    def encryptZero(self) -> Ciphertext:
        return self.EncryptDeterministic(0)  # bgn.go:562-564

    def makeL2(self, ct: Ciphertext) -> Ciphertext:
        """bgn.go:316-321: e(C, P)."""
        return Ciphertext(self.engine.make_l2_batch(self._np(ct)).tobytes(), True)

    def _promote(self, a: Ciphertext, b: Ciphertext) -> Tuple[Ciphertext, Ciphertext]:
        if a.L2 and not b.L2:
            b = self.makeL2(b)
        if not a.L2 and b.L2:
            a = self.makeL2(a)
        return a, b

    def Add(self, ct1: Ciphertext, ct2: Ciphertext, r: Optional[int] = None) -> Ciphertext:
        """bgn.go:442-497."""
        a, b = self._promote(ct1, ct2)
        if a.L2:
            res = self.engine.gt_mul_batch(self._np(a), self._np(b))
            if not self.Deterministic:
                res = self._blind_gt(res, r)
            return Ciphertext(res.tobytes(), True)
        res = self.engine.g1_add_batch(self._np(a), self._np(b))
        if not self.Deterministic:
            res = self._blind_g1(res, r)
        return Ciphertext(res.tobytes(), False)

    def Sub(self, ct1: Ciphertext, ct2: Ciphertext, r: Optional[int] = None) -> Ciphertext:
        """bgn.go:375-433.  (The reference's non-deterministic L2 branch returns L2=false,
        bgn.go:411 -- a flag bug; the flag is reported correctly here.)"""
        a, b = self._promote(ct1, ct2)
        if a.L2:
            res = self.engine.gt_div_batch(self._np(a), self._np(b))
            if not self.Deterministic:
                res = self._blind_gt(res, r)
            return Ciphertext(res.tobytes(), True)
        res = self.engine.g1_sub_batch(self._np(a), self._np(b))
        if not self.Deterministic:
            res = self._blind_g1(res, r)
        return Ciphertext(res.tobytes(), False)

    def Neg(self, c: Ciphertext, r: Optional[int] = None) -> Ciphertext:
        """bgn.go:436-439: Sub(encryptZero(), c) (an L2 `c` promotes the zero: e(O,P) = 1)."""
        return self.Sub(self.encryptZero(), c, r)

    def Mult(self, ct1: Ciphertext, ct2: Ciphertext, r: Optional[int] = None) -> Ciphertext:
        """bgn.go:294-314."""
        if ct1.L2 or ct2.L2:
            raise ValueError("Mult needs two level-1 ciphertexts")
        res = self.engine.pair_batch(self._np(ct1), self._np(ct2))
        if not self.Deterministic:
            res = self._blind_gt(res, r)
        return Ciphertext(res.tobytes(), True)

    def MultConst(self, c: Ciphertext, constant: int, r: Optional[int] = None) -> Ciphertext:
        """bgn.go:253-291."""
        if constant < 0:
            raise ValueError("negative constants: use Neg(MultConst(c, -k))")
        kb = max(1, (constant.bit_length() + 7) // 8)
        k = self.engine.scalars_be([constant], kb)
        if c.L2:
            res = self.engine.gt_pow_batch(self._np(c), k, kb)
            if not self.Deterministic:
                res = self._blind_gt(res, r)
            return Ciphertext(res.tobytes(), True)
        res = self.engine.g1_mulconst_batch(self._np(c), k, kb)
        if not self.Deterministic:
            res = self._blind_g1(res, r)
        return Ciphertext(res.tobytes(), False)

    # ---------------------------------------------------------------- polynomial ciphertexts
    def _split(self, buf, count: int, L2: bool) -> List[Ciphertext]:
        raw = buf.tobytes() if hasattr(buf, "tobytes") else bytes(buf.cpu().numpy().tobytes())
        eb = self.elem_bytes
        return [Ciphertext(raw[i * eb:(i + 1) * eb], L2) for i in range(count)]

    def EncryptPoly(self, pt: PolyPlaintext, rs: Optional[Sequence[int]] = None) -> PolyCiphertext:
        """poly.go:11-29; negative coefficients become -(|c| P + r Q) as there."""
        x = np.array(pt.Coefficients[: pt.Degree], dtype=np.int64)
        if rs is None:
            rs = [secrets.randbelow(self.N) for _ in range(pt.Degree)]
        out = self.engine.encrypt_batch(x, self.engine.scalars_be([r % self.N for r in rs]))
        return PolyCiphertext(self._split(out, pt.Degree, False), pt.Degree, pt.ScaleFactor, False)

    def NegPoly(self, ct: PolyCiphertext) -> PolyCiphertext:
        """poly.go:45-55."""
        buf = self._np(*ct.Coefficients[: ct.Degree])
        res = self.engine.gt_inv_batch(buf) if ct.L2 else self.engine.g1_neg_batch(buf)
        return PolyCiphertext(self._split(res, ct.Degree, ct.L2), ct.Degree, ct.ScaleFactor, ct.L2)

    def MultPoly(self, ct1: PolyCiphertext, ct2: PolyCiphertext) -> PolyCiphertext:
        """poly.go:123-156: result[j] = prod_{i+k=j} e(c1[i], c2[k]); Degree = d1 + d2 slots."""
        if ct1.L2 or ct2.L2:
            raise ValueError("MultPoly needs two level-1 ciphertexts")
        d1, d2 = ct1.Degree, ct2.Degree
        out = self.engine.multpoly_batch(self._np(*ct1.Coefficients[:d1]), d1, self._np(*ct2.Coefficients[:d2]), d2, 1)
        return PolyCiphertext(self._split(out, d1 + d2, True), d1 + d2, ct1.ScaleFactor + ct2.ScaleFactor, True)

    def MakePolyL2(self, ct: PolyCiphertext) -> PolyCiphertext:
        """poly.go:159-163: MultPoly(E(1.0), ct); E(1.0) is one slot, so Degree grows by one."""
        one = self.EncryptPoly(self.NewPolyPlaintext(1.0), rs=[0])
        return self.MultPoly(one, ct)

    def MultConstPoly(self, ct: PolyCiphertext, constant: float) -> PolyCiphertext:
        """poly.go:71-120: schoolbook product with the unbalanced digits (0,1,2) of |constant|."""
        negative = constant < 0
        poly = self.NewUnbalancedPlaintext(abs(constant))
        degree = ct.Degree + poly.Degree
        eb = self.elem_bytes
        ident = self.makeL2(self.encryptZero()) if ct.L2 else self.encryptZero()
        acc = np.frombuffer(ident.C * degree, dtype=np.uint8).copy()
        src = self._np(*ct.Coefficients[: ct.Degree])
        for k, digit in enumerate(poly.Coefficients[: poly.Degree]):
            if digit == 0:
                continue
            kbuf = self.engine.scalars_be([digit] * ct.Degree, 1)
            term = (self.engine.gt_pow_batch(src, kbuf, 1) if ct.L2 else self.engine.g1_mulconst_batch(src, kbuf, 1))
            window = acc[k * eb:(k + ct.Degree) * eb]
            summed = self.engine.gt_mul_batch(window, term) if ct.L2 else self.engine.g1_add_batch(window, term)
            acc[k * eb:(k + ct.Degree) * eb] = summed
        prod = PolyCiphertext(self._split(acc, degree, ct.L2), degree, ct.ScaleFactor + poly.ScaleFactor, ct.L2)
        return self.NegPoly(prod) if negative else prod

    def alignPolyCiphertexts(self, ct1: PolyCiphertext, ct2: PolyCiphertext):
        """poly.go:209-226."""
        if ct1.ScaleFactor > ct2.ScaleFactor:
            diff = ct1.ScaleFactor - ct2.ScaleFactor
            ct2 = self.MultConstPoly(ct2, math.pow(float(self.PolyEncodingParams.FPScaleBase), float(diff)))
            ct2.ScaleFactor = ct1.ScaleFactor
        elif ct2.ScaleFactor > ct1.ScaleFactor:
            return self.alignPolyCiphertexts(ct2, ct1)
        return ct1, ct2

    def AddPoly(self, pct1: PolyCiphertext, pct2: PolyCiphertext) -> PolyCiphertext:
        """poly.go:171-207."""
        if pct1.L2 or pct2.L2:
            if not pct1.L2:
                return self.AddPoly(self.MakePolyL2(pct1), pct2)
            if not pct2.L2:
                return self.AddPoly(pct1, self.MakePolyL2(pct2))
        ct1, ct2 = self.alignPolyCiphertexts(pct1.Copy(), pct2.Copy())
        degree = max(ct1.Degree, ct2.Degree)
        common = min(ct1.Degree, ct2.Degree)
        a, b = self._np(*ct1.Coefficients[:common]), self._np(*ct2.Coefficients[:common])
        res = self.engine.gt_mul_batch(a, b) if ct1.L2 else self.engine.g1_add_batch(a, b)
        coeffs = self._split(res, common, ct1.L2)
        longer = ct1 if ct1.Degree > ct2.Degree else ct2
        coeffs += longer.Coefficients[common:degree]
        return PolyCiphertext(coeffs, degree, ct1.ScaleFactor, ct1.L2)

    def SubPoly(self, ct1: PolyCiphertext, ct2: PolyCiphertext) -> PolyCiphertext:
        return self.AddPoly(ct1, self.NegPoly(ct2))  # poly.go:166-168

    def EvalPoly(self, ct: PolyCiphertext) -> Ciphertext:
        """poly.go:58-68 (Horner in the exponent)."""
        acc = self.EncryptDeterministic(0)
        for c in reversed(ct.Coefficients[: ct.Degree]):
            acc = self.MultConst(acc, self.PolyEncodingParams.PolyBase)
            acc = self.Add(acc, c)
        return acc

    # ---------------------------------------------------------------- batch entry points (new)
    def EncryptPolyBatch(self, coeffs, r_be, ScaleFactor: int = 0) -> PolyCiphertextBatch:
        """coeffs: int64 [count, Degree] (host array or CUDA tensor) of polynomial digits;
        r_be: count*Degree big-endian scalars (scalar_bytes each)."""
        count, degree = coeffs.shape
        flat = coeffs.reshape(-1)
        out = self.engine.encrypt_batch(flat, r_be)
        return PolyCiphertextBatch(out, count, degree, ScaleFactor, False)

    @staticmethod
    def _rows(data, count: int, degree: int, eb: int):
        """a batch buffer as [count, degree * eb] (numpy array or torch tensor alike)"""
        return data.reshape(count, degree * eb)

    @staticmethod
    def _flat(x):
        x = x.contiguous() if hasattr(x, "contiguous") else np.ascontiguousarray(x)
        return x.reshape(-1)

    def AddPolyBatch(self, a: PolyCiphertextBatch, b: PolyCiphertextBatch) -> PolyCiphertextBatch:
        """AddPoly (poly.go:171-207) over two batches of equal count, with everything the scalar
        version does: a level-1 operand is promoted with MakePolyL2 when the other is level 2
        (poly.go:173-182), the operand with the smaller ScaleFactor is multiplied by
        FPScaleBase^diff (alignPolyCiphertexts, poly.go:209-226), the common slots are added and the
        longer operand's tail is passed through (poly.go:191-204)."""
        if a.count != b.count:
            raise ValueError("AddPolyBatch needs batches of equal count")
        if a.L2 != b.L2:
            if not a.L2:
                a = self.MakePolyL2Batch(a)
            else:
                b = self.MakePolyL2Batch(b)
        if a.ScaleFactor != b.ScaleFactor:
            lo, hi = (a, b) if a.ScaleFactor < b.ScaleFactor else (b, a)
            diff = hi.ScaleFactor - lo.ScaleFactor
            up = self.MultConstPolyBatch(lo, math.pow(float(self.PolyEncodingParams.FPScaleBase), float(diff)))
            up.ScaleFactor = hi.ScaleFactor
            a, b = hi, up
        eb = self.elem_bytes
        if a.Degree == b.Degree:
            res = self.engine.gt_mul_batch(a.data, b.data) if a.L2 else self.engine.g1_add_batch(a.data, b.data)
            return PolyCiphertextBatch(res, a.count, a.Degree, a.ScaleFactor, a.L2)
        if a.Degree < b.Degree:
            a, b = b, a  # a is the longer one; addition is commutative
        common = b.Degree
        ra = self._rows(a.data, a.count, a.Degree, eb)
        head = self._flat(ra[:, : common * eb])
        summed = self.engine.gt_mul_batch(head, b.data) if a.L2 else self.engine.g1_add_batch(head, b.data)
        out = ra.clone() if hasattr(ra, "clone") else ra.copy()
        sm = self._rows(summed, a.count, common, eb)
        if hasattr(out, "is_cuda") and not hasattr(sm, "is_cuda"):
            import torch
            sm = torch.from_numpy(sm).to(out.device)
        elif not hasattr(out, "is_cuda") and hasattr(sm, "is_cuda"):
            sm = sm.cpu().numpy()
        out[:, : common * eb] = sm
        return PolyCiphertextBatch(out.reshape(-1), a.count, a.Degree, a.ScaleFactor, a.L2)

    def MultPolyBatch(self, a: PolyCiphertextBatch, b: PolyCiphertextBatch, rs=None) -> PolyCiphertextBatch:
        """MultPoly over a batch.  Non-deterministic keys re-randomise every output slot: the
        reference multiplies e(Q,Q)^r into each of the d1*d2 coefficient pairings (bgn.go:302-311 via
        poly.go:140-152), which per slot is e(Q,Q)^(sum of its r's); `rs` injects one scalar per
        output slot (count * (d1+d2-1) ... the padding slot included: count * (d1+d2))."""
        if a.L2 or b.L2 or a.count != b.count:
            raise ValueError("MultPolyBatch needs two level-1 batches of equal count")
        out = self.engine.multpoly_batch(a.data, a.Degree, b.data, b.Degree, a.count)
        if not self.Deterministic:
            out = self.engine.gt_blind_batch(out, self._rand_be(a.count * (a.Degree + b.Degree), rs))
        return PolyCiphertextBatch(out, a.count, a.Degree + b.Degree, a.ScaleFactor + b.ScaleFactor, True)

    def MakePolyL2Batch(self, a: PolyCiphertextBatch) -> PolyCiphertextBatch:
        """MakePolyL2 (poly.go:159-163) over a batch: e(c_i, P) per slot plus the identity top slot."""
        if a.L2:
            raise ValueError("MakePolyL2Batch needs a level-1 batch")
        out = self.engine.make_poly_l2_batch(a.data, a.Degree, a.count)
        return PolyCiphertextBatch(out, a.count, a.Degree + 1, a.ScaleFactor, True)

    def MultConstPolyBatch(self, a: PolyCiphertextBatch, constant: float) -> PolyCiphertextBatch:
        """MultConstPoly (poly.go:71-120) over a batch: one kernel launch, one thread per output slot."""
        poly = self.NewUnbalancedPlaintext(abs(constant))
        out = self.engine.multconstpoly_batch(a.data, a.Degree, a.L2, poly.Coefficients[: poly.Degree], constant < 0,
                                              a.count)
        return PolyCiphertextBatch(out, a.count, a.Degree + poly.Degree, a.ScaleFactor + poly.ScaleFactor, a.L2)

    def EvalPolyBatch(self, a: PolyCiphertextBatch):
        """EvalPoly (poly.go:58-68) over a batch -> count elements (level of the batch)."""
        return self.engine.evalpoly_batch(a.data, a.Degree, a.L2, self.PolyEncodingParams.PolyBase, a.count)

    def AddPolyBatchRand(self, a: PolyCiphertextBatch, b: PolyCiphertextBatch, rs=None) -> PolyCiphertextBatch:
        """AddPolyBatch followed by the per-coefficient re-randomisation of a non-deterministic key."""
        res = self.AddPolyBatch(a, b)
        if not self.Deterministic:
            blind = self.engine.gt_blind_batch if res.L2 else self.engine.g1_blind_batch
            res.data = blind(res.data, self._rand_be(res.count * res.Degree, rs))
        return res

    def SumPolyBatch(self, a: PolyCiphertextBatch) -> PolyCiphertext:
        """AddPoly folded over a whole L2 batch (one GPU's share of an inner product)."""
        if not a.L2:
            raise ValueError("SumPolyBatch works on level-2 batches")
        out = self.engine.l2_sum_reduce(a.data, a.count, a.Degree)
        return PolyCiphertext(self._split(out, a.Degree, True), a.Degree, a.ScaleFactor, True)

    def InnerProduct(self, u: PolyCiphertextBatch, v: PolyCiphertextBatch) -> PolyCiphertext:
        """sum_i u[i]*v[i] as one L2 polynomial ciphertext (BASELINE.json config 5)."""
        return self.SumPolyBatch(self.MultPolyBatch(u, v))


def ComputeDecryptionPreprocessing(pk: PublicKey, sk: SecretKey) -> None:
    """bgn.go:140-149: the package-level spelling of pk.SetupDecryption(sk)."""
    pk.SetupDecryption(sk)
