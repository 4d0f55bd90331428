"""Host-side mirror of the reference's Go API (bgn.go, poly.go, ciphertext.go).

Same names, argument meaning and error behaviour as the reference so that the
parity tests read like bgn_test.go / poly_test.go; every group operation is one
call into the CUDA library through the C-ABI (bgn_b200.engine.Engine).  The
single-element methods (Encrypt, Add, Mult, ...) are batches of one; the
`*Batch` methods are the new entry points north_star asks for and are what a
production caller uses.

What differs from the reference, on purpose:
  * elements are held as PBC `element_to_bytes` byte strings, not *pbc.Element;
  * randomness is injected (`r=`, `rs=`, or a `PublicKey.rand_source` callable that stands for
    newCryptoRandom, consumed in the reference's sequential program order) or drawn from
    `secrets`; the reference draws from crypto/rand inside the call (bgn.go:567-574);
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
        self._secret_key = None
        # newCryptoRandom(pk.N) (bgn.go:567-574).  None: the OS CSPRNG.  A callable: every draw the
        # reference would make that is not given explicitly (`r=`, `rs=`) is one call of it, in the
        # reference's sequential program order (goroutine bodies in the textual order of their loops,
        # poly.go:101-112, 144-151) -- what the parity tests use to reproduce the oracle byte for byte.
        self.rand_source = None

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
        self._secret_key = sk.Key

    def _need_secret(self, sk: SecretKey) -> Engine:
        if not self._secret_set:
            raise RuntimeError("DL tables not computed!")  # gsbs.go:56-58
        if sk is not None and sk.Key != self._secret_key:
            raise ValueError("this SecretKey is not the one installed by SetupDecryption")
        return self.engine

    # ---------------------------------------------------------------- helpers
    @property
    def elem_bytes(self) -> int:
        return self.engine.elem_bytes

    def _np(self, *cts: Ciphertext) -> np.ndarray:
        return np.frombuffer(b"".join(c.C for c in cts), dtype=np.uint8)

    def _rand(self, r: Optional[int] = None) -> int:
        if r is not None:
            return r
        if self.rand_source is not None:
            return int(self.rand_source()) % self.N
        return secrets.randbelow(self.N)

    def _blind(self, buf, scalars: Sequence[int], L2: bool):
        """the re-randomisation of a non-deterministic key over a whole buffer: slot i gains
        scalars[i]*Q (level 1, bgn.go:264-268, 491-495) or is multiplied by e(Q,Q)^scalars[i] (level 2,
        bgn.go:283-287, 306-310, 469-474); a scalar 0 leaves its slot untouched."""
        r_be = self.engine.scalars_be([int(x) % self.N for x in scalars])
        if type(buf).__module__.startswith("torch") and buf.is_cuda:
            import torch
            r_be = torch.from_numpy(r_be).to(buf.device)
        return self.engine.gt_blind_batch(buf, r_be) if L2 else self.engine.g1_blind_batch(buf, r_be)

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
        if abs(x) >= 1 << 63:  # the reference takes any *big.Int
            return Ciphertext(self._encrypt_big([x], [0]), False)
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
        negative = constant < 0  # k*C = |k|*(-C): the group-theoretic meaning of PowBig by a negative k
        constant = abs(constant)
        kb = max(1, (constant.bit_length() + 7) // 8)
        k = self.engine.scalars_be([constant], kb)
        if c.L2:
            res = self.engine.gt_pow_batch(self._np(c), k, kb)
            if negative:
                res = self.engine.gt_inv_batch(res)
            if not self.Deterministic:
                res = self._blind_gt(res, r)
            return Ciphertext(res.tobytes(), True)
        res = self.engine.g1_mulconst_batch(self._np(c), k, kb)
        if negative:
            res = self.engine.g1_neg_batch(res)
        if not self.Deterministic:
            res = self._blind_g1(res, r)
        return Ciphertext(res.tobytes(), False)

    # ---------------------------------------------------------------- polynomial ciphertexts
    def _split(self, buf, count: int, L2: bool) -> List[Ciphertext]:
        raw = buf.tobytes() if hasattr(buf, "tobytes") else bytes(buf.cpu().numpy().tobytes())
        eb = self.elem_bytes
        return [Ciphertext(raw[i * eb:(i + 1) * eb], L2) for i in range(count)]

    def EncryptPoly(self, pt: PolyPlaintext, rs: Optional[Sequence[int]] = None) -> PolyCiphertext:
        """poly.go:11-29; rs[i] is the randomness of coefficient i's Encrypt.  Negative coefficients are
        Sub(encryptZero(), Encrypt(|c|)) = -(|c| P + r Q) as there; with a non-deterministic key that Sub
        adds its own r' Q (bgn.go:421-432), drawn here when a rand_source is installed or `rs` is absent:
        -(|c| P + (r - r') Q), one engine call either way."""
        coeffs = [int(c) for c in pt.Coefficients[: pt.Degree]]
        x = np.array(coeffs, dtype=np.int64)
        draw_sub = (not self.Deterministic) and (rs is None or self.rand_source is not None)
        eff = []
        for i, c in enumerate(coeffs):  # the reference's loop order (ascending), Encrypt's draw first
            r = self._rand(None if rs is None else rs[i])
            if c < 0 and draw_sub:
                r -= self._rand()
            eff.append(r % self.N)
        out = self.engine.encrypt_batch(x, self.engine.scalars_be(eff))
        return PolyCiphertext(self._split(out, pt.Degree, False), pt.Degree, pt.ScaleFactor, False)

    def NegPoly(self, ct: PolyCiphertext) -> PolyCiphertext:
        """poly.go:45-55: Sub(encryptZero(), c_i), i = Degree-1 .. 0 -- one draw per slot for a
        non-deterministic key."""
        d = ct.Degree
        buf = self._np(*ct.Coefficients[:d])
        res = self.engine.gt_inv_batch(buf) if ct.L2 else self.engine.g1_neg_batch(buf)
        if not self.Deterministic:
            res = self._blind(res, [self._rand() for _ in range(d)][::-1], ct.L2)
        return PolyCiphertext(self._split(res, d, ct.L2), d, ct.ScaleFactor, ct.L2)

    def _multpoly_scalars(self, d1: int, d2: int) -> List[int]:
        """Per output slot, the sum of the draws the reference makes for it in MultPoly: one in Mult
        (bgn.go:302-311) and one in the Add that folds the pairing into its slot (bgn.go:466-474), for
        i = d1-1 .. 0, k = d2-1 .. 0 (poly.go:140-152).  The unused top slot draws nothing and stays 1."""
        R = [0] * (d1 + d2)
        for i in range(d1 - 1, -1, -1):
            for k in range(d2 - 1, -1, -1):
                R[i + k] += self._rand() + self._rand()
        return R

    def MultPoly(self, ct1: PolyCiphertext, ct2: PolyCiphertext) -> PolyCiphertext:
        """poly.go:123-156: result[j] = prod_{i+k=j} e(c1[i], c2[k]); Degree = d1 + d2 slots.  A
        non-deterministic key multiplies e(Q,Q)^(its draws) into every slot but the unused top one."""
        if ct1.L2 or ct2.L2:
            raise ValueError("MultPoly needs two level-1 ciphertexts")
        d1, d2 = ct1.Degree, ct2.Degree
        out = self.engine.multpoly_batch(self._np(*ct1.Coefficients[:d1]), d1, self._np(*ct2.Coefficients[:d2]), d2, 1)
        if not self.Deterministic:
            out = self._blind(out, self._multpoly_scalars(d1, d2), True)
        return PolyCiphertext(self._split(out, d1 + d2, True), d1 + d2, ct1.ScaleFactor + ct2.ScaleFactor, True)

    def _one_randomness(self, r_one: Optional[int]) -> int:
        """The randomness of E(1.0) in MakePolyL2 (poly.go:161: EncryptPoly draws it, also for
        Deterministic keys).  Explicit, else drawn from the rand_source / the CSPRNG; a Deterministic key
        without a rand_source uses 0 -- a valid member of the reference's output distribution, the
        reproducible one, and the one served by the recorded line table of P."""
        if r_one is not None:
            return r_one % self.N
        if self.Deterministic and self.rand_source is None:
            return 0
        return self._rand()

    def MakePolyL2(self, ct: PolyCiphertext, r_one: Optional[int] = None) -> PolyCiphertext:
        """poly.go:159-163: MultPoly(EncryptPoly(1.0), ct); E(1.0) is one slot, so Degree grows by one."""
        one = self.EncryptPoly(self.NewPolyPlaintext(1.0), rs=[self._one_randomness(r_one)])
        return self.MultPoly(one, ct)

    def _multconst_scalars(self, d: int, nd: int, negative: bool) -> List[int]:
        """Per output slot of MultConstPoly, the net re-randomisation scalar of a non-deterministic key:
        every (coefficient i, digit k) pair -- zero digits too -- draws once in MultConst (bgn.go:260-269,
        279-288) and once in Add, for i = d-1 .. 0, k = nd-1 .. 0 (poly.go:97-112); a negative constant
        negates all of that and adds NegPoly's draw per slot (poly.go:116-118)."""
        R = [0] * (d + nd)
        for i in range(d - 1, -1, -1):
            for k in range(nd - 1, -1, -1):
                R[i + k] += self._rand() + self._rand()
        if negative:
            neg = [self._rand() for _ in range(d + nd)][::-1]
            R = [n_ - r_ for n_, r_ in zip(neg, R)]
        return R

    def MultConstPoly(self, ct: PolyCiphertext, constant: float) -> PolyCiphertext:
        """poly.go:71-120: schoolbook product with the unbalanced digits (0,1,2) of |constant|, one
        engine call (bgn_multconstpoly_batch)."""
        negative = constant < 0
        poly = self.NewUnbalancedPlaintext(abs(constant))
        d, nd = ct.Degree, poly.Degree
        out = self.engine.multconstpoly_batch(self._np(*ct.Coefficients[:d]), d, ct.L2, poly.Coefficients[:nd],
                                              negative, 1)
        if not self.Deterministic:
            out = self._blind(out, self._multconst_scalars(d, nd, negative), ct.L2)
        return PolyCiphertext(self._split(out, d + nd, ct.L2), d + nd, ct.ScaleFactor + poly.ScaleFactor, ct.L2)

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
        """poly.go:171-207: the common slots go through Add (re-randomised for a non-deterministic key,
        i = degree-1 .. 0), the longer operand's tail is passed through untouched."""
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
        if not self.Deterministic:
            res = self._blind(res, [self._rand() for _ in range(common)][::-1], ct1.L2)
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

    # The batch forms draw exactly what `count` sequential calls of the scalar form would, unit by unit
    # (so a replayed rand_source reproduces them), and re-randomise with ONE blind call per stage.
    def _flatten(self, per_unit: List[List[int]]) -> List[int]:
        return [x for row in per_unit for x in row]

    def _addpoly_exec(self, a: PolyCiphertextBatch, b: PolyCiphertextBatch, Rs) -> PolyCiphertextBatch:
        """common slots added (and blinded with Rs[u][0..common) when given), longer tail passed through"""
        eb = self.elem_bytes
        if a.Degree < b.Degree:
            a, b = b, a  # a is the longer one; addition is commutative
        common = b.Degree
        if a.Degree == common:
            head = a.data
        else:
            ra = self._rows(a.data, a.count, a.Degree, eb)
            head = self._flat(ra[:, : common * eb])
        summed = self.engine.gt_mul_batch(head, b.data) if a.L2 else self.engine.g1_add_batch(head, b.data)
        if Rs is not None:
            summed = self._blind(summed, self._flatten(Rs), a.L2)
        if a.Degree == common:
            return PolyCiphertextBatch(summed, a.count, a.Degree, a.ScaleFactor, a.L2)
        out = ra.clone() if hasattr(ra, "clone") else ra.copy()
        sm = self._rows(summed, a.count, common, eb)
        if hasattr(out, "is_cuda") and not hasattr(sm, "is_cuda"):
            import torch
            sm = torch.from_numpy(sm).to(out.device)
        elif not hasattr(out, "is_cuda") and hasattr(sm, "is_cuda"):
            sm = sm.cpu().numpy()
        out[:, : common * eb] = sm
        return PolyCiphertextBatch(out.reshape(-1), a.count, a.Degree, a.ScaleFactor, a.L2)

    def AddPolyBatch(self, a: PolyCiphertextBatch, b: PolyCiphertextBatch) -> PolyCiphertextBatch:
        """AddPoly (poly.go:171-207) over two batches of equal count, with everything the scalar
        version does: a level-1 operand is promoted with MakePolyL2 when the other is level 2
        (poly.go:173-182), the operand with the smaller ScaleFactor is multiplied by
        FPScaleBase^diff (alignPolyCiphertexts, poly.go:209-226), the common slots are added --
        re-randomised for a non-deterministic key -- and the longer operand's tail is passed
        through (poly.go:191-204)."""
        if a.count != b.count:
            raise ValueError("AddPolyBatch needs batches of equal count")
        count, nondet = a.count, not self.Deterministic
        promote = None if a.L2 == b.L2 else ("a" if not a.L2 else "b")
        d_prom = 0 if promote is None else (a.Degree if promote == "a" else b.Degree)
        da = a.Degree + (1 if promote == "a" else 0)
        db = b.Degree + (1 if promote == "b" else 0)
        align, const, nd = None, 0.0, 0
        if a.ScaleFactor != b.ScaleFactor:
            align = "a" if a.ScaleFactor < b.ScaleFactor else "b"
            const = math.pow(float(self.PolyEncodingParams.FPScaleBase), float(abs(a.ScaleFactor - b.ScaleFactor)))
            nd = self.NewUnbalancedPlaintext(const).Degree
            if align == "a":
                da += nd
            else:
                db += nd
        common = min(da, db)
        # the draws of unit 0's AddPoly, then unit 1's, ... in the scalar form's order
        r_ones, R_l2, R_mc, R_add = [], [], [], []
        for _ in range(count):
            if promote is not None:
                r_ones.append(self._one_randomness(None))
                R_l2.append(self._multpoly_scalars(1, d_prom) if nondet else None)
            if align is not None and nondet:
                d_al = (a.Degree if align == "a" else b.Degree) + (1 if promote == align else 0)
                R_mc.append(self._multconst_scalars(d_al, nd, False))
            if nondet:
                R_add.append([self._rand() for _ in range(common)][::-1])
        if promote == "a":
            a = self._makepolyl2_exec(a, r_ones, R_l2 if nondet else None)
        elif promote == "b":
            b = self._makepolyl2_exec(b, r_ones, R_l2 if nondet else None)
        if align is not None:
            lo, hi = (a, b) if align == "a" else (b, a)
            up = self._multconstpoly_exec(lo, const, R_mc if nondet else None)
            up.ScaleFactor = hi.ScaleFactor
            a, b = hi, up
        return self._addpoly_exec(a, b, R_add if nondet else None)

    AddPolyBatchRand = AddPolyBatch  # round-1 name: AddPolyBatch itself honours Deterministic=False now

    def NegPolyBatch(self, a: PolyCiphertextBatch) -> PolyCiphertextBatch:
        """NegPoly (poly.go:45-55) over a batch."""
        res = self.engine.gt_inv_batch(a.data) if a.L2 else self.engine.g1_neg_batch(a.data)
        if not self.Deterministic:
            Rs = [[self._rand() for _ in range(a.Degree)][::-1] for _ in range(a.count)]
            res = self._blind(res, self._flatten(Rs), a.L2)
        return PolyCiphertextBatch(res, a.count, a.Degree, a.ScaleFactor, a.L2)

    def SubPolyBatch(self, a: PolyCiphertextBatch, b: PolyCiphertextBatch) -> PolyCiphertextBatch:
        return self.AddPolyBatch(a, self.NegPolyBatch(b))  # poly.go:166-168

    def MultPolyBatch(self, a: PolyCiphertextBatch, b: PolyCiphertextBatch, rs=None) -> PolyCiphertextBatch:
        """MultPoly over a batch.  Non-deterministic keys re-randomise every output slot but the
        unused top one: the reference multiplies e(Q,Q)^r into each of the d1*d2 coefficient pairings
        and again in the Add that folds each into its slot (bgn.go:302-311, 466-474 via
        poly.go:140-152), which per slot is e(Q,Q)^(sum of its draws).  `rs` injects the per-slot
        sums directly: count * (d1+d2) scalars, the top slot's entry ignored (it stays 1,
        poly.go:130-137)."""
        if a.L2 or b.L2 or a.count != b.count:
            raise ValueError("MultPolyBatch needs two level-1 batches of equal count")
        d = a.Degree + b.Degree
        out = self.engine.multpoly_batch(a.data, a.Degree, b.data, b.Degree, a.count)
        if not self.Deterministic:
            if rs is None:
                R = self._flatten([self._multpoly_scalars(a.Degree, b.Degree) for _ in range(a.count)])
            else:
                R = [0 if (i + 1) % d == 0 else int(r) for i, r in enumerate(rs)]
                if len(R) != a.count * d:
                    raise ValueError("rs must hold count * (d1 + d2) scalars")
            out = self._blind(out, R, True)
        return PolyCiphertextBatch(out, a.count, d, a.ScaleFactor + b.ScaleFactor, True)

    def _makepolyl2_exec(self, a: PolyCiphertextBatch, r_ones, Rs) -> PolyCiphertextBatch:
        if all(r == 0 for r in r_ones):
            out = self.engine.make_poly_l2_batch(a.data, a.Degree, a.count)  # e(., P): the recorded line table of P
        else:
            ones = self.engine.encrypt_batch(np.ones(a.count, dtype=np.int64), self.engine.scalars_be(r_ones))
            if type(a.data).__module__.startswith("torch") and a.data.is_cuda:
                import torch
                ones = torch.from_numpy(np.ascontiguousarray(ones)).to(a.data.device)
            out = self.engine.multpoly_batch(ones, 1, a.data, a.Degree, a.count)
        if Rs is not None:
            out = self._blind(out, self._flatten(Rs), True)
        return PolyCiphertextBatch(out, a.count, a.Degree + 1, a.ScaleFactor, True)

    def MakePolyL2Batch(self, a: PolyCiphertextBatch, r_one=None) -> PolyCiphertextBatch:
        """MakePolyL2 (poly.go:159-163) over a batch: e(E(1.0), c_i) per slot plus the identity top slot.
        r_one: the randomness of every polynomial's E(1.0) (a sequence of count scalars, or one int for
        all); see _one_randomness for the default."""
        if a.L2:
            raise ValueError("MakePolyL2Batch needs a level-1 batch")
        nondet = not self.Deterministic
        r_ones, Rs = [], []
        for u in range(a.count):
            ru = None if r_one is None else (r_one if isinstance(r_one, int) else r_one[u])
            r_ones.append(self._one_randomness(ru))
            if nondet:
                Rs.append(self._multpoly_scalars(1, a.Degree))
        return self._makepolyl2_exec(a, r_ones, Rs if nondet else None)

    def _multconstpoly_exec(self, a: PolyCiphertextBatch, constant: float, Rs) -> PolyCiphertextBatch:
        poly = self.NewUnbalancedPlaintext(abs(constant))
        out = self.engine.multconstpoly_batch(a.data, a.Degree, a.L2, poly.Coefficients[: poly.Degree], constant < 0,
                                              a.count)
        if Rs is not None:
            out = self._blind(out, self._flatten(Rs), a.L2)
        return PolyCiphertextBatch(out, a.count, a.Degree + poly.Degree, a.ScaleFactor + poly.ScaleFactor, a.L2)

    def MultConstPolyBatch(self, a: PolyCiphertextBatch, constant: float) -> PolyCiphertextBatch:
        """MultConstPoly (poly.go:71-120) over a batch: one kernel launch, one thread per output slot."""
        Rs = None
        if not self.Deterministic:
            nd = self.NewUnbalancedPlaintext(abs(constant)).Degree
            Rs = [self._multconst_scalars(a.Degree, nd, constant < 0) for _ in range(a.count)]
        return self._multconstpoly_exec(a, constant, Rs)

    def EvalPolyBatch(self, a: PolyCiphertextBatch):
        """EvalPoly (poly.go:58-68) over a batch -> count elements (level of the batch).  Horner draws
        once in MultConst and once in Add per coefficient (i = Degree-1 .. 0); the draws of step i are
        multiplied by base i more times, so the net scalar is sum_i (r_i + r'_i) base^i."""
        out = self.engine.evalpoly_batch(a.data, a.Degree, a.L2, self.PolyEncodingParams.PolyBase, a.count)
        if not self.Deterministic:
            base, R = self.PolyEncodingParams.PolyBase, []
            for _ in range(a.count):
                acc = 0
                for i in range(a.Degree - 1, -1, -1):
                    acc += (self._rand() + self._rand()) * pow(base, i, self.N)
                R.append(acc)
            out = self._blind(out, R, a.L2)
        return out

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
