"""TEST INFRASTRUCTURE: a stand-in for bgn_b200.engine.Engine whose every call is answered by the CPU
oracle (oracle/bgn_oracle.py).  It exists so that the HOST logic of bgn_b200/bgn.py -- level promotion,
scale-factor alignment, tail pass-through, slot bookkeeping, wire formats, error behaviour -- runs in the
`-m "not gpu"` suite.  It doubles as an executable statement of what each C-ABI entry point computes.
Nothing under bgn_b200/ imports it; the product has no CPU path."""
from __future__ import annotations

import numpy as np

from oracle import bgn_oracle as O


class FakeEngine:
    def __init__(self, p, n, l, P_bytes, Q_bytes):
        self.par = O.A1Params(p, n, l)
        self.p, self.n, self.l = p, n, l
        self.P = O.g1_from_bytes(bytes(P_bytes), self.par)
        self.Q = O.g1_from_bytes(bytes(Q_bytes), self.par)
        self.coord_bytes = self.par.coord_bytes
        self.elem_bytes = 2 * self.coord_bytes
        self.scalar_bytes = (n.bit_length() + 7) // 8
        self.pk = O.PublicKey(self.par, self.P, self.Q, 0)
        self.sk = None
        self._qq = None

    # ---- helpers
    def _raw(self, x) -> bytes:
        return bytes(x) if isinstance(x, (bytes, bytearray)) else np.ascontiguousarray(x).tobytes()

    def _g1(self, buf):
        raw, eb = self._raw(buf), self.elem_bytes
        return [O.g1_from_bytes(raw[i:i + eb], self.par) for i in range(0, len(raw), eb)]

    def _gt(self, buf):
        raw, eb = self._raw(buf), self.elem_bytes
        return [O.gt_from_bytes(raw[i:i + eb], self.par) for i in range(0, len(raw), eb)]

    def _scal(self, buf, width):
        raw = self._raw(buf)
        return [int.from_bytes(raw[i:i + width], "big") for i in range(0, len(raw), width)]

    def _out1(self, pts):
        return np.frombuffer(b"".join(O.g1_to_bytes(x, self.par) for x in pts), dtype=np.uint8).copy()

    def _out2(self, vals):
        return np.frombuffer(b"".join(O.gt_to_bytes(x, self.par) for x in vals), dtype=np.uint8).copy()

    def scalars_be(self, ks, width=None):
        width = width or self.scalar_bytes
        return np.frombuffer(b"".join(int(k).to_bytes(width, "big") for k in ks), dtype=np.uint8).copy()

    def close(self):
        pass

    # ---- C-ABI entry points
    def set_secret(self, q1, msg_space, baby_steps=0):
        self.sk = O.SecretKey(q1, 0)
        self.pk.msg_space = msg_space
        O.setup_decryption(self.pk, self.sk)

    def encrypt_batch(self, x, r_be=None, out=None):
        xs = [int(v) for v in np.asarray(x).reshape(-1)]
        rs = self._scal(r_be, self.scalar_bytes) if r_be is not None else [0] * len(xs)
        pts = []
        for v, r in zip(xs, rs):
            c = O.encrypt_with_randomness(self.pk, abs(v), r).C
            pts.append(O.g1_neg(c, self.p) if v < 0 else c)
        return self._out1(pts)

    def g1_add_batch(self, a, b, out=None):
        return self._out1([O.g1_add(x, y, self.p) for x, y in zip(self._g1(a), self._g1(b))])

    def g1_sub_batch(self, a, b, out=None):
        return self._out1([O.g1_add(x, O.g1_neg(y, self.p), self.p) for x, y in zip(self._g1(a), self._g1(b))])

    def g1_neg_batch(self, a, out=None):
        return self._out1([O.g1_neg(x, self.p) for x in self._g1(a)])

    def g1_mulconst_batch(self, a, k_be, kbytes, out=None):
        return self._out1([O.g1_mul(k, x, self.p) for x, k in zip(self._g1(a), self._scal(k_be, kbytes))])

    def gt_mul_batch(self, a, b, out=None):
        return self._out2([O.fp2_mul(x, y, self.p) for x, y in zip(self._gt(a), self._gt(b))])

    def gt_div_batch(self, a, b, out=None):
        return self._out2([O.fp2_mul(x, O.fp2_inv(y, self.p), self.p) for x, y in zip(self._gt(a), self._gt(b))])

    def gt_inv_batch(self, a, out=None):
        return self._out2([O.fp2_inv(x, self.p) for x in self._gt(a)])

    def gt_pow_batch(self, a, k_be, kbytes, out=None):
        return self._out2([O.fp2_pow(x, k, self.p) for x, k in zip(self._gt(a), self._scal(k_be, kbytes))])

    def pair_batch(self, a, b, out=None):
        return self._out2([O.pairing(x, y, self.par) for x, y in zip(self._g1(a), self._g1(b))])

    def make_l2_batch(self, a, out=None):
        return self._out2([O.pairing(x, self.P, self.par) for x in self._g1(a)])

    def multpoly_batch(self, c1, d1, c2, d2, count, out=None):
        A, B = self._g1(c1), self._g1(c2)
        res = []
        for u in range(count):
            acc = [O.GT_ONE] * (d1 + d2)
            for i in range(d1):
                for k in range(d2):
                    acc[i + k] = O.fp2_mul(acc[i + k], O.pairing(A[u * d1 + i], B[u * d2 + k], self.par), self.p)
            res += acc
        return self._out2(res)

    def l2_sum_reduce(self, terms, nterms, ncoeff, out=None):
        T = self._gt(terms)
        acc = [O.GT_ONE] * ncoeff
        for t in range(nterms):
            for c in range(ncoeff):
                acc[c] = O.fp2_mul(acc[c], T[t * ncoeff + c], self.p)
        return self._out2(acc)

    def decrypt_batch(self, cts, is_l2):
        if self.sk is None:
            raise RuntimeError("DL tables not computed!")
        elems = self._gt(cts) if is_l2 else self._g1(cts)
        vals, st = [], []
        for e in elems:
            try:
                vals.append(O.decrypt(self.pk, self.sk, O.Ciphertext(e, bool(is_l2))))
                st.append(0)
            except O.DLError:
                vals.append(0)
                st.append(1)
        return np.array(vals, dtype=np.int64), np.array(st, dtype=np.uint8)

    def g1_blind_batch(self, a, r_be, out=None):
        rs = self._scal(r_be, self.scalar_bytes)
        return self._out1([O.g1_add(x, O.g1_mul(r, self.Q, self.p), self.p) for x, r in zip(self._g1(a), rs)])

    def gt_blind_batch(self, a, r_be, out=None):
        if self._qq is None:
            self._qq = O.pairing(self.Q, self.Q, self.par)
        rs = self._scal(r_be, self.scalar_bytes)
        return self._out2([O.fp2_mul(x, O.fp2_pow(self._qq, r, self.p), self.p) for x, r in zip(self._gt(a), rs)])

    def _conv(self, elems, d, is_l2, w, j_begin, j_count, negate, count):
        res = []
        for u in range(count):
            for jj in range(j_count):
                j = j_begin + jj
                acc = O.GT_ONE if is_l2 else None
                for k, wk in enumerate(w):
                    i = j - k
                    if 0 <= i < d and wk:
                        e = elems[u * d + i]
                        acc = O.fp2_mul(acc, O.fp2_pow(e, wk, self.p), self.p) if is_l2 else O.g1_add(
                            acc, O.g1_mul(wk, e, self.p), self.p)
                if negate:
                    acc = O.fp2_inv(acc, self.p) if is_l2 else O.g1_neg(acc, self.p)
                res.append(acc)
        return self._out2(res) if is_l2 else self._out1(res)

    def multconstpoly_batch(self, cts, d, is_l2, digits, negate, count, out=None):
        elems = self._gt(cts) if is_l2 else self._g1(cts)
        return self._conv(elems, d, is_l2, list(digits), 0, d + len(digits), negate, count)

    def evalpoly_batch(self, cts, d, is_l2, base, count, out=None):
        elems = self._gt(cts) if is_l2 else self._g1(cts)
        return self._conv(elems, d, is_l2, [base ** (d - 1 - k) for k in range(d)], d - 1, 1, False, count)

    def make_poly_l2_batch(self, cts, d, count, out=None):
        A = self._g1(cts)
        res = []
        for u in range(count):
            res += [O.pairing(A[u * d + i], self.P, self.par) for i in range(d)] + [O.GT_ONE]
        return self._out2(res)
