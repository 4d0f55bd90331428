"""CPU oracle for the BGN hot path -- TEST INFRASTRUCTURE ONLY.

This module is the checker, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  ``bgn_b200`` never imports it.

PARITY UNPINNED.  The arithmetic of the reference (sachaservan/bgn) lives in a
third-party dependency that is absent from /root/reference:
``github.com/Nik-U/pbc v0.0.0-20181205041846-3e516ca0c5d6`` (go.mod:5), a cgo
wrapper over libpbc 0.5.14 (README.md:37) on GMP.  Neither Go nor libpbc exist
in this image and the reference's own tests hold no golden vectors
(bgn_test.go:15-85, poly_test.go:68-189 only check round trips with fresh
random keys).  The group / pairing layer below therefore restates libpbc's
*published* type-A1 definitions (a1_param.c, curve.c, fieldquadratic.c,
montfp.c of pbc-0.5.14); every output is the canonical representative of a
mathematically unique value (affine G1 point, reduced Tate pairing value), so
any implementation of the same definitions yields the same bytes.  The scheme
layer follows the reference file:line cited on each function.

Pure Python big-int arithmetic; sized for keyBits <= 1024 on small batches.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

G1Point = Optional[Tuple[int, int]]  # None == point at infinity O
Fp2 = Tuple[int, int]  # (re, im), i^2 = -1

# --------------------------------------------------------------------------
# primality (deterministic Miller-Rabin bases + a few pseudo-random ones)
# --------------------------------------------------------------------------
_SMALL_PRIMES = [2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71]


def is_probable_prime(m: int) -> bool:
    if m < 2:
        return False
    for q in _SMALL_PRIMES:
        if m % q == 0:
            return m == q
    d, s = m - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in _SMALL_PRIMES:
        x = pow(a, d, m)
        if x in (1, m - 1):
            continue
        for _ in range(s - 1):
            x = x * x % m
            if x == m - 1:
                break
        else:
            return False
    return True


# --------------------------------------------------------------------------
# type-A1 parameters  (libpbc a1_param.c: pbc_param_init_a1_gen)
# --------------------------------------------------------------------------
@dataclass(frozen=True)
class A1Params:
    p: int  # field prime, p = l*n - 1, p = 3 (mod 4)
    n: int  # group order (q1*q2)
    l: int  # cofactor, smallest multiple of 4 with l*n-1 prime

    @property
    def coord_bytes(self) -> int:
        """PBC fixed width of one serialised F_p coordinate."""
        return (self.p.bit_length() + 7) // 8

    def pbc_string(self) -> str:
        """PBC param text; the reference parses 'l' from it (bgn.go:583-593)."""
        return "type a1\np %d\nn %d\nl %d\n" % (self.p, self.n, self.l)


def a1_gen(n: int) -> A1Params:
    """l = smallest multiple of 4 such that p = l*n - 1 is prime."""
    l = 4
    while True:
        p = l * n - 1
        if is_probable_prime(p):
            return A1Params(p, n, l)
        l += 4


def a1_from_string(s: str) -> A1Params:
    vals = {}
    for line in s.strip().splitlines():
        k, v = line.split(None, 1)
        vals[k] = v
    assert vals["type"] == "a1"
    return A1Params(int(vals["p"]), int(vals["n"]), int(vals["l"]))


# --------------------------------------------------------------------------
# F_p^2 = F_p[i]/(i^2+1)     (libpbc fieldquadratic.c, "fi" variant)
# --------------------------------------------------------------------------
def fp2_mul(a: Fp2, b: Fp2, p: int) -> Fp2:
    return ((a[0] * b[0] - a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)


def fp2_sqr(a: Fp2, p: int) -> Fp2:
    return ((a[0] * a[0] - a[1] * a[1]) % p, (2 * a[0] * a[1]) % p)


def fp2_conj(a: Fp2, p: int) -> Fp2:
    return (a[0], (-a[1]) % p)


def fp2_inv(a: Fp2, p: int) -> Fp2:
    d = pow((a[0] * a[0] + a[1] * a[1]) % p, -1, p)
    return (a[0] * d % p, (-a[1]) * d % p)


def fp2_pow(a: Fp2, e: int, p: int) -> Fp2:
    if e < 0:
        return fp2_pow(fp2_inv(a, p), -e, p)
    r: Fp2 = (1, 0)
    for bit in bin(e)[2:] if e else "":
        r = fp2_sqr(r, p)
        if bit == "1":
            r = fp2_mul(r, a, p)
    return r


GT_ONE: Fp2 = (1, 0)


# --------------------------------------------------------------------------
# G1: E: y^2 = x^3 + x over F_p, affine    (libpbc curve.c with a=1, b=0)
# --------------------------------------------------------------------------
def g1_on_curve(P: G1Point, p: int) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - (x * x * x + x)) % p == 0


def g1_neg(P: G1Point, p: int) -> G1Point:
    if P is None:
        return None
    return (P[0], (-P[1]) % p)


def g1_dbl(P: G1Point, p: int) -> G1Point:
    if P is None:
        return None
    x, y = P
    if y == 0:
        return None
    lam = (3 * x * x + 1) * pow(2 * y, -1, p) % p
    x3 = (lam * lam - 2 * x) % p
    return (x3, (lam * (x - x3) - y) % p)


def g1_add(P: G1Point, Q: G1Point, p: int) -> G1Point:
    if P is None:
        return Q
    if Q is None:
        return P
    if P[0] == Q[0]:
        if (P[1] + Q[1]) % p == 0:
            return None
        return g1_dbl(P, p)
    lam = (Q[1] - P[1]) * pow(Q[0] - P[0], -1, p) % p
    x3 = (lam * lam - P[0] - Q[0]) % p
    return (x3, (lam * (P[0] - x3) - P[1]) % p)


def g1_mul(k: int, P: G1Point, p: int) -> G1Point:
    """[k]P; PBC's PowBig on G1 (multiplicative notation).  k<0 -> -[|k|]P."""
    if k < 0:
        return g1_neg(g1_mul(-k, P, p), p)
    R: G1Point = None
    for bit in bin(k)[2:] if k else "":
        R = g1_dbl(R, p)
        if bit == "1":
            R = g1_add(R, P, p)
    return R


# --------------------------------------------------------------------------
# reduced Tate pairing on A1   (libpbc a1_param.c: a1_pairing / a1_miller_evalfn)
#   e(P,Q) = f_{n,P}(phi(Q))^((p^2-1)/n),  phi(x,y) = (-x, i*y)
# --------------------------------------------------------------------------
def _line_eval(lam: int, V: Tuple[int, int], Q: Tuple[int, int], p: int) -> Fp2:
    # line through V with slope lam, evaluated at phi(Q) = (-xQ, i*yQ):
    # (i*yQ - yV) - lam*(-xQ - xV)
    return ((lam * (Q[0] + V[0]) - V[1]) % p, Q[1] % p)


def miller_loop(P: Tuple[int, int], Q: Tuple[int, int], par: A1Params) -> Fp2:
    p, n = par.p, par.n
    f: Fp2 = (1, 0)
    V: G1Point = P
    bits = bin(n)[3:]  # MSB-first, top bit consumed by V = P
    for bit in bits:
        assert V is not None and V[1] != 0, "degenerate Miller loop (point order divides a prefix of n)"
        lam = (3 * V[0] * V[0] + 1) * pow(2 * V[1], -1, p) % p
        f = fp2_mul(fp2_sqr(f, p), _line_eval(lam, V, Q, p), p)
        V = g1_dbl(V, p)
        if bit == "1":
            assert V is not None
            if V[0] == P[0]:
                # vertical line (V = -P): value in F_p, killed by the final exponentiation
                assert (V[1] + P[1]) % p == 0, "degenerate Miller loop (V == P)"
                V = None
            else:
                lam = (P[1] - V[1]) * pow(P[0] - V[0], -1, p) % p
                f = fp2_mul(f, _line_eval(lam, V, Q, p), p)
                V = g1_add(V, P, p)
    return f


def final_exp(f: Fp2, par: A1Params) -> Fp2:
    """f^((p^2-1)/n) = (conj(f)/f)^l  (a1_pairing 'Tate exponentiation' trick)."""
    p = par.p
    g = fp2_mul(fp2_conj(f, p), fp2_inv(f, p), p)
    return fp2_pow(g, par.l, p)


def pairing(P: G1Point, Q: G1Point, par: A1Params) -> Fp2:
    """pbc Element.Pair (pairing_apply: identity if either input is O)."""
    if P is None or Q is None:
        return GT_ONE
    return final_exp(miller_loop(P, Q, par), par)


# --------------------------------------------------------------------------
# serialisation   (libpbc element_to_bytes / element_from_bytes)
# --------------------------------------------------------------------------
def fp_to_bytes(x: int, par: A1Params) -> bytes:
    return int(x).to_bytes(par.coord_bytes, "big")


def g1_to_bytes(P: G1Point, par: A1Params) -> bytes:
    """x||y, big-endian fixed width.  O has no canonical PBC encoding
    (SURVEY.md 8(a) note); this engine and oracle use all-zero bytes."""
    if P is None:
        return bytes(2 * par.coord_bytes)
    return fp_to_bytes(P[0], par) + fp_to_bytes(P[1], par)


def g1_from_bytes(b: bytes, par: A1Params) -> G1Point:
    B = par.coord_bytes
    assert len(b) == 2 * B
    x = int.from_bytes(b[:B], "big") % par.p
    y = int.from_bytes(b[B:], "big") % par.p
    if x == 0 and y == 0:
        return None
    if not g1_on_curve((x, y), par.p):
        return None  # curve_from_bytes: not on curve -> O
    return (x, y)


def gt_to_bytes(a: Fp2, par: A1Params) -> bytes:
    return fp_to_bytes(a[0], par) + fp_to_bytes(a[1], par)


def gt_from_bytes(b: bytes, par: A1Params) -> Fp2:
    B = par.coord_bytes
    assert len(b) == 2 * B
    return (int.from_bytes(b[:B], "big") % par.p, int.from_bytes(b[B:], "big") % par.p)


def g1_string(P: G1Point) -> str:
    return "O" if P is None else "[%d, %d]" % P


def gt_string(a: Fp2) -> str:
    return "[%d, %d]" % a


# --------------------------------------------------------------------------
# plaintext encoding  (plaintext.go)
# --------------------------------------------------------------------------
DEGREE_BOUND = 128  # plaintext.go:11


def compute_encoding_table(base: int) -> Tuple[List[int], List[int]]:
    """plaintext.go:105-124 -> (degreeTable, degreeSumTable)."""
    deg = [1]
    sums = [1]
    s = 1
    for i in range(1, DEGREE_BOUND):
        r = base ** i
        s += r
        deg.append(r)
        sums.append(s)
    return deg, sums


def _degree(target: int, deg: List[int], sums: List[int], bound: int, balanced: bool) -> int:
    """plaintext.go:127-151."""
    if target == 1:
        return 0
    if balanced:
        for i in range(1, min(bound, DEGREE_BOUND - 1) + 1):
            if sums[i] >= target:
                return i
    else:
        for i in range(1, min(bound, DEGREE_BOUND - 1) + 1):
            if deg[i] > target:
                return i - 1
    return -1


def unbalanced_encode(target: int, base: int) -> List[int]:
    """plaintext.go:161-207; digits in {0,1,2}; returns coefficient list (len == Degree)."""
    deg, sums = compute_encoding_table(base)
    if target == 0:
        return [0]
    if target < 0:
        raise ValueError("Negative encoding not supported")
    coeffs = [0] * DEGREE_BOUND
    bound = None
    last = DEGREE_BOUND
    while True:
        index = _degree(target, deg, sums, last, False)
        last = index + 1
        if bound is None:
            bound = index + 1
        value = deg[index]
        if 2 * deg[index] <= target:
            value = 2 * deg[index]
            coeffs[index] = 2
        else:
            coeffs[index] = 1
        if value == target:
            return coeffs[: bound + 1]
        target -= value


def balanced_encode(target: int, base: int) -> List[int]:
    """plaintext.go:209-266; digits in {-1,0,1}."""
    deg, sums = compute_encoding_table(base)
    if target == 0:
        return [0]
    neg = target < 0
    if neg:
        target = -target
    coeffs = [0] * DEGREE_BOUND
    bound = None
    last = DEGREE_BOUND
    next_neg = False
    while True:
        index = _degree(target, deg, sums, last, True)
        last = index
        if bound is None:
            bound = index
        coeffs[index] = -1 if next_neg else 1
        if deg[index] == target:
            out = coeffs[: bound + 1]
            return [-c for c in out] if neg else out
        if deg[index] > target:
            next_neg = not next_neg
            target = deg[index] - target
        else:
            target = target - deg[index]


def rationalize(x: float, base: int, precision: float) -> Tuple[int, int]:
    """plaintext.go:269-312 (IEEE doubles, brute-force search)."""
    factor = math.floor(x)
    x = 1.0 + math.remainder(x, 1.0)
    if abs(x) > 1.0:
        x += 1.0
    if x >= 0.0:
        x -= float(int(x))
    elif x <= -0.0:
        x += float(int(x))
    num = 1.0
    pw = 1.0
    qmin = x - precision
    qmax = x + precision
    while True:
        denom = math.pow(float(base), pw)
        rat = num / denom
        if qmin <= rat <= qmax:
            while int(num) % base == 0:
                num = num / float(base)
                pw -= 1
            denom = math.pow(float(base), pw)
            return int(factor * denom + num), int(pw)
        if num + 1 >= denom:
            num = 1.0
            pw += 1
        num += 1


@dataclass
class PolyPlaintext:
    coefficients: List[int]
    scale_factor: int
    poly_base: int
    fp_scale_base: int

    @property
    def degree(self) -> int:  # number of coefficient slots (poly.go:13)
        return len(self.coefficients)

    def poly_eval(self) -> float:
        """plaintext.go:315-335 (Horner); exact rational -> float."""
        acc = 0
        for c in reversed(self.coefficients):
            acc = acc * self.poly_base + c
        if self.scale_factor:
            return acc / (self.fp_scale_base ** self.scale_factor)
        return float(acc)


def _encode_value(m: float, fp_scale_base: int, fp_precision: float) -> Tuple[int, int]:
    """shared prefix of NewPolyPlaintext / NewUnbalancedPlaintext (plaintext.go:41-52, 82-93)."""
    mf = float(m)
    if math.remainder(mf, 1.0) != 0.0:
        numerator, sf = rationalize(mf - math.floor(mf), fp_scale_base, fp_precision)
        m_int = int(mf)  # big.Float.Int truncates toward zero
        m_int = m_int * int(math.pow(float(fp_scale_base), float(sf))) + numerator
        return m_int, sf
    return int(mf), 0


# --------------------------------------------------------------------------
# keys
# --------------------------------------------------------------------------
@dataclass
class PublicKey:
    params: A1Params
    P: G1Point
    Q: G1Point
    msg_space: int
    deterministic: bool = True
    poly_base: int = 3
    fp_scale_base: int = 3
    fp_precision: float = 0.0001
    # decryption tables (gsbs.go keeps them in package globals; per key here)
    table_g1: dict = field(default_factory=dict, repr=False)
    table_gt: dict = field(default_factory=dict, repr=False)
    tables_computed: bool = False
    # Injected randomness: every newCryptoRandom(pk.N) call of the reference (bgn.go:567-574) that is
    # not given an explicit `r` below is one call of rand_source(), in the reference's sequential
    # program order (goroutine bodies in the textual order of their loops).  None: explicit r only.
    rand_source: Optional[Callable[[], int]] = field(default=None, repr=False)

    @property
    def n(self) -> int:
        return self.params.n

    # ---- plaintext constructors (plaintext.go:34-103)
    def new_poly_plaintext(self, m: float) -> PolyPlaintext:
        if m < 0:
            raise ValueError("negative encodings not implemented")
        v, sf = _encode_value(m, self.fp_scale_base, self.fp_precision)
        return PolyPlaintext(balanced_encode(v, self.poly_base), sf, self.poly_base, self.fp_scale_base)

    def new_unbalanced_plaintext(self, m: float) -> PolyPlaintext:
        v, sf = _encode_value(m, self.fp_scale_base, self.fp_precision)
        return PolyPlaintext(unbalanced_encode(v, self.poly_base), sf, self.poly_base, self.fp_scale_base)


@dataclass
class SecretKey:
    key: int  # q1
    R: int
    poly_base: int = 3


@dataclass
class Ciphertext:  # ciphertext.go:12-15
    C: object  # G1Point when not L2, Fp2 when L2
    L2: bool


@dataclass
class PolyCiphertext:  # ciphertext.go:26-31
    coefficients: List[Ciphertext]
    degree: int
    scale_factor: int
    L2: bool


def _rand_prime(rng: random.Random, bits: int) -> int:
    """crypto/rand.Prime behaviour: top two bits set, odd."""
    while True:
        c = rng.getrandbits(bits) | (3 << (bits - 2)) | 1
        if is_probable_prime(c):
            return c


def _rand_point(rng: random.Random, par: A1Params) -> G1Point:
    """pbc curve_random: random x with x^3+x a square, random sign, times cofactor."""
    p = par.p
    while True:
        x = rng.randrange(p)
        rhs = (x * x * x + x) % p
        y = pow(rhs, (p + 1) // 4, p)
        if y * y % p != rhs:
            continue
        if rng.getrandbits(1):
            y = (-y) % p
        return g1_mul(par.l, (x, y), p)


def keygen(key_bits: int, msg_space: int, poly_base: int = 3, fp_scale_base: int = 3,
           fp_precision: float = 0.0001, deterministic: bool = True, seed: int = 0) -> Tuple[PublicKey, SecretKey]:
    """bgn.go:65-138 with a seeded PRNG in place of crypto/rand."""
    assert key_bits >= 16 and key_bits % 2 == 0
    rng = random.Random(seed)
    q1 = _rand_prime(rng, key_bits // 2)
    q2 = _rand_prime(rng, key_bits // 2)
    while q2 == q1:
        q2 = _rand_prime(rng, key_bits // 2)
    if q1 < msg_space or q2 < msg_space:
        raise ValueError("Message space is greater than the group order!")
    n = q1 * q2
    par = a1_gen(n)
    p = par.p
    # findGenerator (bgn.go:170-192)
    while True:
        P = _rand_point(rng, par)
        if g1_mul(q1, P, p) is None or g1_mul(n, P, p) is not None:
            continue
        break
    P = g1_mul(4 * par.l, P, p)  # bgn.go:113
    R = rng.randrange(n)  # bgn.go:117
    Q = g1_mul(q2, g1_mul(R, P, p), p)  # bgn.go:118-119
    pk = PublicKey(par, P, Q, msg_space, deterministic, poly_base, fp_scale_base, fp_precision)
    return pk, SecretKey(q1, R, poly_base)


# --------------------------------------------------------------------------
# scalar scheme  (bgn.go)
# --------------------------------------------------------------------------
def encrypt_deterministic(pk: PublicKey, x: int) -> Ciphertext:
    """bgn.go:325-331."""
    return Ciphertext(g1_mul(x, pk.P, pk.params.p), False)


def encrypt_zero(pk: PublicKey) -> Ciphertext:
    return encrypt_deterministic(pk, 0)  # bgn.go:562-564


def encrypt_with_randomness(pk: PublicKey, x: int, r: int) -> Ciphertext:
    """bgn.go:340-353: C = P^x * Q^r."""
    p = pk.params.p
    return Ciphertext(g1_add(g1_mul(x, pk.P, p), g1_mul(r, pk.Q, p), p), False)


def _gt_div(a: Fp2, b: Fp2, p: int) -> Fp2:
    return fp2_mul(a, fp2_inv(b, p), p)


def make_l2(pk: PublicKey, ct: Ciphertext) -> Ciphertext:
    """bgn.go:316-321: e(C, P^1)."""
    return Ciphertext(pairing(ct.C, encrypt_deterministic(pk, 1).C, pk.params), True)


def _qq(pk: PublicKey) -> Fp2:
    return pairing(pk.Q, pk.Q, pk.params)


def _draw(pk: PublicKey, r: Optional[int] = None) -> int:
    """newCryptoRandom(pk.N) (bgn.go:567-574): the explicit r, else the next value of pk.rand_source."""
    if r is not None:
        return r
    if pk.rand_source is None:
        raise ValueError("non-deterministic operation without injected randomness (r= or pk.rand_source)")
    return pk.rand_source() % pk.params.n


def add(pk: PublicKey, a: Ciphertext, b: Ciphertext, r: Optional[int] = None) -> Ciphertext:
    """bgn.go:442-497.  r is the injected randomness used when !Deterministic."""
    p = pk.params.p
    ct1, ct2 = a, b
    if a.L2 and not b.L2:
        ct2 = make_l2(pk, b)
    if not a.L2 and b.L2:
        ct1 = make_l2(pk, a)
    if ct1.L2 and ct2.L2:
        res = fp2_mul(ct1.C, ct2.C, p)
        if not pk.deterministic:
            res = fp2_mul(res, fp2_pow(_qq(pk), _draw(pk, r), p), p)
        return Ciphertext(res, True)
    res = g1_add(ct1.C, ct2.C, p)
    if not pk.deterministic:
        res = g1_add(res, g1_mul(_draw(pk, r), pk.Q, p), p)
    return Ciphertext(res, ct1.L2)


def sub(pk: PublicKey, a: Ciphertext, b: Ciphertext, r: Optional[int] = None) -> Ciphertext:
    """bgn.go:375-433 (incl. the L2=false flag quirk of the non-deterministic L2 branch, bgn.go:411)."""
    p = pk.params.p
    ct1, ct2 = a, b
    if a.L2 and not b.L2:
        ct2 = make_l2(pk, b)
    if not a.L2 and b.L2:
        ct1 = make_l2(pk, a)
    if ct1.L2 and ct2.L2:
        res = _gt_div(ct1.C, ct2.C, p)
        if pk.deterministic:
            return Ciphertext(res, True)
        res = fp2_mul(res, fp2_pow(_qq(pk), _draw(pk, r), p), p)
        return Ciphertext(res, False)
    res = g1_add(ct1.C, g1_neg(ct2.C, p), p)
    if not pk.deterministic:
        res = g1_add(res, g1_mul(_draw(pk, r), pk.Q, p), p)
    return Ciphertext(res, ct1.L2)


def neg(pk: PublicKey, c: Ciphertext, r: Optional[int] = None) -> Ciphertext:
    """bgn.go:436-439."""
    return sub(pk, encrypt_zero(pk), c, r)


def mult(pk: PublicKey, a: Ciphertext, b: Ciphertext, r: Optional[int] = None) -> Ciphertext:
    """bgn.go:294-314."""
    p = pk.params.p
    res = pairing(a.C, b.C, pk.params)
    if not pk.deterministic:
        res = fp2_mul(res, fp2_pow(_qq(pk), _draw(pk, r), p), p)
    return Ciphertext(res, True)


def mult_const(pk: PublicKey, c: Ciphertext, k: int, r: Optional[int] = None) -> Ciphertext:
    """bgn.go:253-291."""
    p = pk.params.p
    if not c.L2:
        res = g1_mul(k, c.C, p)
        if not pk.deterministic:
            res = g1_add(res, g1_mul(_draw(pk, r), pk.Q, p), p)
        return Ciphertext(res, False)
    res = fp2_pow(c.C, k, p)
    if not pk.deterministic:
        res = fp2_mul(res, fp2_pow(_qq(pk), _draw(pk, r), p), p)
    return Ciphertext(res, True)


# ---- decryption (bgn.go:195-250, 357-372; gsbs.go) ------------------------
def _bsgs_bound(msg_space: int) -> int:
    return int(math.ceil(math.sqrt(float(msg_space))))


def setup_decryption(pk: PublicKey, sk: SecretKey) -> None:
    """bgn.go:195-201 + gsbs.go:17-51: table[gen^(j+1)] = j for j = 0..bound+1."""
    p = pk.params.p
    gen_g1 = g1_mul(sk.key, pk.P, p)
    gen_gt = fp2_pow(pairing(pk.P, pk.P, pk.params), sk.key, p)
    bound = _bsgs_bound(pk.msg_space) + 1
    pk.table_g1.clear()
    pk.table_gt.clear()
    aux_gt = gen_gt
    aux_g1 = gen_g1
    for j in range(bound + 1):
        pk.table_gt[aux_gt] = j
        aux_gt = fp2_mul(aux_gt, gen_gt, p)
    for j in range(bound + 1):
        pk.table_g1[aux_g1] = j
        aux_g1 = g1_add(aux_g1, gen_g1, p)
    pk.tables_computed = True


class DLError(Exception):
    pass


def get_dl(pk: PublicKey, csk, gsk, l2: bool) -> int:
    """gsbs.go:54-106."""
    if not pk.tables_computed:
        raise RuntimeError("DL tables not computed!")
    p = pk.params.p
    bound = _bsgs_bound(pk.msg_space)
    aux = csk
    if l2:
        gamma_inv = fp2_inv(fp2_pow(gsk, bound, p), p)
    else:
        gamma_inv = g1_neg(g1_mul(bound, gsk, p), p)
    for i in range(bound + 1):
        tbl = pk.table_gt if l2 else pk.table_g1
        if aux in tbl:
            return i * bound + tbl[aux] + 1
        aux = fp2_mul(aux, gamma_inv, p) if l2 else g1_add(aux, gamma_inv, p)
    raise DLError("cannot find discrete log; out of bounds")


def decrypt(pk: PublicKey, sk: SecretKey, ct: Ciphertext, _failed: bool = False) -> int:
    """bgn.go:218-250 (raises DLError where the reference returns an error)."""
    p = pk.params.p
    if ct.L2:
        gsk = fp2_pow(pairing(pk.P, pk.P, pk.params), sk.key, p)
        csk = fp2_pow(ct.C, sk.key, p)
        is_id = csk == GT_ONE
    else:
        gsk = g1_mul(sk.key, pk.P, p)
        csk = g1_mul(sk.key, ct.C, p)
        is_id = csk is None
    if is_id:  # recoverMessage, bgn.go:359-363
        return 0
    try:
        return get_dl(pk, csk, gsk, ct.L2)
    except DLError:
        if _failed:
            raise
        det = pk.deterministic
        pk.deterministic = True  # the sign flip itself needs no blinding for the value
        try:
            return -decrypt(pk, sk, neg(pk, ct), True)
        finally:
            pk.deterministic = det


def decrypt_fail_safe(pk: PublicKey, sk: SecretKey, ct: Ciphertext) -> int:
    try:
        return decrypt(pk, sk, ct)
    except DLError:
        return 0


# --------------------------------------------------------------------------
# polynomial ciphertexts (poly.go).  Randomness is injected: the per-coefficient encryption
# randomness as `rs` lists, every other newCryptoRandom draw through pk.rand_source (see _draw), taken
# in the reference's sequential program order: the goroutine bodies of poly.go:101-112 and
# poly.go:144-151 in the textual order of their loops (i and k DEscending).  With
# pk.deterministic = True and no rand_source the functions are the deterministic mode every reference
# test uses (bgn_test.go:13).
# --------------------------------------------------------------------------
def encrypt_poly(pk: PublicKey, pt: PolyPlaintext, rs: Optional[Sequence[int]] = None) -> PolyCiphertext:
    """poly.go:11-29; rs[i] is the randomness of coefficient i's Encrypt (drawn when rs is None).  A
    negative coefficient is Sub(encryptZero(), Encrypt(|c|)), whose own re-randomisation (non-deterministic
    keys, bgn.go:421-432) is drawn from pk.rand_source, or is 0 when no source is installed."""
    out = []
    for i, c in enumerate(pt.coefficients):
        r = rs[i] if rs is not None else _draw(pk)
        if c < 0:
            r2 = None
            if not pk.deterministic:
                r2 = _draw(pk) if pk.rand_source is not None else 0
            out.append(sub(pk, encrypt_zero(pk), encrypt_with_randomness(pk, -c, r), r2))
        else:
            out.append(encrypt_with_randomness(pk, c, r))
    return PolyCiphertext(out, pt.degree, pt.scale_factor, False)


def decrypt_poly(pk: PublicKey, sk: SecretKey, ct: PolyCiphertext) -> PolyPlaintext:
    """poly.go:32-42 (errors surface as None coefficients there; here DLError propagates)."""
    coeffs = [decrypt(pk, sk, c) for c in ct.coefficients[: ct.degree]]
    return PolyPlaintext(coeffs, ct.scale_factor, pk.poly_base, pk.fp_scale_base)


def neg_poly(pk: PublicKey, ct: PolyCiphertext) -> PolyCiphertext:
    """poly.go:45-55: Sub(encryptZero(), c_i) for i = degree-1 .. 0 (one draw each when !Deterministic).
    The coefficient flags are reported correctly (the reference's non-deterministic L2 Sub says L2=false,
    bgn.go:411)."""
    res: List[Optional[Ciphertext]] = [None] * ct.degree
    for i in range(ct.degree - 1, -1, -1):
        c = sub(pk, encrypt_zero(pk), ct.coefficients[i])
        res[i] = Ciphertext(c.C, ct.coefficients[i].L2)
    return PolyCiphertext(res, ct.degree, ct.scale_factor, ct.L2)


def mult_poly(pk: PublicKey, ct1: PolyCiphertext, ct2: PolyCiphertext) -> PolyCiphertext:
    """poly.go:123-156.  Non-deterministic keys: every coefficient pairing draws once in Mult
    (bgn.go:302-311) and once in the Add that folds it into its slot (bgn.go:466-474); the unused top
    slot stays makeL2(encryptZero()) = 1."""
    degree = ct1.degree + ct2.degree
    result = [make_l2(pk, encrypt_zero(pk)) for _ in range(degree)]
    for i in range(ct1.degree - 1, -1, -1):
        for k in range(ct2.degree - 1, -1, -1):
            coeff = mult(pk, ct1.coefficients[i], ct2.coefficients[k])
            result[i + k] = add(pk, result[i + k], coeff)
    return PolyCiphertext(result, degree, ct1.scale_factor + ct2.scale_factor, True)


def make_poly_l2(pk: PublicKey, ct: PolyCiphertext) -> PolyCiphertext:
    """poly.go:159-163: MultPoly(EncryptPoly(1.0), ct).  E(1.0)'s randomness is drawn from pk.rand_source;
    without a source it is 0 (the reference always draws one, so its MakePolyL2 is randomised even for
    Deterministic keys; r = 0 is the reproducible member of that family)."""
    one_pt = pk.new_poly_plaintext(1.0)
    one = encrypt_poly(pk, one_pt, None if pk.rand_source is not None else [0] * one_pt.degree)
    return mult_poly(pk, one, ct)


def mult_const_poly(pk: PublicKey, ct: PolyCiphertext, constant: float) -> PolyCiphertext:
    """poly.go:71-120.  Non-deterministic keys: every (coefficient, digit) pair -- zero digits included --
    draws once in MultConst (bgn.go:260-269, 279-288) and once in Add; a negative constant adds NegPoly's
    draws."""
    is_neg = constant < 0
    if is_neg:
        constant = -constant
    poly = pk.new_unbalanced_plaintext(constant)
    degree = ct.degree + poly.degree
    zero = encrypt_zero(pk)
    if ct.L2:
        zero = make_l2(pk, zero)
    result = [zero] * degree
    for i in range(ct.degree - 1, -1, -1):
        for k in range(poly.degree - 1, -1, -1):
            coeff = mult_const(pk, ct.coefficients[i], poly.coefficients[k])
            result[i + k] = add(pk, result[i + k], coeff)
    prod = PolyCiphertext(result, degree, ct.scale_factor + poly.scale_factor, ct.L2)
    return neg_poly(pk, prod) if is_neg else prod


def _align(pk: PublicKey, ct1: PolyCiphertext, ct2: PolyCiphertext):
    """poly.go:209-226."""
    if ct1.scale_factor > ct2.scale_factor:
        diff = ct1.scale_factor - ct2.scale_factor
        ct2 = mult_const_poly(pk, ct2, math.pow(float(pk.fp_scale_base), float(diff)))
        ct2.scale_factor = ct1.scale_factor
    elif ct2.scale_factor > ct1.scale_factor:
        return _align(pk, ct2, ct1)
    return ct1, ct2


def add_poly(pk: PublicKey, a: PolyCiphertext, b: PolyCiphertext) -> PolyCiphertext:
    """poly.go:171-207: common slots go through Add (i = degree-1 .. 0, one draw each when
    !Deterministic), the longer operand's tail is passed through untouched."""
    if a.L2 or b.L2:
        if not a.L2:
            return add_poly(pk, make_poly_l2(pk, a), b)
        if not b.L2:
            return add_poly(pk, a, make_poly_l2(pk, b))
    ct1, ct2 = _align(pk, a, b)
    degree = max(ct1.degree, ct2.degree)
    out: List[Optional[Ciphertext]] = [None] * degree
    for i in range(degree - 1, -1, -1):
        if i >= ct2.degree:
            out[i] = ct1.coefficients[i]
        elif i >= ct1.degree:
            out[i] = ct2.coefficients[i]
        else:
            out[i] = add(pk, ct1.coefficients[i], ct2.coefficients[i])
    return PolyCiphertext(out, degree, ct1.scale_factor, ct1.L2)


def sub_poly(pk: PublicKey, a: PolyCiphertext, b: PolyCiphertext) -> PolyCiphertext:
    return add_poly(pk, a, neg_poly(pk, b))  # poly.go:166-168


def eval_poly(pk: PublicKey, ct: PolyCiphertext) -> Ciphertext:
    """poly.go:58-68 (Horner in the exponent; MultConst then Add per coefficient, each drawing once when
    !Deterministic)."""
    acc = encrypt_deterministic(pk, 0)
    for c in reversed(ct.coefficients[: ct.degree]):
        acc = mult_const(pk, acc, pk.poly_base)
        acc = add(pk, acc, c)
    return acc


# --------------------------------------------------------------------------
# byte-level helpers used by the parity tests
# --------------------------------------------------------------------------
def ct_bytes(pk: PublicKey, ct: Ciphertext) -> bytes:
    """pbc Element.Bytes() of the ciphertext element (ciphertext.go:79)."""
    return gt_to_bytes(ct.C, pk.params) if ct.L2 else g1_to_bytes(ct.C, pk.params)


def poly_ct_bytes(pk: PublicKey, ct: PolyCiphertext) -> bytes:
    return b"".join(ct_bytes(pk, c) for c in ct.coefficients)


# --------------------------------------------------------------------------
# canonical (unshared) work model, SURVEY.md 8(d)
# --------------------------------------------------------------------------
def canonical_modmuls_per_pairing(par: A1Params) -> int:
    n, l = par.n, par.l
    miller = 23 * n.bit_length() + 18 * bin(n).count("1") - 70
    fexp = 4 + 3 + 2 * l.bit_length() + 3 * bin(l).count("1")
    return miller + fexp + 1


def products_per_modmul(L: int) -> int:
    return 2 * L * L + L
