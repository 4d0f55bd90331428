"""Work model of the kernels: how many F_p Montgomery products (modmuls) the
schedules in bgn_b200/csrc actually execute, and how many 32x32->64 products
(IMAD.WIDE issue slots) that is.  bench.py's roofline numbers come from here;
tests/test_hostsim.py checks these formulas against a counter in the CPU
simulation of the device code, and DESIGN.md states them.

Units:  1 modmul on L limbs = 2*L*L + L products.
"""
from __future__ import annotations

from typing import List


def products_per_modmul(L: int) -> int:
    return 2 * L * L + L


def products_per_sqr(L: int) -> int:
    """the dedicated squaring (arith.cuh: Fp::sqr): L (L + 1) / 2 for the square + L^2 + L for its reduction"""
    return L * (L + 1) // 2 + L * L + L


FUSED_SQR = False  # fused.cuh BGN_FUSED_SQR: measured neutral in k_miller, not shipped


def fused_sqr(L: int) -> bool:
    """fused.cuh MF::SQR: whether the fused Miller routines square with Fp::sqr (an option, off: see there)"""
    return FUSED_SQR and L <= 17


def pick_limbs(p: int) -> int:
    for L in (3, 5, 9, 17, 33):
        if 32 * L >= p.bit_length() + 8:
            return L
    raise ValueError("field too large")


def naf_digits(n: int) -> List[int]:
    d = []
    while n:
        z = 0
        if n & 1:
            z = 2 - (n & 3)
            n -= z
        d.append(z)
        n >>= 1
    return d[::-1]


def fermat_inv_modmuls(p: int, L: int) -> int:
    """F<L>::inv: left-to-right binary exponentiation by p-2 (field.cuh)."""
    e = p - 2
    return (e.bit_length() - 1) + (bin(e).count("1") - 1)


GCD_INV_MODMULS = 2  # F<L>::inv_gcd: the binary GCD itself multiplies nothing; R^3 and the form fix are 2 products


def final_exp_modmuls(p: int, l: int, L: int, nslots: int = 1, team: int = 1) -> int:
    """MillerTeam::finalize for one unit: per slot conj(f)^2 and N(f) (3), g = conj(f)^2 / N (2) and
    g^l; one inversion (binary GCD) per THREAD that owns a slot (thread t owns slots t and t + team),
    a thread with two slots adds 3 products (Montgomery's trick)."""
    pow_l = 2 * (l.bit_length() - 1) + 3 * (bin(l).count("1") - 1)
    owners = min(team, nslots)
    return nslots * (5 + pow_l) + owners * GCD_INV_MODMULS + (nslots - owners) * 3


def line_lazy(L: int) -> bool:
    """pairing.cuh BGN_LINE_LAZY: phase B uses the lazy-reduction line_mul up to 17 limbs."""
    return L <= 17


def miller_unit_squarings(p: int, n: int, l: int, dM: int, dE: int) -> int:
    """of miller_unit_modmuls, how many are dedicated squarings: 6 of dbl_line's 12 (XX, YY, ZZ, ZZ^2,
    (2YY)^2, M^2), 3 of madd_line's 13 (ZZ, (2H)^2, r^2), 2 of fe_prepare's 3 per output slot"""
    L = pick_limbs(p)
    if not fused_sqr(L):
        return 0
    naf = naf_digits(n)
    D = len(naf) - 1
    A = sum(1 for i in range(1, len(naf) - 1) if naf[i] != 0)
    return D * dM * 6 + A * dM * 3 + (dM + dE - 1) * 2


EVAL_NORM = True  # pairing.cuh BGN_EVAL_NORM: evaluation points normalised to (x / y, 1 / y) before the loop


def line_products(L: int, eval_norm: bool = None) -> int:
    """32x32->64 products of ONE line folded into an accumulator in the team kernels (fused.cuh).
    Plain form: two evaluation products and an F_p^2 product -- 5 (2L^2 + L), or with lazy reduction
    (up to 17 limbs) 2 (2L^2 + L) + 3 L^2 + 2 (L^2 + L).  At a normalised evaluation point (round 2) the
    evaluation is one dot product of 3L^2 + L: 9L^2 + 4L, or 8L^2 + 3L with lazy reduction."""
    eval_norm = EVAL_NORM if eval_norm is None else eval_norm
    full = products_per_modmul(L)
    ev = (3 * L * L + L) if eval_norm else 2 * full
    return ev + (3 * L * L + 2 * (L * L + L) if line_lazy(L) else 3 * full)


PARABOLA = True  # pairing.cuh BGN_PARABOLA: a NAF digit != 0 is one parabola step


def parabola_on(L: int, eval_norm: bool = None, parabola: bool = None, dE: int = 11) -> bool:
    """api.cu run_miller: the team kernel merges doubling and addition where a Miller point serves >= 3 evaluation points"""
    eval_norm = EVAL_NORM if eval_norm is None else eval_norm
    parabola = PARABOLA if parabola is None else parabola
    return bool(parabola and eval_norm and dE >= 3)


def para_products(L: int) -> int:
    """one parabola folded into an accumulator (fused.cuh: para_mul_lazy / para_mul): a three-term dot product
    (4L^2 + L) and an F_p^2 product (3 L^2 + 2 (L^2 + L) lazily reduced, else 3 (2L^2 + L))"""
    full = products_per_modmul(L)
    return (4 * L * L + L) + (3 * L * L + 2 * (L * L + L) if line_lazy(L) else 3 * full)


DADD_MULS, DADD_SQRS, DADD_DOT2 = 28, 0, 1  # fused.cuh MF::dadd_para: 22 products + 6 squarings (as products) + one dot product


def miller_unit_products(p: int, n: int, l: int, dM: int, dE: int, eval_norm: bool = None, parabola: bool = None) -> int:
    """32x32->64 products one unit of k_miller executes.  Every F_p product is a fused
    multiply-and-reduce of 2L^2 + L, except the lines (line_products), the squarings where the
    dedicated routine is used (L (L + 1) / 2 + L^2 + L), and -- with normalised evaluation points --
    one inversion (its two products of glue) and one product per evaluation point before the loop.
    With the parabola step a NAF digit != 0 costs dM dadd_para (28 products and a dot product) and dM dE
    para_products instead of a doubling and an addition step with dM dE lines each; the x^2 / y of the
    evaluation points is one more product each before the loop."""
    eval_norm = EVAL_NORM if eval_norm is None else eval_norm
    L = pick_limbs(p)
    full, sq = products_per_modmul(L), products_per_sqr(L)
    naf = naf_digits(n)
    D = len(naf) - 1
    A = sum(1 for i in range(1, len(naf) - 1) if naf[i] != 0)
    nslots = dM + dE - 1
    fe = final_exp_modmuls(p, l, L, nslots, dE)
    prep = dE * (GCD_INV_MODMULS + 1) if eval_norm else 0
    if parabola_on(L, eval_norm, parabola, dE):
        prep += dE   # x^2 / y of every evaluation point
    nsq = miller_unit_squarings(p, n, l, dM, dE)   # dedicated squarings inside the fused routines (FUSED_SQR)
    if parabola_on(L, eval_norm, parabola, dE):
        plain = (D - A) * dM * 12 + (D - 1) * nslots * 2 + fe + prep   # fused products outside lines / parabolas
        return ((plain - nsq + A * dM * DADD_MULS) * full + (nsq + A * dM * DADD_SQRS) * sq + A * dM * DADD_DOT2 * (3 * L * L + L)
                + (D - A) * dM * dE * line_products(L, eval_norm) + A * dM * dE * para_products(L))
    mm = miller_unit_modmuls(p, n, l, dM, dE)
    lines = (D + A) * dM * dE
    return (mm - 5 * lines - nsq + prep) * full + lines * line_products(L, eval_norm) + nsq * sq


def miller_unit_lines(n: int, dM: int, dE: int, L: int = 17) -> int:
    """lines folded into accumulators by one unit of k_miller: one per (Miller point, evaluation point) and step;
    with the parabola step only the steps whose digit is zero have lines"""
    naf = naf_digits(n)
    D = len(naf) - 1
    A = sum(1 for i in range(1, len(naf) - 1) if naf[i] != 0)
    return ((D - A) if parabola_on(L, dE=dE) else (D + A)) * dM * dE


def miller_unit_parabolas(n: int, dM: int, dE: int, L: int = 17) -> int:
    naf = naf_digits(n)
    A = sum(1 for i in range(1, len(naf) - 1) if naf[i] != 0)
    return A * dM * dE if parabola_on(L, dE=dE) else 0


def miller_unit_modmuls(p: int, n: int, l: int, dM: int, dE: int) -> int:
    """One unit of k_miller (dM Miller points x dE evaluation points, dM <= dE; all points finite):
    dbl_line 12, madd_line 13, line_mul 5, sqr2 2 per output slot."""
    L = pick_limbs(p)
    naf = naf_digits(n)
    D = len(naf) - 1
    A = sum(1 for i in range(1, len(naf) - 1) if naf[i] != 0)
    nslots = dM + dE - 1
    per_dbl = dM * 12 + dM * dE * 5
    per_add = dM * 13 + dM * dE * 5
    return D * per_dbl + (D - 1) * nslots * 2 + A * per_add + final_exp_modmuls(p, l, L, nslots, dE)


def miller_fixed_products(p: int, n: int, l: int, parabola: bool = None) -> int:
    """32x32->64 products of one k_miller_fixed thread (pairing with the recorded, normalised table of a
    fixed first argument): per doubling step one line (fused.cuh: line_mul_lazy_f up to 17 limbs -- one
    evaluation product, three double-width products, two reductions -- else line_mul_f, 4 products); per
    doubling-and-addition step one parabola (para_mul_lazy_f: two evaluation products; without the parabola
    step: two lines); on every step after the first one sqr2; then the final exponentiation of one slot."""
    parabola = PARABOLA if parabola is None else parabola
    L = pick_limbs(p)
    naf = naf_digits(n)
    D = len(naf) - 1
    A = sum(1 for i in range(1, len(naf) - 1) if naf[i] != 0)
    full = products_per_modmul(L)
    fp2 = (3 * L * L + 2 * (L * L + L)) if line_lazy(L) else 3 * full
    nsq = 2 if fused_sqr(L) else 0  # fe_prepare: f0^2, f1^2
    folds = (D - A) * (full + fp2) + A * (2 * full + fp2) if parabola else (D + A) * (full + fp2)
    return folds + ((D - 1) * 2 + final_exp_modmuls(p, l, L, 1, 1) - nsq) * full + nsq * products_per_sqr(L)


def miller_fixed_pair_counts(p: int, n: int, l: int, parabola: bool = None):
    """k_miller_fixed_pair (pairlane.cuh), BOTH lanes of one pairing together:
    -> (dot products, plain Montgomery products incl. the inversion's glue).
    Per table entry a dot-product half per lane; a line's real part costs one product and the lanes evaluate
    two consecutive line entries at once; a parabola's real part is one product per lane (csn xB^2 | c1n xB;
    xB^2 is one more product per lane before the loop); a squaring half per lane on every step after the
    first; final exponentiation: f0^2 | f1^2, f0 f1 and the scaling in both lanes, the verified inversion
    (3 products) in both lanes, then g^l."""
    parabola = PARABOLA if parabola is None else parabola
    naf = naf_digits(n)
    N = len(naf)
    D = N - 1
    A = sum(1 for i in range(1, N - 1) if naf[i] != 0)
    lb, lw = l.bit_length() - 1, bin(l).count("1") - 1
    if not parabola:
        dots = 2 * (D + A) + 2 * lw
        muls = 2 * ((D + A + 1) // 2) + 2 * (D - 1) + 2 * (1 + 1 + 1 + 3) + 2 * lb
        return dots, muls
    is_dadd = lambda i: naf[i] != 0 and i != N - 1
    rounds, have = 0, False
    for idx in range(1, N):
        if is_dadd(idx):
            rounds += 1
            continue
        if not have:
            rounds += 1
            have = idx + 1 < N and not is_dadd(idx + 1)
        else:
            have = False
    dots = 2 * D + 2 * lw
    muls = 2 * rounds + 2 + 2 * (D - 1) + 2 * (1 + 1 + 1 + 3) + 2 * lb
    return dots, muls


def miller_fixed_pair_products(p: int, n: int, l: int) -> int:
    """32x32->64 products both lanes of one k_miller_fixed_pair pairing execute: a dot product is
    3L^2 + L (arith.cuh: Fp::dot2), a Montgomery product 2L^2 + L."""
    L = pick_limbs(p)
    dots, muls = miller_fixed_pair_counts(p, n, l)
    return dots * (3 * L * L + L) + muls * products_per_modmul(L)


def duo_loop_default(L: int) -> int:
    """api.cu: pair_duo_variant -- the shipped loop shape of k_pair_duo's products"""
    return 2 if L == 17 else (0 if L < 17 else 4)


def pair_duo_counts(p: int, n: int, l: int, loop: int = None):
    """k_pair_duo (pairwarp.cuh), BOTH warps' lanes of one pairing together:
    -> (fused Montgomery products, double-width products, separate reductions, dedicated squarings).
    Doubling step: X 9 (6 of them squarings up to 17 limbs), F 4 (line at B) + 2 (sqr2, not on the first
    step); addition step: X 11 (3 squarings), F 3; every step one F_p^2 product f * l: 3 double-width
    products + 2 reductions up to 17 limbs (lazy), 3 fused products beyond; final exponentiation of one slot
    as k_miller_fixed (its two squarings dedicated only with unrolled products, loop == 0)."""
    L = pick_limbs(p)
    loop = duo_loop_default(L) if loop is None else loop
    naf = naf_digits(n)
    D = len(naf) - 1
    A = sum(1 for i in range(1, len(naf) - 1) if naf[i] != 0)
    mm = D * (9 + 4) + (D - 1) * 2 + A * (11 + 3) + final_exp_modmuls(p, l, L, 1, 1) + 1  # + the inversion's check
    nsq = (D * 6 + A * 3 if L <= 17 else 0) + (2 if (loop == 0 and fused_sqr(L)) else 0)
    if line_lazy(L):
        return mm - nsq, 3 * (D + A), 2 * (D + A), nsq
    return mm - nsq + 3 * (D + A), 0, 0, nsq


def pair_duo_products(p: int, n: int, l: int, loop: int = None) -> int:
    """32x32->64 products one k_pair_duo pairing executes (both warps)"""
    L = pick_limbs(p)
    mm, mw, rd, nsq = pair_duo_counts(p, n, l, loop)
    return mm * products_per_modmul(L) + mw * L * L + rd * (L * L + L) + nsq * products_per_sqr(L)


def canonical_pairing_modmuls(n: int, l: int) -> int:
    """PBC-like unshared schedule, one full pairing."""
    return 23 * n.bit_length() + 18 * bin(n).count("1") - 70 + 4 + 3 + 2 * l.bit_length() + 3 * bin(l).count("1") + 1


def gt_pow_modmuls(e: int) -> int:
    """GT<L>::pow_fixed: left-to-right binary, sqr2 = 2 products, mul2 = 3 (pairing.cuh)."""
    if e == 0:
        return 0
    return 2 * (e.bit_length() - 1) + 3 * (bin(e).count("1") - 1)


def dec_lucas_modmuls(e: int) -> int:
    """k_dec_lucas (lucas.cuh), per ciphertext: one squaring and one product per exponent bit below
    the top one (one each per lane of the pair), the set-up both lanes run (norm check 2, V_2 1) and
    the two products of the sign test."""
    if e == 0:
        return 0
    return 2 * (e.bit_length() - 1) + 2 * 3 + 2


def madd_products(L: int) -> int:
    """one complete mixed addition (curve.cuh: G::madd): 8 products and 3 dedicated squarings"""
    return 8 * products_per_modmul(L) + 3 * products_per_sqr(L)


def affadd_products(L: int) -> int:
    """one affine addition with a shared inversion (k_g1_affadd): 5 products and 1 squaring (the slope's)"""
    return 5 * products_per_modmul(L) + products_per_sqr(L)


ENC_EDWARDS = True  # api.cu enc_edwards: Encrypt's tables hold twisted Edwards points (curve.cuh: Ed)


def enc_table_bytes(rbytes: int, window_bits: int, L: int, edwards: bool = None) -> int:
    """bytes of the fixed-base table of Q with windows of window_bits (api.cu: tabQw_bytes): two field
    elements per entry, three (u, v, u v) in Edwards form"""
    edwards = ENC_EDWARDS if edwards is None else edwards
    return ((8 * rbytes + window_bits - 1) // window_bits) * ((1 << window_bits) - 1) * (3 if edwards else 2) * L * 4


def enc_window_auto(rbytes: int, L: int, table_max: int = 6 << 30) -> int:
    """the window width `enc_window = 0` picks (api.cu: enc_window_auto): widest of 20 / 18 / 16 bits within table_max"""
    for bits in (20, 18):
        if enc_table_bytes(rbytes, bits, L) <= table_max:
            return bits
    return 16


def encrypt_additions(rbytes: int, window_bits: int, p_x_nonzero: float = 2.0 / 3.0) -> float:
    """k_encrypt, EXPECTED table points per coefficient for uniform r: one per non-zero window digit of r,
    plus one for a non-zero plaintext digit"""
    windows = (8 * rbytes + window_bits - 1) // window_bits
    return windows * (1.0 - 2.0 ** -window_bits) + p_x_nonzero


def encrypt_products(n: int, rbytes: int, window_bits: int, L: int, p_x_nonzero: float = 2.0 / 3.0,
                     edwards: bool = None) -> float:
    """k_encrypt, EXPECTED 32x32->64 products per coefficient.  The first table point is a copy.  Edwards
    form (the default): 8 products per further point and the conversion to Jacobian coordinates (5 products
    and a squaring); Weierstrass tables: one complete mixed addition (8 products, 3 squarings) per further
    point.  Jacobian -> affine (k_normalize) is accounted separately."""
    edwards = ENC_EDWARDS if edwards is None else edwards
    adds = max(0.0, encrypt_additions(rbytes, window_bits, p_x_nonzero) - 1.0)
    if edwards:
        return (8 * adds + 5) * products_per_modmul(L) + products_per_sqr(L)
    return adds * madd_products(L)


def encrypt_modmuls(n: int, rbytes: int, window_bits: int = 8, p_x_nonzero: float = 2.0 / 3.0, edwards: bool = None) -> float:
    """the same as a count of F_p products (squarings counted as products)"""
    edwards = ENC_EDWARDS if edwards is None else edwards
    adds = max(0.0, encrypt_additions(rbytes, window_bits, p_x_nonzero) - 1.0)
    return 8 * adds + 6 if edwards else 11 * adds
