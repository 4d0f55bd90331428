"""Host-side plaintext encoding: the mirror of the reference's plaintext.go.

north_star keeps "the poly.go/plaintext.go encoding on the host"; nothing here
touches the GPU.  Names follow the reference (NewPolyPlaintext,
NewUnbalancedPlaintext, PolyEval); IEEE-double behaviour of `rationalize`
(plaintext.go:269-312) is reproduced with Python floats, which are the same
binary64 values Go's float64 holds.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

DEGREE_BOUND = 128  # plaintext.go:11


@dataclass
class PolyEncodingParams:  # bgn.go:20-24
    PolyBase: int = 3
    FPScaleBase: int = 3
    FPPrecision: float = 0.0001


class EncodingTable:
    """computeEncodingTable (plaintext.go:105-124): powers of the base and their running sums."""

    def __init__(self, base: int):
        self.base = base
        self.degree = [base ** i for i in range(DEGREE_BOUND)]
        self.degree_sum = []
        acc = 0
        for d in self.degree:
            acc += d
            self.degree_sum.append(acc)

    def closest(self, target: int, bound: int, balanced: bool) -> int:
        """degree() (plaintext.go:127-151)."""
        if target == 1:
            return 0
        hi = min(bound, DEGREE_BOUND - 1)
        if balanced:
            for i in range(1, hi + 1):
                if self.degree_sum[i] >= target:
                    return i
        else:
            for i in range(1, hi + 1):
                if self.degree[i] > target:
                    return i - 1
        return -1


def unbalancedEncode(target: int, tab: EncodingTable) -> List[int]:
    """plaintext.go:161-207: greedy digits in {0,1,2}; slot count = highest index + 2."""
    if target == 0:
        return [0]
    if target < 0:
        raise ValueError("Negative encoding not supported")
    digits = [0] * DEGREE_BOUND
    top = None
    last = DEGREE_BOUND
    while True:
        idx = tab.closest(target, last, False)
        last = idx + 1
        if top is None:
            top = idx + 1
        step = tab.degree[idx]
        if 2 * step <= target:
            step *= 2
            digits[idx] = 2
        else:
            digits[idx] = 1
        if step == target:
            return digits[: top + 1]
        target -= step


def balancedEncode(target: int, tab: EncodingTable) -> List[int]:
    """plaintext.go:209-266: digits in {-1,0,1}."""
    if target == 0:
        return [0]
    flip = target < 0
    target = abs(target)
    digits = [0] * DEGREE_BOUND
    top = None
    last = DEGREE_BOUND
    minus = False
    while True:
        idx = tab.closest(target, last, True)
        last = idx
        if top is None:
            top = idx
        digits[idx] = -1 if minus else 1
        pw = tab.degree[idx]
        if pw == target:
            res = digits[: top + 1]
            return [-d for d in res] if flip else res
        if pw > target:
            minus = not minus
            target = pw - target
        else:
            target -= pw


def rationalize(x: float, base: int, precision: float) -> Tuple[int, int]:
    """plaintext.go:269-312 -> (numerator, scaleFactor)."""
    whole = math.floor(x)
    x = 1.0 + math.remainder(x, 1.0)
    if abs(x) > 1.0:
        x += 1.0
    if x >= 0.0:
        x -= float(int(x))
    elif x <= -0.0:
        x += float(int(x))
    lo, hi = x - precision, x + precision
    num, pw = 1.0, 1.0
    while True:
        den = math.pow(float(base), pw)
        if lo <= num / den <= hi:
            while int(num) % base == 0:
                num /= float(base)
                pw -= 1
            den = math.pow(float(base), pw)
            return int(whole * den + num), int(pw)
        if num + 1 >= den:
            num = 1.0
            pw += 1
        num += 1


@dataclass
class PolyPlaintext:  # plaintext.go:14-19
    Coefficients: List[int]
    Degree: int  # number of coefficient slots
    ScaleFactor: int
    params: PolyEncodingParams

    def PolyEval(self) -> float:
        """plaintext.go:315-335 (Horner)."""
        acc = 0
        for c in reversed(self.Coefficients[: self.Degree]):
            acc = acc * self.params.PolyBase + c
        if self.ScaleFactor != 0:
            return acc / (self.params.FPScaleBase ** self.ScaleFactor)
        return float(acc)


def _to_scaled_int(m: float, prm: PolyEncodingParams) -> Tuple[int, int]:
    """common prefix of plaintext.go:41-52 and 82-93."""
    mf = float(m)
    if math.remainder(mf, 1.0) != 0.0:
        numerator, sf = rationalize(mf - math.floor(mf), prm.FPScaleBase, prm.FPPrecision)
        return int(mf) * int(math.pow(float(prm.FPScaleBase), float(sf))) + numerator, sf
    return int(mf), 0


def NewPolyPlaintext(m: float, prm: PolyEncodingParams, tab: EncodingTable) -> PolyPlaintext:
    """plaintext.go:67-103."""
    if m < 0:
        raise ValueError("negative encodings not implemented")
    v, sf = _to_scaled_int(m, prm)
    c = balancedEncode(v, tab)
    return PolyPlaintext(c, len(c), sf, prm)


def NewUnbalancedPlaintext(m: float, prm: PolyEncodingParams, tab: EncodingTable) -> PolyPlaintext:
    """plaintext.go:34-63."""
    v, sf = _to_scaled_int(m, prm)
    c = unbalancedEncode(v, tab)
    return PolyPlaintext(c, len(c), sf, prm)
