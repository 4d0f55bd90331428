"""Prototype (integers mod p): two consecutive doubling steps of the Miller loop as ONE step with the function
g = l_T^2 l_2T / v_2T^2, div g = 4(T) + (-4T) - 5(O), g = a2 x^2 + a1 x + a0 + (x + beta) y  (a reduction of the product
modulo the curve with sympy gave the coefficients).  Checks the pairing value against the oracle."""
import json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import bgn_oracle as O
from bgn_b200.workmodel import naf_digits


def quad_coeffs(x1, y1, p):
    l1 = (3 * x1 * x1 + 1) * pow(2 * y1, -1, p) % p
    x2 = (l1 * l1 - 2 * x1) % p
    y2 = (l1 * (x1 - x2) - y1) % p
    l2 = (3 * x2 * x2 + 1) * pow(2 * y2, -1, p) % p
    a2 = -(2 * l1 + l2) % p
    a1 = (-3 * l1**3 - 2 * l1**2 * l2 + 7 * l1 * x1 + 2 * l2 * x1 - y1) % p
    a0 = (-3 * l1**5 - 2 * l1**4 * l2 + 15 * l1**3 * x1 + 8 * l1**2 * l2 * x1 - l1**2 * y1 - 2 * l1 * l2 * y1
          - 20 * l1 * x1**2 - 2 * l1 - 4 * l2 * x1**2 - l2 + 4 * x1 * y1) % p
    beta = (3 * l1 * l1 + 2 * l1 * l2 - 4 * x1) % p
    x4 = (l2 * l2 - 2 * x2) % p
    y4 = (l2 * (x2 - x4) - y2) % p
    return a2, a1, a0, beta, (x4, y4)


def miller_quad(P, Q, par):
    p, n = par.p, par.n
    xE, yE = Q
    naf = naf_digits(n)
    N = len(naf)
    f = (1, 0)
    T = P
    idx = 1
    nq = 0
    while idx < N:
        d = naf[idx]
        add = d != 0 and idx != N - 1
        nxt_add = idx + 1 < N and naf[idx + 1] != 0 and idx + 1 != N - 1
        if not add and idx + 1 < N and not nxt_add:
            # two doubling steps at once: f <- (f^2)^2 g(phi(Q))
            a2, a1, a0, beta, T4 = quad_coeffs(T[0], T[1], p)
            if idx != 1:
                f = O.fp2_sqr(f, p)
            f = O.fp2_sqr(f, p)
            re = (a2 * xE * xE - a1 * xE + a0) % p      # a(-xE)
            im = yE * (beta - xE) % p                    # (x + beta) y at (-xE, i yE)
            f = O.fp2_mul(f, (re, im), p)
            T = T4
            idx += 2
            nq += 1
            continue
        if idx != 1:
            f = O.fp2_sqr(f, p)
        lam = (3 * T[0] * T[0] + 1) * pow(2 * T[1], -1, p) % p
        f = O.fp2_mul(f, O._line_eval(lam, T, Q, p), p)
        T = O.g1_dbl(T, p)
        if add:
            A = P if d > 0 else O.g1_neg(P, p)
            lam = (A[1] - T[1]) * pow(A[0] - T[0], -1, p) % p
            f = O.fp2_mul(f, O._line_eval(lam, T, Q, p), p)
            T = O.g1_add(T, A, p)
        idx += 1
    return f, nq


def main():
    for kb in (64, 128):
        g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "kb%d.json" % kb)))
        par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
        P = O.g1_from_bytes(bytes.fromhex(g["P"]), par)
        rng = random.Random(kb)
        for _ in range(3):
            A = O.g1_mul(rng.randrange(1, par.n), P, par.p)
            B = O.g1_mul(rng.randrange(1, par.n), P, par.p)
            f, nq = miller_quad(A, B, par)
            assert O.final_exp(f, par) == O.pairing(A, B, par), kb
        print(kb, "ok", nq, "quadruplings of", len(naf_digits(par.n)) - 1, "steps")


if __name__ == "__main__":
    main()
