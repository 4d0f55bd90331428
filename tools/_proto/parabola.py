"""Prototype (integers mod p) of the Miller loop with ELM parabola steps in Jacobian coordinates:
a NAF digit != 0 makes ONE step T <- (T + sP) + T with the parabola through T, T, sP, -(2T + sP) instead of a
doubling step followed by an addition step.  Checks the pairing value against the oracle."""
import json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import bgn_oracle as O
from bgn_b200.workmodel import naf_digits


def miller_parabola(P, Q, par, count=None):
    p, n = par.p, par.n
    xP, yP0 = P
    xE, yE = Q
    v = pow(yE, -1, p); u = xE * v % p; w = xE * u % p      # evaluation constants: 1/y, x/y, x^2/y
    X, Y, Z = xP, yP0, 1
    f = (1, 0)
    naf = naf_digits(n)          # device order: naf[0] is the top digit
    L = len(naf)
    nm = 0
    for idx in range(1, L):
        d = naf[idx]
        if idx != 1:
            f = O.fp2_sqr(f, p)
        if d != 0 and idx != L - 1:
            yP = yP0 if d > 0 else (-yP0) % p
            # R = T + sP (mixed), T rescaled to R's Z: (U, W, ZR)
            ZZ = Z * Z % p
            H = (xP * ZZ - X) % p
            N1 = (yP * Z % p * ZZ - Y) % p
            HH = H * H % p
            H3 = H * HH % p
            U = X * HH % p
            W = Y * H3 % p
            XR = (N1 * N1 - H3 - 2 * U) % p
            YR = (N1 * (U - XR) - W) % p
            ZR = Z * H % p
            # S = R + T (co-Z)
            H2 = (XR - U) % p
            N2 = (YR - W) % p
            A2 = H2 * H2 % p
            B2 = U * A2 % p
            C2 = XR * A2 % p
            XS = (N2 * N2 - B2 - C2) % p
            YS = (N2 * (B2 - XS) - W * (C2 - B2)) % p
            ZS = ZR * H2 % p
            # parabola g Dn^2 = s x^2 + c1 x + c0 + im y  at (-xE, i yE), Dn = ZR ZS
            An = (XR * H2 + N1 * N2) % p
            Sn = (N1 * H2 + N2) % p
            Dn = ZR * ZS % p
            s = Dn * Dn % p
            c1 = An * Dn % p                       # coefficient of x in g Dn^2 (x -> -xE flips its sign)
            c0 = H2 * ((Sn * W - U * ((U * H2 + An) % p)) % p) % p
            im = (-Sn * Dn % p) * ZR % p           # coefficient of y
            re = (s * w - c1 * u + c0 * v) % p     # (s xE^2 - c1 xE + c0) / yE
            f = O.fp2_mul(f, (re, im), p)          # im * yE / yE
            X, Y, Z = XS, YS, ZS
            nm += 31
        else:
            # doubling with tangent (device convention): cR = M X - 2YY, aR = M ZZ, bI = Z3 ZZ
            XX = X * X % p; YY = Y * Y % p; ZZ = Z * Z % p
            M = (3 * XX + ZZ * ZZ) % p
            Z3 = 2 * Y * Z % p
            cR = (M * X - 2 * YY) % p; aR = M * ZZ % p; bI = Z3 * ZZ % p
            Sx = 4 * X * YY % p
            X3 = (M * M - 2 * Sx) % p
            Y3 = (M * (Sx - X3) - 8 * YY * YY) % p
            f = O.fp2_mul(f, ((cR * v + aR * u) % p, bI), p)
            X, Y, Z = X3, Y3, Z3
    return f


def main():
    for kb in (64, 128, 512):
        g = json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "tests", "golden", "kb%d.json" % kb)))
        par = O.A1Params(int(g["p"], 16), int(g["n"], 16), g["l"])
        P = O.g1_from_bytes(bytes.fromhex(g["P"]), par)
        rng = random.Random(kb)
        for _ in range(3):
            A = O.g1_mul(rng.randrange(1, par.n), P, par.p)
            B = O.g1_mul(rng.randrange(1, par.n), P, par.p)
            got = O.final_exp(miller_parabola(A, B, par), par)
            assert got == O.pairing(A, B, par), kb
        print(kb, "ok")


if __name__ == "__main__":
    main()
