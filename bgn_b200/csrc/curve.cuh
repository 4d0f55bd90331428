// curve.cuh -- G1 = E(F_p)[n], E: y^2 = x^3 + x, in Jacobian coordinates
// (x = X/Z^2, y = Y/Z^3; O is Z == 0), plus the tangent / chord line
// coefficients the Miller loop needs.
//
// Replaces libpbc curve.c (element_mul / element_double / element_pow_mpz on
// G1, affine with one inversion per operation) as reached from bgn.go:344-350,
// 482, 419, 258 -- redesigned inversion-free; affine output is recovered by a
// batched inversion (kernels.cuh: normalize_body).  Everything is three-address
// code over element handles (field.cuh): the only arithmetic instantiated is
// F<L>::mul / add / sub.  Line convention (evaluated at the distorted point
// phi(B) = (-xB, i*yB), F_p factors dropped because the final exponentiation
// kills them):
//     l(B) = (cR + aR*xB) + (bI*yB) i
#pragma once
#include "field.cuh"

template <int L>
struct G {
  typedef F<L> FF;

  // P <- 2P and the tangent line at the old P.  12 products.
  // cR, aR, bI receive the line and double as scratch; t0..t2 scratch.
  BGN_DEV static void dbl_line(E X, E Y, E Z, E cR, E aR, E bI, E t0, E t1, E t2) {
    FF::sqr(t0, X);       // XX
    FF::sqr(t1, Y);       // YY
    FF::sqr(t2, Z);       // ZZ
    FF::sqr(cR, t2);      // ZZ^2
    FF::add(aR, t0, t0);
    FF::add(t0, aR, t0);
    FF::add(t0, t0, cR);  // M = 3XX + ZZ^2   (curve a = 1)
    FF::mul(Z, Y, Z);
    FF::add(Z, Z, Z);     // Z3 = 2YZ
    FF::mul(bI, Z, t2);   // bI = Z3*ZZ
    FF::mul(aR, t0, t2);  // aR = M*ZZ
    FF::mul(cR, t0, X);
    FF::sub(cR, cR, t1);
    FF::sub(cR, cR, t1);  // cR = M*X - 2YY
    FF::mul(t2, X, t1);
    FF::add(t2, t2, t2);
    FF::add(t2, t2, t2);  // S = 4 X YY
    FF::sqr(X, t0);
    FF::sub(X, X, t2);
    FF::sub(X, X, t2);    // X3 = M^2 - 2S
    FF::sub(t2, t2, X);
    FF::mul(Y, t0, t2);   // M (S - X3)
    FF::sqr(t1, t1);
    FF::add(t1, t1, t1);
    FF::add(t1, t1, t1);
    FF::add(t1, t1, t1);  // 8 YYYY
    FF::sub(Y, Y, t1);
  }

  // P <- 2P (no line).  9 products; t0..t3 scratch.
  BGN_DEV static void dbl(E X, E Y, E Z, E t0, E t1, E t2, E t3) {
    FF::sqr(t0, X);
    FF::sqr(t1, Y);
    FF::sqr(t2, Z);
    FF::sqr(t2, t2);
    FF::add(t3, t0, t0);
    FF::add(t0, t3, t0);
    FF::add(t0, t0, t2);  // M
    FF::mul(Z, Y, Z);
    FF::add(Z, Z, Z);
    FF::mul(t2, X, t1);
    FF::add(t2, t2, t2);
    FF::add(t2, t2, t2);  // S
    FF::sqr(X, t0);
    FF::sub(X, X, t2);
    FF::sub(X, X, t2);
    FF::sub(t2, t2, X);
    FF::mul(Y, t0, t2);
    FF::sqr(t1, t1);
    FF::add(t1, t1, t1);
    FF::add(t1, t1, t1);
    FF::add(t1, t1, t1);
    FF::sub(Y, Y, t1);
  }

  // P <- P + (xA, sgn*yA) (mixed) and the chord through them.  13 products.  No special cases: the Miller loop never meets them for points of
  // order n except at the very last step, which the schedule drops (vertical line).
  BGN_DEV static void madd_line(E X, E Y, E Z, const uint32_t* xA, const uint32_t* yA, bool negate, E cR, E aR, E bI, E t0, E t1, E t2) {
    FF::sqr(t0, Z);        // ZZ
    FF::mul(t1, t0, xA);
    FF::sub(t1, t1, X);    // H = U2 - X
    FF::mul(t0, Z, t0);
    FF::mul(t0, t0, yA);   // |S2|
    if (negate)
      FF::add(t0, t0, Y), FF::neg(t0, t0);  // S2 - Y with S2 = -(yA Z^3)
    else
      FF::sub(t0, t0, Y);
    FF::add(aR, t0, t0);   // aR = r = 2(S2 - Y)
    FF::mul(Z, Z, t1);
    FF::add(Z, Z, Z);      // Z3 = 2 Z H
    FF::copy(bI, Z);       // bI = Z3
    FF::mul(cR, aR, xA);
    FF::mul(t0, Z, yA);
    if (negate)
      FF::add(cR, cR, t0);
    else
      FF::sub(cR, cR, t0);  // cR = r*xA - (sgn yA)*Z3
    FF::sqr(t0, t1);
    FF::add(t0, t0, t0);
    FF::add(t0, t0, t0);   // I = 4 HH
    FF::mul(t2, t1, t0);   // J = H*I
    FF::mul(t0, X, t0);    // V = X*I
    FF::sqr(X, aR);
    FF::sub(X, X, t2);
    FF::sub(X, X, t0);
    FF::sub(X, X, t0);     // X3 = r^2 - J - 2V
    FF::sub(t0, t0, X);
    FF::mul(t0, aR, t0);   // r (V - X3)
    FF::mul(t2, Y, t2);
    FF::add(t2, t2, t2);   // 2 Y J
    FF::sub(Y, t0, t2);
  }

  // ---- the doubling-and-addition step of the Miller loop as ONE step (round 2):
  //     T <- (T + A) + T = 2T + A,   A = (xA, +-yA) affine,
  // and the PARABOLA through T (twice), A and -(2T + A) (Eisentraeger-Lauter-Montgomery): the tangent at T and
  // the chord through 2T and A multiply to parabola x vertical, and the vertical lies in F_p at the distorted
  // evaluation points, so f <- f^2 * parabola replaces f <- f^2 * tangent * chord -- one F_p^2 product per
  // evaluation point instead of two.  With R = T + A (slope l1), S = R + T (slope l2):
  //     g(x, y) = (x - xT)(x + xT + xR + l1 l2) - (l1 + l2)(y - yT)
  // In Jacobian coordinates, T = (X, Y, Z):  ZZ = Z^2, H = xA ZZ - X, N1 = yA Z^3 - Y, ZR = Z H,
  //     U = X H^2, W = Y H^3 (T on R's Z), XR = N1^2 - H^3 - 2U, YR = N1 (U - XR) - W;
  // S by the co-Z addition: H2 = XR - U, N2 = YR - W, A2 = H2^2, B2 = U A2, C2 = XR A2,
  //     XS = N2^2 - B2 - C2, YS = N2 (B2 - XS) - W (C2 - B2), ZS = ZR H2;
  // and with An = XR H2 + N1 N2, Sn = N1 H2 + N2, Dn = ZR ZS the parabola scaled by Dn^2, at phi(B) = (-xB, i yB):
  //     g Dn^2 = [cs xB^2 + c1 xB + c0] + [ci yB] i,
  //     cs = Dn^2, c1 = -An Dn, c0 = H2 (Sn W - U (U H2 + An)), ci = -Sn Dn ZR.
  // 11 + 7 + 12 = 30 products where dbl_line + madd_line spend 25; three-address code (its share of the loop is
  // small; the evaluation side is where the step pays: fused.cuh para_mul).  No special cases, as madd_line.
  // X, Y, Z are replaced by S; cs, c1, c0, ci receive the coefficients; t0..t7 scratch.
  BGN_DEV static void dadd_para(E X, E Y, E Z, const uint32_t* xA, const uint32_t* yA, bool negate, E cs, E c1, E c0,
                                E ci, E t0, E t1, E t2, E t3, E t4, E t5, E t6, E t7) {
    FF::sqr(t0, Z);            // ZZ
    FF::mul(t1, xA, t0);
    FF::sub(t1, t1, X);        // H
    FF::mul(t0, Z, t0);        // Z^3
    FF::mul(t0, yA, t0);
    if (negate) FF::neg(t0, t0);
    FF::sub(t0, t0, Y);        // N1
    FF::mul(Z, Z, t1);         // ZR
    FF::sqr(t2, t1);           // H^2
    FF::mul(t3, t1, t2);       // H^3
    FF::mul(t2, X, t2);        // U
    FF::mul(t4, Y, t3);        // W
    FF::sqr(X, t0);
    FF::sub(X, X, t3);
    FF::sub(X, X, t2);
    FF::sub(X, X, t2);         // XR
    FF::sub(t3, t2, X);
    FF::mul(Y, t0, t3);
    FF::sub(Y, Y, t4);         // YR
    FF::sub(t1, X, t2);        // H2
    FF::sub(t3, Y, t4);        // N2
    FF::mul(t5, X, t1);        // XR H2
    FF::mul(t6, t0, t3);       // N1 N2
    FF::add(t5, t5, t6);       // An
    FF::mul(t6, t0, t1);
    FF::add(t6, t6, t3);       // Sn                         (N1 is dead: t0 free)
    FF::sqr(t0, t1);           // A2
    FF::mul(t7, t2, t0);       // B2
    FF::mul(t0, X, t0);        // C2                         (XR is dead)
    FF::sqr(X, t3);
    FF::sub(X, X, t7);
    FF::sub(X, X, t0);         // XS
    FF::sub(t0, t0, t7);       // C2 - B2
    FF::sub(t7, t7, X);        // B2 - XS
    FF::mul(Y, t3, t7);
    FF::mul(t0, t4, t0);
    FF::sub(Y, Y, t0);         // YS                         (N2 is dead: t3 free)
    FF::mul(t3, Z, t1);        // ZS = ZR H2
    FF::mul(t7, Z, t3);        // Dn = ZR ZS
    FF::mul(ci, t6, t7);
    FF::mul(ci, ci, Z);
    FF::neg(ci, ci);           // ci = -Sn Dn ZR
    FF::copy(Z, t3);           // Z <- ZS
    FF::sqr(cs, t7);           // cs = Dn^2
    FF::mul(c1, t5, t7);
    FF::neg(c1, c1);           // c1 = -An Dn
    FF::mul(t0, t2, t1);       // U H2
    FF::add(t0, t0, t5);       // U H2 + An
    FF::mul(t0, t2, t0);       // U (U H2 + An)
    FF::mul(c0, t6, t4);       // Sn W
    FF::sub(c0, c0, t0);
    FF::mul(c0, t1, c0);       // c0 = H2 (Sn W - U (U H2 + An))
  }

  // Complete mixed addition P <- P + (xA, sgn*yA) for scalar multiplication and EAdd: handles
  // P == O, P == A (doubling) and P == -A (-> O).  11 products on the common path.
  BGN_DEVNI static void madd(E X, E Y, E Z, const uint32_t* xA, const uint32_t* yA, bool negate, E t0, E t1, E t2, E t3) {
    if (FF::is_zero(Z)) {  // O + A
      FF::copy(X, xA);
      FF::copy(Y, yA);
      if (negate) FF::neg(Y, Y);
      FF::set_one(Z);
      return;
    }
    FF::sqr(t0, Z);       // ZZ
    FF::mul(t1, t0, xA);
    FF::sub(t1, t1, X);   // H
    FF::mul(t0, Z, t0);
    FF::mul(t0, t0, yA);
    if (negate)
      FF::add(t0, t0, Y), FF::neg(t0, t0);
    else
      FF::sub(t0, t0, Y);  // S2 - Y
    if (FF::is_zero(t1)) {
      if (FF::is_zero(t0)) {  // same point: double the affine one
        FF::copy(X, xA);
        FF::copy(Y, yA);
        if (negate) FF::neg(Y, Y);
        FF::set_one(Z);
        dbl(X, Y, Z, t0, t1, t2, t3);
      } else {  // inverse points
        FF::set_zero(Z);
      }
      return;
    }
    FF::add(t3, t0, t0);  // r
    FF::mul(Z, Z, t1);
    FF::add(Z, Z, Z);     // Z3
    FF::sqr(t0, t1);
    FF::add(t0, t0, t0);
    FF::add(t0, t0, t0);  // I
    FF::mul(t2, t1, t0);  // J
    FF::mul(t0, X, t0);   // V
    FF::sqr(X, t3);
    FF::sub(X, X, t2);
    FF::sub(X, X, t0);
    FF::sub(X, X, t0);    // X3
    FF::sub(t0, t0, X);
    FF::mul(t0, t3, t0);
    FF::mul(t2, Y, t2);
    FF::add(t2, t2, t2);
    FF::sub(Y, t0, t2);
  }

  // y^2 == x^3 + x ?  (pbc curve_from_bytes falls back to O otherwise)
  BGN_DEVNI static bool on_curve(const uint32_t* xA, const uint32_t* yA, E t0, E t1) {
    FF::sqr(t0, xA);
    FF::mul(t0, t0, xA);
    FF::add(t0, t0, xA);
    FF::sqr(t1, yA);
    return FF::equal(t0, t1);
  }
};

// ---------------------------------------------------------------------------------------------------
// The same group in twisted Edwards form, for the fixed-base sums of Encrypt (round 2).
//
// y^2 = x^3 + x is the Montgomery curve B y^2 = x^3 + A x^2 + x with A = 0, B = 1, hence birationally
// equivalent to the twisted Edwards curve  a u^2 + v^2 = 1 + d u^2 v^2  with a = (A + 2) / B = 2,
// d = (A - 2) / B = -2  through  (u, v) = (x / y, (x - 1) / (x + 1)),  x = (1 + v) / (1 - v), y = x / u;
// O <-> (0, 1).  The map is a group isomorphism away from the points of order dividing 4; G1 has odd
// order n, and on points of odd order the unified addition law below has no exceptional cases (they
// all involve a point of even order), so it needs no branches for P + P, P - P or the identity.
//
// Extended coordinates (U : V : Z : T), u = U / Z, v = V / Z, T = U V / Z, and a table point given
// affinely as (u2, v2, t2 = u2 v2) [Hisil-Wong-Carter-Dawson 2008, unified mixed addition]:
//     A = U u2, B = V v2, C = d T t2, E = (U + V)(u2 + v2) - A - B, F = Z - C, G = Z + C, H = B - a A,
//     U3 = E F, V3 = G H, T3 = E H, Z3 = F G
// 8 products where the complete mixed Jacobian addition of G<L>::madd spends 8 products + 3 squarings
// and two zero tests.  a A = 2A and d T t2 = -2 T t2 are additions.
// libpbc has no such form; the results leave this file as Jacobian Weierstrass points (to_jac), so
// the bytes Encrypt returns are the ones curve.c produces (bgn.go:340-353).
template <int L>
struct Ed {
  typedef F<L> FF;
  // (U, V, Z, T) <- (U, V, Z, T) + (ent[0..L), ent[L..2L), t = ent[2L..3L)).  t0..t3 scratch.
  BGN_DEVNI static void madd(E U, E V, E Z, E T, const uint32_t* ent, E t0, E t1, E t2, E t3) {
    const uint32_t *u2 = ent, *v2 = ent + L, *tt = ent + 2 * L;
    FF::mul(t0, U, u2);    // A
    FF::mul(t1, V, v2);    // B
    FF::mul(t2, T, tt);
    FF::add(t2, t2, t2);   // -C = 2 T t2
    FF::add(t3, U, V);
    FF::add(T, u2, v2);
    FF::mul(t3, t3, T);
    FF::sub(t3, t3, t0);
    FF::sub(t3, t3, t1);   // E
    FF::add(U, Z, t2);     // F = Z - C
    FF::sub(V, Z, t2);     // G = Z + C
    FF::add(t0, t0, t0);
    FF::sub(t1, t1, t0);   // H = B - 2A
    FF::mul(Z, U, V);      // Z3 = F G
    FF::mul(T, t3, t1);    // T3 = E H
    FF::mul(U, t3, U);     // U3 = E F
    FF::mul(V, V, t1);     // V3 = G H
  }
  // first term of a sum: the table point itself
  BGN_DEVNI static void set(E U, E V, E Z, E T, const uint32_t* ent) {
    FF::copy(U, ent);
    FF::copy(V, ent + L);
    FF::set_one(Z);
    FF::copy(T, ent + 2 * L);
  }
  // Jacobian Weierstrass (X, Y, Zj) of the extended point (U, V, Z, .): with D = (Z - V) U,
  // x = (Z + V) U / D, y = (Z + V) Z / D, so Zj = D, X = (Z + V) U D, Y = (Z + V) Z D^2.  6 products.
  // The identity (U = 0, V = Z) gives Zj = 0 = O.  U, V, Z are overwritten with the result.
  BGN_DEVNI static void to_jac(E U, E V, E Z, E t0, E t1, E t2) {
    FF::sub(t0, Z, V);
    FF::add(t1, Z, V);
    FF::mul(t0, t0, U);    // D
    FF::mul(t2, t1, U);    // (Z + V) U
    FF::mul(t1, t1, Z);    // (Z + V) Z
    FF::mul(U, t2, t0);    // X
    FF::sqr(t2, t0);       // D^2
    FF::mul(V, t1, t2);    // Y
    FF::copy(Z, t0);       // Zj
  }
};

