// curve.cuh -- G1 = E(F_p)[n], E: y^2 = x^3 + x, in Jacobian coordinates
// (x = X/Z^2, y = Y/Z^3; O is Z == 0), plus the tangent / chord line
// coefficients the Miller loop needs.
//
// Replaces libpbc curve.c (element_mul / element_double / element_pow_mpz on
// G1, affine with one inversion per operation) as reached from bgn.go:344-350,
// 482, 419, 258 -- redesigned inversion-free; affine output is recovered by a
// batched inversion (kernels.cu: k_normalize).  Line convention (evaluated at
// the distorted point phi(B) = (-xB, i*yB), F_p factors dropped because the
// final exponentiation kills them):
//     l(B) = (cR + aR*xB) + (bI*yB) i
#pragma once
#include "field.cuh"

template <int L>
struct G {
  typedef Fp<L> P;
  typedef F<L> FF;

  // V <- 2V and tangent line at the old V.  12 products.
  template <bool LINE>
  BGN_DEV static void dbl_impl(V X, V Y, V Z, V cR, V aR, V bI) {
    uint32_t x[L], y[L], z[L], xx[L], yy[L], zz[L], m[L], s[L], t[L], u[L];
    ld<L>(x, X);
    ld<L>(y, Y);
    ld<L>(z, Z);
    P::sqr(xx, x);
    P::sqr(yy, y);
    P::sqr(zz, z);
    P::sqr(t, zz);
    P::add(m, xx, xx);
    P::add(m, m, xx);
    P::add(m, m, t);  // M = 3XX + ZZ^2   (curve a = 1)
    P::mul(t, x, yy);
    P::add(s, t, t);
    P::add(s, s, s);  // S = 4 X YY
    P::mul(t, y, z);
    P::add(u, t, t);  // Z3 = 2YZ
    st<L>(Z, u);
    if (LINE) {
      P::mul(t, u, zz);
      st<L>(bI, t);  // bI = Z3*ZZ
      P::mul(t, m, zz);
      st<L>(aR, t);  // aR = M*ZZ
      P::mul(t, m, x);
      P::add(u, yy, yy);
      P::sub(t, t, u);
      st<L>(cR, t);  // cR = M*X - 2YY
    }
    P::sqr(t, m);
    P::sub(t, t, s);
    P::sub(t, t, s);  // X3 = M^2 - 2S
    st<L>(X, t);
    P::sub(s, s, t);
    P::mul(u, m, s);  // M (S - X3)
    P::sqr(t, yy);
    P::add(t, t, t);
    P::add(t, t, t);
    P::add(t, t, t);  // 8 YYYY
    P::sub(u, u, t);
    st<L>(Y, u);
  }
  BGN_DEVNI static void dbl_line(V X, V Y, V Z, V cR, V aR, V bI) { dbl_impl<true>(X, Y, Z, cR, aR, bI); }
  BGN_DEVNI static void dbl(V X, V Y, V Z) { dbl_impl<false>(X, Y, Z, X, X, X); }

  // V <- V + (xA, sgn*yA) (mixed) and the chord through them.  13 products.
  // No special cases: the Miller loop never meets them for points of order n
  // except at the very last step, which the schedule drops (vertical line).
  BGN_DEVNI static void madd_line(V X, V Y, V Z, V xA, V yA, bool negate, V cR, V aR, V bI) {
    uint32_t x[L], y[L], z[L], ax[L], ay[L], zz[L], h[L], r[L], t[L], u[L], i4[L], j[L], vv[L];
    ld<L>(x, X);
    ld<L>(y, Y);
    ld<L>(z, Z);
    ld<L>(ax, xA);
    ld<L>(t, yA);
    if (negate) {
      BGN_UNROLL
      for (int k = 0; k < L; k++) u[k] = 0;
      P::sub(ay, u, t);
    } else {
      BGN_UNROLL
      for (int k = 0; k < L; k++) ay[k] = t[k];
    }
    P::sqr(zz, z);
    P::mul(t, ax, zz);  // U2
    P::sub(h, t, x);    // H
    P::mul(t, z, zz);
    P::mul(u, ay, t);  // S2
    P::sub(r, u, y);
    P::add(r, r, r);  // r = 2(S2 - Y)
    P::mul(t, z, h);
    P::add(t, t, t);  // Z3 = 2 Z H
    st<L>(Z, t);
    st<L>(bI, t);  // bI = Z3
    st<L>(aR, r);  // aR = r
    P::mul(u, ay, t);
    P::mul(t, r, ax);
    P::sub(t, t, u);
    st<L>(cR, t);  // cR = r*xA - yA*Z3
    P::sqr(t, h);
    P::add(t, t, t);
    P::add(i4, t, t);    // I = 4 HH
    P::mul(j, h, i4);    // J
    P::mul(vv, x, i4);   // V
    P::sqr(t, r);
    P::sub(t, t, j);
    P::sub(t, t, vv);
    P::sub(t, t, vv);  // X3
    st<L>(X, t);
    P::sub(u, vv, t);
    P::mul(t, r, u);
    P::mul(u, y, j);
    P::add(u, u, u);
    P::sub(t, t, u);  // Y3 = r(V - X3) - 2 Y J
    st<L>(Y, t);
  }

  // Complete mixed addition V <- V + (xA, sgn*yA) for scalar multiplication and
  // EAdd: handles V == O, V == A (doubling) and V == -A (-> O).  11 products on
  // the common path.
  BGN_DEVNI static void madd(V X, V Y, V Z, V xA, V yA, bool negate) {
    uint32_t x[L], y[L], z[L], ax[L], ay[L], zz[L], h[L], r[L], t[L], u[L], i4[L], j[L], vv[L];
    ld<L>(z, Z);
    ld<L>(ax, xA);
    ld<L>(t, yA);
    if (negate) {
      BGN_UNROLL
      for (int k = 0; k < L; k++) u[k] = 0;
      P::sub(ay, u, t);
    } else {
      BGN_UNROLL
      for (int k = 0; k < L; k++) ay[k] = t[k];
    }
    P::canon(t, z);
    if (P::is_zero_raw(t)) {  // O + A
      st<L>(X, ax);
      st<L>(Y, ay);
      FF::set_one(Z);
      return;
    }
    ld<L>(x, X);
    ld<L>(y, Y);
    P::sqr(zz, z);
    P::mul(t, ax, zz);
    P::sub(h, t, x);
    P::mul(t, z, zz);
    P::mul(u, ay, t);
    P::sub(r, u, y);
    P::canon(t, h);
    if (P::is_zero_raw(t)) {
      P::canon(t, r);
      if (P::is_zero_raw(t)) {  // same point: double the affine one
        st<L>(X, ax);
        st<L>(Y, ay);
        FF::set_one(Z);
        dbl(X, Y, Z);
      } else {  // inverse points
        FF::set_zero(Z);
      }
      return;
    }
    P::add(r, r, r);
    P::mul(t, z, h);
    P::add(t, t, t);
    st<L>(Z, t);
    P::sqr(t, h);
    P::add(t, t, t);
    P::add(i4, t, t);
    P::mul(j, h, i4);
    P::mul(vv, x, i4);
    P::sqr(t, r);
    P::sub(t, t, j);
    P::sub(t, t, vv);
    P::sub(t, t, vv);
    st<L>(X, t);
    P::sub(u, vv, t);
    P::mul(t, r, u);
    P::mul(u, y, j);
    P::add(u, u, u);
    P::sub(t, t, u);
    st<L>(Y, t);
  }

  // y^2 == x^3 + x ?  (pbc curve_from_bytes falls back to O otherwise)
  BGN_DEVNI static bool on_curve(V xA, V yA) {
    uint32_t x[L], y[L], t[L], u[L];
    ld<L>(x, xA);
    ld<L>(y, yA);
    P::sqr(t, x);
    P::mul(u, t, x);
    P::add(u, u, x);
    P::sqr(t, y);
    P::sub(t, t, u);
    P::canon(u, t);
    return P::is_zero_raw(u);
  }
};
