// field.cuh -- memory-level F_p / F_p^2 operations on strided element handles.
//
// An element handle V = {pointer to limb 0, stride in words}.  The same code
// addresses (i) the limb-major SoA arrays in HBM ([limb][batch], stride =
// batch), (ii) per-thread state slots in shared memory ([slot][limb][thread],
// stride = blockDim) and (iii) thread-local scratch (stride 1).  The heavy ops
// are __noinline__ so a kernel is a short straight-line program of calls and
// the register allocator only ever sees one or a few Montgomery products.
//
// Replaces libpbc montfp.c / fieldquadratic.c behaviour (element_mul, _add,
// _sub, _invert, _square on F_p and F_p[i]); see SURVEY.md 8(a) row a8.
#pragma once
#include "arith.cuh"

struct V {
  uint32_t* p;
  int s;
};
BGN_DEV V mkv(uint32_t* p, int s) {
  V v;
  v.p = p;
  v.s = s;
  return v;
}
BGN_DEV V mkvc(const uint32_t* p, int s) { return mkv(const_cast<uint32_t*>(p), s); }

template <int L>
BGN_DEV void ld(uint32_t (&r)[L], V a) {
  BGN_UNROLL
  for (int j = 0; j < L; j++) r[j] = a.p[(size_t)j * a.s];
}
template <int L>
BGN_DEV void st(V a, const uint32_t (&r)[L]) {
  BGN_UNROLL
  for (int j = 0; j < L; j++) a.p[(size_t)j * a.s] = r[j];
}

// F_p^2 handle: re and im are two handles (usually consecutive slots)
struct V2 {
  V re, im;
};
BGN_DEV V2 mkv2(V re, V im) {
  V2 v;
  v.re = re;
  v.im = im;
  return v;
}

template <int L>
struct F {
  typedef Fp<L> P;

  // ---------------- F_p ----------------
  BGN_DEVNI static void mul(V r, V a, V b) {
    uint32_t x[L], y[L], z[L];
    ld<L>(x, a);
    ld<L>(y, b);
    P::mul(z, x, y);
    st<L>(r, z);
  }
  BGN_DEVNI static void sqr(V r, V a) {
    uint32_t x[L], z[L];
    ld<L>(x, a);
    P::sqr(z, x);
    st<L>(r, z);
  }
  BGN_DEVNI static void add(V r, V a, V b) {
    uint32_t x[L], y[L], z[L];
    ld<L>(x, a);
    ld<L>(y, b);
    P::add(z, x, y);
    st<L>(r, z);
  }
  BGN_DEVNI static void sub(V r, V a, V b) {
    uint32_t x[L], y[L], z[L];
    ld<L>(x, a);
    ld<L>(y, b);
    P::sub(z, x, y);
    st<L>(r, z);
  }
  BGN_DEV static void dbl(V r, V a) { add(r, a, a); }
  BGN_DEVNI static void neg(V r, V a) {
    uint32_t x[L], y[L], z[L];
    BGN_UNROLL
    for (int j = 0; j < L; j++) x[j] = 0;
    ld<L>(y, a);
    P::sub(z, x, y);
    st<L>(r, z);
  }
  BGN_DEVNI static void copy(V r, V a) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) r.p[(size_t)j * r.s] = a.p[(size_t)j * a.s];
  }
  BGN_DEVNI static void set_one(V r) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) r.p[(size_t)j * r.s] = c_fc.one[j];
  }
  BGN_DEVNI static void set_zero(V r) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) r.p[(size_t)j * r.s] = 0;
  }
  // value == 0 (mod p) for a lazy-form element
  BGN_DEVNI static bool is_zero(V a) {
    uint32_t x[L], y[L];
    ld<L>(x, a);
    P::canon(y, x);
    return P::is_zero_raw(y);
  }
  BGN_DEVNI static bool equal(V a, V b) {
    uint32_t x[L], y[L], u[L], w[L];
    ld<L>(x, a);
    ld<L>(y, b);
    P::canon(u, x);
    P::canon(w, y);
    return P::eq_raw(u, w);
  }
  // r = canonical [0,p) of a (stays in Montgomery form)
  BGN_DEVNI static void canon(V r, V a) {
    uint32_t x[L], y[L];
    ld<L>(x, a);
    P::canon(y, x);
    st<L>(r, y);
  }
  // standard integer -> Montgomery form (a < 2^(32L), result lazy)
  BGN_DEVNI static void to_mont(V r, V a) {
    uint32_t x[L], y[L], z[L];
    ld<L>(x, a);
    BGN_UNROLL
    for (int j = 0; j < L; j++) y[j] = c_fc.r2[j];
    P::mul(z, x, y);
    st<L>(r, z);
  }
  // Montgomery form -> canonical standard integer in [0,p)
  BGN_DEVNI static void from_mont(V r, V a) {
    uint32_t x[L], y[L], z[L];
    ld<L>(x, a);
    BGN_UNROLL
    for (int j = 0; j < L; j++) y[j] = 0;
    y[0] = 1;
    P::mul(z, x, y);
    P::canon(x, z);
    st<L>(r, x);
  }

  // r = a^(p-2) (Fermat inverse); 0 -> 0.  4-bit fixed window, uniform control flow.
  BGN_DEVNI static void inv(V r, V a) {
    uint32_t tab[15][L];  // a^1 .. a^15
    uint32_t acc[L], t[L];
    ld<L>(t, a);
    BGN_UNROLL
    for (int j = 0; j < L; j++) tab[0][j] = t[j];
    for (int i = 1; i < 15; i++) {
      uint32_t u[L], w[L];
      for (int j = 0; j < L; j++) u[j] = tab[i - 1][j];
      P::mul(w, u, t);
      for (int j = 0; j < L; j++) tab[i][j] = w[j];
    }
    // exponent e = p - 2 (p = 3 mod 4, so no borrow out of limb 0)
    bool started = false;
    for (int w = 8 * L - 1; w >= 0; w--) {
      uint32_t limb = c_fc.p[w >> 3];
      if ((w >> 3) == 0) limb -= 2;
      uint32_t d = (limb >> ((w & 7) * 4)) & 15u;
      if (started) {
        P::sqr(t, acc);
        P::sqr(acc, t);
        P::sqr(t, acc);
        P::sqr(acc, t);
      }
      if (d) {
        uint32_t u[L];
        for (int j = 0; j < L; j++) u[j] = tab[d - 1][j];
        if (started) {
          P::mul(t, acc, u);
          for (int j = 0; j < L; j++) acc[j] = t[j];
        } else {
          for (int j = 0; j < L; j++) acc[j] = u[j];
          started = true;
        }
      }
    }
    st<L>(r, acc);
  }

  // ---------------- F_p^2 = F_p[i]/(i^2+1) ----------------
  // r = a*b, Karatsuba, 3 products; operands are read before r is written.
  BGN_DEVNI static void mul2(V2 r, V2 a, V2 b) {
    uint32_t a0[L], a1[L], b0[L], b1[L], t0[L], t1[L], t2[L], s[L], u[L];
    ld<L>(a0, a.re);
    ld<L>(a1, a.im);
    ld<L>(b0, b.re);
    ld<L>(b1, b.im);
    P::mul(t0, a0, b0);
    P::mul(t1, a1, b1);
    P::add(s, a0, a1);
    P::add(u, b0, b1);
    P::mul(t2, s, u);
    P::sub(s, t0, t1);
    st<L>(r.re, s);
    P::sub(u, t2, t0);
    P::sub(s, u, t1);
    st<L>(r.im, s);
  }
  // r = a^2: (a0+a1)(a0-a1) + 2 a0 a1 i, 2 products
  BGN_DEVNI static void sqr2(V2 r, V2 a) {
    uint32_t a0[L], a1[L], s[L], d[L], t[L];
    ld<L>(a0, a.re);
    ld<L>(a1, a.im);
    P::add(s, a0, a1);
    P::sub(d, a0, a1);
    P::mul(t, s, d);
    st<L>(r.re, t);
    P::mul(t, a0, a1);
    P::add(s, t, t);
    st<L>(r.im, s);
  }
  BGN_DEV static void conj2(V2 r, V2 a) {
    copy(r.re, a.re);
    neg(r.im, a.im);
  }
  BGN_DEV static void copy2(V2 r, V2 a) {
    copy(r.re, a.re);
    copy(r.im, a.im);
  }
  BGN_DEV static void set_one2(V2 r) {
    set_one(r.re);
    set_zero(r.im);
  }

  // Miller-loop term: f <- f * ((cR + aR*xB) + (bI*yB) i); 5 products, all
  // intermediates in registers (eval 2 + f*line 3).
  BGN_DEVNI static void line_mul(V2 f, V cR, V aR, V bI, V xB, V yB) {
    uint32_t l0[L], l1[L], x[L], y[L], t0[L], t1[L], t2[L];
    ld<L>(x, aR);
    ld<L>(y, xB);
    P::mul(t0, x, y);
    ld<L>(x, cR);
    P::add(l0, x, t0);
    ld<L>(x, bI);
    ld<L>(y, yB);
    P::mul(l1, x, y);
    ld<L>(x, f.re);
    ld<L>(y, f.im);
    P::mul(t0, x, l0);
    P::mul(t1, y, l1);
    P::add(t2, x, y);
    P::add(x, l0, l1);
    P::mul(y, t2, x);
    P::sub(x, t0, t1);
    st<L>(f.re, x);
    P::sub(x, y, t0);
    P::sub(y, x, t1);
    st<L>(f.im, y);
  }
};
