// field.cuh -- memory-level F_p / F_p^2 operations on strided element handles.
//
// An element handle V = {pointer to limb 0, stride in words}.  The same code
// addresses (i) the limb-major SoA arrays in HBM ([limb][batch], stride =
// batch), (ii) per-thread state slots in shared memory ([slot][limb][thread],
// stride = blockDim) and (iii) thread-local scratch (stride 1).
//
// Code-size rule (measured, profiles/r01_miller_v1.md): with the Montgomery
// product inlined at every use the Miller kernel was 625 KB of SASS and spent
// most of its issue slots waiting for instruction fetch.  So there is exactly
// ONE copy of the product per translation unit -- F<L>::mul, __noinline__,
// operands and result in memory -- and everything above it (F_p^2, curve,
// pairing) is written as straight-line "three-address code" over handles with
// caller-provided temporaries.  The whole hot loop is then ~20 KB of SASS and
// stays in the instruction cache; register pressure is that of one product.
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

// thread-local scratch element (lives in local memory because its address is taken)
template <int L>
struct Loc {
  uint32_t w[L];
  BGN_DEV V v() { return mkv(w, 1); }
};

template <int L>
struct F {
  typedef Fp<L> P;

  // ---------------- F_p primitives (the only places that touch limbs) ----------------
  // r = a*b (Montgomery).  r may alias a and/or b.
  BGN_DEVNI static void mul(V r, V a, V b) {
    uint32_t x[L], y[L], z[L];
    ld<L>(x, a);
    ld<L>(y, b);
    P::mul(z, x, y);
    st<L>(r, z);
  }
  BGN_DEV static void sqr(V r, V a) { mul(r, a, a); }
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
  BGN_DEVNI static bool is_one(V a) {
    uint32_t x[L], y[L], o[L], w[L];
    ld<L>(x, a);
    P::canon(y, x);
    BGN_UNROLL
    for (int j = 0; j < L; j++) o[j] = c_fc.one[j];
    P::canon(w, o);
    return P::eq_raw(y, w);
  }
  // r = canonical [0,p) of a (stays in Montgomery form)
  BGN_DEVNI static void canon(V r, V a) {
    uint32_t x[L], y[L];
    ld<L>(x, a);
    P::canon(y, x);
    st<L>(r, y);
  }
  // standard integer -> Montgomery form (a < 2^(32L), result lazy)
  BGN_DEV static void to_mont(V r, V a) { mul(r, a, mkvc(c_fc_r2(), 1)); }
  // Montgomery form -> canonical standard integer in [0,p)
  BGN_DEVNI static void from_mont(V r, V a) {
    uint32_t y[L];
    BGN_UNROLL
    for (int j = 0; j < L; j++) y[j] = 0;
    y[0] = 1;
    mul(r, a, mkv(y, 1));
    canon(r, r);
  }
  BGN_DEV static const uint32_t* c_fc_r2() { return c_fc.r2; }

  // r = a^(p-2) (Fermat inverse); 0 -> 0.  Left-to-right binary, uniform control flow across
  // threads (the exponent is a key constant).  t must not alias r or a; r may alias a.
  BGN_DEVNI static void inv(V r, V a, V t) {
    copy(t, a);
    int top = 32 * L - 1;
    while (top > 0 && !((c_fc.p[top >> 5] >> (top & 31)) & 1)) top--;
    // exponent e = p - 2: p = 3 (mod 4), so subtracting 2 only touches the low limb (no borrow)
    copy(r, t);  // consumes the top bit of e
    for (int bit = top - 1; bit >= 0; bit--) {
      uint32_t limb = c_fc.p[bit >> 5];
      if ((bit >> 5) == 0) limb -= 2;
      sqr(r, r);
      if ((limb >> (bit & 31)) & 1) mul(r, r, t);
    }
  }

  // ---------------- F_p^2 = F_p[i]/(i^2+1), three-address code ----------------
  // r = a*b (Karatsuba, 3 products).  r may alias a or b; t0..t2 are scratch.
  BGN_DEV static void mul2(V2 r, V2 a, V2 b, V t0, V t1, V t2) {
    mul(t0, a.re, b.re);
    mul(t1, a.im, b.im);
    add(t2, a.re, a.im);
    add(r.im, b.re, b.im);
    mul(r.im, t2, r.im);
    sub(r.re, t0, t1);
    sub(r.im, r.im, t0);
    sub(r.im, r.im, t1);
  }
  // r = a^2: (a0+a1)(a0-a1) + 2 a0 a1 i, 2 products.  r may alias a.
  BGN_DEV static void sqr2(V2 r, V2 a, V t0, V t1) {
    add(t0, a.re, a.im);
    sub(t1, a.re, a.im);
    mul(r.im, a.re, a.im);
    add(r.im, r.im, r.im);
    mul(r.re, t0, t1);
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

  // Miller-loop term: f <- f * ((cR + aR*xB) + (bI*yB) i); 5 products
  // (eval 2 + f*line 3).  t0..t3 scratch.
  BGN_DEV static void line_mul(V2 f, V cR, V aR, V bI, V xB, V yB, V t0, V t1, V t2, V t3) {
    mul(t0, aR, xB);
    add(t0, t0, cR);  // l0
    mul(t1, bI, yB);  // l1
    mul(t2, f.re, t0);
    mul(t3, f.im, t1);
    add(f.re, f.re, f.im);
    add(t0, t0, t1);
    mul(f.im, f.re, t0);
    sub(f.re, t2, t3);
    sub(f.im, f.im, t2);
    sub(f.im, f.im, t3);
  }
};
