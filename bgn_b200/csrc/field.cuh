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

// Two kinds of element handle:
//   E  = pointer to L consecutive words (unit stride): thread-local scratch, shared-memory
//        slots, AoS tables.  Limb j is at a compile-time offset, so loads/stores need no
//        address arithmetic (which would otherwise land on the integer-multiply pipe as
//        IMAD.WIDE and compete with the products; measured, profiles/r01_miller_v2_ncu.txt).
//   V  = {pointer, stride in words}: limb-major SoA arrays in HBM (stride = batch size).
// Results and first operands are always E; only the multiplier of mul() and the copy
// helpers take a strided V.
typedef uint32_t* E;
struct V {
  const uint32_t* p;
  int s;
};
BGN_DEV V mkv(const uint32_t* p, int s) {
  V v;
  v.p = p;
  v.s = s;
  return v;
}
BGN_DEV V mkv(const uint32_t* p) { return mkv(p, 1); }

template <int L>
BGN_DEV void ld(uint32_t (&r)[L], const uint32_t* a) {
  BGN_UNROLL
  for (int j = 0; j < L; j++) r[j] = a[j];
}
template <int L>
BGN_DEV void st(E a, const uint32_t (&r)[L]) {
  BGN_UNROLL
  for (int j = 0; j < L; j++) a[j] = r[j];
}

// F_p^2 handle: re and im
struct E2 {
  E re, im;
};
BGN_DEV E2 mke2(E re, E im) {
  E2 v;
  v.re = re;
  v.im = im;
  return v;
}

// thread-local scratch element (lives in local memory because its address is taken)
template <int L>
struct Loc {
  uint32_t w[L];
  BGN_DEV E v() { return w; }
};

template <int L>
struct F {
  typedef Fp<L> P;

  // ---------------- F_p primitives (the only places that touch limbs) ----------------
  // r = a*b (Montgomery).  r may alias a and/or b.  The multiplier b is streamed one limb
  // per row from memory (any stride); a sits in registers.
  BGN_DEVNI static void mul(E r, const uint32_t* a, V b) {
    uint32_t x[L], z[L];
    ld<L>(x, a);
    P::mul_stream(z, x, b.p, b.s);
    st<L>(r, z);
  }
  BGN_DEV static void mul(E r, const uint32_t* a, const uint32_t* b) { mul(r, a, mkv(b, 1)); }
  BGN_DEV static void sqr(E r, const uint32_t* a) { mul(r, a, mkv(a, 1)); }
  BGN_DEVNI static void add(E r, const uint32_t* a, const uint32_t* b) {
    uint32_t x[L], y[L], z[L];
    ld<L>(x, a);
    ld<L>(y, b);
    P::add(z, x, y);
    st<L>(r, z);
  }
  BGN_DEVNI static void sub(E r, const uint32_t* a, const uint32_t* b) {
    uint32_t x[L], y[L], z[L];
    ld<L>(x, a);
    ld<L>(y, b);
    P::sub(z, x, y);
    st<L>(r, z);
  }
  BGN_DEV static void dbl(E r, const uint32_t* a) { add(r, a, a); }
  BGN_DEVNI static void neg(E r, const uint32_t* a) {
    uint32_t x[L], y[L], z[L];
    BGN_UNROLL
    for (int j = 0; j < L; j++) x[j] = 0;
    ld<L>(y, a);
    P::sub(z, x, y);
    st<L>(r, z);
  }
  // strided copies: the only way data moves between SoA arrays and unit-stride elements
  BGN_DEVNI static void load(E r, V a) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = a.p[(size_t)j * a.s];
  }
  BGN_DEVNI static void store(uint32_t* dst, int stride, const uint32_t* a) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) dst[(size_t)j * stride] = a[j];
  }
  BGN_DEV static void copy(E r, const uint32_t* a) { load(r, mkv(a, 1)); }
  BGN_DEVNI static void set_one(E r) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = c_fc.one[j];
  }
  BGN_DEVNI static void set_zero(E r) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = 0;
  }
  BGN_DEVNI static void store_one(uint32_t* dst, int stride) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) dst[(size_t)j * stride] = c_fc.one[j];
  }
  BGN_DEVNI static void store_zero(uint32_t* dst, int stride) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) dst[(size_t)j * stride] = 0;
  }
  // value == 0 (mod p) for a lazy-form element
  BGN_DEVNI static bool is_zero(const uint32_t* a) {
    uint32_t x[L], y[L];
    ld<L>(x, a);
    P::canon(y, x);
    return P::is_zero_raw(y);
  }
  BGN_DEVNI static bool equal(const uint32_t* a, const uint32_t* b) {
    uint32_t x[L], y[L], u[L], w[L];
    ld<L>(x, a);
    ld<L>(y, b);
    P::canon(u, x);
    P::canon(w, y);
    return P::eq_raw(u, w);
  }
  BGN_DEVNI static bool is_one(const uint32_t* a) {
    uint32_t x[L], y[L], o[L], w[L];
    ld<L>(x, a);
    P::canon(y, x);
    BGN_UNROLL
    for (int j = 0; j < L; j++) o[j] = c_fc.one[j];
    P::canon(w, o);
    return P::eq_raw(y, w);
  }
  // r = canonical [0,p) of a (stays in Montgomery form)
  BGN_DEVNI static void canon(E r, const uint32_t* a) {
    uint32_t x[L], y[L];
    ld<L>(x, a);
    P::canon(y, x);
    st<L>(r, y);
  }
  // standard integer -> Montgomery form (a < 2^(32L), result lazy)
  BGN_DEV static void to_mont(E r, const uint32_t* a) { mul(r, a, mkv(c_fc.r2, 1)); }
  // Montgomery form -> canonical standard integer in [0,p)
  BGN_DEVNI static void from_mont(E r, const uint32_t* a) {
    uint32_t y[L];
    BGN_UNROLL
    for (int j = 0; j < L; j++) y[j] = 0;
    y[0] = 1;
    mul(r, a, mkv(y, 1));
    canon(r, r);
  }

  // r = a^(p-2) (Fermat inverse); 0 -> 0.  Left-to-right binary, uniform control flow across
  // threads (the exponent is a key constant).  t must not alias r or a; r may alias a.
  BGN_DEVNI static void inv(E r, const uint32_t* a, E t) {
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
  BGN_DEV static void mul2(E2 r, E2 a, E2 b, E t0, E t1, E t2) {
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
  BGN_DEV static void sqr2(E2 r, E2 a, E t0, E t1) {
    add(t0, a.re, a.im);
    sub(t1, a.re, a.im);
    mul(r.im, a.re, a.im);
    add(r.im, r.im, r.im);
    mul(r.re, t0, t1);
  }
  BGN_DEV static void conj2(E2 r, E2 a) {
    copy(r.re, a.re);
    neg(r.im, a.im);
  }
  BGN_DEV static void copy2(E2 r, E2 a) {
    copy(r.re, a.re);
    copy(r.im, a.im);
  }
  BGN_DEV static void set_one2(E2 r) {
    set_one(r.re);
    set_zero(r.im);
  }

  // Miller-loop term: f <- f * ((cR + aR*xB) + (bI*yB) i); 5 products
  // (eval 2 + f*line 3).  xB, yB may be strided (SoA in HBM).  t0..t2 scratch.
  BGN_DEV static void line_mul(E2 f, const uint32_t* cR, const uint32_t* aR, const uint32_t* bI, V xB, V yB, E t0, E t1,
                               E t2) {
    mul(t0, aR, xB);
    add(t0, t0, cR);  // l0
    mul(t1, bI, yB);  // l1
    mul(t2, f.re, t0);
    add(f.re, f.re, f.im);
    add(t0, t0, t1);
    mul(t1, f.im, t1);  // f1*l1
    mul(f.im, f.re, t0);
    sub(f.re, t2, t1);
    sub(f.im, f.im, t2);
    sub(f.im, f.im, t1);
  }
};
