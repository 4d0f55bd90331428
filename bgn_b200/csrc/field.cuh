// field.cuh -- memory-level F_p / F_p^2 operations on strided element handles.
//
// An element handle is a pointer to L consecutive words; the same code addresses (i) the
// batch arrays in HBM ([N][L]), (ii) per-thread state slots in shared memory
// ([slot][thread][L]) and (iii) thread-local scratch.
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

// Element handle: E = pointer to L consecutive words.  Every device array is array-of-elements
// ([N][L] words, element e at e*L): thread-local scratch, shared-memory slots, tables and the
// batch arrays in HBM alike.  Limb j of any operand is then at a compile-time offset, so loads and
// stores need no address arithmetic -- with a runtime stride (limb-major SoA) ptxas computes each
// limb address with an IMAD.WIDE, i.e. on the very pipe the products saturate (measured:
// tools/primbench.py, 85.0 % -> 89.7 % of the IMAD.WIDE peak for the memory-operand product).
typedef uint32_t* E;

// (ES = element stride: limb j at a[j * ES]; 1 everywhere except the Miller kernel's
// thread-interleaved layout for the 1024-bit field, pairing.cuh)
template <int L, int ES = 1>
BGN_DEV void ld(uint32_t (&r)[L], const uint32_t* a) {
  BGN_SETB(r, BGN_GETB(a));
  BGN_UNROLL
  for (int j = 0; j < L; j++) r[j] = a[j * ES];
}
template <int L, int ES = 1>
BGN_DEV void st(E a, const uint32_t (&r)[L]) {
  BGN_SETB(a, BGN_GETB(r));
  BGN_UNROLL
  for (int j = 0; j < L; j++) a[j * ES] = r[j];
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
  // r = a*b (Montgomery).  r may alias a and/or b.  Both operands are loaded up front (measured
  // 88.5 % of the IMAD.WIDE peak at 2 warps/scheduler vs 85 % when the multiplier is read row by
  // row, tools/primbench.py modes 20 / 10).
  BGN_DEVNI static void mul(E r, const uint32_t* a, const uint32_t* b) {
    uint32_t x[L], y[L], z[L];
    ld<L>(x, a);
    ld<L>(y, b);
    P::mul(z, x, y);
    st<L>(r, z);
  }
  // r = a^2 through the dedicated squaring (arith.cuh: Fp::sqr): 3 of the 11 products of a mixed addition,
  // 6 of the 9 of a doubling, so Encrypt / MultConst / the table builds execute ~8 % / ~15 % fewer products
  BGN_DEVNI static void sqr(E r, const uint32_t* a) {
    uint32_t x[L], z[L];
    ld<L>(x, a);
    P::sqr(z, x);
    st<L>(r, z);
  }
  // measured alternatives (tools/primbench.py modes 10, 22): multiplier streamed from memory / two
  // interleaved products per call
  BGN_DEVNI static void mul_stream(E r, const uint32_t* a, const uint32_t* b) {
    uint32_t x[L], z[L];
    ld<L>(x, a);
    P::mul_stream(z, x, b);
    st<L>(r, z);
  }
  BGN_DEVNI static void mul_pair(E r1, const uint32_t* a1, const uint32_t* b1, E r2, const uint32_t* a2,
                                 const uint32_t* b2) {
    uint32_t x1[L], x2[L], z1[L], z2[L];
    ld<L>(x1, a1);
    ld<L>(x2, a2);
    P::mul_pair(z1, x1, b1, z2, x2, b2);
    st<L>(r1, z1);
    st<L>(r2, z2);
  }
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
    BGN_SETB(x, 0.0);
    BGN_UNROLL
    for (int j = 0; j < L; j++) x[j] = 0;
    ld<L>(y, a);
    P::sub(z, x, y);
    st<L>(r, z);
  }
  BGN_DEVNI static void copy(E r, const uint32_t* a) {
    BGN_SETB(r, BGN_GETB(a));
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = a[j];
  }
  BGN_DEVNI static void set_one(E r) {
    BGN_SETB(r, 1.0);
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = c_fc.one[j];
  }
  BGN_DEVNI static void set_zero(E r) {
    BGN_SETB(r, 0.0);
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = 0;
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
    BGN_SETB(o, 1.0);
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
  BGN_DEV static void to_mont(E r, const uint32_t* a) { mul(r, a, c_fc.r2); }
  // Montgomery form -> canonical standard integer in [0,p)
  BGN_DEVNI static void from_mont(E r, const uint32_t* a) {
    uint32_t y[L];
    BGN_SETB(y, 1.0);
    BGN_UNROLL
    for (int j = 0; j < L; j++) y[j] = 0;
    y[0] = 1;
    mul(r, a, y);
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

  // r = a^-1 (Montgomery in, Montgomery out; 0 -> 0) through the binary GCD of arith.cuh:
  // inv(a R) = a^-1 R^-1 as an integer, times R^3 (= r2 * r2 / R) in a Montgomery product = a^-1 R.
  // r may alias a.  in: a < 2p.  out: r < 2p.
  // RELAXED: the operand may be anything below 8p (the fused routines' range)
  template <bool RELAXED = false, int ES = 1>
  BGN_DEVNI static void inv_gcd(E r, const uint32_t* a) {
    uint32_t x[L], y[L], r2[L], r3[L];
    ld<L, ES>(x, a);
    if (RELAXED) P::norm2p(x, x);
    P::canon(x, x);
    P::inv_bgcd(y, x);
    ld<L>(r2, c_fc.r2);
    P::mul(r3, r2, r2);
    P::mul(x, y, r3);
    st<L, ES>(r, x);
  }
  // The same inversion through the batched division steps (Fp::inv_safegcd, ~4x fewer instructions),
  // VERIFIED with one product (a * a^-1 == 1) and falling back to the plain binary GCD if the check
  // fails -- a flaw in the fast routine can cost time, never correctness.  Used by k_normalize and
  // k_g1_affadd, where the inversion is the bottleneck; the Miller kernels keep inv_gcd (their
  // inversion is ~0.1 % of the work and their register allocation is left alone).
  template <bool RELAXED = false>
  BGN_DEVNI static void inv_gcd_fast(E r, const uint32_t* a) {
    uint32_t x[L], y[L], z[L], chk[L], one[L], r3[L];
    ld<L>(x, a);
    if (RELAXED) P::norm2p(x, x);
    P::canon(x, x);
    ld<L>(one, c_fc.r2);
    P::mul(r3, one, one);
    bool ok = P::inv_safegcd(y, x);
    P::mul(z, y, r3);
    P::mul(chk, z, x);
    P::canon(chk, chk);
    ld<L>(one, c_fc.one);
    P::canon(one, one);
    const bool zero = P::is_zero_raw(x);
    if (zero || (ok && P::eq_raw(chk, one))) {
      if (zero) {
        BGN_UNROLL
        for (int j = 0; j < L; j++) z[j] = 0;
      }
      st<L>(r, z);
      return;
    }
#ifdef BGN_HOSTSIM
    bgnsim::safegcd_fallbacks++;
#endif
    P::inv_bgcd(y, x);
    P::mul(x, y, r3);
    st<L>(r, x);
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
  // (eval 2 + f*line 3).  t0..t2 scratch.
  BGN_DEV static void line_mul(E2 f, const uint32_t* cR, const uint32_t* aR, const uint32_t* bI, const uint32_t* xB, const uint32_t* yB, E t0, E t1,
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
