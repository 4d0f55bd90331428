// fused.cuh -- the Miller loop's four inner routines as FUSED register-level programs.
//
// Why (measured, profiles/r01_miller_v4_ncu.txt + r01_primbench_v4.txt): a warp can issue one
// IMAD.WIDE per ~6 cycles while the pipe accepts one per 4 cycles per scheduler, so the pipe only
// saturates when BOTH warps of a scheduler are inside a product at the same time.  With the
// three-address code of field.cuh every F_p operation is a call that loads its operands from
// shared memory, runs, and stores its result; the loads, the carry chains of the add/sub glue and
// the stores of one operation never overlap the products of the same warp, and the sequences ran
// at 72-76 % of the pipe where the bare product reaches 98 %.  Here each routine is ONE function:
// intermediates stay in registers, the glue sits in the issue slots the multiplier leaves free,
// and additions are "relaxed" (no conditional subtraction):
//
//   values are bounded multiples of p, far below R = 2^(32L) >= 256 p.  A sum is a plain
//   multi-limb add; a difference adds a fixed K p (K in {2,4,8,16}) that keeps it non-negative;
//   the Montgomery product absorbs the slack -- a < A p, b < B p gives a b / R mod p below
//   (A B / 256 + 1) p.  The CPU run of this very code (tests/hostsim) propagates the worst-case
//   bound of every value and fails on any product operand >= R - p or any difference whose
//   subtrahend may exceed its offset, so the ranges are proven, not sampled.
//
// Every product is (register operand) x (memory operand): the multiplier is read from its
// shared-memory slot row by row, so it never occupies registers.  U selects the non-unrolled
// row loop (Fp::mul_loop<U>, 2U rows per iteration, 1 + 2U rows of code per product) or, U = 0,
// the fully unrolled product (Fp::mul_stream); which one ships is decided by measurement
// (tools/primbench.py).
//
// Replaces the same libpbc behaviour as curve.cuh / pairing.cuh (a1_param.c Miller loop steps).
#pragma once
#include "field.cuh"

template <int L, int U, int ES = 1>
struct MF {
  typedef Fp<L> P;
  typedef uint32_t R[L];

  BGN_DEV static void mulm(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t* bp) {
    if (U > 0)
      P::template mul_loop<(U > 0 ? U : 1), ES>(r, a, bp);
    else
      P::template mul_stream<ES>(r, a, bp);
  }
  BGN_DEV static void dbl(uint32_t (&r)[L], const uint32_t (&a)[L]) { P::addn(r, a, a); }
  // r = a^2 where `mem` holds the same value as the register operand a: the product, or (BGN_FUSED_SQR,
  // unrolled products only) the dedicated squaring of arith.cuh -- 459 instead of 595 products at L = 17.
  // Measured in k_miller (profiles/r02_bench_n1_v4.json against _v5): the 2^14-product step takes 596.7 ms
  // with it and 596.9 ms without -- its 136 saved products per squaring are paid back in the zeroing,
  // doubling and merging of its two double-width arrays and in shorter carry chains -- while the executed
  // product count, the numerator of the roofline fraction, drops 1.7 %.  Not shipped in the fused
  // routines; Encrypt's three-address code (field.cuh: F::sqr, +4.5 %) and the two-warp pairing
  // (pairwarp.cuh, +4.7 %) do use it.
#ifndef BGN_FUSED_SQR
#define BGN_FUSED_SQR 0
#endif
  static constexpr bool SQR = (U == 0) && BGN_FUSED_SQR != 0;
  BGN_DEV static void sqrm(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t* mem) {
    if (SQR)
      P::sqr(r, a);
    else
      mulm(r, a, mem);
  }

  // f <- f * ((cR + aR*xB) + (bI*yB) i): the Miller-loop term, 5 products.
  // in: f.re, f.im < 8p; cR < 8p; aR < 64p; bI, xB, yB < 8p.   out: f.re < 4p, f.im < 7p.
  BGN_DEVNI static void line_mul(E fre, E fim, const uint32_t* cR, const uint32_t* aR, const uint32_t* bI,
                                 const uint32_t* xB, const uint32_t* yB) {
    R a, l0, l1, t, u, v;
    ld<L, ES>(a, xB);
    mulm(l0, a, aR);
    ld<L, ES>(a, cR);
    P::addn(l0, l0, a);  // l0 = cR + aR xB
    ld<L, ES>(a, yB);
    mulm(l1, a, bI);     // l1 = bI yB
    mulm(t, l0, fre);    // f0 l0
    mulm(u, l1, fim);    // f1 l1
    ld<L, ES>(a, fre);
    ld<L, ES>(v, fim);
    P::addn(a, a, v);
    st<L, ES>(fre, a);       // f0 + f1 (f0 itself is dead)
    P::addn(l0, l0, l1);
    mulm(v, l0, fre);    // (f0 + f1)(l0 + l1)
    P::subk(a, t, u, c_fc.p2, 2);
    st<L, ES>(fre, a);       // f0 l0 - f1 l1
    P::addn(t, t, u);
    P::subk(v, v, t, c_fc.p4, 4);
    st<L, ES>(fim, v);       // f0 l1 + f1 l0
  }

  // The same term with LAZY REDUCTION: the three products of the F_p^2 multiplication stay
  // double-width and share two Montgomery reductions (re = redc(f0 l0 - f1 l1 + p R),
  // im = redc((f0 + f1)(l0 + l1) - f0 l0 - f1 l1)): 5 multiplications + 4 reductions =
  // 5 L^2 + 4 (L^2 + L) products instead of 5 (2 L^2 + L), i.e. 2669 instead of 2975 at L = 17.
  // in: as line_mul.   out: f.re < 3p, f.im < 2p.
  // KM selects the multiplier: 0 = schoolbook rows, 1 = one level of Karatsuba (arith.cuh: Kara)
  // for the three double-width products, 2 = also for aR xB and bI yB (then reduced separately).
  template <int KM>
  BGN_DEV static void mulwk(uint32_t (&T)[2 * L], const uint32_t (&a)[L], const uint32_t* bp) {
    if (KM > 0 && L >= 9)
      Kara<L>::template mulw<ES>(T, a, bp);
    else
      P::template mulw<ES>(T, a, bp);
  }
  template <int KM>
  BGN_DEVNI static void line_mul_lazy(E fre, E fim, const uint32_t* cR, const uint32_t* aR, const uint32_t* bI,
                                      const uint32_t* xB, const uint32_t* yB) {
    R a, b, c, l0, l1;
    uint32_t T0[2 * L], T1[2 * L], S[2 * L];
    ld<L, ES>(a, xB);
    if (KM > 1 && L >= 9) {
      Kara<L>::template mulw<ES>(T0, a, aR);
      P::redc(l0, T0);
    } else {
      mulm(l0, a, aR);
    }
    ld<L, ES>(a, cR);
    P::addn(l0, l0, a);       // l0 = cR + aR xB
    ld<L, ES>(a, yB);
    if (KM > 1 && L >= 9) {
      Kara<L>::template mulw<ES>(T0, a, bI);
      P::redc(l1, T0);
    } else {
      mulm(l1, a, bI);        // l1 = bI yB
    }
    mulwk<KM>(T0, l0, fre);   // f0 l0
    mulwk<KM>(T1, l1, fim);   // f1 l1
    P::addw(S, T0, T1);
    P::subw_k(T0, T0, T1, c_fc.p, 1);
    P::redc(a, T0);           // re; stays in registers while the f.re slot feeds the last product
    ld<L, ES>(b, fre);
    ld<L, ES>(c, fim);
    P::addn(b, b, c);
    st<L, ES>(fre, b);            // f0 + f1
    P::addn(l0, l0, l1);
    mulwk<KM>(T1, l0, fre);   // (f0 + f1)(l0 + l1)
    P::subw(T1, T1, S);       // = f0 l1 + f1 l0 >= 0
    st<L, ES>(fre, a);
    P::redc(b, T1);
    st<L, ES>(fim, b);
  }

  // ---- the line at a NORMALISED evaluation point (round 2).  F_p factors of a line value die in the
  // final exponentiation, so the line may be divided by yB: with uB = xB / yB and vB = 1 / yB (one
  // inversion per evaluation point, before the loop: MillerTeam::init)
  //     l / yB = (cR vB + aR uB) + bI i
  // -- the imaginary part is the line coefficient itself, and the real part is ONE dot product
  // (arith.cuh: dot2_stream, 3L^2 + L) where the plain form spends two products (4L^2 + 2L):
  // 8L^2 + 3L = 2363 products per line at L = 17 against 2669 (line_mul_lazy_n), 9L^2 + 4L against
  // 10L^2 + 5L without lazy reduction (line_mul_n).  yB != 0: the only point of the curve with y = 0 is
  // (0, 0), which the byte format reads as O.
  // in: f.re, f.im < 8p; cR < 8p; aR < 64p; bI < 8p; uB, vB < 2p.   out: as line_mul / line_mul_lazy.
  BGN_DEVNI static void line_mul_n(E fre, E fim, const uint32_t* cR, const uint32_t* aR, const uint32_t* bI,
                                   const uint32_t* uB, const uint32_t* vB) {
    R a, l0, l1, t, u, v;
    ld<L, ES>(a, vB);
    ld<L, ES>(v, uB);
    P::template dot2_stream<ES>(l0, a, cR, v, aR);  // l0 = cR vB + aR uB
    ld<L, ES>(l1, bI);                              // l1 = bI
    mulm(t, l0, fre);    // f0 l0
    mulm(u, l1, fim);    // f1 l1
    ld<L, ES>(a, fre);
    ld<L, ES>(v, fim);
    P::addn(a, a, v);
    st<L, ES>(fre, a);       // f0 + f1 (f0 itself is dead)
    P::addn(l0, l0, l1);
    mulm(v, l0, fre);    // (f0 + f1)(l0 + l1)
    P::subk(a, t, u, c_fc.p2, 2);
    st<L, ES>(fre, a);       // f0 l0 - f1 l1
    P::addn(t, t, u);
    P::subk(v, v, t, c_fc.p4, 4);
    st<L, ES>(fim, v);       // f0 l1 + f1 l0
  }
  template <int KM>
  BGN_DEVNI static void line_mul_lazy_n(E fre, E fim, const uint32_t* cR, const uint32_t* aR, const uint32_t* bI,
                                        const uint32_t* uB, const uint32_t* vB) {
    R a, b, c, l0, l1;
    uint32_t T0[2 * L], T1[2 * L], S[2 * L];
    ld<L, ES>(a, vB);
    ld<L, ES>(b, uB);
    P::template dot2_stream<ES>(l0, a, cR, b, aR);  // l0 = cR vB + aR uB
    ld<L, ES>(l1, bI);                              // l1 = bI
    mulwk<KM>(T0, l0, fre);   // f0 l0
    mulwk<KM>(T1, l1, fim);   // f1 l1
    P::addw(S, T0, T1);
    P::subw_k(T0, T0, T1, c_fc.p, 1);
    P::redc(a, T0);           // re; stays in registers while the f.re slot feeds the last product
    ld<L, ES>(b, fre);
    ld<L, ES>(c, fim);
    P::addn(b, b, c);
    st<L, ES>(fre, b);            // f0 + f1
    P::addn(l0, l0, l1);
    mulwk<KM>(T1, l0, fre);   // (f0 + f1)(l0 + l1)
    P::subw(T1, T1, S);       // = f0 l1 + f1 l0 >= 0
    st<L, ES>(fre, a);
    P::redc(b, T1);
    st<L, ES>(fim, b);
  }
  // ---- the line of a RECORDED table (fixed first pairing argument: MillerFixed, pairing.cuh).  The table is
  // normalised once per key, every line divided by its third coefficient (cRn = cR / bI, aRn = aR / bI):
  //     l / bI = (cRn + aRn xB) + yB i
  // one product for the real part, none for the imaginary part: 7L^2 + 3L = 2074 products per line at
  // L = 17 with lazy reduction (line_mul_lazy_f), 4 (2L^2 + L) without (line_mul_f).
  // in: f.re, f.im < 8p; cRn, aRn < 2p; xB, yB < 8p.   out: as line_mul / line_mul_lazy.
  BGN_DEVNI static void line_mul_f(E fre, E fim, const uint32_t* cRn, const uint32_t* aRn, const uint32_t* xB,
                                   const uint32_t* yB) {
    R a, l0, l1, t, u, v;
    ld<L, ES>(a, xB);
    mulm(l0, a, aRn);
    ld<L, ES>(a, cRn);
    P::addn(l0, l0, a);  // l0 = cRn + aRn xB
    ld<L, ES>(l1, yB);   // l1 = yB
    mulm(t, l0, fre);    // f0 l0
    mulm(u, l1, fim);    // f1 l1
    ld<L, ES>(a, fre);
    ld<L, ES>(v, fim);
    P::addn(a, a, v);
    st<L, ES>(fre, a);       // f0 + f1 (f0 itself is dead)
    P::addn(l0, l0, l1);
    mulm(v, l0, fre);    // (f0 + f1)(l0 + l1)
    P::subk(a, t, u, c_fc.p2, 2);
    st<L, ES>(fre, a);       // f0 l0 - f1 l1
    P::addn(t, t, u);
    P::subk(v, v, t, c_fc.p4, 4);
    st<L, ES>(fim, v);       // f0 l1 + f1 l0
  }
  template <int KM>
  BGN_DEVNI static void line_mul_lazy_f(E fre, E fim, const uint32_t* cRn, const uint32_t* aRn, const uint32_t* xB,
                                        const uint32_t* yB) {
    R a, b, c, l0, l1;
    uint32_t T0[2 * L], T1[2 * L], S[2 * L];
    ld<L, ES>(a, xB);
    mulm(l0, a, aRn);
    ld<L, ES>(a, cRn);
    P::addn(l0, l0, a);       // l0 = cRn + aRn xB
    ld<L, ES>(l1, yB);        // l1 = yB
    mulwk<KM>(T0, l0, fre);   // f0 l0
    mulwk<KM>(T1, l1, fim);   // f1 l1
    P::addw(S, T0, T1);
    P::subw_k(T0, T0, T1, c_fc.p, 1);
    P::redc(a, T0);
    ld<L, ES>(b, fre);
    ld<L, ES>(c, fim);
    P::addn(b, b, c);
    st<L, ES>(fre, b);            // f0 + f1
    P::addn(l0, l0, l1);
    mulwk<KM>(T1, l0, fre);   // (f0 + f1)(l0 + l1)
    P::subw(T1, T1, S);       // = f0 l1 + f1 l0 >= 0
    st<L, ES>(fre, a);
    P::redc(b, T1);
    st<L, ES>(fim, b);
  }
  // ---- the parabola of a doubling-and-addition step (curve.cuh: G::dadd_para) at a normalised evaluation
  // point: g / yB = (cs wB + c1 uB + c0 vB) + ci i with wB = xB^2 / yB, uB = xB / yB, vB = 1 / yB -- ONE
  // three-term dot product (arith.cuh: dot3_stream, 4L^2 + L) and ONE F_p^2 product where the tangent and the
  // chord of the two separate steps cost two dot products and two F_p^2 products: 9L^2 + 3L = 2652 products
  // per evaluation point instead of 2 x 2363 at L = 17 (para_mul_lazy), 10L^2 + 4L instead of 2 x (9L^2 + 4L)
  // without lazy reduction (para_mul).
  // in: f.re, f.im < 8p; cs, c1, c0, ci < 2p; uB, vB, wB < 2p.   out: as line_mul / line_mul_lazy.
  BGN_DEVNI static void para_mul(E fre, E fim, const uint32_t* cs, const uint32_t* c1, const uint32_t* c0,
                                 const uint32_t* ci, const uint32_t* uB, const uint32_t* vB, const uint32_t* wB) {
    R a, l0, l1, t, u, v;
    ld<L, 1>(a, wB);   // wB lives in a global array (unit stride)
    ld<L, ES>(u, uB);
    ld<L, ES>(v, vB);
    P::template dot3_stream<ES>(l0, a, cs, u, c1, v, c0);
    ld<L, ES>(l1, ci);
    mulm(t, l0, fre);    // f0 l0
    mulm(u, l1, fim);    // f1 l1
    ld<L, ES>(a, fre);
    ld<L, ES>(v, fim);
    P::addn(a, a, v);
    st<L, ES>(fre, a);       // f0 + f1 (f0 itself is dead)
    P::addn(l0, l0, l1);
    mulm(v, l0, fre);    // (f0 + f1)(l0 + l1)
    P::subk(a, t, u, c_fc.p2, 2);
    st<L, ES>(fre, a);       // f0 l0 - f1 l1
    P::addn(t, t, u);
    P::subk(v, v, t, c_fc.p4, 4);
    st<L, ES>(fim, v);       // f0 l1 + f1 l0
  }
  template <int KM>
  BGN_DEVNI static void para_mul_lazy(E fre, E fim, const uint32_t* cs, const uint32_t* c1, const uint32_t* c0,
                                      const uint32_t* ci, const uint32_t* uB, const uint32_t* vB, const uint32_t* wB) {
    R a, b, c, l0, l1;
    uint32_t T0[2 * L], T1[2 * L], S[2 * L];
    ld<L, 1>(a, wB);
    ld<L, ES>(b, uB);
    ld<L, ES>(c, vB);
    P::template dot3_stream<ES>(l0, a, cs, b, c1, c, c0);
    ld<L, ES>(l1, ci);
    mulwk<KM>(T0, l0, fre);   // f0 l0
    mulwk<KM>(T1, l1, fim);   // f1 l1
    P::addw(S, T0, T1);
    P::subw_k(T0, T0, T1, c_fc.p, 1);
    P::redc(a, T0);
    ld<L, ES>(b, fre);
    ld<L, ES>(c, fim);
    P::addn(b, b, c);
    st<L, ES>(fre, b);            // f0 + f1
    P::addn(l0, l0, l1);
    mulwk<KM>(T1, l0, fre);   // (f0 + f1)(l0 + l1)
    P::subw(T1, T1, S);       // = f0 l1 + f1 l0 >= 0
    st<L, ES>(fre, a);
    P::redc(b, T1);
    st<L, ES>(fim, b);
  }
  // dst (unit stride, e.g. a global array) <- a (unit stride) * b (slot)
  BGN_DEVNI static void mul_to_unit(uint32_t* dst, const uint32_t* a, const uint32_t* b) {
    R x, y;
    ld<L, 1>(x, a);
    mulm(y, x, b);
    st<L, 1>(dst, y);
  }
  // bring one slot back to [0, 2p) (entry of the three-address parabola step).  in: < 8p.
  BGN_DEVNI static void norm1(E s) {
    R a;
    ld<L, ES>(a, s);
    P::norm2p(a, a);
    st<L, ES>(s, a);
  }
  // ---- the parabola of a RECORDED table (MillerFixed::record normalises it by its imaginary coefficient):
  //     g / ci = ((csn xB + c1n) xB + c0n) + yB i
  // two products for the real part, none for the imaginary part, ONE F_p^2 product: 9L^2 + 4L = 2669 products
  // where the tangent and the chord of the two separate steps cost 2 x 2074 (para_mul_lazy_f), 5 (2L^2 + L)
  // against 8 without lazy reduction (para_mul_f).
  // in: f.re, f.im < 8p; csn, c1n, c0n < 2p; xB, yB < 8p.   out: as line_mul / line_mul_lazy.
  BGN_DEV static void para_eval_f(uint32_t (&l0)[L], const uint32_t* csn, const uint32_t* c1n, const uint32_t* c0n,
                                  const uint32_t* xB) {
    R a, t;
    ld<L, ES>(a, xB);
    mulm(t, a, csn);
    ld<L, ES>(a, c1n);
    P::addn(t, t, a);     // csn xB + c1n
    mulm(l0, t, xB);
    ld<L, ES>(a, c0n);
    P::addn(l0, l0, a);   // (csn xB + c1n) xB + c0n
  }
  BGN_DEVNI static void para_mul_f(E fre, E fim, const uint32_t* csn, const uint32_t* c1n, const uint32_t* c0n,
                                   const uint32_t* xB, const uint32_t* yB) {
    R a, l0, l1, t, u, v;
    para_eval_f(l0, csn, c1n, c0n, xB);
    ld<L, ES>(l1, yB);
    mulm(t, l0, fre);    // f0 l0
    mulm(u, l1, fim);    // f1 l1
    ld<L, ES>(a, fre);
    ld<L, ES>(v, fim);
    P::addn(a, a, v);
    st<L, ES>(fre, a);
    P::addn(l0, l0, l1);
    mulm(v, l0, fre);    // (f0 + f1)(l0 + l1)
    P::subk(a, t, u, c_fc.p2, 2);
    st<L, ES>(fre, a);
    P::addn(t, t, u);
    P::subk(v, v, t, c_fc.p4, 4);
    st<L, ES>(fim, v);
  }
  template <int KM>
  BGN_DEVNI static void para_mul_lazy_f(E fre, E fim, const uint32_t* csn, const uint32_t* c1n, const uint32_t* c0n,
                                        const uint32_t* xB, const uint32_t* yB) {
    R a, b, c, l0, l1;
    uint32_t T0[2 * L], T1[2 * L], S[2 * L];
    para_eval_f(l0, csn, c1n, c0n, xB);
    ld<L, ES>(l1, yB);
    mulwk<KM>(T0, l0, fre);   // f0 l0
    mulwk<KM>(T1, l1, fim);   // f1 l1
    P::addw(S, T0, T1);
    P::subw_k(T0, T0, T1, c_fc.p, 1);
    P::redc(a, T0);
    ld<L, ES>(b, fre);
    ld<L, ES>(c, fim);
    P::addn(b, b, c);
    st<L, ES>(fre, b);
    P::addn(l0, l0, l1);
    mulwk<KM>(T1, l0, fre);
    P::subw(T1, T1, S);
    st<L, ES>(fre, a);
    P::redc(b, T1);
    st<L, ES>(fim, b);
  }
  // (xB, yB) -> (uB, vB) = (xB / yB, 1 / yB) in place: the inversion is the binary GCD of arith.cuh on the
  // ALU pipe (F::inv_gcd), once per evaluation point and pairing batch.  in: canonical.  out: < 2p.
  BGN_DEVNI static void eval_normalise(E xB, E yB) {
    F<L>::template inv_gcd<true, ES>(yB, yB);
    R x, y;
    ld<L, ES>(x, xB);
    mulm(y, x, yB);
    st<L, ES>(xB, y);
  }

  // A/B candidate (tools/primbench.py mode 83): line_mul_lazy with the two independent double-width
  // products f0 l0 and f1 l1 interleaved row by row (Fp::mulw2), and the two evaluation products likewise
  BGN_DEVNI static void line_mul_lazy_il(E fre, E fim, const uint32_t* cR, const uint32_t* aR, const uint32_t* bI,
                                         const uint32_t* xB, const uint32_t* yB) {
    R a, b, c, l0, l1;
    uint32_t T0[2 * L], T1[2 * L], S[2 * L];
    ld<L, ES>(a, xB);
    ld<L, ES>(b, yB);
    P::mul_pair(l0, a, aR, l1, b, bI);  // aR xB | bI yB
    ld<L, ES>(a, cR);
    P::addn(l0, l0, a);       // l0 = cR + aR xB
    P::template mulw2<ES>(T0, l0, fre, T1, l1, fim);   // f0 l0 | f1 l1
    P::addw(S, T0, T1);
    P::subw_k(T0, T0, T1, c_fc.p, 1);
    P::redc(a, T0);
    ld<L, ES>(b, fre);
    ld<L, ES>(c, fim);
    P::addn(b, b, c);
    st<L, ES>(fre, b);            // f0 + f1
    P::addn(l0, l0, l1);
    P::template mulw<ES>(T1, l0, fre);   // (f0 + f1)(l0 + l1)
    P::subw(T1, T1, S);
    st<L, ES>(fre, a);
    P::redc(b, T1);
    st<L, ES>(fim, b);
  }

  // f <- f^2 = (f0 + f1)(f0 - f1) + 2 f0 f1 i, 2 products.  in: < 8p.  out: < 4p.
  BGN_DEVNI static void sqr2(E fre, E fim) {
    R a, b, s, d, m;
    ld<L, ES>(a, fre);
    ld<L, ES>(b, fim);
    P::addn(s, a, b);
    P::subk(d, a, b, c_fc.p8, 8);
    mulm(m, a, fim);     // f0 f1
    st<L, ES>(fre, d);
    mulm(a, s, fre);     // (f0 + f1)(f0 - f1)
    st<L, ES>(fre, a);
    dbl(m, m);
    st<L, ES>(fim, m);
  }

  // (X, Y, Z) <- 2 (X, Y, Z) on y^2 = x^3 + x (Jacobian) and the tangent at the old point:
  // cR = M X - 2 YY, aR = M ZZ, bI = Z3 ZZ with M = 3 XX + ZZ^2.  12 products.
  // in: X, Y, Z < 9p.   out: X, Y < 6p, Z < 3p, cR < 6p, aR, bI < 2p.
  // The six slots double as scratch for the multipliers, in an order that never overwrites a
  // value still to be read.
  BGN_DEVNI static void dbl_line(E X, E Y, E Z, E cR, E aR, E bI) {
    R x, w, xx, yy, zz, m, s;
    ld<L, ES>(x, X);
    sqrm(xx, x, X);            // XX
    ld<L, ES>(w, Y);
    sqrm(yy, w, Y);            // YY
    dbl(w, w);                 // 2Y
    ld<L, ES>(s, Z);
    sqrm(zz, s, Z);            // ZZ
    st<L, ES>(bI, zz);
    sqrm(m, zz, bI);           // ZZ^2
    P::addn(m, m, xx);
    dbl(xx, xx);
    P::addn(m, m, xx);         // M = 3 XX + ZZ^2
    mulm(s, w, Z);             // Z3 = 2Y * Z
    st<L, ES>(Z, s);
    mulm(w, m, bI);            // aR = M ZZ   (bI still holds ZZ)
    st<L, ES>(aR, w);
    mulm(w, s, bI);            // bI = Z3 ZZ
    st<L, ES>(bI, w);
    dbl(yy, yy);               // 2 YY
    st<L, ES>(cR, yy);
    mulm(w, m, X);             // M X
    P::subk(w, w, yy, c_fc.p4, 4);  // cR = M X - 2 YY   (kept in registers until the slot is free)
    dbl(x, x);
    mulm(s, x, cR);            // S = 2X * 2YY = 4 X YY
    sqrm(zz, yy, cR);          // 4 YY^2
    st<L, ES>(cR, w);
    st<L, ES>(Y, m);
    sqrm(xx, m, Y);            // M^2
    dbl(w, s);
    P::subk(xx, xx, w, c_fc.p4, 4);  // X3 = M^2 - 2S
    st<L, ES>(X, xx);
    P::subk(s, s, xx, c_fc.p8, 8);   // S - X3
    mulm(w, s, Y);             // M (S - X3)
    dbl(zz, zz);               // 8 YY^2
    P::subk(w, w, zz, c_fc.p4, 4);
    st<L, ES>(Y, w);               // Y3
  }

  // (X, Y, Z) <- (X, Y, Z) + (xA, +-yA) (mixed) and the chord through them: aR = r = 2 (S2 - Y),
  // bI = Z3, cR = r xA - (+-yA) Z3.  13 products.  No special cases: the Miller loop never meets
  // them for points of order n except at the very last step, which the schedule drops.
  // in: X, Y, Z < 9p; xA, yA < 2p.   out: X, Y < 9p, Z, bI, cR < 4p, aR < 24p.
  BGN_DEVNI static void madd_line(E X, E Y, E Z, const uint32_t* xA, const uint32_t* yA, bool negate, E cR, E aR,
                                  E bI) {
    R xa, ya, z, h, r, w, v, c;
    ld<L, 1>(xa, xA);  // the affine point lives in the batch arrays (unit stride)
    ld<L, 1>(ya, yA);
    if (negate) P::negk(ya, ya, c_fc.p2, 2);  // 2p - yA = -yA
    ld<L, ES>(z, Z);
    sqrm(w, z, Z);             // ZZ
    st<L, ES>(bI, w);
    mulm(h, xa, bI);           // U2 = xA ZZ
    ld<L, ES>(w, X);
    P::subk(h, h, w, c_fc.p16, 16);  // H = U2 - X
    mulm(w, z, bI);            // Z ZZ
    st<L, ES>(aR, w);
    mulm(r, ya, aR);           // S2 = yA Z^3
    ld<L, ES>(w, Y);
    P::subk(r, r, w, c_fc.p16, 16);
    dbl(r, r);                 // r = 2 (S2 - Y)
    st<L, ES>(aR, r);              // aR = r
    dbl(w, h);
    st<L, ES>(cR, w);              // 2H
    mulm(v, z, cR);            // Z3 = Z * 2H
    st<L, ES>(Z, v);
    st<L, ES>(bI, v);              // bI = Z3
    mulm(c, xa, aR);           // r xA
    mulm(w, ya, Z);            // yA Z3
    P::subk(c, c, w, c_fc.p2, 2);   // cR, kept in registers until the slot is free
    ld<L, ES>(w, cR);
    sqrm(v, w, cR);            // I = (2H)^2
    ld<L, ES>(z, X);               // z now holds X
    st<L, ES>(X, v);
    mulm(w, h, X);             // J = H I
    mulm(v, z, X);             // V = X I
    st<L, ES>(cR, c);
    sqrm(c, r, aR);            // r^2
    dbl(z, v);
    P::addn(z, z, w);          // J + 2V
    P::subk(c, c, z, c_fc.p4, 4);   // X3 = r^2 - J - 2V
    st<L, ES>(X, c);
    mulm(h, w, Y);             // Y J
    P::subk(v, v, c, c_fc.p16, 16); // V - X3
    st<L, ES>(Y, v);
    mulm(w, r, Y);             // r (V - X3)
    dbl(h, h);
    P::subk(w, w, h, c_fc.p4, 4);
    st<L, ES>(Y, w);               // Y3 = r (V - X3) - 2 Y J
  }

  // ---- the doubling-and-addition step T <- (T + A) + T with its parabola (curve.cuh: G::dadd_para has the
  // formulas), fused: 22 products, 6 squarings and one dot product, every product (register) x (memory), the
  // seven slots X, Y, Z, cs, c1, c0, ci doubling as the multipliers' storage in an order that never overwrites
  // a value still to be read.
  // in: X, Y, Z < 9p; xA, yA < 2p.   out: X, Y < 9p, Z < 3p; cs, c1, c0, ci < 8p.
  BGN_DEVNI static void dadd_para(E X, E Y, E Z, const uint32_t* xA, const uint32_t* yA, bool negate, E cs, E c1, E c0,
                                  E ci) {
    R a, b, h, n1, zr, u, w, xr, n2, an, sn, t;
    ld<L, ES>(a, Z);
    sqrm(b, a, Z);               // ZZ
    st<L, ES>(cs, b);
    mulm(t, a, cs);              // Z^3
    st<L, ES>(c1, t);
    ld<L, 1>(a, xA);
    mulm(h, a, cs);              // xA ZZ
    ld<L, ES>(b, X);
    P::subk(h, h, b, c_fc.p16, 16);  // H = xA ZZ - X
    ld<L, 1>(a, yA);
    if (negate) P::negk(a, a, c_fc.p2, 2);
    mulm(n1, a, c1);             // yA Z^3
    ld<L, ES>(a, Y);
    P::subk(n1, n1, a, c_fc.p16, 16);  // N1
    mulm(zr, h, Z);              // ZR = Z H
    st<L, ES>(c0, h);
    sqrm(a, h, c0);              // H^2
    st<L, ES>(ci, a);
    mulm(t, h, ci);              // H^3
    mulm(u, b, ci);              // U = X H^2   (b still holds X)
    st<L, ES>(c1, t);            // H^3 (Z^3 is dead)
    ld<L, ES>(a, Y);
    mulm(w, a, c1);              // W = Y H^3
    st<L, ES>(cs, n1);           // N1 (ZZ is dead)
    sqrm(xr, n1, cs);            // N1^2
    P::subk(xr, xr, t, c_fc.p4, 4);
    dbl(a, u);
    P::subk(xr, xr, a, c_fc.p8, 8);    // XR = N1^2 - H^3 - 2U
    P::subk(a, u, xr, c_fc.p16, 16);   // U - XR
    mulm(t, a, cs);              // N1 (U - XR)
    P::subk(t, t, w, c_fc.p4, 4);      // YR
    P::subk(h, xr, u, c_fc.p4, 4);     // H2 = XR - U   (H is dead)
    P::subk(n2, t, w, c_fc.p4, 4);     // N2 = YR - W
    st<L, ES>(c0, h);            // H2
    st<L, ES>(ci, n2);           // N2
    P::template dot2_stream<ES>(an, xr, c0, n1, ci);   // An = XR H2 + N1 N2
    mulm(sn, n1, c0);            // N1 H2
    P::addn(sn, sn, n2);         // Sn
    sqrm(a, h, c0);              // A2 = H2^2
    st<L, ES>(c1, a);            // A2 (H^3 is dead)
    mulm(b, u, c1);              // B2 = U A2
    mulm(t, xr, c1);             // C2 = XR A2
    sqrm(a, n2, ci);             // N2^2
    P::subk(a, a, b, c_fc.p4, 4);
    P::subk(a, a, t, c_fc.p4, 4);      // XS
    st<L, ES>(X, a);
    P::subk(t, t, b, c_fc.p4, 4);      // C2 - B2
    st<L, ES>(c1, t);
    P::subk(b, b, a, c_fc.p16, 16);    // B2 - XS
    mulm(a, b, ci);              // N2 (B2 - XS)
    mulm(t, w, c1);              // W (C2 - B2)
    P::subk(a, a, t, c_fc.p4, 4);      // YS
    st<L, ES>(Y, a);
    mulm(t, zr, c0);             // ZS = ZR H2
    st<L, ES>(Z, t);
    mulm(a, zr, Z);              // Dn = ZR ZS
    st<L, ES>(ci, a);            // Dn (N2 is dead)
    sqrm(n2, a, ci);             // cs = Dn^2            (kept in registers until its slot is free)
    mulm(b, an, ci);             // An Dn
    P::negk(b, b, c_fc.p4, 4);   // c1 = -An Dn
    mulm(t, sn, ci);             // Sn Dn
    st<L, ES>(c1, zr);           // ZR (C2 - B2 is dead)
    mulm(a, t, c1);              // Sn Dn ZR
    P::negk(a, a, c_fc.p4, 4);   // ci = -Sn Dn ZR
    st<L, ES>(ci, a);
    st<L, ES>(cs, w);            // W (N1 is dead)
    mulm(t, sn, cs);             // Sn W
    mulm(a, u, c0);              // U H2
    P::addn(a, a, an);           // U H2 + An
    st<L, ES>(c1, a);
    mulm(zr, u, c1);             // U (U H2 + An)
    P::subk(t, t, zr, c_fc.p4, 4);
    mulm(a, t, c0);              // c0 = H2 (Sn W - U (U H2 + An))
    st<L, ES>(c0, a);
    st<L, ES>(c1, b);
    st<L, ES>(cs, n2);
  }

  // ---- final exponentiation f^((p^2-1)/n) = (conj(f)/f)^l = (conj(f)^2 / N(f))^l, fused.
  // Step 1: f <- conj(f)^2 = (f0^2 - f1^2) - 2 f0 f1 i and nrm <- N(f) = f0^2 + f1^2.  3 products.
  // in: f < 8p.  out: f.re < 4p, f.im < 4p, nrm < 4p.
  BGN_DEVNI static void fe_prepare(E fre, E fim, E nrm) {
    R a, b, t, u, v;
    ld<L, ES>(a, fre);
    ld<L, ES>(b, fim);
    sqrm(t, a, fre);   // f0^2
    sqrm(u, b, fim);   // f1^2
    mulm(v, a, fim);   // f0 f1
    P::addn(a, t, u);
    st<L, ES>(nrm, a);
    P::subk(t, t, u, c_fc.p2, 2);
    st<L, ES>(fre, t);
    dbl(v, v);
    P::negk(v, v, c_fc.p4, 4);
    st<L, ES>(fim, v);
  }
  // r <- a^(p-2) (Fermat inverse; 0 -> 0) with the running power kept in registers: bits(p) - 1
  // squarings and wt(p-2) - 1 products, no loads or stores in between.  The exponent is a key
  // constant, so control flow is uniform.  r may alias a.  in: a < 16p.  out: r < 2p.
  BGN_DEVNI static void fp_inv(E r, const uint32_t* a) {
    R x, y;
    ld<L, ES>(x, a);
    ld<L, ES>(y, a);
    int top = 32 * L - 1;
    while (top > 0 && !((c_fc.p[top >> 5] >> (top & 31)) & 1)) top--;
    // exponent e = p - 2: p = 3 (mod 4), so subtracting 2 only clears bit 1 (no borrow)
    BGN_UNROLL1
    for (int bit = top - 1; bit >= 0; bit--) {
      uint32_t limb = c_fc.p[bit >> 5];
      if ((bit >> 5) == 0) limb -= 2;
      P::mul(y, y, y);
      if ((limb >> (bit & 31)) & 1) P::mul(y, y, x);
    }
    st<L, ES>(r, y);
  }
  // r <- a * b in F_p (register operand a, memory operand b); r may alias either
  BGN_DEVNI static void fp_mul(E r, const uint32_t* a, const uint32_t* b) {
    R x, y;
    ld<L, ES>(x, a);
    mulm(y, x, b);
    st<L, ES>(r, y);
  }
  // f <- f * s for s in F_p.  2 products.
  BGN_DEVNI static void scale2(E fre, E fim, const uint32_t* s) {
    R x, y;
    ld<L, ES>(x, s);
    mulm(y, x, fre);
    st<L, ES>(fre, y);
    mulm(y, x, fim);
    st<L, ES>(fim, y);
  }
  // f <- f * g in F_p^2 (Karatsuba, 3 products).  in: f, g < 8p.  out: f.re < 4p, f.im < 6p.
  BGN_DEVNI static void mul2(E fre, E fim, const uint32_t* gre, const uint32_t* gim) {
    R a, b, t, u, v;
    ld<L, ES>(a, fre);
    ld<L, ES>(b, fim);
    mulm(t, a, gre);   // f0 g0
    mulm(u, b, gim);   // f1 g1
    P::addn(a, a, b);
    ld<L, ES>(b, gre);
    ld<L, ES>(v, gim);
    P::addn(b, b, v);
    st<L, ES>(fre, b);     // g0 + g1 (f0 is dead)
    mulm(v, a, fre);   // (f0 + f1)(g0 + g1)
    P::subk(a, t, u, c_fc.p2, 2);
    st<L, ES>(fre, a);
    P::addn(t, t, u);
    P::subk(v, v, t, c_fc.p4, 4);
    st<L, ES>(fim, v);
  }
  // bring both coordinates back to [0, 2p) (what the GT kernels downstream expect).  in: < 8p.
  BGN_DEVNI static void norm2(E fre, E fim) {
    R a;
    ld<L, ES>(a, fre);
    P::norm2p(a, a);
    st<L, ES>(fre, a);
    ld<L, ES>(a, fim);
    P::norm2p(a, a);
    st<L, ES>(fim, a);
  }
};
