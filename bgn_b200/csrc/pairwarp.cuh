// pairwarp.cuh -- a general pairing e(A, B) split over TWO WARPS, for batches too small to give
// every scheduler two warps of one-thread-per-pairing work (Mult, bgn.go:294-314, on few operands).
//
// Why a different split than pairlane.cuh: the general Miller loop is point arithmetic (which only
// ever touches A's multiples) feeding an accumulator update (which only ever touches f and B).  These
// are two different instruction streams, so they cannot share a warp without divergence -- but they
// can be two warps of one block that meet once per Miller step:
//
//   X warp, lane i    advances the Miller point of pairing i: doubling 9 products, mixed addition 11
//                     -- the bare point arithmetic, no line coefficients -- and PUBLISHES what the line
//                     needs (doubling: M, ZZ, Z3, X, 2YY; addition: r, Z3) in shared memory;
//   F warp, lane i    one step behind, evaluates the line at B directly from the published values
//                     (l0 = M (X + ZZ xB) - 2YY, l1 = Z3 (ZZ yB): 4 products where forming the
//                     coefficients first and evaluating them costs 5; addition: l0 = r (xA + xB) - yA Z3,
//                     l1 = Z3 yB: 3 instead of 4) and folds it into its accumulator f <- f^2 * l
//                     (fused.cuh: sqr2 and the lazily reduced F_p^2 product).
//
// The published values are double-buffered, so ONE named barrier per step and pair of warps is the
// only synchronisation: at barrier s the X warp has published step s and the F warp has consumed step
// s - 1.  Work per doubling step at L = 17: X 5 355, F 5 049 products (one thread: 10 999); per
// addition step X 6 545, F 3 264 (10 404): the critical path is 0.52 of the one-thread kernel's and
// the total 5.5 % smaller.  The final exponentiation is the F warp's.
//
// Replaces the same libpbc behaviour as pairing.cuh (a1_pairing via Element.Pair, bgn.go:300).
#pragma once
#include "pairing.cuh"

template <int L, int U>
struct MDuo : MF<L, U, 1> {
  typedef MF<L, U, 1> B;
  typedef Fp<L> P;
  typedef uint32_t R[L];
  using B::dbl;
  using B::mulm;
  // the dedicated squaring also with looped products (fused.cuh: sqrm uses it only when U == 0): the
  // X warp's doubling is 6 squarings of 9 products and the X warp is the critical path; measured by A/B
#ifndef BGN_DUO_SQR
#define BGN_DUO_SQR 1
#endif
  BGN_DEV static void sqrm(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t* mem) {
    if (BGN_DUO_SQR && L <= 17)
      P::sqr(r, a);
    else
      mulm(r, a, mem);
  }

  // ---- X warp.  (X, Y, Z) <- 2 (X, Y, Z), 9 products; publishes M = 3 XX + ZZ^2, ZZ, Z3 = 2 Y Z,
  // the old X and 2 YY.   in: X, Y, Z < 9p.   out: X, Y < 6p, Z < 3p; oM < 6p, oZZ < 2p, oZ3 < 3p,
  // oX < 9p, oYY2 < 3p.
  BGN_DEVNI static void dbl_pub(E X, E Y, E Z, E oM, E oZZ, E oZ3, E oX, E oYY2) {
    R x, w, xx, yy, zz, m, s;
    ld<L>(x, X);
    st<L>(oX, x);
    sqrm(xx, x, X);            // XX
    ld<L>(w, Y);
    sqrm(yy, w, Y);            // YY
    dbl(w, w);                 // 2Y
    ld<L>(s, Z);
    sqrm(zz, s, Z);            // ZZ
    st<L>(oZZ, zz);
    sqrm(m, zz, oZZ);          // ZZ^2
    P::addn(m, m, xx);
    dbl(xx, xx);
    P::addn(m, m, xx);         // M = 3 XX + ZZ^2
    st<L>(oM, m);
    mulm(s, w, Z);             // Z3 = 2Y * Z
    st<L>(Z, s);
    st<L>(oZ3, s);
    dbl(yy, yy);               // 2 YY
    st<L>(oYY2, yy);
    dbl(x, x);                 // 2X
    mulm(s, x, oYY2);          // S = 2X * 2YY
    sqrm(zz, yy, oYY2);        // 4 YY^2
    sqrm(xx, m, oM);           // M^2
    dbl(w, s);
    P::subk(xx, xx, w, c_fc.p4, 4);  // X3 = M^2 - 2S
    st<L>(X, xx);
    P::subk(s, s, xx, c_fc.p8, 8);   // S - X3
    mulm(w, s, oM);            // M (S - X3)
    dbl(zz, zz);               // 8 YY^2
    P::subk(w, w, zz, c_fc.p4, 4);
    st<L>(Y, w);               // Y3
  }
  // (X, Y, Z) <- (X, Y, Z) + (xA, +-yA), 11 products; publishes r = 2 (S2 - Y) and Z3; t0, t1 are
  // scratch slots.  No special cases, as madd_line (the same sequence without its two line products).
  // in: X, Y, Z < 9p; xA, yA < 2p.   out: X, Y < 9p, Z < 4p; oR < 24p, oZ3 < 4p.
  BGN_DEVNI static void madd_pub(E X, E Y, E Z, const uint32_t* xA, const uint32_t* yA, bool negate, E oR, E oZ3, E t0,
                                 E t1) {
    R xa, ya, z, h, r, w, v, c;
    ld<L>(xa, xA);
    ld<L>(ya, yA);
    if (negate) P::negk(ya, ya, c_fc.p2, 2);
    ld<L>(z, Z);
    sqrm(w, z, Z);             // ZZ
    st<L>(t0, w);
    mulm(h, xa, t0);           // U2 = xA ZZ
    ld<L>(w, X);
    P::subk(h, h, w, c_fc.p16, 16);  // H = U2 - X
    mulm(w, z, t0);            // Z ZZ
    st<L>(t1, w);
    mulm(r, ya, t1);           // S2 = yA Z^3
    ld<L>(w, Y);
    P::subk(r, r, w, c_fc.p16, 16);
    dbl(r, r);                 // r = 2 (S2 - Y)
    st<L>(oR, r);
    dbl(w, h);                 // 2H
    st<L>(t0, w);
    mulm(v, z, t0);            // Z3 = Z * 2H
    st<L>(Z, v);
    st<L>(oZ3, v);
    sqrm(v, w, t0);            // I = (2H)^2
    ld<L>(z, X);               // z <- X
    st<L>(t1, v);
    mulm(w, h, t1);            // J = H I
    mulm(v, z, t1);            // V = X I
    sqrm(c, r, oR);            // r^2
    dbl(z, v);
    P::addn(z, z, w);          // J + 2V
    P::subk(c, c, z, c_fc.p4, 4);    // X3 = r^2 - J - 2V
    st<L>(X, c);
    mulm(h, w, Y);             // Y J
    P::subk(v, v, c, c_fc.p16, 16);  // V - X3
    st<L>(t0, v);
    mulm(w, r, t0);            // r (V - X3)
    dbl(h, h);
    P::subk(w, w, h, c_fc.p4, 4);
    st<L>(Y, w);               // Y3 = r (V - X3) - 2 Y J
  }

  // ---- F warp.  f <- f * (l0 + l1 i) with l0, l1 in registers (clobbered).
  // in: f < 8p, l0 < 8p, l1 < 8p.   out (lazy): f.re < 3p, f.im < 2p; (plain): f.re < 4p, f.im < 7p.
  BGN_DEV static void fl_mul(E fre, E fim, uint32_t (&l0)[L], uint32_t (&l1)[L]) {
#if BGN_LINE_LAZY
    R a, b, c;
    uint32_t T0[2 * L], T1[2 * L], S[2 * L];
    P::template mulw<1>(T0, l0, fre);  // f0 l0
    P::template mulw<1>(T1, l1, fim);  // f1 l1
    P::addw(S, T0, T1);
    P::subw_k(T0, T0, T1, c_fc.p, 1);
    P::redc(a, T0);                    // re
    ld<L>(b, fre);
    ld<L>(c, fim);
    P::addn(b, b, c);
    st<L>(fre, b);                     // f0 + f1
    P::addn(l0, l0, l1);
    P::template mulw<1>(T1, l0, fre);  // (f0 + f1)(l0 + l1)
    P::subw(T1, T1, S);                // = f0 l1 + f1 l0 >= 0
    st<L>(fre, a);
    P::redc(b, T1);
    st<L>(fim, b);
#else
    R a, t, u, v;
    mulm(t, l0, fre);    // f0 l0
    mulm(u, l1, fim);    // f1 l1
    ld<L>(a, fre);
    ld<L>(v, fim);
    P::addn(a, a, v);
    st<L>(fre, a);       // f0 + f1
    P::addn(l0, l0, l1);
    mulm(v, l0, fre);    // (f0 + f1)(l0 + l1)
    P::subk(a, t, u, c_fc.p2, 2);
    st<L>(fre, a);
    P::addn(t, t, u);
    P::subk(v, v, t, c_fc.p4, 4);
    st<L>(fim, v);
#endif
  }
  // the tangent of a doubling step at B = (xB, yB) from the published values, folded into f
  BGN_DEVNI static void fstep_dbl(E fre, E fim, const uint32_t* pM, const uint32_t* pZZ, const uint32_t* pZ3,
                                  const uint32_t* pX, const uint32_t* pYY2, const uint32_t* xB, const uint32_t* yB) {
    R a, k, l0, l1;
    ld<L>(a, xB);
    mulm(k, a, pZZ);           // ZZ xB
    ld<L>(a, pX);
    P::addn(k, k, a);          // X + ZZ xB
    mulm(l0, k, pM);           // M (X + ZZ xB)
    ld<L>(a, pYY2);
    P::subk(l0, l0, a, c_fc.p4, 4);  // l0 = M X + M ZZ xB - 2 YY
    ld<L>(a, yB);
    mulm(k, a, pZZ);           // ZZ yB
    mulm(l1, k, pZ3);          // l1 = Z3 ZZ yB
    fl_mul(fre, fim, l0, l1);
  }
  // the chord of an addition step: l0 = r (xA + xB) - (+-yA) Z3, l1 = Z3 yB
  BGN_DEVNI static void fstep_add(E fre, E fim, const uint32_t* pR, const uint32_t* pZ3, const uint32_t* xA,
                                  const uint32_t* yA, bool negate, const uint32_t* xB, const uint32_t* yB) {
    R a, b, k, l0, l1;
    ld<L>(a, xA);
    ld<L>(b, xB);
    P::addn(a, a, b);          // xA + xB
    mulm(l0, a, pR);           // r (xA + xB)
    ld<L>(b, yA);
    if (negate) P::negk(b, b, c_fc.p2, 2);
    mulm(k, b, pZ3);           // (+-yA) Z3
    P::subk(l0, l0, k, c_fc.p2, 2);
    ld<L>(a, yB);
    mulm(l1, a, pZ3);          // Z3 yB
    fl_mul(fre, fim, l0, l1);
  }
};

// U: the products' row loop (fused.cuh), 0 = fully unrolled.  Which one ships is measured
// (tools/mapping_ab.py --duo-loop): the two warps of a pair run DIFFERENT straight-line code, so the
// unrolled routines (9.5 KB per product at L = 17) put twice the pressure on the instruction cache
// that the one-thread kernel's lockstep blocks do.
template <int L, int U = BGN_MILLER_LOOP_A>
struct MillerDuo {
  typedef F<L> FF;
  typedef MDuo<L, U> MA;
  // shared-memory slots per pairing, [slot][pairing of the block][L]: the X warp's point, two buffers
  // of published values, the F warp's accumulator and evaluation point
  enum { S_X = 0, S_Y = 1, S_Z = 2, S_PUB = 3, NPUB = 5, S_FR = 13, S_FI = 14, S_EX = 15, S_EY = 16, NSLOT = 17 };
  static BGN_HD size_t smem_words(int np) { return (size_t)NSLOT * L * np + ((size_t)np + 3) / 4; }

  const PairDuoArgs& a;
  uint32_t* smem;
  int np, pt;     // pairings per block, this thread's pairing within the block
  size_t unit;
  bool live;      // a pairing of the batch with both points finite

  BGN_DEV MillerDuo(const PairDuoArgs& a_, uint32_t* smem_, int np_, int pt_, size_t unit_)
      : a(a_), smem(smem_), np(np_), pt(pt_), unit(unit_) {
    live = unit < (size_t)a.count && !a.Minf[unit] && !a.Einf[unit];
  }
  BGN_DEV E slot(int k) const { return smem + ((size_t)k * np + pt) * L; }
  BGN_DEV E pub(int buf, int k) const { return slot(S_PUB + buf * NPUB + k); }

  // number of steps and the operation of step s (the same schedule as MillerTeam::run)
  BGN_DEV static int nsteps() {
    int n = 0;
    for (int idx = 1; idx < c_pc.naf_len; idx++) n += 1 + ((c_pc.naf[idx] != 0 && idx != c_pc.naf_len - 1) ? 1 : 0);
    return n;
  }

  BGN_DEV void x_init() {
    if (!live) return;
    FF::copy(slot(S_X), a.Mx + unit * L);
    FF::copy(slot(S_Y), a.My + unit * L);
    FF::copy(slot(S_Z), c_fc.one);
  }
  BGN_DEV void x_step(int op, int buf) {
    if (!live) return;
    if (op == MOP_DBL)
      MA::dbl_pub(slot(S_X), slot(S_Y), slot(S_Z), pub(buf, 0), pub(buf, 1), pub(buf, 2), pub(buf, 3), pub(buf, 4));
    else
      MA::madd_pub(slot(S_X), slot(S_Y), slot(S_Z), a.Mx + unit * L, a.My + unit * L, op == MOP_SUB, pub(buf, 0),
                   pub(buf, 1), pub(buf, 2), pub(buf, 3));
  }
  BGN_DEV void f_init() {
    if (!live) return;
    FF::copy(slot(S_FR), c_fc.one);
    FF::set_zero(slot(S_FI));
    FF::copy(slot(S_EX), a.Ex + unit * L);
    FF::copy(slot(S_EY), a.Ey + unit * L);
  }
  BGN_DEV void f_step(int op, int buf, bool first) {
    if (!live) return;
    if (op == MOP_DBL) {
      if (!first) MA::sqr2(slot(S_FR), slot(S_FI));
      MA::fstep_dbl(slot(S_FR), slot(S_FI), pub(buf, 0), pub(buf, 1), pub(buf, 2), pub(buf, 3), pub(buf, 4), slot(S_EX),
                    slot(S_EY));
    } else {
      MA::fstep_add(slot(S_FR), slot(S_FI), pub(buf, 0), pub(buf, 1), a.Mx + unit * L, a.My + unit * L, op == MOP_SUB,
                    slot(S_EX), slot(S_EY));
    }
  }
  // final exponentiation (conj(f)^2 / N(f))^l, as MillerFixed::run; the published buffers are scratch
  BGN_DEV void f_finish() {
    if (unit >= (size_t)a.count) return;
    if (!live) {  // e(O, .) = e(., O) = 1
      FF::set_one(a.out_re + unit * L);
      FF::set_zero(a.out_im + unit * L);
      return;
    }
    E fr = slot(S_FR), fi = slot(S_FI), n0 = pub(0, 0), i0 = pub(0, 1), g0 = pub(0, 2), g1 = pub(0, 3);
    MA::fe_prepare(fr, fi, n0);
    FF::template inv_gcd_fast<true>(i0, n0);
    MA::scale2(fr, fi, i0);
    FF::copy(g0, fr);
    FF::copy(g1, fi);
    uint64_t l = c_pc.l;
    int top = 63;
    while (top > 0 && !((l >> top) & 1)) top--;
    for (int bit = top - 1; bit >= 0; bit--) {
      MA::sqr2(fr, fi);
      if ((l >> bit) & 1) MA::mul2(fr, fi, g0, g1);
    }
    MA::norm2(fr, fi);
    FF::copy(a.out_re + unit * L, fr);
    FF::copy(a.out_im + unit * L, fi);
  }

  // the operation of every step, in order: calls fn(step, op)
  template <typename Fn>
  BGN_DEV static void for_steps(Fn fn) {
    const int n = c_pc.naf_len;
    int s = 0;
    for (int idx = 1; idx < n; idx++) {
      fn(s++, (int)MOP_DBL);
      const int d = c_pc.naf[idx];
      if (d != 0 && idx != n - 1) fn(s++, d > 0 ? (int)MOP_ADD : (int)MOP_SUB);
    }
  }
};
