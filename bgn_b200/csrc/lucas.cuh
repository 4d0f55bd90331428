// lucas.cuh -- Decrypt's C^q1 as a Lucas ladder on the trace, two threads per ciphertext.
//
// Replaces element_pow_mpz on GT as reached from bgn.go:223 (csk = C^q1) followed by the table
// search of gsbs.go:54-106, for the common case that the whole message space sits in the baby-step
// table (one giant step).  Redesign:
//
//  * A level-2 ciphertext is a unitary element of F_p^2 (a pairing value: norm 1), so with
//    V_k = C^k + C^-k = 2 Re(C^k) the ladder  V_2k = V_k^2 - 2,  V_2k+1 = V_k V_k+1 - V_1  needs ONE
//    squaring and ONE product in F_p per exponent bit -- 2 modmuls instead of the 3.5 of
//    square-and-multiply in F_p^2 (2 + 3/2).
//  * The two products of a step are independent: a PAIR of adjacent lanes computes them at the same
//    time and swaps the results with one warp shuffle per limb.  A batch of 2^14 decryptions is then
//    2^15 threads (7 warps per SM instead of 3.5), each running 1/3.6 of the sequential work.
//  * The ladder yields Re(C^q1) and Re(C^(q1+1)) but not Im(C^q1).  The baby-step table is keyed by
//    the real part (gsk^m and gsk^-m share it); the sign of m follows from
//    Im(C^q1) Im(C) = Re(C^q1) Re(C) - Re(C^(q1+1)) compared with the table entry's imaginary part,
//    so no inversion is needed.  This is also the Neg(ct) retry of bgn.go:235-241 at no cost.
//
// Non-unitary input is not a ciphertext: it gets status 1 (the reference would run its generic
// exponentiation and then fail the table search).
#pragma once
#include "pairing.cuh"

BGN_DEV uint32_t bsgs_hash_re(const uint32_t* re) {
  uint32_t h = re[0] * 0x9E3779B1u ^ re[1] * 0x85EBCA77u ^ (re[2] >> 7);
  return h ^ (h >> 15);
}

template <int L>
struct Lucas {
  typedef Fp<L> P;
  struct State {
    uint32_t A[L], B[L];    // V_k, V_k+1 (relaxed range, below 6p)
    uint32_t Pc[L], two[L]; // V_1 = 2 Re(C), V_0 = 2
    bool unitary;
  };

  BGN_DEV static void sel(uint32_t (&r)[L], bool c, const uint32_t (&a)[L], const uint32_t (&b)[L]) {
#ifdef BGN_HOSTSIM
    {
      double x = BGN_GETB(a), y = BGN_GETB(b);
      BGN_SETB(r, x > y ? x : y);  // worst case of both branches
    }
#endif
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = c ? a[j] : b[j];
  }

  BGN_DEV static void init(State& st, const DecLucasArgs& a, size_t e, bool active) {
    uint32_t c0[L], c1[L], t[L], u[L], one[L];
    if (active) {
      ld<L>(c0, a.re + e * L);
      ld<L>(c1, a.im + e * L);
    } else {
      BGN_SETB(c0, 0.0);
      BGN_SETB(c1, 0.0);
      BGN_UNROLL
      for (int j = 0; j < L; j++) c0[j] = c1[j] = 0;
    }
    ld<L>(one, c_fc.one);
    P::mul(t, c0, c0);
    P::mul(u, c1, c1);
    P::add(t, t, u);
    P::canon(t, t);
    P::canon(u, one);
    st.unitary = P::eq_raw(t, u);
    P::addn(st.two, one, one);
    P::addn(st.Pc, c0, c0);
    P::mul(t, st.Pc, st.Pc);
    P::subk(st.B, t, st.two, c_fc.p4, 4);  // V_2 = V_1^2 - 2
    sel(st.A, true, st.Pc, st.Pc);         // V_1
  }

  // role s = 0 squares (V_k or V_k+1, whichever the bit doubles), role 1 forms V_k V_k+1 - V_1
  BGN_DEV static void step(uint32_t (&r)[L], const State& st, int bit, int s) {
    uint32_t x[L], y[L], c[L];
    const bool sq = s == 0;
    sel(x, sq && bit, st.B, st.A);
    sel(y, sq && !bit, st.A, st.B);
    sel(c, sq, st.two, st.Pc);
    P::mul(r, x, y);
    P::subk(r, r, c, c_fc.p4, 4);
  }
  BGN_DEV static void update(State& st, const uint32_t (&mine)[L], const uint32_t (&other)[L], int bit, int s) {
    uint32_t sqv[L], pr[L];
    sel(sqv, s == 0, mine, other);
    sel(pr, s == 0, other, mine);
    sel(st.A, bit != 0, pr, sqv);
    sel(st.B, bit != 0, sqv, pr);
  }

  // st holds (V_q1, V_q1+1): recover m with gsk^m = C^q1, |m| <= mmax
  BGN_DEV static void finish(const State& st, const DecLucasArgs& a, size_t e) {
    uint32_t x[L], xp[L], cx[L], t[L], u[L], c0[L], c1[L], one[L];
    a.out[e] = 0;
    a.status[e] = 1;
    if (!st.unitary) return;
    P::norm2p(x, st.A);
    P::halve(x, x);       // Re(C^q1)
    P::canon(cx, x);
    ld<L>(one, c_fc.one);
    P::canon(t, one);
    if (P::eq_raw(cx, t)) {  // identity => 0 (recoverMessage, bgn.go:359-363)
      a.status[e] = 0;
      return;
    }
    uint32_t h = bsgs_hash_re(cx) & a.hmask;
    int64_t j = -1;
    for (;;) {
      uint32_t sl = a.slots[h];
      if (sl == 0) break;
      const uint32_t* el = a.elems + (size_t)(sl - 1) * 2 * L;
      uint32_t diff = 0;
      for (int k = 0; k < L; k++) diff |= el[k] ^ cx[k];
      if (diff == 0) {
        j = (int64_t)(sl - 1);
        break;
      }
      h = (h + 1) & a.hmask;
    }
    if (j < 0 || (uint64_t)(j + 1) > a.mmax) return;
    // Im(C^q1) Im(C) = Re(C^q1) Re(C) - Re(C^(q1+1))
    P::norm2p(xp, st.B);
    P::halve(xp, xp);
    ld<L>(c0, a.re + e * L);
    ld<L>(c1, a.im + e * L);
    P::mul(t, x, c0);
    P::sub(t, t, xp);
    P::canon(t, t);
    ld<L>(u, a.elems + (size_t)j * 2 * L + L);
    BGN_SETB(u, 1.0);
    P::mul(u, u, c1);
    P::canon(u, u);
    if (P::eq_raw(u, t)) {
      a.out[e] = j + 1;
      a.status[e] = 0;
      return;
    }
    uint32_t z[L];
    BGN_SETB(z, 0.0);
    BGN_UNROLL
    for (int k = 0; k < L; k++) z[k] = 0;
    P::sub(u, z, u);
    P::canon(u, u);
    if (P::eq_raw(u, t)) {
      a.out[e] = -(j + 1);
      a.status[e] = 0;
    }
  }

  // the exponent bits below the top one, most significant first
  BGN_DEV static int nbits() { return c_pc.exp_bits; }
  BGN_DEV static int bit(int i) { return (c_pc.exp[i >> 5] >> (i & 31)) & 1; }
};

// ---------------------------------------------------------------------------
// C^q1 as a full F_p^2 element (bgn_gt_pow_secret_batch; the giant-step decrypt), again with a
// pair of lanes per element: square-and-multiply in F_p^2 whose products split evenly --
//   squaring   (r0 + r1)(r0 - r1)  |  2 r0 r1                     one product per lane
//   times a    r0 a0 - r1 a1       |  r0 a1 + r1 a0               two products per lane
// -- so a lane runs 1 + 2 wt/bits sequential products per exponent bit instead of the 2 + 3 wt/bits
// of one thread, and both lanes execute the same instruction stream (operands picked by selects,
// results swapped with one shuffle per limb).  The exponent is a key constant: uniform control flow.
// ---------------------------------------------------------------------------
template <int L>
struct GtPowPair {
  typedef Fp<L> P;
  typedef Lucas<L> LU;
  struct State {
    uint32_t r0[L], r1[L];  // running power, both coordinates in both lanes (relaxed range)
    uint32_t a0[L], a1[L];  // the base
    uint32_t na1[L];        // -Im(a): conj(a) = a^-1 serves the exponent's negative digits
  };

  BGN_DEV static void init(State& st, const uint32_t* re, const uint32_t* im, bool active) {
    if (active) {
      ld<L>(st.a0, re);
      ld<L>(st.a1, im);
    } else {
      BGN_SETB(st.a0, 0.0);
      BGN_SETB(st.a1, 0.0);
      BGN_UNROLL
      for (int j = 0; j < L; j++) st.a0[j] = st.a1[j] = 0;
    }
    P::negk(st.na1, st.a1, c_fc.p2, 2);
    LU::sel(st.r0, true, st.a0, st.a0);
    LU::sel(st.r1, true, st.a1, st.a1);
  }
  // this lane's half of r^2
  BGN_DEV static void sqr_half(uint32_t (&t)[L], const State& st, int s) {
    uint32_t sum[L], dif[L], x[L], y[L], d[L];
    P::addn(sum, st.r0, st.r1);
    P::subk(dif, st.r0, st.r1, c_fc.p8, 8);
    LU::sel(x, s == 0, sum, st.r0);
    LU::sel(y, s == 0, dif, st.r1);
    P::mul(t, x, y);
    P::addn(d, t, t);
    LU::sel(t, s == 0, t, d);  // lane 1 holds 2 r0 r1
  }
  // this lane's half of r * a (neg: r * conj(a))
  BGN_DEV static void mul_half(uint32_t (&t)[L], const State& st, int s, bool neg) {
    uint32_t x[L], y[L], u[L], v[L], dif[L], sum[L], b1[L];
    LU::sel(b1, neg, st.na1, st.a1);
    LU::sel(x, s == 0, st.a0, b1);
    LU::sel(y, s == 0, b1, st.a0);
    P::mul(u, st.r0, x);
    P::mul(v, st.r1, y);
    P::subk(dif, u, v, c_fc.p4, 4);
    P::addn(sum, u, v);
    LU::sel(t, s == 0, dif, sum);
  }
  BGN_DEV static void update(State& st, const uint32_t (&mine)[L], const uint32_t (&other)[L], int s) {
    LU::sel(st.r0, s == 0, mine, other);
    LU::sel(st.r1, s == 0, other, mine);
  }
  // lane s stores its coordinate (re for lane 0, im for lane 1), back in [0, 2p)
  BGN_DEV static void finish(const State& S, uint32_t* ore, uint32_t* oim, int s) {
    uint32_t t[L];
    LU::sel(t, s == 0, S.r0, S.r1);
    P::norm2p(t, t);
    st<L>(s == 0 ? ore : oim, t);
  }
};
