// pairing.cuh -- reduced Tate pairing on the type-A1 curve and GT = mu_n in
// F_p^2, organised for batched polynomial products.
//
// Replaces libpbc a1_param.c (a1_pairing: Miller loop over the bits of n,
// "Tate exponentiation" conj(f)/f then ^l) as reached through Element.Pair from
// bgn.go:300 (Mult), bgn.go:318 (makeL2) and the schoolbook product of
// poly.go:140-152 (MultPoly), and element_pow_mpz on GT (bgn.go:223, 277).
//
// Redesign (results are the same canonical GT values):
//  * signed-digit (NAF) Miller loop in Jacobian coordinates, no inversions;
//  * a "team" of TS threads computes ALL d1*d2 pairings of one polynomial
//    product: thread i advances Miller point A_i once per step and publishes
//    the line coefficients in shared memory; every thread then folds the lines
//    evaluated at its B_k into the <= 2 output slots it owns, so the f^2
//    squarings and the final exponentiation happen once per OUTPUT slot
//    (d1+d2-1 of them), not once per pairing (d1*d2);
//  * a plain pairing batch is the same program with TS = 1.
#pragma once
#include "curve.cuh"

BGN_CONST PairConsts c_pc;

template <int L>
struct GT {
  typedef Fp<L> P;
  typedef F<L> FF;

  // f <- f^((p^2-1)/n) = (conj(f)/f)^l, in place.  conj(f)/f = conj(f)^2 / N(f).
  BGN_DEVNI static void final_exp(V2 f) {
    uint32_t f0[L], f1[L], a[L], b[L], ni[L], g0[L], g1[L], r0[L], r1[L];
    ld<L>(f0, f.re);
    ld<L>(f1, f.im);
    P::sqr(a, f0);
    P::sqr(b, f1);
    P::add(ni, a, b);
    FF::inv(mkv(ni, 1), mkv(ni, 1));
    P::sub(g0, a, b);
    P::mul(g0, g0, ni);
    P::mul(g1, f0, f1);
    P::add(g1, g1, g1);
    BGN_UNROLL
    for (int k = 0; k < L; k++) a[k] = 0;
    P::sub(g1, a, g1);
    P::mul(g1, g1, ni);
    // (g0 + g1 i)^l, MSB-first
    BGN_UNROLL
    for (int k = 0; k < L; k++) {
      r0[k] = g0[k];
      r1[k] = g1[k];
    }
    uint64_t l = c_pc.l;
    int top = 63;
    while (top > 0 && !((l >> top) & 1)) top--;
    V2 r = mkv2(mkv(r0, 1), mkv(r1, 1));
    V2 g = mkv2(mkv(g0, 1), mkv(g1, 1));
    for (int bit = top - 1; bit >= 0; bit--) {
      FF::sqr2(r, r);
      if ((l >> bit) & 1) FF::mul2(r, r, g);
    }
    st<L>(f.re, r0);
    st<L>(f.im, r1);
  }

  // r <- a^e for the fixed exponent c_pc.exp (Decrypt: C^q1, bgn.go:223).
  BGN_DEVNI static void pow_fixed(V2 r, V2 a) {
    uint32_t a0[L], a1[L], r0[L], r1[L];
    ld<L>(a0, a.re);
    ld<L>(a1, a.im);
    BGN_UNROLL
    for (int k = 0; k < L; k++) {
      r0[k] = a0[k];
      r1[k] = a1[k];
    }
    V2 rr = mkv2(mkv(r0, 1), mkv(r1, 1));
    V2 aa = mkv2(mkv(a0, 1), mkv(a1, 1));
    int nb = c_pc.exp_bits;
    if (nb == 0) {
      FF::set_one2(rr);
    }
    for (int bit = nb - 2; bit >= 0; bit--) {
      FF::sqr2(rr, rr);
      if ((c_pc.exp[bit >> 5] >> (bit & 31)) & 1) FF::mul2(rr, rr, aa);
    }
    st<L>(r.re, r0);
    st<L>(r.im, r1);
  }

  // r <- a^e, per-element exponent given as big-endian bytes (MultConst on L2, bgn.go:277).
  BGN_DEVNI static void pow_var(V2 r, V2 a, const uint8_t* e_be, int ebytes) {
    uint32_t a0[L], a1[L], r0[L], r1[L];
    ld<L>(a0, a.re);
    ld<L>(a1, a.im);
    V2 rr = mkv2(mkv(r0, 1), mkv(r1, 1));
    V2 aa = mkv2(mkv(a0, 1), mkv(a1, 1));
    FF::set_one2(rr);
    for (int i = 0; i < ebytes; i++) {
      uint32_t byte = e_be[i];
      for (int bit = 7; bit >= 0; bit--) {
        FF::sqr2(rr, rr);
        if ((byte >> bit) & 1) FF::mul2(rr, rr, aa);
      }
    }
    st<L>(r.re, r0);
    st<L>(r.im, r1);
  }
};

// ---------------------------------------------------------------------------
// Miller-loop team program
// ---------------------------------------------------------------------------

enum { MOP_DBL = 0, MOP_ADD = 1, MOP_SUB = 2 };

template <int L>
struct MillerTeam {
  typedef F<L> FF;
  typedef G<L> GG;
  enum { S_F0 = 0, S_F1 = 2, S_X = 4, S_Y = 5, S_Z = 6, S_BX = 7, S_BY = 8, S_CR = 9, S_AR = 10, S_BI = 11, NSLOT = BGN_MILLER_NSLOT };

  const MillerArgs& a;
  uint32_t* smem;   // NSLOT*L*nt words of state, then nt bytes flagsA, nt bytes flagsB
  int nt, tid, bid;
  int t, team, unit;
  bool active;

  BGN_DEV MillerTeam(const MillerArgs& a_, uint32_t* smem_, int tid_, int bid_, int nt_)
      : a(a_), smem(smem_), nt(nt_), tid(tid_), bid(bid_) {
    int TS = a.dE;
    t = tid % TS;
    team = tid / TS;
    unit = bid * a.teams_per_block + team;
    active = team < a.teams_per_block && unit < a.count;
  }
  static BGN_DEV size_t smem_bytes(int nt) { return (size_t)NSLOT * L * nt * 4 + 2 * (size_t)nt; }
  BGN_DEV V slot(int thread, int k) const { return mkv(smem + (size_t)k * L * nt + thread, nt); }
  BGN_DEV uint8_t* flagsA() const { return reinterpret_cast<uint8_t*>(smem + (size_t)NSLOT * L * nt); }
  BGN_DEV uint8_t* flagsB() const { return flagsA() + nt; }
  BGN_DEV V2 facc(int thread, int s) const { return mkv2(slot(thread, S_F0 + 2 * s), slot(thread, S_F0 + 2 * s + 1)); }

  BGN_DEV void init() {
    flagsA()[tid] = 0;
    flagsB()[tid] = 0;
    if (!active) return;
    FF::set_one2(facc(tid, 0));
    FF::set_one2(facc(tid, 1));
    if (t < a.dM) {
      size_t idx = (size_t)unit * a.dM + t;
      bool inf = a.Minf[idx] != 0;
      flagsA()[tid] = inf ? 0 : 1;
      if (!inf) {
        FF::copy(slot(tid, S_X), mkvc(a.Mx + idx, a.NM));
        FF::copy(slot(tid, S_Y), mkvc(a.My + idx, a.NM));
        FF::set_one(slot(tid, S_Z));
      }
    }
    {
      size_t idx = a.e_bcast ? (size_t)t : (size_t)unit * a.dE + t;
      bool inf = a.Einf[idx] != 0;
      flagsB()[tid] = inf ? 0 : 1;
      if (!inf) {
        FF::copy(slot(tid, S_BX), mkvc(a.Ex + idx, a.NE));
        FF::copy(slot(tid, S_BY), mkvc(a.Ey + idx, a.NE));
      }
    }
  }

  // phase A: advance own Miller point, publish its line; square own accumulators on doubling steps
  BGN_DEV void phaseA(int op, bool first) {
    if (!active) return;
    if (op == MOP_DBL && !first) {
      FF::sqr2(facc(tid, 0), facc(tid, 0));
      if (t + a.dE < a.dM + a.dE - 1) FF::sqr2(facc(tid, 1), facc(tid, 1));
    }
    if (t < a.dM && flagsA()[tid]) {
      if (op == MOP_DBL) {
        GG::dbl_line(slot(tid, S_X), slot(tid, S_Y), slot(tid, S_Z), slot(tid, S_CR), slot(tid, S_AR), slot(tid, S_BI));
      } else {
        size_t idx = (size_t)unit * a.dM + t;
        GG::madd_line(slot(tid, S_X), slot(tid, S_Y), slot(tid, S_Z), mkvc(a.Mx + idx, a.NM), mkvc(a.My + idx, a.NM),
                      op == MOP_SUB, slot(tid, S_CR), slot(tid, S_AR), slot(tid, S_BI));
      }
    }
  }

  // phase B: fold line_i(B_k) into the slot i+k for every Miller point i; this thread owns slots t and t+TS
  BGN_DEV void phaseB() {
    if (!active) return;
    int TS = a.dE;
    int base = tid - t;  // first thread of the team
    for (int i = 0; i < a.dM; i++) {
      int k = t - i;
      int s = 0;
      if (k < 0) {
        k += TS;
        s = 1;
      }
      if (!flagsA()[base + i] || !flagsB()[base + k]) continue;
      FF::line_mul(facc(tid, s), slot(base + i, S_CR), slot(base + i, S_AR), slot(base + i, S_BI), slot(base + k, S_BX),
                   slot(base + k, S_BY));
    }
  }

  BGN_DEV void finalize() {
    if (!active) return;
    int nslots = a.dM + a.dE - 1;
    for (int s = 0; s < 2; s++) {
      int j = t + s * a.dE;
      if (j >= nslots || j >= a.out_slots) continue;
      V2 f = facc(tid, s);
      GT<L>::final_exp(f);
      size_t o = (size_t)unit * a.out_slots + j;
      FF::copy(mkv(a.out_re + o, a.NOUT), f.re);
      FF::copy(mkv(a.out_im + o, a.NOUT), f.im);
    }
    if (t == 0) {
      for (int j = nslots; j < a.out_slots; j++) {  // padding slot(s): GT identity (poly.go:130-137)
        size_t o = (size_t)unit * a.out_slots + j;
        FF::set_one(mkv(a.out_re + o, a.NOUT));
        FF::set_zero(mkv(a.out_im + o, a.NOUT));
      }
    }
  }

  // number of (phaseA, phaseB) steps and their ops, shared by the kernel and the host simulator
  template <typename Sync>
  BGN_DEV void run(Sync sync) {
    init();
    sync();
    int n = c_pc.naf_len;
    for (int idx = 1; idx < n; idx++) {
      phaseA(MOP_DBL, idx == 1);
      sync();
      phaseB();
      sync();
      int d = c_pc.naf[idx];
      if (d != 0 && idx != n - 1) {
        phaseA(d > 0 ? MOP_ADD : MOP_SUB, false);
        sync();
        phaseB();
        sync();
      }
    }
    finalize();
  }
};
