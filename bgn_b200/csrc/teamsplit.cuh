// teamsplit.cuh -- the Miller team program of pairing.cuh with TWO threads per output-slot pair,
// for MultPoly batches of less than one wave (strong scaling: BASELINE config 3 split over 4 or 8
// GPUs leaves 4096 or 2048 products per GPU where one wave of k_miller is 3404).
//
// Why (measured, profiles/r02_bench_n2_v3.json: 8192 products per GPU cost three waves' time, 0.71 of the
// IMAD.WIDE peak; 2048 per GPU are 1.2 warps per scheduler): a team of dE threads walks 16 + 4.5 dM
// dependent products per Miller step, and below one wave nothing else is there to hide them.  Here a
// team is 2 dE threads, thread (t, h), h = 0 | 1:
//
//   phase A   (t, 0) advances Miller point t (dbl_line 12 / madd_line 13 products) -- as before;
//             (t, 1) squares the four partial accumulators of column t (its own two and (t, 0)'s):
//             8 products.  One thread did 16; the longer of the two now does 12.
//   phase B   the dM lines are split between the halves: (t, h) folds lines i of half h into ITS
//             partial accumulators of slots t and t + dE: ceil(dM / 2) line products instead of dM.
//   finalize  (t, 0) multiplies the two partials of each slot and runs the final exponentiation.
//
// Per doubling step at dM = dE = 11: 12 + 6 * 4.49 = 38.9 dependent products instead of 65.4, for
// 42 instead of 21 squarings per step (+5.8 % work).  Shared memory is laid out by need, not by
// thread: 4 accumulator slots per thread, 8 column slots (Miller point, line, evaluation point) per
// (team, t) -- 16 dE slots per team where the uniform layout of pairing.cuh would take 24 dE -- so an
// SM holds 14 teams of 22 threads (10 warps) where k_miller holds 23 teams of 11.
//
// A NAF digit != 0 is one merged doubling-and-addition step with a parabola (pairing.cuh: BGN_PARABOLA), split the
// same way: (t, 0) runs fused.cuh dadd_para while (t, 1) squares, then each half folds its share of the parabolas.
//
// Same results, bit for bit: the product of the two partial Miller values is the Miller value.
// Unit-stride layout only (up to 17 limbs); the 1024-bit field keeps k_miller.
#pragma once
#include "pairing.cuh"

template <int L>
struct MillerSplit {
  typedef F<L> FF;
  typedef MF<L, BGN_MILLER_LOOP, 1> M;     // phase B
  typedef MF<L, BGN_MILLER_LOOP_A, 1> MA;  // phase A
  static constexpr bool PARA = BGN_EVAL_NORM != 0 && BGN_PARABOLA != 0;  // pairing.cuh: doubling-and-addition steps
  enum { A_F0 = 0, A_F1 = 2, NA = 4,                                   // per thread: two partial accumulators
         C_X = 0, C_Y = 1, C_Z = 2, C_CR = 3, C_AR = 4, C_BI = 5, C_EX = 6, C_EY = 7, C_C3 = 8,
         NC = PARA ? 9 : 8 };  // per column (PARA: the parabola's fourth coefficient)

  const MillerArgs& a;
  uint32_t* smem;
  int nt, ncol, tid;
  int t, h, team, unit, col, base_col, ptid;  // ptid: the partner thread (t, 1 - h)
  bool active;

  // A block is teams_per_group teams.  Its first blockDim / 2 threads are the (t, 0) halves of all teams,
  // the second blockDim / 2 the (t, 1) halves (each half padded to whole warps): a warp then holds threads
  // of ONE role, so the Miller-point warps (dbl_line / madd_line) and the squaring warps of phase A run side
  // by side.  With both halves of a team in one warp the two roles diverge and serialise -- measured: 72 ms
  // against 54 ms for a batch of 692 products (profiles/r02_split_ab_v1.json against _v2).
  BGN_DEV MillerSplit(const MillerArgs& a_, uint32_t* smem_, int tid_, int bid_, int nt_)
      : a(a_), smem(smem_), nt(nt_), tid(tid_) {
    const int half = nt >> 1;
    ncol = a.teams_per_group * a.dE;
    h = tid >= half ? 1 : 0;
    const int r = tid - h * half;
    team = r / a.dE;
    t = r - team * a.dE;
    unit = bid_ * a.teams_per_group + team;
    active = team < a.teams_per_group && unit < a.count;
    base_col = team * a.dE;
    col = base_col + t;
    ptid = (1 - h) * half + r;
  }
  static BGN_HD size_t smem_words(int nt, int ncol) {
    return ((size_t)NA * nt + (size_t)NC * ncol) * L + (2 * (size_t)ncol + 3) / 4;
  }
  BGN_DEV E acc(int thread, int k) const { return smem + ((size_t)k * nt + thread) * L; }
  BGN_DEV E cslot(int c, int k) const { return smem + (size_t)NA * nt * L + ((size_t)k * ncol + c) * L; }
  BGN_DEV uint8_t* flagsA() const { return reinterpret_cast<uint8_t*>(smem + ((size_t)NA * nt + (size_t)NC * ncol) * L); }
  BGN_DEV uint8_t* flagsB() const { return flagsA() + ncol; }
  BGN_DEV bool owns2() const { return t + a.dE < a.dM + a.dE - 1; }

  BGN_DEV void init() {
    if (!active) return;
    for (int s = 0; s < 2; s++) {
      FF::copy(acc(tid, A_F0 + 2 * s), c_fc.one);
      FF::set_zero(acc(tid, A_F0 + 2 * s + 1));
    }
    if (h != 0) return;
    flagsA()[col] = 0;
    if (t < a.dM) {
      size_t idx = (size_t)unit * a.dM + t;
      bool inf = a.Minf[idx] != 0;
      flagsA()[col] = inf ? 0 : 1;
      if (!inf) {
        FF::copy(cslot(col, C_X), a.Mx + idx * L);
        FF::copy(cslot(col, C_Y), a.My + idx * L);
        FF::copy(cslot(col, C_Z), c_fc.one);
      }
    }
    size_t e = a.e_bcast ? (size_t)t : (size_t)unit * a.dE + t;
    bool einf = a.Einf[e] != 0;
    if (BGN_EVAL_NORM && !einf && FF::is_zero(a.Ey + e * L)) einf = true;  // (0, 0): pairings are 1 (pairing.cuh: init)
    flagsB()[col] = einf ? 0 : 1;
    if (!einf) {
      FF::copy(cslot(col, C_EX), a.Ex + e * L);
      FF::copy(cslot(col, C_EY), a.Ey + e * L);
      if (BGN_EVAL_NORM) MA::eval_normalise(cslot(col, C_EX), cslot(col, C_EY));  // (x / y, 1 / y): fused.cuh line_mul_n
      if (PARA && a.para) MA::mul_to_unit(a.evw + e * L, a.Ex + e * L, cslot(col, C_EX));  // x^2 / y
    }
  }

  BGN_DEV void phaseA(int op, bool first) {
    if (!active) return;
    if (h == 1) {
      if (op == MOP_DBL && !first) {
        MA::sqr2(acc(tid, A_F0), acc(tid, A_F0 + 1));
        MA::sqr2(acc(ptid, A_F0), acc(ptid, A_F0 + 1));
        if (owns2()) {
          MA::sqr2(acc(tid, A_F1), acc(tid, A_F1 + 1));
          MA::sqr2(acc(ptid, A_F1), acc(ptid, A_F1 + 1));
        }
      }
      return;
    }
    if (t < a.dM && flagsA()[col]) {
      if (op == MOP_DBL) {
        MA::dbl_line(cslot(col, C_X), cslot(col, C_Y), cslot(col, C_Z), cslot(col, C_CR), cslot(col, C_AR), cslot(col, C_BI));
      } else {
        size_t idx = (size_t)unit * a.dM + t;
        MA::madd_line(cslot(col, C_X), cslot(col, C_Y), cslot(col, C_Z), a.Mx + idx * L, a.My + idx * L, op == MOP_SUB,
                      cslot(col, C_CR), cslot(col, C_AR), cslot(col, C_BI));
      }
    }
  }

  // a doubling-and-addition step (pairing.cuh: MillerTeam::phaseA_dadd / phaseB_para) with the same split:
  // (t, 1) squares the four partial accumulators of column t, (t, 0) replaces Miller point t by 2T +- A and
  // publishes the parabola; then each half folds its share of the parabolas
  BGN_DEV void phaseA_dadd(int op) {
    if (!active) return;
    if (h == 1) {
      MA::sqr2(acc(tid, A_F0), acc(tid, A_F0 + 1));
      MA::sqr2(acc(ptid, A_F0), acc(ptid, A_F0 + 1));
      if (owns2()) {
        MA::sqr2(acc(tid, A_F1), acc(tid, A_F1 + 1));
        MA::sqr2(acc(ptid, A_F1), acc(ptid, A_F1 + 1));
      }
      return;
    }
    if (t < a.dM && flagsA()[col]) {
      size_t idx = (size_t)unit * a.dM + t;
      MA::dadd_para(cslot(col, C_X), cslot(col, C_Y), cslot(col, C_Z), a.Mx + idx * L, a.My + idx * L, op == MOP_SUB,
                    cslot(col, C_CR), cslot(col, C_AR), cslot(col, C_BI), cslot(col, C_C3));
    }
  }
  BGN_DEV void phaseB_para() {
    if (!active) return;
    const int TS = a.dE;
    const int mid = (a.dM + 1) / 2;
    // the half that advanced the Miller point (30 products against 8 squarings) folds the smaller share
    const int i0 = h ? a.dM - mid : 0, i1 = h ? a.dM : a.dM - mid;
    for (int i = i0; i < i1; i++) {
      int k = t - i;
      int s = 0;
      if (k < 0) {
        k += TS;
        s = 1;
      }
      if (!flagsA()[base_col + i] || !flagsB()[base_col + k]) continue;
      E fr = acc(tid, A_F0 + 2 * s), fi = acc(tid, A_F0 + 2 * s + 1);
      const size_t e = a.e_bcast ? (size_t)k : (size_t)unit * a.dE + k;
#if BGN_LINE_LAZY
      M::template para_mul_lazy<BGN_LINE_KARATSUBA>(fr, fi, cslot(base_col + i, C_CR), cslot(base_col + i, C_AR),
                                                    cslot(base_col + i, C_BI), cslot(base_col + i, C_C3),
                                                    cslot(base_col + k, C_EX), cslot(base_col + k, C_EY), a.evw + e * L);
#else
      M::para_mul(fr, fi, cslot(base_col + i, C_CR), cslot(base_col + i, C_AR), cslot(base_col + i, C_BI),
                  cslot(base_col + i, C_C3), cslot(base_col + k, C_EX), cslot(base_col + k, C_EY), a.evw + e * L);
#endif
    }
  }

  BGN_DEV void phaseB() {
    if (!active) return;
    const int TS = a.dE;
    const int mid = (a.dM + 1) / 2;
    const int i0 = h ? mid : 0, i1 = h ? a.dM : mid;
    for (int i = i0; i < i1; i++) {
      int k = t - i;
      int s = 0;
      if (k < 0) {
        k += TS;
        s = 1;
      }
      if (!flagsA()[base_col + i] || !flagsB()[base_col + k]) continue;
      E fr = acc(tid, A_F0 + 2 * s), fi = acc(tid, A_F0 + 2 * s + 1);
      const uint32_t *cR = cslot(base_col + i, C_CR), *aR = cslot(base_col + i, C_AR), *bI = cslot(base_col + i, C_BI);
      const uint32_t *ex = cslot(base_col + k, C_EX), *ey = cslot(base_col + k, C_EY);
#if BGN_LINE_LAZY
      if (BGN_EVAL_NORM)
        M::template line_mul_lazy_n<BGN_LINE_KARATSUBA>(fr, fi, cR, aR, bI, ex, ey);
      else
        M::template line_mul_lazy<BGN_LINE_KARATSUBA>(fr, fi, cR, aR, bI, ex, ey);
#else
      if (BGN_EVAL_NORM)
        M::line_mul_n(fr, fi, cR, aR, bI, ex, ey);
      else
        M::line_mul(fr, fi, cR, aR, bI, ex, ey);
#endif
    }
  }

  // (t, 0): slot value = own partial * partner's partial, then (conj(f)^2 / N(f))^l as
  // MillerTeam::finalize; the column's slots are scratch by now.  Called after a barrier.
  BGN_DEV void finalize() {
    if (!active || h != 0) return;
    int nslots = a.dM + a.dE - 1;
    bool own0 = t < nslots && t < a.out_slots;
    bool own1 = t + a.dE < nslots && t + a.dE < a.out_slots;
    E f0r = acc(tid, A_F0), f0i = acc(tid, A_F0 + 1), f1r = acc(tid, A_F1), f1i = acc(tid, A_F1 + 1);
    E n0 = cslot(col, C_X), n1 = cslot(col, C_Y), w = cslot(col, C_Z), i0 = cslot(col, C_CR), i1 = cslot(col, C_AR);
    if (own0) MA::mul2(f0r, f0i, acc(ptid, A_F0), acc(ptid, A_F0 + 1));
    if (own1) MA::mul2(f1r, f1i, acc(ptid, A_F1), acc(ptid, A_F1 + 1));
    if (own0) MA::fe_prepare(f0r, f0i, n0);
    if (own1) MA::fe_prepare(f1r, f1i, n1);
    if (own1) {
      MA::fp_mul(w, n0, n1);
      FF::template inv_gcd<true, 1>(w, w);
      MA::fp_mul(i0, w, n1);
      MA::fp_mul(i1, w, n0);
    } else if (own0) {
      FF::template inv_gcd<true, 1>(i0, n0);
    }
    for (int s = 0; s < 2; s++) {
      if (!(s ? own1 : own0)) continue;
      E fr = s ? f1r : f0r, fi = s ? f1i : f0i;
      MA::scale2(fr, fi, s ? i1 : i0);
      FF::copy(n0, fr);
      FF::copy(n1, fi);
      uint64_t l = c_pc.l;
      int top = 63;
      while (top > 0 && !((l >> top) & 1)) top--;
      for (int bit = top - 1; bit >= 0; bit--) {
        MA::sqr2(fr, fi);
        if ((l >> bit) & 1) MA::mul2(fr, fi, n0, n1);
      }
      MA::norm2(fr, fi);
      size_t o = (size_t)unit * a.out_slots + t + s * a.dE;
      FF::copy(a.out_re + o * L, fr);
      FF::copy(a.out_im + o * L, fi);
    }
    if (t == 0) {
      for (int j = nslots; j < a.out_slots; j++) {  // padding slot(s): GT identity (poly.go:130-137)
        size_t o = (size_t)unit * a.out_slots + j;
        FF::set_one(a.out_re + o * L);
        FF::set_zero(a.out_im + o * L);
      }
    }
  }

  template <typename Sync>
  BGN_DEV void run(Sync sync) {
    init();
    sync();
    int n = c_pc.naf_len;
    for (int idx = 1; idx < n; idx++) {
      int d = c_pc.naf[idx];
      if (PARA && a.para && d != 0 && idx != n - 1) {
        phaseA_dadd(d > 0 ? MOP_ADD : MOP_SUB);
        sync();
        phaseB_para();
        sync();
        continue;
      }
      phaseA(MOP_DBL, idx == 1);
      sync();
      phaseB();
      sync();
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
