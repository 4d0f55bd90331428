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
//  * a plain pairing batch is the same program with TS = 1;
//  * (round 2) line values are scaled by F_p factors that the final exponentiation removes -- evaluation points
//    normalised to (x / y, 1 / y), BGN_EVAL_NORM -- and a NAF digit != 0 is ONE doubling-and-addition step whose
//    tangent and chord are merged into a parabola, BGN_PARABOLA.
#pragma once
#include "curve.cuh"
#include "fused.cuh"

// U of fused.cuh for the phase-B routine (line_mul) and for the phase-A routines (sqr2, dbl_line,
// madd_line): 0 = fully unrolled products, U > 0 = 2U rows per loop iteration.  Measured at L = 17
// (profiles/r01_primbench_v5.txt, 2 warps per scheduler): unrolled 94 % of the IMAD.WIDE peak,
// U = 4 87 %, U = 1 80 %; an unrolled product is L(2L+1) x 16 B of code (9.5 KB at L = 17,
// 35 KB at L = 33), so the big field keeps the loop.
#ifndef BGN_MILLER_LOOP
#define BGN_MILLER_LOOP (BGN_L <= 17 ? 0 : 4)
#endif
// Phase B with lazy reduction (line_mul_lazy: double-width products sharing two reductions).
// Measured at L = 17: +9 % line_mul rate, k_miller 646 -> 611 ms (profiles/r01_ab_v9_lazy.txt).
// Its products are fully unrolled and hold three double-width values in registers, so the
// 1024-bit field (L = 33) keeps the single-width routine.
#ifndef BGN_LINE_LAZY
#define BGN_LINE_LAZY (BGN_L <= 17 ? 1 : 0)
#endif
// Thread-interleaved layout with the thread-private state in global memory (the 1024-bit field):
// 12 slots of 33 limbs are 1.5 KB per thread, so shared memory holds only 4 warps per SM -- one per
// scheduler, where a single warp cannot issue more than ~2/3 of the multiplier's rate.  With
// BGN_MILLER_GP the 7 private slots (accumulators, Miller point) live in an L2-resident scratch
// array and only the 5 slots other threads read (line, evaluation point) stay in shared memory;
// every slot is laid out [slot][limb][thread] with a compile-time thread stride BGN_MILLER_NT, so
// limb addresses are immediates and a warp's access to one limb is one 128-byte line.
#ifndef BGN_MILLER_GP
#define BGN_MILLER_GP (BGN_L > 17 ? 1 : 0)
#endif
#ifndef BGN_MILLER_NT
#define BGN_MILLER_NT 256
#endif
#ifndef BGN_LINE_KARATSUBA
#define BGN_LINE_KARATSUBA 0  // KM of fused.cuh line_mul_lazy
#endif
#ifndef BGN_MILLER_LOOP_A
#define BGN_MILLER_LOOP_A (BGN_L <= 17 ? 0 : 4)
#endif
// Phase B of the TEAM kernel (line_mul: three quarters of its work).  Round 2, measured at 33 limbs on a
// full wave of 8 x 8 products (profiles/r02_ab1024_unroll.txt): 8-row loop in both phases 0.813 of the
// IMAD.WIDE peak, everything unrolled 0.831, phase B unrolled and phase A looped 0.850 -- the five
// unrolled products of line_mul (175 KB) are walked by all warps of the block in lockstep, the twelve
// of dbl_line would not stay cached.  Up to 17 limbs this is BGN_MILLER_LOOP (unrolled) as before.
#ifndef BGN_TEAM_LOOP_B
#define BGN_TEAM_LOOP_B (BGN_L <= 17 ? BGN_MILLER_LOOP : 0)
#endif

// Evaluation points normalised to (x / y, 1 / y) before the loop, so that a line value is one dot product
// and the line's third coefficient (fused.cuh: line_mul_n): 8.7 % fewer products per MultPoly unit at
// 11 x 11 slots.  0 restores the plain form (A/B).
#ifndef BGN_EVAL_NORM
#define BGN_EVAL_NORM 1
#endif

// A NAF digit != 0 as ONE doubling-and-addition step whose tangent and chord are merged into a parabola
// (curve.cuh: G::dadd_para, fused.cuh: para_mul): one F_p^2 product per evaluation point on those steps instead
// of two.  Team kernel, both layouts; needs BGN_EVAL_NORM.  0 restores the two separate steps (A/B).
#ifndef BGN_PARABOLA
#define BGN_PARABOLA (BGN_EVAL_NORM ? 1 : 0)
#endif

BGN_CONST PairConsts c_pc;

template <int L>
struct GT {
  typedef F<L> FF;

  // r <- a^e for the fixed exponent c_pc.exp (Decrypt: C^q1, bgn.go:223).  r must not alias a.
  BGN_DEV static void pow_fixed(E2 r, E2 a, E t0, E t1, E t2) {
    int nb = c_pc.exp_bits;
    if (nb == 0) {
      FF::set_one2(r);
      return;
    }
    FF::copy2(r, a);
    for (int bit = nb - 2; bit >= 0; bit--) {
      FF::sqr2(r, r, t0, t1);
      if ((c_pc.exp[bit >> 5] >> (bit & 31)) & 1) FF::mul2(r, r, a, t0, t1, t2);
    }
  }

  // r <- a^e, per-element exponent given as big-endian bytes (MultConst on L2, bgn.go:277).
  // r must not alias a.
  BGN_DEV static void pow_var(E2 r, E2 a, const uint8_t* e_be, int ebytes, E t0, E t1, E t2) {
    FF::set_one2(r);
    bool started = false;  // squaring 1 is a no-op
    for (int i = 0; i < ebytes; i++) {
      uint32_t byte = e_be[i];
      for (int bit = 7; bit >= 0; bit--) {
        if (started) FF::sqr2(r, r, t0, t1);
        if ((byte >> bit) & 1) {
          FF::mul2(r, r, a, t0, t1, t2);
          started = true;
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------
// Miller-loop team program
// ---------------------------------------------------------------------------

enum { MOP_DBL = 0, MOP_ADD = 1, MOP_SUB = 2 };

// EG ("evaluation points in global memory", round 2): the two evaluation-point slots are not copied into
// shared memory -- line_mul reads xB, yB straight from the batch arrays (34 cached loads per 2 669
// products) -- so a thread's state is 10 slots instead of 12 and an SM holds 28 teams of 11 in 10 warps
// where 12 slots allow 23 in 8 (k_miller_wide, kernels.cuh).  Unit-stride layout only.
template <int L, bool EG = false>
struct MillerTeam {
  typedef F<L> FF;
  typedef G<L> GG;
  static constexpr bool GP = BGN_MILLER_GP != 0;
  static_assert(!(EG && GP), "the wide variant uses the unit-stride layout");
  static constexpr int NT = BGN_MILLER_NT;
  static constexpr int ES = GP ? NT : 1;          // element stride of every slot
  static constexpr bool NORM = !EG && BGN_EVAL_NORM != 0;  // slots S_EX, S_EY hold (x / y, 1 / y)
  static constexpr bool PARA = NORM && BGN_PARABOLA != 0;  // doubling-and-addition steps use the parabola
  typedef MF<L, BGN_TEAM_LOOP_B, ES> M;      // phase B
  typedef MF<L, BGN_MILLER_LOOP_A, ES> MA;   // phase A
  // element slots per thread in shared memory: two GT accumulators, the thread's Miller point,
  // the line it publishes, its evaluation point.  The loop's routines are fused (fused.cuh) and
  // keep their temporaries in registers.
  // (PARA: a 13th slot for the parabola's fourth coefficient; its x^2 / y values live in a global array)
  enum { S_F0 = 0, S_F1 = 2, S_X = 4, S_Y = 5, S_Z = 6, S_CR = 7, S_AR = 8, S_BI = 9, S_EX = 10, S_EY = 11, S_C3 = 12,
         NSLOT = EG ? 10 : (PARA ? 13 : BGN_MILLER_NSLOT), NPRIV = BGN_MILLER_NPRIV };  // slots < NPRIV are thread-private

  const MillerArgs& a;
  uint32_t* smem;   // the block's slots (layout: slot()), then one flagsA and one flagsB byte per thread
  int nt, tid, bid;
  int t, team, unit, group;
  bool active;

  // A block is `groups` barrier groups of a.group_threads threads; each group holds whole teams and
  // synchronises on its own named barrier, so the groups drift independently: while the warps of one
  // group run the add/sub/load glue between two products, the warps of the other group (which share
  // the same schedulers) keep the multiply pipe busy (profiles/r01_miller_v3_ncu.txt: in lockstep the
  // pipe idles during the glue of BOTH warps of a scheduler).
  BGN_DEV MillerTeam(const MillerArgs& a_, uint32_t* smem_, int tid_, int bid_, int nt_)
      : a(a_), smem(smem_), nt(nt_), tid(tid_), bid(bid_) {
    int TS = a.dE;
    int groups = nt / a.group_threads;
    group = tid / a.group_threads;
    int ltid = tid - group * a.group_threads;
    t = ltid % TS;
    team = ltid / TS;
    unit = (bid * groups + group) * a.teams_per_group + team;
    active = team < a.teams_per_group && unit < a.count;
  }
  // shared-memory words a block of nt threads needs (host side: api.cu, hostsim.cpp)
  static BGN_HD size_t smem_words(int nt) {
    return GP ? (size_t)(NSLOT - NPRIV) * L * NT + (2 * (size_t)NT + 3) / 4 : (size_t)NSLOT * L * nt + (2 * (size_t)nt + 3) / 4;
  }
  // global scratch words per block (GP only)
  static BGN_HD size_t priv_words() { return GP ? (size_t)NPRIV * L * NT : 0; }
  static BGN_HD int fixed_threads() { return GP ? NT : 0; }  // interleaved layout: blockDim is fixed
  // unit-stride layout: thread stride L words is odd for every supported L, so a warp touching one
  // limb of one slot hits 32 different banks; interleaved layout: consecutive threads, consecutive words
  BGN_DEV E slot(int thread, int k) const {
    if (GP) {
      if (k < NPRIV) return a.priv + ((size_t)bid * NPRIV + k) * L * NT + thread;
      return smem + (size_t)(k - NPRIV) * L * NT + thread;
    }
    return smem + ((size_t)k * nt + thread) * L;
  }
  BGN_DEV uint8_t* flagsA() const {
    return reinterpret_cast<uint8_t*>(smem + (GP ? (size_t)(NSLOT - NPRIV) * L * NT : (size_t)NSLOT * L * nt));
  }
  BGN_DEV uint8_t* flagsB() const { return flagsA() + (GP ? NT : nt); }
  BGN_DEV size_t eidx(int k) const { return a.e_bcast ? (size_t)k : (size_t)unit * a.dE + k; }
  // element moves between slots (stride ES) and the batch arrays (unit stride)
  BGN_DEV static void s_in(E dst, const uint32_t* src) {
    BGN_SETB(dst, BGN_GETB(src));
    for (int j = 0; j < L; j++) dst[j * ES] = src[j];
  }
  BGN_DEV static void s_out(uint32_t* dst, const uint32_t* src) {
    BGN_SETB(dst, BGN_GETB(src));
    for (int j = 0; j < L; j++) dst[j] = src[j * ES];
  }
  BGN_DEV static void s_copy(E dst, const uint32_t* src) {
    BGN_SETB(dst, BGN_GETB(src));
    for (int j = 0; j < L; j++) dst[j * ES] = src[j * ES];
  }
  BGN_DEV static void s_zero(E dst) {
    BGN_SETB(dst, 0.0);
    for (int j = 0; j < L; j++) dst[j * ES] = 0;
  }

  BGN_DEV void init() {
    flagsA()[tid] = 0;
    flagsB()[tid] = 0;
    if (!active) return;
    for (int s = 0; s < 2; s++) {
      s_in(slot(tid, S_F0 + 2 * s), c_fc.one);
      s_zero(slot(tid, S_F0 + 2 * s + 1));
    }
    if (t < a.dM) {
      size_t idx = (size_t)unit * a.dM + t;
      bool inf = a.Minf[idx] != 0;
      flagsA()[tid] = inf ? 0 : 1;
      if (!inf) {
        s_in(slot(tid, S_X), a.Mx + idx * L);
        s_in(slot(tid, S_Y), a.My + idx * L);
        s_in(slot(tid, S_Z), c_fc.one);
      }
    }
    bool einf = a.Einf[eidx(t)] != 0;
    // A finite evaluation point with y = 0 is the point of order 2, (0, 0); every line value at it lies in F_p, so
    // its pairings are 1 -- what skipping it like O yields, and what keeps 1 / y out of the normalisation.  (The
    // byte format already reads (0, 0) as O; only a handle can carry it.)
    if (NORM && !einf && FF::is_zero(a.Ey + eidx(t) * L)) einf = true;
    flagsB()[tid] = einf ? 0 : 1;
    if (!einf && !EG) {
      s_in(slot(tid, S_EX), a.Ex + eidx(t) * L);
      s_in(slot(tid, S_EY), a.Ey + eidx(t) * L);
      if (NORM) MA::eval_normalise(slot(tid, S_EX), slot(tid, S_EY));
      if (PARA && a.para) MA::mul_to_unit(a.evw + eidx(t) * L, a.Ex + eidx(t) * L, slot(tid, S_EX));  // x^2 / y = x * (x / y)
    }
  }

  // phase A of a doubling-and-addition step: square own accumulators, replace own Miller point T by 2T +- A and
  // publish the parabola (slots S_CR, S_AR, S_BI, S_C3 = cs, c1, c0, ci)
  BGN_DEV void phaseA_dadd(int op) {
    if (!active) return;
    MA::sqr2(slot(tid, S_F0), slot(tid, S_F0 + 1));
    if (t + a.dE < a.dM + a.dE - 1) MA::sqr2(slot(tid, S_F1), slot(tid, S_F1 + 1));
    if (t < a.dM && flagsA()[tid]) {
      size_t idx = (size_t)unit * a.dM + t;
      MA::dadd_para(slot(tid, S_X), slot(tid, S_Y), slot(tid, S_Z), a.Mx + idx * L, a.My + idx * L, op == MOP_SUB,
                    slot(tid, S_CR), slot(tid, S_AR), slot(tid, S_BI), slot(tid, S_C3));
    }
  }
  // phase B of such a step: fold parabola_i(B_k) into the slot i+k
  BGN_DEV void phaseB_para() {
    if (!active) return;
    int TS = a.dE;
    int base = tid - t;
    for (int i = 0; i < a.dM; i++) {
      int k = t - i;
      int s = 0;
      if (k < 0) {
        k += TS;
        s = 1;
      }
      if (!flagsA()[base + i] || !flagsB()[base + k]) continue;
      E fr = slot(tid, S_F0 + 2 * s), fi = slot(tid, S_F0 + 2 * s + 1);
      const uint32_t *cs = slot(base + i, S_CR), *c1 = slot(base + i, S_AR), *c0 = slot(base + i, S_BI),
                     *ci = slot(base + i, S_C3);
      const uint32_t *eu = slot(base + k, S_EX), *ev = slot(base + k, S_EY), *ew = a.evw + eidx(k) * L;
#if BGN_LINE_LAZY
      M::template para_mul_lazy<BGN_LINE_KARATSUBA>(fr, fi, cs, c1, c0, ci, eu, ev, ew);
#else
      M::para_mul(fr, fi, cs, c1, c0, ci, eu, ev, ew);
#endif
    }
  }

  // phase A: advance own Miller point, publish its line; square own accumulators on doubling steps
  BGN_DEV void phaseA(int op, bool first) {
    if (!active) return;
    if (op == MOP_DBL && !first) {
      MA::sqr2(slot(tid, S_F0), slot(tid, S_F0 + 1));
      if (t + a.dE < a.dM + a.dE - 1) MA::sqr2(slot(tid, S_F1), slot(tid, S_F1 + 1));
    }
    if (t < a.dM && flagsA()[tid]) {
      if (op == MOP_DBL) {
        MA::dbl_line(slot(tid, S_X), slot(tid, S_Y), slot(tid, S_Z), slot(tid, S_CR), slot(tid, S_AR), slot(tid, S_BI));
      } else {
        size_t idx = (size_t)unit * a.dM + t;
        MA::madd_line(slot(tid, S_X), slot(tid, S_Y), slot(tid, S_Z), (a.Mx + (size_t)(idx) * L), (a.My + (size_t)(idx) * L),
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
      const uint32_t* ex = EG ? a.Ex + eidx(k) * L : slot(base + k, S_EX);
      const uint32_t* ey = EG ? a.Ey + eidx(k) * L : slot(base + k, S_EY);
      E fr = slot(tid, S_F0 + 2 * s), fi = slot(tid, S_F0 + 2 * s + 1);
      const uint32_t *cR = slot(base + i, S_CR), *aR = slot(base + i, S_AR), *bI = slot(base + i, S_BI);
#if BGN_LINE_LAZY
      if (NORM)
        M::template line_mul_lazy_n<BGN_LINE_KARATSUBA>(fr, fi, cR, aR, bI, ex, ey);
      else
        M::template line_mul_lazy<BGN_LINE_KARATSUBA>(fr, fi, cR, aR, bI, ex, ey);
#else
      if (NORM)
        M::line_mul_n(fr, fi, cR, aR, bI, ex, ey);
      else
        M::line_mul(fr, fi, cR, aR, bI, ex, ey);
#endif
    }
  }

  // Final exponentiation of the <= 2 slots this thread owns, (conj(f)^2 / N(f))^l, with fused
  // routines (fused.cuh) and ONE inversion per thread: a thread that owns two slots inverts
  // N0 N1 and recovers both inverses with three products (Montgomery's trick).  The thread's
  // Miller point and line slots are dead by now and serve as scratch.
  BGN_DEV void finalize() {
    if (!active) return;
    int nslots = a.dM + a.dE - 1;
    bool own0 = t < nslots && t < a.out_slots;
    bool own1 = t + a.dE < nslots && t + a.dE < a.out_slots;
    E f0r = slot(tid, S_F0), f0i = slot(tid, S_F0 + 1), f1r = slot(tid, S_F1), f1i = slot(tid, S_F1 + 1);
    E n0 = slot(tid, S_X), n1 = slot(tid, S_Y), w = slot(tid, S_Z), i0 = slot(tid, S_CR), i1 = slot(tid, S_AR);
    if (own0) MA::fe_prepare(f0r, f0i, n0);
    if (own1) MA::fe_prepare(f1r, f1i, n1);
    // the inversion is the binary GCD of arith.cuh (ALU pipe), not a Fermat power: measured 610.0 ->
    // 597.2 ms for the 2^14-pair batch (BGN_MILLER_FERMAT_INV restores the power)
#ifdef BGN_MILLER_FERMAT_INV
#define BGN_TEAM_INV(r, a) MA::fp_inv(r, a)
#else
#define BGN_TEAM_INV(r, a) FF::template inv_gcd<true, ES>(r, a)
#endif
    if (own1) {
      MA::fp_mul(w, n0, n1);
      BGN_TEAM_INV(w, w);
      MA::fp_mul(i0, w, n1);
      MA::fp_mul(i1, w, n0);
    } else if (own0) {
      BGN_TEAM_INV(i0, n0);
    }
    for (int s = 0; s < 2; s++) {
      if (!(s ? own1 : own0)) continue;
      E fr = s ? f1r : f0r, fi = s ? f1i : f0i;
      MA::scale2(fr, fi, s ? i1 : i0);  // g = conj(f)^2 / N(f) = f^(p-1)
      s_copy(n0, fr);
      s_copy(n1, fi);
      uint64_t l = c_pc.l;
      int top = 63;
      while (top > 0 && !((l >> top) & 1)) top--;
      for (int bit = top - 1; bit >= 0; bit--) {  // f = g^l, MSB first
        MA::sqr2(fr, fi);
        if ((l >> bit) & 1) MA::mul2(fr, fi, n0, n1);
      }
      MA::norm2(fr, fi);
      size_t o = (size_t)unit * a.out_slots + t + s * a.dE;
      s_out(a.out_re + o * L, fr);
      s_out(a.out_im + o * L, fi);
    }
    if (t == 0) {
      for (int j = nslots; j < a.out_slots; j++) {  // padding slot(s): GT identity (poly.go:130-137)
        size_t o = (size_t)unit * a.out_slots + j;
        FF::set_one(a.out_re + (size_t)(o) * L);
        FF::set_zero(a.out_im + (size_t)(o) * L);
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
      int d = c_pc.naf[idx];
      const bool add = d != 0 && idx != n - 1;
      if (PARA && a.para && add) {  // never the first step: a NAF has no two adjacent non-zero digits
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
      if (add) {
        phaseA(d > 0 ? MOP_ADD : MOP_SUB, false);
        sync();
        phaseB();
        sync();
      }
    }
    finalize();
  }
};

// ---------------------------------------------------------------------------
// Fixed-first-argument pairing: replay of a recorded line table (types.h: MillerFixedArgs)
// ---------------------------------------------------------------------------
template <int L>
struct MillerFixed {
  typedef MF<L, BGN_MILLER_LOOP, 1> M;
  typedef MF<L, BGN_MILLER_LOOP_A, 1> MA;
  // unit-stride shared-memory slots [slot][thread][L]: the accumulator, the evaluation point, and
  // two more for the final exponentiation (the evaluation point's slots are dead by then)
  enum { S_FR = 0, S_FI = 1, S_EX = 2, S_EY = 3, S_G0 = 4, S_G1 = 5, NSLOT = 6 };
  static BGN_HD size_t smem_words(int nt) { return (size_t)NSLOT * L * nt; }
  // number of lines a key's table holds: one per doubling and one per non-zero digit below the top
  // one, the last digit's chord dropped (vertical line)
  static BGN_HD int nsteps(const PairConsts& pc) {
    int n = 0;
    for (int idx = 1; idx < pc.naf_len; idx++) {
      n++;
      if (pc.naf[idx] != 0 && idx != pc.naf_len - 1) n++;
    }
    return n;
  }

  // a NAF digit != 0 below the top one is ONE doubling-and-addition step with a parabola entry (BGN_PARABOLA)
  static constexpr bool PARA = BGN_PARABOLA != 0;
  static BGN_HD bool is_dadd(const PairConsts& pc, int idx) { return PARA && pc.naf[idx] != 0 && idx != pc.naf_len - 1; }

  // The table of the affine point (px, py): the Miller loop's point arithmetic alone, one thread, in the step
  // order of MillerTeam::run, every entry NORMALISED by its imaginary coefficient -- F_p factors of a line
  // value die in the final exponentiation:
  //     doubling step            [cR / bI | aR / bI]             line at (xB, yB): (cRn + aRn xB) + yB i
  //     doubling-and-addition    [cs / ci | c1 / ci | c0 / ci]   parabola: ((csn xB + c1n) xB + c0n) + yB i
  // (without BGN_PARABOLA: a doubling and an addition entry of two elements each).  The divisions share one
  // inversion (Montgomery's trick over all steps); scratch holds [step][denominator | prefix product].  *ok = 0
  // if some denominator is zero -- only possible when (px, py) is not a point of odd order -- and the table is
  // then unusable (api.cu falls back to the general kernel).  At most nsteps() * 2 elements either way.
  BGN_DEV static void record(const uint32_t* px, const uint32_t* py, uint32_t* lines, uint32_t* scratch, int* ok) {
    typedef F<L> FF;
    Loc<L> X, Y, Z, cR, aR, bI, c3, acc, t;
    Loc<L> t0, t1, t2, t3, t4, t5, t6, t7;
    FF::copy(X.v(), px);
    FF::copy(Y.v(), py);
    FF::copy(Z.v(), c_fc.one);
    FF::copy(acc.v(), c_fc.one);
    const int n = c_pc.naf_len;
    int ns = 0;
    uint32_t* ln = lines;
    auto emit = [&](int nnum, const uint32_t* den) {  // numerators cR, aR[, bI]; denominator den
      uint32_t* sc = scratch + (size_t)ns * 2 * L;
      FF::copy(ln, cR.v());
      FF::copy(ln + L, aR.v());
      if (nnum == 3) FF::copy(ln + 2 * L, bI.v());
      ln += (size_t)nnum * L;
      FF::copy(sc, den);
      FF::copy(sc + L, acc.v());
      FF::mul(acc.v(), acc.v(), den);
      ns++;
    };
    for (int idx = 1; idx < n; idx++) {
      int d = c_pc.naf[idx];
      if (is_dadd(c_pc, idx)) {
        MA::norm1(X.v());
        MA::norm1(Y.v());
        MA::norm1(Z.v());
        G<L>::dadd_para(X.v(), Y.v(), Z.v(), px, py, d < 0, cR.v(), aR.v(), bI.v(), c3.v(), t0.v(), t1.v(), t2.v(), t3.v(),
                        t4.v(), t5.v(), t6.v(), t7.v());
        emit(3, c3.v());
        continue;
      }
      MA::dbl_line(X.v(), Y.v(), Z.v(), cR.v(), aR.v(), bI.v());
      emit(2, bI.v());
      if (d != 0 && idx != n - 1) {
        MA::madd_line(X.v(), Y.v(), Z.v(), px, py, d < 0, cR.v(), aR.v(), bI.v());
        emit(2, bI.v());
      }
    }
    *ok = FF::is_zero(acc.v()) ? 0 : 1;
    if (!*ok) return;
    FF::template inv_gcd<true>(acc.v(), acc.v());
    // backwards over the entries: the same walk, from the last step
    int k = ns - 1;
    for (int idx = n - 1; idx >= 1; idx--) {
      int d = c_pc.naf[idx];
      const int entries = is_dadd(c_pc, idx) ? 1 : ((d != 0 && idx != n - 1) ? 2 : 1);
      for (int q = 0; q < entries; q++, k--) {
        const int nnum = is_dadd(c_pc, idx) ? 3 : 2;
        ln -= (size_t)nnum * L;
        uint32_t* sc = scratch + (size_t)k * 2 * L;
        FF::mul(t.v(), acc.v(), sc + L);    // 1 / den_k
        FF::mul(acc.v(), acc.v(), sc);      // drop den_k from the running inverse
        for (int j = 0; j < nnum; j++) FF::mul(ln + (size_t)j * L, ln + (size_t)j * L, t.v());
      }
    }
  }

  BGN_DEV static void run(const MillerFixedArgs& a, uint32_t* smem, int tid, int nt, size_t e) {
    typedef F<L> FF;
    if (e >= (size_t)a.count) return;
    auto slot = [&](int k) -> E { return smem + ((size_t)k * nt + tid) * L; };
    E fr = slot(S_FR), fi = slot(S_FI), ex = slot(S_EX), ey = slot(S_EY);
    if (a.Einf[e]) {  // e(., O) = 1
      FF::set_one(a.out_re + e * L);
      FF::set_zero(a.out_im + e * L);
      return;
    }
    FF::copy(fr, c_fc.one);
    FF::set_zero(fi);
    FF::copy(ex, a.Ex + e * L);
    FF::copy(ey, a.Ey + e * L);
    const uint32_t* ln = a.lines;
    const int n = c_pc.naf_len;
    auto fold = [&]() {
#if BGN_LINE_LAZY
      M::template line_mul_lazy_f<BGN_LINE_KARATSUBA>(fr, fi, ln, ln + L, ex, ey);
#else
      M::line_mul_f(fr, fi, ln, ln + L, ex, ey);
#endif
      ln += 2 * L;
    };
    BGN_UNROLL1
    for (int idx = 1; idx < n; idx++) {
      if (idx != 1) MA::sqr2(fr, fi);
      if (is_dadd(c_pc, idx)) {
#if BGN_LINE_LAZY
        M::template para_mul_lazy_f<BGN_LINE_KARATSUBA>(fr, fi, ln, ln + L, ln + 2 * L, ex, ey);
#else
        M::para_mul_f(fr, fi, ln, ln + L, ln + 2 * L, ex, ey);
#endif
        ln += 3 * L;
        continue;
      }
      fold();
      if (c_pc.naf[idx] != 0 && idx != n - 1) fold();
    }
    // final exponentiation (conj(f)^2 / N(f))^l, as MillerTeam::finalize for one slot
    E n0 = ex, i0 = ey, g0 = slot(S_G0), g1 = slot(S_G1);
    MA::fe_prepare(fr, fi, n0);
    // 1 / N(f) by the verified division-step GCD (ALU pipe; field.cuh: inv_gcd_fast)
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
    FF::copy(a.out_re + e * L, fr);
    FF::copy(a.out_im + e * L, fi);
  }
};
