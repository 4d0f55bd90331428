// pairlane.cuh -- pairings split over a PAIR OF ADJACENT LANES, for batches too small to give every
// scheduler two warps of one-thread-per-pairing work.
//
// Why (measured, profiles/r01_ops_v27_ncu.txt): a single warp per scheduler issues one IMAD.WIDE per
// ~6 cycles where the pipe takes one per 4, and a batch of 2^14 pairings is 0.87 warps per scheduler
// when a pairing is one thread -- k_miller_fixed sat at 0.52 of the IMAD.WIDE peak there.  Two lanes
// per pairing double the warps and halve the dependent chain of each.
//
// The split is the one lucas.cuh uses: both lanes run the SAME instruction stream (no divergence),
// pick their operands with selects, and swap results with one warp shuffle per limb.  Everything is
// register-resident; nothing lives in shared or local memory.
//
//   e(C, P) through the recorded line table of P (MillerFixedPair; replaces k_miller_fixed below a
//   batch-size threshold chosen by measurement, api.cu: run_miller_fixed).  The table's lines are
//   normalised by their third coefficient (pairing.cuh: MillerFixed::record), so the line of step k at
//   the evaluation point (xB, yB) is l = (cRn_k + aRn_k xB) + yB i: only its real part costs a product,
//   and it does not depend on the accumulator -- the two lanes evaluate the lines of TWO consecutive
//   steps in one round:
//
//                       lane 0                              lane 1
//     lines (every      l0_k = cRn_k + aRn_k xB            l0_k+1 = cRn_k+1 + aRn_k+1 xB      1 product each
//      other step)
//     f^2            (f0 + f1)(f0 - f1)                   2 f0 f1                  1 product each
//     f * l          re = f0 l0 + (-f1) yB                im = f0 yB + f1 l0       1 dot product each
//
//   1.5 (2L^2 + L) + (3L^2 + L) = 1776 products per lane and doubling step at L = 17 against the
//   3264 one thread spends (fused.cuh: sqr2 + line_mul_lazy_f), 2.5 shuffle exchanges per step.
//   A doubling-and-addition step is ONE table entry, a parabola [csn | c1n | c0n] (pairing.cuh:
//   MillerFixed::record): lane 0 forms csn xB^2 + c0n, lane 1 c1n xB, their sum is the real part, the
//   imaginary part is yB -- 2 (2L^2 + L) + (3L^2 + L) per lane where a doubling step followed by an
//   addition step cost 2 (2L^2 + L) + 2 (3L^2 + L).
//   The dot product (arith.cuh: Fp::dot2) accumulates both multiplicands row by row into one CIOS
//   window, so the F_p^2 product needs no double-width temporaries.
//
// Replaces the same libpbc behaviour as pairing.cuh (a1_pairing with a fixed first argument,
// reached from bgn.go:316-321 makeL2, poly.go:159-163 MakePolyL2 and the level-1 branch of
// bgn.go:218-250 decrypt).
#pragma once
#include "lucas.cuh"

template <int L>
struct MillerFixedPair {
  typedef Fp<L> P;
  typedef Lucas<L> LU;
  struct State {
    uint32_t f0[L], f1[L];  // accumulator, both coordinates in both lanes (relaxed range, below 4p)
    uint32_t ex[L], ey[L];  // the evaluation point, in both lanes
    uint32_t e2[L];         // lane 0: xB^2, lane 1: xB (the multiplicand of this lane's half of a parabola's real part)
  };

  BGN_DEV static void init(State& st, const uint32_t* ex, const uint32_t* ey, int s, bool active) {
    if (active) {
      ld<L>(st.ex, ex);
      ld<L>(st.ey, ey);
    } else {
      BGN_SETB(st.ex, 0.0);
      BGN_SETB(st.ey, 0.0);
      BGN_UNROLL
      for (int j = 0; j < L; j++) st.ex[j] = st.ey[j] = 0;
    }
    ld<L>(st.f0, c_fc.one);
    BGN_SETB(st.f1, 0.0);
    BGN_UNROLL
    for (int j = 0; j < L; j++) st.f1[j] = 0;
    if (MillerFixed<L>::PARA) {
      uint32_t sq[L];
      P::mul(sq, st.ex, st.ex);
      LU::sel(st.e2, s == 0, sq, st.ex);
    }
  }
  // this lane's half of f^2.  in: f < 4p.  out: < 3p.
  BGN_DEV static void sqr_half(uint32_t (&t)[L], const State& st, int s) {
    uint32_t sum[L], dif[L], x[L], y[L], d[L];
    P::addn(sum, st.f0, st.f1);
    P::subk(dif, st.f0, st.f1, c_fc.p4, 4);
    LU::sel(x, s == 0, sum, st.f0);
    LU::sel(y, s == 0, dif, st.f1);
    P::mul(t, x, y);
    P::addn(d, t, t);
    LU::sel(t, s == 0, t, d);  // lane 1 holds 2 f0 f1
  }
  // real part of the normalised line ln = [cRn | aRn] at the evaluation point: cRn + aRn xB.  out: < 4p.
  BGN_DEV static void eval_line(uint32_t (&t)[L], const State& st, const uint32_t* ln) {
    uint32_t c[L];
    P::mul_stream(t, st.ex, ln + L);
    ld<L>(c, ln);
    P::addn(t, t, c);
  }
  // this lane's half of the real part of the normalised parabola ln = [csn | c1n | c0n] at the evaluation
  // point: lane 0 csn xB^2 + c0n, lane 1 c1n xB; the sum of the halves is the real part.  out: < 4p.
  BGN_DEV static void eval_para_half(uint32_t (&t)[L], const State& st, const uint32_t* ln, int s) {
    uint32_t c[L], z[L];
    P::mul_stream(t, st.e2, ln + (s == 0 ? 0 : L));
    ld<L>(c, ln + 2 * L);
    BGN_SETB(z, 0.0);
    BGN_UNROLL
    for (int j = 0; j < L; j++) z[j] = 0;
    LU::sel(c, s == 0, c, z);
    P::addn(t, t, c);
  }
  // this lane's coordinate of f * (l0 + l1 i), `mine` being the coordinate of l this lane evaluated:
  // lane 0 f0 l0 + (4p - f1) l1, lane 1 f0 l1 + f1 l0.  in: f < 4p, l < 8p.  out: < 2p.
  BGN_DEV static void mul_half(uint32_t (&t)[L], const State& st, const uint32_t (&mine)[L], const uint32_t (&other)[L],
                               int s) {
    uint32_t nf1[L], a1[L];
    P::negk(nf1, st.f1, c_fc.p4, 4);
    LU::sel(a1, s == 0, nf1, st.f1);
    P::dot2(t, st.f0, mine, a1, other);
  }
  BGN_DEV static void update(State& st, const uint32_t (&mine)[L], const uint32_t (&other)[L], int s) {
    LU::sel(st.f0, s == 0, mine, other);
    LU::sel(st.f1, s == 0, other, mine);
  }

  // ---- final exponentiation (conj(f)^2 / N(f))^l on the pair
  // stage 1: lane 0 f0^2, lane 1 f1^2
  BGN_DEV static void fe_sq(uint32_t (&t)[L], const State& st, int s) {
    uint32_t x[L];
    LU::sel(x, s == 0, st.f0, st.f1);
    P::mul(t, x, x);  // (once per pairing: the dedicated squaring's two double-width arrays would cost this kernel spills)
  }
  // stage 2 (after the exchange: u = f0^2, v = f1^2 in both lanes): both lanes form N = u + v, invert it
  // (verified division-step GCD, ALU pipe) and scale their coordinate of conj(f)^2 = (u - v) - 2 f0 f1 i.
  // out: lane 0 Re(g), lane 1 Im(g), below 2p.
  BGN_DEV static void fe_scale(uint32_t (&t)[L], const State& st, const uint32_t (&mine)[L], const uint32_t (&other)[L],
                               int s) {
    uint32_t u[L], v[L], nrm[L], re[L], w[L], im[L], inv[L], c[L];
    LU::sel(u, s == 0, mine, other);
    LU::sel(v, s == 0, other, mine);
    P::addn(nrm, u, v);
    P::subk(re, u, v, c_fc.p2, 2);
    P::mul(w, st.f0, st.f1);
    P::addn(w, w, w);
    P::negk(im, w, c_fc.p4, 4);
    F<L>::template inv_gcd_fast<true>(inv, nrm);
    LU::sel(c, s == 0, re, im);
    P::mul(t, c, inv);
  }
  // this lane's coordinate of f * g for g = (g0, g1) held by both lanes
  BGN_DEV static void mulg_half(uint32_t (&t)[L], const State& st, const uint32_t (&g0)[L], const uint32_t (&g1)[L],
                                int s) {
    uint32_t nf1[L], a1[L], b0[L], b1[L];
    P::negk(nf1, st.f1, c_fc.p4, 4);
    LU::sel(a1, s == 0, nf1, st.f1);
    LU::sel(b0, s == 0, g0, g1);
    LU::sel(b1, s == 0, g1, g0);
    P::dot2(t, st.f0, b0, a1, b1);
  }
  // lane s stores its coordinate (re for lane 0, im for lane 1), back in [0, 2p)
  BGN_DEV static void finish(const State& S, uint32_t* ore, uint32_t* oim, int s) {
    uint32_t t[L];
    LU::sel(t, s == 0, S.f0, S.f1);
    P::norm2p(t, t);
    st<L>(s == 0 ? ore : oim, t);
  }
  BGN_DEV static void finish_one(uint32_t* ore, uint32_t* oim, int s) {  // e(., O) = 1
    uint32_t t[L];
    if (s == 0) {
      ld<L>(t, c_fc.one);
    } else {
      BGN_SETB(t, 0.0);
      BGN_UNROLL
      for (int j = 0; j < L; j++) t[j] = 0;
    }
    st<L>(s == 0 ? ore : oim, t);
  }

  // The whole program of one lane; xchg(other, mine) hands every lane its partner's `mine`.  The CUDA
  // kernel passes a warp-shuffle exchange; the CPU simulation runs the two lanes as two threads of
  // control meeting at each exchange (tests/hostsim).
  template <typename Xchg>
  BGN_DEV static void run(const MillerFixedArgs& a, size_t e, int s, bool active, Xchg xchg) {
    const size_t ee = active ? e : 0;
    const bool inf = active && a.Einf[ee] != 0;
    State st;
    init(st, a.Ex + ee * L, a.Ey + ee * L, s, active && !inf);
    const uint32_t* ln = a.lines;
    const int n = c_pc.naf_len;
    int left = 0;  // lines still to fold
    for (int idx = 1; idx < n; idx++) left += (c_pc.naf[idx] != 0 && idx != n - 1) ? 2 : 1;
    uint32_t mine[L], other[L], lm[L], lo[L], cur[L], nxt[L];
    bool have = false;  // nxt holds the real part of the next line (uniform over the batch: one key)
    BGN_UNROLL1
    for (int idx = 1; idx < n; idx++) {
      const int folds = (c_pc.naf[idx] != 0 && idx != n - 1) ? 2 : 1;
      if (idx != 1) {
        sqr_half(mine, st, s);
        xchg(other, mine);
        update(st, mine, other, s);
      }
      if (MillerFixed<L>::is_dadd(c_pc, idx)) {
        // doubling-and-addition step: one parabola entry, its real part evaluated in halves by the two lanes
        eval_para_half(lm, st, ln, s);
        xchg(lo, lm);
        P::addn(cur, lm, lo);
        LU::sel(lm, s == 0, cur, st.ey);
        LU::sel(lo, s == 0, st.ey, cur);
        mul_half(mine, st, lm, lo, s);
        xchg(other, mine);
        update(st, mine, other, s);
        ln += 3 * L;
        left -= 2;
        continue;
      }
      BGN_UNROLL1
      for (int k = 0; k < folds; k++) {
        if (!have) {  // lane 0 evaluates this line, lane 1 the next entry if that is a line too
          const bool pair = MillerFixed<L>::PARA ? (idx + 1 < n && !MillerFixed<L>::is_dadd(c_pc, idx + 1)) : left > 1;
          eval_line(lm, st, ln + (pair ? s : 0) * 2 * L);
          xchg(lo, lm);
          LU::sel(cur, s == 0, lm, lo);
          LU::sel(nxt, s == 0, lo, lm);
          have = pair;
        } else {
          LU::sel(cur, true, nxt, nxt);
          have = false;
        }
        // l = cur + yB i: lane 0 multiplies by (cur, yB), lane 1 by (yB, cur)
        LU::sel(lm, s == 0, cur, st.ey);
        LU::sel(lo, s == 0, st.ey, cur);
        mul_half(mine, st, lm, lo, s);
        xchg(other, mine);
        update(st, mine, other, s);
        ln += 2 * L;
        left--;
      }
    }
    fe_sq(mine, st, s);
    xchg(other, mine);
    fe_scale(lm, st, mine, other, s);
    xchg(lo, lm);
    uint32_t g0[L], g1[L];
    LU::sel(g0, s == 0, lm, lo);
    LU::sel(g1, s == 0, lo, lm);
    LU::sel(st.f0, true, g0, g0);
    LU::sel(st.f1, true, g1, g1);
    const uint64_t l = c_pc.l;
    int top = 63;
    while (top > 0 && !((l >> top) & 1)) top--;
    BGN_UNROLL1
    for (int bit = top - 1; bit >= 0; bit--) {
      sqr_half(mine, st, s);
      xchg(other, mine);
      update(st, mine, other, s);
      if ((l >> bit) & 1) {
        mulg_half(mine, st, g0, g1, s);
        xchg(other, mine);
        update(st, mine, other, s);
      }
    }
    if (!active) return;
    if (inf)
      finish_one(a.out_re + e * L, a.out_im + e * L, s);
    else
      finish(st, a.out_re + e * L, a.out_im + e * L, s);
  }
};
