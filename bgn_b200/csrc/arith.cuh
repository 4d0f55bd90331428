// arith.cuh -- multi-limb Montgomery arithmetic over the type-A1 field F_p
// (p = l*n - 1) on 32-bit limbs, written for sm_100a.
//
// Replaces libpbc montfp.c / GMP mpn_* ("native #2/#3");
// nothing here is derived from their sources.  One 32x32->64 product is one
// IMAD.WIDE.U32[.X] issue slot: every (mad.lo.cc, madc.hi.cc) pair below lands
// on an aligned 64-bit register pair so ptxas fuses it, and carries travel in
// predicate registers so independent chains interleave (checked with
// cuobjdump -sass; see DESIGN.md "K1").
//
// Representation: L limbs, little-endian limb order, Montgomery form with
// R = 2^(32L).  Values are kept in the redundant range [0, 2p) ("lazy" form);
// fp_canon() brings them to [0, p).  Requires 32L >= bits(p) + 3.
//
// When BGN_HOSTSIM is defined the same code compiles as plain C++ with the
// carry flag emulated in software.  That build exists ONLY for the CPU-side
// unit tests of the device logic (tests/hostsim); the shipped library never
// contains it.
#pragma once
#include <stdint.h>
#include "types.h"

#ifdef BGN_HOSTSIM
#define BGN_DEV inline
#define BGN_HD inline
#define BGN_DEVNI
#define BGN_CONST static
#define BGN_UNROLL
#define BGN_UNROLL1
#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <unordered_map>
namespace bgnsim {
static thread_local uint32_t cc = 0;
static uint64_t nmul = 0;  // Montgomery products executed (work model check, tests only)
static uint64_t nmulw = 0, nredc = 0;  // double-width products / separate reductions executed
static uint64_t nmulk = 0;             // of which Karatsuba products (counted apart from nmulw)
static uint64_t nsqrw = 0;             // dedicated double-width squarings (L (L + 1) / 2 products each)
static uint64_t ndot3 = 0;             // one-pass three-term dot products (4L^2 + L products each)
static uint64_t ndot2 = 0;             // one-pass dot products a0 b0 + a1 b1 (3L^2 + L products each)
static uint64_t safegcd_fallbacks = 0;  // F<L>::inv_gcd<SAFE>: times the verified fast inversion fell back
// Range tracker (tests only): every element written by the arithmetic below carries an upper
// bound in multiples of p, keyed by its address, so the CPU run of the device programs PROVES
// (by worst-case interval propagation, not by the sampled values) that the relaxed-range code
// never overflows a product and never lets a difference go negative.
static std::unordered_map<const void*, double> bnd;
static double headroom = 128.0;  // floor(2^(32L) / p), set with the constants
// double-width values (lazy reduction): value < bw * p^2 + kk * p * R
struct WB {
  double bw, kk;
};
static std::unordered_map<const void*, WB> bndw;
static double max_bound = 0.0;   // largest bound ever attached (reported to the tests)
static uint64_t unknown = 0;     // reads of untracked addresses (treated as < 2p)
static uint64_t violations = 0;
inline double getb(const void* a) {
  auto it = bnd.find(a);
  if (it == bnd.end()) {
    unknown++;
    return 2.0;
  }
  return it->second;
}
inline void setb(const void* a, double b) {
  bnd[a] = b;
  if (b > max_bound) max_bound = b;
}
inline void check(bool ok, const char* what) {
  if (!ok) {
    if (violations++ < 10) fprintf(stderr, "bgnsim: range violation: %s\n", what);
  }
}
}
#define BGN_SETB(a, b) bgnsim::setb((const void*)(a), (b))
#define BGN_GETB(a) bgnsim::getb((const void*)(a))
#define BGN_CHECK(c, w) bgnsim::check((c), (w))
namespace bgnsim {
inline WB getw(const void* a) {
  auto it = bndw.find(a);
  if (it == bndw.end()) {
    unknown++;
    return WB{4.0, 0.0};
  }
  return it->second;
}
inline void setw(const void* a, double bw, double kk) {
  bndw[a] = WB{bw, kk};
  // must fit 2L limbs: bw p^2 + kk p R < R^2  <=>  bw / H^2 + kk / H < 1
  check(bw / (headroom * headroom) + kk / headroom < 1.0, "double-width value too large");
}
}
BGN_DEV void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
}
BGN_DEV void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  unsigned __int128 t = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)hi << 32) | lo);
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
  bgnsim::cc = (uint32_t)(t >> 64);
}
BGN_DEV void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  unsigned __int128 t = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)hi << 32) | lo) + bgnsim::cc;
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
  bgnsim::cc = (uint32_t)(t >> 64);
}
BGN_DEV void madc_wide_cc3(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  unsigned __int128 t = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)chi << 32) | clo) + bgnsim::cc;
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
  bgnsim::cc = (uint32_t)(t >> 64);
}
BGN_DEV void add_cc(uint32_t& r, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b;
  r = (uint32_t)t;
  bgnsim::cc = (uint32_t)(t >> 32);
}
BGN_DEV void addc_cc(uint32_t& r, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b + bgnsim::cc;
  r = (uint32_t)t;
  bgnsim::cc = (uint32_t)(t >> 32);
}
BGN_DEV void addc(uint32_t& r, uint32_t a, uint32_t b) {
  r = a + b + bgnsim::cc;
}
BGN_DEV void sub_cc(uint32_t& r, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b;
  r = (uint32_t)t;
  bgnsim::cc = (uint32_t)((t >> 32) & 1);  // PTX ISA: CC.CF = borrow-out
}
BGN_DEV void subc_cc(uint32_t& r, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b - bgnsim::cc;
  r = (uint32_t)t;
  bgnsim::cc = (uint32_t)((t >> 32) & 1);
}
BGN_DEV void subc(uint32_t& r, uint32_t a, uint32_t b) {
  r = a - b - bgnsim::cc;
}
#else
#define BGN_SETB(a, b)
#define BGN_GETB(a) 0.0
#define BGN_CHECK(c, w)
#define BGN_DEV __device__ __forceinline__
#define BGN_HD __host__ __device__ __forceinline__
#define BGN_DEVNI __device__ __noinline__
#define BGN_CONST __constant__
#define BGN_UNROLL _Pragma("unroll")
#define BGN_UNROLL1 _Pragma("unroll 1")
BGN_DEV void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
BGN_DEV void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
BGN_DEV void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
BGN_DEV void madc_wide_cc3(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
  asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
               : "=&r"(lo), "=r"(hi)
               : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
BGN_DEV void add_cc(uint32_t& r, uint32_t a, uint32_t b) {
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
}
BGN_DEV void addc_cc(uint32_t& r, uint32_t a, uint32_t b) {
  asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
}
BGN_DEV void addc(uint32_t& r, uint32_t a, uint32_t b) {
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
}
BGN_DEV void sub_cc(uint32_t& r, uint32_t a, uint32_t b) {
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
}
BGN_DEV void subc_cc(uint32_t& r, uint32_t a, uint32_t b) {
  asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
}
BGN_DEV void subc(uint32_t& r, uint32_t a, uint32_t b) {
  asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
}
#endif

// ---------------------------------------------------------------------------
// Field constants.  One key is active per device at a time; the host uploads
// these before launching (api.cu: ctx_activate).  Limb counts up to 33
// (1024-bit keys) -- arrays are padded to BGN_MAXL.
// ---------------------------------------------------------------------------
BGN_CONST FieldConsts c_fc;

// ---------------------------------------------------------------------------
// Register-level primitives.  All loops have compile-time bounds.
// ---------------------------------------------------------------------------
template <int L>
struct Fp {
  static constexpr int W = 2 * ((L + 2) / 2);  // accumulator registers per parity array
  static constexpr int KE = (L + 1) / 2;       // even-index limbs of an operand
  static constexpr int KO = L / 2;             // odd-index limbs

#ifdef BGN_HOSTSIM
  static void trk_mul(const void* r, const void* a, const void* b) {
    double A = BGN_GETB(a), B = BGN_GETB(b);
    BGN_CHECK(A + 1.0 <= bgnsim::headroom && B <= bgnsim::headroom, "product operand too large");
    BGN_SETB(r, A * B / bgnsim::headroom + 1.0);
    bgnsim::nmul++;
  }
#else
  BGN_DEV static void trk_mul(const void*, const void*, const void*) {}
#endif

  // one CIOS row: acc += a*s (then caller reduces).  X is the array aligned at
  // limb 0, Y sits one limb higher; on entry (not FIRST) Y is the previous
  // row's X whose limb 0 is zero and whose limb 1 is the pending carry limb.
  template <bool FIRST>
  BGN_DEV static void row(uint32_t (&X)[W], uint32_t (&Y)[W], const uint32_t (&a)[L], uint32_t s,
                          const uint32_t* __restrict__ pm, uint32_t np0) {
    if (FIRST) {
      BGN_UNROLL
      for (int k = 0; k < KO; k++) mul_wide(Y[2 * k], Y[2 * k + 1], a[2 * k + 1], s);
      Y[W - 2] = 0;
      Y[W - 1] = 0;
      BGN_UNROLL
      for (int k = 0; k < KE; k++) mul_wide(X[2 * k], X[2 * k + 1], a[2 * k], s);
      if (2 * KE < W) {
        X[W - 2] = 0;
        X[W - 1] = 0;
      }
    } else {
      add_cc(X[0], X[0], Y[1]);
      BGN_UNROLL
      for (int k = 0; k < KO; k++) madc_wide_cc3(Y[2 * k], Y[2 * k + 1], a[2 * k + 1], s, Y[2 * k + 2], Y[2 * k + 3]);
      addc(Y[W - 2], 0, 0);
      Y[W - 1] = 0;
      mad_wide_cc(X[0], X[1], a[0], s);
      BGN_UNROLL
      for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], a[2 * k], s);
      if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
    }
    uint32_t m = X[0] * np0;
    mad_wide_cc(Y[0], Y[1], pm[1], m);
    BGN_UNROLL
    for (int k = 1; k < KO; k++) madc_wide_cc(Y[2 * k], Y[2 * k + 1], pm[2 * k + 1], m);
    addc(Y[W - 2], Y[W - 2], 0);
    mad_wide_cc(X[0], X[1], pm[0], m);
    BGN_UNROLL
    for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], pm[2 * k], m);
    if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
  }

  // r = a*b/R mod p, r in [0,2p) for a,b in [0,2p).  2L^2+L products.
  BGN_DEV static void mul(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    uint32_t X[W], Y[W];
    trk_mul(r, a, b);
    const uint32_t* pm = c_fc.p;
    const uint32_t np0 = c_fc.np0;
    row<true>(X, Y, a, b[0], pm, np0);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      row<false>(Y, X, a, b[i], pm, np0);
      row<false>(X, Y, a, b[i + 1], pm, np0);
    }
    if ((L & 1) == 0) {
      row<false>(Y, X, a, b[L - 1], pm, np0);
      merge(r, X, Y);  // last row had (aligned=Y, other=X): after shift X is aligned
    } else {
      merge(r, Y, X);
    }
  }

  // Same product with the multiplier streamed from memory: row i reads b[i] when it needs it, so b
  // never occupies registers and its loads overlap the previous rows.
  // (ES = element stride of the memory operand: limb i at bp[i * ES])
  template <int ES = 1>
  BGN_DEV static void mul_stream(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t* bp) {
    uint32_t X[W], Y[W];
    trk_mul(r, a, bp);
    const uint32_t* pm = c_fc.p;
    const uint32_t np0 = c_fc.np0;
    row<true>(X, Y, a, bp[0], pm, np0);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      row<false>(Y, X, a, bp[i * ES], pm, np0);
      row<false>(X, Y, a, bp[(i + 1) * ES], pm, np0);
    }
    if ((L & 1) == 0) {
      row<false>(Y, X, a, bp[(L - 1) * ES], pm, np0);
      merge(r, X, Y);
    } else {
      merge(r, Y, X);
    }
  }

  // Two independent products with their rows interleaved in program order (ILP 2): the carry
  // chains of one product fill the dependency gaps of the other.
  BGN_DEV static void mul_pair(uint32_t (&r1)[L], const uint32_t (&a1)[L], const uint32_t* b1, uint32_t (&r2)[L],
                               const uint32_t (&a2)[L], const uint32_t* b2) {
    uint32_t X1[W], Y1[W], X2[W], Y2[W];
    trk_mul(r1, a1, b1);
    trk_mul(r2, a2, b2);
    const uint32_t* pm = c_fc.p;
    const uint32_t np0 = c_fc.np0;
    row<true>(X1, Y1, a1, b1[0], pm, np0);
    row<true>(X2, Y2, a2, b2[0], pm, np0);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      row<false>(Y1, X1, a1, b1[i], pm, np0);
      row<false>(Y2, X2, a2, b2[i], pm, np0);
      row<false>(X1, Y1, a1, b1[i + 1], pm, np0);
      row<false>(X2, Y2, a2, b2[i + 1], pm, np0);
    }
    if ((L & 1) == 0) {
      row<false>(Y1, X1, a1, b1[L - 1], pm, np0);
      row<false>(Y2, X2, a2, b2[L - 1], pm, np0);
      merge(r1, X1, Y1);
      merge(r2, X2, Y2);
    } else {
      merge(r1, Y1, X1);
      merge(r2, Y2, X2);
    }
  }

  // ---- dot product r = (a0*b0 + a1*b1) / R mod p in ONE pass: every CIOS row accumulates both
  // multiplicands before its reduction half, so the sum of two products costs 3L^2 + L products (what
  // two double-width products sharing one reduction cost) without a double-width temporary: the
  // accumulator window is the same W + W registers as one product's.  Used where an F_p^2 product is
  // split over a lane pair (pairlane.cuh): re = f0 l0 + (-f1) l1 on one lane, im = f0 l1 + f1 l0 on the
  // other.  Requires a0 + a1 + p < R / 2 (window head-room); r < (a0 b0 + a1 b1) / R + p.
  template <bool FIRST>
  BGN_DEV static void row2(uint32_t (&X)[W], uint32_t (&Y)[W], const uint32_t (&a0)[L], uint32_t s0,
                           const uint32_t (&a1)[L], uint32_t s1, const uint32_t* __restrict__ pm, uint32_t np0) {
    if (FIRST) {
      BGN_UNROLL
      for (int k = 0; k < KO; k++) mul_wide(Y[2 * k], Y[2 * k + 1], a0[2 * k + 1], s0);
      Y[W - 2] = 0;
      Y[W - 1] = 0;
      BGN_UNROLL
      for (int k = 0; k < KE; k++) mul_wide(X[2 * k], X[2 * k + 1], a0[2 * k], s0);
      if (2 * KE < W) {
        X[W - 2] = 0;
        X[W - 1] = 0;
      }
    } else {
      add_cc(X[0], X[0], Y[1]);
      BGN_UNROLL
      for (int k = 0; k < KO; k++) madc_wide_cc3(Y[2 * k], Y[2 * k + 1], a0[2 * k + 1], s0, Y[2 * k + 2], Y[2 * k + 3]);
      addc(Y[W - 2], 0, 0);
      Y[W - 1] = 0;
      mad_wide_cc(X[0], X[1], a0[0], s0);
      BGN_UNROLL
      for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], a0[2 * k], s0);
      if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
    }
    if (KO > 0) {
      mad_wide_cc(Y[0], Y[1], a1[1], s1);
      BGN_UNROLL
      for (int k = 1; k < KO; k++) madc_wide_cc(Y[2 * k], Y[2 * k + 1], a1[2 * k + 1], s1);
      addc(Y[W - 2], Y[W - 2], 0);
    }
    mad_wide_cc(X[0], X[1], a1[0], s1);
    BGN_UNROLL
    for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], a1[2 * k], s1);
    if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
    uint32_t m = X[0] * np0;
    mad_wide_cc(Y[0], Y[1], pm[1], m);
    BGN_UNROLL
    for (int k = 1; k < KO; k++) madc_wide_cc(Y[2 * k], Y[2 * k + 1], pm[2 * k + 1], m);
    addc(Y[W - 2], Y[W - 2], 0);
    mad_wide_cc(X[0], X[1], pm[0], m);
    BGN_UNROLL
    for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], pm[2 * k], m);
    if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
  }
  BGN_DEV static void dot2(uint32_t (&r)[L], const uint32_t (&a0)[L], const uint32_t (&b0)[L], const uint32_t (&a1)[L],
                           const uint32_t (&b1)[L]) {
    uint32_t X[W], Y[W];
#ifdef BGN_HOSTSIM
    {
      double A0 = BGN_GETB(a0), B0 = BGN_GETB(b0), A1 = BGN_GETB(a1), B1 = BGN_GETB(b1);
      BGN_CHECK(2.0 * (A0 + A1 + 1.0) <= bgnsim::headroom, "dot product multiplicands too large");
      BGN_CHECK(B0 <= bgnsim::headroom && B1 <= bgnsim::headroom, "dot product multiplier too large");
      BGN_SETB(r, (A0 * B0 + A1 * B1) / bgnsim::headroom + 1.0);
      bgnsim::ndot2++;
    }
#endif
    const uint32_t* pm = c_fc.p;
    const uint32_t np0 = c_fc.np0;
    row2<true>(X, Y, a0, b0[0], a1, b1[0], pm, np0);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      row2<false>(Y, X, a0, b0[i], a1, b1[i], pm, np0);
      row2<false>(X, Y, a0, b0[i + 1], a1, b1[i + 1], pm, np0);
    }
    if ((L & 1) == 0) {
      row2<false>(Y, X, a0, b0[L - 1], a1, b1[L - 1], pm, np0);
      merge(r, X, Y);
    } else {
      merge(r, Y, X);
    }
  }

  // The dot product with both multipliers streamed from memory (limb i at bp[i * ES]), as mul_stream does for
  // the product: the multiplicands a0, a1 stay in registers.  Used by the Miller loop's line evaluation at a
  // normalised point (fused.cuh: line_mul_n): cR * (1/yB) + aR * (xB/yB) in one pass.
  template <int ES = 1>
  BGN_DEV static void dot2_stream(uint32_t (&r)[L], const uint32_t (&a0)[L], const uint32_t* b0, const uint32_t (&a1)[L],
                                  const uint32_t* b1) {
    uint32_t X[W], Y[W];
#ifdef BGN_HOSTSIM
    {
      double A0 = BGN_GETB(a0), B0 = BGN_GETB(b0), A1 = BGN_GETB(a1), B1 = BGN_GETB(b1);
      BGN_CHECK(2.0 * (A0 + A1 + 1.0) <= bgnsim::headroom, "dot product multiplicands too large");
      BGN_CHECK(B0 <= bgnsim::headroom && B1 <= bgnsim::headroom, "dot product multiplier too large");
      BGN_SETB(r, (A0 * B0 + A1 * B1) / bgnsim::headroom + 1.0);
      bgnsim::ndot2++;
    }
#endif
    const uint32_t* pm = c_fc.p;
    const uint32_t np0 = c_fc.np0;
    row2<true>(X, Y, a0, b0[0], a1, b1[0], pm, np0);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      row2<false>(Y, X, a0, b0[i * ES], a1, b1[i * ES], pm, np0);
      row2<false>(X, Y, a0, b0[(i + 1) * ES], a1, b1[(i + 1) * ES], pm, np0);
    }
    if ((L & 1) == 0) {
      row2<false>(Y, X, a0, b0[(L - 1) * ES], a1, b1[(L - 1) * ES], pm, np0);
      merge(r, X, Y);
    } else {
      merge(r, Y, X);
    }
  }

  // ---- three-term dot product r = (a0*b0 + a1*b1 + a2*b2) / R mod p in one CIOS pass: 4L^2 + L products.
  // Multiplicands in registers, multipliers streamed from memory.  Used by the Miller loop's parabola step
  // (fused.cuh: para_mul): s (x^2/y) + c1 (x/y) + c0 (1/y) at a normalised evaluation point.
  // Requires a0 + a1 + a2 + p < R / 2 (window head-room).
  template <bool FIRST>
  BGN_DEV static void row3(uint32_t (&X)[W], uint32_t (&Y)[W], const uint32_t (&a0)[L], uint32_t s0,
                           const uint32_t (&a1)[L], uint32_t s1, const uint32_t (&a2)[L], uint32_t s2,
                           const uint32_t* __restrict__ pm, uint32_t np0) {
    if (FIRST) {
      BGN_UNROLL
      for (int k = 0; k < KO; k++) mul_wide(Y[2 * k], Y[2 * k + 1], a0[2 * k + 1], s0);
      Y[W - 2] = 0;
      Y[W - 1] = 0;
      BGN_UNROLL
      for (int k = 0; k < KE; k++) mul_wide(X[2 * k], X[2 * k + 1], a0[2 * k], s0);
      if (2 * KE < W) {
        X[W - 2] = 0;
        X[W - 1] = 0;
      }
    } else {
      add_cc(X[0], X[0], Y[1]);
      BGN_UNROLL
      for (int k = 0; k < KO; k++) madc_wide_cc3(Y[2 * k], Y[2 * k + 1], a0[2 * k + 1], s0, Y[2 * k + 2], Y[2 * k + 3]);
      addc(Y[W - 2], 0, 0);
      Y[W - 1] = 0;
      mad_wide_cc(X[0], X[1], a0[0], s0);
      BGN_UNROLL
      for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], a0[2 * k], s0);
      if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
    }
    BGN_UNROLL
    for (int t = 0; t < 2; t++) {
      const uint32_t (&a)[L] = t == 0 ? a1 : a2;
      const uint32_t s = t == 0 ? s1 : s2;
      if (KO > 0) {
        mad_wide_cc(Y[0], Y[1], a[1], s);
        BGN_UNROLL
        for (int k = 1; k < KO; k++) madc_wide_cc(Y[2 * k], Y[2 * k + 1], a[2 * k + 1], s);
        addc(Y[W - 2], Y[W - 2], 0);
      }
      mad_wide_cc(X[0], X[1], a[0], s);
      BGN_UNROLL
      for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], a[2 * k], s);
      if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
    }
    uint32_t m = X[0] * np0;
    mad_wide_cc(Y[0], Y[1], pm[1], m);
    BGN_UNROLL
    for (int k = 1; k < KO; k++) madc_wide_cc(Y[2 * k], Y[2 * k + 1], pm[2 * k + 1], m);
    addc(Y[W - 2], Y[W - 2], 0);
    mad_wide_cc(X[0], X[1], pm[0], m);
    BGN_UNROLL
    for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], pm[2 * k], m);
    if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
  }
  template <int ES = 1>
  BGN_DEV static void dot3_stream(uint32_t (&r)[L], const uint32_t (&a0)[L], const uint32_t* b0, const uint32_t (&a1)[L],
                                  const uint32_t* b1, const uint32_t (&a2)[L], const uint32_t* b2) {
    uint32_t X[W], Y[W];
#ifdef BGN_HOSTSIM
    {
      double A0 = BGN_GETB(a0), B0 = BGN_GETB(b0), A1 = BGN_GETB(a1), B1 = BGN_GETB(b1), A2 = BGN_GETB(a2), B2 = BGN_GETB(b2);
      BGN_CHECK(2.0 * (A0 + A1 + A2 + 1.0) <= bgnsim::headroom, "dot product multiplicands too large");
      BGN_CHECK(B0 <= bgnsim::headroom && B1 <= bgnsim::headroom && B2 <= bgnsim::headroom, "dot product multiplier too large");
      BGN_SETB(r, (A0 * B0 + A1 * B1 + A2 * B2) / bgnsim::headroom + 1.0);
      bgnsim::ndot3++;
    }
#endif
    const uint32_t* pm = c_fc.p;
    const uint32_t np0 = c_fc.np0;
    row3<true>(X, Y, a0, b0[0], a1, b1[0], a2, b2[0], pm, np0);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      row3<false>(Y, X, a0, b0[i * ES], a1, b1[i * ES], a2, b2[i * ES], pm, np0);
      row3<false>(X, Y, a0, b0[(i + 1) * ES], a1, b1[(i + 1) * ES], a2, b2[(i + 1) * ES], pm, np0);
    }
    if ((L & 1) == 0) {
      row3<false>(Y, X, a0, b0[(L - 1) * ES], a1, b1[(L - 1) * ES], a2, b2[(L - 1) * ES], pm, np0);
      merge(r, X, Y);
    } else {
      merge(r, Y, X);
    }
  }

  // Same product as mul_stream with the rows in a NON-unrolled loop of 2U rows per iteration (the
  // accumulators swap roles every row and move down one register pair every two rows, which a
  // loop pays for with ~2L register moves per iteration; unrolled code renames for free): 1 + 2U
  // rows of code instead of L, so many products can be fused into one routine without leaving the
  // instruction cache.  The multiplicand a stays in registers, the multiplier is read from memory
  // row by row (dynamic index).
  template <int U, int ES = 1>
  BGN_DEV static void mul_loop(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t* bp) {
    uint32_t X[W], Y[W];
    trk_mul(r, a, bp);
    const uint32_t* pm = c_fc.p;
    const uint32_t np0 = c_fc.np0;
    constexpr int NI = (L - 1) / (2 * U);  // whole iterations
    constexpr int T0 = 1 + NI * 2 * U;     // first row of the unrolled tail
    row<true>(X, Y, a, bp[0], pm, np0);
    if (NI > 0) {
      const uint32_t* q = bp + ES;
      BGN_UNROLL1
      for (int it = 0; it < NI; it++) {
        BGN_UNROLL
        for (int k = 0; k < U; k++) {
          row<false>(Y, X, a, q[(2 * k) * ES], pm, np0);
          row<false>(X, Y, a, q[(2 * k + 1) * ES], pm, np0);
        }
        q += 2 * U * ES;
      }
    }
    BGN_UNROLL
    for (int i = T0; i + 1 < L; i += 2) {
      row<false>(Y, X, a, bp[i * ES], pm, np0);
      row<false>(X, Y, a, bp[(i + 1) * ES], pm, np0);
    }
    if ((L & 1) == 0) {
      row<false>(Y, X, a, bp[(L - 1) * ES], pm, np0);
      merge(r, X, Y);
    } else {
      merge(r, Y, X);
    }
  }

  // ---- double-width arithmetic for lazy reduction (fused.cuh: line_mul).  A product and its
  // Montgomery reduction are separated so that sums of products share ONE reduction:
  //   mulw : T[2L] = a * b           (L^2 products: the CIOS rows without their reduction half; the
  //                                   sliding window retires one finished limb of T per row)
  //   redc : r = T / R mod p         (L^2 + L products: the rows' reduction half on T's low L limbs,
  //                                   then + T's high L limbs); r < T/R + p
  // One multiplication row without reduction; returns the finished lowest limb.
  template <bool FIRST>
  BGN_DEV static uint32_t mrow(uint32_t (&X)[W], uint32_t (&Y)[W], const uint32_t (&a)[L], uint32_t s) {
    if (FIRST) {
      BGN_UNROLL
      for (int k = 0; k < KO; k++) mul_wide(Y[2 * k], Y[2 * k + 1], a[2 * k + 1], s);
      Y[W - 2] = 0;
      Y[W - 1] = 0;
      BGN_UNROLL
      for (int k = 0; k < KE; k++) mul_wide(X[2 * k], X[2 * k + 1], a[2 * k], s);
      if (2 * KE < W) {
        X[W - 2] = 0;
        X[W - 1] = 0;
      }
    } else {
      add_cc(X[0], X[0], Y[1]);
      BGN_UNROLL
      for (int k = 0; k < KO; k++) madc_wide_cc3(Y[2 * k], Y[2 * k + 1], a[2 * k + 1], s, Y[2 * k + 2], Y[2 * k + 3]);
      addc(Y[W - 2], 0, 0);
      Y[W - 1] = 0;
      mad_wide_cc(X[0], X[1], a[0], s);
      BGN_UNROLL
      for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], a[2 * k], s);
      if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
    }
    return X[0];
  }
  // T = a * b, multiplier streamed from memory
  template <int ES = 1>
  BGN_DEV static void mulw(uint32_t (&T)[2 * L], const uint32_t (&a)[L], const uint32_t* bp) {
    uint32_t X[W], Y[W];
#ifdef BGN_HOSTSIM
    {
      double A = BGN_GETB(a), B = BGN_GETB(bp);
      BGN_CHECK(A <= bgnsim::headroom && B <= bgnsim::headroom, "wide product operand too large");
      bgnsim::setw(T, A * B, 0.0);
      bgnsim::nmulw++;
    }
#endif
    T[0] = mrow<true>(X, Y, a, bp[0]);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      T[i] = mrow<false>(Y, X, a, bp[i * ES]);
      T[i + 1] = mrow<false>(X, Y, a, bp[(i + 1) * ES]);
    }
    uint32_t hi[L];
    if ((L & 1) == 0) {
      T[L - 1] = mrow<false>(Y, X, a, bp[(L - 1) * ES]);
      merge(hi, X, Y);
    } else {
      merge(hi, Y, X);
    }
    BGN_UNROLL
    for (int j = 0; j < L; j++) T[L + j] = hi[j];
  }
  // Two independent double-width products with their rows interleaved in program order (ILP 2): the carry
  // chains of one fill the dependency gaps of the other (A/B candidate for line_mul_lazy; fused.cuh)
  template <int ES = 1>
  BGN_DEV static void mulw2(uint32_t (&T0)[2 * L], const uint32_t (&a0)[L], const uint32_t* b0, uint32_t (&T1)[2 * L],
                            const uint32_t (&a1)[L], const uint32_t* b1) {
    uint32_t X0[W], Y0[W], X1[W], Y1[W];
#ifdef BGN_HOSTSIM
    {
      double A = BGN_GETB(a0), B = BGN_GETB(b0), C = BGN_GETB(a1), D = BGN_GETB(b1);
      BGN_CHECK(A <= bgnsim::headroom && B <= bgnsim::headroom && C <= bgnsim::headroom && D <= bgnsim::headroom,
                "wide product operand too large");
      bgnsim::setw(T0, A * B, 0.0);
      bgnsim::setw(T1, C * D, 0.0);
      bgnsim::nmulw += 2;
    }
#endif
    T0[0] = mrow<true>(X0, Y0, a0, b0[0]);
    T1[0] = mrow<true>(X1, Y1, a1, b1[0]);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      T0[i] = mrow<false>(Y0, X0, a0, b0[i * ES]);
      T1[i] = mrow<false>(Y1, X1, a1, b1[i * ES]);
      T0[i + 1] = mrow<false>(X0, Y0, a0, b0[(i + 1) * ES]);
      T1[i + 1] = mrow<false>(X1, Y1, a1, b1[(i + 1) * ES]);
    }
    uint32_t h0[L], h1[L];
    if ((L & 1) == 0) {
      T0[L - 1] = mrow<false>(Y0, X0, a0, b0[(L - 1) * ES]);
      T1[L - 1] = mrow<false>(Y1, X1, a1, b1[(L - 1) * ES]);
      merge(h0, X0, Y0);
      merge(h1, X1, Y1);
    } else {
      merge(h0, Y0, X0);
      merge(h1, Y1, X1);
    }
    BGN_UNROLL
    for (int j = 0; j < L; j++) T0[L + j] = h0[j], T1[L + j] = h1[j];
  }
  // plain integer product T[2L] = a * b of two L-limb numbers held in registers (no field
  // semantics, no range bookkeeping): the building block of the Karatsuba product below
  BGN_DEV static void mulw_raw(uint32_t (&T)[2 * L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    uint32_t X[W], Y[W];
    T[0] = mrow<true>(X, Y, a, b[0]);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      T[i] = mrow<false>(Y, X, a, b[i]);
      T[i + 1] = mrow<false>(X, Y, a, b[i + 1]);
    }
    uint32_t hi[L];
    if ((L & 1) == 0) {
      T[L - 1] = mrow<false>(Y, X, a, b[L - 1]);
      merge(hi, X, Y);
    } else {
      merge(hi, Y, X);
    }
    BGN_UNROLL
    for (int j = 0; j < L; j++) T[L + j] = hi[j];
  }

  // one reduction row on the sliding window (no multiplication half)
  template <bool FIRST>
  BGN_DEV static void rrow(uint32_t (&X)[W], uint32_t (&Y)[W], const uint32_t* __restrict__ pm, uint32_t np0) {
    if (FIRST) {
      uint32_t m = X[0] * np0;
      BGN_UNROLL
      for (int k = 0; k < KO; k++) mul_wide(Y[2 * k], Y[2 * k + 1], pm[2 * k + 1], m);
      Y[W - 2] = 0;
      Y[W - 1] = 0;
      mad_wide_cc(X[0], X[1], pm[0], m);
      BGN_UNROLL
      for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], pm[2 * k], m);
      if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
    } else {
      add_cc(X[0], X[0], Y[1]);
      uint32_t m = X[0] * np0;
      BGN_UNROLL
      for (int k = 0; k < KO; k++) madc_wide_cc3(Y[2 * k], Y[2 * k + 1], pm[2 * k + 1], m, Y[2 * k + 2], Y[2 * k + 3]);
      addc(Y[W - 2], 0, 0);
      Y[W - 1] = 0;
      mad_wide_cc(X[0], X[1], pm[0], m);
      BGN_UNROLL
      for (int k = 1; k < KE; k++) madc_wide_cc(X[2 * k], X[2 * k + 1], pm[2 * k], m);
      if (2 * KE < W) addc(X[2 * KE], X[2 * KE], 0);
    }
  }
  // r = T / R mod p (Montgomery reduction of a double-width value), r < T/R + p
  BGN_DEV static void redc(uint32_t (&r)[L], const uint32_t (&T)[2 * L]) {
    uint32_t X[W], Y[W];
#ifdef BGN_HOSTSIM
    {
      bgnsim::WB w = bgnsim::getw(T);
      BGN_SETB(r, w.bw / bgnsim::headroom + w.kk + 1.0);
      bgnsim::nredc++;
    }
#endif
    const uint32_t* pm = c_fc.p;
    const uint32_t np0 = c_fc.np0;
    BGN_UNROLL
    for (int j = 0; j < L; j++) X[j] = T[j];
    BGN_UNROLL
    for (int j = L; j < W; j++) X[j] = 0;
    rrow<true>(X, Y, pm, np0);
    BGN_UNROLL
    for (int i = 1; i + 1 < L; i += 2) {
      rrow<false>(Y, X, pm, np0);
      rrow<false>(X, Y, pm, np0);
    }
    uint32_t u[L];
    if ((L & 1) == 0) {
      rrow<false>(Y, X, pm, np0);
      merge(u, X, Y);
    } else {
      merge(u, Y, X);
    }
    // + high half of T
    add_cc(r[0], u[0], T[L]);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) addc_cc(r[j], u[j], T[L + j]);
    addc(r[L - 1], u[L - 1], T[2 * L - 1]);
  }
  // S = A + B (double width)
  BGN_DEV static void addw(uint32_t (&S)[2 * L], const uint32_t (&A)[2 * L], const uint32_t (&B)[2 * L]) {
#ifdef BGN_HOSTSIM
    {
      bgnsim::WB a = bgnsim::getw(A), b = bgnsim::getw(B);
      bgnsim::setw(S, a.bw + b.bw, a.kk + b.kk);
    }
#endif
    add_cc(S[0], A[0], B[0]);
    BGN_UNROLL
    for (int j = 1; j < 2 * L - 1; j++) addc_cc(S[j], A[j], B[j]);
    addc(S[2 * L - 1], A[2 * L - 1], B[2 * L - 1]);
  }
  // D = A - B + K p R (double width; K p is added to the high half); requires B <= K p R
  BGN_DEV static void subw_k(uint32_t (&D)[2 * L], const uint32_t (&A)[2 * L], const uint32_t (&B)[2 * L],
                             const uint32_t* kp, int K) {
#ifdef BGN_HOSTSIM
    {
      bgnsim::WB a = bgnsim::getw(A), b = bgnsim::getw(B);
      // B < bw p^2 + kk p R <= (bw / H + kk) p R with H >= 256
      BGN_CHECK(b.bw / bgnsim::headroom + b.kk <= (double)K, "double-width difference may go negative");
      bgnsim::setw(D, a.bw, a.kk + K);
    }
#endif
    sub_cc(D[0], A[0], B[0]);
    BGN_UNROLL
    for (int j = 1; j < 2 * L - 1; j++) subc_cc(D[j], A[j], B[j]);
    subc(D[2 * L - 1], A[2 * L - 1], B[2 * L - 1]);
    add_cc(D[L], D[L], kp[0]);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) addc_cc(D[L + j], D[L + j], kp[j]);
    addc(D[2 * L - 1], D[2 * L - 1], kp[L - 1]);
  }
  // D = A - B (double width) for B <= A by construction (caller's algebra)
  BGN_DEV static void subw(uint32_t (&D)[2 * L], const uint32_t (&A)[2 * L], const uint32_t (&B)[2 * L]) {
#ifdef BGN_HOSTSIM
    {
      bgnsim::WB a = bgnsim::getw(A);
      bgnsim::setw(D, a.bw, a.kk);
    }
#endif
    sub_cc(D[0], A[0], B[0]);
    BGN_UNROLL
    for (int j = 1; j < 2 * L - 1; j++) subc_cc(D[j], A[j], B[j]);
    subc(D[2 * L - 1], A[2 * L - 1], B[2 * L - 1]);
  }

  // ---- relaxed-range helpers (fused.cuh): values are bounded multiples of p far below
  // R = 2^(32L) >= 256 p, so sums need no reduction and differences only a fixed offset K*p
  // that keeps them non-negative; the product absorbs the slack: a < A p, b < B p gives
  // a*b/R mod p < (A*B/256 + 1) p.  tests/hostsim tracks the bounds of every value.
  BGN_DEV static void addn(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
#ifdef BGN_HOSTSIM
    {
      double s = BGN_GETB(a) + BGN_GETB(b);
      BGN_CHECK(s <= bgnsim::headroom, "sum too large");
      BGN_SETB(r, s);
    }
#endif
    add_cc(r[0], a[0], b[0]);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) addc_cc(r[j], a[j], b[j]);
    addc(r[L - 1], a[L - 1], b[L - 1]);
  }
  // r = a - b + K p, kp = the limbs of K p (K in {2, 4, 8, 16}); requires b <= K p
  BGN_DEV static void subk(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L], const uint32_t* kp, int K) {
#ifdef BGN_HOSTSIM
    {
      double A = BGN_GETB(a), B = BGN_GETB(b);
      BGN_CHECK(B <= (double)K, "difference may go negative");
      BGN_CHECK(A + K <= bgnsim::headroom, "difference too large");
      BGN_SETB(r, A + K);
    }
#endif
    uint32_t t[L];
    add_cc(t[0], a[0], kp[0]);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) addc_cc(t[j], a[j], kp[j]);
    addc(t[L - 1], a[L - 1], kp[L - 1]);
    sub_cc(r[0], t[0], b[0]);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) subc_cc(r[j], t[j], b[j]);
    subc(r[L - 1], t[L - 1], b[L - 1]);
  }

  // r = a - {0, 2, 4, 6} p, whichever lies in [0, 2p); requires a < 8p.  Brings a relaxed-range
  // value back to the range every other kernel expects.
  BGN_DEV static void norm2p(uint32_t (&r)[L], const uint32_t (&a)[L]) {
#ifdef BGN_HOSTSIM
    BGN_CHECK(BGN_GETB(a) <= 8.0, "norm2p expects an operand below 8p");
    BGN_SETB(r, 2.0);
#endif
    uint32_t u[L], v[L], bw;
    sub_cc(u[0], a[0], c_fc.p4[0]);
    BGN_UNROLL
    for (int j = 1; j < L; j++) subc_cc(u[j], a[j], c_fc.p4[j]);
    subc(bw, 0, 0);
    BGN_UNROLL
    for (int j = 0; j < L; j++) v[j] = bw ? a[j] : u[j];  // < 4p
    bw = sub_p2(u, v);
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = bw ? v[j] : u[j];
  }

  // r = K p - a; requires a <= K p
  BGN_DEV static void negk(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t* kp, int K) {
#ifdef BGN_HOSTSIM
    BGN_CHECK(BGN_GETB(a) <= (double)K, "negation may go negative");
    BGN_SETB(r, (double)K);
#endif
    sub_cc(r[0], kp[0], a[0]);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) subc_cc(r[j], kp[j], a[j]);
    subc(r[L - 1], kp[L - 1], a[L - 1]);
  }

  // after the final row with roles (X=aligned, Y=offset): result = Y + (X >> 32)
  BGN_DEV static void merge(uint32_t (&r)[L], const uint32_t (&A)[W], const uint32_t (&B)[W]) {
    add_cc(r[0], A[0], B[1]);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) addc_cc(r[j], A[j], B[j + 1]);
    addc(r[L - 1], A[L - 1], B[L]);
  }

  // ---- dedicated squaring (round 2): r = a^2 / R mod p with L (L + 1) / 2 products for the square
  // (every cross product a_i a_j, i < j, once, against the doubled operand 2a; the L squares a_i^2 lead
  // the even chains) plus the L^2 + L of one Montgomery reduction: 459 instead of 595 at L = 17.  The
  // double-width square is accumulated in two arrays as in the product rows -- XE holds the pairs
  // aligned at even limb positions, XO those at odd positions (XO[k] is position k + 1) -- so every
  // (mad.lo.cc, madc.hi.cc) pair still fuses into one IMAD.WIDE.  Row i touches positions 2i and up
  // only; its carry-out lands on a limb no earlier row has used for more than a carry.
  // Requires 2a < R (a < 128 p); r < a^2 / R + p.
  template <int I>
  BGN_DEV static void sq_row(uint32_t (&XE)[2 * L + 2], uint32_t (&XO)[2 * L + 2], const uint32_t (&a)[L],
                             const uint32_t (&a2)[L]) {
    const uint32_t s = a[I];
    // even chain: a_I^2 at position 2I, then a_I * (2 a_j) for j = I + 2, I + 4, ...
    mad_wide_cc(XE[2 * I], XE[2 * I + 1], a[I], s);
    BGN_UNROLL
    for (int j = I + 2; j < L; j += 2) madc_wide_cc(XE[I + j], XE[I + j + 1], a2[j], s);
    {
      constexpr int JL = I + 2 * ((L - 1 - I) / 2);  // last j of the even chain
      addc(XE[I + JL + 2], XE[I + JL + 2], 0);
    }
    // odd chain: a_I * (2 a_j) for j = I + 1, I + 3, ...: position I + j is odd, XO index I + j - 1
    if (I + 1 < L) {
      // the first limb of 2 * (a >> 32 (I + 1)) is a_{I+1} << 1 WITHOUT the bit a_I shifts out (that bit
      // belongs to the doubling of the part at and below limb I, which this row does not multiply)
      mad_wide_cc(XO[2 * I], XO[2 * I + 1], a[I + 1 < L ? I + 1 : 0] << 1, s);
      BGN_UNROLL
      for (int j = I + 3; j < L; j += 2) madc_wide_cc(XO[I + j - 1], XO[I + j], a2[j], s);
      constexpr int JL = I + 1 + 2 * ((L - 2 - I) / 2 < 0 ? 0 : (L - 2 - I) / 2);  // last j of the odd chain
      addc(XO[I + JL + 1], XO[I + JL + 1], 0);
    }
  }
  template <int I, bool END = (I >= L)>
  struct SqRows {
    BGN_DEV static void run(uint32_t (&XE)[2 * L + 2], uint32_t (&XO)[2 * L + 2], const uint32_t (&a)[L],
                            const uint32_t (&a2)[L]) {
      sq_row<I>(XE, XO, a, a2);
      SqRows<I + 1>::run(XE, XO, a, a2);
    }
  };
  template <int I>
  struct SqRows<I, true> {
    BGN_DEV static void run(uint32_t (&)[2 * L + 2], uint32_t (&)[2 * L + 2], const uint32_t (&)[L], const uint32_t (&)[L]) {}
  };
  // T = a^2 (double width)
  BGN_DEV static void sqrw(uint32_t (&T)[2 * L], const uint32_t (&a)[L]) {
    uint32_t XE[2 * L + 2], XO[2 * L + 2], a2[L];
#ifdef BGN_HOSTSIM
    {
      double A = BGN_GETB(a);
      BGN_CHECK(2.0 * A <= bgnsim::headroom, "squaring operand too large");
      bgnsim::setw(T, A * A, 0.0);
      bgnsim::nsqrw++;
    }
#endif
    BGN_UNROLL
    for (int j = 0; j < 2 * L + 2; j++) XE[j] = 0, XO[j] = 0;
    BGN_UNROLL
    for (int j = L - 1; j > 0; j--) a2[j] = (a[j] << 1) | (a[j - 1] >> 31);
    a2[0] = a[0] << 1;
    SqRows<0>::run(XE, XO, a, a2);
    // T = XE + (XO << 32)
    T[0] = XE[0];
    add_cc(T[1], XE[1], XO[0]);
    BGN_UNROLL
    for (int j = 2; j < 2 * L - 1; j++) addc_cc(T[j], XE[j], XO[j - 1]);
    addc(T[2 * L - 1], XE[2 * L - 1], XO[2 * L - 2]);
  }
  BGN_DEV static void sqr(uint32_t (&r)[L], const uint32_t (&a)[L]) {
#ifdef BGN_NO_DEDICATED_SQR
    mul(r, a, a);
#else
    uint32_t T[2 * L];
    sqrw(T, a);
    redc(r, T);
#endif
  }

  // r = a + b, kept in [0,2p)
  BGN_DEV static void add(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    uint32_t t[L], u[L];
#ifdef BGN_HOSTSIM
    BGN_CHECK(BGN_GETB(a) + BGN_GETB(b) <= 4.0, "reduced add expects operands below 2p");
    BGN_SETB(r, 2.0);
#endif
    add_cc(t[0], a[0], b[0]);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) addc_cc(t[j], a[j], b[j]);
    addc(t[L - 1], a[L - 1], b[L - 1]);
    uint32_t bw = sub_p2(u, t);
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = bw ? t[j] : u[j];
  }

  // u = t - 2p; returns all-ones when the subtraction borrowed (t < 2p)
  BGN_DEV static uint32_t sub_p2(uint32_t (&u)[L], const uint32_t (&t)[L]) {
    uint32_t bw;
    sub_cc(u[0], t[0], c_fc.p2[0]);
    BGN_UNROLL
    for (int j = 1; j < L; j++) subc_cc(u[j], t[j], c_fc.p2[j]);
    subc(bw, 0, 0);  // 0 - 0 - CF: all-ones mask on borrow
    return bw;
  }

  // r = a - b, kept in [0,2p)
  BGN_DEV static void sub(uint32_t (&r)[L], const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    uint32_t t[L], mask;
#ifdef BGN_HOSTSIM
    BGN_CHECK(BGN_GETB(a) <= 2.0 && BGN_GETB(b) <= 2.0, "reduced sub expects operands below 2p");
    BGN_SETB(r, 2.0);
#endif
    sub_cc(t[0], a[0], b[0]);
    BGN_UNROLL
    for (int j = 1; j < L; j++) subc_cc(t[j], a[j], b[j]);
    subc(mask, 0, 0);  // borrow -> add 2p
    add_cc(r[0], t[0], c_fc.p2[0] & mask);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) addc_cc(r[j], t[j], c_fc.p2[j] & mask);
    addc(r[L - 1], t[L - 1], c_fc.p2[L - 1] & mask);
  }

  // canonical representative in [0,p) of a value in [0,2p]
  BGN_DEV static void canon(uint32_t (&r)[L], const uint32_t (&a)[L]) {
    uint32_t t[L], u[L], bw;
#ifdef BGN_HOSTSIM
    BGN_CHECK(BGN_GETB(a) <= 2.0, "canon expects an operand of at most 2p");
    BGN_SETB(r, 1.0);
#endif
    sub_cc(t[0], a[0], c_fc.p[0]);
    BGN_UNROLL
    for (int j = 1; j < L; j++) subc_cc(t[j], a[j], c_fc.p[j]);
    subc(bw, 0, 0);
    BGN_UNROLL
    for (int j = 0; j < L; j++) u[j] = bw ? a[j] : t[j];
    // a may equal 2p exactly only for inputs outside the invariant; one more pass is cheap and exact
    sub_cc(t[0], u[0], c_fc.p[0]);
    BGN_UNROLL
    for (int j = 1; j < L; j++) subc_cc(t[j], u[j], c_fc.p[j]);
    subc(bw, 0, 0);
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = bw ? u[j] : t[j];
  }

  // r = a / 2 mod p for a < 2p: an odd a first gains p (3p < R), then one right shift.  r < 1.5p.
  BGN_DEV static void halve(uint32_t (&r)[L], const uint32_t (&a)[L]) {
#ifdef BGN_HOSTSIM
    BGN_CHECK(BGN_GETB(a) <= 2.0, "halve expects an operand below 2p");
    BGN_SETB(r, 1.5);
#endif
    uint32_t t[L], mask = 0u - (a[0] & 1u);
    add_cc(t[0], a[0], c_fc.p[0] & mask);
    BGN_UNROLL
    for (int j = 1; j < L - 1; j++) addc_cc(t[j], a[j], c_fc.p[j] & mask);
    addc(t[L - 1], a[L - 1], c_fc.p[L - 1] & mask);
    BGN_UNROLL
    for (int j = 0; j < L - 1; j++) r[j] = (t[j] >> 1) | (t[j + 1] << 31);
    r[L - 1] = t[L - 1] >> 1;
  }

  // r = x^-1 mod p as plain integers, x canonical in [0, p); 0 -> 0.  Constant-time binary extended
  // GCD on the ALU pipe (adds, selects, shifts -- no multiplications): with a = u x, b = v x (mod p),
  // b odd, every iteration halves a (after a <- |a - b| when a is odd), so the product a b at least
  // halves and 2 bits(p) iterations reach a = 0, b = 1, v = x^-1.  About 12 L plain-ALU
  // instructions per iteration against the ~1.2 bits(p) Montgomery products of the Fermat inversion
  // (F<L>::inv): a quarter of the issue slots, none of them on the multiply pipe, and a sixth of the
  // dependent-chain latency, which is what bounds the batched inversion of k_normalize.
  BGN_DEV static void inv_bgcd(uint32_t (&r)[L], const uint32_t (&x)[L]) {
#ifdef BGN_HOSTSIM
    BGN_CHECK(BGN_GETB(x) <= 1.0, "inv_bgcd expects a canonical operand");
    BGN_SETB(r, 1.0);
#endif
    uint32_t a[L], b[L], u[L], v[L];
    BGN_UNROLL
    for (int j = 0; j < L; j++) {
      a[j] = x[j];
      b[j] = c_fc.p[j];
      u[j] = 0;
      v[j] = 0;
    }
    u[0] = 1;
    int top = 32 * L - 1;
    while (top > 0 && !((c_fc.p[top >> 5] >> (top & 31)) & 1)) top--;
    const int iters = 2 * (top + 1);
    BGN_UNROLL1
    for (int it = 0; it < iters; it++) {
      const uint32_t odd = 0u - (a[0] & 1u);
      uint32_t d[L], nd[L], lt;
      sub_cc(d[0], a[0], b[0]);
      BGN_UNROLL
      for (int j = 1; j < L; j++) subc_cc(d[j], a[j], b[j]);
      subc(lt, 0, 0);  // all-ones when a < b
      sub_cc(nd[0], 0, d[0]);
      BGN_UNROLL
      for (int j = 1; j < L - 1; j++) subc_cc(nd[j], 0, d[j]);
      subc(nd[L - 1], 0, d[L - 1]);  // b - a
      const uint32_t sw = odd & lt;
      BGN_UNROLL
      for (int j = 0; j < L; j++) {
        const uint32_t diff = lt ? nd[j] : d[j];
        const uint32_t aj = a[j];
        a[j] = odd ? diff : aj;
        b[j] = sw ? aj : b[j];
      }
      BGN_UNROLL
      for (int j = 0; j < L - 1; j++) a[j] = (a[j] >> 1) | (a[j + 1] << 31);
      a[L - 1] >>= 1;
      // the same steps on the cofactors, mod p
      BGN_UNROLL
      for (int j = 0; j < L; j++) {
        const uint32_t uj = u[j];
        u[j] = sw ? v[j] : uj;
        v[j] = sw ? uj : v[j];
      }
      uint32_t w[L], m;
      sub_cc(w[0], u[0], v[0]);
      BGN_UNROLL
      for (int j = 1; j < L; j++) subc_cc(w[j], u[j], v[j]);
      subc(m, 0, 0);  // borrow: add p back
      add_cc(w[0], w[0], c_fc.p[0] & m);
      BGN_UNROLL
      for (int j = 1; j < L - 1; j++) addc_cc(w[j], w[j], c_fc.p[j] & m);
      addc(w[L - 1], w[L - 1], c_fc.p[L - 1] & m);
      BGN_UNROLL
      for (int j = 0; j < L; j++) u[j] = odd ? w[j] : u[j];
      const uint32_t h = 0u - (u[0] & 1u);  // u / 2 mod p: an odd u first gains p (u + p < 2p < R)
      add_cc(u[0], u[0], c_fc.p[0] & h);
      BGN_UNROLL
      for (int j = 1; j < L - 1; j++) addc_cc(u[j], u[j], c_fc.p[j] & h);
      addc(u[L - 1], u[L - 1], c_fc.p[L - 1] & h);
      BGN_UNROLL
      for (int j = 0; j < L - 1; j++) u[j] = (u[j] >> 1) | (u[j + 1] << 31);
      u[L - 1] >>= 1;
    }
    BGN_UNROLL
    for (int j = 0; j < L; j++) r[j] = v[j];
  }

  // ---- r = x^-1 mod p by Bernstein-Yang "safegcd" division steps, 30 at a time -------------------
  // (delta, f, g) <- (1 - delta, g, (g - f)/2) if delta > 0 and g odd, else (1 + delta, f, (g + (g mod 2) f)/2),
  // started at (1, p, x): the decisions of 30 consecutive steps depend on delta and the low 30 bits of
  // f and g only, so they are taken on one machine word and give a 2x2 integer matrix t with
  // (f, g) <- t (f, g) / 2^30; the same matrix drives the cofactors (d, e) <- t (d, e) / 2^30 mod p
  // with d x = f, e x = g (mod p).  After floor((49 bits + 57) / 17) steps (Bernstein-Yang, Thm 11.2)
  // g = 0, f = +-1 and d = +-x^-1.  Numbers are 30-bit signed limbs in 32-bit words; the products are
  // 32 x 32 -> 64 multiply-adds: about 250 of them and ~750 plain instructions per round, 50 rounds at
  // 519 bits -- a quarter of inv_bgcd's instruction count.  The CALLER verifies the result with one
  // Montgomery product and falls back to inv_bgcd (F<L>::inv_gcd), so this routine's correctness is
  // checked on every use, not assumed.  x canonical in [0, p).  Returns false if the ladder did not
  // end in g = 0, |f| = 1.
  static constexpr int N30 = (32 * L + 29) / 30 + 1;
  BGN_DEV static void to30(int32_t (&v)[N30], const uint32_t (&x)[L]) {
    BGN_UNROLL
    for (int i = 0; i < N30; i++) {
      const int off = 30 * i, w = off >> 5, sh = off & 31;
      uint64_t acc = 0;
      if (w < L) acc = x[w];
      if (w + 1 < L) acc |= (uint64_t)x[w + 1] << 32;
      v[i] = (int32_t)((acc >> sh) & 0x3fffffffu);
    }
  }
  // v: limbs in [0, 2^30), value below 2^(32 L)
  BGN_DEV static void from30(uint32_t (&x)[L], const int32_t (&v)[N30]) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) {
      const int off = 32 * j, k = off / 30, sh = off % 30;
      uint64_t acc = (uint64_t)(uint32_t)v[k] >> sh;
      if (k + 1 < N30) acc |= (uint64_t)(uint32_t)v[k + 1] << (30 - sh);
      if (k + 2 < N30) acc |= (uint64_t)(uint32_t)v[k + 2] << (60 - sh);
      x[j] = (uint32_t)acc;
    }
  }
  BGN_DEV static bool inv_safegcd(uint32_t (&r)[L], const uint32_t (&x)[L]) {
#ifdef BGN_HOSTSIM
    BGN_CHECK(BGN_GETB(x) <= 1.0, "inv_safegcd expects a canonical operand");
    BGN_SETB(r, 1.0);
#endif
    const int32_t M30 = 0x3fffffff;
    int32_t d[N30], e[N30], f[N30], g[N30], pm[N30];
    {
      uint32_t pc[L];
      BGN_UNROLL
      for (int j = 0; j < L; j++) pc[j] = c_fc.p[j];
      to30(pm, pc);
    }
    to30(g, x);
    BGN_UNROLL
    for (int i = 0; i < N30; i++) {
      f[i] = pm[i];
      d[i] = 0;
      e[i] = 0;
    }
    e[0] = 1;
    const uint32_t pinv30 = (0u - c_fc.np0) & (uint32_t)M30;  // p^-1 mod 2^30 (np0 = -p^-1 mod 2^32)
    int top = 32 * L - 1;
    while (top > 0 && !((c_fc.p[top >> 5] >> (top & 31)) & 1)) top--;
    const int bits = top + 1;
    const int steps = bits >= 46 ? (49 * bits + 57) / 17 : (49 * bits + 80) / 17;
    const int rounds = (steps + 29) / 30;
    int32_t delta = 1;
    BGN_UNROLL1
    for (int rd = 0; rd < rounds; rd++) {
      // 30 division steps on the low words
      uint32_t fl = (uint32_t)f[0] | ((uint32_t)f[1] << 30), gl = (uint32_t)g[0] | ((uint32_t)g[1] << 30);
      int32_t u = 1, v = 0, q = 0, w = 1;
      BGN_UNROLL1
      for (int i = 0; i < 30; i++) {
        const uint32_t odd = 0u - (gl & 1u);
        const uint32_t pos = (uint32_t)((int32_t)(0 - delta) >> 31);  // all ones iff delta > 0
        const uint32_t sw = odd & pos;
        const uint32_t fs = (fl ^ sw) - sw;                 // -f when swapping
        const int32_t us = (u ^ (int32_t)sw) - (int32_t)sw;
        const int32_t vs = (v ^ (int32_t)sw) - (int32_t)sw;
        const uint32_t g2 = gl + (fs & odd);
        const int32_t q2 = q + (us & (int32_t)odd), w2 = w + (vs & (int32_t)odd);
        fl = sw ? gl : fl;
        u = 2 * (sw ? q : u);
        v = 2 * (sw ? w : v);
        gl = g2 >> 1;
        q = q2;
        w = w2;
        delta = sw ? 1 - delta : 1 + delta;
      }
      // (d, e) <- t (d, e) / 2^30 mod p, kept in (-2p, p)
      {
        const int32_t sd = d[N30 - 1] >> 31, se = e[N30 - 1] >> 31;
        int32_t md = (u & sd) + (v & se), me = (q & sd) + (w & se);
        int64_t cd = (int64_t)u * d[0] + (int64_t)v * e[0];
        int64_t ce = (int64_t)q * d[0] + (int64_t)w * e[0];
        md -= (int32_t)((pinv30 * (uint32_t)cd + (uint32_t)md) & (uint32_t)M30);
        me -= (int32_t)((pinv30 * (uint32_t)ce + (uint32_t)me) & (uint32_t)M30);
        cd += (int64_t)pm[0] * md;
        ce += (int64_t)pm[0] * me;
        cd >>= 30;
        ce >>= 30;
        BGN_UNROLL
        for (int i = 1; i < N30; i++) {
          cd += (int64_t)u * d[i] + (int64_t)v * e[i] + (int64_t)pm[i] * md;
          ce += (int64_t)q * d[i] + (int64_t)w * e[i] + (int64_t)pm[i] * me;
          d[i - 1] = (int32_t)cd & M30;
          cd >>= 30;
          e[i - 1] = (int32_t)ce & M30;
          ce >>= 30;
        }
        d[N30 - 1] = (int32_t)cd;
        e[N30 - 1] = (int32_t)ce;
      }
      // (f, g) <- t (f, g) / 2^30 (exact)
      {
        int64_t cf = (int64_t)u * f[0] + (int64_t)v * g[0];
        int64_t cg = (int64_t)q * f[0] + (int64_t)w * g[0];
        cf >>= 30;
        cg >>= 30;
        BGN_UNROLL
        for (int i = 1; i < N30; i++) {
          cf += (int64_t)u * f[i] + (int64_t)v * g[i];
          cg += (int64_t)q * f[i] + (int64_t)w * g[i];
          f[i - 1] = (int32_t)cf & M30;
          cf >>= 30;
          g[i - 1] = (int32_t)cg & M30;
          cg >>= 30;
        }
        f[N30 - 1] = (int32_t)cf;
        g[N30 - 1] = (int32_t)cg;
      }
    }
    // g == 0 and f == +-1 ?
    int32_t gz = 0, fp1 = f[0] ^ 1, fm1 = f[0] ^ M30;
    BGN_UNROLL
    for (int i = 0; i < N30; i++) gz |= g[i];
    BGN_UNROLL
    for (int i = 1; i < N30 - 1; i++) {
      fp1 |= f[i];
      fm1 |= f[i] ^ M30;
    }
    fp1 |= f[N30 - 1];
    fm1 |= f[N30 - 1] ^ -1;
    const bool ok = gz == 0 && (fp1 == 0 || fm1 == 0);
    // d in (-2p, p): add p if negative, negate if f = -1, add p if negative -> [0, p)
    auto add_p_if_neg = [&]() {
      const int32_t m = d[N30 - 1] >> 31;
      int32_t c = 0;
      BGN_UNROLL
      for (int i = 0; i < N30 - 1; i++) {
        c += d[i] + (pm[i] & m);
        d[i] = c & M30;
        c >>= 30;
      }
      d[N30 - 1] += (pm[N30 - 1] & m) + c;
    };
    add_p_if_neg();
    {
      const int32_t m = f[N30 - 1] >> 31;  // f = -1
      int32_t c = 0;
      BGN_UNROLL
      for (int i = 0; i < N30 - 1; i++) {
        c += (d[i] ^ m) - m;
        d[i] = c & M30;
        c >>= 30;
      }
      d[N30 - 1] = ((d[N30 - 1] ^ m) - m) + c;
    }
    add_p_if_neg();
    from30(r, d);
    return ok;
  }

  BGN_DEV static bool is_zero_raw(const uint32_t (&a)[L]) {
    uint32_t o = 0;
    BGN_UNROLL
    for (int j = 0; j < L; j++) o |= a[j];
    return o == 0;
  }
  BGN_DEV static bool eq_raw(const uint32_t (&a)[L], const uint32_t (&b)[L]) {
    uint32_t o = 0;
    BGN_UNROLL
    for (int j = 0; j < L; j++) o |= a[j] ^ b[j];
    return o == 0;
  }
};

// ---------------------------------------------------------------------------
// One level of Karatsuba on the double-width product: T = a * b with the operands split into a
// low part of N0 = L/2 limbs and a high part of N1 = L - N0 limbs,
//   a b = z0 + (zm - z0 - z2) B^N0 + z2 B^(2 N0),  z0 = a0 b0, z2 = a1 b1, zm = (a0 + a1)(b0 + b1):
// N0^2 + 2 N1^2 products instead of L^2 (226 instead of 289 at L = 17) for ~5L additions that sit
// in the issue slots the multiplier leaves free.  The sums a0 + a1 and b0 + b1 need no carry
// limb: operands are below R / 4, so their top limb is far from full (checked by the tracker).
// ---------------------------------------------------------------------------
template <int L>
struct Kara {
  static constexpr int N0 = L / 2, N1 = L - N0;
  typedef Fp<L> P;

  template <int ES = 1>
  BGN_DEV static void mulw(uint32_t (&T)[2 * L], const uint32_t (&a)[L], const uint32_t* bp) {
#ifdef BGN_HOSTSIM
    {
      double A = BGN_GETB(a), B = BGN_GETB(bp);
      // a < R/4 keeps the top limb of a below 2^30: a0 + a1 fits N1 limbs
      BGN_CHECK(4.0 * A <= bgnsim::headroom && 4.0 * B <= bgnsim::headroom, "Karatsuba operand too large");
      bgnsim::setw(T, A * B, 0.0);
      bgnsim::nmulk++;
    }
#endif
    uint32_t a0[N0], a1[N1], b0[N0], b1[N1], sa[N1], sb[N1];
    BGN_UNROLL
    for (int j = 0; j < N0; j++) a0[j] = a[j], b0[j] = bp[j * ES];
    BGN_UNROLL
    for (int j = 0; j < N1; j++) a1[j] = a[N0 + j], b1[j] = bp[(N0 + j) * ES];
    // sa = a1 + a0, sb = b1 + b0 (N1 limbs)
    add_cc(sa[0], a1[0], a0[0]);
    BGN_UNROLL
    for (int j = 1; j < N1; j++) {
      if (j < N1 - 1)
        addc_cc(sa[j], a1[j], j < N0 ? a0[j] : 0u);
      else
        addc(sa[j], a1[j], j < N0 ? a0[j] : 0u);
    }
    add_cc(sb[0], b1[0], b0[0]);
    BGN_UNROLL
    for (int j = 1; j < N1; j++) {
      if (j < N1 - 1)
        addc_cc(sb[j], b1[j], j < N0 ? b0[j] : 0u);
      else
        addc(sb[j], b1[j], j < N0 ? b0[j] : 0u);
    }
    uint32_t z0[2 * N0], z2[2 * N1], zm[2 * N1];
    Fp<N0>::mulw_raw(z0, a0, b0);
    Fp<N1>::mulw_raw(z2, a1, b1);
    Fp<N1>::mulw_raw(zm, sa, sb);
    // zm <- zm - z0 - z2  (= a0 b1 + a1 b0 >= 0)
    sub_cc(zm[0], zm[0], z0[0]);
    BGN_UNROLL
    for (int j = 1; j < 2 * N1; j++) {
      if (j < 2 * N1 - 1)
        subc_cc(zm[j], zm[j], j < 2 * N0 ? z0[j] : 0u);
      else
        subc(zm[j], zm[j], j < 2 * N0 ? z0[j] : 0u);
    }
    sub_cc(zm[0], zm[0], z2[0]);
    BGN_UNROLL
    for (int j = 1; j < 2 * N1; j++) {
      if (j < 2 * N1 - 1)
        subc_cc(zm[j], zm[j], z2[j]);
      else
        subc(zm[j], zm[j], z2[j]);
    }
    // T = z0 | z2 (concatenation), then += zm at limb N0
    BGN_UNROLL
    for (int j = 0; j < 2 * N0; j++) T[j] = z0[j];
    BGN_UNROLL
    for (int j = 0; j < 2 * N1; j++) T[2 * N0 + j] = z2[j];
    add_cc(T[N0], T[N0], zm[0]);
    BGN_UNROLL
    for (int j = 1; j < 2 * L - N0; j++) {
      if (j < 2 * L - N0 - 1)
        addc_cc(T[N0 + j], T[N0 + j], j < 2 * N1 ? zm[j] : 0u);
      else
        addc(T[N0 + j], T[N0 + j], j < 2 * N1 ? zm[j] : 0u);
    }
  }
};
