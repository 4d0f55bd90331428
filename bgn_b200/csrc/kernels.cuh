// kernels.cuh -- the sm_100a kernels of the batched BGN engine (templated on
// the limb count L).  Device data layout: arrays of elements, array[N][L] of
// u32 in Montgomery form ("lazy" range [0,2p)); G1 arrays carry an extra
// byte-per-element infinity flag.  See DESIGN.md for the per-kernel roofline.
//
// Every kernel is a thin __global__ wrapper around a `*_body(args, index)`
// function so the CPU-side unit tests (tests/hostsim, BGN_HOSTSIM) can run the
// very same per-thread program; the wrappers only exist in the CUDA build.
#pragma once
#include "pairing.cuh"
#include "lucas.cuh"
#include "pairlane.cuh"
#include "pairwarp.cuh"
#include "teamsplit.cuh"

#ifdef BGN_HOSTSIM
static inline uint32_t atomicCAS(uint32_t* p, uint32_t cmp, uint32_t val) {
  uint32_t old = *p;
  if (old == cmp) *p = val;
  return old;
}
#else
#define BGN_GID(T) ((T)blockIdx.x * blockDim.x + threadIdx.x)
#endif

// ------------------------------------------------------------------ helpers
template <int L>
BGN_DEV void be_bytes_to_limbs(uint32_t (&x)[L], const uint8_t* b, int B) {
#ifdef BGN_HOSTSIM
  {  // raw integer below 2^(8B): bound in multiples of p
    long double pd = 0;
    for (int j = L - 1; j >= 0; j--) pd = pd * 4294967296.0L + c_fc.p[j];
    long double v = 1;
    for (int i = 0; i < B; i++) v *= 256.0L;
    BGN_SETB(x, (double)(v / pd));
  }
#endif
  BGN_UNROLL
  for (int j = 0; j < L; j++) {
    uint32_t w = 0;
    BGN_UNROLL
    for (int k = 0; k < 4; k++) {
      int pos = B - 1 - (4 * j + k);
      if (pos >= 0) w |= (uint32_t)b[pos] << (8 * k);
    }
    x[j] = w;
  }
}
template <int L>
BGN_DEV void limbs_to_be_bytes(uint8_t* b, int B, const uint32_t (&x)[L]) {
  BGN_UNROLL
  for (int j = 0; j < L; j++) {
    BGN_UNROLL
    for (int k = 0; k < 4; k++) {
      int pos = B - 1 - (4 * j + k);
      if (pos >= 0) b[pos] = (uint8_t)(x[j] >> (8 * k));
    }
  }
}

// ---------------------------------------------------- (de)serialisation
// PBC element_from_bytes on G1: x||y big-endian, reduced mod p; all-zero bytes
// or a pair not on the curve becomes O (SURVEY.md 8(a) identity note).
template <int L>
BGN_DEV void g1_from_bytes_body(const uint8_t* in, int B, size_t count, uint32_t* x, uint32_t* y, uint8_t* inf,
                                size_t N, size_t e) {
  typedef F<L> FF;
  if (e >= count) return;
  Loc<L> a, b, t0, t1;
  be_bytes_to_limbs<L>(a.w, in + e * 2 * B, B);
  be_bytes_to_limbs<L>(b.w, in + e * 2 * B + B, B);
  uint32_t o = 0;
  BGN_UNROLL
  for (int j = 0; j < L; j++) o |= a.w[j] | b.w[j];
  FF::to_mont(a.v(), a.v());
  FF::to_mont(b.v(), b.v());
  bool isinf = (o == 0);
  if (!isinf) isinf = !G<L>::on_curve(a.v(), b.v(), t0.v(), t1.v());
  FF::copy(x + (size_t)(e) * L, a.v());
  FF::copy(y + (size_t)(e) * L, b.v());
  inf[e] = isinf ? 1 : 0;
}

template <int L>
BGN_DEV void g1_to_bytes_body(const uint32_t* x, const uint32_t* y, const uint8_t* inf, size_t N, size_t count,
                              uint8_t* out, int B, size_t e) {
  typedef F<L> FF;
  if (e >= count) return;
  Loc<L> a, b;
  if (inf[e]) {
    FF::set_zero(a.v());
    FF::set_zero(b.v());
  } else {
    FF::copy(a.v(), (x + (size_t)(e) * L));
    FF::copy(b.v(), (y + (size_t)(e) * L));
    FF::from_mont(a.v(), a.v());
    FF::from_mont(b.v(), b.v());
  }
  limbs_to_be_bytes<L>(out + e * 2 * B, B, a.w);
  limbs_to_be_bytes<L>(out + e * 2 * B + B, B, b.w);
}

template <int L>
BGN_DEV void fp2_from_bytes_body(const uint8_t* in, int B, size_t count, uint32_t* re, uint32_t* im, size_t N,
                                 size_t e) {
  typedef F<L> FF;
  if (e >= count) return;
  Loc<L> a, b;
  be_bytes_to_limbs<L>(a.w, in + e * 2 * B, B);
  be_bytes_to_limbs<L>(b.w, in + e * 2 * B + B, B);
  FF::to_mont(a.v(), a.v());
  FF::to_mont(b.v(), b.v());
  FF::copy(re + (size_t)(e) * L, a.v());
  FF::copy(im + (size_t)(e) * L, b.v());
}

// `count` output elements.  grp == 0: element e of the array.  grp > 0: the output is grouped as
// polynomials of grp + pad slots, the first grp taken from consecutive array elements and the pad
// slots written as the GT identity (MakePolyL2's unused top slot, poly.go:130-137, 159-163).
template <int L>
BGN_DEV void fp2_to_bytes_body(const uint32_t* re, const uint32_t* im, size_t N, size_t count, uint8_t* out, int B,
                               int grp, int pad, size_t e) {
  typedef F<L> FF;
  if (e >= count) return;
  Loc<L> a, b;
  size_t src = e;
  if (grp > 0) {
    size_t u = e / (size_t)(grp + pad), s = e % (size_t)(grp + pad);
    if (s >= (size_t)grp) {
      uint8_t* o = out + e * 2 * B;
      for (int i = 0; i < 2 * B; i++) o[i] = (i == B - 1) ? 1 : 0;
      return;
    }
    src = u * grp + s;
  }
  FF::copy(a.v(), (re + (size_t)(src) * L));
  FF::copy(b.v(), (im + (size_t)(src) * L));
  FF::from_mont(a.v(), a.v());
  FF::from_mont(b.v(), b.v());
  limbs_to_be_bytes<L>(out + e * 2 * B, B, a.w);
  limbs_to_be_bytes<L>(out + e * 2 * B + B, B, b.w);
}

// ------------------------------------------------------------ G1 kernels
// Fixed-base window tables: entry (win, d) = d * 2^(8 win) * Base, affine
// Montgomery, AoS: tab[(win*255 + d-1) * 2L + {0..L-1: x, L..2L-1: y}].
// C = x*P + r*Q  (EncryptWithRandomness, bgn.go:340-353), x < 0: C = -(|x|*P + r*Q); Jacobian out.
template <int L>
BGN_DEV uint32_t enc_digit(const uint8_t* r, int rbytes, int win, int wbits) {
  const int bit = win * wbits;            // offset of the digit from the least significant bit of r
  const int lo = rbytes - 1 - (bit >> 3);
  uint32_t v = r[lo];                     // four bytes cover 7 + 24 bits
  for (int k = 1; k < 4; k++)
    if (lo - k >= 0) v |= (uint32_t)r[lo - k] << (8 * k);
  return (v >> (bit & 7)) & ((1u << wbits) - 1u);
}
template <int L>
BGN_DEV void encrypt_body(const EncArgs& a, size_t e) {
  typedef F<L> FF;
  if (e >= a.count) return;
  Loc<L> X, Y, Z, t0, t1, t2, t3;
  int64_t xs = a.x ? a.x[e] : 0;
  bool neg = xs < 0;
  uint64_t xm = neg ? (uint64_t)(-(xs + 1)) + 1u : (uint64_t)xs;
  // fixed-base windows of wbitsQ bits (8 .. 24, any width): one table point added per non-zero window digit
  // of r, plus one per non-zero byte of |x|.  The wide tables (2^wbitsQ - 1 points per window: 285 MB at 16
  // bits, 3.7 GB at 20, 50 GB at 24 for 512-bit keys and Weierstrass points) live in HBM and each lookup is
  // one read of an entry at a random address.
  const int wbits = a.wbitsQ;
  const size_t ents = ((size_t)1 << wbits) - 1;
  const int nw = a.r_be ? (8 * a.rbytes + wbits - 1) / wbits : 0;
  const uint8_t* r = a.r_be ? a.r_be + e * a.rbytes : nullptr;
  if (a.edw) {
    // twisted Edwards form (curve.cuh: Ed): 8 products per table point, no special cases; the sum is
    // converted to Jacobian Weierstrass coordinates (6 products) before the starting point, if any, is added
    Loc<L> T;
    bool started = false;
    for (int win = 0; win < nw; win++) {
      const uint32_t d = enc_digit<L>(r, a.rbytes, win, wbits);
      if (!d) continue;
      const uint32_t* ent = a.tabQ + ((size_t)win * ents + (d - 1)) * 3 * L;
      if (started)
        Ed<L>::madd(X.v(), Y.v(), Z.v(), T.v(), ent, t0.v(), t1.v(), t2.v(), t3.v());
      else
        Ed<L>::set(X.v(), Y.v(), Z.v(), T.v(), ent);
      started = true;
    }
    for (int win = 0; win < 8; win++) {
      const uint32_t d = (uint32_t)(xm >> (8 * win)) & 255u;
      if (!d) continue;
      const uint32_t* ent = a.tabP + ((size_t)win * 255 + (d - 1)) * 3 * L;
      if (started)
        Ed<L>::madd(X.v(), Y.v(), Z.v(), T.v(), ent, t0.v(), t1.v(), t2.v(), t3.v());
      else
        Ed<L>::set(X.v(), Y.v(), Z.v(), T.v(), ent);
      started = true;
    }
    if (started) {
      Ed<L>::to_jac(X.v(), Y.v(), Z.v(), t0.v(), t1.v(), t2.v());
    } else {
      FF::set_zero(X.v());
      FF::set_zero(Y.v());
      FF::set_zero(Z.v());
    }
    if (a.bx && !a.binf[e])  // re-randomisation: the ciphertext + r*Q
      G<L>::madd(X.v(), Y.v(), Z.v(), a.bx + e * L, a.by + e * L, false, t0.v(), t1.v(), t2.v(), t3.v());
  } else {
    FF::set_zero(X.v());
    FF::set_zero(Y.v());
    FF::set_zero(Z.v());
    if (a.bx && !a.binf[e]) {  // re-randomisation: start from the ciphertext instead of O
      FF::copy(X.v(), a.bx + e * L);
      FF::copy(Y.v(), a.by + e * L);
      FF::set_one(Z.v());
    }
    // Weierstrass tables: one complete mixed Jacobian addition per table point
    for (int win = 0; win < nw; win++) {
      const uint32_t d = enc_digit<L>(r, a.rbytes, win, wbits);
      if (d) {
        const uint32_t* ent = a.tabQ + ((size_t)win * ents + (d - 1)) * 2 * L;
        G<L>::madd(X.v(), Y.v(), Z.v(), ent, ent + L, false, t0.v(), t1.v(), t2.v(), t3.v());
      }
    }
    for (int win = 0; win < 8; win++) {
      uint32_t d = (uint32_t)(xm >> (8 * win)) & 255u;
      if (d) {
        const uint32_t* ent = a.tabP + ((size_t)win * 255 + (d - 1)) * 2 * L;
        G<L>::madd(X.v(), Y.v(), Z.v(), ent, ent + L, false, t0.v(), t1.v(), t2.v(), t3.v());
      }
    }
  }
  // negative plaintext: -(|x|*P + r*Q), the Sub(encryptZero(), Encrypt(|c|)) of poly.go:17-21
  if (neg) FF::neg(Y.v(), Y.v());
  FF::copy(a.X + (size_t)(e) * L, X.v());
  FF::copy(a.Y + (size_t)(e) * L, Y.v());
  FF::copy(a.Z + (size_t)(e) * L, Z.v());
}

// Jacobian -> affine with one inversion per thread (Montgomery's trick over
// the elements g, g+G, g+2G, ... of thread g).  Output element e goes to
// ox/oy + e*o_estride, so the same kernel fills x[]/y[] arrays and x||y tables.
template <int L>
BGN_DEV void normalize_body(const NormArgs& a, size_t g) {
  typedef F<L> FF;
  if (g >= (size_t)a.G || g >= a.count) return;
  Loc<L> acc, z, zi, zz, t;
  FF::set_one(acc.v());
  for (size_t e = g; e < a.count; e += a.G) {
    FF::copy(z.v(), (a.Z + (size_t)(e) * L));
    if (FF::is_zero(z.v())) continue;
    FF::copy(a.scratch + (size_t)(e) * L, acc.v());
    FF::mul(acc.v(), acc.v(), z.v());
  }
#ifdef BGN_NORMALIZE_FERMAT
  FF::inv(acc.v(), acc.v(), t.v());
#else
  FF::inv_gcd_fast(acc.v(), acc.v());
#endif
  size_t last = ((a.count - 1 - g) / a.G) * a.G + g;  // largest e = g (mod G) below count
  for (size_t e = last;; e -= a.G) {
    FF::copy(z.v(), (a.Z + (size_t)(e) * L));
    bool isinf = FF::is_zero(z.v());
    uint32_t* ox = a.ox + e * a.o_estride;
    uint32_t* oy = a.oy + e * a.o_estride;
    if (a.inf) a.inf[e] = isinf ? 1 : 0;
    if (isinf) {
      FF::set_zero(ox);
      FF::set_zero(oy);
    } else {
      FF::mul(zi.v(), acc.v(), (a.scratch + (size_t)(e) * L));  // 1/Z_e
      FF::mul(acc.v(), acc.v(), z.v());                          // drop Z_e from the running inverse
      FF::sqr(zz.v(), zi.v());
      FF::mul(t.v(), zz.v(), (a.X + (size_t)(e) * L));
      FF::copy(ox, t.v());
      FF::mul(zz.v(), zz.v(), zi.v());
      FF::mul(t.v(), zz.v(), (a.Y + (size_t)(e) * L));
      FF::copy(oy, t.v());
    }
    if (e < (size_t)a.G) break;
  }
}

// EAdd / ESub on L1: affine + affine -> Jacobian (bgn.go:482, 419)
template <int L>
BGN_DEV void g1_add_body(const G1AddArgs& a, size_t e) {
  typedef F<L> FF;
  if (e >= a.count) return;
  size_t e1 = a.bcast1 ? 0 : e;
  Loc<L> X, Y, Z, t0, t1, t2, t3;
  if (a.inf1[e1]) {
    FF::set_zero(X.v());
    FF::set_zero(Y.v());
    FF::set_zero(Z.v());
  } else {
    FF::copy(X.v(), (a.x1 + (size_t)(e1) * L));
    FF::copy(Y.v(), (a.y1 + (size_t)(e1) * L));
    FF::set_one(Z.v());
  }
  if (!a.inf2[e])
    G<L>::madd(X.v(), Y.v(), Z.v(), (a.x2 + (size_t)(e) * L), (a.y2 + (size_t)(e) * L), a.subtract != 0, t0.v(),
               t1.v(), t2.v(), t3.v());
  FF::copy(a.X + (size_t)(e) * L, X.v());
  FF::copy(a.Y + (size_t)(e) * L, Y.v());
  FF::copy(a.Z + (size_t)(e) * L, Z.v());
}

// EAdd / ESub in affine coordinates with a shared inversion (types.h: G1AffAddArgs).
// den(e): the denominator of the chord / tangent slope of element e, or 1 where no slope is needed
// (an operand is O, the operands are inverse to each other, or a 2-torsion point is doubled).
// kind: 0 = result O, 1 = copy operand 2 (signed), 2 = copy operand 1, 3 = chord, 4 = tangent.
template <int L>
BGN_DEV int g1_affadd_den(const G1AffAddArgs& a, size_t e, E den, E b2y, E t) {
  typedef F<L> FF;
  const size_t e1 = a.bcast1 ? 0 : e;
  const bool i1 = a.inf1[e1] != 0, i2 = a.inf2[e] != 0;
  FF::set_one(den);
  if (!i2) {
    FF::copy(b2y, a.y2 + e * L);
    if (a.subtract) FF::neg(b2y, b2y);
  }
  if (i1) return i2 ? 0 : 1;
  if (i2) return 2;
  FF::sub(t, a.x2 + e * L, a.x1 + e1 * L);
  if (!FF::is_zero(t)) {
    FF::copy(den, t);
    return 3;
  }
  if (FF::equal(a.y1 + e1 * L, b2y) && !FF::is_zero(b2y)) {
    FF::add(den, b2y, b2y);
    return 4;
  }
  return 0;
}
template <int L>
BGN_DEV void g1_affadd_body(const G1AffAddArgs& a, size_t g) {
  typedef F<L> FF;
  if (g >= (size_t)a.G || g >= a.count) return;
  Loc<L> acc, den, b2y, t, lam, x3, y3;
  FF::set_one(acc.v());
  for (size_t e = g; e < a.count; e += a.G) {
    g1_affadd_den<L>(a, e, den.v(), b2y.v(), t.v());
    FF::copy(a.scratch + e * L, acc.v());
    FF::mul(acc.v(), acc.v(), den.v());
  }
  FF::inv_gcd_fast(acc.v(), acc.v());
  const size_t last = ((a.count - 1 - g) / a.G) * a.G + g;
  for (size_t e = last;; e -= a.G) {
    const size_t e1 = a.bcast1 ? 0 : e;
    const int kind = g1_affadd_den<L>(a, e, den.v(), b2y.v(), t.v());
    FF::mul(lam.v(), acc.v(), a.scratch + e * L);  // 1 / den(e)
    FF::mul(acc.v(), acc.v(), den.v());            // drop den(e) from the running inverse
    uint32_t* ox = a.ox + e * L;
    uint32_t* oy = a.oy + e * L;
    a.oinf[e] = kind == 0 ? 1 : 0;
    if (kind == 0) {
      FF::set_zero(ox);
      FF::set_zero(oy);
    } else if (kind == 1) {
      FF::copy(ox, a.x2 + e * L);
      FF::copy(oy, b2y.v());
    } else if (kind == 2) {
      FF::copy(ox, a.x1 + e1 * L);
      FF::copy(oy, a.y1 + e1 * L);
    } else {
      const uint32_t* x1 = a.x1 + e1 * L;
      const uint32_t* y1 = a.y1 + e1 * L;
      if (kind == 3) {
        FF::sub(t.v(), b2y.v(), y1);  // y2 - y1
      } else {
        FF::sqr(t.v(), x1);           // 3 x1^2 + 1  (curve y^2 = x^3 + x)
        FF::add(x3.v(), t.v(), t.v());
        FF::add(t.v(), x3.v(), t.v());
        FF::set_one(x3.v());
        FF::add(t.v(), t.v(), x3.v());
      }
      FF::mul(lam.v(), lam.v(), t.v());  // slope
      FF::sqr(x3.v(), lam.v());
      FF::sub(x3.v(), x3.v(), x1);
      FF::sub(x3.v(), x3.v(), kind == 3 ? a.x2 + e * L : x1);
      FF::sub(t.v(), x1, x3.v());
      FF::mul(y3.v(), lam.v(), t.v());
      FF::sub(y3.v(), y3.v(), y1);
      FF::copy(ox, x3.v());
      FF::copy(oy, y3.v());
    }
    if (e < (size_t)a.G) break;
  }
}

// MultConst on L1: k*C, per-element big-endian scalar (bgn.go:258)
template <int L>
BGN_DEV void g1_mulvar_body(const G1MulArgs& a, size_t e) {
  typedef F<L> FF;
  if (e >= a.count) return;
  Loc<L> X, Y, Z, ax, ay, t0, t1, t2, t3;
  FF::set_zero(X.v());
  FF::set_zero(Y.v());
  FF::set_zero(Z.v());
  if (!a.inf[e]) {
    FF::copy(ax.v(), (a.x + (size_t)(e) * L));
    FF::copy(ay.v(), (a.y + (size_t)(e) * L));
    const uint8_t* k = a.k_be + e * a.kbytes;
    bool started = false;  // leading zero bits: doubling O is a no-op
    for (int i = 0; i < a.kbytes; i++) {
      uint32_t byte = k[i];
      for (int bit = 7; bit >= 0; bit--) {
        if (started) G<L>::dbl(X.v(), Y.v(), Z.v(), t0.v(), t1.v(), t2.v(), t3.v());
        if ((byte >> bit) & 1) {
          G<L>::madd(X.v(), Y.v(), Z.v(), ax.w, ay.w, false, t0.v(), t1.v(), t2.v(), t3.v());
          started = true;
        }
      }
    }
  }
  FF::copy(a.X + (size_t)(e) * L, X.v());
  FF::copy(a.Y + (size_t)(e) * L, Y.v());
  FF::copy(a.Z + (size_t)(e) * L, Z.v());
}

// table construction -----------------------------------------------------
// bases[win] = 2^(hb win) * Base in Jacobian coordinates (single thread: hb*nwin doublings)
template <int L>
BGN_DEV void tab_bases_body(const uint32_t* bx, const uint32_t* by, int nwin, int hb, uint32_t* X, uint32_t* Y,
                            uint32_t* Z, size_t N, size_t g) {
  typedef F<L> FF;
  if (g != 0) return;
  Loc<L> x, y, z, t0, t1, t2, t3;
  FF::copy(x.v(), bx);
  FF::copy(y.v(), by);
  FF::set_one(z.v());
  for (int win = 0; win < nwin; win++) {
    FF::copy(X + (size_t)(win) * L, x.v());
    FF::copy(Y + (size_t)(win) * L, y.v());
    FF::copy(Z + (size_t)(win) * L, z.v());
    for (int i = 0; i < hb; i++) G<L>::dbl(x.v(), y.v(), z.v(), t0.v(), t1.v(), t2.v(), t3.v());
  }
}
// entries (win, d), d = 1 .. 2^hb - 1, by repeated addition of the affine base of the window
template <int L>
BGN_DEV void tab_fill_body(const uint32_t* ax, const uint32_t* ay, const uint8_t* ainf, size_t Nb, int nwin, int hb,
                           uint32_t* X, uint32_t* Y, uint32_t* Z, size_t N, size_t g) {
  typedef F<L> FF;
  int win = (int)g;
  if (g >= (size_t)nwin) return;
  const int ents = (1 << hb) - 1;
  Loc<L> x, y, z, t0, t1, t2, t3;
  FF::set_zero(x.v());
  FF::set_zero(y.v());
  FF::set_zero(z.v());
  for (int d = 1; d <= ents; d++) {
    if (!ainf[win])
      G<L>::madd(x.v(), y.v(), z.v(), (ax + (size_t)(win) * L), (ay + (size_t)(win) * L), false, t0.v(), t1.v(), t2.v(),
                 t3.v());
    size_t o = (size_t)win * ents + (d - 1);
    FF::copy(X + (size_t)(o) * L, x.v());
    FF::copy(Y + (size_t)(o) * L, y.v());
    FF::copy(Z + (size_t)(o) * L, z.v());
  }
}

// Wide-window table from a narrow one: a window of nsub * hb bits is nsub windows of hb bits, so entry
// (w, d) is the sum of the narrow entries of d's hb-bit digits, Th[nsub w + k][digit k of d] (nsub - 1
// complete mixed additions per entry, all entries in parallel; Jacobian out).  `first` is the global
// index of this launch's first entry (large tables are built in chunks).  The narrow table has nwin_h
// windows; a wide window that reaches past them leaves the entries with a non-zero digit there as O:
// they lie beyond the scalar's top bit and are never addressed.
template <int L>
BGN_DEV void tabw_fill_body(const uint32_t* tabh, int nwin_h, int nsub, int hb, uint32_t* X, uint32_t* Y, uint32_t* Z,
                            size_t first, size_t nent, size_t id) {
  typedef F<L> FF;
  if (id >= nent) return;
  const size_t ents = ((size_t)1 << (hb * nsub)) - 1;
  const uint32_t mask = (1u << hb) - 1u;
  const size_t gidx = first + id;
  const int w = (int)(gidx / ents);
  const uint32_t d = (uint32_t)(gidx % ents) + 1;
  Loc<L> x, y, z, t0, t1, t2, t3;
  FF::set_zero(x.v());
  FF::set_zero(y.v());
  FF::set_zero(z.v());
  bool valid = true;
  for (int k = 0; k < nsub; k++)
    if (((d >> (hb * k)) & mask) && w * nsub + k >= nwin_h) valid = false;
  if (valid) {
    for (int k = 0; k < nsub; k++) {
      uint32_t dig = (d >> (hb * k)) & mask;
      if (dig) {
        const uint32_t* ent = tabh + ((size_t)(w * nsub + k) * mask + (dig - 1)) * 2 * L;
        G<L>::madd(x.v(), y.v(), z.v(), ent, ent + L, false, t0.v(), t1.v(), t2.v(), t3.v());
      }
    }
  }
  FF::copy(X + id * L, x.v());
  FF::copy(Y + id * L, y.v());
  FF::copy(Z + id * L, z.v());
}

// Weierstrass table -> twisted Edwards table (curve.cuh: Ed): entry (x, y) -> (u, v, u v) with u = x / y,
// v = (x - 1) / (x + 1); `count` entries, AoS x || y in, u || v || t out.  Thread g converts the entries
// g, g + G, ... with one inversion (Montgomery's trick on y (x + 1)).  All-zero entries (O: windows beyond
// the scalar's top bit, never addressed) are written as zeros.  A finite entry with y (x + 1) = 0 is a point
// of order 2 or 4 -- the base point is not of odd order -- and raises *bad: the caller then keeps the
// Weierstrass tables.
template <int L>
BGN_DEV void tab_edwards_body(const uint32_t* tabw, uint32_t* tabe, uint32_t* scratch, size_t count, int G_,
                              int* bad, size_t g) {
  typedef F<L> FF;
  if (g >= (size_t)G_ || g >= count) return;
  Loc<L> acc, den, one, t, u, v;
  FF::set_one(acc.v());
  FF::set_one(one.v());
  auto skip = [&](size_t e) {  // den(e) into `den`; true where the entry takes no part in the shared inversion
    const uint32_t* x = tabw + e * 2 * L;
    FF::add(t.v(), x, one.v());
    FF::mul(den.v(), t.v(), x + L);
    if (!FF::is_zero(den.v())) return false;
    if (!(FF::is_zero(x) && FF::is_zero(x + L))) *bad = 1;
    return true;
  };
  for (size_t e = g; e < count; e += G_) {
    if (skip(e)) continue;
    FF::copy(scratch + e * L, acc.v());
    FF::mul(acc.v(), acc.v(), den.v());
  }
  FF::inv_gcd_fast(acc.v(), acc.v());
  const size_t last = ((count - 1 - g) / G_) * G_ + g;
  for (size_t e = last;; e -= G_) {
    const uint32_t* x = tabw + e * 2 * L;
    uint32_t* o = tabe + e * 3 * L;
    if (skip(e)) {
      FF::set_zero(o);
      FF::set_zero(o + L);
      FF::set_zero(o + 2 * L);
    } else {
      FF::mul(t.v(), acc.v(), scratch + e * L);   // 1 / (y (x + 1))
      FF::mul(acc.v(), acc.v(), den.v());
      FF::add(u.v(), x, one.v());
      FF::mul(u.v(), u.v(), t.v());               // 1 / y
      FF::mul(u.v(), u.v(), x);                   // u = x / y
      FF::mul(v.v(), t.v(), x + L);               // 1 / (x + 1)
      FF::sub(t.v(), x, one.v());
      FF::mul(v.v(), v.v(), t.v());               // v = (x - 1) / (x + 1)
      FF::mul(t.v(), u.v(), v.v());
      FF::canon(o, u.v());
      FF::canon(o + L, v.v());
      FF::canon(o + 2 * L, t.v());
    }
    if (e < (size_t)G_) break;
  }
}

// ------------------------------------------------------------ GT kernels
template <int L>
BGN_DEV void gt_mul_body(const GtBinArgs& a, size_t e) {
  typedef F<L> FF;
  if (e >= a.count) return;
  Loc<L> a0, a1, b0, b1, t0, t1, t2;
  FF::copy(a0.v(), (a.are + (size_t)(e) * L));
  FF::copy(a1.v(), (a.aim + (size_t)(e) * L));
  FF::copy(b0.v(), (a.bre + (size_t)(e) * L));
  FF::copy(b1.v(), (a.bim + (size_t)(e) * L));
  if (a.conj_b) FF::neg(b1.v(), b1.v());
  E2 r = mke2(a0.v(), a1.v());
  FF::mul2(r, r, mke2(b0.v(), b1.v()), t0.v(), t1.v(), t2.v());
  FF::copy(a.ore + (size_t)(e) * L, r.re);
  FF::copy(a.oim + (size_t)(e) * L, r.im);
}

template <int L>
BGN_DEV void gt_pow_body(const GtPowArgs& a, size_t e) {
  typedef F<L> FF;
  if (e >= a.count) return;
  Loc<L> a0, a1, r0, r1, t0, t1, t2;
  E2 in = mke2(a0.v(), a1.v()), r = mke2(r0.v(), r1.v());
  FF::copy(in.re, (a.re + (size_t)(e) * L));
  FF::copy(in.im, (a.im + (size_t)(e) * L));
  if (a.mode == 1) {
    GT<L>::pow_fixed(r, in, t0.v(), t1.v(), t2.v());
  } else if (a.mode == 2) {
    FF::conj2(r, in);
  } else {
    GT<L>::pow_var(r, in, a.e_be + e * a.ebytes, a.ebytes, t0.v(), t1.v(), t2.v());
    if (a.mode == 3) FF::neg(r.im, r.im);
  }
  FF::copy(a.ore + (size_t)(e) * L, r.re);
  FF::copy(a.oim + (size_t)(e) * L, r.im);
}

// One pass of the GT product tree of an L2 sum (bgn.go:460 folded over terms):
// in[t*ncoeff + c], t < nterms  ->  out[g*ncoeff + c] = prod_{t = g (mod G)} in[t][c]
template <int L>
BGN_DEV void gt_reduce_body(const uint32_t* re, const uint32_t* im, size_t Nin, size_t nterms, int ncoeff, int G,
                            uint32_t* ore, uint32_t* oim, size_t N, size_t id) {
  typedef F<L> FF;
  if (id >= (size_t)G * ncoeff) return;
  size_t g = id / ncoeff;
  int c = (int)(id % ncoeff);
  Loc<L> r0, r1, b0, b1, t0, t1, t2;
  E2 acc = mke2(r0.v(), r1.v()), b = mke2(b0.v(), b1.v());
  FF::set_one2(acc);
  for (size_t t = g; t < nterms; t += G) {
    size_t e = t * ncoeff + c;
    FF::copy(b.re, (re + (size_t)(e) * L));
    FF::copy(b.im, (im + (size_t)(e) * L));
    FF::mul2(acc, acc, b, t0.v(), t1.v(), t2.v());
  }
  FF::copy(ore + (size_t)(id) * L, acc.re);
  FF::copy(oim + (size_t)(id) * L, acc.im);
}

// Level-2 re-randomisation: out = a * E^r, E = e(Q,Q), through the fixed-base table of E
// (bgn.go:283-287, 306-310, 469-474; the reference computes the pairing e(Q,Q) anew per call).
template <int L>
BGN_DEV void gt_blind_body(const GtBlindArgs& a, size_t e) {
  typedef F<L> FF;
  if (e >= a.count) return;
  Loc<L> r0, r1, t0, t1, t2;
  E2 acc = mke2(r0.v(), r1.v());
  FF::copy(acc.re, a.re + e * L);
  FF::copy(acc.im, a.im + e * L);
  const uint8_t* r = a.r_be + e * a.rbytes;
  for (int win = 0; win < a.rbytes; win++) {
    uint32_t d = r[a.rbytes - 1 - win];
    if (d) {
      uint32_t* ent = const_cast<uint32_t*>(a.tabE) + ((size_t)win * 255 + (d - 1)) * 2 * L;
      FF::mul2(acc, acc, mke2(ent, ent + L), t0.v(), t1.v(), t2.v());
    }
  }
  FF::copy(a.ore + e * L, acc.re);
  FF::copy(a.oim + e * L, acc.im);
}

// Fixed-base table of a GT element: bases[w] = gen^(256^w) (single thread), then entry
// (w, d) = bases[w]^d for d = 1..255 (one thread per window), AoS [w][d-1][re||im], canonical.
template <int L>
BGN_DEV void gt_tab_bases_body(const uint32_t* gen, int nwin, uint32_t* bases, size_t g) {
  typedef F<L> FF;
  if (g != 0) return;
  Loc<L> r0, r1, t0, t1;
  E2 acc = mke2(r0.v(), r1.v());
  FF::copy(acc.re, gen);
  FF::copy(acc.im, gen + L);
  for (int w = 0; w < nwin; w++) {
    FF::canon(bases + (size_t)w * 2 * L, acc.re);
    FF::canon(bases + (size_t)w * 2 * L + L, acc.im);
    for (int i = 0; i < 8; i++) FF::sqr2(acc, acc, t0.v(), t1.v());
  }
}
template <int L>
BGN_DEV void gt_tab_fill_body(const uint32_t* bases, int nwin, uint32_t* tab, size_t g) {
  typedef F<L> FF;
  if (g >= (size_t)nwin) return;
  Loc<L> r0, r1, b0, b1, t0, t1, t2;
  E2 acc = mke2(r0.v(), r1.v()), b = mke2(b0.v(), b1.v());
  FF::copy(b.re, bases + g * 2 * L);
  FF::copy(b.im, bases + g * 2 * L + L);
  FF::copy2(acc, b);
  for (int d = 1; d <= 255; d++) {
    uint32_t* dst = tab + (g * 255 + (d - 1)) * 2 * L;
    FF::canon(dst, acc.re);
    FF::canon(dst + L, acc.im);
    FF::mul2(acc, acc, b, t0.v(), t1.v(), t2.v());
  }
}

BGN_DEV bool conv_bit(const PolyConvArgs& a, int k, int b) {
  return ((b < 64 ? a.w[k] >> b : a.whi[k] >> (b - 64)) & 1) != 0;
}
// Integer-weighted correlation (types.h: PolyConvArgs), one thread per output slot; bit-plane
// Horner over the weights: acc <- 2 acc + sum_{k: bit b of w[k]} in[j-k].  Every addition is a
// complete mixed addition with an affine input, so no general Jacobian addition is needed.
template <int L>
BGN_DEV void g1_polyconv_body(const PolyConvArgs& a, size_t id) {
  typedef F<L> FF;
  if (id >= a.count * (size_t)a.j_count) return;
  size_t u = id / (size_t)a.j_count;
  int j = a.j_begin + (int)(id % (size_t)a.j_count);
  Loc<L> X, Y, Z, t0, t1, t2, t3;
  FF::set_zero(X.v());
  FF::set_zero(Y.v());
  FF::set_zero(Z.v());
  bool started = false;  // doubling O is a no-op
  for (int b = a.top_bit; b >= 0; b--) {
    if (started) G<L>::dbl(X.v(), Y.v(), Z.v(), t0.v(), t1.v(), t2.v(), t3.v());
    for (int k = 0; k < a.nw; k++) {
      int i = j - k;
      if (i < 0 || i >= a.d || !conv_bit(a, k, b)) continue;
      size_t e = u * a.d + i;
      if (a.inf[e]) continue;
      G<L>::madd(X.v(), Y.v(), Z.v(), a.x + e * L, a.y + e * L, false, t0.v(), t1.v(), t2.v(), t3.v());
      started = true;
    }
  }
  if (a.negate) FF::neg(Y.v(), Y.v());
  FF::copy(a.X + id * L, X.v());
  FF::copy(a.Y + id * L, Y.v());
  FF::copy(a.Z + id * L, Z.v());
}
template <int L>
BGN_DEV void gt_polyconv_body(const PolyConvArgs& a, size_t id) {
  typedef F<L> FF;
  if (id >= a.count * (size_t)a.j_count) return;
  size_t u = id / (size_t)a.j_count;
  int j = a.j_begin + (int)(id % (size_t)a.j_count);
  Loc<L> r0, r1, t0, t1, t2;
  E2 acc = mke2(r0.v(), r1.v());
  FF::set_one2(acc);
  bool started = false;
  for (int b = a.top_bit; b >= 0; b--) {
    if (started) FF::sqr2(acc, acc, t0.v(), t1.v());
    for (int k = 0; k < a.nw; k++) {
      int i = j - k;
      if (i < 0 || i >= a.d || !conv_bit(a, k, b)) continue;
      size_t e = u * a.d + i;
      FF::mul2(acc, acc, mke2(const_cast<uint32_t*>(a.x) + e * L, const_cast<uint32_t*>(a.y) + e * L), t0.v(), t1.v(),
               t2.v());
      started = true;
    }
  }
  if (a.negate) FF::neg(acc.im, acc.im);  // unitary: conj = inverse
  FF::copy(a.X + id * L, acc.re);
  FF::copy(a.Y + id * L, acc.im);
}

// ------------------------------------------------------------ BSGS (gsbs.go)
// Baby steps: elems[j] = gen^(j+1), j < S, canonical Montgomery AoS [S][2L];
// open-addressing hash table slots[hmask+1] holding j+1 (0 = empty).
// hashed on the real part only: gsk^m and gsk^-m share it and only positive m are stored, so the
// same table serves the full-element search below and the trace search of lucas.cuh
BGN_DEV uint32_t bsgs_hash(const uint32_t* re, const uint32_t* im) {
  (void)im;
  return bsgs_hash_re(re);
}

template <int L>
BGN_DEV void bsgs_build_body(const BsgsBuildArgs& a, size_t g) {
  typedef F<L> FF;
  uint64_t j0 = (uint64_t)g * a.chunk;
  if (j0 >= a.S) return;
  Loc<L> g0, g1, r0, r1, t0, t1, t2;
  FF::copy(g0.v(), a.gen);
  FF::copy(g1.v(), a.gen + L);
  E2 gg = mke2(g0.v(), g1.v()), rr = mke2(r0.v(), r1.v());
  // rr = gen^(j0+1)
  uint64_t ex = j0 + 1;
  FF::set_one2(rr);
  for (int bit = 40; bit >= 0; bit--) {
    FF::sqr2(rr, rr, t0.v(), t1.v());
    if ((ex >> bit) & 1) FF::mul2(rr, rr, gg, t0.v(), t1.v(), t2.v());
  }
  for (int i = 0; i < a.chunk && j0 + i < a.S; i++) {
    uint32_t j = (uint32_t)(j0 + i);
    uint32_t* dst = a.elems + (size_t)j * 2 * L;
    FF::canon(dst, rr.re);
    FF::canon(dst + L, rr.im);
    uint32_t h = bsgs_hash(dst, dst + L) & a.hmask;
    while (atomicCAS(a.slots + h, 0u, j + 1) != 0u) h = (h + 1) & a.hmask;
    FF::mul2(rr, rr, gg, t0.v(), t1.v(), t2.v());
  }
}

// index of the canonical element (c0, c1) in the baby-step table, or -1
template <int L>
BGN_DEVNI int64_t bsgs_probe(const BsgsLookupArgs& a, const uint32_t* c0, const uint32_t* c1) {
  uint32_t h = bsgs_hash(c0, c1) & a.hmask;
  for (;;) {
    uint32_t s = a.slots[h];
    if (s == 0) return -1;
    const uint32_t* el = a.elems + (size_t)(s - 1) * 2 * L;
    uint32_t diff = 0;
    for (int k = 0; k < L; k++) diff |= (el[k] ^ c0[k]) | (el[L + k] ^ c1[k]);
    if (diff == 0) return (int64_t)(s - 1);
    h = (h + 1) & a.hmask;
  }
}

template <int L>
BGN_DEV void bsgs_lookup_body(const BsgsLookupArgs& a, size_t e) {
  typedef F<L> FF;
  if (e >= a.count) return;
  Loc<L> p0, p1, n0, n1, gi0, gi1, c0, c1, t0, t1, t2;
  FF::copy(p0.v(), (a.re + (size_t)(e) * L));
  FF::copy(p1.v(), (a.im + (size_t)(e) * L));
  // identity => 0 (recoverMessage, bgn.go:359-363)
  if (FF::is_one(p0.v()) && FF::is_zero(p1.v())) {
    a.out[e] = 0;
    a.status[e] = 0;
    return;
  }
  // conj(csk) = csk^-1 (GT is unitary) is the Neg(ct) retry of bgn.go:235-241: the positive and
  // the negated candidate walk the giant steps together, so a negative plaintext costs no retry
  FF::copy(n0.v(), p0.v());
  FF::neg(n1.v(), p1.v());
  FF::copy(gi0.v(), a.ginv);
  FF::copy(gi1.v(), a.ginv + L);
  E2 vp = mke2(p0.v(), p1.v()), vn = mke2(n0.v(), n1.v()), vg = mke2(gi0.v(), gi1.v());
  for (uint32_t i = 0; i < a.giant_steps; i++) {
    FF::canon(c0.v(), vp.re);
    FF::canon(c1.v(), vp.im);
    int64_t j = bsgs_probe<L>(a, c0.w, c1.w);
    if (j >= 0) {
      uint64_t m = (uint64_t)i * a.S + (uint64_t)j + 1;
      if (m <= a.mmax) {
        a.out[e] = (int64_t)m;
        a.status[e] = 0;
        return;
      }
    }
    FF::canon(c0.v(), vn.re);
    FF::canon(c1.v(), vn.im);
    j = bsgs_probe<L>(a, c0.w, c1.w);
    if (j >= 0) {
      uint64_t m = (uint64_t)i * a.S + (uint64_t)j + 1;
      if (m <= a.mmax) {
        a.out[e] = -(int64_t)m;
        a.status[e] = 0;
        return;
      }
    }
    if (i + 1 < a.giant_steps) {
      FF::mul2(vp, vp, vg, t0.v(), t1.v(), t2.v());
      FF::mul2(vn, vn, vg, t0.v(), t1.v(), t2.v());
    }
  }
  a.out[e] = 0;
  a.status[e] = 1;
}

// Decrypt = Lucas ladder for C^q1 + table search, a pair of lanes per ciphertext (lucas.cuh).
// CPU simulation: both roles of one pair in lockstep.
#ifdef BGN_HOSTSIM
template <int L>
void dec_lucas_pair_sim(const DecLucasArgs& a, size_t e) {
  typedef Lucas<L> LU;
  typename LU::State s0, s1;
  LU::init(s0, a, e, true);
  LU::init(s1, a, e, true);
  for (int i = LU::nbits() - 2; i >= 0; i--) {
    uint32_t r0[L], r1[L];
    LU::step(r0, s0, LU::bit(i), 0);
    LU::step(r1, s1, LU::bit(i), 1);
    LU::update(s0, r0, r1, LU::bit(i), 0);
    LU::update(s1, r1, r0, LU::bit(i), 1);
  }
  LU::finish(s0, a, e);
}
#endif

#ifdef BGN_HOSTSIM
// CPU simulation of k_gt_pow_pair: both lanes of one pair in lockstep
template <int L>
void gt_pow_pair_sim(const GtPowArgs& a, size_t e) {
  typedef GtPowPair<L> GP;
  typename GP::State s0, s1;
  GP::init(s0, a.re + e * L, a.im + e * L, true);
  GP::init(s1, a.re + e * L, a.im + e * L, true);
  for (int i = 1; i < c_pc.exp_naf_len; i++) {
    uint32_t t0[L], t1[L];
    GP::sqr_half(t0, s0, 0);
    GP::sqr_half(t1, s1, 1);
    GP::update(s0, t0, t1, 0);
    GP::update(s1, t1, t0, 1);
    const int dg = c_pc.exp_naf[i];
    if (dg != 0) {
      GP::mul_half(t0, s0, 0, dg < 0);
      GP::mul_half(t1, s1, 1, dg < 0);
      GP::update(s0, t0, t1, 0);
      GP::update(s1, t1, t0, 1);
    }
  }
  GP::finish(s0, a.ore + e * L, a.oim + e * L, 0);
  GP::finish(s1, a.ore + e * L, a.oim + e * L, 1);
}
#endif

// =========================================================== CUDA wrappers
#ifndef BGN_HOSTSIM
#define BGN_KERNEL_1D(NAME, ARGT)                                                   \
  template <int L>                                                                  \
  __global__ void __launch_bounds__(128) k_##NAME(const __grid_constant__ ARGT a) { \
    NAME##_body<L>(a, BGN_GID(size_t));                                             \
  }
BGN_KERNEL_1D(encrypt, EncArgs)
BGN_KERNEL_1D(normalize, NormArgs)
BGN_KERNEL_1D(g1_add, G1AddArgs)
BGN_KERNEL_1D(g1_mulvar, G1MulArgs)
BGN_KERNEL_1D(g1_affadd, G1AffAddArgs)
BGN_KERNEL_1D(gt_mul, GtBinArgs)
BGN_KERNEL_1D(gt_pow, GtPowArgs)
BGN_KERNEL_1D(bsgs_build, BsgsBuildArgs)
BGN_KERNEL_1D(bsgs_lookup, BsgsLookupArgs)
BGN_KERNEL_1D(gt_blind, GtBlindArgs)
BGN_KERNEL_1D(g1_polyconv, PolyConvArgs)
BGN_KERNEL_1D(gt_polyconv, PolyConvArgs)
template <int L>
__global__ void k_gt_tab_bases(const uint32_t* gen, int nwin, uint32_t* bases) {
  gt_tab_bases_body<L>(gen, nwin, bases, BGN_GID(size_t));
}
template <int L>
__global__ void k_gt_tab_fill(const uint32_t* bases, int nwin, uint32_t* tab) {
  gt_tab_fill_body<L>(bases, nwin, tab, BGN_GID(size_t));
}

template <int L>
__global__ void __launch_bounds__(64) k_dec_lucas(const __grid_constant__ DecLucasArgs a) {
  typedef Lucas<L> LU;
  const size_t gid = BGN_GID(size_t);
  const size_t e = gid >> 1;
  const int s = (int)(gid & 1);
  const bool active = e < a.count;
  typename LU::State st;
  LU::init(st, a, e, active);
  BGN_UNROLL1
  for (int i = LU::nbits() - 2; i >= 0; i--) {
    const int b = LU::bit(i);
    uint32_t mine[L], other[L];
    LU::step(mine, st, b, s);
#pragma unroll
    for (int j = 0; j < L; j++) other[j] = __shfl_xor_sync(0xffffffffu, mine[j], 1);
    LU::update(st, mine, other, b, s);
  }
  if (active && s == 0) LU::finish(st, a, e);
}

// a^q1 for the fixed exponent c_pc.exp, a lane pair per element (GtPowArgs mode 1 only)
template <int L>
__global__ void __launch_bounds__(64) k_gt_pow_pair(const __grid_constant__ GtPowArgs a) {
  typedef GtPowPair<L> GP;
  const size_t gid = BGN_GID(size_t);
  const size_t e = gid >> 1;
  const int s = (int)(gid & 1);
  const bool active = e < a.count;
  const size_t ee = active ? e : 0;
  typename GP::State st;
  GP::init(st, a.re + ee * L, a.im + ee * L, active);
  BGN_UNROLL1
  for (int i = 1; i < c_pc.exp_naf_len; i++) {  // signed digits, MSB first (digit 0 is the leading 1)
    uint32_t mine[L], other[L];
    GP::sqr_half(mine, st, s);
#pragma unroll
    for (int j = 0; j < L; j++) other[j] = __shfl_xor_sync(0xffffffffu, mine[j], 1);
    GP::update(st, mine, other, s);
    const int dg = c_pc.exp_naf[i];
    if (dg != 0) {
      GP::mul_half(mine, st, s, dg < 0);
#pragma unroll
      for (int j = 0; j < L; j++) other[j] = __shfl_xor_sync(0xffffffffu, mine[j], 1);
      GP::update(st, mine, other, s);
    }
  }
  if (active) GP::finish(st, a.ore + e * L, a.oim + e * L, s);
}

// (De)serialisation kernels: a block stages its elements' bytes in shared memory so that global
// memory sees whole 32-bit words at consecutive addresses (the byte format has an odd element size,
// 2B = 130 at 512 bit: one thread per element touching its bytes directly is a stride-130 byte
// access, 32 sectors per warp instruction).  Block b owns elements [b*nt, (b+1)*nt); its byte range
// starts at b*nt*2B, a multiple of 4 for every block size that is a multiple of 2.
BGN_DEV void stage_bytes_in(uint8_t* sm, const uint8_t* g, size_t nbytes) {
  // g is 4-byte aligned when the caller's buffer is (cudaMalloc / arena / torch); fall back to bytes otherwise
  if ((reinterpret_cast<uintptr_t>(g) & 3) == 0) {
    const uint32_t* gw = reinterpret_cast<const uint32_t*>(g);
    uint32_t* sw = reinterpret_cast<uint32_t*>(sm);
    for (size_t i = threadIdx.x; i < nbytes / 4; i += blockDim.x) sw[i] = gw[i];
    for (size_t i = (nbytes & ~(size_t)3) + threadIdx.x; i < nbytes; i += blockDim.x) sm[i] = g[i];
  } else {
    for (size_t i = threadIdx.x; i < nbytes; i += blockDim.x) sm[i] = g[i];
  }
}
BGN_DEV void stage_bytes_out(uint8_t* g, const uint8_t* sm, size_t nbytes) {
  if ((reinterpret_cast<uintptr_t>(g) & 3) == 0) {
    uint32_t* gw = reinterpret_cast<uint32_t*>(g);
    const uint32_t* sw = reinterpret_cast<const uint32_t*>(sm);
    for (size_t i = threadIdx.x; i < nbytes / 4; i += blockDim.x) gw[i] = sw[i];
    for (size_t i = (nbytes & ~(size_t)3) + threadIdx.x; i < nbytes; i += blockDim.x) g[i] = sm[i];
  } else {
    for (size_t i = threadIdx.x; i < nbytes; i += blockDim.x) g[i] = sm[i];
  }
}
// elements of this block and their first index
BGN_DEV size_t block_first() { return (size_t)blockIdx.x * blockDim.x; }
BGN_DEV size_t block_count(size_t count) {
  size_t first = block_first();
  return first >= count ? 0 : (count - first < blockDim.x ? count - first : blockDim.x);
}

template <int L>
__global__ void k_g1_from_bytes(const uint8_t* __restrict__ in, int B, size_t count, uint32_t* x, uint32_t* y,
                                uint8_t* inf, size_t N) {
  extern __shared__ __align__(16) uint8_t smem_io[];
  const size_t first = block_first(), n = block_count(count);
  stage_bytes_in(smem_io, in + first * 2 * B, n * 2 * B);
  __syncthreads();
  g1_from_bytes_body<L>(smem_io, B, n, x + first * L, y + first * L, inf + first, N, threadIdx.x);
}
template <int L>
__global__ void k_g1_to_bytes(const uint32_t* x, const uint32_t* y, const uint8_t* inf, size_t N, size_t count,
                              uint8_t* __restrict__ out, int B) {
  extern __shared__ __align__(16) uint8_t smem_io[];
  const size_t first = block_first(), n = block_count(count);
  g1_to_bytes_body<L>(x + first * L, y + first * L, inf + first, N, n, smem_io, B, threadIdx.x);
  __syncthreads();
  stage_bytes_out(out + first * 2 * B, smem_io, n * 2 * B);
}
template <int L>
__global__ void k_fp2_from_bytes(const uint8_t* __restrict__ in, int B, size_t count, uint32_t* re, uint32_t* im,
                                 size_t N) {
  extern __shared__ __align__(16) uint8_t smem_io[];
  const size_t first = block_first(), n = block_count(count);
  stage_bytes_in(smem_io, in + first * 2 * B, n * 2 * B);
  __syncthreads();
  fp2_from_bytes_body<L>(smem_io, B, n, re + first * L, im + first * L, N, threadIdx.x);
}
// the grouped form (grp > 0: identity padding slots) indexes source elements irregularly: unstaged
template <int L>
__global__ void k_fp2_to_bytes(const uint32_t* re, const uint32_t* im, size_t N, size_t count,
                               uint8_t* __restrict__ out, int B, int grp, int pad) {
  extern __shared__ __align__(16) uint8_t smem_io[];
  if (grp > 0) {
    fp2_to_bytes_body<L>(re, im, N, count, out, B, grp, pad, BGN_GID(size_t));
    return;
  }
  const size_t first = block_first(), n = block_count(count);
  fp2_to_bytes_body<L>(re + first * L, im + first * L, N, n, smem_io, B, 0, 0, threadIdx.x);
  __syncthreads();
  stage_bytes_out(out + first * 2 * B, smem_io, n * 2 * B);
}
template <int L>
__global__ void k_tab_bases(const uint32_t* bx, const uint32_t* by, int nwin, int hb, uint32_t* X, uint32_t* Y,
                            uint32_t* Z, size_t N) {
  tab_bases_body<L>(bx, by, nwin, hb, X, Y, Z, N, BGN_GID(size_t));
}
template <int L>
__global__ void k_tab_fill(const uint32_t* ax, const uint32_t* ay, const uint8_t* ainf, size_t Nb, int nwin, int hb,
                           uint32_t* X, uint32_t* Y, uint32_t* Z, size_t N) {
  tab_fill_body<L>(ax, ay, ainf, Nb, nwin, hb, X, Y, Z, N, BGN_GID(size_t));
}
template <int L>
__global__ void __launch_bounds__(128) k_tabw_fill(const uint32_t* tabh, int nwin_h, int nsub, int hb, uint32_t* X,
                                                   uint32_t* Y, uint32_t* Z, size_t first, size_t nent) {
  tabw_fill_body<L>(tabh, nwin_h, nsub, hb, X, Y, Z, first, nent, BGN_GID(size_t));
}
template <int L>
__global__ void __launch_bounds__(128) k_tab_edwards(const uint32_t* tabw, uint32_t* tabe, uint32_t* scratch, size_t count,
                                                     int G_, int* bad) {
  tab_edwards_body<L>(tabw, tabe, scratch, count, G_, bad, BGN_GID(size_t));
}
template <int L>
__global__ void __launch_bounds__(128) k_gt_reduce(const uint32_t* re, const uint32_t* im, size_t Nin, size_t nterms,
                                                   int ncoeff, int G, uint32_t* ore, uint32_t* oim, size_t N) {
  gt_reduce_body<L>(re, im, Nin, nterms, ncoeff, G, ore, oim, N, BGN_GID(size_t));
}

// the Miller team kernel: 1 or 2 barrier groups of 128 threads per block; blocks per SM are bounded
// by shared memory (13 element slots per thread)
template <int L>
__global__ void __launch_bounds__(256, 1) k_miller(const __grid_constant__ MillerArgs a) {
  extern __shared__ uint32_t smem_dyn[];
  MillerTeam<L> T(a, smem_dyn, threadIdx.x, blockIdx.x, blockDim.x);
  const int bar_id = 1 + T.group, bar_n = a.group_threads;
  if ((T.group & 1) && a.skew_cycles > 0) {
    long long t0 = clock64();
    while (clock64() - t0 < a.skew_cycles) {
    }
  }
  T.run([=] { asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory"); });
}

// the team kernel with 10 instead of 12 shared-memory slots per thread (evaluation points read from the
// batch arrays): 10 warps per SM instead of 8, one barrier group (pairing.cuh: MillerTeam<L, true>)
template <int L>
__global__ void __launch_bounds__(320) k_miller_wide(const __grid_constant__ MillerArgs a) {
  extern __shared__ uint32_t smem_dyn[];
  MillerTeam<L, true> T(a, smem_dyn, threadIdx.x, blockIdx.x, blockDim.x);
  T.run([] { __syncthreads(); });
}

// the team kernel with two threads per output-slot pair (teamsplit.cuh): batches below one wave
template <int L>
__global__ void __launch_bounds__(384) k_miller_split(const __grid_constant__ MillerArgs a) {
  extern __shared__ uint32_t smem_dyn[];
  MillerSplit<L> T(a, smem_dyn, threadIdx.x, blockIdx.x, blockDim.x);
  T.run([] { __syncthreads(); });
}

// fixed-first-argument pairing: one thread per evaluation point, no barriers
template <int L>
__global__ void __launch_bounds__(256, 1) k_miller_fixed(const __grid_constant__ MillerFixedArgs a) {
  extern __shared__ uint32_t smem_dyn[];
  MillerFixed<L>::run(a, smem_dyn, threadIdx.x, blockDim.x, BGN_GID(size_t));
}

// the same pairing on a lane pair per evaluation point (pairlane.cuh): registers only, three shuffle
// exchanges per Miller step
template <int L>
__global__ void __launch_bounds__(64) k_miller_fixed_pair(const __grid_constant__ MillerFixedArgs a) {
  const size_t gid = BGN_GID(size_t);
  const size_t e = gid >> 1;
  const int s = (int)(gid & 1);
  MillerFixedPair<L>::run(a, e, s, e < (size_t)a.count, [](uint32_t (&other)[L], const uint32_t (&mine)[L]) {
#pragma unroll
    for (int j = 0; j < L; j++) other[j] = __shfl_xor_sync(0xffffffffu, mine[j], 1);
  });
}

// a general pairing on two warps (pairwarp.cuh): even warps advance the Miller points of 32 pairings
// and publish the line ingredients, odd warps fold the lines into the accumulators one step behind;
// one barrier per step.  BLOCKBAR: the barrier spans the block, so all its warp pairs stay in lockstep
// and share one instruction stream per role (otherwise each pair has its own named barrier).
template <int L, int U, bool BLOCKBAR>
__global__ void __launch_bounds__(256) k_pair_duo(const __grid_constant__ PairDuoArgs a) {
  extern __shared__ uint32_t smem_dyn[];
  typedef MillerDuo<L, U> T_;
  const int np = blockDim.x >> 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pt = (warp >> 1) * 32 + lane;
  T_ T(a, smem_dyn, np, pt, (size_t)blockIdx.x * np + pt);
  const int bar_id = BLOCKBAR ? 0 : 1 + (warp >> 1), bar_n = BLOCKBAR ? (int)blockDim.x : 64;
  auto sync = [=] {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_n) : "memory");
  };
  // Iteration s: the X warps compute step s while the F warps consume step s - 1; then ONE barrier, at
  // one call site for both roles (every thread executes the same sequence of barrier instructions:
  // compute-sanitizer synccheck is clean).  Step s publishes into buffer s & 1, which the F warps
  // finished reading before the previous barrier.
  const bool is_x = (warp & 1) == 0;
  if (is_x)
    T.x_init();
  else
    T.f_init();
  int prev_op = -1, s = 0;
  auto iter = [&](int op_x) {
    if (is_x) {
      if (op_x >= 0) T.x_step(op_x, s & 1);
    } else if (prev_op >= 0) {
      T.f_step(prev_op, (s - 1) & 1, s == 1);
    }
    sync();
    prev_op = op_x;
    s++;
  };
  T_::for_steps([&](int, int op) { iter(op); });
  iter(-1);
  if (!is_x) T.f_finish();
}

template <int L>
__global__ void __launch_bounds__(32) k_miller_record(const uint32_t* px, const uint32_t* py, uint32_t* lines,
                                                      uint32_t* scratch, int* ok) {
  if (BGN_GID(size_t) == 0) MillerFixed<L>::record(px, py, lines, scratch, ok);
}

// Register-resident Montgomery products: `iters` dependent modmuls per thread on
// `ILP` independent chains.  Used by bench.py to relate the pairing kernels to
// the best the mulmod itself reaches.
template <int L, int ILP>
__global__ void __launch_bounds__(128) k_mulmod_bench(uint32_t* io, size_t N, int iters) {
  typedef Fp<L> P;
  size_t e = BGN_GID(size_t);
  uint32_t a[ILP][L], b[L];
#pragma unroll
  for (int c = 0; c < ILP; c++)
#pragma unroll
    for (int j = 0; j < L; j++) a[c][j] = io[(size_t)j * N + e];
#pragma unroll
  for (int j = 0; j < L; j++) b[j] = io[(size_t)j * N + e];
#pragma unroll
  for (int c = 0; c < ILP; c++) a[c][0] += c;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < ILP; c++) P::mul(a[c], a[c], b);
  }
#pragma unroll
  for (int c = 1; c < ILP; c++)
#pragma unroll
    for (int j = 0; j < L; j++) a[0][j] ^= a[c][j];
#pragma unroll
  for (int j = 0; j < L; j++) io[(size_t)j * N + e] = a[0][j];
}
// The same products through the memory-operand primitives (F<L>::mul etc. on shared-memory slots,
// laid out as in k_miller): mode 10 = mul only, 11 = the line_mul sequence (5 mul + 6 add/sub),
// 12 = sqr2, 13 = dbl_line.  Counts `iters` sequences per thread; used by tools/microbench.py to
// separate call / load-store / add-sub overhead from the register-resident product.
template <int L>
__global__ void __launch_bounds__(256, 1) k_prim_bench(uint32_t* io, size_t N, int iters, int mode) {
  extern __shared__ uint32_t smem_dyn[];
  typedef F<L> FF;
  const int nt = blockDim.x, tid = threadIdx.x;
  auto slot = [&](int k) -> E { return smem_dyn + ((size_t)k * nt + tid) * L; };
  size_t e = BGN_GID(size_t);
  for (int k = 0; k < 13; k++) {
    for (int j = 0; j < L; j++) slot(k)[j] = io[(size_t)j * N + e];
    slot(k)[0] += k;
  }
  for (int i = 0; i < iters; i++) {
    if (mode == 10) {
      FF::mul_stream(slot(0), slot(0), slot(1));
    } else if (mode == 20) {
      FF::mul(slot(0), slot(0), slot(1));
    } else if (mode == 22) {
      FF::mul_pair(slot(0), slot(0), slot(1), slot(2), slot(2), slot(3));
    } else if (mode == 11) {
      FF::line_mul(mke2(slot(0), slot(1)), slot(7), slot(8), slot(9), slot(4), slot(5), slot(10), slot(11),
                   slot(12));
    } else if (mode == 12) {
      FF::sqr2(mke2(slot(0), slot(1)), mke2(slot(0), slot(1)), slot(10), slot(11));
#define BGN_PRIM_FUSED(BASE, UU)                                                                             \
    } else if (mode == BASE) {                                                                                 \
      MF<L, UU>::line_mul(slot(0), slot(1), slot(7), slot(8), slot(9), slot(4), slot(5));                       \
    } else if (mode == BASE + 2) {                                                                             \
      MF<L, UU>::sqr2(slot(0), slot(1));                                                                        \
    } else if (mode == BASE + 3) {                                                                             \
      MF<L, UU>::dbl_line(slot(4), slot(5), slot(6), slot(7), slot(8), slot(9));                                \
    } else if (mode == BASE + 4) {                                                                             \
      MF<L, UU>::madd_line(slot(4), slot(5), slot(6), slot(10), slot(11), (i & 1) != 0, slot(7), slot(8), slot(9));
    // fused routines (fused.cuh): the variant the Miller kernel ships, and, for L = 17, the others
    BGN_PRIM_FUSED(30, BGN_MILLER_LOOP)
#ifdef BGN_PRIM_UNROLLED
    BGN_PRIM_FUSED(40, 0)
    BGN_PRIM_FUSED(50, 2)
    BGN_PRIM_FUSED(60, 4)
    BGN_PRIM_FUSED(70, 1)
    } else if (mode == 80) {  // lazy reduction: 5 products + 4 reductions
      MF<L, 0>::template line_mul_lazy<0>(slot(0), slot(1), slot(7), slot(8), slot(9), slot(4), slot(5));
    } else if (mode == 81) {  // + Karatsuba on the three double-width products
      MF<L, 0>::template line_mul_lazy<1>(slot(0), slot(1), slot(7), slot(8), slot(9), slot(4), slot(5));
    } else if (mode == 82) {  // + Karatsuba on all five multiplications
      MF<L, 0>::template line_mul_lazy<2>(slot(0), slot(1), slot(7), slot(8), slot(9), slot(4), slot(5));
    } else if (mode == 83) {  // lazy reduction with the independent products interleaved (ILP 2)
      MF<L, 0>::line_mul_lazy_il(slot(0), slot(1), slot(7), slot(8), slot(9), slot(4), slot(5));
#endif
    } else {
      G<L>::dbl_line(slot(4), slot(5), slot(6), slot(7), slot(8), slot(9), slot(10), slot(11), slot(12));
    }
  }
  FF::add(slot(0), slot(0), slot(4));
  FF::add(slot(0), slot(0), slot(1));
  for (int j = 0; j < L; j++) io[(size_t)j * N + e] = slot(0)[j];
}
#endif  // !BGN_HOSTSIM
