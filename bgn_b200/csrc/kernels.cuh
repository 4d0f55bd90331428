// kernels.cuh -- the sm_100a kernels of the batched BGN engine (templated on
// the limb count L).  Device data layout: limb-major SoA, array[limb][N] of
// u32 in Montgomery form ("lazy" range [0,2p)); G1 arrays carry an extra
// byte-per-element infinity flag.  See DESIGN.md for the per-kernel roofline.
//
// Every kernel is a thin __global__ wrapper around a `*_body(args, index)`
// function so the CPU-side unit tests (tests/hostsim, BGN_HOSTSIM) can run the
// very same per-thread program; the wrappers only exist in the CUDA build.
#pragma once
#include "pairing.cuh"

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
  if (e >= count) return;
  uint32_t a[L], b[L];
  be_bytes_to_limbs<L>(a, in + e * 2 * B, B);
  be_bytes_to_limbs<L>(b, in + e * 2 * B + B, B);
  uint32_t o = 0;
  BGN_UNROLL
  for (int j = 0; j < L; j++) o |= a[j] | b[j];
  V vx = mkv(x + e, (int)N), vy = mkv(y + e, (int)N);
  F<L>::to_mont(vx, mkv(a, 1));
  F<L>::to_mont(vy, mkv(b, 1));
  bool isinf = (o == 0);
  if (!isinf) isinf = !G<L>::on_curve(vx, vy);
  inf[e] = isinf ? 1 : 0;
}

template <int L>
BGN_DEV void g1_to_bytes_body(const uint32_t* x, const uint32_t* y, const uint8_t* inf, size_t N, size_t count,
                              uint8_t* out, int B, size_t e) {
  if (e >= count) return;
  uint32_t a[L], b[L];
  if (inf[e]) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) a[j] = b[j] = 0;
  } else {
    F<L>::from_mont(mkv(a, 1), mkvc(x + e, (int)N));
    F<L>::from_mont(mkv(b, 1), mkvc(y + e, (int)N));
  }
  limbs_to_be_bytes<L>(out + e * 2 * B, B, a);
  limbs_to_be_bytes<L>(out + e * 2 * B + B, B, b);
}

template <int L>
BGN_DEV void fp2_from_bytes_body(const uint8_t* in, int B, size_t count, uint32_t* re, uint32_t* im, size_t N,
                                 size_t e) {
  if (e >= count) return;
  uint32_t a[L], b[L];
  be_bytes_to_limbs<L>(a, in + e * 2 * B, B);
  be_bytes_to_limbs<L>(b, in + e * 2 * B + B, B);
  F<L>::to_mont(mkv(re + e, (int)N), mkv(a, 1));
  F<L>::to_mont(mkv(im + e, (int)N), mkv(b, 1));
}

template <int L>
BGN_DEV void fp2_to_bytes_body(const uint32_t* re, const uint32_t* im, size_t N, size_t count, uint8_t* out, int B,
                               size_t e) {
  if (e >= count) return;
  uint32_t a[L], b[L];
  F<L>::from_mont(mkv(a, 1), mkvc(re + e, (int)N));
  F<L>::from_mont(mkv(b, 1), mkvc(im + e, (int)N));
  limbs_to_be_bytes<L>(out + e * 2 * B, B, a);
  limbs_to_be_bytes<L>(out + e * 2 * B + B, B, b);
}

// ------------------------------------------------------------ G1 kernels
// Fixed-base window tables: entry (win, d) = d * 2^(8 win) * Base, affine
// Montgomery, AoS: tab[(win*255 + d-1) * 2L + {0..L-1: x, L..2L-1: y}].
// C = x*P + r*Q  (EncryptWithRandomness, bgn.go:340-353), x < 0: C = -(|x|*P + r*Q); Jacobian out.
template <int L>
BGN_DEV void encrypt_body(const EncArgs& a, size_t e) {
  if (e >= a.count) return;
  uint32_t X[L], Y[L], Z[L];
  V vX = mkv(X, 1), vY = mkv(Y, 1), vZ = mkv(Z, 1);
  BGN_UNROLL
  for (int j = 0; j < L; j++) X[j] = Y[j] = Z[j] = 0;
  if (a.r_be) {
    const uint8_t* r = a.r_be + e * a.rbytes;
    for (int win = 0; win < a.rbytes; win++) {
      uint32_t d = r[a.rbytes - 1 - win];
      if (d) {
        const uint32_t* ent = a.tabQ + ((size_t)win * 255 + (d - 1)) * 2 * L;
        G<L>::madd(vX, vY, vZ, mkvc(ent, 1), mkvc(ent + L, 1), false);
      }
    }
  }
  int64_t xs = a.x[e];
  bool neg = xs < 0;
  uint64_t xm = neg ? (uint64_t)(-(xs + 1)) + 1u : (uint64_t)xs;
  for (int win = 0; win < 8; win++) {
    uint32_t d = (uint32_t)(xm >> (8 * win)) & 255u;
    if (d) {
      const uint32_t* ent = a.tabP + ((size_t)win * 255 + (d - 1)) * 2 * L;
      G<L>::madd(vX, vY, vZ, mkvc(ent, 1), mkvc(ent + L, 1), false);
    }
  }
  // negative plaintext: -(|x|*P + r*Q), the Sub(encryptZero(), Encrypt(|c|)) of poly.go:17-21
  if (neg) F<L>::neg(vY, vY);
  st<L>(mkv(a.X + e, (int)a.N), X);
  st<L>(mkv(a.Y + e, (int)a.N), Y);
  st<L>(mkv(a.Z + e, (int)a.N), Z);
}

// Jacobian -> affine with one inversion per thread (Montgomery's trick over
// the elements g, g+G, g+2G, ... of thread g).  Output layout is handle-based
// so the same kernel fills SoA arrays and AoS tables.
template <int L>
BGN_DEV void normalize_body(const NormArgs& a, size_t g) {
  typedef Fp<L> P;
  if (g >= (size_t)a.G || g >= a.count) return;
  uint32_t acc[L], z[L], t[L];
  BGN_UNROLL
  for (int j = 0; j < L; j++) acc[j] = c_fc.one[j];
  for (size_t e = g; e < a.count; e += a.G) {
    ld<L>(z, mkvc(a.Z + e, (int)a.N));
    P::canon(t, z);
    if (P::is_zero_raw(t)) continue;
    st<L>(mkv(a.scratch + e, (int)a.N), acc);
    P::mul(t, acc, z);
    BGN_UNROLL
    for (int j = 0; j < L; j++) acc[j] = t[j];
  }
  F<L>::inv(mkv(acc, 1), mkv(acc, 1));
  size_t last = ((a.count - 1 - g) / a.G) * a.G + g;  // largest e = g (mod G) below count
  for (size_t e = last;; e -= a.G) {
    ld<L>(z, mkvc(a.Z + e, (int)a.N));
    P::canon(t, z);
    bool isinf = P::is_zero_raw(t);
    V ox = mkv(a.ox + e * a.o_estride, (int)a.o_lstride), oy = mkv(a.oy + e * a.o_estride, (int)a.o_lstride);
    if (a.inf) a.inf[e] = isinf ? 1 : 0;
    if (isinf) {
      F<L>::set_zero(ox);
      F<L>::set_zero(oy);
    } else {
      uint32_t zi[L], zz[L], u[L];
      ld<L>(t, mkvc(a.scratch + e, (int)a.N));
      P::mul(zi, acc, t);  // 1/Z_e
      P::mul(t, acc, z);   // drop Z_e from the running inverse
      BGN_UNROLL
      for (int j = 0; j < L; j++) acc[j] = t[j];
      P::sqr(zz, zi);
      ld<L>(u, mkvc(a.X + e, (int)a.N));
      P::mul(t, u, zz);
      st<L>(ox, t);
      P::mul(u, zz, zi);
      ld<L>(zz, mkvc(a.Y + e, (int)a.N));
      P::mul(t, zz, u);
      st<L>(oy, t);
    }
    if (e < (size_t)a.G) break;
  }
}

// EAdd / ESub on L1: affine + affine -> Jacobian (bgn.go:482, 419)
template <int L>
BGN_DEV void g1_add_body(const G1AddArgs& a, size_t e) {
  if (e >= a.count) return;
  size_t e1 = a.bcast1 ? 0 : e;
  uint32_t X[L], Y[L], Z[L];
  V vX = mkv(X, 1), vY = mkv(Y, 1), vZ = mkv(Z, 1);
  if (a.inf1[e1]) {
    BGN_UNROLL
    for (int j = 0; j < L; j++) X[j] = Y[j] = Z[j] = 0;
  } else {
    ld<L>(X, mkvc(a.x1 + e1, (int)a.N1));
    ld<L>(Y, mkvc(a.y1 + e1, (int)a.N1));
    BGN_UNROLL
    for (int j = 0; j < L; j++) Z[j] = c_fc.one[j];
  }
  if (!a.inf2[e]) G<L>::madd(vX, vY, vZ, mkvc(a.x2 + e, (int)a.N2), mkvc(a.y2 + e, (int)a.N2), a.subtract != 0);
  st<L>(mkv(a.X + e, (int)a.N), X);
  st<L>(mkv(a.Y + e, (int)a.N), Y);
  st<L>(mkv(a.Z + e, (int)a.N), Z);
}

// MultConst on L1: k*C, per-element big-endian scalar (bgn.go:258)
template <int L>
BGN_DEV void g1_mulvar_body(const G1MulArgs& a, size_t e) {
  if (e >= a.count) return;
  uint32_t X[L], Y[L], Z[L];
  V vX = mkv(X, 1), vY = mkv(Y, 1), vZ = mkv(Z, 1);
  BGN_UNROLL
  for (int j = 0; j < L; j++) X[j] = Y[j] = Z[j] = 0;
  if (!a.inf[e]) {
    V ax = mkvc(a.x + e, (int)a.Nin), ay = mkvc(a.y + e, (int)a.Nin);
    const uint8_t* k = a.k_be + e * a.kbytes;
    for (int i = 0; i < a.kbytes; i++) {
      uint32_t byte = k[i];
      for (int bit = 7; bit >= 0; bit--) {
        G<L>::dbl(vX, vY, vZ);
        if ((byte >> bit) & 1) G<L>::madd(vX, vY, vZ, ax, ay, false);
      }
    }
  }
  st<L>(mkv(a.X + e, (int)a.N), X);
  st<L>(mkv(a.Y + e, (int)a.N), Y);
  st<L>(mkv(a.Z + e, (int)a.N), Z);
}

// table construction -----------------------------------------------------
// bases[win] = 2^(8 win) * Base in Jacobian coordinates (single thread: 8*nwin doublings)
template <int L>
BGN_DEV void tab_bases_body(const uint32_t* bx, const uint32_t* by, int nwin, uint32_t* X, uint32_t* Y, uint32_t* Z,
                            size_t N, size_t g) {
  if (g != 0) return;
  uint32_t x[L], y[L], z[L];
  ld<L>(x, mkvc(bx, 1));
  ld<L>(y, mkvc(by, 1));
  BGN_UNROLL
  for (int j = 0; j < L; j++) z[j] = c_fc.one[j];
  V vX = mkv(x, 1), vY = mkv(y, 1), vZ = mkv(z, 1);
  for (int win = 0; win < nwin; win++) {
    st<L>(mkv(X + win, (int)N), x);
    st<L>(mkv(Y + win, (int)N), y);
    st<L>(mkv(Z + win, (int)N), z);
    for (int i = 0; i < 8; i++) G<L>::dbl(vX, vY, vZ);
  }
}
// entries (win, d), d = 1..255, by repeated addition of the affine base of the window
template <int L>
BGN_DEV void tab_fill_body(const uint32_t* ax, const uint32_t* ay, const uint8_t* ainf, size_t Nb, int nwin,
                           uint32_t* X, uint32_t* Y, uint32_t* Z, size_t N, size_t g) {
  int win = (int)g;
  if (g >= (size_t)nwin) return;
  uint32_t x[L], y[L], z[L];
  BGN_UNROLL
  for (int j = 0; j < L; j++) x[j] = y[j] = z[j] = 0;
  V vX = mkv(x, 1), vY = mkv(y, 1), vZ = mkv(z, 1);
  for (int d = 1; d <= 255; d++) {
    if (!ainf[win]) G<L>::madd(vX, vY, vZ, mkvc(ax + win, (int)Nb), mkvc(ay + win, (int)Nb), false);
    size_t o = (size_t)win * 255 + (d - 1);
    st<L>(mkv(X + o, (int)N), x);
    st<L>(mkv(Y + o, (int)N), y);
    st<L>(mkv(Z + o, (int)N), z);
  }
}

// ------------------------------------------------------------ GT kernels
template <int L>
BGN_DEV void gt_mul_body(const GtBinArgs& a, size_t e) {
  if (e >= a.count) return;
  uint32_t b0[L], b1[L];
  ld<L>(b0, mkvc(a.bre + e, (int)a.Nb));
  ld<L>(b1, mkvc(a.bim + e, (int)a.Nb));
  if (a.conj_b) F<L>::neg(mkv(b1, 1), mkv(b1, 1));
  F<L>::mul2(mkv2(mkv(a.ore + e, (int)a.N), mkv(a.oim + e, (int)a.N)),
             mkv2(mkvc(a.are + e, (int)a.Na), mkvc(a.aim + e, (int)a.Na)), mkv2(mkv(b0, 1), mkv(b1, 1)));
}

template <int L>
BGN_DEV void gt_pow_body(const GtPowArgs& a, size_t e) {
  if (e >= a.count) return;
  V2 in = mkv2(mkvc(a.re + e, (int)a.Nin), mkvc(a.im + e, (int)a.Nin));
  V2 out = mkv2(mkv(a.ore + e, (int)a.N), mkv(a.oim + e, (int)a.N));
  if (a.mode == 1) {
    GT<L>::pow_fixed(out, in);
  } else if (a.mode == 2) {
    F<L>::conj2(out, in);
  } else {
    GT<L>::pow_var(out, in, a.e_be + e * a.ebytes, a.ebytes);
    if (a.mode == 3) F<L>::neg(out.im, out.im);
  }
}

// One pass of the GT product tree of an L2 sum (bgn.go:460 folded over terms):
// in[t*ncoeff + c], t < nterms  ->  out[g*ncoeff + c] = prod_{t = g (mod G)} in[t][c]
template <int L>
BGN_DEV void gt_reduce_body(const uint32_t* re, const uint32_t* im, size_t Nin, size_t nterms, int ncoeff, int G,
                            uint32_t* ore, uint32_t* oim, size_t N, size_t id) {
  if (id >= (size_t)G * ncoeff) return;
  size_t g = id / ncoeff;
  int c = (int)(id % ncoeff);
  uint32_t r0[L], r1[L];
  V2 acc = mkv2(mkv(r0, 1), mkv(r1, 1));
  F<L>::set_one2(acc);
  for (size_t t = g; t < nterms; t += G) {
    size_t e = t * ncoeff + c;
    F<L>::mul2(acc, acc, mkv2(mkvc(re + e, (int)Nin), mkvc(im + e, (int)Nin)));
  }
  st<L>(mkv(ore + id, (int)N), r0);
  st<L>(mkv(oim + id, (int)N), r1);
}

// ------------------------------------------------------------ BSGS (gsbs.go)
// Baby steps: elems[j] = gen^(j+1), j < S, canonical Montgomery AoS [S][2L];
// open-addressing hash table slots[hmask+1] holding j+1 (0 = empty).
template <int L>
BGN_DEV uint32_t bsgs_hash(const uint32_t (&re)[L], const uint32_t (&im)[L]) {
  uint32_t h = re[0] * 0x9E3779B1u ^ im[0] * 0x85EBCA77u ^ (re[1] >> 7);
  return h ^ (h >> 15);
}

template <int L>
BGN_DEV void bsgs_build_body(const BsgsBuildArgs& a, size_t g) {
  typedef Fp<L> P;
  uint64_t j0 = (uint64_t)g * a.chunk;
  if (j0 >= a.S) return;
  uint32_t g0[L], g1[L], r0[L], r1[L];
  ld<L>(g0, mkvc(a.gen, 1));
  ld<L>(g1, mkvc(a.gen + L, 1));
  V2 gg = mkv2(mkv(g0, 1), mkv(g1, 1)), rr = mkv2(mkv(r0, 1), mkv(r1, 1));
  // rr = gen^(j0+1)
  uint64_t ex = j0 + 1;
  F<L>::set_one2(rr);
  for (int bit = 40; bit >= 0; bit--) {
    F<L>::sqr2(rr, rr);
    if ((ex >> bit) & 1) F<L>::mul2(rr, rr, gg);
  }
  for (int i = 0; i < a.chunk && j0 + i < a.S; i++) {
    uint32_t c0[L], c1[L];
    P::canon(c0, r0);
    P::canon(c1, r1);
    uint32_t j = (uint32_t)(j0 + i);
    uint32_t* dst = a.elems + (size_t)j * 2 * L;
    BGN_UNROLL
    for (int k = 0; k < L; k++) {
      dst[k] = c0[k];
      dst[L + k] = c1[k];
    }
    uint32_t h = bsgs_hash<L>(c0, c1) & a.hmask;
    while (atomicCAS(a.slots + h, 0u, j + 1) != 0u) h = (h + 1) & a.hmask;
    F<L>::mul2(rr, rr, gg);
  }
}

template <int L>
BGN_DEVNI int64_t bsgs_probe(const BsgsLookupArgs& a, const uint32_t (&r0)[L], const uint32_t (&r1)[L]) {
  typedef Fp<L> P;
  uint32_t c0[L], c1[L];
  P::canon(c0, r0);
  P::canon(c1, r1);
  uint32_t h = bsgs_hash<L>(c0, c1) & a.hmask;
  for (;;) {
    uint32_t s = a.slots[h];
    if (s == 0) return -1;
    const uint32_t* el = a.elems + (size_t)(s - 1) * 2 * L;
    uint32_t diff = 0;
    BGN_UNROLL
    for (int k = 0; k < L; k++) diff |= (el[k] ^ c0[k]) | (el[L + k] ^ c1[k]);
    if (diff == 0) return (int64_t)(s - 1);
    h = (h + 1) & a.hmask;
  }
}

template <int L>
BGN_DEV void bsgs_lookup_body(const BsgsLookupArgs& a, size_t e) {
  typedef Fp<L> P;
  if (e >= a.count) return;
  uint32_t p0[L], p1[L], n0[L], n1[L], gi0[L], gi1[L], t[L];
  ld<L>(p0, mkvc(a.re + e, (int)a.Nin));
  ld<L>(p1, mkvc(a.im + e, (int)a.Nin));
  // identity => 0 (recoverMessage, bgn.go:359-363)
  {
    uint32_t c0[L], c1[L], one[L];
    P::canon(c0, p0);
    P::canon(c1, p1);
    BGN_UNROLL
    for (int k = 0; k < L; k++) one[k] = c_fc.one[k];
    P::canon(t, one);
    if (P::eq_raw(c0, t) && P::is_zero_raw(c1)) {
      a.out[e] = 0;
      a.status[e] = 0;
      return;
    }
  }
  BGN_UNROLL
  for (int k = 0; k < L; k++) {
    n0[k] = p0[k];
    t[k] = 0;
  }
  P::sub(n1, t, p1);  // conj(csk) = csk^-1 (GT is unitary): the Neg(ct) retry of bgn.go:235-241
  ld<L>(gi0, mkvc(a.ginv, 1));
  ld<L>(gi1, mkvc(a.ginv + L, 1));
  V2 vp = mkv2(mkv(p0, 1), mkv(p1, 1)), vn = mkv2(mkv(n0, 1), mkv(n1, 1)), vg = mkv2(mkv(gi0, 1), mkv(gi1, 1));
  for (uint32_t i = 0; i < a.giant_steps; i++) {
    int64_t j = bsgs_probe<L>(a, p0, p1);
    if (j >= 0) {
      uint64_t m = (uint64_t)i * a.S + (uint64_t)j + 1;
      if (m <= a.mmax) {
        a.out[e] = (int64_t)m;
        a.status[e] = 0;
        return;
      }
    }
    j = bsgs_probe<L>(a, n0, n1);
    if (j >= 0) {
      uint64_t m = (uint64_t)i * a.S + (uint64_t)j + 1;
      if (m <= a.mmax) {
        a.out[e] = -(int64_t)m;
        a.status[e] = 0;
        return;
      }
    }
    if (i + 1 < a.giant_steps) {
      F<L>::mul2(vp, vp, vg);
      F<L>::mul2(vn, vn, vg);
    }
  }
  a.out[e] = 0;
  a.status[e] = 1;
}

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
BGN_KERNEL_1D(gt_mul, GtBinArgs)
BGN_KERNEL_1D(gt_pow, GtPowArgs)
BGN_KERNEL_1D(bsgs_build, BsgsBuildArgs)
BGN_KERNEL_1D(bsgs_lookup, BsgsLookupArgs)

template <int L>
__global__ void k_g1_from_bytes(const uint8_t* __restrict__ in, int B, size_t count, uint32_t* x, uint32_t* y,
                                uint8_t* inf, size_t N) {
  g1_from_bytes_body<L>(in, B, count, x, y, inf, N, BGN_GID(size_t));
}
template <int L>
__global__ void k_g1_to_bytes(const uint32_t* x, const uint32_t* y, const uint8_t* inf, size_t N, size_t count,
                              uint8_t* __restrict__ out, int B) {
  g1_to_bytes_body<L>(x, y, inf, N, count, out, B, BGN_GID(size_t));
}
template <int L>
__global__ void k_fp2_from_bytes(const uint8_t* __restrict__ in, int B, size_t count, uint32_t* re, uint32_t* im,
                                 size_t N) {
  fp2_from_bytes_body<L>(in, B, count, re, im, N, BGN_GID(size_t));
}
template <int L>
__global__ void k_fp2_to_bytes(const uint32_t* re, const uint32_t* im, size_t N, size_t count,
                               uint8_t* __restrict__ out, int B) {
  fp2_to_bytes_body<L>(re, im, N, count, out, B, BGN_GID(size_t));
}
template <int L>
__global__ void k_tab_bases(const uint32_t* bx, const uint32_t* by, int nwin, uint32_t* X, uint32_t* Y, uint32_t* Z,
                            size_t N) {
  tab_bases_body<L>(bx, by, nwin, X, Y, Z, N, BGN_GID(size_t));
}
template <int L>
__global__ void k_tab_fill(const uint32_t* ax, const uint32_t* ay, const uint8_t* ainf, size_t Nb, int nwin,
                           uint32_t* X, uint32_t* Y, uint32_t* Z, size_t N) {
  tab_fill_body<L>(ax, ay, ainf, Nb, nwin, X, Y, Z, N, BGN_GID(size_t));
}
template <int L>
__global__ void __launch_bounds__(128) k_gt_reduce(const uint32_t* re, const uint32_t* im, size_t Nin, size_t nterms,
                                                   int ncoeff, int G, uint32_t* ore, uint32_t* oim, size_t N) {
  gt_reduce_body<L>(re, im, Nin, nterms, ncoeff, G, ore, oim, N, BGN_GID(size_t));
}

// the Miller team kernel: <= 144 threads per block, 2 blocks per SM (shared-memory bound)
template <int L>
__global__ void __launch_bounds__(144, 2) k_miller(const __grid_constant__ MillerArgs a) {
  extern __shared__ uint32_t smem_dyn[];
  MillerTeam<L> T(a, smem_dyn, threadIdx.x, blockIdx.x, blockDim.x);
  T.run([] { __syncthreads(); });
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
  for (int c = 0; c < ILP; c++) ld<L>(a[c], mkvc(io + e, (int)N));
  ld<L>(b, mkvc(io + e, (int)N));
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
  st<L>(mkv(io + e, (int)N), a[0]);
}
#endif  // !BGN_HOSTSIM
