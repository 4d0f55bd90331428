/* cpu_ref.c -- multi-threaded C restatement of the BGN hot path on the CPU.
 *
 * TEST / BASELINE INFRASTRUCTURE ONLY (same rules as oracle/bgn_oracle.py): used by
 * tests/ as a fast checker and by bench.py's cpu_baseline / --impl reference legs as
 * the "port" of the reference's CPU path.  libbgn_b200.so never links or calls it.
 *
 * PARITY UNPINNED: the reference's arithmetic is github.com/Nik-U/pbc@3e516ca0c5d6 ->
 * libpbc 0.5.14 -> GMP (go.mod:5, README.md:13-45), none of which is in /root/reference or
 * in this image.  This file restates the same *definitions* (type-A1 curve y^2 = x^3 + x,
 * F_p^2 = F_p[i], reduced Tate pairing with distortion map (x,y) -> (-x, i y), final
 * exponent (p^2-1)/n, PBC element_to_bytes layout) with the *algorithmic choices* libpbc
 * makes on a CPU: word-size (64-bit limb) Montgomery F_p, one Miller loop + one final
 * exponentiation per pairing in Jacobian coordinates over all bits of n (no sharing between
 * pairings), square-and-multiply exponentiation.  It is checked byte-for-byte against
 * bgn_oracle.py and the golden vectors (tests/test_cpu_ref.py).
 *
 * Scheme-level call sites it mirrors: Mult = Pair (bgn.go:294-314), MultPoly's d1*d2 pairings
 * and GT accumulation (poly.go:123-156), EncryptWithRandomness (bgn.go:340-353), C^q1
 * (bgn.go:223).
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;
#define MAXW 18

typedef struct {
  int w;        /* 64-bit limbs of p */
  int B;        /* serialised bytes per coordinate */
  u64 p[MAXW], one[MAXW], r2[MAXW];
  u64 np0;
  u64 n[MAXW];
  int nbits;
  u64 l;
} cpuref;

/* ---------------------------------------------------------------- F_p (Montgomery, canonical range) */
static int geq(const u64* a, const u64* b, int w) {
  for (int i = w - 1; i >= 0; i--)
    if (a[i] != b[i]) return a[i] > b[i];
  return 1;
}
static u64 sub_n(u64* r, const u64* a, const u64* b, int w) {
  u64 br = 0;
  for (int i = 0; i < w; i++) {
    u128 t = (u128)a[i] - b[i] - br;
    r[i] = (u64)t;
    br = (u64)(t >> 64) & 1;
  }
  return br;
}
static u64 add_n(u64* r, const u64* a, const u64* b, int w) {
  u64 c = 0;
  for (int i = 0; i < w; i++) {
    u128 t = (u128)a[i] + b[i] + c;
    r[i] = (u64)t;
    c = (u64)(t >> 64);
  }
  return c;
}
static void fp_add(const cpuref* F, u64* r, const u64* a, const u64* b) {
  u64 c = add_n(r, a, b, F->w);
  if (c || geq(r, F->p, F->w)) sub_n(r, r, F->p, F->w);
}
static void fp_sub(const cpuref* F, u64* r, const u64* a, const u64* b) {
  if (sub_n(r, a, b, F->w)) add_n(r, r, F->p, F->w);
}
static void fp_neg(const cpuref* F, u64* r, const u64* a) {
  u64 z[MAXW] = {0};
  fp_sub(F, r, z, a);
}
static void fp_mul(const cpuref* F, u64* r, const u64* a, const u64* b) {
  const int W = F->w;
  u64 t[MAXW + 2];
  memset(t, 0, sizeof(u64) * (W + 2));
  for (int i = 0; i < W; i++) {
    u128 c = 0;
    for (int j = 0; j < W; j++) {
      c += (u128)a[j] * b[i] + t[j];
      t[j] = (u64)c;
      c >>= 64;
    }
    c += t[W];
    t[W] = (u64)c;
    t[W + 1] = (u64)(c >> 64);
    u64 m = t[0] * F->np0;
    c = (u128)m * F->p[0] + t[0];
    c >>= 64;
    for (int j = 1; j < W; j++) {
      c += (u128)m * F->p[j] + t[j];
      t[j - 1] = (u64)c;
      c >>= 64;
    }
    c += t[W];
    t[W - 1] = (u64)c;
    t[W] = t[W + 1] + (u64)(c >> 64);
  }
  if (t[W] || geq(t, F->p, W)) sub_n(t, t, F->p, W);
  memcpy(r, t, sizeof(u64) * W);
}
static void fp_sqr(const cpuref* F, u64* r, const u64* a) { fp_mul(F, r, a, a); }
static int fp_is_zero(const cpuref* F, const u64* a) {
  u64 o = 0;
  for (int i = 0; i < F->w; i++) o |= a[i];
  return o == 0;
}
static int fp_eq(const cpuref* F, const u64* a, const u64* b) { return memcmp(a, b, sizeof(u64) * F->w) == 0; }
static void fp_set(const cpuref* F, u64* r, const u64* a) { memcpy(r, a, sizeof(u64) * F->w); }
/* r = a^e, e given as little-endian words, plain square-and-multiply (MSB first) */
static void fp_pow(const cpuref* F, u64* r, const u64* a, const u64* e, int ew) {
  u64 acc[MAXW], base[MAXW];
  fp_set(F, acc, F->one);
  fp_set(F, base, a);
  int started = 0;
  for (int i = ew * 64 - 1; i >= 0; i--) {
    int bit = (int)((e[i >> 6] >> (i & 63)) & 1);
    if (started) fp_sqr(F, acc, acc);
    if (bit) {
      fp_mul(F, acc, acc, base);
      started = 1;
    }
  }
  fp_set(F, r, acc);
}
static void fp_inv(const cpuref* F, u64* r, const u64* a) {
  u64 e[MAXW], two[MAXW] = {2};
  sub_n(e, F->p, two, F->w);
  fp_pow(F, r, a, e, F->w);
}
static void fp_from_be(const cpuref* F, u64* r, const uint8_t* b) {
  u64 t[MAXW] = {0};
  for (int i = 0; i < F->B; i++) {
    int pos = F->B - 1 - i; /* significance of byte b[i] */
    t[pos >> 3] |= (u64)b[i] << (8 * (pos & 7));
  }
  while (geq(t, F->p, F->w)) sub_n(t, t, F->p, F->w); /* from_bytes reduces mod p */
  fp_mul(F, r, t, F->r2);
}
static void fp_to_be(const cpuref* F, uint8_t* b, const u64* a) {
  u64 t[MAXW], o[MAXW] = {1};
  fp_mul(F, t, a, o);
  for (int i = 0; i < F->B; i++) {
    int pos = F->B - 1 - i;
    b[i] = (uint8_t)(t[pos >> 3] >> (8 * (pos & 7)));
  }
}

/* ---------------------------------------------------------------- F_p^2 */
typedef struct {
  u64 re[MAXW], im[MAXW];
} fp2;
static void fp2_mul(const cpuref* F, fp2* r, const fp2* a, const fp2* b) {
  u64 t0[MAXW], t1[MAXW], s[MAXW], u[MAXW], t2[MAXW];
  fp_mul(F, t0, a->re, b->re);
  fp_mul(F, t1, a->im, b->im);
  fp_add(F, s, a->re, a->im);
  fp_add(F, u, b->re, b->im);
  fp_mul(F, t2, s, u);
  fp_sub(F, r->re, t0, t1);
  fp_sub(F, t2, t2, t0);
  fp_sub(F, r->im, t2, t1);
}
static void fp2_sqr(const cpuref* F, fp2* r, const fp2* a) {
  u64 s[MAXW], d[MAXW], t[MAXW];
  fp_add(F, s, a->re, a->im);
  fp_sub(F, d, a->re, a->im);
  fp_mul(F, t, a->re, a->im);
  fp_mul(F, r->re, s, d);
  fp_add(F, r->im, t, t);
}
static void fp2_one(const cpuref* F, fp2* r) {
  fp_set(F, r->re, F->one);
  memset(r->im, 0, sizeof(r->im));
}
static void fp2_pow(const cpuref* F, fp2* r, const fp2* a, const u64* e, int ew) {
  fp2 acc, base = *a;
  fp2_one(F, &acc);
  int started = 0;
  for (int i = ew * 64 - 1; i >= 0; i--) {
    int bit = (int)((e[i >> 6] >> (i & 63)) & 1);
    if (started) fp2_sqr(F, &acc, &acc);
    if (bit) {
      fp2_mul(F, &acc, &acc, &base);
      started = 1;
    }
  }
  *r = acc;
}

/* ---------------------------------------------------------------- G1 (affine in/out, inf flag) */
typedef struct {
  u64 x[MAXW], y[MAXW];
  int inf;
} g1;
static int g1_on_curve(const cpuref* F, const g1* P) {
  u64 t[MAXW], u[MAXW];
  fp_sqr(F, t, P->x);
  fp_mul(F, u, t, P->x);
  fp_add(F, u, u, P->x);
  fp_sqr(F, t, P->y);
  return fp_eq(F, t, u);
}
static void g1_from_bytes(const cpuref* F, g1* P, const uint8_t* b) {
  int allz = 1;
  for (int i = 0; i < 2 * F->B; i++)
    if (b[i]) allz = 0;
  fp_from_be(F, P->x, b);
  fp_from_be(F, P->y, b + F->B);
  P->inf = allz || !g1_on_curve(F, P);
}
static void g1_to_bytes(const cpuref* F, uint8_t* b, const g1* P) {
  if (P->inf) {
    memset(b, 0, 2 * F->B);
    return;
  }
  fp_to_be(F, b, P->x);
  fp_to_be(F, b + F->B, P->y);
}
/* affine doubling / addition with one inversion each (libpbc curve.c behaviour) */
static void g1_dbl(const cpuref* F, g1* R, const g1* P) {
  if (P->inf || fp_is_zero(F, P->y)) {
    R->inf = 1;
    return;
  }
  u64 lam[MAXW], t[MAXW], u[MAXW], x3[MAXW];
  fp_sqr(F, t, P->x);
  fp_add(F, u, t, t);
  fp_add(F, t, u, t);
  fp_add(F, t, t, F->one); /* 3x^2 + 1 */
  fp_add(F, u, P->y, P->y);
  fp_inv(F, u, u);
  fp_mul(F, lam, t, u);
  fp_sqr(F, x3, lam);
  fp_sub(F, x3, x3, P->x);
  fp_sub(F, x3, x3, P->x);
  fp_sub(F, t, P->x, x3);
  fp_mul(F, t, lam, t);
  fp_sub(F, R->y, t, P->y);
  fp_set(F, R->x, x3);
  R->inf = 0;
}
static void g1_add(const cpuref* F, g1* R, const g1* P, const g1* Q) {
  if (P->inf) {
    *R = *Q;
    return;
  }
  if (Q->inf) {
    *R = *P;
    return;
  }
  if (fp_eq(F, P->x, Q->x)) {
    if (fp_eq(F, P->y, Q->y)) {
      g1_dbl(F, R, P);
    } else {
      R->inf = 1;
    }
    return;
  }
  u64 lam[MAXW], t[MAXW], u[MAXW], x3[MAXW];
  fp_sub(F, t, Q->y, P->y);
  fp_sub(F, u, Q->x, P->x);
  fp_inv(F, u, u);
  fp_mul(F, lam, t, u);
  fp_sqr(F, x3, lam);
  fp_sub(F, x3, x3, P->x);
  fp_sub(F, x3, x3, Q->x);
  fp_sub(F, t, P->x, x3);
  fp_mul(F, t, lam, t);
  fp_sub(F, u, t, P->y);
  fp_set(F, R->y, u);
  fp_set(F, R->x, x3);
  R->inf = 0;
}
static void g1_mul_be(const cpuref* F, g1* R, const g1* P, const uint8_t* k_be, int kbytes) {
  g1 acc;
  acc.inf = 1;
  for (int i = 0; i < kbytes; i++)
    for (int bit = 7; bit >= 0; bit--) {
      g1_dbl(F, &acc, &acc);
      if ((k_be[i] >> bit) & 1) g1_add(F, &acc, &acc, P);
    }
  *R = acc;
}

/* ---------------------------------------------------------------- reduced Tate pairing (type A1) */
/* f <- f * ((cR + aR*xB) + (bI*yB) i) */
static void line_mul(const cpuref* F, fp2* f, const u64* cR, const u64* aR, const u64* bI, const g1* Bp) {
  fp2 ln;
  u64 t[MAXW];
  fp_mul(F, t, aR, Bp->x);
  fp_add(F, ln.re, cR, t);
  fp_mul(F, ln.im, bI, Bp->y);
  fp2_mul(F, f, f, &ln);
}
static void pairing(const cpuref* F, fp2* out, const g1* A, const g1* Bp) {
  if (A->inf || Bp->inf) {
    fp2_one(F, out);
    return;
  }
  fp2 f;
  fp2_one(F, &f);
  u64 X[MAXW], Y[MAXW], Z[MAXW];
  fp_set(F, X, A->x);
  fp_set(F, Y, A->y);
  fp_set(F, Z, F->one);
  for (int i = F->nbits - 2; i >= 0; i--) {
    /* tangent at V and V <- 2V (Jacobian, a = 1) */
    u64 xx[MAXW], yy[MAXW], zz[MAXW], m[MAXW], s[MAXW], t[MAXW], u[MAXW], cR[MAXW], aR[MAXW], bI[MAXW];
    fp_sqr(F, xx, X);
    fp_sqr(F, yy, Y);
    fp_sqr(F, zz, Z);
    fp_sqr(F, t, zz);
    fp_add(F, m, xx, xx);
    fp_add(F, m, m, xx);
    fp_add(F, m, m, t);
    fp_mul(F, t, X, yy);
    fp_add(F, s, t, t);
    fp_add(F, s, s, s);
    fp_mul(F, t, Y, Z);
    fp_add(F, u, t, t); /* Z3 */
    fp_mul(F, bI, u, zz);
    fp_mul(F, aR, m, zz);
    fp_mul(F, t, m, X);
    fp_sub(F, t, t, yy);
    fp_sub(F, cR, t, yy);
    fp2_sqr(F, &f, &f);
    line_mul(F, &f, cR, aR, bI, Bp);
    fp_set(F, Z, u);
    fp_sqr(F, t, m);
    fp_sub(F, t, t, s);
    fp_sub(F, X, t, s);
    fp_sub(F, s, s, X);
    fp_mul(F, u, m, s);
    fp_sqr(F, t, yy);
    fp_add(F, t, t, t);
    fp_add(F, t, t, t);
    fp_add(F, t, t, t);
    fp_sub(F, Y, u, t);
    if (((F->n[i >> 6] >> (i & 63)) & 1) && i != 0) {
      /* chord through V and A, V <- V + A (mixed); the last addition (i == 0) lands on O with a
         vertical line whose value lies in F_p and is killed by the final exponentiation */
      u64 h[MAXW], r[MAXW], i4[MAXW], j[MAXW], vv[MAXW], z3[MAXW];
      fp_sqr(F, zz, Z);
      fp_mul(F, t, A->x, zz);
      fp_sub(F, h, t, X);
      fp_mul(F, t, Z, zz);
      fp_mul(F, u, A->y, t);
      fp_sub(F, r, u, Y);
      fp_add(F, r, r, r);
      fp_mul(F, t, Z, h);
      fp_add(F, z3, t, t);
      fp_mul(F, u, A->y, z3);
      fp_mul(F, t, r, A->x);
      fp_sub(F, cR, t, u);
      line_mul(F, &f, cR, r, z3, Bp);
      fp_sqr(F, t, h);
      fp_add(F, t, t, t);
      fp_add(F, i4, t, t);
      fp_mul(F, j, h, i4);
      fp_mul(F, vv, X, i4);
      fp_sqr(F, t, r);
      fp_sub(F, t, t, j);
      fp_sub(F, t, t, vv);
      fp_sub(F, t, t, vv);
      fp_sub(F, u, vv, t);
      fp_set(F, X, t);
      fp_mul(F, t, r, u);
      fp_mul(F, u, Y, j);
      fp_add(F, u, u, u);
      fp_sub(F, Y, t, u);
      fp_set(F, Z, z3);
    }
  }
  /* f^((p^2-1)/n) = (conj(f)/f)^l */
  fp2 g, inv;
  u64 nrm[MAXW], t[MAXW];
  fp_sqr(F, nrm, f.re);
  fp_sqr(F, t, f.im);
  fp_add(F, nrm, nrm, t);
  fp_inv(F, nrm, nrm);
  fp_mul(F, inv.re, f.re, nrm);
  fp_mul(F, t, f.im, nrm);
  fp_neg(F, inv.im, t);
  g = f;
  fp_neg(F, g.im, f.im);
  fp2_mul(F, &g, &g, &inv);
  u64 e[1] = {F->l};
  fp2_pow(F, out, &g, e, 1);
}
static void gt_to_bytes(const cpuref* F, uint8_t* b, const fp2* a) {
  fp_to_be(F, b, a->re);
  fp_to_be(F, b + F->B, a->im);
}
static void gt_from_bytes(const cpuref* F, fp2* a, const uint8_t* b) {
  fp_from_be(F, a->re, b);
  fp_from_be(F, a->im, b + F->B);
}

/* ---------------------------------------------------------------- batch drivers (pthreads) */
typedef struct {
  const cpuref* F;
  int op;
  size_t lo, hi;
  const uint8_t *a, *b;
  const int64_t* x;
  const uint8_t* kbe;
  int kbytes, d1, d2;
  const g1 *P, *Q;
  uint8_t* out;
} job;

static void* worker(void* arg) {
  job* J = (job*)arg;
  const cpuref* F = J->F;
  const size_t E = 2 * (size_t)F->B;
  for (size_t i = J->lo; i < J->hi; i++) {
    switch (J->op) {
      case 0: { /* Pair */
        g1 A, Bp;
        fp2 e;
        g1_from_bytes(F, &A, J->a + i * E);
        g1_from_bytes(F, &Bp, J->b + i * E);
        pairing(F, &e, &A, &Bp);
        gt_to_bytes(F, J->out + i * E, &e);
      } break;
      case 1: { /* MultPoly: d1*d2 full pairings, products accumulated per slot (poly.go:140-152) */
        int d1 = J->d1, d2 = J->d2, ns = d1 + d2;
        fp2* acc = (fp2*)malloc(sizeof(fp2) * ns);
        g1* c2 = (g1*)malloc(sizeof(g1) * d2);
        for (int s = 0; s < ns; s++) fp2_one(F, &acc[s]);
        for (int k = 0; k < d2; k++) g1_from_bytes(F, &c2[k], J->b + (i * d2 + k) * E);
        for (int u = 0; u < d1; u++) {
          g1 A;
          g1_from_bytes(F, &A, J->a + (i * d1 + u) * E);
          for (int k = 0; k < d2; k++) {
            fp2 e;
            pairing(F, &e, &A, &c2[k]);
            fp2_mul(F, &acc[u + k], &acc[u + k], &e);
          }
        }
        for (int s = 0; s < ns; s++) gt_to_bytes(F, J->out + (i * ns + s) * E, &acc[s]);
        free(acc);
        free(c2);
      } break;
      case 2: { /* EncryptWithRandomness: |x| P + r Q, negated for x < 0 (poly.go:17-21) */
        int64_t xs = J->x[i];
        u64 xm = xs < 0 ? (u64)(-(xs + 1)) + 1u : (u64)xs;
        uint8_t xb[8];
        for (int k = 0; k < 8; k++) xb[k] = (uint8_t)(xm >> (8 * (7 - k)));
        g1 G, H, C;
        g1_mul_be(F, &G, J->P, xb, 8);
        if (J->kbe) {
          g1_mul_be(F, &H, J->Q, J->kbe + i * J->kbytes, J->kbytes);
          g1_add(F, &C, &G, &H);
        } else {
          C = G;
        }
        if (xs < 0 && !C.inf) fp_neg(F, C.y, C.y);
        g1_to_bytes(F, J->out + i * E, &C);
      } break;
      case 3: { /* GT power with a shared exponent (C^q1, bgn.go:223) */
        fp2 c, r;
        gt_from_bytes(F, &c, J->a + i * E);
        u64 e[MAXW * 2] = {0};
        for (int k = 0; k < J->kbytes; k++) {
          int pos = J->kbytes - 1 - k;
          e[pos >> 3] |= (u64)J->kbe[k] << (8 * (pos & 7));
        }
        fp2_pow(F, &r, &c, e, (J->kbytes + 7) / 8);
        gt_to_bytes(F, J->out + i * E, &r);
      } break;
    }
  }
  return 0;
}

static void run(job* proto, size_t count, int threads) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > count) threads = (int)(count ? count : 1);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
  job* jobs = (job*)malloc(sizeof(job) * threads);
  for (int t = 0; t < threads; t++) {
    jobs[t] = *proto;
    jobs[t].lo = count * t / threads;
    jobs[t].hi = count * (t + 1) / threads;
    pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], 0);
  free(th);
  free(jobs);
}

/* ---------------------------------------------------------------- exported API */
static void words_from_be(u64* r, int w, const uint8_t* b, size_t len) {
  memset(r, 0, sizeof(u64) * w);
  for (size_t i = 0; i < len; i++) {
    size_t pos = len - 1 - i;
    if ((int)(pos >> 3) < w) r[pos >> 3] |= (u64)b[i] << (8 * (pos & 7));
  }
}
static int bits_of(const u64* a, int w) {
  for (int i = w - 1; i >= 0; i--)
    if (a[i]) return 64 * i + 64 - __builtin_clzll(a[i]);
  return 0;
}

cpuref* cpuref_create(const uint8_t* p_be, size_t plen, const uint8_t* n_be, size_t nlen, uint64_t l) {
  cpuref* F = (cpuref*)calloc(1, sizeof(cpuref));
  words_from_be(F->p, MAXW, p_be, plen);
  int pb = bits_of(F->p, MAXW);
  F->w = (pb + 63) / 64;
  F->B = (pb + 7) / 8;
  if (F->w > MAXW - 1 || pb < 40) {
    free(F);
    return 0;
  }
  words_from_be(F->n, MAXW, n_be, nlen);
  F->nbits = bits_of(F->n, MAXW);
  F->l = l;
  u64 inv = F->p[0];
  for (int i = 0; i < 6; i++) inv *= 2 - F->p[0] * inv;
  F->np0 = (u64)0 - inv;
  /* R mod p and R^2 mod p by repeated doubling */
  u64 x[MAXW] = {1};
  for (int i = 0; i < 2 * 64 * F->w; i++) {
    u64 c = add_n(x, x, x, F->w);
    if (c || geq(x, F->p, F->w)) sub_n(x, x, F->p, F->w);
    if (i == 64 * F->w - 1) memcpy(F->one, x, sizeof(u64) * F->w);
  }
  memcpy(F->r2, x, sizeof(u64) * F->w);
  return F;
}
void cpuref_destroy(cpuref* F) { free(F); }
int cpuref_coord_bytes(const cpuref* F) { return F->B; }

void cpuref_pair_batch(const cpuref* F, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out, int threads) {
  job J = {0};
  J.F = F;
  J.op = 0;
  J.a = a;
  J.b = b;
  J.out = out;
  run(&J, count, threads);
}
void cpuref_multpoly_batch(const cpuref* F, const uint8_t* c1, int d1, const uint8_t* c2, int d2, size_t count,
                           uint8_t* out, int threads) {
  job J = {0};
  J.F = F;
  J.op = 1;
  J.a = c1;
  J.b = c2;
  J.d1 = d1;
  J.d2 = d2;
  J.out = out;
  run(&J, count, threads);
}
void cpuref_encrypt_batch(const cpuref* F, const uint8_t* P_bytes, const uint8_t* Q_bytes, const int64_t* x,
                          const uint8_t* r_be, int rbytes, size_t count, uint8_t* out, int threads) {
  g1 P, Q;
  g1_from_bytes(F, &P, P_bytes);
  g1_from_bytes(F, &Q, Q_bytes);
  job J = {0};
  J.F = F;
  J.op = 2;
  J.x = x;
  J.kbe = r_be;
  J.kbytes = rbytes;
  J.P = &P;
  J.Q = &Q;
  J.out = out;
  run(&J, count, threads);
}
void cpuref_gt_pow_batch(const cpuref* F, const uint8_t* a, const uint8_t* e_be, int ebytes, size_t count,
                         uint8_t* out, int threads) {
  job J = {0};
  J.F = F;
  J.op = 3;
  J.a = a;
  J.kbe = e_be;
  J.kbytes = ebytes;
  J.out = out;
  run(&J, count, threads);
}
