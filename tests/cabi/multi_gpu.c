/* ONE process driving SEVERAL contexts, one host thread each -- the shape a Go integration has (one
 * goroutine per device around the cgo calls; the reference fans goroutines out in poly.go:129-153).
 * Each thread owns one bgn_ctx on its device, encrypts its contiguous shard of two vectors of
 * polynomials, multiplies them (MultPoly) and reduces its products to one partial L2 sum
 * (bgn_l2_sum_reduce).  The main thread folds the partials with the same call on context 0, decrypts,
 * and checks (i) the plaintext inner product and (ii) that the folded bytes equal what ONE context
 * computes over the whole batch.  No Python, no torch, no C++.
 *
 *   multi_gpu <devices, e.g. 0,1 or 0,0> <terms> [repeat]
 *
 * With `repeat` > 0 the shard work is repeated and the aggregate EMult/s printed (bench.py
 * --single-process reads it).  The key comes from vectors.h (written by the test / bench). */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "bgn_b200.h"
#include "vectors.h"

#define MAXDEV 16

typedef struct {
  int device, index, ndev, repeat;
  size_t terms, lo, hi; /* shard [lo, hi) of the terms */
  bgn_ctx* ctx;
  uint8_t* partial; /* (D1 + D2) elements */
  size_t eb, sb;
  double seconds;
  int rc;
  char err[256];
} worker_t;

static bgn_params key_params(void) {
  bgn_params prm;
  prm.p_be = KEY_P;
  prm.p_len = sizeof(KEY_P);
  prm.n_be = KEY_N;
  prm.n_len = sizeof(KEY_N);
  prm.l = KEY_L;
  prm.P_bytes = KEY_GEN_P;
  prm.Q_bytes = KEY_GEN_Q;
  return prm;
}

/* plaintext digits and randomness of term t, slot i of vector v: small deterministic functions */
static int64_t digit(size_t t, int i, int v) { return (int64_t)((t * 7 + (size_t)i * 3 + (size_t)v * 5) % 3) - 1; }
static void fill_r(uint8_t* r, size_t sb, size_t t, int i, int v) {
  size_t k;
  uint32_t s = (uint32_t)(t * 2654435761u + (uint32_t)i * 40503u + (uint32_t)v * 977u + 12345u);
  for (k = 0; k < sb; k++) {
    s = s * 1664525u + 1013904223u;
    r[k] = (uint8_t)(s >> 24);
  }
  r[0] &= 0x3f; /* below n */
}

#define WCHECK(call)                                                                         \
  do {                                                                                       \
    int st_ = (call);                                                                        \
    if (st_ != 0) {                                                                          \
      snprintf(w->err, sizeof(w->err), "%s -> %d: %s", #call, st_,                           \
               w->ctx ? bgn_last_error(w->ctx) : bgn_global_last_error());                   \
      w->rc = st_;                                                                           \
      goto done;                                                                             \
    }                                                                                        \
  } while (0)

static double now(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static void* work(void* arg) {
  worker_t* w = (worker_t*)arg;
  bgn_params prm = key_params();
  size_t cnt = w->hi - w->lo, t;
  int limbs, cb, sb, i, rep;
  int64_t *x1 = 0, *x2 = 0;
  uint8_t *r1 = 0, *r2 = 0, *c1 = 0, *c2 = 0, *prod = 0;
  double t0;
  w->rc = 0;
  WCHECK(bgn_ctx_create(&prm, w->device, &w->ctx));
  WCHECK(bgn_ctx_info(w->ctx, &limbs, &cb, &sb));
  w->eb = 2 * (size_t)cb;
  w->sb = (size_t)sb;
  w->partial = (uint8_t*)malloc((D1 + D2) * w->eb);
  if (cnt == 0) { /* an empty shard contributes the identity */
    WCHECK(bgn_l2_sum_reduce(w->ctx, 0, 0, D1 + D2, w->partial));
    goto done;
  }
  x1 = (int64_t*)malloc(cnt * D1 * sizeof(int64_t));
  x2 = (int64_t*)malloc(cnt * D2 * sizeof(int64_t));
  r1 = (uint8_t*)malloc(cnt * D1 * w->sb);
  r2 = (uint8_t*)malloc(cnt * D2 * w->sb);
  c1 = (uint8_t*)malloc(cnt * D1 * w->eb);
  c2 = (uint8_t*)malloc(cnt * D2 * w->eb);
  prod = (uint8_t*)malloc(cnt * (D1 + D2) * w->eb);
  for (t = 0; t < cnt; t++) {
    for (i = 0; i < D1; i++) {
      x1[t * D1 + (size_t)i] = digit(w->lo + t, i, 0);
      fill_r(r1 + (t * D1 + (size_t)i) * w->sb, w->sb, w->lo + t, i, 0);
    }
    for (i = 0; i < D2; i++) {
      x2[t * D2 + (size_t)i] = digit(w->lo + t, i, 1);
      fill_r(r2 + (t * D2 + (size_t)i) * w->sb, w->sb, w->lo + t, i, 1);
    }
  }
  WCHECK(bgn_encrypt_batch(w->ctx, x1, r1, cnt * D1, c1));
  WCHECK(bgn_encrypt_batch(w->ctx, x2, r2, cnt * D2, c2));
  t0 = now();
  for (rep = 0; rep < (w->repeat > 0 ? w->repeat : 1); rep++) {
    WCHECK(bgn_multpoly_batch(w->ctx, c1, D1, c2, D2, cnt, prod));
    WCHECK(bgn_l2_sum_reduce(w->ctx, prod, cnt, D1 + D2, w->partial));
  }
  w->seconds = now() - t0;
done:
  free(x1);
  free(x2);
  free(r1);
  free(r2);
  free(c1);
  free(c2);
  free(prod);
  return 0;
}

static int run(const int* devs, int ndev, size_t terms, int repeat, worker_t* ws) {
  pthread_t th[MAXDEV];
  int i;
  for (i = 0; i < ndev; i++) {
    size_t base = terms / (size_t)ndev, rem = terms % (size_t)ndev;
    memset(&ws[i], 0, sizeof(ws[i]));
    ws[i].device = devs[i];
    ws[i].index = i;
    ws[i].ndev = ndev;
    ws[i].repeat = repeat;
    ws[i].terms = terms;
    ws[i].lo = (size_t)i * base + ((size_t)i < rem ? (size_t)i : rem);
    ws[i].hi = ws[i].lo + base + ((size_t)i < rem ? 1 : 0);
    if (pthread_create(&th[i], 0, work, &ws[i]) != 0) return 20;
  }
  for (i = 0; i < ndev; i++) pthread_join(th[i], 0);
  for (i = 0; i < ndev; i++)
    if (ws[i].rc != 0) {
      fprintf(stderr, "worker %d (device %d): %s\n", i, ws[i].device, ws[i].err);
      return 21;
    }
  return 0;
}

int main(int argc, char** argv) {
  int devs[MAXDEV], ndev = 0, repeat = 0, i, k, rc;
  size_t terms, t, eb;
  worker_t ws[MAXDEV], one[1];
  uint8_t *parts, *total, status[D1 + D2];
  int64_t vals[D1 + D2], plain[D1 + D2];
  int zero = 0;
  double slowest = 0;
  char* tok;
  if (argc < 3) {
    fprintf(stderr, "usage: multi_gpu <devices> <terms> [repeat]\n");
    return 2;
  }
  for (tok = strtok(argv[1], ","); tok && ndev < MAXDEV; tok = strtok(0, ",")) devs[ndev++] = atoi(tok);
  terms = (size_t)strtoull(argv[2], 0, 10);
  if (argc > 3) repeat = atoi(argv[3]);
  if (ndev < 1) return 2;
  rc = run(devs, ndev, terms, repeat, ws);
  if (rc) return rc;
  eb = ws[0].eb;
  /* fold the per-context partials on context 0: the same reduction, nterms = ndev */
  parts = (uint8_t*)malloc((size_t)ndev * (D1 + D2) * eb);
  total = (uint8_t*)malloc((D1 + D2) * eb);
  for (i = 0; i < ndev; i++) memcpy(parts + (size_t)i * (D1 + D2) * eb, ws[i].partial, (D1 + D2) * eb);
  if (bgn_l2_sum_reduce(ws[0].ctx, parts, (size_t)ndev, D1 + D2, total) != 0) {
    fprintf(stderr, "fold: %s\n", bgn_last_error(ws[0].ctx));
    return 22;
  }
  if (bgn_ctx_set_secret(ws[0].ctx, KEY_Q1, sizeof(KEY_Q1), MSG_SPACE, 0) != 0 ||
      bgn_decrypt_batch(ws[0].ctx, total, 1, D1 + D2, vals, status) != 0) {
    fprintf(stderr, "decrypt: %s\n", bgn_last_error(ws[0].ctx));
    return 23;
  }
  memset(plain, 0, sizeof(plain));
  for (t = 0; t < terms; t++)
    for (i = 0; i < D1; i++)
      for (k = 0; k < D2; k++) plain[i + k] += digit(t, i, 0) * digit(t, k, 1);
  for (i = 0; i < D1 + D2; i++)
    if (status[i] != 0 || vals[i] != plain[i]) {
      fprintf(stderr, "slot %d: status %u value %lld, expected %lld\n", i, status[i], (long long)vals[i], (long long)plain[i]);
      return 24;
    }
  if (repeat == 0) { /* the whole batch on ONE context must give the same bytes */
    rc = run(&zero, 1, terms, 0, one);
    if (rc) return rc;
    if (memcmp(one[0].partial, total, (D1 + D2) * eb) != 0) {
      fprintf(stderr, "folded partials differ from the single-context result\n");
      return 25;
    }
    free(one[0].partial);
    bgn_ctx_destroy(one[0].ctx);
  }
  for (i = 0; i < ndev; i++)
    if (ws[i].seconds > slowest) slowest = ws[i].seconds;
  printf("multi_gpu: %d contexts (devices", ndev);
  for (i = 0; i < ndev; i++) printf(" %d", devs[i]);
  printf("), %lu terms: decrypt(fold(partials)) == plaintext inner product%s\n", (unsigned long)terms,
         repeat == 0 ? ", bytes == single context" : "");
  if (repeat > 0 && slowest > 0)
    printf("{\"contexts\": %d, \"terms\": %lu, \"repeat\": %d, \"seconds_slowest_thread\": %.6f, \"emult_per_s\": %.3f}\n", ndev,
           (unsigned long)terms, repeat, slowest, (double)terms * repeat / slowest);
  for (i = 0; i < ndev; i++) {
    free(ws[i].partial);
    bgn_ctx_destroy(ws[i].ctx);
  }
  free(parts);
  free(total);
  return 0;
}
