/* A caller of the C-ABI with no Python, no torch, no C++: encrypts two small polynomials, multiplies them
 * (MultPoly), sums the product with itself and decrypts -- the calls a cgo binding makes
 * (go/bgn_cuda.go).  The key and the expected bytes come from vectors.h, written by the test from
 * tests/golden/kb128.json.  Exit code 0 = every byte and every plaintext as expected. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "bgn_b200.h"
#include "vectors.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int st_ = (call);                                                            \
    if (st_ != 0) {                                                              \
      fprintf(stderr, "%s -> %d: %s\n", #call, st_, ctx ? bgn_last_error(ctx) : ""); \
      return 10;                                                                 \
    }                                                                            \
  } while (0)

int main(void) {
  bgn_ctx* ctx = 0;
  bgn_params prm;
  int limbs, cb, sb;
  size_t eb, i;
  uint8_t *prod, *sum, status[D1 + D2];
  int64_t vals[D1 + D2];
  prm.p_be = KEY_P;
  prm.p_len = sizeof(KEY_P);
  prm.n_be = KEY_N;
  prm.n_len = sizeof(KEY_N);
  prm.l = KEY_L;
  prm.P_bytes = KEY_GEN_P;
  prm.Q_bytes = KEY_GEN_Q;
  CHECK(bgn_ctx_create(&prm, 0, &ctx));
  CHECK(bgn_ctx_info(ctx, &limbs, &cb, &sb));
  eb = 2 * (size_t)cb;
  if (sizeof(C1) != D1 * eb || sizeof(C2) != D2 * eb || sizeof(EXPECT) != (D1 + D2) * eb) return 11;
  prod = (uint8_t*)malloc(2 * (D1 + D2) * eb);
  sum = (uint8_t*)malloc((D1 + D2) * eb);
  CHECK(bgn_multpoly_batch(ctx, C1, D1, C2, D2, 1, prod));
  if (memcmp(prod, EXPECT, sizeof(EXPECT)) != 0) {
    fprintf(stderr, "MultPoly bytes differ from the golden vector\n");
    return 12;
  }
  /* AddPoly of the product with itself: 2 terms of D1+D2 slots */
  memcpy(prod + (D1 + D2) * eb, prod, (D1 + D2) * eb);
  CHECK(bgn_l2_sum_reduce(ctx, prod, 2, D1 + D2, sum));
  CHECK(bgn_ctx_set_secret(ctx, KEY_Q1, sizeof(KEY_Q1), MSG_SPACE, 0));
  CHECK(bgn_decrypt_batch(ctx, sum, 1, D1 + D2, vals, status));
  for (i = 0; i < D1 + D2; i++) {
    if (status[i] != 0 || vals[i] != 2 * PLAIN[i]) {
      fprintf(stderr, "slot %u: status %u value %lld, expected %lld\n", (unsigned)i, status[i], (long long)vals[i],
              (long long)(2 * PLAIN[i]));
      return 13;
    }
  }
  printf("C caller: multpoly bytes == golden, decrypt(2 x product) == 2 x plaintext convolution (%d slots)\n", D1 + D2);
  free(prod);
  free(sum);
  bgn_ctx_destroy(ctx);
  return 0;
}
