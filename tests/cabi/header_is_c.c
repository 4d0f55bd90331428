/* The drop-in boundary is plain C: this file is compiled as C11 with -Wall -Werror -pedantic, takes the
 * address of every entry point include/bgn_b200.h declares and links against libbgn_b200.so.  It makes
 * no compute call (it runs in the CPU test-suite, where there is no GPU). */
#include <stdio.h>

#include "bgn_b200.h"

int main(void) {
  typedef void (*fn)(void);
  fn syms[] = {
      (fn)bgn_ctx_create,          (fn)bgn_ctx_destroy,        (fn)bgn_last_error,
      (fn)bgn_ctx_set_option,      (fn)bgn_ctx_info,           (fn)bgn_ctx_set_secret,
      (fn)bgn_encrypt_batch,       (fn)bgn_g1_add_batch,       (fn)bgn_g1_sub_batch,
      (fn)bgn_g1_neg_batch,        (fn)bgn_g1_mulconst_batch,  (fn)bgn_gt_mul_batch,
      (fn)bgn_gt_div_batch,        (fn)bgn_gt_inv_batch,       (fn)bgn_gt_pow_batch,
      (fn)bgn_pair_batch,          (fn)bgn_make_l2_batch,      (fn)bgn_multpoly_batch,
      (fn)bgn_l2_sum_reduce,       (fn)bgn_gt_pow_secret_batch, (fn)bgn_decrypt_batch,
      (fn)bgn_g1_blind_batch,      (fn)bgn_gt_blind_batch,     (fn)bgn_multconstpoly_batch,
      (fn)bgn_evalpoly_batch,      (fn)bgn_make_poly_l2_batch, (fn)bgn_timing_enable,
      (fn)bgn_timing_reset,        (fn)bgn_timing_get,         (fn)bgn_timing_last_call,
      (fn)bgn_bench_mulmod,        (fn)bgn_bench_imad_peak,    (fn)bgn_global_last_error,
      (fn)bgn_bench_issue_mix,     (fn)bgn_buf_import,         (fn)bgn_buf_export,
      (fn)bgn_buf_info,            (fn)bgn_buf_free,           (fn)bgn_encrypt_h,
      (fn)bgn_g1_add_h,            (fn)bgn_gt_mul_h,           (fn)bgn_pair_h,
      (fn)bgn_multpoly_h,          (fn)bgn_l2_sum_reduce_h,    (fn)bgn_decrypt_h,
      (fn)bgn_group_create,        (fn)bgn_group_destroy,      (fn)bgn_group_size,
      (fn)bgn_group_ctx,           (fn)bgn_group_last_error,   (fn)bgn_group_set_secret,
      (fn)bgn_group_set_option,    (fn)bgn_group_encrypt_batch, (fn)bgn_group_g1_add_batch,
      (fn)bgn_group_multpoly_batch, (fn)bgn_group_decrypt_batch, (fn)bgn_group_inner_product};
  unsigned n = (unsigned)(sizeof(syms) / sizeof(syms[0])), ok = 0, i;
  for (i = 0; i < n; i++) ok += syms[i] != 0;
  /* a null context is rejected with a status, never a crash */
  if (bgn_ctx_info(0, 0, 0, 0) != BGN_E_BADARG) return 2;
  if (bgn_ctx_set_option(0, "enc_window", 16) != BGN_E_BADARG) return 3;
  printf("%u/%u symbols\n", ok, n);
  return ok == n ? 0 : 1;
}
