/* bgn_b200.h -- C-ABI of the B200-native batched BGN engine.
 *
 * This is the drop-in boundary for the arithmetic the reference (sachaservan/bgn)
 * obtains, one element at a time, from github.com/Nik-U/pbc (cgo -> libpbc -> GMP).
 * Each entry point names the reference code it replaces (file:line in the
 * reference tree).  Everything crossing the boundary is a plain pointer + size:
 *
 *   - group elements travel in PBC element_to_bytes format: G1 = x||y, GT = re||im,
 *     each coordinate big-endian, fixed width B = ceil(bits(p)/8) bytes
 *     (bgn_ctx_info reports B).  The point at infinity is all-zero bytes.
 *   - scalars are big-endian byte strings of a caller-stated fixed width.
 *   - every data pointer may be a HOST pointer or a CUDA DEVICE pointer on the
 *     context's device (detected with cudaPointerGetAttributes); host buffers are
 *     copied in/out inside the call.  Device buffers are read and written on the
 *     context's own stream, a blocking stream: it is ordered after work already queued
 *     on the legacy default stream (stream 0) and every call returns only after its
 *     device work has finished.  Inputs still being produced on ANY OTHER stream must
 *     be synchronised by the caller before the call.
 *   - the caller owns all buffers; the library keeps no caller pointer after return.
 *   - all functions return 0 on success or a negative bgn_status; none throws or aborts.
 *   - a context is thread-compatible: calls are synchronous (they return after the
 *     device work has finished) and serialised per DEVICE -- contexts on different
 *     GPUs may be driven concurrently from one process, one thread each.
 *
 * Homomorphic operations implement the reference's Deterministic=true behaviour
 * (bgn_test.go:13).  The non-deterministic mode is the same operation followed by
 * bgn_g1_blind_batch / bgn_gt_blind_batch with caller-supplied randomness (the
 * reference draws it from crypto/rand inside the call, bgn.go:567-574).
 */
#ifndef BGN_B200_H
#define BGN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bgn_ctx bgn_ctx;

typedef enum {
  BGN_OK = 0,
  BGN_E_BADARG = -1,    /* programmer error; the reference panics (bgn.go:67-73, 87-89, 389) */
  BGN_E_CUDA = -2,      /* CUDA runtime failure; see bgn_last_error */
  BGN_E_NOTSETUP = -3,  /* decrypt before bgn_ctx_set_secret: "DL tables not computed!" (gsbs.go:56-58) */
  BGN_E_NOMEM = -4,
  BGN_E_UNSUPPORTED = -5
} bgn_status;

/* Public parameters: what PublicKey{P, Q, N, PairingParams} carries (bgn.go:28-41).
 * p, n: big-endian magnitudes of the PBC "type a1" parameters (bgn.go:93-94);
 * l: the cofactor the reference parses out of the param string (bgn.go:583-593);
 * P, Q: Element.Bytes() of the generators (bgn.go:605-607), 2*B bytes each. */
typedef struct {
  const uint8_t* p_be;
  size_t p_len;
  const uint8_t* n_be;
  size_t n_len;
  uint64_t l;
  const uint8_t* P_bytes;
  const uint8_t* Q_bytes;
} bgn_params;

/* pbc.NewPairingFromString + G1.SetBytes (bgn.go:640-653): validates the
 * parameters, uploads Montgomery constants, builds the fixed-base window tables
 * for P and Q on the device.  device = CUDA ordinal. */
int bgn_ctx_create(const bgn_params* prm, int device, bgn_ctx** out);
void bgn_ctx_destroy(bgn_ctx* ctx);
const char* bgn_last_error(const bgn_ctx* ctx);
/* Message of the last failure that had no context to report through -- bgn_ctx_create, or a call on a
 * NULL context -- on the calling thread ("" if none).  bgn_last_error(NULL) returns the same string. */
const char* bgn_global_last_error(void);

/* Tuning knobs of a context (none changes any result):
 *   "enc_window"   0 | 8 | 16 | 18 | 20 | 22 | 24
 *                                fixed-base window of Q for Encrypt / level-1 re-randomisation, in bits.  0 (default):
 *                                the widest of 16 / 18 / 20 whose table stays within "enc_table_max_mb" (20 bits,
 *                                5.6 GB, at 512-bit keys; 18 bits, 5.9 GB, at 1024).  16: 428 MB at 512-bit keys;
 *                                22: 20.5 GB; 24: 75 GB, a third fewer additions than 16 (+42 % Encrypt
 *                                throughput), ~2 s to build -- for long-lived contexts.  A table that does not fit
 *                                the free device memory falls back to 16 bits.  The table is (re)built on the next
 *                                randomised encryption.
 *   "enc_table_max_mb"           bound of the automatic choice in MiB (default 6144)
 *   "enc_edwards"  0 | 1         Encrypt sums its table points in twisted Edwards form (8 products per point
 *                                instead of 11; tables hold 3 field elements per point instead of 2).
 *                                Default 1; 0 keeps Weierstrass tables.  Off by itself for a key whose P or Q
 *                                is not of odd order.
 *   "dec_lucas"    0 | 1         Decrypt through the Lucas ladder when one giant step suffices (default 1)
 *   "fixed_lines"  0 | 1         e(., P) through the recorded line table (default 1)
 *   "split_para"   -1 | 0 | 1    the sub-wave MultPoly kernel with merged doubling-and-addition steps (default on; 0: A/B)
 *   "dec_pair_q1"  0 | 1         level-1 Decrypt as ONE pairing e(C, q1 P) through a line table of q1*P built by
 *                                bgn_ctx_set_secret (default 1), instead of e(C, P) followed by the exponentiation
 *   "fixed_pair"   -1 | 0 | 1    e(., P) with one pairing split over a pair of lanes: -1 (default) below the
 *                                measured batch-size crossover, 0 never, 1 always
 *   "miller_split" -1 | 0 | 1    MultPoly work below one wave (a small batch, or the remainder after the full waves) on
 *                                the team kernel with two threads per output-slot pair: -1 (default) where the
 *                                measured time model says it is faster, 0 never, 1 always
 *   "pair_duo"     -1 | 0 | 1    e(a, b) (bgn_pair_batch) with one pairing split over two warps: -1 (default)
 *                                for batches within one wave of that kernel, 0 never, 1 always                */
int bgn_ctx_set_option(bgn_ctx* ctx, const char* name, long value);

/* limbs: 32-bit limbs of the field; coord_bytes: B; scalar_bytes: ceil(bits(n)/8). */
int bgn_ctx_info(const bgn_ctx* ctx, int* limbs, int* coord_bytes, int* scalar_bytes);

/* SetupDecryption / ComputeDecryptionPreprocessing (bgn.go:142-149, 195-201) +
 * PrecomputeTables (gsbs.go:41-51): q1 = SecretKey.Key, msg_space = PublicKey.MsgSpace.
 * Builds gsk = e(P,P)^q1 and the baby-step table as a device hash table.
 * baby_steps = 0 lets the library choose (it may hold more baby steps than the
 * reference's ceil(sqrt(T))+2; the set of decryptable values is kept identical). */
int bgn_ctx_set_secret(bgn_ctx* ctx, const uint8_t* q1_be, size_t q1_len, uint64_t msg_space, uint32_t baby_steps);

/* EncryptWithRandomness over a batch (bgn.go:340-353; the negative branch of
 * EncryptPoly, poly.go:17-21, is x < 0):  out[i] = x[i]*P + r[i]*Q for x >= 0 and
 * -(|x[i]|*P + r[i]*Q) for x < 0 (Sub(encryptZero(), Encrypt(|x|)), as EncryptPoly does).
 * r_be: count scalars of scalar_bytes each, or NULL for EncryptDeterministic
 * (bgn.go:325-331).  out: count G1 elements. */
int bgn_encrypt_batch(bgn_ctx* ctx, const int64_t* x, const uint8_t* r_be, size_t count, uint8_t* out);

/* Add / Sub / Neg on level-1 ciphertexts (bgn.go:482, 419, 436-439). */
int bgn_g1_add_batch(bgn_ctx* ctx, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out);
int bgn_g1_sub_batch(bgn_ctx* ctx, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out);
int bgn_g1_neg_batch(bgn_ctx* ctx, const uint8_t* a, size_t count, uint8_t* out);
/* MultConst on level 1 (bgn.go:258): out[i] = k[i]*a[i]; k_be: count scalars of kbytes each. */
int bgn_g1_mulconst_batch(bgn_ctx* ctx, const uint8_t* a, const uint8_t* k_be, size_t kbytes, size_t count,
                          uint8_t* out);

/* Add / Sub / Neg / MultConst on level-2 ciphertexts (bgn.go:460, 397, 277). */
int bgn_gt_mul_batch(bgn_ctx* ctx, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out);
int bgn_gt_div_batch(bgn_ctx* ctx, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out);
int bgn_gt_inv_batch(bgn_ctx* ctx, const uint8_t* a, size_t count, uint8_t* out);
int bgn_gt_pow_batch(bgn_ctx* ctx, const uint8_t* a, const uint8_t* k_be, size_t kbytes, size_t count, uint8_t* out);

/* Mult (bgn.go:294-314): out[i] = e(a[i], b[i]).  makeL2 (bgn.go:316-321): e(a[i], P). */
int bgn_pair_batch(bgn_ctx* ctx, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out);
int bgn_make_l2_batch(bgn_ctx* ctx, const uint8_t* a, size_t count, uint8_t* out);

/* MultPoly (poly.go:123-156) over a batch of `count` polynomial pairs:
 * c1: count*d1 G1 elements (coefficient-major per polynomial), c2: count*d2.
 * out: count*(d1+d2) GT elements, out[u][j] = prod_{i+k=j} e(c1[u][i], c2[u][k]);
 * the last slot of each polynomial is the GT identity, as in the reference. */
int bgn_multpoly_batch(bgn_ctx* ctx, const uint8_t* c1, size_t d1, const uint8_t* c2, size_t d2, size_t count,
                       uint8_t* out);

/* L2 sum of `nterms` polynomials of `ncoeff` slots each (AddPoly folded, poly.go:191-204
 * -> bgn.go:460): out[c] = prod_t in[t*ncoeff + c].  One GPU's share of an encrypted
 * inner product; partial results from several GPUs are folded with the same call. */
int bgn_l2_sum_reduce(bgn_ctx* ctx, const uint8_t* in, size_t nterms, size_t ncoeff, uint8_t* out);

/* csk = C^q1 (bgn.go:223), level 2 only: count GT in, count GT out. */
int bgn_gt_pow_secret_batch(bgn_ctx* ctx, const uint8_t* in, size_t count, uint8_t* out);

/* decrypt (bgn.go:218-250) + recoverMessage (bgn.go:357-372) + getDL (gsbs.go:54-106),
 * including the negate-and-retry path.  status[i]: 0 ok, 1 "cannot find discrete
 * log; out of bounds" (gsbs.go:105), in which case out[i] = 0 (DecryptFailSafe). */
int bgn_decrypt_batch(bgn_ctx* ctx, const uint8_t* in, int is_l2, size_t count, int64_t* out, uint8_t* status);

/* ---- non-deterministic mode: the re-randomisation every homomorphic operation appends when
 * !pk.Deterministic.  r_be: count scalars of scalar_bytes each (the reference's newCryptoRandom(N)).
 * Level 1:  out[i] = a[i] + r[i]*Q              (bgn.go:260-269, 421-432, 488-495)
 * Level 2:  out[i] = a[i] * e(Q,Q)^r[i]         (bgn.go:279-288, 302-311, 404-411, 466-474)
 * e(Q,Q) -- a full pairing per call in the reference (bgn.go:283, 306, 406, 469) -- is computed once
 * per context and kept as a fixed-base table.  MultPoly in this mode (one r per coefficient pairing,
 * poly.go:140-152) is bgn_multpoly_batch followed by one blind per slot j with
 * r_j = sum_{i+k=j} r_ik mod n. */
int bgn_g1_blind_batch(bgn_ctx* ctx, const uint8_t* a, const uint8_t* r_be, size_t count, uint8_t* out);
int bgn_gt_blind_batch(bgn_ctx* ctx, const uint8_t* a, const uint8_t* r_be, size_t count, uint8_t* out);

/* ---- polynomial-ciphertext helpers beside MultPoly.
 * MultConstPoly (poly.go:71-120) over `count` polynomials of d slots (level 1 or 2): digits = the nd
 * coefficients of NewUnbalancedPlaintext(|constant|) (values 0..base-1, poly.go:78-80);
 * out: count*(d+nd) elements, out[u][j] = sum_{i+k=j} digits[k]*in[u][i] (top slot = identity);
 * negate != 0 applies NegPoly (negative constant, poly.go:116-118). */
int bgn_multconstpoly_batch(bgn_ctx* ctx, const uint8_t* in, size_t d, int is_l2, const uint8_t* digits, size_t nd,
                            int negate, size_t count, uint8_t* out);
/* EvalPoly (poly.go:58-68): out[u] = sum_i base^i * in[u][i]  (Horner with MultConst/Add in the
 * reference); base = PolyEncodingParams.PolyBase; d <= 64 and base^(d-1) < 2^128
 * (base 3: any d <= 64). */
int bgn_evalpoly_batch(bgn_ctx* ctx, const uint8_t* in, size_t d, int is_l2, uint32_t base, size_t count, uint8_t* out);
/* MakePolyL2 (poly.go:159-163) in deterministic mode: MultPoly(E(1.0), ct) with E(1.0) = [P]:
 * out: count*(d+1) GT elements, out[u][i] = e(in[u][i], P), out[u][d] = identity. */
int bgn_make_poly_l2_batch(bgn_ctx* ctx, const uint8_t* in, size_t d, size_t count, uint8_t* out);

/* ---- device-resident batches: chained operations without the byte format in between.
 * A bgn_buf holds `count` elements of one group on the context's device in the kernels' own form.  The
 * byte entry points above convert, and for G1 check the curve equation, on every call -- for a level-1
 * addition that costs as much as the addition; a pipeline Encrypt -> EAdd -> EMult -> L2 sum -> Decrypt on
 * handles runs kernel to kernel and meets the PBC byte format only at bgn_buf_import / bgn_buf_export.
 * Same results as the byte forms, bit for bit.  Output handles: pass the address of a NULL handle to have
 * one allocated, or of an earlier result to reuse its memory (it is reallocated if too small); an output
 * must not alias an operand unless stated.  Handles belong to their context; free them before it. */
typedef struct bgn_buf bgn_buf;
enum { BGN_KIND_G1 = 1, BGN_KIND_GT = 2 };
int bgn_buf_import(bgn_ctx* ctx, int kind, const uint8_t* bytes, size_t count, bgn_buf** out); /* Element.SetBytes */
int bgn_buf_export(bgn_ctx* ctx, const bgn_buf* buf, uint8_t* bytes_out);                      /* Element.Bytes */
int bgn_buf_info(const bgn_buf* buf, int* kind, size_t* count);
void bgn_buf_free(bgn_buf* buf);
/* bgn_encrypt_batch / g1_add|sub / gt_mul|div / pair (b = NULL: makeL2, e(a, P)) / multpoly / l2_sum_reduce /
 * decrypt (level taken from the handle's kind) on handles */
int bgn_encrypt_h(bgn_ctx* ctx, const int64_t* x, const uint8_t* r_be, size_t count, bgn_buf** out);
int bgn_g1_add_h(bgn_ctx* ctx, const bgn_buf* a, const bgn_buf* b, int subtract, bgn_buf** out);
int bgn_gt_mul_h(bgn_ctx* ctx, const bgn_buf* a, const bgn_buf* b, int divide, bgn_buf** out);
int bgn_pair_h(bgn_ctx* ctx, const bgn_buf* a, const bgn_buf* b, bgn_buf** out);
int bgn_multpoly_h(bgn_ctx* ctx, const bgn_buf* c1, size_t d1, const bgn_buf* c2, size_t d2, size_t count, bgn_buf** out);
int bgn_l2_sum_reduce_h(bgn_ctx* ctx, const bgn_buf* in, size_t nterms, size_t ncoeff, bgn_buf** out);
int bgn_decrypt_h(bgn_ctx* ctx, const bgn_buf* in, int64_t* out, uint8_t* status);

/* ---- several GPUs behind one handle ("multi-GPU is driven inside one call by per-device host
 * threads").  A bgn_group owns one context per listed device (a device may be listed more than once: its contexts
 * then take turns); every batch call cuts the batch into contiguous shards -- the reference's independent units
 * (poly.go:15, 37, 140-141) -- and runs each shard on its device from its own host thread.  Buffers are HOST
 * pointers.  Results are identical to the single-context calls.  Per-member access (options, handles, timing):
 * bgn_group_ctx(grp, i). */
typedef struct bgn_group bgn_group;
int bgn_group_create(const bgn_params* prm, int ndev, const int* devs, bgn_group** out);
void bgn_group_destroy(bgn_group* grp);
int bgn_group_size(const bgn_group* grp);
bgn_ctx* bgn_group_ctx(bgn_group* grp, int i);
const char* bgn_group_last_error(const bgn_group* grp);
int bgn_group_set_secret(bgn_group* grp, const uint8_t* q1_be, size_t q1_len, uint64_t msg_space, uint32_t baby_steps);
int bgn_group_set_option(bgn_group* grp, const char* name, long value);
int bgn_group_encrypt_batch(bgn_group* grp, const int64_t* x, const uint8_t* r_be, size_t count, uint8_t* out);
int bgn_group_g1_add_batch(bgn_group* grp, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out);
int bgn_group_multpoly_batch(bgn_group* grp, const uint8_t* c1, size_t d1, const uint8_t* c2, size_t d2, size_t count,
                             uint8_t* out);
int bgn_group_decrypt_batch(bgn_group* grp, const uint8_t* in, int is_l2, size_t count, int64_t* out, uint8_t* status);
/* sum_i c1[i] * c2[i] as one polynomial ciphertext of d1 + d2 level-2 slots (AddPoly folded over MultPoly,
 * poly.go:123-156, 191-204): per-device MultPoly + GT product tree, partials folded on the first device. */
int bgn_group_inner_product(bgn_group* grp, const uint8_t* c1, size_t d1, const uint8_t* c2, size_t d2, size_t count,
                            uint8_t* out);

/* ---- instrumentation (bench.py) ---- */
/* When enabled, every kernel launch is bracketed by CUDA events on the context's
 * stream; bgn_timing_get returns the accumulated device time and launch count of
 * kernels whose name starts with `prefix` ("" = all) since the last reset. */
int bgn_timing_enable(bgn_ctx* ctx, int on);
int bgn_timing_reset(bgn_ctx* ctx);
int bgn_timing_get(bgn_ctx* ctx, const char* prefix, double* ms_total, uint64_t* launches);
/* Device time of the last C-ABI call on this context, first copy-in to last copy-out
 * (CUDA events on the context's stream); valid while timing is enabled. */
int bgn_timing_last_call(bgn_ctx* ctx, double* ms);
/* Register-resident Montgomery-product microbenchmark: blocks*threads threads,
 * `iters` dependent products on each of `ilp` (1, 2 or 4) chains; returns device ms. */
int bgn_bench_mulmod(bgn_ctx* ctx, int ilp, int iters, int blocks, int threads, float* ms);
/* Peak rate of the 32x32+64 multiply-add instruction (IMAD.WIDE.U32) on `device`:
 * blocks*threads threads issue iters*64 independent-chain instructions each. */
int bgn_bench_imad_peak(int device, int iters, int blocks, int threads, float* ms, double* instr_per_thread);

/* Issue-mix microbenchmark on `device`: which instruction classes share the integer-multiply pipe.
 * mix: 0 IMAD.WIDE | 1 IMAD.LO+IMAD.HI | 2 both | 3 FFMA | 4 IMAD.WIDE+FFMA | 5 DFMA | 6 IMAD.WIDE+DFMA |
 *      7..9 other ratios | 10 IMAD.LO | 11 IMAD.HI | 12 IMAD.WIDE+IMAD.LO | 13 IMAD.WIDE+IMAD.HI (api.cu).
 * per_thread[5] receives the instructions of each class one thread issued:
 * {IMAD.WIDE, IMAD.LO, IMAD.HI, FFMA, DFMA}; *ms the device time. */
int bgn_bench_issue_mix(int device, int mix, int iters, int blocks, int threads, float* ms, double* per_thread);

#ifdef __cplusplus
}
#endif
#endif /* BGN_B200_H */
