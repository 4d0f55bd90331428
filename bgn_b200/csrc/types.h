// types.h -- plain structs shared by the host orchestration (api.cpp) and the
// per-limb-count kernel translation units (inst.cu).
#pragma once
#include <stddef.h>
#include <stdint.h>

#define BGN_MAXL 34
struct FieldConsts {
  uint32_t p[BGN_MAXL];    // modulus
  uint32_t p2[BGN_MAXL];   // 2p
  uint32_t one[BGN_MAXL];  // R mod p   (Montgomery 1)
  uint32_t r2[BGN_MAXL];   // R^2 mod p (to-Montgomery factor)
  uint32_t p4[BGN_MAXL];   // 4p, 8p, 16p: offsets that keep relaxed-range differences non-negative
  uint32_t p8[BGN_MAXL];
  uint32_t p16[BGN_MAXL];
  uint32_t np0;            // -p^{-1} mod 2^32
  uint32_t pad[3];
};

#define BGN_MAX_NAF 1100
#define BGN_MAX_EXPW 34
struct PairConsts {
  uint64_t l;                    // cofactor, p + 1 = l*n
  int32_t naf_len;               // number of signed digits of n (MSB first, naf[0] == 1)
  int32_t exp_bits;              // bit length of the fixed GT exponent (secret q1)
  uint32_t exp[BGN_MAX_EXPW];    // fixed exponent, little-endian words
  int8_t naf[BGN_MAX_NAF];
  // appended (keeps every earlier offset): signed digits of the fixed exponent, MSB first, for
  // k_gt_pow_pair -- inverses in GT are conjugates, so a digit -1 costs what a digit +1 does
  int32_t exp_naf_len;
  int8_t exp_naf[BGN_MAX_NAF];
};

#define BGN_MILLER_NSLOT 12  // F_p slots per thread of the Miller team kernel
#define BGN_MILLER_NPRIV 7   // of which thread-private (accumulators, Miller point): slots 0..6
struct MillerArgs {
  const uint32_t* Mx;  // Miller-side points, Montgomery [L][NM], index unit*dM + i
  const uint32_t* My;
  const uint8_t* Minf;  // 1 = point at infinity
  const uint32_t* Ex;   // evaluation-side points, SoA [L][NE], index unit*dE + k
  const uint32_t* Ey;
  const uint8_t* Einf;
  uint32_t* priv;    // global scratch for the thread-private slots (interleaved layout only, else null)
  uint32_t* out_re;  // GT out, Montgomery [L][NOUT], index unit*out_slots + j
  uint32_t* out_im;
  int NM, NE, NOUT;
  int e_bcast;      // evaluation side is ONE polynomial shared by all units (makeL2: B = P)
  int dM, dE;       // dM <= dE; team size TS = dE
  int out_slots;    // slots written per unit (dM+dE for MultPoly: last one is the identity; 1 for Pair)
  int count;        // units
  int teams_per_group;  // whole teams in one barrier group
  int group_threads;    // threads per barrier group (128 on the GPU); blockDim = groups * group_threads
  int skew_cycles;      // start-up delay of odd groups (decorrelates the two warps of a scheduler)
  // global scratch of the parabola steps (pairing.cuh: BGN_PARABOLA): x^2 / y of every evaluation point,
  // [count * dE][L] ([dE][L] with e_bcast); written by the kernel's init, read in its phase B
  uint32_t* evw;
  int para;  // use the parabola steps: k_miller when a Miller point serves >= 3 evaluation points (below that the
             // 5 extra products of the merged step outweigh the saved F_p^2 products), k_miller_split by block geometry
};

// Pairing with a FIXED first argument (makeL2: e(C, P) = e(P, C), bgn.go:316-321; level-1 decrypt;
// MakePolyL2): the Miller point's multiples and therefore the line coefficients of every step do
// not depend on the batch, so they are computed once per key (MillerFixed::record) and a thread only
// squares its accumulator and folds in the line evaluated at its own point: 2 + 5 products per
// step instead of 12..13 + 2 + 5, no inter-thread traffic, no barriers.
struct MillerFixedArgs {
  const uint32_t* lines;  // [nsteps][2][L]: cR / bI, aR / bI of every line (pairing.cuh: MillerFixed::record)
  const uint32_t *Ex, *Ey;  // evaluation points, affine Montgomery [count][L]
  const uint8_t* Einf;
  uint32_t *out_re, *out_im;  // [count][L]
  int count;
};

// General pairing on two warps per 32 pairings (pairwarp.cuh): out[i] = e(M[i], E[i])
struct PairDuoArgs {
  const uint32_t *Mx, *My;  // Miller-side points, affine Montgomery [count][L]
  const uint8_t* Minf;
  const uint32_t *Ex, *Ey;  // evaluation-side points
  const uint8_t* Einf;
  uint32_t *out_re, *out_im;  // [count][L]
  int count;
};

struct EncArgs {
  const int64_t* x;       // plaintext scalars (signed, |x| < 2^63)
  const uint8_t* r_be;    // randomness, big-endian, rbytes each (may be null: r = 0)
  int rbytes;
  const uint32_t* tabP;   // 8 windows
  const uint32_t* tabQ;   // ceil(8 rbytes / wbitsQ) windows of wbitsQ bits
  int wbitsQ;             // 8 .. 24: window width of tabQ (entries per window = 2^wbitsQ - 1)
  uint32_t *X, *Y, *Z;    // Jacobian out, [N][L]
  size_t count, N;
  // optional starting point per element (affine, [count][L] + infinity flags): out = base + r*Q,
  // the re-randomisation of the non-deterministic mode (bgn.go:264-268, 491-495).  x may then be null.
  const uint32_t *bx, *by;
  const uint8_t* binf;
  // edw != 0: tabP and tabQ hold twisted Edwards points, 3L words per entry (u | v | u v) (curve.cuh: Ed);
  // the sum is formed in extended coordinates and converted to Jacobian before the starting point is added
  int edw;
};

struct NormArgs {
  const uint32_t *X, *Y, *Z;  // [N][L]
  uint32_t* scratch;          // [N][L] prefix products
  size_t count, N;
  int G;                      // number of worker threads
  uint32_t* ox;               // out x: element e, limb j at ox[e*o_estride + j*o_lstride]
  uint32_t* oy;
  size_t o_estride, o_lstride;
  uint8_t* inf;               // out flags (may be null)
};

struct G1AddArgs {
  const uint32_t *x1, *y1;
  const uint8_t* inf1;
  const uint32_t *x2, *y2;
  const uint8_t* inf2;
  size_t N1, N2;
  int bcast1;   // operand 1 is a single element (Neg: O - c)
  int subtract;
  uint32_t *X, *Y, *Z;
  size_t count, N;
};

// EAdd / ESub entirely in affine coordinates: one inversion (binary GCD) per thread shared by the
// elements g, g+G, g+2G, ... of thread g (Montgomery's trick on the denominators x2 - x1, or 2 y1
// where the operands coincide), 6 products per element instead of the 11 + 6 of a mixed Jacobian
// addition followed by the batched normalisation.
struct G1AffAddArgs {
  const uint32_t *x1, *y1;
  const uint8_t* inf1;
  const uint32_t *x2, *y2;
  const uint8_t* inf2;
  int bcast1;    // operand 1 is a single element (Neg: O - c)
  int subtract;
  uint32_t *ox, *oy;
  uint8_t* oinf;
  uint32_t* scratch;  // [count][L] prefix products
  size_t count;
  int G;              // worker threads
};

struct G1MulArgs {
  const uint32_t *x, *y;
  const uint8_t* inf;
  size_t Nin;
  const uint8_t* k_be;
  int kbytes;
  uint32_t *X, *Y, *Z;
  size_t count, N;
};

struct GtBinArgs {
  const uint32_t *are, *aim, *bre, *bim;
  size_t Na, Nb;
  int conj_b;  // division of unitary elements: a * conj(b)   (bgn.go:397)
  uint32_t *ore, *oim;
  size_t count, N;
};

struct GtPowArgs {
  const uint32_t *re, *im;
  size_t Nin;
  const uint8_t* e_be;  // per-element exponents (modes 0 and 3)
  int ebytes;
  int mode;             // 0: a^e[i]; 1: a^q1 (c_pc.exp); 2: conj(a) = a^-1 (unitary); 3: conj(a^e[i])
  uint32_t *ore, *oim;
  size_t count, N;
};

// out[i] = a[i] * E^r[i] with E = e(Q,Q): the level-2 re-randomisation (bgn.go:283-287, 306-310,
// 469-474) through a fixed-base table of E (8-bit windows, AoS [win][d-1][re||im], canonical)
struct GtBlindArgs {
  const uint32_t *re, *im;
  const uint8_t* r_be;
  int rbytes;
  const uint32_t* tabE;
  uint32_t *ore, *oim;
  size_t count;
};

// Integer-weighted correlation of coefficient vectors: out[u][jj] = sum_k w[k] * in[u][j_begin + jj - k]
// over the k with 0 <= j_begin + jj - k < d.  MultConstPoly (poly.go:71-120) is w = the unbalanced
// digits of the constant, j_begin = 0, j_count = d + nw; EvalPoly (poly.go:58-68) is w[k] =
// base^(d-1-k), j_begin = d - 1, j_count = 1.  G1 inputs are affine (x, y, inf) and the result is
// Jacobian (X, Y, Z); GT inputs/outputs use (x, y) / (X, Y) as (re, im).
#define BGN_CONV_MAXW 64
struct PolyConvArgs {
  const uint32_t *x, *y;
  const uint8_t* inf;  // G1 only
  int d, nw, j_begin, j_count;
  int top_bit;         // highest set bit over all weights
  int negate;          // NegPoly of the result (negative constant)
  uint64_t w[BGN_CONV_MAXW];    // weights, low 64 bits
  uint64_t whi[BGN_CONV_MAXW];  // bits 64..127 (EvalPoly of long polynomials: base^(d-1) up to 2^128)
  uint32_t *X, *Y, *Z;
  size_t count;        // polynomials
};

struct BsgsBuildArgs {
  const uint32_t* gen;   // generator gsk, AoS [2L] Montgomery
  uint32_t* elems;       // [S][2L]
  uint32_t* slots;
  uint32_t hmask;
  uint32_t S;
  int chunk;             // baby steps per thread
};

struct BsgsLookupArgs {
  const uint32_t *re, *im;  // csk = C^q1, SoA [L][Nin]
  size_t Nin, count;
  const uint32_t* elems;
  const uint32_t* slots;
  uint32_t hmask;
  uint32_t S;
  const uint32_t* ginv;     // gen^(-S), AoS [2L]
  uint32_t giant_steps;     // ceil(Mmax / S)
  uint64_t mmax;            // largest |m| the reference's table/loop bounds can return
  int64_t* out;
  uint8_t* status;          // 0 ok, 1 out of bounds (gsbs.go:105)
};

// Decrypt with the whole message space in the baby-step table (lucas.cuh)
struct DecLucasArgs {
  const uint32_t *re, *im;  // C, Montgomery [count][L], values below 2p
  size_t count;
  const uint32_t* elems;    // baby steps gsk^(j+1), canonical Montgomery AoS [S][2L]
  const uint32_t* slots;    // open addressing on the real part, value j+1 (0 = empty)
  uint32_t hmask;
  uint32_t S;
  uint64_t mmax;
  int64_t* out;
  uint8_t* status;
};

