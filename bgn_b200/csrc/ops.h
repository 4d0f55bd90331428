// ops.h -- per-limb-count launcher tables.  Each supported limb count L is
// compiled in its own translation units (inst_a.cu: pairing / GT kernels,
// inst_b.cu: G1, (de)serialisation and BSGS kernels) so the build parallelises;
// every unit owns a copy of the __constant__ key material and exports `upload`.
#pragma once
#include <cuda_runtime.h>

#include "types.h"

struct LaunchCfg {
  unsigned grid, block;
  size_t smem;
  cudaStream_t stream;
};

struct LOpsA {
  int L;
  cudaError_t (*upload)(const FieldConsts*, const PairConsts*, cudaStream_t);
  cudaError_t (*miller_set_smem)(size_t smem);
  size_t (*miller_smem_bytes)(int nt);   // shared memory a block of nt threads needs
  size_t (*miller_priv_bytes)();         // global scratch per block (0: all state in shared memory)
  int (*miller_fixed_threads)();         // != 0: the layout fixes blockDim to this value
  void (*miller)(LaunchCfg, const MillerArgs&);
  void (*gt_mul)(LaunchCfg, const GtBinArgs&);
  void (*gt_pow)(LaunchCfg, const GtPowArgs&);
  void (*gt_reduce)(LaunchCfg, const uint32_t* re, const uint32_t* im, size_t Nin, size_t nterms, int ncoeff, int G,
                    uint32_t* ore, uint32_t* oim, size_t N);
  void (*fp2_from_bytes)(LaunchCfg, const uint8_t* in, int B, size_t count, uint32_t* re, uint32_t* im, size_t N);
  void (*fp2_to_bytes)(LaunchCfg, const uint32_t* re, const uint32_t* im, size_t N, size_t count, uint8_t* out, int B,
                       int grp, int pad);
  void (*bsgs_build)(LaunchCfg, const BsgsBuildArgs&);
  void (*bsgs_lookup)(LaunchCfg, const BsgsLookupArgs&);
  void (*mulmod_bench)(LaunchCfg, int ilp, uint32_t* io, size_t N, int iters);
};

// inst_c.cu: the kernels added after the headline path (re-randomisation, polynomial helpers,
// Lucas decrypt, fixed-argument pairing).  Their own translation unit keeps the register allocation
// of k_miller in inst_a.cu independent of them (measured: sharing a unit cost k_miller 1.5 %).
struct LOpsC {
  int L;
  cudaError_t (*upload)(const FieldConsts*, const PairConsts*, cudaStream_t);
  void (*gt_blind)(LaunchCfg, const GtBlindArgs&);
  void (*gt_tab_bases)(LaunchCfg, const uint32_t* gen, int nwin, uint32_t* bases);
  void (*gt_tab_fill)(LaunchCfg, const uint32_t* bases, int nwin, uint32_t* tab);
  void (*gt_polyconv)(LaunchCfg, const PolyConvArgs&);
  void (*dec_lucas)(LaunchCfg, const DecLucasArgs&);
  void (*gt_pow_pair)(LaunchCfg, const GtPowArgs&);
  cudaError_t (*miller_fixed_set_smem)(size_t smem);
  size_t (*miller_fixed_smem_bytes)(int nt);
  void (*miller_fixed)(LaunchCfg, const MillerFixedArgs&);
  void (*miller_record)(LaunchCfg, const uint32_t* px, const uint32_t* py, uint32_t* lines, uint32_t* scratch, int* ok);
};

// inst_d.cu: kernels that split one item over a pair of lanes (pairlane.cuh), added in round 2 for
// batches below one wave; their own translation unit for the same reason as inst_c.cu.
struct LOpsD {
  int L;
  cudaError_t (*upload)(const FieldConsts*, const PairConsts*, cudaStream_t);
  void (*miller_fixed_pair)(LaunchCfg, const MillerFixedArgs&);
  int (*miller_fixed_pair_blocks_per_sm)();  // resident 64-thread blocks of k_miller_fixed_pair per SM (register-bound)
  // general pairing on two warps (pairwarp.cuh); blocks are `pairs` warp pairs = 32 * pairs pairings
  // variant = loop shape of the products (0 unrolled, 1 / 2 / 4 = 2U rows per iteration; only the
  // shipped one and the A/B candidates are instantiated) + 8 if the barrier spans the whole block
  size_t (*pair_duo_smem_bytes)(int pairings_per_block);
  cudaError_t (*pair_duo_set_smem)(int variant, size_t smem);
  int (*pair_duo_blocks_per_sm)(int variant, int threads, size_t smem);
  bool (*pair_duo)(int variant, LaunchCfg, const PairDuoArgs&);  // false: variant not instantiated
};

// inst_e.cu: the team kernel with two threads per output-slot pair (teamsplit.cuh).  It runs 12 warps per
// SM, which caps it at 168 registers; ptxas applies a unit's tightest cap to the device functions its
// kernels share, so it lives alone (in inst_d.cu it cost k_miller_fixed_pair 80 bytes of spills).
struct LOpsE {
  int L;
  cudaError_t (*upload)(const FieldConsts*, const PairConsts*, cudaStream_t);
  size_t (*miller_split_smem_bytes)(int nt, int ncol);  // null above 17 limbs
  cudaError_t (*miller_split_set_smem)(size_t smem);
  void (*miller_split)(LaunchCfg, const MillerArgs&);
  // the team kernel with 10 slots per thread, 10 warps per SM (k_miller_wide); null above 17 limbs
  size_t (*miller_wide_smem_bytes)(int nt);
  cudaError_t (*miller_wide_set_smem)(size_t smem);
  void (*miller_wide)(LaunchCfg, const MillerArgs&);
};

struct LOpsB {
  int L;
  cudaError_t (*upload)(const FieldConsts*, const PairConsts*, cudaStream_t);
  void (*g1_from_bytes)(LaunchCfg, const uint8_t* in, int B, size_t count, uint32_t* x, uint32_t* y, uint8_t* inf,
                        size_t N);
  void (*g1_to_bytes)(LaunchCfg, const uint32_t* x, const uint32_t* y, const uint8_t* inf, size_t N, size_t count,
                      uint8_t* out, int B);
  void (*encrypt)(LaunchCfg, const EncArgs&);
  void (*normalize)(LaunchCfg, const NormArgs&);
  void (*g1_add)(LaunchCfg, const G1AddArgs&);
  void (*g1_mulvar)(LaunchCfg, const G1MulArgs&);
  void (*tab_bases)(LaunchCfg, const uint32_t* bx, const uint32_t* by, int nwin, int hb, uint32_t* X, uint32_t* Y,
                    uint32_t* Z, size_t N);
  void (*tab_fill)(LaunchCfg, const uint32_t* ax, const uint32_t* ay, const uint8_t* ainf, size_t Nb, int nwin, int hb,
                   uint32_t* X, uint32_t* Y, uint32_t* Z, size_t N);
  void (*tabw_fill)(LaunchCfg, const uint32_t* tabh, int nwin_h, int nsub, int hb, uint32_t* X, uint32_t* Y, uint32_t* Z,
                    size_t first, size_t nent);
  void (*tab_edwards)(LaunchCfg, const uint32_t* tabw, uint32_t* tabe, uint32_t* scratch, size_t count, int G, int* bad);
  void (*g1_polyconv)(LaunchCfg, const PolyConvArgs&);
  void (*g1_affadd)(LaunchCfg, const G1AffAddArgs&);
};

#define BGN_DECL_OPS(L)               \
  extern "C" const LOpsA* bgn_opsA_##L(); \
  extern "C" const LOpsB* bgn_opsB_##L(); \
  extern "C" const LOpsC* bgn_opsC_##L(); \
  extern "C" const LOpsD* bgn_opsD_##L(); \
  extern "C" const LOpsE* bgn_opsE_##L();
BGN_DECL_OPS(3)
BGN_DECL_OPS(5)
BGN_DECL_OPS(9)
BGN_DECL_OPS(17)
BGN_DECL_OPS(33)
