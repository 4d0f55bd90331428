// inst_d.cu -- lane-pair kernels (pairlane.cuh) for one limb count (compile with -DBGN_L=<L>).
// See ops.h: LOpsD.
#define BGN_GROUP_D 1
#include "kernels.cuh"
#include "ops.h"
#ifndef BGN_L
#error "compile with -DBGN_L=<limbs>"
#endif
namespace {
constexpr int LL = BGN_L;
#define CFG cfg.grid, cfg.block, cfg.smem, cfg.stream
cudaError_t upload(const FieldConsts* fc, const PairConsts* pc, cudaStream_t s) {
  cudaError_t e = cudaMemcpyToSymbolAsync(c_fc, fc, sizeof(FieldConsts), 0, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbolAsync(c_pc, pc, sizeof(PairConsts), 0, cudaMemcpyHostToDevice, s);
}
void miller_fixed_pair(LaunchCfg cfg, const MillerFixedArgs& a) { k_miller_fixed_pair<LL><<<CFG>>>(a); }
int miller_fixed_pair_blocks_per_sm() {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_miller_fixed_pair<LL>, 64, 0) != cudaSuccess) n = 0;
  return n;
}
size_t pair_duo_smem_bytes(int np) { return MillerDuo<LL>::smem_words(np) * 4; }
// the variants compiled in: every limb count has its default loop shape; the 512-bit field also carries
// the A/B candidates (tools/mapping_ab.py --which general)
#if BGN_L == 17
#define BGN_DUO_VARIANTS(X) X(0) X(1) X(2) X(4)
#elif BGN_L <= 17
#define BGN_DUO_VARIANTS(X) X(0)
#else
#define BGN_DUO_VARIANTS(X) X(4) X(2)
#endif
template <typename Fn>
bool duo_dispatch(int variant, Fn fn) {
  const bool blockbar = (variant & 8) != 0;
  switch (variant & 7) {
#define BGN_DUO_CASE(U)                            \
  case U:                                          \
    if (blockbar)                                  \
      fn(k_pair_duo<LL, U, true>);                 \
    else                                           \
      fn(k_pair_duo<LL, U, false>);                \
    return true;
    BGN_DUO_VARIANTS(BGN_DUO_CASE)
#undef BGN_DUO_CASE
    default:
      return false;
  }
}
cudaError_t pair_duo_set_smem(int variant, size_t smem) {
  cudaError_t e = cudaErrorInvalidValue;
  duo_dispatch(variant, [&](auto k) { e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); });
  return e;
}
int pair_duo_blocks_per_sm(int variant, int threads, size_t smem) {
  int n = 0;
  duo_dispatch(variant, [&](auto k) {
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k, threads, smem) != cudaSuccess) n = 0;
  });
  return n;
}
bool pair_duo(int variant, LaunchCfg cfg, const PairDuoArgs& a) {
  return duo_dispatch(variant, [&](auto k) { k<<<CFG>>>(a); });
}
const LOpsD ops = {LL, upload, miller_fixed_pair, miller_fixed_pair_blocks_per_sm, pair_duo_smem_bytes, pair_duo_set_smem,
                   pair_duo_blocks_per_sm, pair_duo};
}  // namespace
#define BGN_CAT2(a, b) a##b
#define BGN_CAT(a, b) BGN_CAT2(a, b)
extern "C" const LOpsD* BGN_CAT(bgn_opsD_, BGN_L)() { return &ops; }
