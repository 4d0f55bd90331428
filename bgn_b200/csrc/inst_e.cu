// inst_e.cu -- the split team kernel (teamsplit.cuh) for one limb count (compile with -DBGN_L=<L>).
// See ops.h: LOpsE.
#define BGN_GROUP_E 1
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
#if BGN_L <= 17
size_t miller_split_smem_bytes(int nt, int ncol) { return MillerSplit<LL>::smem_words(nt, ncol) * 4; }
cudaError_t miller_split_set_smem(size_t smem) {
  return cudaFuncSetAttribute(k_miller_split<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
void miller_split(LaunchCfg cfg, const MillerArgs& a) { k_miller_split<LL><<<CFG>>>(a); }
size_t miller_wide_smem_bytes(int nt) { return MillerTeam<LL, true>::smem_words(nt) * 4; }
cudaError_t miller_wide_set_smem(size_t smem) {
  return cudaFuncSetAttribute(k_miller_wide<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
void miller_wide(LaunchCfg cfg, const MillerArgs& a) { k_miller_wide<LL><<<CFG>>>(a); }
const LOpsE ops = {LL, upload, miller_split_smem_bytes, miller_split_set_smem, miller_split,
                   miller_wide_smem_bytes, miller_wide_set_smem, miller_wide};
#else
const LOpsE ops = {LL, upload, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
#endif
}  // namespace
#define BGN_CAT2(a, b) a##b
#define BGN_CAT(a, b) BGN_CAT2(a, b)
extern "C" const LOpsE* BGN_CAT(bgn_opsE_, BGN_L)() { return &ops; }
