// inst_c.cu -- re-randomisation, polynomial helpers, Lucas decrypt and the fixed-argument pairing
// for one limb count (compile with -DBGN_L=<L>).  See ops.h: LOpsC.
#define BGN_GROUP_C 1
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
void gt_blind(LaunchCfg cfg, const GtBlindArgs& a) { k_gt_blind<LL><<<CFG>>>(a); }
void gt_tab_bases(LaunchCfg cfg, const uint32_t* gen, int nwin, uint32_t* bases) {
  k_gt_tab_bases<LL><<<CFG>>>(gen, nwin, bases);
}
void gt_tab_fill(LaunchCfg cfg, const uint32_t* bases, int nwin, uint32_t* tab) {
  k_gt_tab_fill<LL><<<CFG>>>(bases, nwin, tab);
}
void gt_polyconv(LaunchCfg cfg, const PolyConvArgs& a) { k_gt_polyconv<LL><<<CFG>>>(a); }
void dec_lucas(LaunchCfg cfg, const DecLucasArgs& a) { k_dec_lucas<LL><<<CFG>>>(a); }
void gt_pow_pair(LaunchCfg cfg, const GtPowArgs& a) { k_gt_pow_pair<LL><<<CFG>>>(a); }
cudaError_t miller_fixed_set_smem(size_t smem) {
  return cudaFuncSetAttribute(k_miller_fixed<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
size_t miller_fixed_smem_bytes(int nt) { return MillerFixed<LL>::smem_words(nt) * 4; }
void miller_fixed(LaunchCfg cfg, const MillerFixedArgs& a) { k_miller_fixed<LL><<<CFG>>>(a); }
void miller_record(LaunchCfg cfg, const uint32_t* px, const uint32_t* py, uint32_t* lines, uint32_t* scratch, int* ok) {
  k_miller_record<LL><<<CFG>>>(px, py, lines, scratch, ok);
}
const LOpsC ops = {LL,          upload,    gt_blind,
                   gt_tab_bases, gt_tab_fill, gt_polyconv,
                   dec_lucas,   gt_pow_pair, miller_fixed_set_smem, miller_fixed_smem_bytes,
                   miller_fixed, miller_record};
}  // namespace
#define BGN_CAT2(a, b) a##b
#define BGN_CAT(a, b) BGN_CAT2(a, b)
extern "C" const LOpsC* BGN_CAT(bgn_opsC_, BGN_L)() { return &ops; }
