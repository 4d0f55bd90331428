// inst_a.cu -- pairing / GT kernels for one limb count (compile with -DBGN_L=<L>).
#define BGN_GROUP_A 1
#if BGN_L == 17
#define BGN_PRIM_UNROLLED 1  // primitive microbenchmarks of the fully unrolled fused routines (tools/primbench.py)
#endif
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
cudaError_t miller_set_smem(size_t smem) {
  return cudaFuncSetAttribute(k_miller<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}
size_t miller_smem_bytes(int nt) { return MillerTeam<LL>::smem_words(nt) * 4; }
size_t miller_priv_bytes() { return MillerTeam<LL>::priv_words() * 4; }
int miller_fixed_threads() { return MillerTeam<LL>::fixed_threads(); }
void miller(LaunchCfg cfg, const MillerArgs& a) { k_miller<LL><<<CFG>>>(a); }
void gt_mul(LaunchCfg cfg, const GtBinArgs& a) { k_gt_mul<LL><<<CFG>>>(a); }
void gt_pow(LaunchCfg cfg, const GtPowArgs& a) { k_gt_pow<LL><<<CFG>>>(a); }
void gt_reduce(LaunchCfg cfg, const uint32_t* re, const uint32_t* im, size_t Nin, size_t nterms, int ncoeff, int G,
               uint32_t* ore, uint32_t* oim, size_t N) {
  k_gt_reduce<LL><<<CFG>>>(re, im, Nin, nterms, ncoeff, G, ore, oim, N);
}
void fp2_from_bytes(LaunchCfg cfg, const uint8_t* in, int B, size_t count, uint32_t* re, uint32_t* im, size_t N) {
  k_fp2_from_bytes<LL><<<CFG>>>(in, B, count, re, im, N);
}
void fp2_to_bytes(LaunchCfg cfg, const uint32_t* re, const uint32_t* im, size_t N, size_t count, uint8_t* out, int B,
                  int grp, int pad) {
  k_fp2_to_bytes<LL><<<CFG>>>(re, im, N, count, out, B, grp, pad);
}
void bsgs_build(LaunchCfg cfg, const BsgsBuildArgs& a) { k_bsgs_build<LL><<<CFG>>>(a); }
void bsgs_lookup(LaunchCfg cfg, const BsgsLookupArgs& a) { k_bsgs_lookup<LL><<<CFG>>>(a); }
void mulmod_bench(LaunchCfg cfg, int ilp, uint32_t* io, size_t N, int iters) {
  if (ilp >= 10) {
    size_t smem = (size_t)13 * LL * 4 * cfg.block;
    cudaFuncSetAttribute(k_prim_bench<LL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_prim_bench<LL><<<cfg.grid, cfg.block, smem, cfg.stream>>>(io, N, iters, ilp);
  } else if (ilp == 2)
    k_mulmod_bench<LL, 2><<<CFG>>>(io, N, iters);
  else
    k_mulmod_bench<LL, 1><<<CFG>>>(io, N, iters);
}
const LOpsA ops = {LL,        upload,         miller_set_smem, miller_smem_bytes, miller_priv_bytes, miller_fixed_threads,
                   miller,     gt_mul,      gt_pow,
                   gt_reduce, fp2_from_bytes, fp2_to_bytes,    bsgs_build, bsgs_lookup, mulmod_bench};
}  // namespace
#define BGN_CAT2(a, b) a##b
#define BGN_CAT(a, b) BGN_CAT2(a, b)
extern "C" const LOpsA* BGN_CAT(bgn_opsA_, BGN_L)() { return &ops; }
