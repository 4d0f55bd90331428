// inst_b.cu -- G1 / (de)serialisation kernels for one limb count (compile with -DBGN_L=<L>).
#define BGN_GROUP_B 1
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
void g1_from_bytes(LaunchCfg cfg, const uint8_t* in, int B, size_t count, uint32_t* x, uint32_t* y, uint8_t* inf,
                   size_t N) {
  k_g1_from_bytes<LL><<<CFG>>>(in, B, count, x, y, inf, N);
}
void g1_to_bytes(LaunchCfg cfg, const uint32_t* x, const uint32_t* y, const uint8_t* inf, size_t N, size_t count,
                 uint8_t* out, int B) {
  k_g1_to_bytes<LL><<<CFG>>>(x, y, inf, N, count, out, B);
}
void encrypt(LaunchCfg cfg, const EncArgs& a) { k_encrypt<LL><<<CFG>>>(a); }
void normalize(LaunchCfg cfg, const NormArgs& a) { k_normalize<LL><<<CFG>>>(a); }
void g1_add(LaunchCfg cfg, const G1AddArgs& a) { k_g1_add<LL><<<CFG>>>(a); }
void g1_mulvar(LaunchCfg cfg, const G1MulArgs& a) { k_g1_mulvar<LL><<<CFG>>>(a); }
void tab_bases(LaunchCfg cfg, const uint32_t* bx, const uint32_t* by, int nwin, int hb, uint32_t* X, uint32_t* Y,
               uint32_t* Z, size_t N) {
  k_tab_bases<LL><<<CFG>>>(bx, by, nwin, hb, X, Y, Z, N);
}
void tab_fill(LaunchCfg cfg, const uint32_t* ax, const uint32_t* ay, const uint8_t* ainf, size_t Nb, int nwin, int hb,
              uint32_t* X, uint32_t* Y, uint32_t* Z, size_t N) {
  k_tab_fill<LL><<<CFG>>>(ax, ay, ainf, Nb, nwin, hb, X, Y, Z, N);
}
void tabw_fill(LaunchCfg cfg, const uint32_t* tabh, int nwin_h, int nsub, int hb, uint32_t* X, uint32_t* Y, uint32_t* Z,
               size_t first, size_t nent) {
  k_tabw_fill<LL><<<CFG>>>(tabh, nwin_h, nsub, hb, X, Y, Z, first, nent);
}
void tab_edwards(LaunchCfg cfg, const uint32_t* tabw, uint32_t* tabe, uint32_t* scratch, size_t count, int G, int* bad) {
  k_tab_edwards<LL><<<CFG>>>(tabw, tabe, scratch, count, G, bad);
}
void g1_polyconv(LaunchCfg cfg, const PolyConvArgs& a) { k_g1_polyconv<LL><<<CFG>>>(a); }
void g1_affadd(LaunchCfg cfg, const G1AffAddArgs& a) { k_g1_affadd<LL><<<CFG>>>(a); }
const LOpsB ops = {LL,     upload,    g1_from_bytes, g1_to_bytes, encrypt,    normalize,
                   g1_add, g1_mulvar, tab_bases,     tab_fill,    tabw_fill,  tab_edwards, g1_polyconv, g1_affadd};
}  // namespace
#define BGN_CAT2(a, b) a##b
#define BGN_CAT(a, b) BGN_CAT2(a, b)
extern "C" const LOpsB* BGN_CAT(bgn_opsB_, BGN_L)() { return &ops; }
