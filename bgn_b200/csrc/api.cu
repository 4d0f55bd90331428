// api.cu -- C-ABI (include/bgn_b200.h), context management and host-side
// orchestration of the kernels in kernels.cuh.  No CPU arithmetic fallback: the
// host only computes the handful of Montgomery constants a key needs (shifts,
// adds and compares on small big-integers) and drives the device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bgn_b200.h"
#include "ops.h"

// ---------------------------------------------------------------- host bigint
namespace {
typedef std::vector<uint32_t> Big;  // little-endian limbs, fixed length per use

Big big_from_be(const uint8_t* b, size_t len, size_t limbs) {
  Big r(limbs, 0);
  for (size_t i = 0; i < len; i++) {
    size_t pos = len - 1 - i;  // byte significance
    if (pos / 4 < limbs) r[pos / 4] |= (uint32_t)b[i] << (8 * (pos % 4));
  }
  return r;
}
int big_bits(const Big& a) {
  for (int i = (int)a.size() - 1; i >= 0; i--)
    if (a[i]) return 32 * i + 32 - __builtin_clz(a[i]);
  return 0;
}
int big_cmp(const Big& a, const Big& b) {
  for (int i = (int)a.size() - 1; i >= 0; i--)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return 0;
}
uint32_t big_add(Big& r, const Big& a, const Big& b) {
  uint64_t c = 0;
  for (size_t i = 0; i < r.size(); i++) {
    c += (uint64_t)a[i] + b[i];
    r[i] = (uint32_t)c;
    c >>= 32;
  }
  return (uint32_t)c;
}
void big_sub(Big& r, const Big& a, const Big& b) {
  int64_t c = 0;
  for (size_t i = 0; i < r.size(); i++) {
    c += (int64_t)a[i] - b[i];
    r[i] = (uint32_t)c;
    c >>= 32;
  }
}
// r = 2a mod p (a < p)
void big_dbl_mod(Big& a, const Big& p) {
  Big t(a.size());
  uint32_t carry = big_add(t, a, a);
  if (carry || big_cmp(t, p) >= 0) big_sub(t, t, p);
  a = t;
}
// signed-digit (NAF) expansion, most significant digit first
std::vector<int8_t> big_naf(Big n) {
  std::vector<int8_t> d;
  n.push_back(0);
  auto is_zero = [&]() {
    for (uint32_t w : n)
      if (w) return false;
    return true;
  };
  while (!is_zero()) {
    int8_t z = 0;
    if (n[0] & 1) {
      z = (int8_t)(2 - (int)(n[0] & 3));
      // n -= z
      if (z > 0) {
        uint64_t i = 0;
        while (n[i] == 0) n[i++] = 0xffffffffu;
        n[i] -= 1;
      } else {
        size_t i = 0;
        while (++n[i] == 0) i++;
      }
    }
    d.push_back(z);
    for (size_t i = 0; i + 1 < n.size(); i++) n[i] = (n[i] >> 1) | (n[i + 1] << 31);
    n.back() >>= 1;
  }
  std::reverse(d.begin(), d.end());
  return d;
}

// floor(a / b) for b != 0 (bitwise long division; key set-up only)
Big big_div(const Big& a, const Big& b) {
  size_t nl = a.size();
  Big q(nl, 0), r(nl + 1, 0), bb(nl + 1, 0);
  for (size_t i = 0; i < nl && i < b.size(); i++) bb[i] = b[i];
  for (int bit = big_bits(a) - 1; bit >= 0; bit--) {
    for (size_t i = nl; i > 0; i--) r[i] = (r[i] << 1) | (r[i - 1] >> 31);
    r[0] = (r[0] << 1) | ((a[bit >> 5] >> (bit & 31)) & 1u);
    if (big_cmp(r, bb) >= 0) {
      big_sub(r, r, bb);
      q[bit >> 5] |= 1u << (bit & 31);
    }
  }
  return q;
}

const int kSupportedL[] = {3, 5, 9, 17, 33};

// One lock per device: the kernels read their key material from __constant__ memory, of which there
// is one copy per device, so calls on contexts of the SAME device are serialised (thread-compatible,
// like the reference's pk.mu, bgn.go:40) while contexts on different devices run concurrently -- a
// single process may drive all GPUs of a box from one thread each.
constexpr int kMaxDevices = 64;
std::mutex g_dev_mu[kMaxDevices];
const bgn_ctx* g_active[kMaxDevices] = {};  // device -> context whose constants are resident (under its lock)
std::mutex& dev_mu(int device) { return g_dev_mu[(unsigned)device % kMaxDevices]; }
}  // namespace

// ---------------------------------------------------------------- context
struct KTime {
  double ms = 0;
  uint64_t launches = 0;
};

struct bgn_ctx {
  int device = 0;
  int L = 0, B = 0, nbytes = 0;
  std::vector<uint32_t> n_limbs;  // group order n, little-endian, BGN_MAXL limbs
  const LOpsA* A = nullptr;
  const LOpsB* Bo = nullptr;
  const LOpsC* Co = nullptr;
  const LOpsD* Do = nullptr;
  const LOpsE* Eo = nullptr;
  FieldConsts fc;
  PairConsts pc;
  cudaStream_t stream = nullptr;
  // generators (affine Montgomery, SoA with N = 1) and window tables
  uint32_t *dPx = nullptr, *dPy = nullptr, *dQx = nullptr, *dQy = nullptr;
  uint8_t *dPinf = nullptr, *dQinf = nullptr;
  uint32_t *tabP = nullptr, *tabQ = nullptr;
  uint32_t* tabPe = nullptr;   // tabP as twisted Edwards points (u | v | u v), 3L words per entry (curve.cuh: Ed)
  bool enc_edwards = true;     // Encrypt sums table points in Edwards form (option enc_edwards; cleared if P or Q has even order)
  bool tabQw_edw = false;      // tabQw holds Edwards points
  uint32_t* tabQw = nullptr;   // 16- or 24-bit windows of Q, built on the first randomised encryption
  int tabQw_bits = 0;
  uint32_t* tabE = nullptr;    // 8-bit windows of e(Q,Q) in GT, built on the first level-2 re-randomisation
  uint32_t* linesP = nullptr;  // line table of the Miller loop of P (MillerFixedArgs::lines), built on first use
  uint32_t* linesPq = nullptr; // line table of q1*P, built with the secret: level-1 Decrypt is e(C, q1 P) = e(C, P)^q1
  bool dec_pair_q1 = true;     // level-1 Decrypt through linesPq (option dec_pair_q1)
  int split_para = -1;         // k_miller_split with parabola steps: -1 / 1 on (default), 0 off (option split_para)
  bool fixed_lines = true;     // e(., P) through the line table (BGN_FIXED_LINES=0: the general kernel)
  int enc_window = 0;          // window bits of Q's table: 0 = widest of 16/18/20 within enc_table_max; 8 | 16 | 18 | 20 | 22 | 24 (BGN_ENC_WINDOW)
  size_t enc_table_max = (size_t)6 << 30;  // bound of the automatic choice (option enc_table_max_mb)
  int norm_per_thread = 8;     // lower bound of elements per inversion in k_normalize (BGN_NORM_PER_THREAD)
  int norm_threads = 148 * 256;  // threads k_normalize aims at (BGN_NORM_THREADS)
  bool affine_add = true;      // EAdd / ESub / Neg in affine coordinates with shared inversions (BGN_AFFINE_ADD=0: Jacobian + normalise)
  size_t pair_duo_cap = (size_t)-1;  // pairings one wave of k_pair_duo holds (occupancy query, cached)
  int miller_wide = 0;         // teams per block of k_miller_wide (0 = the 8-warp k_miller); BGN_MILLER_WIDE
  int miller_wide_min = 1;     // smallest batch that uses it
  int miller_split = -1;       // MultPoly below one wave on the split team kernel (teamsplit.cuh): -1 = by the time model, 0 = never, 1 = always (BGN_MILLER_SPLIT)
  int pair_duo_loop = -1;      // products' row loop of k_pair_duo: -1 = the key size's default, 0 / 1 / 2 / 4 (A/B knob)
  int pair_duo_pairs = 2;      // most warp pairs per block of k_pair_duo (measured: 2 beats 1 and 4 at 2^14 pairings)
  bool pair_duo_blockbar = true;  // blocks of several pairs synchronise as a whole (lockstep: one instruction stream per role)
  int pair_duo = -1;           // general pairings on two warps each (pairwarp.cuh): -1 = by batch size, 0 = never, 1 = always (BGN_PAIR_DUO)
  size_t fixed_pair_cap = 0;   // pairings one wave of k_miller_fixed_pair holds (occupancy query, cached)
  int fixed_pair = -1;         // e(., P) on a lane pair per point (pairlane.cuh): -1 = by batch size, 0 = never, 1 = always (BGN_FIXED_PAIR)
  bool dec_lucas = true;       // Decrypt through the Lucas ladder when one giant step suffices (BGN_DEC_LUCAS=0: off)
  // decryption
  bool has_secret = false;
  uint32_t *bs_elems = nullptr, *bs_slots = nullptr, *bs_ginv = nullptr;
  uint32_t bs_S = 0, bs_hmask = 0, bs_giant = 0;
  uint64_t bs_mmax = 0;
  // workspace arena (grow-only)
  uint8_t* arena = nullptr;
  size_t arena_cap = 0, arena_off = 0;
  // instrumentation
  bool timing = false;
  int miller_skew = 20000;  // cycles; see MillerArgs::skew_cycles
  int miller_groups = 0;    // 0 = choose per batch, 1 / 2 = force (tuning knob BGN_MILLER_GROUPS)
  int miller_tail = 0;      // block-size granularity of a separate launch for a partial last wave (0 = off)
  std::map<std::string, KTime> ktimes;
  struct Pending {
    std::string name;
    cudaEvent_t a, b;
  };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> ev_pool;
  uint64_t total_launches = 0;
  cudaEvent_t call_a = nullptr, call_b = nullptr;  // bracket the device work of the last C-ABI call
  double last_call_ms = 0;
  std::string err;
};

namespace {

struct CudaErr {
  std::string msg;
};
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      throw CudaErr{std::string(#call) + ": " + cudaGetErrorString(e_) + " @" + std::to_string(__LINE__)}; \
  } while (0)
struct ArgErr {
  std::string msg;
};

void activate(bgn_ctx* c) {
  CK(cudaSetDevice(c->device));
  if (g_active[c->device] != c) {
    CK(c->A->upload(&c->fc, &c->pc, c->stream));
    CK(c->Bo->upload(&c->fc, &c->pc, c->stream));
    CK(c->Co->upload(&c->fc, &c->pc, c->stream));
    CK(c->Do->upload(&c->fc, &c->pc, c->stream));
    CK(c->Eo->upload(&c->fc, &c->pc, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    g_active[c->device] = c;
  }
}
void reupload_pc(bgn_ctx* c) {
  CK(c->A->upload(&c->fc, &c->pc, c->stream));
  CK(c->Bo->upload(&c->fc, &c->pc, c->stream));
  CK(c->Co->upload(&c->fc, &c->pc, c->stream));
  CK(c->Do->upload(&c->fc, &c->pc, c->stream));
  CK(c->Eo->upload(&c->fc, &c->pc, c->stream));
  CK(cudaStreamSynchronize(c->stream));
}

// ---- arena
void arena_reset(bgn_ctx* c) { c->arena_off = 0; }
void arena_reserve(bgn_ctx* c, size_t bytes) {
  if (bytes <= c->arena_cap) return;
  CK(cudaStreamSynchronize(c->stream));
  if (c->arena) CK(cudaFree(c->arena));
  c->arena = nullptr;
  c->arena_cap = 0;
  size_t cap = bytes + bytes / 8 + (1 << 20);
  CK(cudaMalloc(&c->arena, cap));
  c->arena_cap = cap;
}
template <typename T>
T* arena_get(bgn_ctx* c, size_t n) {
  size_t bytes = (n * sizeof(T) + 255) & ~(size_t)255;
  if (c->arena_off + bytes > c->arena_cap) throw CudaErr{"internal: arena overflow"};
  T* p = reinterpret_cast<T*>(c->arena + c->arena_off);
  c->arena_off += bytes;
  return p;
}
size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

bool is_device_ptr(const void* p) {
  cudaPointerAttributes at;
  cudaError_t e = cudaPointerGetAttributes(&at, p);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ---- instrumented launch
struct Timer {
  bgn_ctx* c;
  bool on;
  cudaEvent_t a = nullptr, b = nullptr;
  std::string name;
  Timer(bgn_ctx* c_, const char* name_) : c(c_), on(c_->timing), name(name_) {
    c->total_launches++;
    if (!on) return;
    auto get = [&]() {
      cudaEvent_t e;
      if (!c->ev_pool.empty()) {
        e = c->ev_pool.back();
        c->ev_pool.pop_back();
      } else {
        CK(cudaEventCreate(&e));
      }
      return e;
    };
    a = get();
    b = get();
    CK(cudaEventRecord(a, c->stream));
  }
  void done() {
    CK(cudaGetLastError());
    if (!on) return;
    CK(cudaEventRecord(b, c->stream));
    c->pending.push_back({name, a, b});
  }
};
void finish(bgn_ctx* c) {
  CK(cudaStreamSynchronize(c->stream));
  for (auto& p : c->pending) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, p.a, p.b));
    KTime& k = c->ktimes[p.name];
    k.ms += ms;
    k.launches++;
    c->ev_pool.push_back(p.a);
    c->ev_pool.push_back(p.b);
  }
  c->pending.clear();
}

LaunchCfg cfg(bgn_ctx* c, size_t grid, unsigned block, size_t smem = 0) {
  return LaunchCfg{(unsigned)grid, block, smem, c->stream};
}
unsigned nblk(size_t n, int t) { return (unsigned)((n + t - 1) / t); }

// ---- device arrays
struct G1Arr {
  uint32_t *x = nullptr, *y = nullptr;
  uint8_t* inf = nullptr;
  size_t N = 0;
};
struct JacArr {
  uint32_t *X = nullptr, *Y = nullptr, *Z = nullptr;
  size_t N = 0;
};
struct GtArr {
  uint32_t *re = nullptr, *im = nullptr;
  size_t N = 0;
};
size_t g1_bytes(bgn_ctx* c, size_t n) { return 2 * pad256(n * c->L * 4) + pad256(n); }
size_t jac_bytes(bgn_ctx* c, size_t n) { return 3 * pad256(n * c->L * 4); }
size_t gt_bytes(bgn_ctx* c, size_t n) { return 2 * pad256(n * c->L * 4); }
size_t io_bytes(bgn_ctx* c, size_t n) { return pad256(n * 2 * c->B); }

G1Arr g1_alloc(bgn_ctx* c, size_t n) {
  G1Arr a;
  a.N = n;
  a.x = arena_get<uint32_t>(c, n * c->L);
  a.y = arena_get<uint32_t>(c, n * c->L);
  a.inf = arena_get<uint8_t>(c, n);
  return a;
}
JacArr jac_alloc(bgn_ctx* c, size_t n) {
  JacArr a;
  a.N = n;
  a.X = arena_get<uint32_t>(c, n * c->L);
  a.Y = arena_get<uint32_t>(c, n * c->L);
  a.Z = arena_get<uint32_t>(c, n * c->L);
  return a;
}
GtArr gt_alloc(bgn_ctx* c, size_t n) {
  GtArr a;
  a.N = n;
  a.re = arena_get<uint32_t>(c, n * c->L);
  a.im = arena_get<uint32_t>(c, n * c->L);
  return a;
}

// stage a caller buffer on the device (no copy if it already is a device pointer)
const uint8_t* stage_in(bgn_ctx* c, const void* p, size_t bytes) {
  if (bytes == 0) return nullptr;
  if (is_device_ptr(p)) return static_cast<const uint8_t*>(p);
  uint8_t* d = arena_get<uint8_t>(c, bytes);
  CK(cudaMemcpyAsync(d, p, bytes, cudaMemcpyHostToDevice, c->stream));
  return d;
}
struct OutBuf {
  uint8_t* dev;
  void* host;  // null when the caller gave a device pointer
  size_t bytes;
};
OutBuf stage_out(bgn_ctx* c, void* p, size_t bytes) {
  OutBuf o{nullptr, nullptr, bytes};
  if (bytes == 0) return o;
  if (is_device_ptr(p)) {
    o.dev = static_cast<uint8_t*>(p);
  } else {
    o.dev = arena_get<uint8_t>(c, bytes);
    o.host = p;
  }
  return o;
}
void commit_out(bgn_ctx* c, const OutBuf& o) {
  if (o.host && o.bytes) CK(cudaMemcpyAsync(o.host, o.dev, o.bytes, cudaMemcpyDeviceToHost, c->stream));
}

// ---- kernel wrappers
// shared memory of the (de)serialisation kernels: the block's elements in byte format, word-padded
size_t io_smem(bgn_ctx* c, int nt) { return ((size_t)nt * 2 * c->B + 15) & ~(size_t)15; }
void g1_from_bytes(bgn_ctx* c, const uint8_t* d_in, size_t count, const G1Arr& a) {
  if (!count) return;
  Timer t(c, "k_g1_from_bytes");
  c->Bo->g1_from_bytes(cfg(c, nblk(count, 128), 128, io_smem(c, 128)), d_in, c->B, count, a.x, a.y, a.inf, a.N);
  t.done();
}
void g1_to_bytes(bgn_ctx* c, const G1Arr& a, size_t count, uint8_t* d_out) {
  if (!count) return;
  Timer t(c, "k_g1_to_bytes");
  c->Bo->g1_to_bytes(cfg(c, nblk(count, 128), 128, io_smem(c, 128)), a.x, a.y, a.inf, a.N, count, d_out, c->B);
  t.done();
}
void gt_from_bytes(bgn_ctx* c, const uint8_t* d_in, size_t count, const GtArr& a) {
  if (!count) return;
  Timer t(c, "k_fp2_from_bytes");
  c->A->fp2_from_bytes(cfg(c, nblk(count, 128), 128, io_smem(c, 128)), d_in, c->B, count, a.re, a.im, a.N);
  t.done();
}
// grp > 0: the output is polynomials of grp + pad slots, the pad slots written as the GT identity
void gt_to_bytes(bgn_ctx* c, const GtArr& a, size_t count, uint8_t* d_out, int grp = 0, int pad = 0) {
  if (!count) return;
  Timer t(c, "k_fp2_to_bytes");
  c->A->fp2_to_bytes(cfg(c, nblk(count, 128), 128, io_smem(c, 128)), a.re, a.im, a.N, count, d_out, c->B, grp, pad);
  t.done();
}

// Threads for a kernel whose threads share one inversion among their elements (k_normalize,
// k_g1_affadd).  The binary-GCD inversion costs about as much as 85 products and each element ~6, so
// ~10 elements per thread already amortise it; fewer elements per thread mean more threads, which is
// what a latency-bound chain needs.  Aim at two warps per scheduler (148 x 256 threads), between 8 and
// 64 elements per thread; small batches favour parallelism over shared inversions.
size_t shared_inversion_threads(bgn_ctx* c, size_t count) {
  size_t per = std::min<size_t>(64, std::max<size_t>(c->norm_per_thread, (count + c->norm_threads - 1) / c->norm_threads));
  size_t G = (count + per - 1) / per;
  if (G < 4096) G = std::min<size_t>(count, 4096);
  return G;
}

// Jacobian -> affine; output either SoA (G1Arr) or AoS table
void normalize(bgn_ctx* c, const JacArr& j, size_t count, uint32_t* scratch, uint32_t* ox, uint32_t* oy,
               size_t estride, size_t lstride, uint8_t* inf) {
  if (!count) return;
  NormArgs a;
  a.X = j.X;
  a.Y = j.Y;
  a.Z = j.Z;
  a.scratch = scratch;
  a.count = count;
  a.N = j.N;
  size_t G = shared_inversion_threads(c, count);
  a.G = (int)G;
  a.ox = ox;
  a.oy = oy;
  a.o_estride = estride;
  a.o_lstride = lstride;
  a.inf = inf;
  Timer t(c, "k_normalize");
  c->Bo->normalize(cfg(c, nblk(G, 128), 128, 0), a);
  t.done();
}
void normalize_soa(bgn_ctx* c, const JacArr& j, size_t count, uint32_t* scratch, const G1Arr& o) {
  normalize(c, j, count, scratch, o.x, o.y, (size_t)c->L, 1, o.inf);
}

// global scratch the Miller kernel needs for `count` units with teams of dE threads (0 unless
// the key's limb count uses the interleaved layout); callers add it to their arena reservation
size_t miller_scratch(bgn_ctx* c, size_t count, int dE) {
  size_t pb = c->A->miller_priv_bytes();
  int fixed = c->A->miller_fixed_threads();
  if (!count || dE <= 0) return 0;
  // x^2 / y of every evaluation point for the parabola steps (MillerArgs::evw); a batch may be served by several
  // launches (full waves, remainder, balanced split waves), each with its own padded slice
  const size_t evw = pad256(count * (size_t)dE * c->L * 4) + 16384;
  if (!pb || !fixed) return evw;
  if (dE > fixed) return 0;
  size_t upb = (size_t)(fixed / dE);
  return evw + pad256(((count + upb - 1) / upb + 160) * pb) + 256;  // + one block per SM: small batches are spread out
}

// ---- the split team kernel (teamsplit.cuh): two threads per output-slot pair, for sub-wave batches
struct SplitGeom {
  int tpb = 0;       // teams (units) per block at most
  size_t cap = 0;    // units one wave holds
};
SplitGeom miller_split_geom(bgn_ctx* c, int dM, int dE) {
  SplitGeom g;
  if (!c->Eo->miller_split || dM < 2 || dE > 192) return g;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  int tpb = 192 / dE;  // each half of a block (one thread role) is at most 6 warps
  while (tpb > 0) {
    int nt = 2 * ((tpb * dE + 31) / 32 * 32);
    if (c->Eo->miller_split_smem_bytes(nt, tpb * dE) + 16 <= 227 * 1024 - 64) break;
    tpb--;
  }
  g.tpb = tpb;
  g.cap = (size_t)sms * tpb;
  return g;
}
// one launch: `count` units spread over the SMs, at most g.tpb per block
void launch_miller_split(bgn_ctx* c, const SplitGeom& g, const G1Arr& M, int dM, const G1Arr& E, int dE, int e_bcast,
                         size_t count, int out_slots, const GtArr& out) {
  if (!count) return;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  int tpb = (int)std::min<size_t>((size_t)g.tpb, std::max<size_t>(1, (count + sms - 1) / sms));
  int nt = 2 * ((tpb * dE + 31) / 32 * 32);  // two halves of whole warps: one thread role per warp
  size_t smem = c->Eo->miller_split_smem_bytes(nt, tpb * dE) + 16;
  MillerArgs a;
  memset(&a, 0, sizeof(a));
  a.Mx = M.x;
  a.My = M.y;
  a.Minf = M.inf;
  a.NM = (int)M.N;
  a.Ex = E.x;
  a.Ey = E.y;
  a.Einf = E.inf;
  a.NE = (int)E.N;
  a.e_bcast = e_bcast;
  a.priv = nullptr;
  a.out_re = out.re;
  a.out_im = out.im;
  a.NOUT = (int)out.N;
  a.dM = dM;
  a.dE = dE;
  a.out_slots = out_slots;
  a.count = (int)count;
  a.teams_per_group = tpb;
  a.group_threads = nt;
  a.evw = arena_get<uint32_t>(c, (e_bcast ? (size_t)dE : count * (size_t)dE) * c->L);
  // parabola steps: faster at every block geometry since the merged step is a fused routine (measured,
  // profiles/r02_split_para_ab.json: 1 - 6 warps per role gain 4 - 11 %); split_para = 0 is the A/B switch
  a.para = c->split_para >= 0 ? c->split_para : 1;
  Timer t(c, "k_miller_split");
  CK(c->Eo->miller_split_set_smem(smem));
  c->Eo->miller_split(cfg(c, (count + tpb - 1) / tpb, nt, smem), a);
  t.done();
}

// the Miller team kernel; dM <= dE
void run_miller(bgn_ctx* c, const G1Arr& M, int dM, const G1Arr& E, int dE, int e_bcast, size_t count, int out_slots,
                const GtArr& out, bool is_tail = false) {
  if (!count) return;
  int TS = dE;
  const size_t smem_max = 227 * 1024 - 64;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  int groups = 1, tpg = 0, GT_ = 0;
  size_t smem = 0;
  if (int fixed = c->A->miller_fixed_threads()) {
    // interleaved layout (1024-bit field): blockDim is the layout's thread stride; the private
    // slots of every block live in a global scratch array (L2-resident)
    if (TS > fixed) throw ArgErr{"polynomial has too many coefficients for one thread group"};
    GT_ = fixed;
    tpg = fixed / TS;
    smem = c->A->miller_smem_bytes(fixed) + 16;
    if (smem > smem_max) throw ArgErr{"field too large for the Miller kernel's shared-memory state"};
  } else {
    size_t per_thread = c->A->miller_smem_bytes(256) / 256;  // bytes per thread (slots + 2 flag bytes)
    // A block is either ONE barrier group of up to 256 threads or TWO independent groups of 128
    // (whole teams per group).  Pick whichever needs less time for this batch:
    // waves x time-per-wave, waves = ceil(units / (SMs x blocks/SM x units/block)).
    if (TS > 128) throw ArgErr{"polynomial has too many coefficients for one thread group"};
    if (128 * per_thread + 16 > smem_max) throw ArgErr{"field too large for the Miller kernel's shared-memory state"};
    int nt1 = (int)std::min<size_t>(256, (smem_max - 16) / per_thread) / 32 * 32;  // one group: whole warps
    int tpg1 = nt1 / TS;
    if (tpg1 == 0) throw ArgErr{"polynomial has too many coefficients for the shared-memory state at this key size"};
    bool can2 = 2 * 128 * per_thread + 16 <= smem_max;
    int tpg2 = 128 / TS;
    auto cost = [&](int ngroups, int tpgx, int nt, double twave) {
      size_t bps = std::max<size_t>(1, std::min<size_t>(8, smem_max / (per_thread * nt + 16)));  // blocks per SM
      size_t per_wave = (size_t)sms * bps * ngroups * tpgx;
      return (double)((count + per_wave - 1) / per_wave) * twave;
    };
    tpg = tpg1;
    GT_ = nt1;
    if (can2 && c->miller_groups == 2 && cost(2, tpg2, 256, 1.0) <= cost(1, tpg1, nt1, 1.0)) {
      groups = 2;  // tuning knob only: with the fused routines two groups no longer beat one
      tpg = tpg2;
      GT_ = 128;
    }
    smem = c->A->miller_smem_bytes(groups * GT_) + 16;
  }
  // The wide team kernel (10 warps per SM: evaluation points read from the batch arrays, pairing.cuh:
  // MillerTeam<L, true>) for batches of several waves: miller_wide = teams per block (A/B knob; 0 = off)
  if (groups == 1 && !is_tail && !c->A->miller_fixed_threads() && c->miller_wide > 0 && c->Eo->miller_wide &&
      count >= (size_t)c->miller_wide_min) {
    int tpbw = std::min(c->miller_wide, 320 / TS);
    int ntw = (tpbw * TS + 31) / 32 * 32;
    size_t smw = c->Eo->miller_wide_smem_bytes(ntw) + 16;
    if (tpbw > 0 && smw <= smem_max) {
      MillerArgs a;
      memset(&a, 0, sizeof(a));
      a.Mx = M.x;
      a.My = M.y;
      a.Minf = M.inf;
      a.NM = (int)M.N;
      a.Ex = E.x;
      a.Ey = E.y;
      a.Einf = E.inf;
      a.NE = (int)E.N;
      a.e_bcast = e_bcast;
      a.out_re = out.re;
      a.out_im = out.im;
      a.NOUT = (int)out.N;
      a.dM = dM;
      a.dE = dE;
      a.out_slots = out_slots;
      a.count = (int)count;
      a.teams_per_group = tpbw;
      a.group_threads = ntw;
      Timer t(c, "k_miller_wide");
      CK(c->Eo->miller_wide_set_smem(smw));
      c->Eo->miller_wide(cfg(c, (count + tpbw - 1) / tpbw, ntw, smw), a);
      t.done();
      return;
    }
  }
  // Sub-wave work -- a whole batch below one wave, or the remainder after the full waves -- goes to the
  // split kernel (two threads per output-slot pair: 0.6 of the dependent chain per Miller step) when it
  // fits ONE wave of that kernel; several balanced split waves replace a nearly empty last wave pair.
  // Policy constants measured: profiles/r02_split_ab.json.
  if (groups == 1 && !is_tail && !c->A->miller_fixed_threads() && c->miller_split != 0) {
    SplitGeom g = miller_split_geom(c, dM, dE);
    const size_t cap_std = (size_t)sms * tpg;
    if (g.cap) {
      auto g1_off = [&](const G1Arr& a, size_t n) { return G1Arr{a.x + n * c->L, a.y + n * c->L, a.inf + n, a.N - n}; };
      auto gt_off = [&](size_t units) { return GtArr{out.re + units * out_slots * c->L, out.im + units * out_slots * c->L, out.N - units * out_slots}; };
      const size_t full = count / cap_std, rem = count % cap_std;
      // Time model in units of one full wave of k_miller (measured at 11 x 11 slots, 17 limbs:
      // profiles/r02_split_para_ab.json, r02_split_ab_v7.json).  A split wave's time depends on the warps each
      // thread role occupies per SM: 1 - 2: 0.56, 3 - 4: 0.63, 5: 0.90, 6 (full): 0.84.  The team
      // kernel: a batch within half a wave (one warp per scheduler) 0.76, anything else whole waves.
      auto t_split = [&](size_t n) {
        const size_t u = (n + sms - 1) / sms, w = (u * (size_t)dE + 31) / 32;
        static const double tw[7] = {0.56, 0.56, 0.56, 0.63, 0.63, 0.90, 0.84};
        return tw[std::min<size_t>(w, 6)];
      };
      const double t_a = count > cap_std ? (double)((count + cap_std - 1) / cap_std) : (count * 2 <= cap_std ? 0.76 : 1.0);
      const double t_b = rem && rem <= g.cap ? (double)full + t_split(rem) : 1e30;            // full waves + split remainder
      const size_t kw = (count + g.cap - 1) / g.cap;
      const double t_c = (double)kw * t_split((count + kw - 1) / kw);                          // balanced split waves
      int choice = 0;
      if (c->miller_split > 0) choice = (count <= g.cap || !full) ? 2 : 1;
      else if (t_b < t_a && t_b <= t_c) choice = 1;
      else if (t_c < t_a) choice = 2;
      if (choice == 1 && rem && rem <= g.cap) {
        if (full) run_miller(c, M, dM, E, dE, e_bcast, full * cap_std, out_slots, out, true);
        launch_miller_split(c, g, g1_off(M, full * cap_std * dM), dM, e_bcast ? E : g1_off(E, full * cap_std * dE), dE, e_bcast, rem,
                            out_slots, gt_off(full * cap_std));
        return;
      }
      if (choice == 2) {
        size_t done = 0;
        for (size_t w = 0; w < kw; w++) {
          size_t n = (count - done + (kw - w) - 1) / (kw - w);
          launch_miller_split(c, g, g1_off(M, done * dM), dM, e_bcast ? E : g1_off(E, done * dE), dE, e_bcast, n, out_slots, gt_off(done));
          done += n;
        }
        return;
      }
    }
  }
  // A batch smaller than one full wave: use the fewest warps per scheduler that still fit the batch
  // in one wave, in blocks of whole scheduler rounds (128 threads = one warp on each of the four
  // schedulers).  Measured: 1 warp per scheduler finishes ~25 % sooner than 2, while an unbalanced
  // 7-warp block is SLOWER than an 8-warp one (profiles/r01_ops1024_v13.json vs v12).
  // Several waves with a partial last one: run the full waves, then the remainder as its own launch
  // spread over all SMs (BGN_MILLER_TAIL: granularity of that launch's block size in threads, 0 = off).
  if (groups == 1 && c->miller_tail > 0 && !is_tail && count > (size_t)sms * tpg && count % ((size_t)sms * tpg) != 0) {
    size_t full = count / ((size_t)sms * tpg) * ((size_t)sms * tpg), rem = count - full;
    auto g1_off = [&](const G1Arr& a, size_t n) { return G1Arr{a.x + n * c->L, a.y + n * c->L, a.inf + n, a.N - n}; };
    GtArr o2{out.re + full * out_slots * c->L, out.im + full * out_slots * c->L, out.N - full * out_slots};
    run_miller(c, M, dM, E, dE, e_bcast, full, out_slots, out, false);
    run_miller(c, g1_off(M, full * dM), dM, e_bcast ? E : g1_off(E, full * dE), dE, e_bcast, rem, out_slots, o2, true);
    return;
  }
  if (groups == 1 && count < (size_t)sms * tpg) {
    size_t thr_per_sm = (count * TS + sms - 1) / sms;
    size_t gran = is_tail ? (size_t)c->miller_tail : 128;
    int ntM = (int)std::min<size_t>(GT_, std::max<size_t>((thr_per_sm + gran - 1) / gran * gran, (size_t)(TS + 31) / 32 * 32));
    tpg = std::max(1, ntM / TS);
    GT_ = ntM;
    if (!c->A->miller_fixed_threads()) smem = c->A->miller_smem_bytes(GT_) + 16;
  }
  size_t units_per_block = (size_t)groups * tpg;
  int nt = groups * GT_;
  size_t nblocks = (count + units_per_block - 1) / units_per_block;
  uint32_t* priv = nullptr;
  if (size_t pb = c->A->miller_priv_bytes()) priv = reinterpret_cast<uint32_t*>(arena_get<uint8_t>(c, nblocks * pb));  // see miller_scratch()
  MillerArgs a;
  a.Mx = M.x;
  a.My = M.y;
  a.Minf = M.inf;
  a.NM = (int)M.N;
  a.Ex = E.x;
  a.Ey = E.y;
  a.Einf = E.inf;
  a.NE = (int)E.N;
  a.e_bcast = e_bcast;
  a.priv = priv;
  a.out_re = out.re;
  a.out_im = out.im;
  a.NOUT = (int)out.N;
  a.dM = dM;
  a.dE = dE;
  a.out_slots = out_slots;
  a.count = (int)count;
  a.teams_per_group = tpg;
  a.group_threads = GT_;
  a.skew_cycles = c->miller_skew;
  a.evw = arena_get<uint32_t>(c, (e_bcast ? (size_t)dE : count * (size_t)dE) * c->L);
  a.para = dE >= 3 ? 1 : 0;
  Timer t(c, "k_miller");
  CK(c->A->miller_set_smem(smem));
  c->A->miller(cfg(c, nblocks, nt, smem), a);
  t.done();
}

// out[i] = e(A[i], B[i]) with two warps per 32 pairings (pairwarp.cuh)
int pair_duo_variant(bgn_ctx* c, int pairs_per_block) {
  // measured (profiles/r02_duo_knobs.json): at 17 limbs the 4-rows-per-iteration loop beats the unrolled
  // products (17.8 against 18.3 ms at 2^14 pairings, 15.0 against 17.9 at 2^12): the two roles run different
  // code, and the smaller bodies ease the instruction cache
  int loop = c->pair_duo_loop >= 0 ? c->pair_duo_loop : (c->L == 17 ? 2 : (c->L < 17 ? 0 : 4));
  return loop + ((c->pair_duo_blockbar && pairs_per_block > 1) ? 8 : 0);
}
void run_pair_duo(bgn_ctx* c, const G1Arr& A, const G1Arr& Bv, size_t count, const GtArr& out) {
  if (!count) return;
  PairDuoArgs a;
  a.Mx = A.x;
  a.My = A.y;
  a.Minf = A.inf;
  a.Ex = Bv.x;
  a.Ey = Bv.y;
  a.Einf = Bv.inf;
  a.out_re = out.re;
  a.out_im = out.im;
  a.count = (int)count;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  // warp pairs per block: the fewest that keep the batch within one block per SM (small batches are
  // spread pair by pair), at most pair_duo_pairs
  int pairs = (int)std::min<size_t>((size_t)c->pair_duo_pairs, std::max<size_t>(1, (count + (size_t)sms * 32 - 1) / ((size_t)sms * 32)));
  while (pairs > 1 && c->Do->pair_duo_smem_bytes(32 * pairs) + 16 > 227 * 1024 - 64) pairs--;
  const int np = 32 * pairs;
  const int variant = pair_duo_variant(c, pairs);
  size_t smem = c->Do->pair_duo_smem_bytes(np) + 16;
  Timer t(c, "k_pair_duo");
  CK(c->Do->pair_duo_set_smem(variant, smem));
  if (!c->Do->pair_duo(variant, cfg(c, nblk(count, np), 2 * np, smem), a)) throw ArgErr{"pair_duo variant not built for this key size"};
  t.done();
}
// pairings one wave of k_pair_duo holds (0: the kernel does not fit this key's shared-memory state)
size_t pair_duo_capacity(bgn_ctx* c) {
  if (c->pair_duo_cap == (size_t)-1) {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
    int pairs = c->pair_duo_pairs;
    while (pairs > 1 && c->Do->pair_duo_smem_bytes(32 * pairs) + 16 > 227 * 1024 - 64) pairs--;
    size_t smem = c->Do->pair_duo_smem_bytes(32 * pairs) + 16;
    const int variant = pair_duo_variant(c, pairs);
    c->pair_duo_cap = 0;
    if (smem <= 227 * 1024 - 64 && c->Do->pair_duo_set_smem(variant, smem) == cudaSuccess)
      c->pair_duo_cap = (size_t)sms * std::max(0, c->Do->pair_duo_blocks_per_sm(variant, 64 * pairs, smem)) * 32 * pairs;
    cudaGetLastError();
  }
  return c->pair_duo_cap;
}

// lines of the Miller loop of P, recorded once per key (k_miller_record: point arithmetic only)
int miller_nsteps(const bgn_ctx* c) {
  int n = 0;
  for (int idx = 1; idx < c->pc.naf_len; idx++) {
    n++;
    if (c->pc.naf[idx] != 0 && idx != c->pc.naf_len - 1) n++;
  }
  return n;
}
// normalised line table of the affine point (px, py) (device, Montgomery); nullptr if it has none
uint32_t* record_lines(bgn_ctx* c, const uint32_t* px, const uint32_t* py) {
  const size_t words = (size_t)miller_nsteps(c) * 2 * c->L;
  uint32_t *tab = nullptr, *scratch = nullptr;
  int ok = 0;
  try {
    CK(cudaMalloc(&tab, words * 4));
    CK(cudaMalloc(&scratch, words * 4 + 256));
    int* dok = reinterpret_cast<int*>(scratch + words);
    Timer t(c, "k_miller_record");
    c->Co->miller_record(cfg(c, 1, 32, 0), px, py, tab, scratch, dok);
    t.done();
    CK(cudaMemcpyAsync(&ok, dok, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    finish(c);
  } catch (...) {
    cudaFree(tab);
    cudaFree(scratch);
    throw;
  }
  cudaFree(scratch);
  if (!ok) {
    cudaFree(tab);
    return nullptr;
  }
  return tab;
}
void ensure_linesP(bgn_ctx* c) {
  if (c->linesP || !c->fixed_lines) return;
  c->linesP = record_lines(c, c->dPx, c->dPy);
  // P is not a point of odd order: no normalised table; e(., P) goes through the general kernel
  if (!c->linesP) c->fixed_lines = false;
}
// out[i] = e(E[i], P) through the line table
void run_miller_fixed(bgn_ctx* c, const G1Arr& E, size_t count, const GtArr& out, const uint32_t* lines = nullptr) {
  if (!count) return;
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
  MillerFixedArgs a;
  a.lines = lines ? lines : c->linesP;
  a.Ex = E.x;
  a.Ey = E.y;
  a.Einf = E.inf;
  a.out_re = out.re;
  a.out_im = out.im;
  a.count = (int)count;
  // Which mapping (measured A/B: profiles/r02_fixedpair_ab_512.json, _1024.json): a lane pair per pairing
  // (pairlane.cuh, registers only) is the faster one at every batch size -- 1.8x for a single pairing,
  // 0.77 against 0.53 of the IMAD.WIDE peak at 2^14, 0.91 against 0.78 at 2^17 -- except where the batch
  // is just over one wave of the pair kernel but still fits ONE wave of the one-thread kernel (148 x 256
  // threads); there the second, nearly empty wave of the pair kernel costs more than it saves.
  if (c->fixed_pair_cap == 0) c->fixed_pair_cap = (size_t)sms * std::max(1, c->Do->miller_fixed_pair_blocks_per_sm()) * 32;
  const bool pair_auto = count <= c->fixed_pair_cap || count > (size_t)sms * 256;
  if (c->fixed_pair > 0 || (c->fixed_pair < 0 && pair_auto)) {
    Timer t(c, "k_miller_fixed_pair");
    c->Do->miller_fixed_pair(cfg(c, nblk(2 * count, 64), 64, 0), a);
    t.done();
    return;
  }
  // one block per SM up to 256 threads; smaller batches use whole scheduler rounds (128 threads)
  size_t per_sm = (count + sms - 1) / sms;
  int nt = (int)std::min<size_t>(256, std::max<size_t>(128, (per_sm + 127) / 128 * 128));
  size_t smem = c->Co->miller_fixed_smem_bytes(nt);
  while (smem > 227 * 1024 - 64 && nt > 32) {
    nt -= 32;
    smem = c->Co->miller_fixed_smem_bytes(nt);
  }
  Timer t(c, "k_miller_fixed");
  CK(c->Co->miller_fixed_set_smem(smem));
  c->Co->miller_fixed(cfg(c, nblk(count, nt), nt, smem), a);
  t.done();
}

void check_count(size_t count) {
  if (count > ((size_t)1 << 27)) throw ArgErr{"batch too large (max 2^27 elements per call)"};
}

// Table of nwin windows of hb bits for the base point (bx, by) (device, Montgomery, N = 1): entry (w, d),
// d = 1 .. 2^hb - 1, is d * 2^(hb w) * base, affine x || y.  The caller provides the temporaries: Jacobian
// and affine arrays of nwin points for the window bases, a Jacobian array of nwin (2^hb - 1) points and
// as many field elements of scratch.
void build_table_with(bgn_ctx* c, const uint32_t* bx, const uint32_t* by, int nwin, int hb, uint32_t* tab, const JacArr& jb,
                      const G1Arr& ab, const JacArr& je, uint32_t* scratch) {
  size_t nent = (size_t)nwin * (((size_t)1 << hb) - 1);
  {
    Timer t(c, "k_tab_bases");
    c->Bo->tab_bases(cfg(c, 1, 32, 0), bx, by, nwin, hb, jb.X, jb.Y, jb.Z, jb.N);
    t.done();
  }
  normalize_soa(c, jb, nwin, scratch, ab);
  {
    Timer t(c, "k_tab_fill");
    c->Bo->tab_fill(cfg(c, nblk(nwin, 32), 32, 0), ab.x, ab.y, ab.inf, ab.N, nwin, hb, je.X, je.Y, je.Z, je.N);
    t.done();
  }
  normalize(c, je, nent, scratch, tab, tab + c->L, 2 * (size_t)c->L, 1, nullptr);
}
// the 255-entry-per-window tables of P and Q built with the key (temporaries from the arena)
void build_table(bgn_ctx* c, const uint32_t* bx, const uint32_t* by, int nwin, uint32_t* tab) {
  size_t nent = (size_t)nwin * 255;
  arena_reset(c);
  arena_reserve(c, jac_bytes(c, nwin) + g1_bytes(c, nwin) + jac_bytes(c, nent) + 2 * pad256(nent * c->L * 4) + 4096);
  JacArr jb = jac_alloc(c, nwin);
  G1Arr ab = g1_alloc(c, nwin);
  JacArr je = jac_alloc(c, nent);
  uint32_t* scratch = arena_get<uint32_t>(c, nent * c->L);
  build_table_with(c, bx, by, nwin, 8, tab, jb, ab, je, scratch);
  finish(c);
}

// Wide-window table of Q for Encrypt, sized for HBM rather than shared memory: ceil(8 nbytes / w) windows of
// w bits, 2^w - 1 affine points each.  At 512-bit keys: w = 16: 285 MB, 32 additions per encryption;
// w = 20: 3.7 GB, 26; w = 22: 13.7 GB, 24; w = 24: 50 GB, 22.  A window is built as nsub windows of hb
// bits (16 = 2 x 8, 18 = 2 x 9, 20 = 2 x 10, 22 = 2 x 11, 24 = 3 x 8) from a narrow table -- the 8-bit
// table of the key, or a temporary one -- in chunks whose temporaries are released again.  enc_window = 0
// (the default) takes the widest of 16 / 18 / 20 bits whose table stays within enc_table_max bytes
// (4 GiB); a table that does not fit the free memory falls back to 16 bits.
size_t tabQw_bytes(const bgn_ctx* c, int bits) {
  const size_t per_entry = (c->enc_edwards && c->tabPe ? 3 : 2) * (size_t)c->L * 4;
  return (size_t)((8 * c->nbytes + bits - 1) / bits) * (((size_t)1 << bits) - 1) * per_entry;
}
// Weierstrass table (x || y per entry) -> Edwards table (u || v || u v); returns false if some finite entry
// has no Edwards image (the base point is not of odd order)
bool table_to_edwards(bgn_ctx* c, const uint32_t* tabw, uint32_t* tabe, uint32_t* scratch, int* dbad, size_t count) {
  if (!count) return true;
  CK(cudaMemsetAsync(dbad, 0, sizeof(int), c->stream));
  size_t G = shared_inversion_threads(c, count);
  {
    Timer t(c, "k_tab_edwards");
    c->Bo->tab_edwards(cfg(c, nblk(G, 128), 128, 0), tabw, tabe, scratch, count, (int)G, dbad);
    t.done();
  }
  int bad = 0;
  CK(cudaMemcpyAsync(&bad, dbad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  finish(c);
  return bad == 0;
}
int enc_window_auto(const bgn_ctx* c) {
  for (int bits = 20; bits > 16; bits -= 2)
    if (tabQw_bytes(c, bits) <= c->enc_table_max) return bits;
  return 16;
}
void ensure_tabQw(bgn_ctx* c) {
  if (c->tabQw || c->enc_window == 8) return;
  int wbits = c->enc_window ? c->enc_window : enc_window_auto(c);
  const bool edw = c->enc_edwards && c->tabPe;
  size_t ew = (size_t)c->L * 4;
  const size_t chunk_max = (size_t)1 << 22;
  if (wbits > 16) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    // leave room for the callers' batches next to the table (a quarter of it, at least 1 GiB, at most 16)
    size_t room = std::min<size_t>((size_t)16 << 30, std::max<size_t>((size_t)1 << 30, tabQw_bytes(c, wbits) / 4));
    if (tabQw_bytes(c, wbits) + chunk_max * 6 * ew + room > free_b) wbits = 16;
  }
  const int nsub = wbits % 8 == 0 ? wbits / 8 : 2;
  const int hb = wbits / nsub;
  const int nwin = (8 * c->nbytes + wbits - 1) / wbits;
  const size_t ents = ((size_t)1 << wbits) - 1;
  const size_t nent = (size_t)nwin * ents;
  const size_t chunk = std::min(nent, chunk_max);
  uint32_t *tab = nullptr, *tmp = nullptr, *narrow = nullptr;
  bool ok = true;
  try {
    CK(cudaMalloc(&tab, nent * (edw ? 3 : 2) * ew));
    CK(cudaMalloc(&tmp, chunk * (edw ? 6 : 4) * ew + 256));
    const uint32_t* tabh = c->tabQ;
    int nwin_h = c->nbytes;
    if (hb != 8) {
      // temporary narrow table: nwin * nsub windows of hb bits
      nwin_h = nwin * nsub;
      const size_t nent_h = (size_t)nwin_h * (((size_t)1 << hb) - 1);
      const size_t words = (6 * nent_h + 5 * (size_t)nwin_h) * c->L;
      CK(cudaMalloc(&narrow, words * 4 + pad256(nwin_h) + 256));
      uint32_t* w = narrow + 2 * nent_h * c->L;
      JacArr je{w, w + nent_h * c->L, w + 2 * nent_h * c->L, nent_h};
      w += 3 * nent_h * c->L;
      uint32_t* scratch = w;
      w += nent_h * c->L;
      JacArr jb{w, w + (size_t)nwin_h * c->L, w + 2 * (size_t)nwin_h * c->L, (size_t)nwin_h};
      w += 3 * (size_t)nwin_h * c->L;
      G1Arr ab{w, w + (size_t)nwin_h * c->L, reinterpret_cast<uint8_t*>(w + 2 * (size_t)nwin_h * c->L), (size_t)nwin_h};
      build_table_with(c, c->dQx, c->dQy, nwin_h, hb, narrow, jb, ab, je, scratch);
      tabh = narrow;
    }
    JacArr j;
    j.X = tmp;
    j.Y = tmp + chunk * c->L;
    j.Z = tmp + 2 * chunk * c->L;
    j.N = chunk;
    uint32_t* scratch = tmp + 3 * chunk * c->L;
    uint32_t* aff = tmp + 4 * chunk * c->L;                       // Edwards build: the chunk's affine points
    int* dbad = reinterpret_cast<int*>(tmp + (edw ? 6 : 4) * chunk * c->L);
    for (size_t first = 0; first < nent && ok; first += chunk) {
      size_t cnt = std::min(chunk, nent - first);
      {
        Timer t(c, "k_tabw_fill");
        c->Bo->tabw_fill(cfg(c, nblk(cnt, 128), 128, 0), tabh, nwin_h, nsub, hb, j.X, j.Y, j.Z, first, cnt);
        t.done();
      }
      if (edw) {
        normalize(c, j, cnt, scratch, aff, aff + c->L, 2 * (size_t)c->L, 1, nullptr);
        ok = table_to_edwards(c, aff, tab + first * 3 * c->L, scratch, dbad, cnt);
      } else {
        uint32_t* dst = tab + first * 2 * c->L;
        normalize(c, j, cnt, scratch, dst, dst + c->L, 2 * (size_t)c->L, 1, nullptr);
      }
    }
    finish(c);
  } catch (...) {
    cudaFree(tab);
    cudaFree(tmp);
    cudaFree(narrow);
    throw;
  }
  CK(cudaFree(tmp));
  if (narrow) CK(cudaFree(narrow));
  if (!ok) {  // Q is not of odd order: stay with Weierstrass tables
    cudaFree(tab);
    c->enc_edwards = false;
    ensure_tabQw(c);
    return;
  }
  c->tabQw = tab;
  c->tabQw_bits = wbits;
  c->tabQw_edw = edw;
}
// tables, window width and form for one Encrypt / re-randomisation launch
void enc_tables(bgn_ctx* c, EncArgs& ea, bool randomised) {
  if (randomised) ensure_tabQw(c);
  const bool wide = randomised && c->tabQw;
  const bool edw = wide ? c->tabQw_edw : (!randomised && c->enc_edwards && c->tabPe);
  ea.tabP = edw ? c->tabPe : c->tabP;
  ea.tabQ = wide ? c->tabQw : c->tabQ;
  ea.wbitsQ = wide ? c->tabQw_bits : 8;
  ea.edw = edw ? 1 : 0;
}

// Fixed-base table of E = e(Q,Q) for the level-2 re-randomisation `* e(Q,Q)^r` (bgn.go:283-287,
// 306-310, 469-474): nbytes windows of 8 bits, 255 canonical GT elements each (2.2 MB at 512-bit
// keys: L2-resident).  The reference computes the pairing e(Q,Q) anew in every such call.  Uses and
// releases the arena: call before carving a call's buffers.
void ensure_tabE(bgn_ctx* c) {
  if (c->tabE) return;
  int nwin = c->nbytes;
  size_t ew = (size_t)c->L * 4;
  arena_reset(c);
  arena_reserve(c, gt_bytes(c, 1) + miller_scratch(c, 1, 1) + pad256(2 * ew) + pad256((size_t)nwin * 2 * ew) + 8192);
  G1Arr Qv{c->dQx, c->dQy, c->dQinf, 1};
  GtArr e = gt_alloc(c, 1);
  run_miller(c, Qv, 1, Qv, 1, 0, 1, 1, e);
  uint32_t* gen = arena_get<uint32_t>(c, 2 * c->L);
  uint32_t* bases = arena_get<uint32_t>(c, (size_t)nwin * 2 * c->L);
  CK(cudaMemcpyAsync(gen, e.re, ew, cudaMemcpyDeviceToDevice, c->stream));
  CK(cudaMemcpyAsync(gen + c->L, e.im, ew, cudaMemcpyDeviceToDevice, c->stream));
  uint32_t* tab = nullptr;
  CK(cudaMalloc(&tab, (size_t)nwin * 255 * 2 * ew));
  try {
    {
      Timer t(c, "k_gt_tab_bases");
      c->Co->gt_tab_bases(cfg(c, 1, 32, 0), gen, nwin, bases);
      t.done();
    }
    {
      Timer t(c, "k_gt_tab_fill");
      c->Co->gt_tab_fill(cfg(c, nblk(nwin, 32), 32, 0), bases, nwin, tab);
      t.done();
    }
    finish(c);
  } catch (...) {
    cudaFree(tab);
    throw;
  }
  c->tabE = tab;
  arena_reset(c);
}

// failures with no context to hold their message (bgn_ctx_create, NULL contexts): per calling thread
thread_local std::string g_last_error;

// releases everything a context owns (bgn_ctx_destroy and the failure paths of bgn_ctx_create); the
// device lock is held by the caller
void ctx_free(bgn_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  const bool resident = g_active[c->device] == c;
  if (resident) g_active[c->device] = nullptr;
  // the secret exponent must not outlive its context: scrub the host copy and, if this context's
  // constants are the resident ones, the device copy
  memset(c->pc.exp, 0, sizeof(c->pc.exp));
  memset(c->pc.exp_naf, 0, sizeof(c->pc.exp_naf));
  c->pc.exp_bits = c->pc.exp_naf_len = 0;
  if (resident && c->has_secret && c->stream && c->A && c->Bo && c->Co) {
    c->A->upload(&c->fc, &c->pc, c->stream);
    c->Bo->upload(&c->fc, &c->pc, c->stream);
    c->Co->upload(&c->fc, &c->pc, c->stream);
    if (c->Do) c->Do->upload(&c->fc, &c->pc, c->stream);
    if (c->Eo) c->Eo->upload(&c->fc, &c->pc, c->stream);
    cudaStreamSynchronize(c->stream);
  }
  cudaFree(c->dPx);
  cudaFree(c->dPinf);
  cudaFree(c->tabP);
  cudaFree(c->tabPe);
  cudaFree(c->tabQ);
  cudaFree(c->tabQw);
  cudaFree(c->tabE);
  cudaFree(c->linesP);
  cudaFree(c->linesPq);
  cudaFree(c->bs_elems);
  cudaFree(c->bs_slots);
  cudaFree(c->bs_ginv);
  cudaFree(c->arena);
  for (auto& pnd : c->pending) {
    cudaEventDestroy(pnd.a);
    cudaEventDestroy(pnd.b);
  }
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->call_a) cudaEventDestroy(c->call_a);
  if (c->call_b) cudaEventDestroy(c->call_b);
  if (c->stream) cudaStreamDestroy(c->stream);
  cudaGetLastError();
  delete c;
}

template <typename Fn>
int guarded(bgn_ctx* c, Fn fn) {
  if (!c) {
    g_last_error = "null context";
    return BGN_E_BADARG;
  }
  std::lock_guard<std::mutex> lk(dev_mu(c->device));
  try {
    activate(c);
    arena_reset(c);
    if (c->timing) {
      if (!c->call_a) {
        CK(cudaEventCreate(&c->call_a));
        CK(cudaEventCreate(&c->call_b));
      }
      CK(cudaEventRecord(c->call_a, c->stream));
    }
    fn();
    if (c->timing) CK(cudaEventRecord(c->call_b, c->stream));
    finish(c);
    if (c->timing) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, c->call_a, c->call_b));
      c->last_call_ms = ms;
    }
    return BGN_OK;
  } catch (const CudaErr& e) {
    c->err = e.msg;
    cudaGetLastError();
    c->pending.clear();
    return BGN_E_CUDA;
  } catch (const ArgErr& e) {
    c->err = e.msg;
    return BGN_E_BADARG;
  } catch (const std::bad_alloc&) {
    c->err = "host allocation failed";
    c->pending.clear();
    return BGN_E_NOMEM;
  } catch (const std::exception& e) {  // nothing may cross extern "C" as an exception
    c->err = std::string("internal error: ") + e.what();
    c->pending.clear();
    return BGN_E_BADARG;
  } catch (...) {
    c->err = "internal error: unknown exception";
    c->pending.clear();
    return BGN_E_BADARG;
  }
}

}  // namespace

// Issue-mix microbenchmark ("is the IMAD.WIDE pipe the whole roofline?"): per inner
// iteration NW IMAD.WIDE.U32, NLH (IMAD.LO, IMAD.HI) pairs on different operands (so ptxas cannot fuse
// them into one IMAD.WIDE), NF FFMA and ND DFMA, each class on its own independent accumulators,
// interleaved in program order.  Comparing a mix's time with the times of its parts shows which
// classes share an issue pipe: pipes that co-issue give max(parts), a shared pipe gives sum(parts).
template <int NW, int NLO, int NHI, int NF, int ND>
__global__ void __launch_bounds__(256) k_issue_mix(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x, b2 = b + 77u;
  uint32_t w[16], lo[8], hi[8];
  float f[8], fx = (float)(threadIdx.x & 7) * 1.0e-3f + 1.0f, fy = 0.999f;
  double d[8], dx = (double)(threadIdx.x & 7) * 1.0e-3 + 1.0, dy = 0.999;
#pragma unroll
  for (int i = 0; i < 16; i++) w[i] = a + i;
#pragma unroll
  for (int i = 0; i < 8; i++) lo[i] = b + i, hi[i] = a - i, f[i] = (float)i, d[i] = (double)i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 8; rep++) {
      constexpr int M0 = NW > NLO ? NW : NLO, M1 = NHI > NF ? NHI : NF, M2 = M0 > M1 ? M0 : M1, M = M2 > ND ? M2 : ND;
#pragma unroll
      for (int k = 0; k < M; k++) {
        if (k < NW)
          asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;"
                       : "+r"(w[2 * (k & 7)]), "+r"(w[2 * (k & 7) + 1])
                       : "r"(a), "r"(b));
        // the accumulator is also the multiplicand, so neither product is loop-invariant (ptxas hoists
        // an invariant mad.hi out of the loop and leaves only its additions)
        if (k < NLO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(lo[k & 7]) : "r"(b), "r"(a));
        if (k < NHI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(hi[k & 7]) : "r"(b2), "r"(a));
        if (k < NF) asm volatile("fma.rn.f32 %0, %1, %2, %0;" : "+f"(f[k & 7]) : "f"(fx), "f"(fy));
        if (k < ND) asm volatile("fma.rn.f64 %0, %1, %2, %0;" : "+d"(d[k & 7]) : "d"(dx), "d"(dy));
      }
    }
  }
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) o ^= w[i];
#pragma unroll
  for (int i = 0; i < 8; i++) o ^= lo[i] ^ hi[i] ^ __float_as_uint(f[i]) ^ (uint32_t)__double_as_longlong(d[i]);
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = o;
}

// ================================================================== C-ABI
extern "C" {

int bgn_ctx_create(const bgn_params* prm, int device, bgn_ctx** out) {
  g_last_error.clear();
  if (!prm || !out || !prm->p_be || !prm->n_be || !prm->P_bytes || !prm->Q_bytes || prm->l == 0) {
    g_last_error = "bgn_ctx_create: null argument or l == 0";
    return BGN_E_BADARG;
  }
  *out = nullptr;
  if (device < 0 || device >= kMaxDevices) {
    g_last_error = "bgn_ctx_create: bad device ordinal";
    return BGN_E_BADARG;
  }
  std::lock_guard<std::mutex> lk(dev_mu(device));
  bgn_ctx* c = nullptr;
  try {
    c = new bgn_ctx();
    c->device = device;
    if (const char* sk = getenv("BGN_MILLER_SKEW")) c->miller_skew = atoi(sk);  // tuning knob (cycles)
    if (const char* gr = getenv("BGN_MILLER_GROUPS")) c->miller_groups = atoi(gr);
    if (const char* tl = getenv("BGN_MILLER_TAIL")) c->miller_tail = atoi(tl);
    if (const char* ew = getenv("BGN_ENC_WINDOW")) {
      int w = atoi(ew);
      c->enc_window = (w == 8 || (w >= 16 && w <= 24 && w % 2 == 0)) ? w : 0;
    }
    if (const char* ee = getenv("BGN_ENC_EDWARDS")) c->enc_edwards = atoi(ee) != 0;
    if (const char* dl = getenv("BGN_DEC_LUCAS")) c->dec_lucas = atoi(dl) != 0;
    if (const char* af = getenv("BGN_AFFINE_ADD")) c->affine_add = atoi(af) != 0;
    if (const char* np = getenv("BGN_NORM_PER_THREAD")) c->norm_per_thread = std::max(1, atoi(np));
    if (const char* nt = getenv("BGN_NORM_THREADS")) c->norm_threads = std::max(128, atoi(nt));
    if (const char* fl = getenv("BGN_FIXED_LINES")) c->fixed_lines = atoi(fl) != 0;
    if (const char* fp = getenv("BGN_FIXED_PAIR")) c->fixed_pair = atoi(fp);
    if (const char* pd = getenv("BGN_PAIR_DUO")) c->pair_duo = atoi(pd);
    if (const char* ms = getenv("BGN_MILLER_SPLIT")) c->miller_split = atoi(ms);
    if (const char* mw = getenv("BGN_MILLER_WIDE")) c->miller_wide = atoi(mw);
    Big p0 = big_from_be(prm->p_be, prm->p_len, BGN_MAXL);
    int pbits = big_bits(p0);
    if (pbits < 40 || (p0[0] & 3) != 3) throw ArgErr{"p must be a prime = 3 (mod 4) of at least 40 bits"};
    int L = 0;
    for (int cand : kSupportedL)
      if (32 * cand >= pbits + 8) {
        L = cand;
        break;
      }
    if (!L) throw ArgErr{"field too large (max 1049-bit p)"};
    c->L = L;
    switch (L) {
#define BGN_PICK(n)          \
  case n:                    \
    c->A = bgn_opsA_##n();   \
    c->Bo = bgn_opsB_##n();  \
    c->Co = bgn_opsC_##n();  \
    c->Do = bgn_opsD_##n();  \
    c->Eo = bgn_opsE_##n();  \
    break;
#ifdef BGN_HAVE_L3
      BGN_PICK(3)
#endif
#ifdef BGN_HAVE_L5
      BGN_PICK(5)
#endif
#ifdef BGN_HAVE_L9
      BGN_PICK(9)
#endif
#ifdef BGN_HAVE_L17
      BGN_PICK(17)
#endif
#ifdef BGN_HAVE_L33
      BGN_PICK(33)
#endif
#undef BGN_PICK
      default:
        break;
    }
    if (!c->A || !c->Bo || !c->Co || !c->Do || !c->Eo) throw ArgErr{"this build of libbgn_b200 does not instantiate the limb count this key needs"};
    c->B = (pbits + 7) / 8;
    Big p(p0.begin(), p0.begin() + L);
    Big n = big_from_be(prm->n_be, prm->n_len, BGN_MAXL);
    int nbits = big_bits(n);
    if (nbits < 16 || !(n[0] & 1) || nbits > pbits) throw ArgErr{"bad group order n"};
    c->nbytes = (nbits + 7) / 8;
    c->n_limbs = n;
    // p + 1 == l * n  (type a1)
    {
      Big acc(BGN_MAXL + 2, 0);
      uint64_t carry = 0;
      uint64_t l_lo = prm->l & 0xffffffffu, l_hi = prm->l >> 32;
      // acc = n * l (schoolbook with two 32-bit digits)
      for (int i = 0; i < BGN_MAXL; i++) {
        unsigned __int128 t = (unsigned __int128)n[i] * l_lo + acc[i] + carry;
        acc[i] = (uint32_t)t;
        carry = (uint64_t)(t >> 32);
      }
      acc[BGN_MAXL] = (uint32_t)carry;
      carry = 0;
      for (int i = 0; i < BGN_MAXL; i++) {
        unsigned __int128 t = (unsigned __int128)n[i] * l_hi + acc[i + 1] + carry;
        acc[i + 1] = (uint32_t)t;
        carry = (uint64_t)(t >> 32);
      }
      Big pp(BGN_MAXL + 2, 0);
      for (int i = 0; i < BGN_MAXL; i++) pp[i] = p0[i];
      size_t i = 0;
      while (++pp[i] == 0) i++;  // p + 1
      if (pp != acc) throw ArgErr{"parameters are not type a1: p + 1 != l * n"};
    }
    memset(&c->fc, 0, sizeof(c->fc));
    memset(&c->pc, 0, sizeof(c->pc));
    for (int i = 0; i < L; i++) c->fc.p[i] = p[i];
    Big p2(L);
    big_add(p2, p, p);
    for (int i = 0; i < L; i++) c->fc.p2[i] = p2[i];
    {
      Big p4(L), p8(L), p16(L);
      big_add(p4, p2, p2);
      big_add(p8, p4, p4);
      big_add(p16, p8, p8);
      for (int i = 0; i < L; i++) c->fc.p4[i] = p4[i], c->fc.p8[i] = p8[i], c->fc.p16[i] = p16[i];
    }
    Big x(L, 0);
    x[0] = 1;
    for (int i = 0; i < 32 * L; i++) big_dbl_mod(x, p);
    for (int i = 0; i < L; i++) c->fc.one[i] = x[i];
    for (int i = 0; i < 32 * L; i++) big_dbl_mod(x, p);
    for (int i = 0; i < L; i++) c->fc.r2[i] = x[i];
    uint32_t inv = p[0];
    for (int i = 0; i < 5; i++) inv *= 2 - p[0] * inv;
    c->fc.np0 = (uint32_t)(0u - inv);
    c->pc.l = prm->l;
    std::vector<int8_t> naf = big_naf(Big(n.begin(), n.begin() + (nbits + 31) / 32));
    if (naf.size() > BGN_MAX_NAF) throw ArgErr{"group order too large"};
    c->pc.naf_len = (int)naf.size();
    for (size_t i = 0; i < naf.size(); i++) c->pc.naf[i] = naf[i];

    CK(cudaSetDevice(device));
    // A BLOCKING stream: it orders itself after work already queued on the legacy default stream
    // (where a caller that hands over device pointers -- e.g. PyTorch -- normally produced them), and
    // every call ends with a stream synchronise, so results are complete when the call returns.
    // Producers on other streams must be synchronised by the caller (include/bgn_b200.h).
    CK(cudaStreamCreate(&c->stream));
    g_active[device] = nullptr;
    activate(c);
    // generators
    size_t lw = (size_t)L * 4;
    CK(cudaMalloc(&c->dPx, 4 * lw));
    c->dPy = c->dPx + L;
    c->dQx = c->dPx + 2 * L;
    c->dQy = c->dPx + 3 * L;
    CK(cudaMalloc(&c->dPinf, 2));
    c->dQinf = c->dPinf + 1;
    arena_reset(c);
    arena_reserve(c, 4 * io_bytes(c, 1) + 4096);
    {
      uint8_t* dp = arena_get<uint8_t>(c, 2 * c->B);
      uint8_t* dq = arena_get<uint8_t>(c, 2 * c->B);
      CK(cudaMemcpyAsync(dp, prm->P_bytes, 2 * c->B, cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemcpyAsync(dq, prm->Q_bytes, 2 * c->B, cudaMemcpyHostToDevice, c->stream));
      G1Arr aP{c->dPx, c->dPy, c->dPinf, 1}, aQ{c->dQx, c->dQy, c->dQinf, 1};
      g1_from_bytes(c, dp, 1, aP);
      g1_from_bytes(c, dq, 1, aQ);
      finish(c);
      uint8_t flags[2];
      CK(cudaMemcpy(flags, c->dPinf, 2, cudaMemcpyDeviceToHost));
      if (flags[0] || flags[1]) throw ArgErr{"generator P or Q is not a point on the curve"};
    }
    size_t tabP_words = (size_t)8 * 255 * 2 * L, tabQ_words = (size_t)c->nbytes * 255 * 2 * L;
    CK(cudaMalloc(&c->tabP, tabP_words * 4));
    CK(cudaMalloc(&c->tabQ, tabQ_words * 4));
    build_table(c, c->dPx, c->dPy, 8, c->tabP);
    build_table(c, c->dQx, c->dQy, c->nbytes, c->tabQ);
    if (c->enc_edwards) {
      // The Edwards form's unified addition has no exceptional cases on points of ODD order only: check
      // n P = n Q = O (n is odd).  Every key NewKeyGen makes passes (bgn.go:112-119); other generators keep
      // the Weierstrass tables and their complete addition.
      uint8_t* dn = nullptr;
      CK(cudaMalloc(&dn, pad256(prm->n_len) + 6 * (size_t)L * 4));
      bool odd = true;
      try {
        uint32_t* jz = reinterpret_cast<uint32_t*>(dn + pad256(prm->n_len));
        CK(cudaMemcpyAsync(dn, prm->n_be, prm->n_len, cudaMemcpyHostToDevice, c->stream));
        for (int w = 0; w < 2; w++) {
          G1MulArgs ma;
          ma.x = w ? c->dQx : c->dPx;
          ma.y = w ? c->dQy : c->dPy;
          ma.inf = w ? c->dQinf : c->dPinf;
          ma.Nin = 1;
          ma.k_be = dn;
          ma.kbytes = (int)prm->n_len;
          ma.X = jz;
          ma.Y = jz + L;
          ma.Z = jz + 2 * L + w * L;  // Z of P's multiple, then Z of Q's
          ma.count = 1;
          ma.N = 1;
          c->Bo->g1_mulvar(cfg(c, 1, 32, 0), ma);
        }
        std::vector<uint32_t> z(2 * (size_t)L);
        CK(cudaMemcpyAsync(z.data(), jz + 2 * L, 2 * (size_t)L * 4, cudaMemcpyDeviceToHost, c->stream));
        finish(c);
        for (int w = 0; w < 2; w++) {
          bool zero = true, is_p = true;
          for (int i = 0; i < L; i++) {
            zero = zero && z[(size_t)w * L + i] == 0;
            is_p = is_p && z[(size_t)w * L + i] == p[i];
          }
          odd = odd && (zero || is_p);
        }
      } catch (...) {
        cudaFree(dn);
        throw;
      }
      cudaFree(dn);
      if (!odd) c->enc_edwards = false;
    }
    if (c->enc_edwards) {  // P's table in Edwards form (Q's wide table is built on the first randomised encryption)
      const size_t nP = (size_t)8 * 255;
      uint32_t* scr = nullptr;
      CK(cudaMalloc(&c->tabPe, nP * 3 * L * 4));
      CK(cudaMalloc(&scr, nP * L * 4 + 256));
      bool ok = false;
      try {
        ok = table_to_edwards(c, c->tabP, c->tabPe, scr, reinterpret_cast<int*>(scr + nP * L), nP);
      } catch (...) {
        cudaFree(scr);
        throw;
      }
      cudaFree(scr);
      if (!ok) {
        cudaFree(c->tabPe);
        c->tabPe = nullptr;
        c->enc_edwards = false;
      }
    }
    *out = c;
    return BGN_OK;
  } catch (const CudaErr& e) {
    g_last_error = "bgn_ctx_create: " + e.msg;
    cudaGetLastError();
    ctx_free(c);
    return BGN_E_CUDA;
  } catch (const ArgErr& e) {
    g_last_error = "bgn_ctx_create: " + e.msg;
    ctx_free(c);
    return BGN_E_BADARG;
  } catch (const std::bad_alloc&) {
    g_last_error = "bgn_ctx_create: host allocation failed";
    ctx_free(c);
    return BGN_E_NOMEM;
  } catch (const std::exception& e) {
    g_last_error = std::string("bgn_ctx_create: internal error: ") + e.what();
    ctx_free(c);
    return BGN_E_BADARG;
  } catch (...) {
    g_last_error = "bgn_ctx_create: internal error: unknown exception";
    ctx_free(c);
    return BGN_E_BADARG;
  }
}

const char* bgn_global_last_error(void) { return g_last_error.c_str(); }

void bgn_ctx_destroy(bgn_ctx* c) {
  if (!c) return;
  std::lock_guard<std::mutex> lk(dev_mu(c->device));
  ctx_free(c);
}

const char* bgn_last_error(const bgn_ctx* c) { return c ? c->err.c_str() : g_last_error.c_str(); }

int bgn_ctx_set_option(bgn_ctx* c, const char* name, long value) {
  if (!c || !name) return BGN_E_BADARG;
  std::lock_guard<std::mutex> lk(dev_mu(c->device));
  std::string k(name);
  if (k == "enc_window") {
    if (value != 0 && value != 8 && !(value >= 16 && value <= 24 && value % 2 == 0)) {
      c->err = "enc_window must be 0 (automatic), 8, 16, 18, 20, 22 or 24";
      return BGN_E_BADARG;
    }
    if (c->enc_window != (int)value) {
      cudaSetDevice(c->device);
      if (c->stream) cudaStreamSynchronize(c->stream);
      cudaFree(c->tabQw);
      c->tabQw = nullptr;
      c->tabQw_bits = 0;
    }
    c->enc_window = (int)value;
  } else if (k == "enc_edwards") {
    const bool on = value != 0 && c->tabPe != nullptr;
    if (on != c->enc_edwards) {
      cudaSetDevice(c->device);
      if (c->stream) cudaStreamSynchronize(c->stream);
      cudaFree(c->tabQw);
      c->tabQw = nullptr;
      c->tabQw_bits = 0;
    }
    c->enc_edwards = on;
  } else if (k == "enc_table_max_mb") {
    c->enc_table_max = (size_t)std::max<long>(0, value) << 20;
    if (c->enc_window == 0 && c->tabQw) {
      cudaSetDevice(c->device);
      if (c->stream) cudaStreamSynchronize(c->stream);
      cudaFree(c->tabQw);
      c->tabQw = nullptr;
      c->tabQw_bits = 0;
    }
  } else if (k == "dec_lucas") {
    c->dec_lucas = value != 0;
  } else if (k == "fixed_lines") {
    c->fixed_lines = value != 0;
  } else if (k == "split_para") {
    c->split_para = value < 0 ? -1 : (value != 0);
  } else if (k == "dec_pair_q1") {
    c->dec_pair_q1 = value != 0;  // takes effect for tables built by the next bgn_ctx_set_secret; 0 also stops using one
  } else if (k == "fixed_pair") {
    c->fixed_pair = value < 0 ? -1 : (value != 0);
  } else if (k == "pair_duo") {
    c->pair_duo = value < 0 ? -1 : (value != 0);
  } else if (k == "miller_wide") {
    c->miller_wide = (int)std::max<long>(0, value);
  } else if (k == "miller_wide_min") {
    c->miller_wide_min = (int)std::max<long>(1, value);
  } else if (k == "miller_split") {
    c->miller_split = value < 0 ? -1 : (value != 0);
  } else if (k == "pair_duo_loop") {
    c->pair_duo_loop = (int)value;
    c->pair_duo_cap = (size_t)-1;
  } else if (k == "pair_duo_pairs") {
    c->pair_duo_pairs = (int)std::max<long>(1, std::min<long>(4, value));
    c->pair_duo_cap = (size_t)-1;
  } else if (k == "pair_duo_blockbar") {
    c->pair_duo_blockbar = value != 0;
  } else {
    c->err = "unknown option " + k;
    return BGN_E_BADARG;
  }
  return BGN_OK;
}

int bgn_ctx_info(const bgn_ctx* c, int* limbs, int* coord_bytes, int* scalar_bytes) {
  if (!c) return BGN_E_BADARG;
  if (limbs) *limbs = c->L;
  if (coord_bytes) *coord_bytes = c->B;
  if (scalar_bytes) *scalar_bytes = c->nbytes;
  return BGN_OK;
}

int bgn_encrypt_batch(bgn_ctx* c, const int64_t* x, const uint8_t* r_be, size_t count, uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!x || !out) throw ArgErr{"null buffer"};
    check_count(count);
    arena_reserve(c, pad256(count * 8) + pad256(count * c->nbytes) + jac_bytes(c, count) + g1_bytes(c, count) +
                         pad256(count * c->L * 4) + io_bytes(c, count) + 4096);
    const int64_t* dx = reinterpret_cast<const int64_t*>(stage_in(c, x, count * 8));
    const uint8_t* dr = r_be ? stage_in(c, r_be, count * c->nbytes) : nullptr;
    OutBuf ob = stage_out(c, out, count * 2 * c->B);
    JacArr j = jac_alloc(c, count);
    G1Arr a = g1_alloc(c, count);
    uint32_t* scratch = arena_get<uint32_t>(c, count * c->L);
    EncArgs ea;
    ea.x = dx;
    ea.r_be = dr;
    ea.rbytes = c->nbytes;
    enc_tables(c, ea, dr != nullptr);
    ea.X = j.X;
    ea.Y = j.Y;
    ea.Z = j.Z;
    ea.count = count;
    ea.N = count;
    ea.bx = ea.by = nullptr;
    ea.binf = nullptr;
    {
      Timer t(c, "k_encrypt");
      c->Bo->encrypt(cfg(c, nblk(count, 128), 128, 0), ea);
      t.done();
    }
    normalize_soa(c, j, count, scratch, a);
    g1_to_bytes(c, a, count, ob.dev);
    commit_out(c, ob);
  });
}

// R = A + B / A - B / O - B (A a single O element when bcast1) on device arrays; arena: scratch only
static void g1_add_core(bgn_ctx* c, const G1Arr& A, const G1Arr& Bv, size_t count, int subtract, int bcast1, const G1Arr& R) {
  if (c->affine_add) {
    // affine + affine -> affine with one binary-GCD inversion per thread (types.h: G1AffAddArgs)
    uint32_t* scratch = arena_get<uint32_t>(c, count * c->L);
    G1AffAddArgs aa;
    aa.x1 = A.x;
    aa.y1 = A.y;
    aa.inf1 = A.inf;
    aa.x2 = Bv.x;
    aa.y2 = Bv.y;
    aa.inf2 = Bv.inf;
    aa.bcast1 = bcast1;
    aa.subtract = subtract;
    aa.ox = R.x;
    aa.oy = R.y;
    aa.oinf = R.inf;
    aa.scratch = scratch;
    aa.count = count;
    aa.G = (int)shared_inversion_threads(c, count);
    Timer t(c, "k_g1_affadd");
    c->Bo->g1_affadd(cfg(c, nblk(aa.G, 128), 128, 0), aa);
    t.done();
    return;
  }
  JacArr j = jac_alloc(c, count);
  uint32_t* scratch = arena_get<uint32_t>(c, count * c->L);
  G1AddArgs ga;
  ga.x1 = A.x;
  ga.y1 = A.y;
  ga.inf1 = A.inf;
  ga.N1 = A.N;
  ga.x2 = Bv.x;
  ga.y2 = Bv.y;
  ga.inf2 = Bv.inf;
  ga.N2 = Bv.N;
  ga.bcast1 = bcast1;
  ga.subtract = subtract;
  ga.X = j.X;
  ga.Y = j.Y;
  ga.Z = j.Z;
  ga.count = count;
  ga.N = j.N;
  {
    Timer t(c, "k_g1_add");
    c->Bo->g1_add(cfg(c, nblk(count, 128), 128, 0), ga);
    t.done();
  }
  normalize_soa(c, j, count, scratch, R);
}

static int g1_binop(bgn_ctx* c, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out, int subtract,
                    int neg_only) {
  return guarded(c, [&] {
    if (!count) return;
    if ((!neg_only && !a) || !b || !out) throw ArgErr{"null buffer"};
    check_count(count);
    arena_reserve(c, 3 * io_bytes(c, count) + 3 * g1_bytes(c, count) + jac_bytes(c, count) + pad256(count * c->L * 4) + 8192);
    OutBuf ob = stage_out(c, out, count * 2 * c->B);
    G1Arr A, Bv = g1_alloc(c, count), R = g1_alloc(c, count);
    if (!neg_only) {
      const uint8_t* da = stage_in(c, a, count * 2 * c->B);
      A = g1_alloc(c, count);
      g1_from_bytes(c, da, count, A);
    } else {
      A = g1_alloc(c, 1);
      CK(cudaMemsetAsync(A.inf, 1, 1, c->stream));  // O
    }
    const uint8_t* db = stage_in(c, b, count * 2 * c->B);
    g1_from_bytes(c, db, count, Bv);
    g1_add_core(c, A, Bv, count, subtract, neg_only, R);
    g1_to_bytes(c, R, count, ob.dev);
    commit_out(c, ob);
  });
}
int bgn_g1_add_batch(bgn_ctx* c, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out) {
  return g1_binop(c, a, b, count, out, 0, 0);
}
int bgn_g1_sub_batch(bgn_ctx* c, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out) {
  return g1_binop(c, a, b, count, out, 1, 0);
}
int bgn_g1_neg_batch(bgn_ctx* c, const uint8_t* a, size_t count, uint8_t* out) {
  return g1_binop(c, nullptr, a, count, out, 1, 1);
}

int bgn_g1_mulconst_batch(bgn_ctx* c, const uint8_t* a, const uint8_t* k_be, size_t kbytes, size_t count,
                          uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!a || !k_be || !out || !kbytes || kbytes > 4096) throw ArgErr{"bad argument"};
    check_count(count);
    arena_reserve(c, 2 * io_bytes(c, count) + pad256(count * kbytes) + 2 * g1_bytes(c, count) + jac_bytes(c, count) +
                         pad256(count * c->L * 4) + 8192);
    OutBuf ob = stage_out(c, out, count * 2 * c->B);
    const uint8_t* da = stage_in(c, a, count * 2 * c->B);
    const uint8_t* dk = stage_in(c, k_be, count * kbytes);
    G1Arr A = g1_alloc(c, count), R = g1_alloc(c, count);
    g1_from_bytes(c, da, count, A);
    JacArr j = jac_alloc(c, count);
    uint32_t* scratch = arena_get<uint32_t>(c, count * c->L);
    G1MulArgs ma;
    ma.x = A.x;
    ma.y = A.y;
    ma.inf = A.inf;
    ma.Nin = A.N;
    ma.k_be = dk;
    ma.kbytes = (int)kbytes;
    ma.X = j.X;
    ma.Y = j.Y;
    ma.Z = j.Z;
    ma.count = count;
    ma.N = j.N;
    {
      Timer t(c, "k_g1_mulvar");
      c->Bo->g1_mulvar(cfg(c, nblk(count, 128), 128, 0), ma);
      t.done();
    }
    normalize_soa(c, j, count, scratch, R);
    g1_to_bytes(c, R, count, ob.dev);
    commit_out(c, ob);
  });
}

static int gt_binop(bgn_ctx* c, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out, int conj_b) {
  return guarded(c, [&] {
    if (!count) return;
    if (!a || !b || !out) throw ArgErr{"null buffer"};
    check_count(count);
    arena_reserve(c, 3 * io_bytes(c, count) + 3 * gt_bytes(c, count) + 8192);
    OutBuf ob = stage_out(c, out, count * 2 * c->B);
    const uint8_t* da = stage_in(c, a, count * 2 * c->B);
    const uint8_t* db = stage_in(c, b, count * 2 * c->B);
    GtArr A = gt_alloc(c, count), Bv = gt_alloc(c, count), R = gt_alloc(c, count);
    gt_from_bytes(c, da, count, A);
    gt_from_bytes(c, db, count, Bv);
    GtBinArgs ga;
    ga.are = A.re;
    ga.aim = A.im;
    ga.Na = A.N;
    ga.bre = Bv.re;
    ga.bim = Bv.im;
    ga.Nb = Bv.N;
    ga.conj_b = conj_b;
    ga.ore = R.re;
    ga.oim = R.im;
    ga.count = count;
    ga.N = R.N;
    {
      Timer t(c, "k_gt_mul");
      c->A->gt_mul(cfg(c, nblk(count, 128), 128, 0), ga);
      t.done();
    }
    gt_to_bytes(c, R, count, ob.dev);
    commit_out(c, ob);
  });
}
int bgn_gt_mul_batch(bgn_ctx* c, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out) {
  return gt_binop(c, a, b, count, out, 0);
}
int bgn_gt_div_batch(bgn_ctx* c, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out) {
  return gt_binop(c, a, b, count, out, 1);
}

static int gt_pow_impl(bgn_ctx* c, const uint8_t* a, const uint8_t* k_be, size_t kbytes, size_t count, uint8_t* out,
                       int mode /*0 var, 1 secret, 2 inverse*/) {
  return guarded(c, [&] {
    if (!count) return;
    if (!a || !out) throw ArgErr{"null buffer"};
    if (mode == 0 && (!k_be || !kbytes || kbytes > 4096)) throw ArgErr{"bad exponent buffer"};
    if (mode == 1 && !c->has_secret) throw ArgErr{"secret key not set"};
    check_count(count);
    arena_reserve(c, 2 * io_bytes(c, count) + pad256(count * (kbytes + 1)) + 2 * gt_bytes(c, count) + 8192);
    OutBuf ob = stage_out(c, out, count * 2 * c->B);
    const uint8_t* da = stage_in(c, a, count * 2 * c->B);
    GtArr A = gt_alloc(c, count), R = gt_alloc(c, count);
    gt_from_bytes(c, da, count, A);
    GtPowArgs pa;
    pa.re = A.re;
    pa.im = A.im;
    pa.Nin = A.N;
    pa.e_be = nullptr;
    pa.ebytes = 0;
    pa.mode = mode;
    if (mode == 0) {
      pa.e_be = stage_in(c, k_be, count * kbytes);
      pa.ebytes = (int)kbytes;
    }
    pa.ore = R.re;
    pa.oim = R.im;
    pa.count = count;
    pa.N = R.N;
    if (mode == 1 && c->dec_lucas) {  // fixed exponent q1: a lane pair per element (lucas.cuh: GtPowPair)
      Timer t(c, "k_gt_pow_pair");
      c->Co->gt_pow_pair(cfg(c, nblk(2 * count, 64), 64, 0), pa);
      t.done();
    } else {
      Timer t(c, "k_gt_pow");
      c->A->gt_pow(cfg(c, nblk(count, 128), 128, 0), pa);
      t.done();
    }
    gt_to_bytes(c, R, count, ob.dev);
    commit_out(c, ob);
  });
}
int bgn_gt_pow_batch(bgn_ctx* c, const uint8_t* a, const uint8_t* k_be, size_t kbytes, size_t count, uint8_t* out) {
  return gt_pow_impl(c, a, k_be, kbytes, count, out, 0);
}
int bgn_gt_pow_secret_batch(bgn_ctx* c, const uint8_t* in, size_t count, uint8_t* out) {
  return gt_pow_impl(c, in, nullptr, 0, count, out, 1);
}
int bgn_gt_inv_batch(bgn_ctx* c, const uint8_t* a, size_t count, uint8_t* out) {
  return gt_pow_impl(c, a, nullptr, 0, count, out, 2);
}

static void pair_common(bgn_ctx* c, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out) {
  if (!b) ensure_linesP(c);
  arena_reserve(c, 3 * io_bytes(c, count) + 2 * g1_bytes(c, count) + gt_bytes(c, count) + miller_scratch(c, count, 1) + 8192);
  OutBuf ob = stage_out(c, out, count * 2 * c->B);
  const uint8_t* da = stage_in(c, a, count * 2 * c->B);
  G1Arr A = g1_alloc(c, count);
  g1_from_bytes(c, da, count, A);
  GtArr R = gt_alloc(c, count);
  if (b) {
    const uint8_t* db = stage_in(c, b, count * 2 * c->B);
    G1Arr Bv = g1_alloc(c, count);
    g1_from_bytes(c, db, count, Bv);
    // a batch that fits one wave of the two-warp kernel leaves the one-thread kernel short of warps
    // (measured A/B: profiles/r02_pair_duo_ab_*.json)
    const size_t cap = pair_duo_capacity(c);
    if (cap && (c->pair_duo > 0 || (c->pair_duo < 0 && count <= cap)))
      run_pair_duo(c, A, Bv, count, R);
    else
      run_miller(c, A, 1, Bv, 1, 0, count, 1, R);
  } else if (c->linesP && c->fixed_lines) {
    run_miller_fixed(c, A, count, R);
  } else {
    G1Arr Pv{c->dPx, c->dPy, c->dPinf, 1};
    run_miller(c, A, 1, Pv, 1, 1, count, 1, R);
  }
  gt_to_bytes(c, R, count, ob.dev);
  commit_out(c, ob);
}
int bgn_pair_batch(bgn_ctx* c, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!a || !b || !out) throw ArgErr{"null buffer"};
    check_count(count);
    pair_common(c, a, b, count, out);
  });
}
int bgn_make_l2_batch(bgn_ctx* c, const uint8_t* a, size_t count, uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!a || !out) throw ArgErr{"null buffer"};
    check_count(count);
    pair_common(c, a, nullptr, count, out);
  });
}

int bgn_multpoly_batch(bgn_ctx* c, const uint8_t* c1, size_t d1, const uint8_t* c2, size_t d2, size_t count,
                       uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!c1 || !c2 || !out || !d1 || !d2 || d1 > 128 || d2 > 128) throw ArgErr{"bad argument"};
    check_count(count * (d1 + d2));
    size_t n1 = count * d1, n2 = count * d2, no = count * (d1 + d2);
    arena_reserve(c, io_bytes(c, n1) + io_bytes(c, n2) + io_bytes(c, no) + g1_bytes(c, n1) + g1_bytes(c, n2) +
                         gt_bytes(c, no) + miller_scratch(c, count, (int)std::max(d1, d2)) + 8192);
    OutBuf ob = stage_out(c, out, no * 2 * c->B);
    const uint8_t* da = stage_in(c, c1, n1 * 2 * c->B);
    const uint8_t* db = stage_in(c, c2, n2 * 2 * c->B);
    G1Arr A = g1_alloc(c, n1), Bv = g1_alloc(c, n2);
    g1_from_bytes(c, da, n1, A);
    g1_from_bytes(c, db, n2, Bv);
    GtArr R = gt_alloc(c, no);
    // the pairing is symmetric: the shorter polynomial supplies the Miller points
    if (d1 <= d2)
      run_miller(c, A, (int)d1, Bv, (int)d2, 0, count, (int)(d1 + d2), R);
    else
      run_miller(c, Bv, (int)d2, A, (int)d1, 0, count, (int)(d1 + d2), R);
    gt_to_bytes(c, R, no, ob.dev);
    commit_out(c, ob);
  });
}

// product tree over terms; leaves `ncoeff` elements in the returned array
static GtArr reduce_tree(bgn_ctx* c, GtArr cur, size_t nterms, size_t ncoeff) {
  while (nterms > 1) {
    size_t G = (nterms + 15) / 16;
    if (nterms <= 64) G = 1;
    GtArr nxt = gt_alloc(c, G * ncoeff);
    Timer t(c, "k_gt_reduce");
    c->A->gt_reduce(cfg(c, nblk(G * ncoeff, 128), 128, 0), cur.re, cur.im, cur.N, nterms, (int)ncoeff, (int)G, nxt.re, nxt.im, nxt.N);
    t.done();
    cur = nxt;
    nterms = G;
  }
  return cur;
}

int bgn_l2_sum_reduce(bgn_ctx* c, const uint8_t* in, size_t nterms, size_t ncoeff, uint8_t* out) {
  return guarded(c, [&] {
    if (!ncoeff) return;
    if (!out || (nterms && !in) || ncoeff > (1u << 20)) throw ArgErr{"bad argument"};
    check_count(nterms * ncoeff + ncoeff);
    size_t n = nterms * ncoeff;
    arena_reserve(c, io_bytes(c, n) + io_bytes(c, ncoeff) + 3 * gt_bytes(c, n + ncoeff) + 65536);
    OutBuf ob = stage_out(c, out, ncoeff * 2 * c->B);
    GtArr R;
    if (nterms == 0) {
      R = gt_alloc(c, ncoeff);
      Timer t(c, "k_gt_reduce");
      c->A->gt_reduce(cfg(c, nblk(ncoeff, 128), 128, 0), R.re, R.im, R.N, 0, (int)ncoeff, 1, R.re, R.im, R.N);
      t.done();
    } else {
      const uint8_t* di = stage_in(c, in, n * 2 * c->B);
      GtArr A = gt_alloc(c, n);
      gt_from_bytes(c, di, n, A);
      R = reduce_tree(c, A, nterms, ncoeff);
    }
    gt_to_bytes(c, R, ncoeff, ob.dev);
    commit_out(c, ob);
  });
}

// ---- non-deterministic mode ------------------------------------------------
int bgn_g1_blind_batch(bgn_ctx* c, const uint8_t* a, const uint8_t* r_be, size_t count, uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!a || !r_be || !out) throw ArgErr{"null buffer"};
    check_count(count);
    arena_reserve(c, 2 * io_bytes(c, count) + pad256(count * c->nbytes) + jac_bytes(c, count) + 2 * g1_bytes(c, count) +
                         pad256(count * c->L * 4) + 8192);
    OutBuf ob = stage_out(c, out, count * 2 * c->B);
    const uint8_t* da = stage_in(c, a, count * 2 * c->B);
    const uint8_t* dr = stage_in(c, r_be, count * c->nbytes);
    G1Arr A = g1_alloc(c, count), R = g1_alloc(c, count);
    g1_from_bytes(c, da, count, A);
    JacArr j = jac_alloc(c, count);
    uint32_t* scratch = arena_get<uint32_t>(c, count * c->L);
    EncArgs ea;
    ea.x = nullptr;
    ea.r_be = dr;
    ea.rbytes = c->nbytes;
    enc_tables(c, ea, true);
    ea.X = j.X;
    ea.Y = j.Y;
    ea.Z = j.Z;
    ea.count = count;
    ea.N = count;
    ea.bx = A.x;
    ea.by = A.y;
    ea.binf = A.inf;
    {
      Timer t(c, "k_encrypt");
      c->Bo->encrypt(cfg(c, nblk(count, 128), 128, 0), ea);
      t.done();
    }
    normalize_soa(c, j, count, scratch, R);
    g1_to_bytes(c, R, count, ob.dev);
    commit_out(c, ob);
  });
}

int bgn_gt_blind_batch(bgn_ctx* c, const uint8_t* a, const uint8_t* r_be, size_t count, uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!a || !r_be || !out) throw ArgErr{"null buffer"};
    check_count(count);
    ensure_tabE(c);
    arena_reserve(c, 2 * io_bytes(c, count) + pad256(count * c->nbytes) + 2 * gt_bytes(c, count) + 8192);
    OutBuf ob = stage_out(c, out, count * 2 * c->B);
    const uint8_t* da = stage_in(c, a, count * 2 * c->B);
    const uint8_t* dr = stage_in(c, r_be, count * c->nbytes);
    GtArr A = gt_alloc(c, count), R = gt_alloc(c, count);
    gt_from_bytes(c, da, count, A);
    GtBlindArgs ba;
    ba.re = A.re;
    ba.im = A.im;
    ba.r_be = dr;
    ba.rbytes = c->nbytes;
    ba.tabE = c->tabE;
    ba.ore = R.re;
    ba.oim = R.im;
    ba.count = count;
    {
      Timer t(c, "k_gt_blind");
      c->Co->gt_blind(cfg(c, nblk(count, 128), 128, 0), ba);
      t.done();
    }
    gt_to_bytes(c, R, count, ob.dev);
    commit_out(c, ob);
  });
}

// ---- polynomial-ciphertext helpers ------------------------------------------
// out[u][jj] = sum_k w[k] * in[u][j_begin + jj - k]   (types.h: PolyConvArgs)
static void polyconv(bgn_ctx* c, const uint8_t* in, size_t d, int is_l2, const unsigned __int128* w, size_t nw,
                     int j_begin, int j_count, int negate, size_t count, uint8_t* out) {
  size_t nin = count * d, nout = count * (size_t)j_count;
  check_count(nin);
  check_count(nout);
  arena_reserve(c, io_bytes(c, nin) + io_bytes(c, nout) + g1_bytes(c, nin) + g1_bytes(c, nout) + jac_bytes(c, nout) +
                       pad256(nout * c->L * 4) + 8192);
  OutBuf ob = stage_out(c, out, nout * 2 * c->B);
  const uint8_t* di = stage_in(c, in, nin * 2 * c->B);
  PolyConvArgs pa;
  memset(&pa, 0, sizeof(pa));
  pa.d = (int)d;
  pa.nw = (int)nw;
  pa.j_begin = j_begin;
  pa.j_count = j_count;
  pa.negate = negate;
  pa.count = count;
  unsigned __int128 any = 0;
  for (size_t k = 0; k < nw; k++) {
    pa.w[k] = (uint64_t)w[k];
    pa.whi[k] = (uint64_t)(w[k] >> 64);
    any |= w[k];
  }
  pa.top_bit = -1;
  if (any >> 64)
    pa.top_bit = 127 - __builtin_clzll((uint64_t)(any >> 64));
  else if (any)
    pa.top_bit = 63 - __builtin_clzll((uint64_t)any);
  if (is_l2) {
    GtArr A = gt_alloc(c, nin), R = gt_alloc(c, nout);
    gt_from_bytes(c, di, nin, A);
    pa.x = A.re;
    pa.y = A.im;
    pa.X = R.re;
    pa.Y = R.im;
    {
      Timer t(c, "k_gt_polyconv");
      c->Co->gt_polyconv(cfg(c, nblk(nout, 128), 128, 0), pa);
      t.done();
    }
    gt_to_bytes(c, R, nout, ob.dev);
  } else {
    G1Arr A = g1_alloc(c, nin), R = g1_alloc(c, nout);
    g1_from_bytes(c, di, nin, A);
    JacArr j = jac_alloc(c, nout);
    uint32_t* scratch = arena_get<uint32_t>(c, nout * c->L);
    pa.x = A.x;
    pa.y = A.y;
    pa.inf = A.inf;
    pa.X = j.X;
    pa.Y = j.Y;
    pa.Z = j.Z;
    {
      Timer t(c, "k_g1_polyconv");
      c->Bo->g1_polyconv(cfg(c, nblk(nout, 128), 128, 0), pa);
      t.done();
    }
    normalize_soa(c, j, nout, scratch, R);
    g1_to_bytes(c, R, nout, ob.dev);
  }
  commit_out(c, ob);
}

int bgn_multconstpoly_batch(bgn_ctx* c, const uint8_t* in, size_t d, int is_l2, const uint8_t* digits, size_t nd,
                            int negate, size_t count, uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!in || !out || !digits || !d || !nd || nd > BGN_CONV_MAXW || d > 4096) throw ArgErr{"bad argument"};
    unsigned __int128 w[BGN_CONV_MAXW];
    for (size_t k = 0; k < nd; k++) w[k] = digits[k];
    polyconv(c, in, d, is_l2, w, nd, 0, (int)(d + nd), negate, count, out);
  });
}

int bgn_evalpoly_batch(bgn_ctx* c, const uint8_t* in, size_t d, int is_l2, uint32_t base, size_t count, uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!in || !out || !d || d > BGN_CONV_MAXW || base < 2) throw ArgErr{"bad argument"};
    // sum_i base^i c_i as a correlation: w[k] = base^(d-1-k), output slot j = d-1 only
    unsigned __int128 w[BGN_CONV_MAXW];
    unsigned __int128 pw = 1;
    const unsigned __int128 lim = ~(unsigned __int128)0 / base;
    for (size_t i = 0; i < d; i++) {
      w[d - 1 - i] = pw;
      if (i + 1 < d) {
        if (pw > lim) throw ArgErr{"base^(d-1) does not fit 128 bits"};
        pw *= base;
      }
    }
    polyconv(c, in, d, is_l2, w, d, (int)d - 1, 1, 0, count, out);
  });
}

int bgn_make_poly_l2_batch(bgn_ctx* c, const uint8_t* in, size_t d, size_t count, uint8_t* out) {
  return guarded(c, [&] {
    if (!count) return;
    if (!in || !out || !d) throw ArgErr{"bad argument"};
    size_t nin = count * d, nout = count * (d + 1);
    check_count(nout);
    ensure_linesP(c);
    arena_reserve(c, io_bytes(c, nin) + io_bytes(c, nout) + g1_bytes(c, nin) + gt_bytes(c, nin) +
                         miller_scratch(c, nin, 1) + 8192);
    OutBuf ob = stage_out(c, out, nout * 2 * c->B);
    const uint8_t* da = stage_in(c, in, nin * 2 * c->B);
    G1Arr A = g1_alloc(c, nin);
    g1_from_bytes(c, da, nin, A);
    GtArr R = gt_alloc(c, nin);
    if (c->linesP && c->fixed_lines) {
      run_miller_fixed(c, A, nin, R);
    } else {
      G1Arr Pv{c->dPx, c->dPy, c->dPinf, 1};
      run_miller(c, A, 1, Pv, 1, 1, nin, 1, R);
    }
    gt_to_bytes(c, R, nout, ob.dev, (int)d, 1);
    commit_out(c, ob);
  });
}

int bgn_ctx_set_secret(bgn_ctx* c, const uint8_t* q1_be, size_t q1_len, uint64_t msg_space, uint32_t baby_steps) {
  return guarded(c, [&] {
    if (!q1_be || !q1_len || q1_len > 4 * (BGN_MAX_EXPW - 1) || msg_space == 0 || msg_space > ((uint64_t)1 << 50))
      throw ArgErr{"bad secret key / message space"};
    Big q = big_from_be(q1_be, q1_len, BGN_MAX_EXPW);
    int qb = big_bits(q);
    if (qb == 0) throw ArgErr{"q1 is zero"};
    for (int i = 0; i < BGN_MAX_EXPW; i++) c->pc.exp[i] = q[i];
    c->pc.exp_bits = qb;
    {
      std::vector<int8_t> qn = big_naf(Big(q.begin(), q.begin() + (qb + 31) / 32));
      if (qn.size() > BGN_MAX_NAF) throw ArgErr{"secret exponent too large"};
      c->pc.exp_naf_len = (int)qn.size();
      for (size_t i = 0; i < qn.size(); i++) c->pc.exp_naf[i] = qn[i];
    }
    reupload_pc(c);
    c->has_secret = false;
    // bound = ceil(sqrt(T)) exactly as gsbs.go:60 (float64 sqrt of an int64)
    uint64_t bound = (uint64_t)ceil(sqrt((double)(int64_t)msg_space));
    uint64_t mmax = bound * bound + bound + 2;  // i <= bound, table value j+1 <= bound+2 (gsbs.go:44, 77-98)
    uint64_t S = baby_steps ? baby_steps : std::min<uint64_t>(mmax, (uint64_t)1 << 21);
    // gsk = e(P,P)^q1 has order q2 = n / q1, and gsk^j, gsk^(q2-j) share their real part -- the key the
    // table is hashed on and the only thing the Lucas path sees.  Keep the table to j <= (q2-1)/2 so that
    // no two entries share a real part (and none repeats); the giant steps cover the rest.  Only toy keys
    // or message spaces near q2 are affected: for q2 >= 2^26 the table (at most 2^24 entries) never gets there.
    {
      Big nn = c->n_limbs;
      Big qq(BGN_MAXL, 0);
      for (int i = 0; i < BGN_MAX_EXPW && i < BGN_MAXL; i++) qq[i] = q[i];
      Big q2 = big_div(nn, qq);
      if (big_bits(q2) <= 26) {
        uint64_t q2v = q2[0];
        uint64_t lim = q2v > 2 ? (q2v - 1) / 2 : 1;
        if (S > lim) S = lim;
      }
    }
    if (S < 1) S = 1;
    if (S > ((uint64_t)1 << 24)) throw ArgErr{"baby_steps too large"};
    uint64_t hs = 1;
    while (hs < 2 * S) hs <<= 1;
    cudaFree(c->bs_elems);
    cudaFree(c->bs_slots);
    cudaFree(c->bs_ginv);
    c->bs_elems = c->bs_slots = c->bs_ginv = nullptr;
    CK(cudaMalloc(&c->bs_elems, S * 2 * c->L * 4));
    CK(cudaMalloc(&c->bs_slots, hs * 4));
    CK(cudaMalloc(&c->bs_ginv, 4 * (size_t)c->L * 4));
    CK(cudaMemsetAsync(c->bs_slots, 0, hs * 4, c->stream));
    // gsk = e(P,P)^q1   (bgn.go:198-199)
    arena_reserve(c, 4 * gt_bytes(c, 1) + miller_scratch(c, 1, 1) + 8192);
    G1Arr Pv{c->dPx, c->dPy, c->dPinf, 1};
    GtArr e = gt_alloc(c, 1), gsk = gt_alloc(c, 1), gs = gt_alloc(c, 1);
    run_miller(c, Pv, 1, Pv, 1, 0, 1, 1, e);
    GtPowArgs pa;
    pa.re = e.re;
    pa.im = e.im;
    pa.Nin = 1;
    pa.e_be = nullptr;
    pa.ebytes = 0;
    pa.mode = 1;
    pa.ore = gsk.re;
    pa.oim = gsk.im;
    pa.count = 1;
    pa.N = 1;
    {
      Timer t(c, "k_gt_pow");
      c->A->gt_pow(cfg(c, 1, 32, 0), pa);
      t.done();
    }
    // generator in AoS for the table kernels: bs_ginv[0..2L) temporarily holds gsk, then gen^-S
    uint32_t* gen_aos = c->bs_ginv + 2 * c->L;
    CK(cudaMemcpyAsync(gen_aos, gsk.re, c->L * 4, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(gen_aos + c->L, gsk.im, c->L * 4, cudaMemcpyDeviceToDevice, c->stream));
    BsgsBuildArgs ba;
    ba.gen = gen_aos;
    ba.elems = c->bs_elems;
    ba.slots = c->bs_slots;
    ba.hmask = (uint32_t)(hs - 1);
    ba.S = (uint32_t)S;
    ba.chunk = 64;
    size_t nthreads = (S + ba.chunk - 1) / ba.chunk;
    {
      Timer t(c, "k_bsgs_build");
      c->A->bsgs_build(cfg(c, nblk(nthreads, 128), 128, 0), ba);
      t.done();
    }
    // ginv = conj(gsk^S): exponent S as 4 big-endian bytes
    uint8_t sbe[4] = {(uint8_t)(S >> 24), (uint8_t)(S >> 16), (uint8_t)(S >> 8), (uint8_t)S};
    uint8_t* dsbe = arena_get<uint8_t>(c, 4);
    CK(cudaMemcpyAsync(dsbe, sbe, 4, cudaMemcpyHostToDevice, c->stream));
    pa.re = gsk.re;
    pa.im = gsk.im;
    pa.e_be = dsbe;
    pa.ebytes = 4;
    pa.mode = 3;  // variable exponent, conjugated result
    pa.ore = gs.re;
    pa.oim = gs.im;
    {
      Timer t(c, "k_gt_pow");
      c->A->gt_pow(cfg(c, 1, 32, 0), pa);
      t.done();
    }
    CK(cudaMemcpyAsync(c->bs_ginv, gs.re, c->L * 4, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(c->bs_ginv + c->L, gs.im, c->L * 4, cudaMemcpyDeviceToDevice, c->stream));
    c->bs_S = (uint32_t)S;
    c->bs_hmask = (uint32_t)(hs - 1);
    c->bs_mmax = mmax;
    c->bs_giant = (uint32_t)((mmax + S - 1) / S);
    // Level-1 Decrypt needs e(C, P)^q1 = e(C, q1 P): with the line table of q1*P the exponentiation is part of
    // the pairing.  (The reference forms gsk = P^q1 too, bgn.go:222; its table search then runs in G1.)
    cudaFree(c->linesPq);
    c->linesPq = nullptr;
    if (c->fixed_lines && c->dec_pair_q1) {
      uint32_t* tmp = nullptr;  // k_be | Jacobian (3L) | scratch (L) | affine x, y (2L) | flag
      const size_t kb = pad256(q1_len);
      CK(cudaMalloc(&tmp, kb + 6 * (size_t)c->L * 4 + 256));
      try {
        uint8_t* dk = reinterpret_cast<uint8_t*>(tmp);
        uint32_t* w = reinterpret_cast<uint32_t*>(dk + kb);
        JacArr j{w, w + c->L, w + 2 * c->L, 1};
        uint32_t* scr = w + 3 * c->L;
        G1Arr gq{w + 4 * c->L, w + 5 * c->L, reinterpret_cast<uint8_t*>(w + 6 * c->L), 1};
        CK(cudaMemcpyAsync(dk, q1_be, q1_len, cudaMemcpyHostToDevice, c->stream));
        G1MulArgs ma;
        ma.x = c->dPx;
        ma.y = c->dPy;
        ma.inf = c->dPinf;
        ma.Nin = 1;
        ma.k_be = dk;
        ma.kbytes = (int)q1_len;
        ma.X = j.X;
        ma.Y = j.Y;
        ma.Z = j.Z;
        ma.count = 1;
        ma.N = 1;
        c->Bo->g1_mulvar(cfg(c, 1, 32, 0), ma);
        normalize_soa(c, j, 1, scr, gq);
        uint8_t isinf = 1;
        CK(cudaMemcpyAsync(&isinf, gq.inf, 1, cudaMemcpyDeviceToHost, c->stream));
        finish(c);
        if (!isinf) c->linesPq = record_lines(c, gq.x, gq.y);
      } catch (...) {
        cudaFree(tmp);
        throw;
      }
      cudaFree(tmp);
    }
    c->has_secret = true;
  });
}

// R = C^q1 (count elements) -> plaintexts by the baby-step table and the giant steps (gsbs.go:54-106)
static void lookup_core(bgn_ctx* c, const GtArr& R, size_t count, int64_t* d_out, uint8_t* d_status) {
  BsgsLookupArgs la;
  la.re = R.re;
  la.im = R.im;
  la.Nin = R.N;
  la.count = count;
  la.elems = c->bs_elems;
  la.slots = c->bs_slots;
  la.hmask = c->bs_hmask;
  la.S = c->bs_S;
  la.ginv = c->bs_ginv;
  la.giant_steps = c->bs_giant;
  la.mmax = c->bs_mmax;
  la.out = d_out;
  la.status = d_status;
  Timer t(c, "k_bsgs_lookup");
  c->A->bsgs_lookup(cfg(c, nblk(count, 128), 128, 0), la);
  t.done();
}
// level-2 elements A (count) -> plaintexts: Lucas ladder + one table probe, or C^q1 + giant steps
static void decrypt_core(bgn_ctx* c, const GtArr& A, size_t count, int64_t* d_out, uint8_t* d_status) {
  if (c->bs_giant == 1 && c->dec_lucas) {
    // the whole message space is in the baby-step table: Lucas ladder on the trace, a pair of
    // lanes per ciphertext, search by real part (lucas.cuh)
    DecLucasArgs da;
    da.re = A.re;
    da.im = A.im;
    da.count = count;
    da.elems = c->bs_elems;
    da.slots = c->bs_slots;
    da.hmask = c->bs_hmask;
    da.S = c->bs_S;
    da.mmax = c->bs_mmax;
    da.out = d_out;
    da.status = d_status;
    Timer t(c, "k_dec_lucas");
    c->Co->dec_lucas(cfg(c, nblk(2 * count, 64), 64, 0), da);
    t.done();
    return;
  }
  GtArr R = gt_alloc(c, count);
  GtPowArgs pa;
  pa.re = A.re;
  pa.im = A.im;
  pa.Nin = A.N;
  pa.e_be = nullptr;
  pa.ebytes = 0;
  pa.mode = 1;
  pa.ore = R.re;
  pa.oim = R.im;
  pa.count = count;
  pa.N = R.N;
  {
    Timer t(c, "k_gt_pow");
    c->A->gt_pow(cfg(c, nblk(count, 128), 128, 0), pa);
    t.done();
  }
  lookup_core(c, R, count, d_out, d_status);
}
// level-1 elements -> plaintexts: e(C, q1 P) through the line table of q1*P and the table search, or
// e(C, P) and the level-2 path
static void make_l2_core(bgn_ctx* c, const G1Arr& C1, size_t count, const GtArr& out);
static void decrypt_l1_core(bgn_ctx* c, const G1Arr& C1, size_t count, const GtArr& A, int64_t* d_out, uint8_t* d_status) {
  if (c->linesPq && c->fixed_lines && c->dec_pair_q1) {
    run_miller_fixed(c, C1, count, A, c->linesPq);
    lookup_core(c, A, count, d_out, d_status);
    return;
  }
  make_l2_core(c, C1, count, A);
  decrypt_core(c, A, count, d_out, d_status);
}
// e(C[i], P) for level-1 elements (line table of P when enabled)
static void make_l2_core(bgn_ctx* c, const G1Arr& C1, size_t count, const GtArr& out) {
  if (c->linesP && c->fixed_lines) {
    run_miller_fixed(c, C1, count, out);
  } else {
    G1Arr Pv{c->dPx, c->dPy, c->dPinf, 1};
    run_miller(c, C1, 1, Pv, 1, 1, count, 1, out);
  }
}

int bgn_decrypt_batch(bgn_ctx* c, const uint8_t* in, int is_l2, size_t count, int64_t* out, uint8_t* status) {
  if (c && !c->has_secret) {
    c->err = "DL tables not computed!";
    return BGN_E_NOTSETUP;
  }
  return guarded(c, [&] {
    if (!count) return;
    if (!in || !out || !status) throw ArgErr{"null buffer"};
    check_count(count);
    if (!is_l2) ensure_linesP(c);
    arena_reserve(c, io_bytes(c, count) + g1_bytes(c, count) + 3 * gt_bytes(c, count) + pad256(count * 8) +
                         pad256(count) + miller_scratch(c, count, 1) + 8192);
    const uint8_t* di = stage_in(c, in, count * 2 * c->B);
    OutBuf oo = stage_out(c, out, count * 8);
    OutBuf os = stage_out(c, status, count);
    GtArr A = gt_alloc(c, count);
    if (is_l2) {
      gt_from_bytes(c, di, count, A);
    } else {
      // level 1: e(C, P)^q1 = e(P,P)^(q1 m); same m as the reference's G1 table search (bgn.go:222-223)
      G1Arr C1 = g1_alloc(c, count);
      g1_from_bytes(c, di, count, C1);
      decrypt_l1_core(c, C1, count, A, reinterpret_cast<int64_t*>(oo.dev), os.dev);
    }
    if (is_l2) decrypt_core(c, A, count, reinterpret_cast<int64_t*>(oo.dev), os.dev);
    commit_out(c, oo);
    commit_out(c, os);
  });
}

// ---------------------------------------------------------------- device-resident batches
// A bgn_buf is a batch of group elements kept on the device in the kernels' own form (Montgomery limbs,
// [count][L] per coordinate, G1 with its infinity flags).  Chained operations -- Encrypt -> EAdd -> EMult
// -> L2 sum -> Decrypt -- then run kernel to kernel: the byte format, its Montgomery conversion and the
// on-curve check (k_g1_from_bytes alone costs as much as the addition it feeds) are paid once at the edge.
}  // extern "C"  (the struct and helpers below are C++)

struct bgn_buf {
  bgn_ctx* owner = nullptr;
  int kind = 0;  // BGN_KIND_G1 / BGN_KIND_GT
  size_t count = 0, cap = 0;
  uint32_t *a = nullptr, *b = nullptr;  // x, y  /  re, im
  uint8_t* inf = nullptr;               // G1 only
};

namespace {
void buf_release(bgn_buf* h) {
  if (!h) return;
  cudaFree(h->a);
  cudaFree(h->inf);
  delete h;
}
// *out: reused when it belongs to this context, has the kind and is large enough; else (re)allocated
bgn_buf* buf_prepare(bgn_ctx* c, bgn_buf** out, int kind, size_t count) {
  if (!out) throw ArgErr{"null output handle"};
  bgn_buf* h = *out;
  if (h && (h->owner != c || h->kind != kind || h->cap < count)) {
    if (h->owner != c) throw ArgErr{"output handle belongs to another context"};
    buf_release(h);
    h = nullptr;
    *out = nullptr;
  }
  if (!h) {
    h = new bgn_buf();
    h->owner = c;
    h->kind = kind;
    h->cap = std::max<size_t>(count, 1);
    size_t words = h->cap * (size_t)c->L;
    cudaError_t e = cudaMalloc(&h->a, 2 * words * 4);
    if (e == cudaSuccess && kind == BGN_KIND_G1) e = cudaMalloc(&h->inf, h->cap);
    if (e != cudaSuccess) {
      buf_release(h);
      cudaGetLastError();
      throw CudaErr{std::string("cudaMalloc (handle): ") + cudaGetErrorString(e)};
    }
    h->b = h->a + words;
    *out = h;
  }
  h->count = count;
  return h;
}
const bgn_buf* buf_check(bgn_ctx* c, const bgn_buf* h, int kind, const char* what) {
  if (!h) throw ArgErr{std::string("null handle: ") + what};
  if (h->owner != c) throw ArgErr{std::string("handle of another context: ") + what};
  if (h->kind != kind) throw ArgErr{std::string("handle holds the other group: ") + what};
  return h;
}
G1Arr as_g1(const bgn_buf* h) { return G1Arr{h->a, h->b, h->inf, h->count}; }
GtArr as_gt(const bgn_buf* h) { return GtArr{h->a, h->b, h->count}; }
}  // namespace

extern "C" {
int bgn_buf_import(bgn_ctx* c, int kind, const uint8_t* bytes, size_t count, bgn_buf** out) {
  return guarded(c, [&] {
    if (kind != BGN_KIND_G1 && kind != BGN_KIND_GT) throw ArgErr{"kind must be BGN_KIND_G1 or BGN_KIND_GT"};
    if (count && !bytes) throw ArgErr{"null buffer"};
    check_count(count);
    bgn_buf* h = buf_prepare(c, out, kind, count);
    if (!count) return;
    arena_reserve(c, io_bytes(c, count) + 4096);
    const uint8_t* d = stage_in(c, bytes, count * 2 * c->B);
    if (kind == BGN_KIND_G1)
      g1_from_bytes(c, d, count, as_g1(h));
    else
      gt_from_bytes(c, d, count, as_gt(h));
  });
}
int bgn_buf_export(bgn_ctx* c, const bgn_buf* h, uint8_t* bytes_out) {
  return guarded(c, [&] {
    if (!h || h->owner != c) throw ArgErr{"bad handle"};
    if (!h->count) return;
    if (!bytes_out) throw ArgErr{"null buffer"};
    arena_reserve(c, io_bytes(c, h->count) + 4096);
    OutBuf ob = stage_out(c, bytes_out, h->count * 2 * c->B);
    if (h->kind == BGN_KIND_G1)
      g1_to_bytes(c, as_g1(h), h->count, ob.dev);
    else
      gt_to_bytes(c, as_gt(h), h->count, ob.dev);
    commit_out(c, ob);
  });
}
int bgn_buf_info(const bgn_buf* h, int* kind, size_t* count) {
  if (!h) return BGN_E_BADARG;
  if (kind) *kind = h->kind;
  if (count) *count = h->count;
  return BGN_OK;
}
void bgn_buf_free(bgn_buf* h) {
  if (!h) return;
  bgn_ctx* c = h->owner;
  std::lock_guard<std::mutex> lk(dev_mu(c->device));
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  buf_release(h);
}

int bgn_encrypt_h(bgn_ctx* c, const int64_t* x, const uint8_t* r_be, size_t count, bgn_buf** out) {
  return guarded(c, [&] {
    if (count && !x) throw ArgErr{"null buffer"};
    check_count(count);
    bgn_buf* h = buf_prepare(c, out, BGN_KIND_G1, count);
    if (!count) return;
    arena_reserve(c, pad256(count * 8) + pad256(count * c->nbytes) + jac_bytes(c, count) + pad256(count * c->L * 4) + 4096);
    const int64_t* dx = reinterpret_cast<const int64_t*>(stage_in(c, x, count * 8));
    const uint8_t* dr = r_be ? stage_in(c, r_be, count * c->nbytes) : nullptr;
    JacArr j = jac_alloc(c, count);
    uint32_t* scratch = arena_get<uint32_t>(c, count * c->L);
    EncArgs ea;
    ea.x = dx;
    ea.r_be = dr;
    ea.rbytes = c->nbytes;
    enc_tables(c, ea, dr != nullptr);
    ea.X = j.X;
    ea.Y = j.Y;
    ea.Z = j.Z;
    ea.count = count;
    ea.N = count;
    ea.bx = ea.by = nullptr;
    ea.binf = nullptr;
    {
      Timer t(c, "k_encrypt");
      c->Bo->encrypt(cfg(c, nblk(count, 128), 128, 0), ea);
      t.done();
    }
    normalize_soa(c, j, count, scratch, as_g1(h));
  });
}

int bgn_g1_add_h(bgn_ctx* c, const bgn_buf* a, const bgn_buf* b, int subtract, bgn_buf** out) {
  return guarded(c, [&] {
    buf_check(c, a, BGN_KIND_G1, "a");
    buf_check(c, b, BGN_KIND_G1, "b");
    if (a->count != b->count) throw ArgErr{"operands differ in count"};
    if (out && (*out == a || *out == b)) throw ArgErr{"the output handle must not be an operand"};
    bgn_buf* h = buf_prepare(c, out, BGN_KIND_G1, a->count);
    if (!a->count) return;
    arena_reserve(c, jac_bytes(c, a->count) + 2 * pad256(a->count * c->L * 4) + 8192);
    g1_add_core(c, as_g1(a), as_g1(b), a->count, subtract ? 1 : 0, 0, as_g1(h));
  });
}

int bgn_gt_mul_h(bgn_ctx* c, const bgn_buf* a, const bgn_buf* b, int divide, bgn_buf** out) {
  return guarded(c, [&] {
    buf_check(c, a, BGN_KIND_GT, "a");
    buf_check(c, b, BGN_KIND_GT, "b");
    if (a->count != b->count) throw ArgErr{"operands differ in count"};
    bgn_buf* h = buf_prepare(c, out, BGN_KIND_GT, a->count);
    if (!a->count) return;
    GtBinArgs ga;
    ga.are = a->a;
    ga.aim = a->b;
    ga.Na = a->count;
    ga.bre = b->a;
    ga.bim = b->b;
    ga.Nb = b->count;
    ga.conj_b = divide ? 1 : 0;
    ga.ore = h->a;
    ga.oim = h->b;
    ga.count = a->count;
    ga.N = h->count;
    Timer t(c, "k_gt_mul");
    c->A->gt_mul(cfg(c, nblk(a->count, 128), 128, 0), ga);
    t.done();
  });
}

int bgn_pair_h(bgn_ctx* c, const bgn_buf* a, const bgn_buf* b, bgn_buf** out) {
  return guarded(c, [&] {
    buf_check(c, a, BGN_KIND_G1, "a");
    if (b) buf_check(c, b, BGN_KIND_G1, "b");
    if (b && a->count != b->count) throw ArgErr{"operands differ in count"};
    bgn_buf* h = buf_prepare(c, out, BGN_KIND_GT, a->count);
    if (!a->count) return;
    if (!b) ensure_linesP(c);
    arena_reserve(c, miller_scratch(c, a->count, 1) + 8192);
    if (!b) {
      make_l2_core(c, as_g1(a), a->count, as_gt(h));  // makeL2: e(a, P)
      return;
    }
    const size_t cap = pair_duo_capacity(c);
    if (cap && (c->pair_duo > 0 || (c->pair_duo < 0 && a->count <= cap)))
      run_pair_duo(c, as_g1(a), as_g1(b), a->count, as_gt(h));
    else
      run_miller(c, as_g1(a), 1, as_g1(b), 1, 0, a->count, 1, as_gt(h));
  });
}

int bgn_multpoly_h(bgn_ctx* c, const bgn_buf* c1, size_t d1, const bgn_buf* c2, size_t d2, size_t count, bgn_buf** out) {
  return guarded(c, [&] {
    buf_check(c, c1, BGN_KIND_G1, "c1");
    buf_check(c, c2, BGN_KIND_G1, "c2");
    if (!d1 || !d2 || d1 > 128 || d2 > 128) throw ArgErr{"bad slot count"};
    if (c1->count != count * d1 || c2->count != count * d2) throw ArgErr{"handle sizes do not match count x slots"};
    check_count(count * (d1 + d2));
    bgn_buf* h = buf_prepare(c, out, BGN_KIND_GT, count * (d1 + d2));
    if (!count) return;
    arena_reserve(c, miller_scratch(c, count, (int)std::max(d1, d2)) + 8192);
    if (d1 <= d2)
      run_miller(c, as_g1(c1), (int)d1, as_g1(c2), (int)d2, 0, count, (int)(d1 + d2), as_gt(h));
    else
      run_miller(c, as_g1(c2), (int)d2, as_g1(c1), (int)d1, 0, count, (int)(d1 + d2), as_gt(h));
  });
}

int bgn_l2_sum_reduce_h(bgn_ctx* c, const bgn_buf* in, size_t nterms, size_t ncoeff, bgn_buf** out) {
  return guarded(c, [&] {
    buf_check(c, in, BGN_KIND_GT, "in");
    if (!ncoeff || ncoeff > (1u << 20) || in->count != nterms * ncoeff) throw ArgErr{"handle size does not match nterms x ncoeff"};
    if (out && *out == in) throw ArgErr{"the output handle must not be the operand"};
    bgn_buf* h = buf_prepare(c, out, BGN_KIND_GT, ncoeff);
    arena_reserve(c, 3 * gt_bytes(c, nterms * ncoeff / 8 + ncoeff + 64) + 65536);
    GtArr R;
    if (nterms == 0) {
      R = gt_alloc(c, ncoeff);
      Timer t(c, "k_gt_reduce");
      c->A->gt_reduce(cfg(c, nblk(ncoeff, 128), 128, 0), R.re, R.im, R.N, 0, (int)ncoeff, 1, R.re, R.im, R.N);
      t.done();
    } else {
      R = reduce_tree(c, as_gt(in), nterms, ncoeff);
    }
    CK(cudaMemcpyAsync(h->a, R.re, ncoeff * c->L * 4, cudaMemcpyDeviceToDevice, c->stream));
    CK(cudaMemcpyAsync(h->b, R.im, ncoeff * c->L * 4, cudaMemcpyDeviceToDevice, c->stream));
  });
}

int bgn_decrypt_h(bgn_ctx* c, const bgn_buf* in, int64_t* out, uint8_t* status) {
  if (c && !c->has_secret) {
    c->err = "DL tables not computed!";
    return BGN_E_NOTSETUP;
  }
  return guarded(c, [&] {
    if (!in || in->owner != c) throw ArgErr{"bad handle"};
    const size_t count = in->count;
    if (!count) return;
    if (!out || !status) throw ArgErr{"null buffer"};
    const bool l2 = in->kind == BGN_KIND_GT;
    if (!l2) ensure_linesP(c);
    arena_reserve(c, 3 * gt_bytes(c, count) + pad256(count * 8) + pad256(count) + miller_scratch(c, count, 1) + 8192);
    OutBuf oo = stage_out(c, out, count * 8);
    OutBuf os = stage_out(c, status, count);
    GtArr A = l2 ? as_gt(in) : gt_alloc(c, count);
    if (l2)
      decrypt_core(c, A, count, reinterpret_cast<int64_t*>(oo.dev), os.dev);
    else
      decrypt_l1_core(c, as_g1(in), count, A, reinterpret_cast<int64_t*>(oo.dev), os.dev);
    commit_out(c, oo);
    commit_out(c, os);
  });
}

// ---------------------------------------------------------------- several GPUs behind one handle
// SURVEY.md 8(b) proposed bgn_ctx_create(params, ndev, devs): "multi-GPU is driven inside one call by
// per-device host threads".  A bgn_group is that: one context per listed device, and batch entry points
// that cut the batch into contiguous shards (the reference's independent units, poly.go:15, 37, 140-141),
// run every shard on its device from its own host thread, and -- for the one operation with a cross-unit
// dependency, the L2 sum behind an inner product -- fold the per-device partial products on device 0.
// Buffers are HOST pointers (a device pointer belongs to one GPU).
}  // extern "C"

#include <thread>

struct bgn_group {
  std::vector<bgn_ctx*> ctx;
  std::string err;
};

namespace {
// shard [lo, hi) of `count` units for member g of n: sizes differ by at most one
void group_shard(size_t count, size_t g, size_t n, size_t* lo, size_t* hi) {
  size_t base = count / n, rem = count % n;
  *lo = g * base + std::min(g, rem);
  *hi = *lo + base + (g < rem ? 1 : 0);
}
// run fn(g, ctx) on every member from its own thread; the first failing status wins
template <typename Fn>
int group_run(bgn_group* grp, Fn fn) {
  if (!grp || grp->ctx.empty()) return BGN_E_BADARG;
  const size_t n = grp->ctx.size();
  std::vector<int> st(n, BGN_OK);
  try {
    std::vector<std::thread> th;
    th.reserve(n);
    for (size_t g = 1; g < n; g++) th.emplace_back([&, g] { st[g] = fn(g, grp->ctx[g]); });
    st[0] = fn(0, grp->ctx[0]);
    for (auto& t : th) t.join();
  } catch (...) {
    grp->err = "could not start the per-device host threads";
    return BGN_E_NOMEM;
  }
  for (size_t g = 0; g < n; g++)
    if (st[g] != BGN_OK) {
      grp->err = "device " + std::to_string(grp->ctx[g]->device) + ": " + grp->ctx[g]->err;
      return st[g];
    }
  return BGN_OK;
}
}  // namespace

extern "C" {
int bgn_group_create(const bgn_params* prm, int ndev, const int* devs, bgn_group** out) {
  g_last_error.clear();
  if (!out || ndev < 1 || ndev > kMaxDevices || !devs) {
    g_last_error = "bgn_group_create: bad argument";
    return BGN_E_BADARG;
  }
  *out = nullptr;
  bgn_group* grp = nullptr;
  try {
    grp = new bgn_group();
  } catch (...) {
    return BGN_E_NOMEM;
  }
  // contexts are built one after the other: table construction is seconds at most, and the per-thread
  // error string of a failed bgn_ctx_create stays readable on this thread
  for (int i = 0; i < ndev; i++) {
    bgn_ctx* c = nullptr;
    int st = bgn_ctx_create(prm, devs[i], &c);
    if (st != BGN_OK) {
      for (bgn_ctx* d : grp->ctx) bgn_ctx_destroy(d);
      delete grp;
      return st;
    }
    try {
      grp->ctx.push_back(c);
    } catch (...) {
      bgn_ctx_destroy(c);
      for (bgn_ctx* d : grp->ctx) bgn_ctx_destroy(d);
      delete grp;
      return BGN_E_NOMEM;
    }
  }
  *out = grp;
  return BGN_OK;
}
void bgn_group_destroy(bgn_group* grp) {
  if (!grp) return;
  for (bgn_ctx* c : grp->ctx) bgn_ctx_destroy(c);
  delete grp;
}
int bgn_group_size(const bgn_group* grp) { return grp ? (int)grp->ctx.size() : 0; }
bgn_ctx* bgn_group_ctx(bgn_group* grp, int i) {
  return (grp && i >= 0 && (size_t)i < grp->ctx.size()) ? grp->ctx[(size_t)i] : nullptr;
}
const char* bgn_group_last_error(const bgn_group* grp) { return grp ? grp->err.c_str() : g_last_error.c_str(); }

int bgn_group_set_secret(bgn_group* grp, const uint8_t* q1_be, size_t q1_len, uint64_t msg_space, uint32_t baby_steps) {
  return group_run(grp, [&](size_t, bgn_ctx* c) { return bgn_ctx_set_secret(c, q1_be, q1_len, msg_space, baby_steps); });
}
int bgn_group_set_option(bgn_group* grp, const char* name, long value) {
  return group_run(grp, [&](size_t, bgn_ctx* c) { return bgn_ctx_set_option(c, name, value); });
}

int bgn_group_encrypt_batch(bgn_group* grp, const int64_t* x, const uint8_t* r_be, size_t count, uint8_t* out) {
  return group_run(grp, [&](size_t g, bgn_ctx* c) {
    size_t lo, hi;
    group_shard(count, g, grp->ctx.size(), &lo, &hi);
    return bgn_encrypt_batch(c, x + lo, r_be ? r_be + lo * c->nbytes : nullptr, hi - lo, out + lo * 2 * c->B);
  });
}
int bgn_group_g1_add_batch(bgn_group* grp, const uint8_t* a, const uint8_t* b, size_t count, uint8_t* out) {
  return group_run(grp, [&](size_t g, bgn_ctx* c) {
    size_t lo, hi;
    group_shard(count, g, grp->ctx.size(), &lo, &hi);
    const size_t eb = 2 * (size_t)c->B;
    return bgn_g1_add_batch(c, a + lo * eb, b + lo * eb, hi - lo, out + lo * eb);
  });
}
int bgn_group_multpoly_batch(bgn_group* grp, const uint8_t* c1, size_t d1, const uint8_t* c2, size_t d2, size_t count,
                             uint8_t* out) {
  return group_run(grp, [&](size_t g, bgn_ctx* c) {
    size_t lo, hi;
    group_shard(count, g, grp->ctx.size(), &lo, &hi);
    const size_t eb = 2 * (size_t)c->B;
    return bgn_multpoly_batch(c, c1 + lo * d1 * eb, d1, c2 + lo * d2 * eb, d2, hi - lo, out + lo * (d1 + d2) * eb);
  });
}
int bgn_group_decrypt_batch(bgn_group* grp, const uint8_t* in, int is_l2, size_t count, int64_t* out, uint8_t* status) {
  return group_run(grp, [&](size_t g, bgn_ctx* c) {
    size_t lo, hi;
    group_shard(count, g, grp->ctx.size(), &lo, &hi);
    return bgn_decrypt_batch(c, in + lo * 2 * (size_t)c->B, is_l2, hi - lo, out + lo, status + lo);
  });
}
// sum_i c1[i] * c2[i] as ONE polynomial ciphertext of d1 + d2 level-2 slots (BASELINE config 5): every device
// multiplies its shard (MultPoly) and reduces it to one partial (GT product tree); the partials -- d1 + d2
// elements per device -- are folded on device 0.
int bgn_group_inner_product(bgn_group* grp, const uint8_t* c1, size_t d1, const uint8_t* c2, size_t d2, size_t count,
                            uint8_t* out) {
  if (!grp || grp->ctx.empty() || !out) return BGN_E_BADARG;
  const size_t n = grp->ctx.size(), nc = d1 + d2, eb = 2 * (size_t)grp->ctx[0]->B;
  std::vector<uint8_t> parts, prod;
  try {
    parts.resize(n * nc * eb);
    prod.resize(count * nc * eb);
  } catch (...) {
    grp->err = "host allocation failed";
    return BGN_E_NOMEM;
  }
  int st = group_run(grp, [&](size_t g, bgn_ctx* c) {
    size_t lo, hi;
    group_shard(count, g, n, &lo, &hi);
    uint8_t* pr = prod.data() + lo * nc * eb;
    int s1 = (hi > lo) ? bgn_multpoly_batch(c, c1 + lo * d1 * eb, d1, c2 + lo * d2 * eb, d2, hi - lo, pr) : BGN_OK;
    if (s1 != BGN_OK) return s1;
    return bgn_l2_sum_reduce(c, hi > lo ? pr : nullptr, hi - lo, nc, parts.data() + g * nc * eb);
  });
  if (st != BGN_OK) return st;
  st = bgn_l2_sum_reduce(grp->ctx[0], parts.data(), n, nc, out);
  if (st != BGN_OK) grp->err = grp->ctx[0]->err;
  return st;
}

// ---------------------------------------------------------------- instrumentation
int bgn_timing_enable(bgn_ctx* c, int on) {
  if (!c) return BGN_E_BADARG;
  std::lock_guard<std::mutex> lk(dev_mu(c->device));
  c->timing = on != 0;
  return BGN_OK;
}
int bgn_timing_reset(bgn_ctx* c) {
  if (!c) return BGN_E_BADARG;
  std::lock_guard<std::mutex> lk(dev_mu(c->device));
  c->ktimes.clear();
  c->total_launches = 0;
  return BGN_OK;
}
int bgn_timing_get(bgn_ctx* c, const char* prefix, double* ms_total, uint64_t* launches) {
  if (!c || !prefix) return BGN_E_BADARG;
  std::lock_guard<std::mutex> lk(dev_mu(c->device));
  double ms = 0;
  uint64_t n = 0;
  size_t pl = strlen(prefix);
  for (auto& kv : c->ktimes)
    if (kv.first.compare(0, pl, prefix) == 0) {
      ms += kv.second.ms;
      n += kv.second.launches;
    }
  if (pl == 0) n = c->total_launches;
  if (ms_total) *ms_total = ms;
  if (launches) *launches = n;
  return BGN_OK;
}

int bgn_timing_last_call(bgn_ctx* c, double* ms) {
  if (!c || !ms) return BGN_E_BADARG;
  std::lock_guard<std::mutex> lk(dev_mu(c->device));
  *ms = c->last_call_ms;
  return BGN_OK;
}

int bgn_bench_mulmod(bgn_ctx* c, int ilp, int iters, int blocks, int threads, float* ms) {
  return guarded(c, [&] {
    if (!ms || iters <= 0 || blocks <= 0 || threads <= 0 || threads > (ilp >= 10 ? 256 : 128))
      throw ArgErr{"bad argument"};
    size_t N = (size_t)blocks * threads;
    arena_reserve(c, pad256(N * c->L * 4) + 4096);
    uint32_t* io = arena_get<uint32_t>(c, N * c->L);
    CK(cudaMemsetAsync(io, 0x5a, N * c->L * 4, c->stream));
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    auto launch = [&](int it) {
      c->total_launches++;
      if (ilp != 1 && ilp != 2 && (ilp < 10 || ilp > 89)) throw ArgErr{"ilp must be 1, 2 or a primitive mode 10..89"};
      c->A->mulmod_bench(cfg(c, blocks, threads, 0), ilp, io, N, it);
    };
    launch(4);  // warm-up
    CK(cudaEventRecord(a, c->stream));
    launch(iters);
    CK(cudaEventRecord(b, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaGetLastError());
    CK(cudaEventElapsedTime(ms, a, b));
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  });
}

__global__ void __launch_bounds__(256) k_imad_peak(uint32_t* out, int iters, uint32_t seed) {
  uint32_t a = seed + threadIdx.x, b = seed * 3 + blockIdx.x;
  uint32_t r[16];
#pragma unroll
  for (int i = 0; i < 16; i++) r[i] = a + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int rep = 0; rep < 8; rep++) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        // 8 independent 64-bit accumulators, one IMAD.WIDE.U32 each
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;"
                     : "+r"(r[2 * k]), "+r"(r[2 * k + 1])
                     : "r"(a), "r"(b));
      }
    }
  }
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) o ^= r[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = o;
}

// mix: 0 WIDE x8 | 1 (LO,HI) x8 | 2 WIDE x8 + (LO,HI) x8 | 3 FFMA x8 | 4 WIDE x8 + FFMA x8 | 5 DFMA x8 |
//      6 WIDE x8 + DFMA x8 | 7 WIDE x8 + (LO,HI) x4 | 8 WIDE x8 + DFMA x4 | 9 WIDE x4 + DFMA x8 |
//      10 LO x8 | 11 HI x8 | 12 WIDE x8 + LO x8 | 13 WIDE x8 + HI x8
// per_thread[5] = instructions of each class a thread issues: {IMAD.WIDE, IMAD.LO, IMAD.HI, FFMA, DFMA}
int bgn_bench_issue_mix(int device, int mix, int iters, int blocks, int threads, float* ms, double* per_thread) {
  if (!ms || !per_thread || iters <= 0 || blocks <= 0 || threads <= 0 || threads > 256 || device < 0) return BGN_E_BADARG;
  if (mix < 0 || mix > 13) return BGN_E_BADARG;
  std::lock_guard<std::mutex> lk(dev_mu(device));
  if (cudaSetDevice(device) != cudaSuccess) return BGN_E_CUDA;
  uint32_t* d = nullptr;
  if (cudaMalloc(&d, (size_t)blocks * threads * 4) != cudaSuccess) return BGN_E_CUDA;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  int nw = 0, nlo = 0, nhi = 0, nf = 0, nd = 0;
  auto go = [&](int it) {
    switch (mix) {
#define BGN_MIX(id, W_, LO_, HI_, F_, D_)                                  \
  case id:                                                                 \
    nw = W_, nlo = LO_, nhi = HI_, nf = F_, nd = D_;                       \
    k_issue_mix<W_, LO_, HI_, F_, D_><<<blocks, threads>>>(d, it, 12345u); \
    break;
      BGN_MIX(0, 8, 0, 0, 0, 0)
      BGN_MIX(1, 0, 8, 8, 0, 0)
      BGN_MIX(2, 8, 8, 8, 0, 0)
      BGN_MIX(3, 0, 0, 0, 8, 0)
      BGN_MIX(4, 8, 0, 0, 8, 0)
      BGN_MIX(5, 0, 0, 0, 0, 8)
      BGN_MIX(6, 8, 0, 0, 0, 8)
      BGN_MIX(7, 8, 4, 4, 0, 0)
      BGN_MIX(8, 8, 0, 0, 0, 4)
      BGN_MIX(9, 4, 0, 0, 0, 8)
      BGN_MIX(10, 0, 8, 0, 0, 0)
      BGN_MIX(11, 0, 0, 8, 0, 0)
      BGN_MIX(12, 8, 8, 0, 0, 0)
      BGN_MIX(13, 8, 0, 8, 0, 0)
#undef BGN_MIX
      default:
        break;
    }
  };
  go(8);
  cudaEventRecord(a);
  go(iters);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  cudaEventElapsedTime(ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  per_thread[0] = (double)iters * 8 * nw;
  per_thread[1] = (double)iters * 8 * nlo;
  per_thread[2] = (double)iters * 8 * nhi;
  per_thread[3] = (double)iters * 8 * nf;
  per_thread[4] = (double)iters * 8 * nd;
  return e == cudaSuccess ? BGN_OK : BGN_E_CUDA;
}

int bgn_bench_imad_peak(int device, int iters, int blocks, int threads, float* ms, double* instr_per_thread) {
  if (!ms || iters <= 0 || blocks <= 0 || threads <= 0 || threads > 256 || device < 0) return BGN_E_BADARG;
  std::lock_guard<std::mutex> lk(dev_mu(device));
  if (cudaSetDevice(device) != cudaSuccess) return BGN_E_CUDA;
  uint32_t* d = nullptr;
  if (cudaMalloc(&d, (size_t)blocks * threads * 4) != cudaSuccess) return BGN_E_CUDA;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k_imad_peak<<<blocks, threads>>>(d, 16, 12345u);
  cudaEventRecord(a);
  k_imad_peak<<<blocks, threads>>>(d, iters, 12345u);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  cudaEventElapsedTime(ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  cudaFree(d);
  if (instr_per_thread) *instr_per_thread = (double)iters * 64.0;
  return e == cudaSuccess ? BGN_OK : BGN_E_CUDA;
}

}  // extern "C"
