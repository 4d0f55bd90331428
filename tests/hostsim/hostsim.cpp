// hostsim.cpp -- CPU execution of the *device* programs (bgn_b200/csrc/*.cuh
// compiled with BGN_HOSTSIM: carry flag, atomics and the thread grid emulated
// in software).  TEST INFRASTRUCTURE ONLY: it lets the CPU test-suite check the
// kernels' per-thread logic against the oracle without a GPU.  It is never
// linked into libbgn_b200.so and nothing in bgn_b200/ can reach it.
#define BGN_HOSTSIM 1
#ifndef BGN_L
#define BGN_L 17  // only selects the default loop shape of the fused routines (pairing.cuh)
#endif
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "../../bgn_b200/csrc/kernels.cuh"

namespace {
// Two lanes of a lane-pair kernel (pairlane.cuh) as two threads that run STRICTLY one at a time: a lane
// holds the mutex while it computes and gives it up only inside an exchange, where it waits for its
// partner's value -- the warp shuffle of the CUDA kernel.  The program text is the device's own
// (`run` with an exchange functor); only the exchange differs.  Bounds travel with the values.
template <int L>
struct LanePair {
  std::mutex m;
  std::condition_variable cv;
  uint32_t box[2][2][L];
  double bnd[2][2];
  long deposited[2] = {0, 0};
  void xchg(int s, uint32_t (&other)[L], const uint32_t (&mine)[L], std::unique_lock<std::mutex>& lk) {
    const long r = deposited[s];
    for (int j = 0; j < L; j++) box[s][r & 1][j] = mine[j];
    bnd[s][r & 1] = BGN_GETB(mine);
    deposited[s] = r + 1;
    cv.notify_all();
    cv.wait(lk, [&] { return deposited[1 - s] > r; });
    for (int j = 0; j < L; j++) other[j] = box[1 - s][r & 1][j];
    BGN_SETB(other, bnd[1 - s][r & 1]);
  }
  template <typename Fn>
  void run(Fn lane_program) {
    auto body = [&](int s) {
      std::unique_lock<std::mutex> lk(m);
      lane_program(s, [&](uint32_t (&other)[L], const uint32_t (&mine)[L]) { xchg(s, other, mine, lk); });
      cv.notify_all();
    };
    std::thread t1(body, 1);
    body(0);
    t1.join();
  }
};

template <int L, bool EG = false>
void sim_miller(const MillerArgs& a0, int nblocks, int nt) {
  std::vector<uint32_t> smem(MillerTeam<L, EG>::smem_words(nt) + 8);
  std::vector<uint32_t> priv((size_t)nblocks * MillerTeam<L, EG>::priv_words() + 8);
  MillerArgs a = a0;
  a.priv = priv.data();
  std::vector<uint32_t> evw((size_t)(a.e_bcast ? a.dE : a.count * a.dE) * L + 8);
  a.evw = evw.data();
  for (int b = 0; b < nblocks; b++) {
    std::vector<MillerTeam<L, EG>> T;
    T.reserve(nt);
    for (int tid = 0; tid < nt; tid++) T.emplace_back(a, smem.data(), tid, b, nt);
    for (auto& t : T) t.init();
    int n = c_pc.naf_len;
    for (int idx = 1; idx < n; idx++) {
      int d = c_pc.naf[idx];
      if (MillerTeam<L, EG>::PARA && a.para && d != 0 && idx != n - 1) {
        for (auto& t : T) t.phaseA_dadd(d > 0 ? MOP_ADD : MOP_SUB);
        for (auto& t : T) t.phaseB_para();
        continue;
      }
      for (auto& t : T) t.phaseA(MOP_DBL, idx == 1);
      for (auto& t : T) t.phaseB();
      if (d != 0 && idx != n - 1) {
        for (auto& t : T) t.phaseA(d > 0 ? MOP_ADD : MOP_SUB, false);
        for (auto& t : T) t.phaseB();
      }
    }
    for (auto& t : T) t.finalize();
  }
}
// k_miller_split (teamsplit.cuh): the same phase schedule as sim_miller, 2 dE threads per team
template <int L>
void sim_miller_split(const MillerArgs& a0, int nblocks, int nt) {
  MillerArgs a = a0;
  std::vector<uint32_t> smem(MillerSplit<L>::smem_words(nt, a.teams_per_group * a.dE) + 8);
  std::vector<uint32_t> evw((size_t)(a.e_bcast ? a.dE : a.count * a.dE) * L + 8);
  a.evw = evw.data();
  for (int b = 0; b < nblocks; b++) {
    std::vector<MillerSplit<L>> T;
    T.reserve(nt);
    for (int tid = 0; tid < nt; tid++) T.emplace_back(a, smem.data(), tid, b, nt);
    for (auto& t : T) t.init();
    int n = c_pc.naf_len;
    for (int idx = 1; idx < n; idx++) {
      int d = c_pc.naf[idx];
      if (MillerSplit<L>::PARA && a.para && d != 0 && idx != n - 1) {
        for (auto& t : T) t.phaseA_dadd(d > 0 ? MOP_ADD : MOP_SUB);
        for (auto& t : T) t.phaseB_para();
        continue;
      }
      for (auto& t : T) t.phaseA(MOP_DBL, idx == 1);
      for (auto& t : T) t.phaseB();
      if (d != 0 && idx != n - 1) {
        for (auto& t : T) t.phaseA(d > 0 ? MOP_ADD : MOP_SUB, false);
        for (auto& t : T) t.phaseB();
      }
    }
    for (auto& t : T) t.finalize();
  }
}
// k_pair_duo (pairwarp.cuh): the X warp runs one step AHEAD of the F warp -- the most the one
// barrier per step allows -- so the run also checks the double buffering of the published values
template <int L>
void sim_pair_duo(const PairDuoArgs& a, int np) {
  std::vector<uint32_t> smem(MillerDuo<L>::smem_words(np) + 8);
  std::vector<int> ops;
  MillerDuo<L>::for_steps([&](int, int op) { ops.push_back(op); });  // (default loop shape of this build)
  const int S = (int)ops.size();
  for (int b = 0; b * np < a.count; b++) {
    std::vector<MillerDuo<L>> T;
    T.reserve(np);
    for (int pt = 0; pt < np; pt++) T.emplace_back(a, smem.data(), np, pt, (size_t)b * np + pt);
    for (auto& t : T) t.x_init();
    for (auto& t : T) t.f_init();
    for (auto& t : T) t.x_step(ops[0], 0);
    for (int s = 0; s < S; s++) {
      if (s + 1 < S)
        for (auto& t : T) t.x_step(ops[s + 1], (s + 1) & 1);
      for (auto& t : T) t.f_step(ops[s], s & 1, s == 0);
    }
    for (auto& t : T) t.f_finish();
  }
}
}  // namespace

#define FOR_L(L, ...)                          \
  switch (L) {                                 \
    case 3: { constexpr int LL = 3; __VA_ARGS__; } break;   \
    case 5: { constexpr int LL = 5; __VA_ARGS__; } break;   \
    case 17: { constexpr int LL = 17; __VA_ARGS__; } break; \
    default: return -1;                        \
  }                                            \
  return 0;

extern "C" {
void hs_set_consts(const FieldConsts* fc, const PairConsts* pc, int L) {
  c_fc = *fc;
  c_pc = *pc;
  long double pd = 0, R = 1;
  for (int j = L - 1; j >= 0; j--) pd = pd * 4294967296.0L + c_fc.p[j];
  for (int j = 0; j < L; j++) R *= 4294967296.0L;
  bgnsim::bnd.clear();
  // the limb-count rule (32L >= bits(p) + 8) guarantees R/p >= 256; the tracker proves the ranges
  // for that worst case, not just for this key's (usually much larger) headroom
  bgnsim::headroom = (double)(R / pd) < 256.0 ? (double)(R / pd) : 256.0;
  bgnsim::setb(c_fc.p, 1.0);
  bgnsim::setb(c_fc.p2, 2.0);
  bgnsim::setb(c_fc.p4, 4.0);
  bgnsim::setb(c_fc.p8, 8.0);
  bgnsim::setb(c_fc.p16, 16.0);
  bgnsim::setb(c_fc.one, 1.0);
  bgnsim::setb(c_fc.r2, 1.0);
  bgnsim::max_bound = 0;
  bgnsim::unknown = bgnsim::violations = 0;
}
// range tracker report: out[0] = largest bound attached (multiples of p), out[1] = headroom R/p,
// out[2] = reads of untracked addresses, out[3] = violations
void hs_range_report(double* out, int reset) {
  out[0] = bgnsim::max_bound;
  out[1] = bgnsim::headroom;
  out[2] = (double)bgnsim::unknown;
  out[3] = (double)bgnsim::violations;
  if (reset) {
    bgnsim::max_bound = 0;
    bgnsim::unknown = bgnsim::violations = 0;
  }
}
// inputs built by the test driver (Montgomery arrays of canonical values): bound 1 per element
void hs_track_array(const uint32_t* a, size_t count, int L, double bound) {
  for (size_t e = 0; e < count; e++) bgnsim::setb(a + e * L, bound);
}
int hs_miller(int L, const MillerArgs* a, int nblocks, int nt) { FOR_L(L, sim_miller<LL>(*a, nblocks, nt)) }
#if !BGN_MILLER_GP
int hs_miller_wide(int L, const MillerArgs* a, int nblocks, int nt) { FOR_L(L, sim_miller<LL, true>(*a, nblocks, nt)) }
#else
int hs_miller_wide(int, const MillerArgs*, int, int) { return -1; }
#endif
int hs_miller_split(int L, const MillerArgs* a, int nblocks, int nt) { FOR_L(L, sim_miller_split<LL>(*a, nblocks, nt)) }
int hs_encrypt(int L, const EncArgs* a) { FOR_L(L, for (size_t e = 0; e < a->count; e++) encrypt_body<LL>(*a, e)) }
int hs_normalize(int L, const NormArgs* a) { FOR_L(L, for (size_t g = 0; g < (size_t)a->G; g++) normalize_body<LL>(*a, g)) }
int hs_g1_add(int L, const G1AddArgs* a) { FOR_L(L, for (size_t e = 0; e < a->count; e++) g1_add_body<LL>(*a, e)) }
int hs_g1_affadd(int L, const G1AffAddArgs* a) { FOR_L(L, for (size_t g = 0; g < (size_t)a->G; g++) g1_affadd_body<LL>(*a, g)) }
int hs_g1_mulvar(int L, const G1MulArgs* a) { FOR_L(L, for (size_t e = 0; e < a->count; e++) g1_mulvar_body<LL>(*a, e)) }
int hs_gt_mul(int L, const GtBinArgs* a) { FOR_L(L, for (size_t e = 0; e < a->count; e++) gt_mul_body<LL>(*a, e)) }
int hs_gt_pow(int L, const GtPowArgs* a) { FOR_L(L, for (size_t e = 0; e < a->count; e++) gt_pow_body<LL>(*a, e)) }
int hs_gt_reduce(int L, const uint32_t* re, const uint32_t* im, size_t Nin, size_t nterms, int ncoeff, int G,
                 uint32_t* ore, uint32_t* oim, size_t N) {
  FOR_L(L, for (size_t id = 0; id < (size_t)G * ncoeff; id++)
               gt_reduce_body<LL>(re, im, Nin, nterms, ncoeff, G, ore, oim, N, id))
}
int hs_bsgs_build(int L, const BsgsBuildArgs* a) {
  FOR_L(L, for (size_t g = 0; g * a->chunk < a->S; g++) bsgs_build_body<LL>(*a, g))
}
int hs_bsgs_lookup(int L, const BsgsLookupArgs* a) {
  FOR_L(L, for (size_t e = 0; e < a->count; e++) bsgs_lookup_body<LL>(*a, e))
}
int hs_tab_bases(int L, const uint32_t* bx, const uint32_t* by, int nwin, int hb, uint32_t* X, uint32_t* Y, uint32_t* Z,
                 size_t N) {
  FOR_L(L, tab_bases_body<LL>(bx, by, nwin, hb, X, Y, Z, N, 0))
}
int hs_tab_fill(int L, const uint32_t* ax, const uint32_t* ay, const uint8_t* ainf, size_t Nb, int nwin, int hb, uint32_t* X,
                uint32_t* Y, uint32_t* Z, size_t N) {
  FOR_L(L, for (int w = 0; w < nwin; w++) tab_fill_body<LL>(ax, ay, ainf, Nb, nwin, hb, X, Y, Z, N, w))
}
int hs_tabw_fill(int L, const uint32_t* tabh, int nwin_h, int nsub, int hb, uint32_t* X, uint32_t* Y, uint32_t* Z,
                 size_t first, size_t nent) {
  FOR_L(L, for (size_t id = 0; id < nent; id++) tabw_fill_body<LL>(tabh, nwin_h, nsub, hb, X, Y, Z, first, nent, id))
}
int hs_tab_edwards(int L, const uint32_t* tabw, uint32_t* tabe, uint32_t* scratch, size_t count, int G, int* bad) {
  FOR_L(L, for (int g = 0; g < G; g++) tab_edwards_body<LL>(tabw, tabe, scratch, count, G, bad, (size_t)g))
}
int hs_g1_from_bytes(int L, const uint8_t* in, int B, size_t count, uint32_t* x, uint32_t* y, uint8_t* inf, size_t N) {
  FOR_L(L, for (size_t e = 0; e < count; e++) g1_from_bytes_body<LL>(in, B, count, x, y, inf, N, e))
}
int hs_g1_to_bytes(int L, const uint32_t* x, const uint32_t* y, const uint8_t* inf, size_t N, size_t count, uint8_t* out,
                   int B) {
  FOR_L(L, for (size_t e = 0; e < count; e++) g1_to_bytes_body<LL>(x, y, inf, N, count, out, B, e))
}
int hs_fp2_from_bytes(int L, const uint8_t* in, int B, size_t count, uint32_t* re, uint32_t* im, size_t N) {
  FOR_L(L, for (size_t e = 0; e < count; e++) fp2_from_bytes_body<LL>(in, B, count, re, im, N, e))
}
int hs_fp2_to_bytes(int L, const uint32_t* re, const uint32_t* im, size_t N, size_t count, uint8_t* out, int B, int grp,
                    int pad) {
  FOR_L(L, for (size_t e = 0; e < count; e++) fp2_to_bytes_body<LL>(re, im, N, count, out, B, grp, pad, e))
}
int hs_gt_blind(int L, const GtBlindArgs* a) { FOR_L(L, for (size_t e = 0; e < a->count; e++) gt_blind_body<LL>(*a, e)) }
int hs_gt_tab_bases(int L, const uint32_t* gen, int nwin, uint32_t* bases) {
  FOR_L(L, gt_tab_bases_body<LL>(gen, nwin, bases, 0))
}
int hs_gt_tab_fill(int L, const uint32_t* bases, int nwin, uint32_t* tab) {
  FOR_L(L, for (int w = 0; w < nwin; w++) gt_tab_fill_body<LL>(bases, nwin, tab, w))
}
int hs_g1_polyconv(int L, const PolyConvArgs* a) {
  FOR_L(L, for (size_t id = 0; id < a->count * (size_t)a->j_count; id++) g1_polyconv_body<LL>(*a, id))
}
int hs_gt_polyconv(int L, const PolyConvArgs* a) {
  FOR_L(L, for (size_t id = 0; id < a->count * (size_t)a->j_count; id++) gt_polyconv_body<LL>(*a, id))
}
int hs_miller_fixed(int L, const MillerFixedArgs* a, int nt) {
  FOR_L(L, {
    std::vector<uint32_t> smem(MillerFixed<LL>::smem_words(nt) + 8);
    for (int e = 0; e < a->count; e++) MillerFixed<LL>::run(*a, smem.data(), e % nt, nt, (size_t)e);
  })
}
int hs_miller_fixed_pair(int L, const MillerFixedArgs* a) {
  FOR_L(L, for (int e = 0; e < a->count; e++) {
    LanePair<LL> lp;
    lp.run([&](int s, auto xchg) { MillerFixedPair<LL>::run(*a, (size_t)e, s, true, xchg); });
  })
}
int hs_pair_duo(int L, const PairDuoArgs* a, int np) { FOR_L(L, sim_pair_duo<LL>(*a, np)) }
int hs_miller_record(int L, const uint32_t* px, const uint32_t* py, uint32_t* lines, uint32_t* scratch, int* ok) {
  FOR_L(L, MillerFixed<LL>::record(px, py, lines, scratch, ok))
}
int hs_miller_nsteps(int L) { FOR_L(L, return MillerFixed<LL>::nsteps(c_pc)) }
int hs_gt_pow_pair(int L, const GtPowArgs* a) { FOR_L(L, for (size_t e = 0; e < a->count; e++) gt_pow_pair_sim<LL>(*a, e)) }
int hs_dec_lucas(int L, const DecLucasArgs* a) {
  FOR_L(L, for (size_t e = 0; e < a->count; e++) dec_lucas_pair_sim<LL>(*a, e))
}
// tracker self-test: a difference whose subtrahend may exceed its offset must be flagged
int hs_selftest_violation() {
  uint64_t before = bgnsim::violations;
  uint32_t a[3] = {5, 0, 0}, b[3] = {1, 0, 0}, r[3];
  bgnsim::setb(a, 1.0);
  bgnsim::setb(b, 3.0);  // b may be as large as 3p, offset is only 2p
  Fp<3>::subk(r, a, b, c_fc.p2, 2);
  int fired = bgnsim::violations > before;
  bgnsim::violations = before;
  return fired;
}
// double-width products / separate reductions executed (lazy reduction, fused.cuh)
void hs_wide_count(uint64_t* out, int reset) {
  out[0] = bgnsim::nmulw;
  out[1] = bgnsim::nredc;
  out[2] = bgnsim::nmulk;
  out[3] = bgnsim::ndot2;
  out[4] = bgnsim::nsqrw;
  out[5] = bgnsim::ndot3;
  if (reset) bgnsim::nmulw = bgnsim::nredc = bgnsim::nmulk = bgnsim::ndot2 = bgnsim::nsqrw = bgnsim::ndot3 = 0;
}
uint64_t hs_mul_count(int reset) {
  uint64_t v = bgnsim::nmul;
  if (reset) bgnsim::nmul = 0;
  return v;
}
// r[i] = a[i]^2 through the dedicated squaring and through the product: both must agree limb for limb
int hs_fp_sqr(int L, uint32_t* r_sqr, uint32_t* r_mul, const uint32_t* a, size_t count) {
  FOR_L(L, for (size_t i = 0; i < count; i++) {
    uint32_t x[LL], y[LL], z[LL];
    ld<LL>(x, a + i * LL);
    Fp<LL>::sqr(y, x);
    Fp<LL>::mul(z, x, x);
    st<LL>(r_sqr + i * LL, y);
    st<LL>(r_mul + i * LL, z);
  })
}
int hs_fp_inv(int L, uint32_t* r, const uint32_t* a) { FOR_L(L, Loc<LL> t; F<LL>::inv(r, a, t.v())) }
int hs_fp_inv_gcd(int L, uint32_t* r, const uint32_t* a) { FOR_L(L, F<LL>::inv_gcd(r, a)) }
// the verified fast inversion; *fallbacks receives how often it had to fall back since the last call
int hs_fp_inv_safe(int L, uint32_t* r, const uint32_t* a, size_t count, uint64_t* fallbacks) {
  bgnsim::safegcd_fallbacks = 0;
  int rc = 0;
  switch (L) {
    case 3: for (size_t i = 0; i < count; i++) F<3>::inv_gcd_fast(r + i * 3, a + i * 3); break;
    case 5: for (size_t i = 0; i < count; i++) F<5>::inv_gcd_fast(r + i * 5, a + i * 5); break;
    case 17: for (size_t i = 0; i < count; i++) F<17>::inv_gcd_fast(r + i * 17, a + i * 17); break;
    default: rc = -1;
  }
  *fallbacks = bgnsim::safegcd_fallbacks;
  return rc;
}
}
