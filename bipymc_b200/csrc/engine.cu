// Host side of the C-ABI (include/bipymc_b200.h): owns the per-handle workspace,
// turns a generation of DeMcMpi._mcmc_run (bipymc/demc.py:79-135) into kernel launches
// on the caller's stream, and never touches the CPU for arithmetic.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <utility>
#include <vector>

#include "../../include/bipymc_b200.h"
#include "kernels_generic.cuh"
#include "kernels_fused.cuh"
#include "diagnostics.cuh"
#include "gauss_dmma.cuh"
#include "sync.cuh"
#include <cub/device/device_radix_sort.cuh>

namespace {

thread_local std::string g_err;

int fail(const std::string& m) {
  g_err = m;
  return 1;
}

#define CU_TRY(expr)                                                                       \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return fail(std::string(#expr) + ": " + cudaGetErrorString(e__));                    \
  } while (0)

#define BPM_TRY(expr)      \
  do {                     \
    int r__ = (expr);      \
    if (r__) return r__;   \
  } while (0)

inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

inline int pick_lpc(int d) {
  if (d <= 4) return 1;
  if (d <= 16) return 4;
  if (d <= 32) return 8;
  return 32;
}

}  // namespace

struct bpm_engine {
  bpm_config cfg;
  int nA;  // ceil(N / 2): np.array_split gives the first half the extra element
  // workspace (device)
  int32_t* perm = nullptr;
  int32_t* flip = nullptr;
  int32_t* inv = nullptr;        // inverse permutation (only when the packed lists are in use)
  int32_t* loc_list = nullptr;   // local chains of both halves, packed in chain order (kernels_generic.cuh)
  int32_t* loc_cnt = nullptr;    // [2]
  int32_t* phase_cnt = nullptr;  // [1] rows of the phase being stepped (selected from loc_cnt on the device)
  int32_t* cmp_blk = nullptr;    // [2][nblk][2] block counts / offsets
  // Second set of split buffers + a side stream: inside a multi-generation call the shuffle / list packing of
  // generation g+1 -- functions of (seed, generation) only -- are enqueued on the side stream while generation g
  // runs, so they execute in the tail of its last fused launch instead of between two launches (11 us of a
  // 250 us generation on one GPU, ~30 us of 340 on eight).  BIPYMC_B200_NO_SIDE=1 switches it off.
  struct SplitBuf { int32_t *perm = nullptr, *flip = nullptr, *inv = nullptr, *loc_list = nullptr, *loc_cnt = nullptr,
                             *cmp_blk = nullptr; } alt;
  cudaStream_t side = nullptr;
  cudaEvent_t ev_start = nullptr, ev_split = nullptr, ev_done = nullptr;
  cudaEvent_t ev_chunk[8] = {nullptr};       // sharded host entry: chunked host copy -> peer forwarding
  void swap_split() {
    std::swap(perm, alt.perm); std::swap(flip, alt.flip); std::swap(inv, alt.inv);
    std::swap(loc_list, alt.loc_list); std::swap(loc_cnt, alt.loc_cnt); std::swap(cmp_blk, alt.cmp_blk);
  }
  int side_setup() {
    if (side) return 0;
    const int N = cfg.n_chains;
    CU_TRY(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
    CU_TRY(cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&ev_split, cudaEventDisableTiming));
    CU_TRY(cudaEventCreateWithFlags(&ev_done, cudaEventDisableTiming));
    CU_TRY(cudaMalloc(&alt.perm, sizeof(int32_t) * N));
    CU_TRY(cudaMalloc(&alt.flip, sizeof(int32_t)));
    CU_TRY(cudaMemset(alt.flip, 0, sizeof(int32_t)));
    if (packed()) {
      const int nblk = cdiv(cfg.chain_hi - cfg.chain_lo, bpm::kCompactBlock);
      CU_TRY(cudaMalloc(&alt.inv, sizeof(int32_t) * N));
      CU_TRY(cudaMalloc(&alt.loc_list, sizeof(int32_t) * N));
      CU_TRY(cudaMalloc(&alt.loc_cnt, sizeof(int32_t) * 2));
      CU_TRY(cudaMalloc(&alt.cmp_blk, sizeof(int32_t) * 4 * nblk));
    }
    return 0;
  }
  bool serial() const { return cfg.algo == BPM_ALGO_DEMC_SERIAL; }
  bool sharded() const { return cfg.chain_lo != 0 || cfg.chain_hi != cfg.n_chains; }
  // packed, chain-ordered phase lists: always when sharded; unsharded for the 16/24/32-byte rows
  // of d <= 4, where list order would waste half of every sector (serial DE-MC walks chains in order anyway)
  bool packed() const { return sharded() || (cfg.dim <= 4 && !serial()); }
  double* prop = nullptr;
  double* lnl_prop = nullptr;
  double* cr_delta = nullptr;
  int32_t* cr_pick = nullptr;
  double* p_cr = nullptr;
  double* cr_dm = nullptr;
  double* cr_cnt = nullptr;
  double* cr_part = nullptr;
  double* cr_block = nullptr;
  unsigned int* cr_ticket = nullptr;
  unsigned long long* counters = nullptr;  // [0] accepted, [1] rejected
  int32_t* nan_flag = nullptr;
  // target
  int target = BPM_TARGET_EXTERNAL;
  bpm::BananaParams banana;
  bpm::BimodalParams bimodal;
  double* tparams = nullptr;  // device copy of gauss / linefit parameters
  double* wfrag = nullptr;    // Gaussian target: W in DMMA fragment order (w_fragment_kernel)
  int64_t n_tparams = 0;
  int gauss_r = 0, gauss_logpdf_flag = 0, gauss_mu_zero = 0;
  double gauss_c0 = 0.0;
  int linefit_M = 0;
  bpm_lnl_fn user_fn = nullptr;
  void* user_ptr = nullptr;
  // split-path generation context
  bpm::PhaseArgs cur;
  bool cur_replay = false;
  bool cur_lazy = false;       // this generation's bpm_phase launches follow the lazy protocol
  int cur_phases_run = 0;
  bool in_generation = false;
  int64_t cur_k_gen = 0;
  // host-entry buffers
  double* hX = nullptr;
  double* hL = nullptr;
  double* hMean = nullptr;       // host entry: per-chain running moments kept on the device across calls
  double* hM2 = nullptr;         // (derived state: the reference recomputes np.std(chain.chain), dream.py:128)
  int64_t h_mom_len = 0;         // rows they cover; 0 = restart at the next call
  int64_t h_pending = 0;
  int32_t* h_accept = nullptr;   // host entry: accept flags of a generation / rows that moved in this call
  int32_t* h_changed = nullptr;
  unsigned long long* h_nrows = nullptr;
  uint64_t last_d2h_bytes = 0;
  int fused_ok = 1;  // allow the fused fast paths
  double* peers[BPM_MAX_PEERS] = {nullptr};   // other ranks' X replicas mapped here (bpm_set_peers)
  int n_peers = 0;
  // peer-memory barrier / CR exchange (sync.cuh, bpm_set_sync): replaces the per-generation NCCL collectives
  bpm::SyncArgs sync;
  bool sync_on = false;
  unsigned long long sync_epoch = 0, sync_cr_seq = 0;
  int32_t* sync_err = nullptr;
  // diagnostics (diagnostics.cuh): Omega tracking, IQR reset, R-hat scratch
  double* omega_sum = nullptr;   // [N] sum of lnL per chain since tracking started
  double* omega_buf = nullptr;   // [2][N] Omega means / sorted copy
  double* diag_out = nullptr;    // [3] threshold, Q1, Q3
  int32_t* diag_i = nullptr;     // [0] best chain, [1] number of resets
  void* sort_tmp = nullptr;
  size_t sort_tmp_bytes = 0;
  double* rh_mean = nullptr;     // [n_local][ld] scratch for history moments
  double* rh_m2 = nullptr;
  double* rh_out = nullptr;      // [dim]
  bool omega_on = false;
  int64_t omega_cnt = 0;
  double* cov_acc = nullptr;     // [d + d*d] streaming sum / cross moments (bpm_cov_track)
  bool cov_on = false;
  int64_t cov_rows = 0;          // chain states accumulated (per local chain: generations)
  // optional per-kernel timing (bpm_profile): CUDA events around every launch, by kind
  struct Rec { int kind; cudaEvent_t a, b; };
  bool prof_on = false;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> ev_pool;
  cudaEvent_t get_event() {
    if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void prof_begin(int kind, cudaStream_t s) {
    if (!prof_on) return;
    Rec r; r.kind = kind; r.a = get_event(); r.b = get_event();
    cudaEventRecord(r.a, s);
    recs.push_back(r);
  }
  void prof_end(cudaStream_t s) {
    if (!prof_on) return;
    cudaEventRecord(recs.back().b, s);
  }

  ~bpm_engine() {
    cudaFree(inv); cudaFree(loc_list); cudaFree(loc_cnt); cudaFree(phase_cnt); cudaFree(cmp_blk); cudaFree(cr_ticket);
    cudaFree(alt.perm); cudaFree(alt.flip); cudaFree(alt.inv); cudaFree(alt.loc_list); cudaFree(alt.loc_cnt);
    cudaFree(alt.cmp_blk);
    if (side) cudaStreamDestroy(side);
    if (ev_start) { cudaEventDestroy(ev_start); cudaEventDestroy(ev_split); cudaEventDestroy(ev_done); }
    for (auto e : ev_chunk) if (e) cudaEventDestroy(e);
    cudaFree(perm); cudaFree(flip); cudaFree(prop); cudaFree(lnl_prop); cudaFree(cr_delta);
    cudaFree(cr_pick); cudaFree(p_cr); cudaFree(cr_dm); cudaFree(cr_cnt); cudaFree(cr_part); cudaFree(cr_block);
    cudaFree(counters); cudaFree(nan_flag); cudaFree(tparams); cudaFree(wfrag); cudaFree(hX); cudaFree(hL); cudaFree(hMean); cudaFree(hM2);
    cudaFree(h_accept); cudaFree(h_changed); cudaFree(h_nrows);
    cudaFree(omega_sum); cudaFree(omega_buf); cudaFree(diag_out); cudaFree(diag_i); cudaFree(sort_tmp);
    cudaFree(rh_mean); cudaFree(rh_m2); cudaFree(rh_out); cudaFree(sync_err); cudaFree(cov_acc);
    for (auto& r : recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    for (auto e : ev_pool) cudaEventDestroy(e);
  }

  int init() {
    const int N = cfg.n_chains;
    nA = serial() ? N : (N + 1) / 2;   // serial DE-MC: one "half" holding every chain
    CU_TRY(cudaSetDevice(cfg.device));
    CU_TRY(cudaMalloc(&perm, sizeof(int32_t) * N));
    CU_TRY(cudaMalloc(&flip, sizeof(int32_t)));
    if (packed()) {
      const int nblk = cdiv(cfg.chain_hi - cfg.chain_lo, bpm::kCompactBlock);
      CU_TRY(cudaMalloc(&inv, sizeof(int32_t) * N));
      CU_TRY(cudaMalloc(&loc_list, sizeof(int32_t) * N));
      CU_TRY(cudaMalloc(&loc_cnt, sizeof(int32_t) * 2));
      CU_TRY(cudaMalloc(&phase_cnt, sizeof(int32_t)));
      CU_TRY(cudaMalloc(&cmp_blk, sizeof(int32_t) * 4 * nblk));
    }
    CU_TRY(cudaMalloc(&prop, sizeof(double) * (size_t)nA * cfg.ld));
    CU_TRY(cudaMalloc(&lnl_prop, sizeof(double) * nA));
    CU_TRY(cudaMalloc(&cr_delta, sizeof(double) * N));
    CU_TRY(cudaMalloc(&cr_pick, sizeof(int32_t) * N));
    CU_TRY(cudaMalloc(&p_cr, sizeof(double) * BPM_MAX_CR));
    CU_TRY(cudaMalloc(&cr_dm, sizeof(double) * BPM_MAX_CR));
    CU_TRY(cudaMalloc(&cr_cnt, sizeof(double) * BPM_MAX_CR));
    CU_TRY(cudaMalloc(&cr_part, sizeof(double) * 2 * BPM_MAX_CR));
    CU_TRY(cudaMalloc(&cr_block, sizeof(double) * 2 * BPM_MAX_CR * bpm::kCrBlocks));
    CU_TRY(cudaMalloc(&cr_ticket, sizeof(unsigned int)));
    CU_TRY(cudaMemset(cr_ticket, 0, sizeof(unsigned int)));
    CU_TRY(cudaMalloc(&counters, sizeof(unsigned long long) * 2));
    CU_TRY(cudaMalloc(&nan_flag, sizeof(int32_t)));
    CU_TRY(cudaMemset(flip, 0, sizeof(int32_t)));
    CU_TRY(cudaMemset(cr_pick, 0xFF, sizeof(int32_t) * N));
    CU_TRY(cudaMemset(cr_delta, 0, sizeof(double) * N));
    CU_TRY(cudaMemset(counters, 0, sizeof(unsigned long long) * 2));
    CU_TRY(cudaMemset(nan_flag, 0, sizeof(int32_t)));
    CU_TRY(cudaMemset(cr_dm, 0, sizeof(double) * BPM_MAX_CR));
    CU_TRY(cudaMemset(cr_cnt, 0, sizeof(double) * BPM_MAX_CR));
    std::vector<double> p(BPM_MAX_CR, 0.0);
    for (int m = 0; m < cfg.n_cr; ++m) p[m] = 1.0 / cfg.n_cr;  // dream.py:114
    CU_TRY(cudaMemcpy(p_cr, p.data(), sizeof(double) * BPM_MAX_CR, cudaMemcpyHostToDevice));
    return 0;
  }

  // Parameter block of one half-phase of generation k_gen.
  bpm::PhaseArgs make_args(const bpm_state* st, int64_t k_gen, int phase, const bpm_replay* rp,
                           const bpm_trace_out* tr, bool lazy = false, bool fly = false) const {
    bpm::PhaseArgs a;
    memset(&a, 0, sizeof(a));
    a.X = st->X; a.lnl = st->lnl; a.mean = st->mean; a.m2 = st->m2;
    a.hist_base = (rp && st->history) ? st->history : nullptr;
    const size_t hstride = (size_t)(cfg.chain_hi - cfg.chain_lo) * cfg.ld;
    a.hist_row = st->history ? st->history + (size_t)st->hist_len * hstride : nullptr;
    a.lazy = lazy ? 1 : 0;
    a.pending = (lazy && st->pending) ? 1 : 0;
    a.hist_cur = (a.pending && st->history && st->hist_len > 0) ? st->history + (size_t)(st->hist_len - 1) * hstride
                                                                 : nullptr;
    a.perm = perm; a.flip = flip; a.phase = phase;
    a.loc_list = loc_list; a.loc_cnt = loc_cnt;     // both nullptr on an unsharded handle
    a.N = cfg.n_chains; a.nA = nA; a.d = cfg.dim; a.ld = cfg.ld;
    a.chain_lo = cfg.chain_lo; a.chain_hi = cfg.chain_hi;
    a.algo = serial() ? BPM_ALGO_DEMC : cfg.algo;     // same proposal arithmetic (demc.py:180-182 == samplers.py:284-286)
    a.serial = serial() ? 1 : 0;
    a.del_pairs = cfg.algo == BPM_ALGO_DREAM ? cfg.del_pairs : 1;
    a.n_cr = cfg.n_cr;
    if (cfg.algo == BPM_ALGO_DREAM) {
      a.gamma_jump = (k_gen % 5) == 0;   // dream.py:77
      a.gamma_p0 = 0.20;
    } else {
      a.gamma_jump = (k_gen % 10) == 0;  // demc.py:174
      a.gamma_p0 = 0.1;
    }
    if (serial()) a.gamma_jump = 0;                   // samplers.py:284 uses gamma as is
    a.gamma_fixed = cfg.gamma > 0.0 ? cfg.gamma : 2.38 / sqrt(2.0 * cfg.dim);   // demc.py:162
    a.gamma_num = cfg.gamma_scale * 2.38;                                        // dream.py:61
    a.eps = cfg.epsilon > 0.0 ? sqrt(cfg.epsilon * cfg.epsilon) : 0.0;           // util.py:11-14
    a.u_eps = cfg.u_epsilon > 0.0 ? cfg.u_epsilon : 0.0;
    a.adapt = cfg.algo == BPM_ALGO_DREAM && cfg.burnin_gen > k_gen && st->hist_len > cfg.n_cr_gen &&
              st->m2 != nullptr;                                                 // dream.py:92,124
    a.hist_len = st->hist_len;
    a.mom_len = st->mom_len > 0 ? st->mom_len : st->hist_len;
    a.inv_mom = 1.0 / (double)a.mom_len;
    a.inv_n1 = 1.0 / (double)(a.mom_len + 1);
    a.p_cr = p_cr; a.cr_delta = cr_delta; a.cr_pick = cr_pick;
    a.prop = prop; a.lnl_prop = lnl_prop;
    for (int p = 0; p < n_peers; ++p) a.peers[p] = peers[p];
    a.n_peers = n_peers;
    // MEASUREMENT ONLY (BIPYMC_B200_NO_PEER_STORES=1): drop the in-kernel peer stores of a sharded run to
    // separate their cost from that of the larger replicated population; the replicas then diverge.
    static const bool no_peer_stores = [] {
      const char* e = getenv("BIPYMC_B200_NO_PEER_STORES");
      return e && e[0] == '1';
    }();
    if (no_peer_stores) a.n_peers = 0;
    a.n_acc = counters; a.n_rej = counters + 1; a.nan_flag = nan_flag;
    a.rng = bpm::make_rng(cfg.seed, (uint64_t)st->hist_len);
    if (!serial()) {          // what begin() makes the device flag: the replayed coin, or the native one
      a.flip_known = 1;
      a.flip_val = rp ? (rp->flip ? 1 : 0) : host_flip(a.rng);
    }
    if (fly && !rp && !serial()) fill_fly(a);
    if (rp) a.rp = *rp;
    if (tr) a.tr = *tr;
    return a;
  }

  // bpm_generations_host: state block over the device copies of the caller's host arrays.  DREAM keeps
  // the chains' running moments here between calls, so crossover adaptation (dream.py:92,119-140) runs
  // end to end exactly as in a device-resident run; they restart when the entry is (re)started.
  int host_entry_state(bpm_state* st, int64_t g_abs0, size_t nx, cudaStream_t s) {
    memset(st, 0, sizeof(*st));
    st->X = hX; st->lnl = hL; st->hist_len = g_abs0;
    if (hMean) {
      if (h_mom_len == 0) {          // first call: the moments cover the one row the caller handed in
        CU_TRY(cudaMemcpyAsync(hMean, hX, nx, cudaMemcpyDeviceToDevice, s));
        CU_TRY(cudaMemsetAsync(hM2, 0, nx, s));
        h_mom_len = 1; h_pending = 0;
      }
      st->mean = hMean; st->m2 = hM2; st->mom_len = h_mom_len; st->pending = h_pending;
    }
    return 0;
  }

  // Does phase() run the lazy-protocol kernel for this handle?  (kernels_fused.cuh: fused_plan_is_v3)
  bool lazy_plan() const {
    return fused_ok && !serial() && bpm::fused_plan_is_v3(target, cfg.dim, cfg.ld, gauss_r, fused_ok);
  }
  // Native-RNG generations that never materialise the shuffle ("fly" mode, step.cuh) -- d <= 4 fused kernel on a
  // handle that owns every chain.  MEASURED AND SWITCHED OFF (profiles/r2/r2h_*): with the balanced 6-round
  // Feistel network and cycle walking (2.6 walks at N = 10^5) one list entry costs ~160 instructions, seven of
  // them per chain-step; the line-fit generation went 91 -> 132 us (and the 100-D kernel, which tried the same,
  // 114 -> 160 us).  A table lookup in the materialised shuffle (one 8 us split kernel per generation) is
  // cheaper.  BIPYMC_B200_FLY=1 re-enables the mode for d <= 4.
  bool fly_plan() const {
    static const bool on = [] {
      const char* e = getenv("BIPYMC_B200_FLY");
      return e && e[0] == '1';
    }();
    return on && fused_ok == 1 && !serial() && !sharded() && cfg.dim <= 4 &&
           (target == BPM_TARGET_BANANA || target == BPM_TARGET_BIMODAL || target == BPM_TARGET_LINEFIT);
  }
  // bpm_flush: history row hist_len - 1 and its moment sample, left pending by a lazy generation
  int flush(bpm_state* st, cudaStream_t s) {
    if (!st->pending) return 0;
    const int nloc = cfg.chain_hi - cfg.chain_lo;
    const int64_t mom = st->mom_len > 0 ? st->mom_len : st->hist_len;
    double* hist_cur = (st->history && st->hist_len > 0)
                           ? st->history + (size_t)(st->hist_len - 1) * nloc * cfg.ld : nullptr;
    if (st->mean || hist_cur) {
      bpm::flush_pending_kernel<<<cdiv((int64_t)nloc * cfg.ld, 256), 256, 0, s>>>(
          st->X, st->mean, st->m2, hist_cur, cfg.chain_lo, cfg.chain_hi, cfg.dim, cfg.ld, 1.0 / (double)mom);
      CU_TRY(cudaGetLastError());
    }
    st->pending = 0;
    return 0;
  }

  // fly mode: the generation's flip and permutation key, evaluated on the host (demc.py:81-86; the same
  // Philox calls and the same IEEE arithmetic split_native_kernel makes on the device)
  // the generation's flip coin on the host: the same Philox call and IEEE arithmetic as split_native_kernel
  int host_flip(const bpm::RngCtx& rng) const {
    const bpm::Philox4 q = bpm::draw4(rng, 0xFFFFFFFFu, bpm::RNG_GEN, 0);
    const double thr = cfg.flip / (cfg.flip + (1.0 - cfg.flip));
    return bpm::u53(q.x, q.y) < thr ? 1 : 0;
  }
  void fill_fly(bpm::PhaseArgs& a) const {
    a.fly = 1;
    a.flip_val = host_flip(a.rng);
    a.fly_shuffle = cfg.shuffle ? 1 : 0;
    if (cfg.shuffle) a.fk = bpm::make_feistel(a.rng, (uint32_t)cfg.n_chains);
  }

  int begin(const bpm_state* st, const bpm_replay* rp, cudaStream_t s, bool fly = false) {
    const int N = cfg.n_chains;
    if (fly && !rp && !serial()) return 0;      // nothing to materialise (unsharded d <= 4 only, see fly_plan)
    prof_begin(0, s);
    bool inv_on_the_fly = false;
    if (serial()) {
      bpm::identity_split_kernel<<<cdiv(N, 256), 256, 0, s>>>(perm, inv, flip, N);
    } else if (rp) {
      CU_TRY(cudaMemcpyAsync(perm, rp->shuffle_idx, sizeof(int32_t) * N, cudaMemcpyDeviceToDevice, s));
      bpm::set_flag_kernel<<<1, 1, 0, s>>>(flip, rp->flip ? 1 : 0);
      if (inv) bpm::invert_perm_kernel<<<cdiv(N, 256), 256, 0, s>>>(perm, inv, N);
    } else {
      bpm::RngCtx rng = bpm::make_rng(cfg.seed, (uint64_t)st->hist_len);
      // (Evaluating the INVERSE permutation on the fly inside the list packing instead of scattering `inv` was
      // measured and dropped: the split bucket went 31 -> 41 us at 8 GPUs and 34 -> 63 us at 4, 91 -> 122 us per
      // line-fit generation on one -- profiles/r2/r2n_*, r2o_*.  ListPos keeps the hook.)  A rank only ever looks
      // up the list positions of ITS chains, so the scatter is limited to [chain_lo, chain_hi).
      bpm::split_native_kernel<<<cdiv(N, 256), 256, 0, s>>>(perm, inv, flip, N, cfg.shuffle, cfg.flip, rng,
                                                            cfg.chain_lo, cfg.chain_hi);
    }
    if (packed()) {
      const int nblk = cdiv(cfg.chain_hi - cfg.chain_lo, bpm::kCompactBlock);
      bpm::ListPos pos;
      memset(&pos, 0, sizeof(pos));
      pos.inv = inv;
      if (inv_on_the_fly) {         // native shuffle: the key is a function of (seed, generation), evaluated here
        pos.inv = nullptr; pos.fly = 1; pos.shuffle = cfg.shuffle ? 1 : 0;
        if (cfg.shuffle) pos.fk = bpm::make_feistel(bpm::make_rng(cfg.seed, (uint64_t)st->hist_len), (uint32_t)N);
      }
      bpm::compact_count_kernel<<<nblk, bpm::kCompactThreads, 0, s>>>(pos, nA, cfg.chain_lo, cfg.chain_hi, cmp_blk);
      bpm::compact_scan_kernel<<<1, 1024, 0, s>>>(cmp_blk, nblk, cmp_blk + 2 * nblk, loc_cnt);
      bpm::compact_write_kernel<<<nblk, bpm::kCompactThreads, 0, s>>>(pos, nA, cfg.chain_lo, cfg.chain_hi,
                                                                     cmp_blk + 2 * nblk, loc_list);
    }
    prof_end(s);
    CU_TRY(cudaGetLastError());
    return 0;
  }

  template <bool REPLAY>
  int launch_propose(const bpm::PhaseArgs& a, cudaStream_t s) {
    const int lpc = pick_lpc(cfg.dim);
    const int grid = cdiv((int64_t)nA * lpc, bpm::kThreads);
    switch (lpc) {
      case 1: bpm::propose_kernel<REPLAY, 1><<<grid, bpm::kThreads, 0, s>>>(a); break;
      case 4: bpm::propose_kernel<REPLAY, 4><<<grid, bpm::kThreads, 0, s>>>(a); break;
      case 8: bpm::propose_kernel<REPLAY, 8><<<grid, bpm::kThreads, 0, s>>>(a); break;
      default: bpm::propose_kernel<REPLAY, 32><<<grid, bpm::kThreads, 0, s>>>(a); break;
    }
    CU_TRY(cudaGetLastError());
    return 0;
  }
  template <bool REPLAY>
  int launch_accept(const bpm::PhaseArgs& a, cudaStream_t s) {
    const int lpc = pick_lpc(cfg.dim);
    const int grid = cdiv((int64_t)nA * lpc, bpm::kThreads);
    switch (lpc) {
      case 1: bpm::accept_kernel<REPLAY, 1><<<grid, bpm::kThreads, 0, s>>>(a); break;
      case 4: bpm::accept_kernel<REPLAY, 4><<<grid, bpm::kThreads, 0, s>>>(a); break;
      case 8: bpm::accept_kernel<REPLAY, 8><<<grid, bpm::kThreads, 0, s>>>(a); break;
      default: bpm::accept_kernel<REPLAY, 32><<<grid, bpm::kThreads, 0, s>>>(a); break;
    }
    CU_TRY(cudaGetLastError());
    return 0;
  }

  // n_dev (optional): device-side number of valid rows (a packed phase list counted on the
  // device); n is then only the launch bound and the large-d kernel skips the tiles beyond it
  int eval_lnl(const double* P, int n, int ld, double* out, cudaStream_t s, const int32_t* n_dev = nullptr) {
    if (n <= 0) return 0;
    switch (target) {
      case BPM_TARGET_BANANA:
        if (cfg.dim != 2) return fail("banana target needs dim == 2");
        bpm::lnl_banana_kernel<<<cdiv(n, 256), 256, 0, s>>>(P, n, ld, banana, out);
        break;
      case BPM_TARGET_BIMODAL:
        if (cfg.dim != 2) return fail("bimodal target needs dim == 2");
        bpm::lnl_bimodal_kernel<<<cdiv(n, 256), 256, 0, s>>>(P, n, ld, bimodal, out);
        break;
      case BPM_TARGET_GAUSS: {
        const double* mu = tparams;
        const double* W = tparams + cfg.dim;
        if (bpm::gauss_rows_supported(cfg.dim, gauss_r)) {
          if (bpm::launch_gauss_rows(P, n, ld, cfg.dim, gauss_r, mu, W, gauss_c0, gauss_logpdf_flag,
                                     gauss_mu_zero, out, s))
            return fail("gauss rows kernel launch failed");
        }
        else if (bpm::gauss_dmma_supported(cfg.dim, gauss_r, ld)) {   // large d: FP64 tensor-pipe GEMM
          if (bpm::launch_gauss_dmma(P, n, ld, cfg.dim, gauss_r, mu, W, gauss_c0, gauss_logpdf_flag,
                                     gauss_mu_zero, out, n_dev, s))
            return fail("gauss dmma kernel launch failed");
        } else
          bpm::lnl_gauss_tiled_kernel<<<cdiv(n, 64), 256, 0, s>>>(P, n, ld, cfg.dim, gauss_r, mu, W,
                                                                  gauss_c0, gauss_logpdf_flag, out);
        break;
      }
      case BPM_TARGET_LINEFIT:
        if (cfg.dim != 3) return fail("linefit target needs dim == 3");
        bpm::lnl_linefit_kernel<<<cdiv(n, 128), 128, sizeof(double) * 3 * linefit_M, s>>>(
            P, n, ld, tparams, linefit_M, out);
        break;
      case BPM_TARGET_EXPFIT:
        if (cfg.dim != 5) return fail("expfit target needs dim == 5");
        bpm::lnl_expfit_kernel<<<cdiv(n, 128), 128, sizeof(double) * 2 * linefit_M, s>>>(
            P, n, ld, tparams, linefit_M, out);
        break;
      default:
        if (!user_fn) return fail("no likelihood: set a built-in target or a batched callback, or "
                                  "drive the split bpm_propose / bpm_accept API");
        if (user_fn(P, n, cfg.dim, ld, out, user_ptr, (bpm_stream)s))
          return fail("batched ln_like callback returned non-zero");
    }
    {
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return fail(std::string("likelihood launch: ") + cudaGetErrorString(e));
    }
    return 0;
  }

  int track_omega(const bpm_state* st, cudaStream_t s) {
    if (cov_on) {
      bpm::cross_moment_kernel<<<cfg.dim, 128, 0, s>>>(st->X, cfg.chain_lo, cfg.chain_hi, cfg.dim, cfg.ld, cov_acc,
                                                       cov_acc + cfg.dim);
      CU_TRY(cudaGetLastError());
      cov_rows += 1;
    }
    if (!omega_on) return 0;
    const int nloc = cfg.chain_hi - cfg.chain_lo;
    bpm::omega_accum_kernel<<<cdiv(nloc, 256), 256, 0, s>>>(st->lnl, omega_sum, cfg.chain_lo, cfg.chain_hi);
    CU_TRY(cudaGetLastError());
    omega_cnt += 1;
    return 0;
  }

  int end(cudaStream_t s) {
    if (cfg.algo != BPM_ALGO_DREAM) return 0;
    prof_begin(5, s);
    const int nloc = cfg.chain_hi - cfg.chain_lo;
    int nb = cdiv(nloc, 2048);
    if (nb > bpm::kCrBlocks) nb = bpm::kCrBlocks;
    // multi-rank hosts all-reduce cr_part first, then call bpm_apply_cr
    bpm::cr_update_kernel<<<nb, 256, 0, s>>>(cr_delta, cr_pick, cfg.chain_lo, cfg.chain_hi, cfg.n_cr, cr_block,
                                             cr_ticket, cr_part, sharded() ? 0 : 1, cr_dm, cr_cnt, p_cr);
    CU_TRY(cudaGetLastError());
    prof_end(s);
    return 0;
  }

  // cross-rank barrier between half-phases (demc.py:93,116,135) over peer memory
  int peer_barrier(cudaStream_t s) {
    prof_begin(6, s);
    bpm::peer_barrier_kernel<<<1, 32, 0, s>>>(sync, ++sync_epoch);
    CU_TRY(cudaGetLastError());
    prof_end(s);
    return 0;
  }
  // ... and the one that closes a DREAM generation, carrying the CR all-reduce
  int peer_cr_exchange(cudaStream_t s) {
    prof_begin(6, s);
    bpm::peer_cr_exchange_kernel<<<1, 64, 0, s>>>(sync, ++sync_epoch, (int)(sync_cr_seq++ & 1ull), cr_part, cfg.n_cr,
                                                  cr_dm, cr_cnt, p_cr);
    CU_TRY(cudaGetLastError());
    prof_end(s);
    return 0;
  }

  template <bool REPLAY>
  int phase(const bpm_state* st, int64_t k_gen, int ph, const bpm_replay* rp, const bpm_trace_out* tr,
            cudaStream_t s, bool lazy, bool fly) {
    bpm::PhaseArgs a = make_args(st, k_gen, ph, rp, tr, lazy, fly);
    if (a.fly && !sharded()) { a.loc_list = nullptr; a.loc_cnt = nullptr; }   // nothing was packed
    if (fused_ok && !serial()) {
      int done = 0;
      prof_begin(4, s);
      if (bpm::try_fused_phase<REPLAY>(*this_target(), a, s, fused_ok, &done))
        return fail(std::string("fused half-phase launch failed: ") + cudaGetErrorString(cudaGetLastError()));
      if (done) { prof_end(s); return 0; }
      if (prof_on) { ev_pool.push_back(recs.back().a); ev_pool.push_back(recs.back().b); recs.pop_back(); }
    }
    if (lazy || a.fly) return fail("internal: lazy / fly mode planned but the fused kernel did not launch");
    prof_begin(1, s);
    BPM_TRY(launch_propose<REPLAY>(a, s));
    prof_end(s);
    prof_begin(2, s);
    {
      // packed lists hold at most this rank's chains; their exact number lives on the device
      int n_rows = nA;
      const int32_t* n_dev = nullptr;
      if (packed()) {
        const int n_local = cfg.chain_hi - cfg.chain_lo;
        n_rows = n_local < nA ? n_local : nA;
        bpm::phase_count_kernel<<<1, 1, 0, s>>>(loc_cnt, flip, ph, serial() ? 1 : 0, phase_cnt);
        n_dev = phase_cnt;
      }
      BPM_TRY(eval_lnl(prop, n_rows, cfg.ld, lnl_prop, s, n_dev));
    }
    prof_end(s);
    prof_begin(3, s);
    BPM_TRY(launch_accept<REPLAY>(a, s));
    prof_end(s);
    return 0;
  }

  // View of the target parameters the fused kernels need.
  bpm::TargetView tv;
  const bpm::TargetView* this_target() {
    tv.target = target;
    tv.banana = banana;
    tv.bimodal = bimodal;
    tv.mu = tparams;
    tv.W = tparams ? tparams + cfg.dim : nullptr;
    tv.Wf = target == BPM_TARGET_GAUSS ? wfrag : nullptr;
    tv.r = gauss_r;
    tv.c0 = gauss_c0;
    tv.log_of_pdf = gauss_logpdf_flag;
    tv.mu_is_zero = gauss_mu_zero;
    tv.linefit = tparams;
    tv.linefit_M = linefit_M;
    return &tv;
  }

  // d <= 4, native RNG, unsharded: every generation of the call in ONE persistent cooperative launch
  // (kernels_fused.cuh: small_generations_kernel).  Returns done = 0 when the configuration is not covered.
  int coop_blocks = 0;           // co-resident 256-thread blocks of small_generations_kernel (0 = not asked yet)
  int try_small_generations(bpm_state* st, int64_t k_gen0, int n_gen, cudaStream_t s, int* done) {
    *done = 0;
    if (fused_ok != 6 || serial() || sharded() || cfg.dim > 4 || cov_on || n_peers > 0) return 0;   // experimental
    if (!(target == BPM_TARGET_BANANA || target == BPM_TARGET_BIMODAL || target == BPM_TARGET_LINEFIT)) return 0;
    if (st->pending) BPM_TRY(flush(st, s));
    const size_t sm = target == BPM_TARGET_LINEFIT ? sizeof(double) * 3 * linefit_M : 0;
    void* fn = target == BPM_TARGET_BANANA ? (void*)bpm::small_generations_kernel<BPM_TARGET_BANANA>
             : target == BPM_TARGET_BIMODAL ? (void*)bpm::small_generations_kernel<BPM_TARGET_BIMODAL>
                                            : (void*)bpm::small_generations_kernel<BPM_TARGET_LINEFIT>;
    if (coop_blocks == 0) {
      int coop = 0, per_sm = 0, sms = 0;
      CU_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, cfg.device));
      CU_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg.device));
      CU_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, 256, sm));
      coop_blocks = coop ? sms * (per_sm < 4 ? per_sm : 4) : -1;
    }
    if (coop_blocks <= 0) return 0;
    // one resident thread per chain or not at all: a thread that has to walk several chains serialises their
    // (latency-bound) likelihoods, and the per-phase launches win (measured: 10^6 bimodal chains 370 vs 197 us)
    if ((int64_t)cfg.n_chains > (int64_t)coop_blocks * 256) return 0;
    bpm::PhaseArgs a = make_args(st, k_gen0, 0, nullptr, nullptr);
    a.loc_list = nullptr; a.loc_cnt = nullptr;        // chains are walked in order; no packed lists
    bpm::SmallGens q;
    memset(&q, 0, sizeof(q));
    q.k_gen0 = k_gen0; q.n_gen = n_gen; q.burnin_gen = cfg.burnin_gen; q.n_cr_gen = cfg.n_cr_gen;
    q.jump_mod = cfg.algo == BPM_ALGO_DREAM ? 5 : 10;
    q.shuffle = cfg.shuffle; q.flip_p = cfg.flip; q.seed = cfg.seed;
    q.hist0 = st->history; q.omega_sum = omega_on ? omega_sum : nullptr;
    q.cr_block = cr_block; q.cr_part = cr_part; q.cr_dm = cr_dm; q.cr_cnt = cr_cnt; q.p_cr = p_cr;
    int nb = cdiv(cfg.n_chains, 2048);
    if (nb > bpm::kCrBlocks) nb = bpm::kCrBlocks;
    int grid = cdiv(cfg.n_chains, 256);
    if (grid > coop_blocks) grid = coop_blocks;
    if (grid < nb) nb = grid;                           // (the partial sums stay in block order either way)
    q.cr_blocks = nb;
    const bpm::TargetView* tvp = this_target();
    void* args[] = {(void*)&a, (void*)tvp, (void*)&q};
    prof_begin(4, s);
    CU_TRY(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(256), args, sm, s));
    prof_end(s);
    st->hist_len += n_gen;
    if (st->mom_len > 0) st->mom_len += n_gen;
    if (omega_on) omega_cnt += n_gen;
    st->pending = 0;
    *done = 1;
    return 0;
  }

  template <bool REPLAY>
  int generation(bpm_state* st, int64_t k_gen, const bpm_replay* rp, const bpm_trace_out* tr,
                 cudaStream_t s, bool split_ready = false) {
    // lazy protocol: the v3 kernel folds the row the previous generation left pending and leaves its own
    // pending; every other path (and a replay step, whose exact np.std walks the stored history) needs
    // the row materialised first
    const bool lazy = lazy_plan();
    const bool fly = !REPLAY && fly_plan();
    if (st->pending && (!lazy || REPLAY)) BPM_TRY(flush(st, s));
    if (!split_ready) BPM_TRY(begin(st, rp, s, fly));
    BPM_TRY(phase<REPLAY>(st, k_gen, 0, rp, tr, s, lazy, fly));
    if (!serial()) {
      if (sync_on) BPM_TRY(peer_barrier(s));          // every rank's phase-a rows are in every replica
      BPM_TRY(phase<REPLAY>(st, k_gen, 1, rp, tr, s, lazy, fly));
    }
    BPM_TRY(end(s));
    if (sync_on) {
      if (cfg.algo == BPM_ALGO_DREAM) BPM_TRY(peer_cr_exchange(s));
      else BPM_TRY(peer_barrier(s));
    }
    BPM_TRY(track_omega(st, s));
    st->hist_len += 1;
    if (st->mom_len > 0) st->mom_len += 1;
    st->pending = lazy ? 1 : 0;
    return 0;
  }
};

// ===================================================================================
extern "C" {

const char* bpm_last_error(void) { return g_err.c_str(); }
int bpm_version(void) { return 100; }

int bpm_create(const bpm_config* cfg, bpm_handle* out) {
  if (!cfg || !out) return fail("bpm_create: null argument");
  if (cfg->n_chains < 4) return fail("n_chains >= 4 required (samplers.py:249)");
  if (cfg->dim < 1 || cfg->ld < cfg->dim) return fail("bad dim / ld");
  if (cfg->dim > 4 * bpm::kMaxBlocksPerLane * 32) return fail("dim > 1024 not supported yet");
  if (cfg->algo != BPM_ALGO_DEMC && cfg->algo != BPM_ALGO_DREAM && cfg->algo != BPM_ALGO_DEMC_SERIAL)
    return fail("bad algo");
  if (cfg->algo == BPM_ALGO_DREAM && (cfg->del_pairs < 1 || cfg->del_pairs > BPM_MAX_PAIRS))
    return fail("del_pairs must be in [1, 8]");
  if (cfg->n_cr < 1 || cfg->n_cr > BPM_MAX_CR) return fail("n_cr must be in [1, 16]");
  if (cfg->chain_lo < 0 || cfg->chain_hi > cfg->n_chains || cfg->chain_lo >= cfg->chain_hi)
    return fail("bad chain shard");
  if (cfg->n_chains / 2 < 2) return fail("pool too small");
  bpm_engine* e = new (std::nothrow) bpm_engine();
  if (!e) return fail("out of host memory");
  e->cfg = *cfg;
  if (e->init()) {
    delete e;
    return 1;
  }
  *out = e;
  return 0;
}

int bpm_destroy(bpm_handle h) {
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  delete h;
  return 0;
}

int bpm_set_run_params(bpm_handle h, double flip, int32_t shuffle, double epsilon, double u_epsilon,
                       double gamma) {
  if (!h) return fail("null handle");
  h->cfg.flip = flip < 0.0 ? 0.0 : (flip > 1.0 ? 1.0 : flip);
  h->cfg.shuffle = shuffle;
  h->cfg.epsilon = epsilon;
  h->cfg.u_epsilon = u_epsilon;
  h->cfg.gamma = gamma;
  return 0;
}

int bpm_set_fused(bpm_handle h, int32_t on) {
  if (!h) return fail("null handle");
  h->fused_ok = on;
  return 0;
}

static void fill_mvn2(bpm::Mvn2& g, const double* p) {
  g.mu0 = p[0]; g.mu1 = p[1]; g.U00 = p[2]; g.U01 = p[3]; g.U10 = p[4]; g.U11 = p[5]; g.c0 = p[6];
}

int bpm_set_target(bpm_handle h, int32_t target, const double* params, int64_t n) {
  if (!h) return fail("null handle");
  CU_TRY(cudaSetDevice(h->cfg.device));
  const int d = h->cfg.dim;
  switch (target) {
    case BPM_TARGET_EXTERNAL: break;
    case BPM_TARGET_BANANA:
      if (n != 10) return fail("banana: 10 parameters expected");
      h->banana.log_of_pdf = params[0]; h->banana.a = params[1]; h->banana.b = params[2];
      fill_mvn2(h->banana.g, params + 3);
      break;
    case BPM_TARGET_BIMODAL:
      if (n != 17) return fail("bimodal: 17 parameters expected");
      h->bimodal.log_of_pdf = params[0]; h->bimodal.w1 = params[1]; h->bimodal.w2 = params[2];
      fill_mvn2(h->bimodal.g1, params + 3);
      fill_mvn2(h->bimodal.g2, params + 10);
      h->bimodal.lw1 = log(h->bimodal.w1); h->bimodal.lw2 = log(h->bimodal.w2);
      break;
    case BPM_TARGET_GAUSS: {
      // [log_of_pdf, c0, r, mu[d], W[d][r]]
      if (n < 3) return fail("gauss: header missing");
      const int r = (int)params[2];
      if (r < 1 || n != 3 + d + (int64_t)d * r) return fail("gauss: parameter count mismatch");
      h->gauss_logpdf_flag = params[0] != 0.0; h->gauss_c0 = params[1]; h->gauss_r = r;
      h->gauss_mu_zero = 1;
      for (int i = 0; i < d; ++i) if (params[3 + i] != 0.0) h->gauss_mu_zero = 0;
      cudaFree(h->tparams); h->tparams = nullptr;
      CU_TRY(cudaMalloc(&h->tparams, sizeof(double) * (n - 3)));
      CU_TRY(cudaMemcpy(h->tparams, params + 3, sizeof(double) * (n - 3), cudaMemcpyHostToDevice));
      // W once more in DMMA fragment order, for the fused kernels' single bulk copy (kernels_fused.cuh)
      cudaFree(h->wfrag); h->wfrag = nullptr;
      if ((d % 4) == 0 && bpm::gauss_rows_supported(d, r)) {
        const size_t nf = (size_t)(d >> 2) * bpm::dmma_ntiles(r) * 32;
        CU_TRY(cudaMalloc(&h->wfrag, sizeof(double) * nf));
        bpm::w_fragment_kernel<<<cdiv((int64_t)nf, 256), 256>>>(h->wfrag, h->tparams + d, d, r);
        CU_TRY(cudaGetLastError());
        CU_TRY(cudaDeviceSynchronize());
      }
      break;
    }
    case BPM_TARGET_LINEFIT: {
      // [M, x[M], y[M], yerr[M]]
      if (n < 1) return fail("linefit: header missing");
      const int M = (int)params[0];
      if (M < 1 || n != 1 + 3 * (int64_t)M || M > 2000) return fail("linefit: parameter count mismatch");
      h->linefit_M = M;
      cudaFree(h->tparams); h->tparams = nullptr;
      CU_TRY(cudaMalloc(&h->tparams, sizeof(double) * 3 * M));
      CU_TRY(cudaMemcpy(h->tparams, params + 1, sizeof(double) * 3 * M, cudaMemcpyHostToDevice));
      break;
    }
    case BPM_TARGET_EXPFIT: {
      // [M, t[M], y[M]]
      if (n < 1) return fail("expfit: header missing");
      const int M = (int)params[0];
      if (M < 1 || n != 1 + 2 * (int64_t)M || M > 3000) return fail("expfit: parameter count mismatch");
      h->linefit_M = M;
      cudaFree(h->tparams); h->tparams = nullptr;
      CU_TRY(cudaMalloc(&h->tparams, sizeof(double) * 2 * M));
      CU_TRY(cudaMemcpy(h->tparams, params + 1, sizeof(double) * 2 * M, cudaMemcpyHostToDevice));
      break;
    }
    default: return fail("unknown target id");
  }
  h->target = target;
  return 0;
}

int bpm_set_batched_lnl(bpm_handle h, bpm_lnl_fn fn, void* user) {
  if (!h) return fail("null handle");
  h->user_fn = fn;
  h->user_ptr = user;
  h->target = BPM_TARGET_EXTERNAL;
  return 0;
}

int bpm_eval_lnl(bpm_handle h, const double* X, int32_t n, double* lnl, bpm_stream stream) {
  if (!h) return fail("null handle");
  CU_TRY(cudaSetDevice(h->cfg.device));
  return h->eval_lnl(X, n, h->cfg.ld, lnl, (cudaStream_t)stream);
}

int bpm_set_cr_state(bpm_handle h, const double* p_cr, const double* dm, const double* cnt) {
  if (!h) return fail("null handle");
  CU_TRY(cudaSetDevice(h->cfg.device));
  const size_t b = sizeof(double) * h->cfg.n_cr;
  if (p_cr) CU_TRY(cudaMemcpy(h->p_cr, p_cr, b, cudaMemcpyHostToDevice));
  if (dm) CU_TRY(cudaMemcpy(h->cr_dm, dm, b, cudaMemcpyHostToDevice));
  if (cnt) CU_TRY(cudaMemcpy(h->cr_cnt, cnt, b, cudaMemcpyHostToDevice));
  return 0;
}

int bpm_get_cr_state(bpm_handle h, double* p_cr, double* dm, double* cnt) {
  if (!h) return fail("null handle");
  CU_TRY(cudaSetDevice(h->cfg.device));
  const size_t b = sizeof(double) * h->cfg.n_cr;
  if (p_cr) CU_TRY(cudaMemcpy(p_cr, h->p_cr, b, cudaMemcpyDeviceToHost));
  if (dm) CU_TRY(cudaMemcpy(dm, h->cr_dm, b, cudaMemcpyDeviceToHost));
  if (cnt) CU_TRY(cudaMemcpy(cnt, h->cr_cnt, b, cudaMemcpyDeviceToHost));
  return 0;
}

// Multi-rank CR adaptation: device pointer to this rank's per-generation partial sums
// (2 * n_cr doubles: jump statistics, then counts) for the host's all-reduce, and the
// p_cr update from the reduced values.
int bpm_cr_partials(bpm_handle h, double** dev_ptr) {
  if (!h || !dev_ptr) return fail("null argument");
  *dev_ptr = h->cr_part;
  return 0;
}
int bpm_apply_cr(bpm_handle h, bpm_stream stream) {
  if (!h) return fail("null handle");
  CU_TRY(cudaSetDevice(h->cfg.device));
  bpm::cr_apply_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->cr_part, h->cfg.n_cr, h->cr_dm,
                                                          h->cr_cnt, h->p_cr);
  CU_TRY(cudaGetLastError());
  return 0;
}

int bpm_get_counters(bpm_handle h, uint64_t* acc, uint64_t* rej, int32_t* nan_alpha) {
  if (!h) return fail("null handle");
  CU_TRY(cudaSetDevice(h->cfg.device));
  unsigned long long c[2];
  int32_t f;
  CU_TRY(cudaMemcpy(c, h->counters, sizeof(c), cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(&f, h->nan_flag, sizeof(f), cudaMemcpyDeviceToHost));
  if (acc) *acc = c[0];
  if (rej) *rej = c[1];
  if (nan_alpha) *nan_alpha = f;
  return 0;
}

int bpm_reset_counters(bpm_handle h) {
  if (!h) return fail("null handle");
  CU_TRY(cudaSetDevice(h->cfg.device));
  CU_TRY(cudaMemset(h->counters, 0, sizeof(unsigned long long) * 2));
  CU_TRY(cudaMemset(h->nan_flag, 0, sizeof(int32_t)));
  return 0;
}

int bpm_step_generations(bpm_handle h, bpm_state* st, int64_t k_gen0, int32_t n_gen,
                         bpm_stream stream) {
  if (!h || !st) return fail("null argument");
  if (!st->X || !st->lnl) return fail("state needs X and lnl");
  if (h->target == BPM_TARGET_EXTERNAL && !h->user_fn)
    return fail("bpm_step_generations needs a built-in target or a batched callback");
  if (h->sharded() && !h->sync_on)
    return fail("bpm_step_generations on a sharded handle needs bpm_set_sync (peer-memory barrier); without it "
                "the host drives the half-phases: bpm_begin_generation / bpm_phase / exchange / bpm_end_generation");
  if (h->sharded() && h->serial())
    return fail("serial DE-MC steps every chain against the frozen population: sharded handles must use the "
                "split API with an all-gather after the sweep");
  CU_TRY(cudaSetDevice(h->cfg.device));
  if (n_gen > 0) {
    int done = 0;
    BPM_TRY(h->try_small_generations(st, k_gen0, n_gen, (cudaStream_t)stream, &done));
    if (done) return 0;
  }
  static const bool no_side = [] {
    const char* e = getenv("BIPYMC_B200_NO_SIDE");
    return e && e[0] == '1';
  }();
  cudaStream_t s = (cudaStream_t)stream;
  const bool side_ok = !no_side && n_gen >= 2 && !h->serial() && !h->fly_plan();
  if (side_ok) {
    BPM_TRY(h->side_setup());
    CU_TRY(cudaEventRecord(h->ev_start, s));
  }
  bool have_next = false;        // the split of the generation about to run was prepared on the side stream
  for (int g = 0; g < n_gen; ++g) {
    if (have_next) {
      h->swap_split();                                   // current <- the buffers the side stream filled
      CU_TRY(cudaStreamWaitEvent(s, h->ev_split, 0));
    }
    const bool ready = have_next;
    have_next = side_ok && g + 1 < n_gen;
    if (have_next) {
      // generation g+1's shuffle into the other buffer set, free once generation g-1 (its last user) is done
      CU_TRY(cudaStreamWaitEvent(h->side, g >= 1 ? h->ev_done : h->ev_start, 0));
      bpm_state nxt = *st;
      nxt.hist_len = st->hist_len + 1;
      h->swap_split();
      const int rc = h->begin(&nxt, nullptr, h->side, false);
      h->swap_split();
      if (rc) return rc;
      CU_TRY(cudaEventRecord(h->ev_split, h->side));
    }
    BPM_TRY(h->generation<false>(st, k_gen0 + g, nullptr, nullptr, s, ready));
    if (have_next) CU_TRY(cudaEventRecord(h->ev_done, s));
  }
  return 0;
}

int bpm_step_generation_replay(bpm_handle h, bpm_state* st, const bpm_replay* rp, int64_t k_gen,
                               const bpm_trace_out* trace, bpm_stream stream) {
  if (!h || !st || !rp) return fail("null argument");
  if (!st->X || !st->lnl) return fail("state needs X and lnl");
  if (h->target == BPM_TARGET_EXTERNAL && !h->user_fn)
    return fail("replay step needs a built-in target or a batched callback");
  CU_TRY(cudaSetDevice(h->cfg.device));
  return h->generation<true>(st, k_gen, rp, trace, (cudaStream_t)stream);
}

int bpm_dump_draws(bpm_handle h, const bpm_state* st, int64_t k_gen, bpm_replay* out, int32_t* flip,
                   bpm_stream stream) {
  if (!h || !st || !out) return fail("null argument");
  CU_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  BPM_TRY(h->begin(st, nullptr, s));
  bpm::PhaseArgs a = h->make_args(st, k_gen, 0, nullptr, nullptr);
  bpm::dump_draws_kernel<<<cdiv(h->cfg.n_chains, 128), 128, 0, s>>>(a, *out);
  CU_TRY(cudaGetLastError());
  if (flip) {
    CU_TRY(cudaStreamSynchronize(s));
    CU_TRY(cudaMemcpy(flip, h->flip, sizeof(int32_t), cudaMemcpyDeviceToHost));
    out->flip = *flip;
  }
  return 0;
}

int bpm_begin_generation(bpm_handle h, bpm_state* st, int64_t k_gen, const bpm_replay* rp,
                         bpm_stream stream) {
  if (!h || !st) return fail("null argument");
  CU_TRY(cudaSetDevice(h->cfg.device));
  // bpm_phase may run the lazy-protocol kernel; bpm_propose / bpm_accept are eager
  h->cur_lazy = h->lazy_plan() && !(h->target == BPM_TARGET_EXTERNAL);
  h->cur_phases_run = 0;
  if (st->pending && (!h->cur_lazy || rp)) BPM_TRY(h->flush(st, (cudaStream_t)stream));
  BPM_TRY(h->begin(st, rp, (cudaStream_t)stream, false));
  h->cur = h->make_args(st, k_gen, 0, rp, nullptr);
  h->cur_replay = rp != nullptr;
  h->cur_k_gen = k_gen;
  h->in_generation = true;
  return 0;
}

int bpm_propose(bpm_handle h, bpm_state* st, int32_t phase, double** prop, int32_t* n_phase,
                bpm_stream stream) {
  if (!h || !st) return fail("null argument");
  if (!h->in_generation) return fail("bpm_propose outside bpm_begin_generation / bpm_end_generation");
  CU_TRY(cudaSetDevice(h->cfg.device));
  if (h->cur_lazy) {     // the caller drives a built-in target through the split API after all: go eager
    if (h->cur_phases_run) return fail("bpm_propose after bpm_phase in the same generation");
    BPM_TRY(h->flush(st, (cudaStream_t)stream));
    h->cur_lazy = false;
    h->cur = h->make_args(st, h->cur_k_gen, 0, h->cur_replay ? &h->cur.rp : nullptr, nullptr);
  }
  h->cur.phase = phase;
  h->cur.X = st->X; h->cur.lnl = st->lnl;
  if (h->cur_replay) BPM_TRY(h->launch_propose<true>(h->cur, (cudaStream_t)stream));
  else BPM_TRY(h->launch_propose<false>(h->cur, (cudaStream_t)stream));
  if (prop) *prop = h->prop;
  if (n_phase) {
    // the phase size depends on the flip when N is odd: read the flag back
    int32_t f = 0;
    CU_TRY(cudaStreamSynchronize((cudaStream_t)stream));
    CU_TRY(cudaMemcpy(&f, h->flip, sizeof(f), cudaMemcpyDeviceToHost));
    const bool first = ((phase ^ (f != 0)) == 0);
    *n_phase = first ? h->nA : h->cfg.n_chains - h->nA;
    if (h->packed()) {    // only this rank's chains of the half were proposed, densely packed
      int32_t cnt[2];
      CU_TRY(cudaMemcpy(cnt, h->loc_cnt, sizeof(cnt), cudaMemcpyDeviceToHost));
      *n_phase = cnt[first ? 0 : 1];
    }
  }
  return 0;
}

int bpm_accept(bpm_handle h, bpm_state* st, int32_t phase, const double* lnl_prop,
               const bpm_trace_out* trace, bpm_stream stream) {
  if (!h || !st || !lnl_prop) return fail("null argument");
  if (!h->in_generation) return fail("bpm_accept outside a generation");
  CU_TRY(cudaSetDevice(h->cfg.device));
  bpm::PhaseArgs a = h->cur;
  a.phase = phase;
  a.lnl_prop = const_cast<double*>(lnl_prop);
  if (trace) a.tr = *trace;
  if (h->cur_replay) BPM_TRY(h->launch_accept<true>(a, (cudaStream_t)stream));
  else BPM_TRY(h->launch_accept<false>(a, (cudaStream_t)stream));
  return 0;
}

int bpm_phase(bpm_handle h, bpm_state* st, int32_t phase, bpm_stream stream) {
  if (!h || !st) return fail("null argument");
  if (!h->in_generation) return fail("bpm_phase outside a generation");
  if (h->target == BPM_TARGET_EXTERNAL && !h->user_fn)
    return fail("bpm_phase needs a built-in target or a batched callback");
  CU_TRY(cudaSetDevice(h->cfg.device));
  const bpm_replay* rp = h->cur_replay ? &h->cur.rp : nullptr;
  const int64_t k_gen = h->cur_k_gen;
  h->cur_phases_run += 1;
  if (h->cur_replay) return h->phase<true>(st, k_gen, phase, rp, nullptr, (cudaStream_t)stream, h->cur_lazy, false);
  return h->phase<false>(st, k_gen, phase, nullptr, nullptr, (cudaStream_t)stream, h->cur_lazy, false);
}

int bpm_end_generation(bpm_handle h, bpm_state* st, bpm_stream stream) {
  if (!h || !st) return fail("null argument");
  if (!h->in_generation) return fail("bpm_end_generation without bpm_begin_generation");
  CU_TRY(cudaSetDevice(h->cfg.device));
  BPM_TRY(h->end((cudaStream_t)stream));
  BPM_TRY(h->track_omega(st, (cudaStream_t)stream));
  st->hist_len += 1;
  if (st->mom_len > 0) st->mom_len += 1;
  st->pending = h->cur_lazy ? 1 : 0;
  h->in_generation = false;
  return 0;
}

int bpm_generations_host(bpm_handle h, double* X_host, double* lnl_host, int64_t k_gen0,
                         int64_t g_abs0, int32_t n_gen) {
  if (!h || !X_host || !lnl_host) return fail("null argument");
  if (h->cfg.chain_lo != 0 || h->cfg.chain_hi != h->cfg.n_chains)
    return fail("host entry is single-rank");
  CU_TRY(cudaSetDevice(h->cfg.device));
  const int N = h->cfg.n_chains;
  const size_t nx = sizeof(double) * (size_t)N * h->cfg.ld;
  const size_t nl = sizeof(double) * (size_t)N;
  if (!h->hX) {
    CU_TRY(cudaMalloc(&h->hX, nx));
    CU_TRY(cudaMalloc(&h->hL, nl));
    CU_TRY(cudaMalloc(&h->h_accept, sizeof(int32_t) * N));
    CU_TRY(cudaMalloc(&h->h_changed, sizeof(int32_t) * N));
    CU_TRY(cudaMalloc(&h->h_nrows, sizeof(unsigned long long)));
    if (h->cfg.algo == BPM_ALGO_DREAM) {
      CU_TRY(cudaMalloc(&h->hMean, nx));
      CU_TRY(cudaMalloc(&h->hM2, nx));
    }
    h->h_mom_len = 0;
  }
  // pinned (mapped) host buffers can be written by the device directly: only rows that moved go back
  double *X_map = nullptr, *L_map = nullptr;
  {
    cudaPointerAttributes ax, al;
    if (cudaPointerGetAttributes(&ax, X_host) == cudaSuccess && cudaPointerGetAttributes(&al, lnl_host) == cudaSuccess &&
        ax.type == cudaMemoryTypeHost && al.type == cudaMemoryTypeHost && ax.devicePointer && al.devicePointer) {
      X_map = (double*)ax.devicePointer;
      L_map = (double*)al.devicePointer;
    } else {
      cudaGetLastError();
    }
  }
  cudaStream_t s = 0;
  // Default (BIPYMC_B200_HOST_PEER=0 selects the older changed-rows pass below): the mapped host population is
  // handed to the phase kernels as one more peer replica, so every accepted row is stored to the host from
  // inside the kernels (overlapped with compute) and no changed-rows pass follows the generation; the
  // cached likelihoods (8 N bytes) come back with one plain copy.
  static const bool host_peer_opt = [] {
    const char* e = getenv("BIPYMC_B200_HOST_PEER");     // default on; "0" = the changed-rows pass after the generation
    return !(e && e[0] == '0');
  }();
  if (X_map && host_peer_opt && h->n_peers < BPM_MAX_PEERS) {
    unsigned long long acc0 = 0, acc1 = 0;
    CU_TRY(cudaMemcpyAsync(&acc0, h->counters, sizeof(acc0), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(h->hX, X_host, nx, cudaMemcpyHostToDevice, s));
    CU_TRY(cudaMemcpyAsync(h->hL, lnl_host, nl, cudaMemcpyHostToDevice, s));
    bpm_state st;
    BPM_TRY(h->host_entry_state(&st, g_abs0, nx, s));
    h->peers[h->n_peers++] = X_map;
    int rc = 0;
    for (int g = 0; g < n_gen && rc == 0; ++g) rc = h->generation<false>(&st, k_gen0 + g, nullptr, nullptr, s);
    h->peers[--h->n_peers] = nullptr;
    if (rc) return rc;
    h->h_mom_len = st.mom_len; h->h_pending = st.pending;
    CU_TRY(cudaMemcpyAsync(lnl_host, h->hL, nl, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(&acc1, h->counters, sizeof(acc1), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    h->last_d2h_bytes = (acc1 - acc0) * (sizeof(double) * h->cfg.ld) + nl + 2 * sizeof(acc0);
    return 0;
  }
  CU_TRY(cudaMemcpyAsync(h->hX, X_host, nx, cudaMemcpyHostToDevice, s));
  CU_TRY(cudaMemcpyAsync(h->hL, lnl_host, nl, cudaMemcpyHostToDevice, s));
  bpm_state st;
  BPM_TRY(h->host_entry_state(&st, g_abs0, nx, s));
  bpm_trace_out tr;
  memset(&tr, 0, sizeof(tr));
  tr.accept = h->h_accept;
  if (X_map) {
    CU_TRY(cudaMemsetAsync(h->h_changed, 0, sizeof(int32_t) * N, s));
    CU_TRY(cudaMemsetAsync(h->h_nrows, 0, sizeof(unsigned long long), s));
  }
  for (int g = 0; g < n_gen; ++g) {
    BPM_TRY(h->generation<false>(&st, k_gen0 + g, nullptr, X_map ? &tr : nullptr, s));
    if (X_map) bpm::or_flags_kernel<<<cdiv(N, 256), 256, 0, s>>>(h->h_changed, h->h_accept, N);
  }
  h->h_mom_len = st.mom_len; h->h_pending = st.pending;
  if (X_map) {
    bpm::scatter_changed_rows_kernel<<<cdiv((int64_t)N * 32, 256), 256, 0, s>>>(h->hX, h->hL, h->h_changed, X_map,
                                                                               L_map, N, h->cfg.ld, h->h_nrows);
    CU_TRY(cudaGetLastError());
    unsigned long long rows = 0;
    CU_TRY(cudaMemcpyAsync(&rows, h->h_nrows, sizeof(rows), cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    h->last_d2h_bytes = rows * (sizeof(double) * (h->cfg.ld + 1)) + sizeof(rows);
  } else {
    CU_TRY(cudaMemcpyAsync(X_host, h->hX, nx, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaMemcpyAsync(lnl_host, h->hL, nl, cudaMemcpyDeviceToHost, s));
    CU_TRY(cudaStreamSynchronize(s));
    h->last_d2h_bytes = nx + nl;
  }
  return 0;
}

int bpm_generations_host_sharded(bpm_handle h, bpm_state* st, double* X_host, double* lnl_host, int64_t k_gen0,
                                 int32_t n_gen, bpm_stream stream) {
  if (!h || !st || !st->X || !st->lnl || !X_host || !lnl_host) return fail("null argument");
  if (!h->sharded()) return fail("bpm_generations_host_sharded: the handle owns every chain; use bpm_generations_host");
  if (!h->sync_on || h->n_peers != h->sync.world - 1)
    return fail("bpm_generations_host_sharded needs the peer replicas (bpm_set_peers) and the peer barrier (bpm_set_sync)");
  if (h->serial()) return fail("serial DE-MC is not offered on the sharded host entry");
  if (h->n_peers >= BPM_MAX_PEERS) return fail("no peer slot left for the host array");
  if (h->target == BPM_TARGET_EXTERNAL && !h->user_fn) return fail("needs a built-in target or a batched callback");
  CU_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int lo = h->cfg.chain_lo, nloc = h->cfg.chain_hi - lo, ld = h->cfg.ld;
  if (((size_t)ld * sizeof(double)) % 16 != 0) return fail("rows must be multiples of 16 bytes");
  double *X_map = nullptr;
  {
    cudaPointerAttributes ax;
    if (cudaPointerGetAttributes(&ax, X_host) != cudaSuccess || ax.type != cudaMemoryTypeHost || !ax.devicePointer) {
      cudaGetLastError();
      return fail("bpm_generations_host_sharded: X_host must be pinned (device-mapped) host memory");
    }
    X_map = (double*)ax.devicePointer;
  }
  const size_t off = (size_t)lo * ld;
  bpm::ShardInArgs sa;
  memset(&sa, 0, sizeof(sa));
  sa.src = reinterpret_cast<const double2*>(X_map);
  sa.dst[0] = reinterpret_cast<double2*>(st->X + off);
  for (int p = 0; p < h->n_peers; ++p) sa.dst[1 + p] = reinterpret_cast<double2*>(h->peers[p] + off);
  sa.n_dst = 1 + h->n_peers;
  sa.n2 = (int64_t)nloc * ld / 2;
  unsigned long long acc0 = 0, acc1 = 0;
  CU_TRY(cudaMemcpyAsync(&acc0, h->counters, sizeof(acc0), cudaMemcpyDeviceToHost, s));
  static const bool shard_dma_opt = [] {
    const char* e = getenv("BIPYMC_B200_SHARD_DMA");
    return e && e[0] == '1';
  }();
  h->prof_begin(7, s);
  if (!shard_dma_opt) {
    // Default, one kernel: zero-copy PCIe reads, every 16-byte piece stored into every replica (SM-issued NVLink
    // stores riding under the PCIe read)
    bpm::shard_in_kernel<<<296, 512, 0, s>>>(sa);
    CU_TRY(cudaGetLastError());
  } else {
    // BIPYMC_B200_SHARD_DMA=1: copy engines.  The shard comes in as kChunks DMA copies on the caller's stream; as
    // soon as a chunk has landed the side stream forwards it to every peer replica with peer copies (NVLink copy
    // engines), under the next chunk's host copy.  Measured against the one-kernel form (profiles/r2/r2v_*, r2w_*):
    // 2.04 vs 2.09 ms per end-to-end step at 2 GPUs, 3.87 vs 3.69 at 4, 5.06 vs 4.65 at 8 -- the step is bound by
    // the ranks' concurrent reads of host memory, and 7 x 4 peer copies per rank only add submission work.
    constexpr int kChunks = 4;
    BPM_TRY(h->side_setup());
    if (!h->ev_chunk[0])
      for (int k = 0; k <= kChunks; ++k) CU_TRY(cudaEventCreateWithFlags(&h->ev_chunk[k], cudaEventDisableTiming));
    const size_t total = (size_t)nloc * ld, per = ((total / kChunks) + 1) & ~(size_t)1;
    CU_TRY(cudaEventRecord(h->ev_chunk[kChunks], s));
    CU_TRY(cudaStreamWaitEvent(h->side, h->ev_chunk[kChunks], 0));     // the side stream starts after the caller's prior work
    for (int k = 0; k < kChunks; ++k) {
      const size_t o0 = (size_t)k * per;
      if (o0 >= total) break;
      const size_t n = (o0 + per <= total ? per : total - o0) * sizeof(double);
      CU_TRY(cudaMemcpyAsync(st->X + off + o0, X_host + o0, n, cudaMemcpyHostToDevice, s));
      CU_TRY(cudaEventRecord(h->ev_chunk[k], s));
      CU_TRY(cudaStreamWaitEvent(h->side, h->ev_chunk[k], 0));
      for (int p = 0; p < h->n_peers; ++p)
        CU_TRY(cudaMemcpyAsync(h->peers[p] + off + o0, st->X + off + o0, n, cudaMemcpyDefault, h->side));
    }
    CU_TRY(cudaEventRecord(h->ev_chunk[kChunks], h->side));
    CU_TRY(cudaStreamWaitEvent(s, h->ev_chunk[kChunks], 0));           // every peer copy of this rank is done
  }
  h->prof_end(s);
  CU_TRY(cudaMemcpyAsync(st->lnl + lo, lnl_host, sizeof(double) * nloc, cudaMemcpyHostToDevice, s));
  BPM_TRY(h->peer_barrier(s));                  // every replica holds every shard
  // the caller's host array is one more "replica": accepted rows are stored into it from inside the phase
  // kernels (indexed by global chain id, hence the offset base; only this rank's rows are ever written)
  h->peers[h->n_peers++] = X_map - off;
  int rc = 0;
  for (int g = 0; g < n_gen && rc == 0; ++g) rc = h->generation<false>(st, k_gen0 + g, nullptr, nullptr, s);
  h->peers[--h->n_peers] = nullptr;
  if (rc) return rc;
  CU_TRY(cudaMemcpyAsync(lnl_host, st->lnl + lo, sizeof(double) * nloc, cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaMemcpyAsync(&acc1, h->counters, sizeof(acc1), cudaMemcpyDeviceToHost, s));
  CU_TRY(cudaStreamSynchronize(s));
  h->last_d2h_bytes = (acc1 - acc0) * (sizeof(double) * ld) + sizeof(double) * nloc + 2 * sizeof(acc0);
  return 0;
}

int bpm_host_entry_restart(bpm_handle h) {
  if (!h) return fail("null handle");
  h->h_mom_len = 0; h->h_pending = 0;
  return 0;
}

int bpm_last_d2h_bytes(bpm_handle h, uint64_t* bytes) {
  if (!h || !bytes) return fail("null argument");
  *bytes = h->last_d2h_bytes;
  return 0;
}

int bpm_dev_alloc(int32_t device, uint64_t bytes, void** dev_ptr) {
  if (!dev_ptr) return fail("null argument");
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaMalloc(dev_ptr, (size_t)bytes));
  return 0;
}
int bpm_dev_free(int32_t device, void* dev_ptr) {
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaFree(dev_ptr));
  return 0;
}
int bpm_ipc_export(int32_t device, const void* dev_ptr, unsigned char handle64[64]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (!dev_ptr || !handle64) return fail("null argument");
  CU_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t hd;
  CU_TRY(cudaIpcGetMemHandle(&hd, const_cast<void*>(dev_ptr)));
  memcpy(handle64, &hd, 64);
  return 0;
}
int bpm_ipc_open(int32_t device, const unsigned char handle64[64], void** dev_ptr) {
  if (!dev_ptr || !handle64) return fail("null argument");
  CU_TRY(cudaSetDevice(device));
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle64, 64);
  CU_TRY(cudaIpcOpenMemHandle(dev_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int bpm_ipc_close(int32_t device, void* dev_ptr) {
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaIpcCloseMemHandle(dev_ptr));
  return 0;
}
int bpm_peer_copy(int32_t device, void* dst, const void* src, uint64_t bytes, bpm_stream stream) {
  if (!dst || !src) return fail("null argument");
  CU_TRY(cudaSetDevice(device));
  CU_TRY(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDefault, (cudaStream_t)stream));
  return 0;
}
int bpm_set_peers(bpm_handle h, double* const* peer_X, int32_t n_peers) {
  if (!h) return fail("null handle");
  if (n_peers < 0 || n_peers > BPM_MAX_PEERS) return fail("n_peers must be in [0, 15]");
  if (n_peers > 0 && !peer_X) return fail("null peer array");
  for (int p = 0; p < n_peers; ++p) {
    if (!peer_X[p]) return fail("null peer pointer");
    h->peers[p] = peer_X[p];
  }
  h->n_peers = n_peers;
  return 0;
}

int bpm_sync_bytes(uint64_t* bytes) {
  if (!bytes) return fail("null argument");
  *bytes = (uint64_t)bpm::kSyncBytes;
  return 0;
}

int bpm_set_sync(bpm_handle h, void* const* blocks, int32_t rank, int32_t world) {
  if (!h) return fail("null handle");
  if (world == 0) { h->sync_on = false; return 0; }
  if (!blocks || world < 2 || world > bpm::kSyncRanks || rank < 0 || rank >= world) return fail("bad sync arguments");
  CU_TRY(cudaSetDevice(h->cfg.device));
  for (int r = 0; r < world; ++r) {
    if (!blocks[r]) return fail("null sync block");
    h->sync.blocks[r] = (unsigned char*)blocks[r];
  }
  if (!h->sync_err) {
    CU_TRY(cudaMalloc(&h->sync_err, sizeof(int32_t)));
    CU_TRY(cudaMemset(h->sync_err, 0, sizeof(int32_t)));
  }
  h->sync.rank = rank; h->sync.world = world; h->sync.err = h->sync_err;
  h->sync_epoch = 0; h->sync_cr_seq = 0;
  h->sync_on = true;
  return 0;
}

int bpm_peer_barrier(bpm_handle h, bpm_stream stream) {
  if (!h || !h->sync_on) return fail("bpm_peer_barrier: bpm_set_sync first");
  CU_TRY(cudaSetDevice(h->cfg.device));
  return h->peer_barrier((cudaStream_t)stream);
}

int bpm_sync_error(bpm_handle h, int32_t* err) {
  if (!h || !err) return fail("null argument");
  *err = 0;
  if (!h->sync_err) return 0;
  CU_TRY(cudaSetDevice(h->cfg.device));
  CU_TRY(cudaMemcpy(err, h->sync_err, sizeof(int32_t), cudaMemcpyDeviceToHost));
  return 0;
}

int bpm_flush(bpm_handle h, bpm_state* st, bpm_stream stream) {
  if (!h || !st || !st->X) return fail("null argument");
  CU_TRY(cudaSetDevice(h->cfg.device));
  return h->flush(st, (cudaStream_t)stream);
}

int bpm_moments_from_history(bpm_handle h, bpm_state* st, bpm_stream stream) {
  if (!h || !st || !st->history || !st->mean || !st->m2) return fail("null argument");
  CU_TRY(cudaSetDevice(h->cfg.device));
  BPM_TRY(h->flush(st, (cudaStream_t)stream));       // the last history row may still be pending
  const int64_t tot = (int64_t)(h->cfg.chain_hi - h->cfg.chain_lo) * h->cfg.ld;
  bpm::moments_from_history_kernel<<<cdiv(tot, 256), 256, 0, (cudaStream_t)stream>>>(
      st->history, st->hist_len, h->cfg.n_chains, h->cfg.dim, h->cfg.ld, h->cfg.chain_lo,
      h->cfg.chain_hi, st->mean, st->m2);
  CU_TRY(cudaGetLastError());
  return 0;
}

int bpm_profile(bpm_handle h, int32_t on) {
  if (!h) return fail("null handle");
  h->prof_on = on != 0;
  return 0;
}

int bpm_profile_read(bpm_handle h, double* ms_by_kind, int64_t* launches_by_kind) {
  if (!h || !ms_by_kind || !launches_by_kind) return fail("null argument");
  CU_TRY(cudaSetDevice(h->cfg.device));
  CU_TRY(cudaDeviceSynchronize());
  for (int k = 0; k < 8; ++k) { ms_by_kind[k] = 0.0; launches_by_kind[k] = 0; }
  for (auto& r : h->recs) {
    float ms = 0.f;
    CU_TRY(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_by_kind[r.kind] += ms;
    launches_by_kind[r.kind] += 1;
    h->ev_pool.push_back(r.a);
    h->ev_pool.push_back(r.b);
  }
  h->recs.clear();
  return 0;
}

int bpm_test_philox(const uint32_t* ctr, const uint32_t* key, uint32_t* out) {
  bpm::Philox4 q = bpm::philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1]);
  out[0] = q.x; out[1] = q.y; out[2] = q.z; out[3] = q.w;
  return 0;
}

int bpm_test_permutation(uint64_t seed, uint64_t g_abs, int32_t n, int32_t* out_perm) {
  if (n < 1 || !out_perm) return fail("bad argument");
  bpm::RngCtx r = bpm::make_rng(seed, g_abs);
  bpm::FeistelKey f = bpm::make_feistel(r, (uint32_t)n);
  for (int32_t j = 0; j < n; ++j) out_perm[j] = (int32_t)bpm::feistel_perm(f, (uint32_t)j);
  return 0;
}

int bpm_omega_track(bpm_handle h, int32_t on) {
  if (!h) return fail("null handle");
  CU_TRY(cudaSetDevice(h->cfg.device));
  const int N = h->cfg.n_chains;
  if (on) {
    if (!h->omega_sum) CU_TRY(cudaMalloc(&h->omega_sum, sizeof(double) * N));
    CU_TRY(cudaMemset(h->omega_sum, 0, sizeof(double) * N));
    h->omega_cnt = 0;
  }
  h->omega_on = on != 0;
  return 0;
}

int bpm_cov_track(bpm_handle h, int32_t on) {
  if (!h) return fail("null handle");
  CU_TRY(cudaSetDevice(h->cfg.device));
  const size_t n = (size_t)h->cfg.dim * (h->cfg.dim + 1);
  if (on) {
    if (!h->cov_acc) CU_TRY(cudaMalloc(&h->cov_acc, sizeof(double) * n));
    CU_TRY(cudaMemset(h->cov_acc, 0, sizeof(double) * n));
    h->cov_rows = 0;
  }
  h->cov_on = on != 0;
  return 0;
}

int bpm_cov_read(bpm_handle h, double* sum_host, double* cross_host, int64_t* generations) {
  if (!h || !sum_host || !cross_host || !generations) return fail("null argument");
  if (!h->cov_acc) return fail("bpm_cov_read: nothing tracked (bpm_cov_track)");
  CU_TRY(cudaSetDevice(h->cfg.device));
  CU_TRY(cudaDeviceSynchronize());
  const int d = h->cfg.dim;
  CU_TRY(cudaMemcpy(sum_host, h->cov_acc, sizeof(double) * d, cudaMemcpyDeviceToHost));
  CU_TRY(cudaMemcpy(cross_host, h->cov_acc + d, sizeof(double) * d * d, cudaMemcpyDeviceToHost));
  *generations = h->cov_rows;
  return 0;
}

int bpm_omega(bpm_handle h, double** sum_dev, int64_t* count) {
  if (!h) return fail("null handle");
  if (sum_dev) *sum_dev = h->omega_sum;
  if (count) *count = h->omega_cnt;
  return 0;
}

int bpm_outlier_reset(bpm_handle h, bpm_state* st, const double* omega, int32_t* flags_dev,
                      int32_t* n_reset, double* stats_host, bpm_stream stream) {
  if (!h || !st || !st->X || !st->lnl) return fail("null argument");
  CU_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  BPM_TRY(h->flush(st, s));        // a reset overwrites X: the pending row is the PRE-reset state
  const int N = h->cfg.n_chains;
  const bool sharded = h->cfg.chain_lo != 0 || h->cfg.chain_hi != N;
  if (!h->omega_buf) {
    CU_TRY(cudaMalloc(&h->omega_buf, sizeof(double) * 2 * (size_t)N));
    CU_TRY(cudaMalloc(&h->diag_out, sizeof(double) * 3));
    CU_TRY(cudaMalloc(&h->diag_i, sizeof(int32_t) * 2));
  }
  double* om = h->omega_buf;          // Omega of every chain
  double* sorted = h->omega_buf + N;
  if (omega) {
    CU_TRY(cudaMemcpyAsync(om, omega, sizeof(double) * N, cudaMemcpyDeviceToDevice, s));
  } else {
    if (sharded) return fail("bpm_outlier_reset: a sharded handle needs the all-gathered omega array");
    if (!h->omega_sum || h->omega_cnt < 1) return fail("bpm_outlier_reset: no Omega tracked (bpm_omega_track)");
    bpm::omega_mean_kernel<<<cdiv(N, 256), 256, 0, s>>>(h->omega_sum, 1.0 / (double)h->omega_cnt, N, om);
  }
  size_t need = 0;
  CU_TRY(cub::DeviceRadixSort::SortKeys(nullptr, need, om, sorted, N, 0, 64, s));
  if (need > h->sort_tmp_bytes) {
    CU_TRY(cudaStreamSynchronize(s));
    cudaFree(h->sort_tmp); h->sort_tmp = nullptr;
    CU_TRY(cudaMalloc(&h->sort_tmp, need));
    h->sort_tmp_bytes = need;
  }
  CU_TRY(cub::DeviceRadixSort::SortKeys(h->sort_tmp, need, om, sorted, N, 0, 64, s));
  bpm::iqr_threshold_kernel<<<1, 32, 0, s>>>(sorted, N, h->diag_out);
  bpm::argmax_kernel<<<1, 1024, 0, s>>>(om, N, h->diag_i);
  CU_TRY(cudaMemsetAsync(h->diag_i + 1, 0, sizeof(int32_t), s));
  const int nloc = h->cfg.chain_hi - h->cfg.chain_lo;
  bpm::outlier_reset_kernel<<<cdiv((int64_t)nloc * 32, 256), 256, 0, s>>>(
      st->X, st->lnl, omega ? nullptr : h->omega_sum, om, h->diag_out, h->diag_i, h->cfg.chain_lo,
      h->cfg.chain_hi, h->cfg.dim, h->cfg.ld, flags_dev, h->diag_i + 1);
  CU_TRY(cudaGetLastError());
  if (n_reset || stats_host) {
    CU_TRY(cudaStreamSynchronize(s));
    if (n_reset) CU_TRY(cudaMemcpy(n_reset, h->diag_i + 1, sizeof(int32_t), cudaMemcpyDeviceToHost));
    if (stats_host) {
      int32_t best = 0;
      CU_TRY(cudaMemcpy(stats_host, h->diag_out, sizeof(double) * 3, cudaMemcpyDeviceToHost));
      CU_TRY(cudaMemcpy(&best, h->diag_i, sizeof(int32_t), cudaMemcpyDeviceToHost));
      stats_host[3] = (double)best;
    }
  }
  return 0;
}

int bpm_rhat(bpm_handle h, const bpm_state* st, int64_t t0, double* rhat_host, bpm_stream stream) {
  if (!h || !st || !rhat_host) return fail("null argument");
  if (st->pending) return fail("bpm_rhat: the state has a pending row; call bpm_flush first");
  CU_TRY(cudaSetDevice(h->cfg.device));
  cudaStream_t s = (cudaStream_t)stream;
  const int nloc = h->cfg.chain_hi - h->cfg.chain_lo, d = h->cfg.dim, ld = h->cfg.ld;
  if (nloc < 2) return fail("bpm_rhat: at least two chains needed");
  if (!h->rh_out) CU_TRY(cudaMalloc(&h->rh_out, sizeof(double) * d));
  const double* mean = st->mean;
  const double* m2 = st->m2;
  double rows;
  if (t0 < 0) {   // streaming: the running moments cover mom_len rows
    if (!mean || !m2) return fail("bpm_rhat: running moments missing");
    rows = (double)(st->mom_len > 0 ? st->mom_len : st->hist_len);
  } else {
    if (!st->history) return fail("bpm_rhat: history missing (pass t0 < 0 for the running moments)");
    if (t0 >= st->hist_len) return fail("bpm_rhat: t0 beyond the history");
    if (!h->rh_mean) {
      CU_TRY(cudaMalloc(&h->rh_mean, sizeof(double) * (size_t)nloc * ld));
      CU_TRY(cudaMalloc(&h->rh_m2, sizeof(double) * (size_t)nloc * ld));
    }
    bpm::history_moments_kernel<<<cdiv((int64_t)nloc * ld, 256), 256, 0, s>>>(
        st->history, t0, st->hist_len, nloc, d, ld, h->rh_mean, h->rh_m2);
    mean = h->rh_mean; m2 = h->rh_m2;
    rows = (double)(st->hist_len - t0);
  }
  if (rows < 2.0) return fail("bpm_rhat: at least two rows needed");
  bpm::rhat_kernel<<<d, 256, 0, s>>>(mean, m2, nloc, ld, rows, h->rh_out);
  CU_TRY(cudaGetLastError());
  CU_TRY(cudaStreamSynchronize(s));
  CU_TRY(cudaMemcpy(rhat_host, h->rh_out, sizeof(double) * d, cudaMemcpyDeviceToHost));
  return 0;
}

}  // extern "C"
