// Cross-GPU synchronisation of one sharded generation over peer memory (no NCCL on the data path).
//
// The reference puts an MPI Allgather before each half-phase and a Barrier at the end of a
// generation (bipymc/demc.py:93,116,135).  Here accepted rows already travel inside the phase
// kernels (store_peers*, step.cuh); what is left between half-phases is a BARRIER, and after the
// second half-phase of a DREAM generation the all-reduce of 2 n_cr CR statistics.  Both are one tiny
// kernel on the generation's own stream, signalling through flag words that every rank keeps in an
// IPC-mapped "sync block":
//
//   struct layout of a sync block (bytes):
//     [0, 16 * 8)              flags[src]     uint64 epoch last signalled by rank src   (BPM_MAX_PEERS + 1 slots)
//     [128, 128 + 2*16*32*8)   cr[par][src][2 * BPM_MAX_CR] doubles: rank src's CR partial sums of the
//                              generation with parity par (double-buffered: a fast rank may already
//                              publish generation g+1 while a slow one still sums generation g)
//
// Rank r signals epoch e by storing e into flags[r] of EVERY rank's block (its own included) with a
// system-scope release, then waits until all flags of its own block reach e (system-scope acquire).
// Stream order puts the phase kernel's peer stores before the signalling kernel; the release /
// acquire pair orders them before anything the waiting rank launches afterwards.  Epochs only grow,
// so no reset is ever needed; all ranks run the same sequence of barriers, so they agree on e.
// A rank that never arrives must not hang the box: the wait gives up after ~4 s and raises a sticky
// error flag the host turns into an exception.
#pragma once
#include <stdint.h>
#include "../../include/bipymc_b200.h"
#include "kernels_generic.cuh"   // cr_apply

namespace bpm {

constexpr int kSyncRanks = BPM_MAX_PEERS + 1;
constexpr size_t kSyncFlagBytes = 128;
constexpr size_t kSyncCrDoubles = 2 * (size_t)kSyncRanks * 2 * BPM_MAX_CR;
constexpr size_t kSyncBytes = kSyncFlagBytes + sizeof(double) * kSyncCrDoubles;

struct SyncArgs {
  unsigned char* blocks[kSyncRanks];   // sync block of every rank, indexed by rank (own block included)
  int32_t rank, world;
  int32_t* err;                        // sticky: 1 = a peer did not arrive in time
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// threads 0 .. world-1 of one block; returns after every rank signalled `epoch`
__device__ __forceinline__ void sync_signal_and_wait(const SyncArgs& a, unsigned long long epoch) {
  const int t = threadIdx.x;
  __threadfence_system();
  __syncthreads();
  if (t < a.world)
    st_release_sys(reinterpret_cast<unsigned long long*>(a.blocks[t]) + a.rank, epoch);
  if (t < a.world) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(a.blocks[a.rank]) + t;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) < epoch) {
      if (clock64() - t0 > 8000000000ll) {     // ~4 s at 1.9 GHz
        *a.err = 1;
        break;
      }
      __nanosleep(200);
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(32) peer_barrier_kernel(const SyncArgs a, unsigned long long epoch) {
  sync_signal_and_wait(a, epoch);
}

// CR all-reduce folded into the barrier that closes a DREAM generation: publish this rank's partial
// sums (part[2 n_cr], written by cr_update_kernel) into every rank's block, barrier, add the world's
// partials in RANK order (so every rank gets bit-identical sums), apply the p_cr update
// (dream.py:132-140).
__global__ void __launch_bounds__(64) peer_cr_exchange_kernel(const SyncArgs a, unsigned long long epoch, int par,
                                                              double* __restrict__ part, int n_cr,
                                                              double* __restrict__ dm, double* __restrict__ cnt,
                                                              double* __restrict__ p_cr) {
  const int nv = 2 * n_cr;
  const size_t slot = ((size_t)par * kSyncRanks + a.rank) * 2 * BPM_MAX_CR;
  for (int idx = threadIdx.x; idx < a.world * nv; idx += blockDim.x) {
    const int p = idx / nv, v = idx - p * nv;
    reinterpret_cast<double*>(a.blocks[p] + kSyncFlagBytes)[slot + v] = part[v];
  }
  sync_signal_and_wait(a, epoch);
  __shared__ double red[2 * BPM_MAX_CR];
  if ((int)threadIdx.x < nv) {
    const volatile double* own = reinterpret_cast<const volatile double*>(a.blocks[a.rank] + kSyncFlagBytes);
    double w = 0.0;
    for (int r = 0; r < a.world; ++r) w += own[((size_t)par * kSyncRanks + r) * 2 * BPM_MAX_CR + threadIdx.x];
    red[threadIdx.x] = w;
    part[threadIdx.x] = w;            // the reduced sums replace the local partials (bpm_cr_partials readers)
  }
  __syncthreads();
  if (threadIdx.x == 0) cr_apply(red, n_cr, dm, cnt, p_cr);
}

// Sharded end-to-end entry: fused "host -> device copy + all-gather".  Every rank streams ITS shard of
// the population out of the caller's pinned (mapped) host array -- 16-byte zero-copy reads over PCIe,
// many in flight per thread -- and stores each chunk into its own replica and into every peer replica
// (NVLink stores), so the replicas are complete when the PCIe read ends; the reference's
// comm.Allgather of demc.py:93 rides under the host copy instead of following it.
struct ShardInArgs {
  const double2* src;             // mapped host shard [n2] double2
  double2* dst[kSyncRanks];       // the shard's rows in every replica (own first)
  int32_t n_dst;
  int64_t n2;
};
__global__ void __launch_bounds__(512) shard_in_kernel(const ShardInArgs a) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < a.n2; i += 4 * stride) {
    double2 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = a.src[i + u * stride];
    for (int p = 0; p < a.n_dst; ++p)
#pragma unroll
      for (int u = 0; u < 4; ++u) a.dst[p][i + u * stride] = v[u];
  }
  for (; i < a.n2; i += stride) {
    const double2 v = a.src[i];
    for (int p = 0; p < a.n_dst; ++p) a.dst[p][i] = v;
  }
}

}  // namespace bpm
