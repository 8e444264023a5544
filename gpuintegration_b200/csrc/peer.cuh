// In-kernel all-reduce over NVLink peer memory (several ranks, one process per GPU).
//
// Every rank owns a small mailbox in its HBM that all peers can write (mapped with cudaIpc over
// NVLink 5 / NVSwitch).  For exchange number s a rank stores its contribution into slot
// [s & 1][own rank] of EVERY peer's mailbox, fences, then raises that slot's flag to s; it then waits
// until all slots of its own mailbox carry flag s and combines the contributions in rank order, so
// every rank obtains bit-identical totals without the host and without a separate collective launch.
// Two slots per peer are enough: a rank can only start exchange s + 1 after it has read every
// contribution of exchange s, and a peer needs that rank's s + 1 before it can move on to s + 2.
//
// Two independent channels share one mailbox allocation:
//   search channel  (3 doubles)  per probe of the heuristic threshold search (threshold_search.cuh)
//   stats channel   (32 doubles) the iteration statistics, fused into the last block of the
//                                classify / threshold kernels (stages.cuh, dev::finalize)
// The reference has no multi-GPU path; NCCL remains the transport for bulk region rebalancing.
#ifndef PGN_PEER_CUH
#define PGN_PEER_CUH

#include <cuda_runtime.h>

#include <cstddef>

namespace pgn {

constexpr int kMaxPeers = 16;
constexpr int kBoxDoubles = 4;     // search channel, per slot
constexpr int kStatsDoubles = 32;  // stats channel, per slot (= kStatsLen of stages.cuh)

// byte offsets inside a mailbox
constexpr size_t kBoxDataOff = 0;
constexpr size_t kBoxFlagOff = kBoxDataOff + sizeof(double) * 2 * kMaxPeers * kBoxDoubles;
constexpr size_t kStatsDataOff = kBoxFlagOff + sizeof(unsigned long long) * 2 * kMaxPeers;
constexpr size_t kStatsFlagOff = kStatsDataOff + sizeof(double) * 2 * kMaxPeers * kStatsDoubles;
constexpr size_t kMailboxBytes = kStatsFlagOff + sizeof(unsigned long long) * 2 * kMaxPeers;

struct PeerBox {
  char* base[kMaxPeers];  // base[q]: peer q's mailbox (own rank: the local allocation)
  int nranks, rank;
  long long spin_limit;   // clock64() ticks before giving up (a lost peer must not hang the GPU for ever)
};

__device__ __forceinline__ double*
box_data(const PeerBox& b, int q)
{
  return reinterpret_cast<double*>(b.base[q] + kBoxDataOff);
}
__device__ __forceinline__ unsigned long long*
box_flag(const PeerBox& b, int q)
{
  return reinterpret_cast<unsigned long long*>(b.base[q] + kBoxFlagOff);
}
__device__ __forceinline__ double*
stats_data(const PeerBox& b, int q)
{
  return reinterpret_cast<double*>(b.base[q] + kStatsDataOff);
}
__device__ __forceinline__ unsigned long long*
stats_flag(const PeerBox& b, int q)
{
  return reinterpret_cast<unsigned long long*>(b.base[q] + kStatsFlagOff);
}

// Sum of three doubles over the ranks.  Block-collective (any block size >= nranks); returns false
// (in every thread) when a peer did not answer in time.
__device__ __forceinline__ bool
peer_allreduce3(const PeerBox& box, unsigned long long seq, double& a, double& b, double& c, int* s_ok)
{
  const int par = static_cast<int>(seq & 1ull);
  if (threadIdx.x == 0)
    *s_ok = 1;
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < box.nranks) {
    const int q = threadIdx.x;
    volatile double* slot = box_data(box, q) + (par * kMaxPeers + box.rank) * kBoxDoubles;
    slot[0] = a;
    slot[1] = b;
    slot[2] = c;
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(box_flag(box, q) + par * kMaxPeers + box.rank) = seq;
    // wait for peer q's contribution to arrive in OUR mailbox
    volatile unsigned long long* mine = box_flag(box, box.rank) + par * kMaxPeers + q;
    const long long t0 = clock64();
    while (*mine < seq) {
      if (clock64() - t0 > box.spin_limit) {
        *s_ok = 0;
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  const bool ok = *s_ok != 0;
  double ta = 0.0, tb = 0.0, tc = 0.0;
  if (ok) {
    const volatile double* base = box_data(box, box.rank) + par * kMaxPeers * kBoxDoubles;
    for (int q = 0; q < box.nranks; ++q) {
      ta += base[q * kBoxDoubles + 0];
      tb += base[q * kBoxDoubles + 1];
      tc += base[q * kBoxDoubles + 2];
    }
  }
  __syncthreads();
  a = ta;
  b = tb;
  c = tc;
  return ok;
}

// All-reduce of the kStatsDoubles-vector `local` into `global` (both in this rank's memory): SUM in
// rank order for every slot except [max_lo, max_hi), which take the maximum.  Block-collective, the
// block must have at least 32 threads.  On a time-out the vector is filled with NaN (the host driver
// notices) and false is returned.
__device__ __forceinline__ bool
peer_allreduce_stats(const PeerBox& box, unsigned long long seq, const double* local, double* global,
                     int max_lo, int max_hi, int* s_ok)
{
  const int par = static_cast<int>(seq & 1ull);
  if (threadIdx.x == 0)
    *s_ok = 1;
  __syncthreads();
  // thread t writes element t % 32 to peers t / 32, t / 32 + blockDim / 32, ...
  const int elem = threadIdx.x & 31;
  const double mine_v = local[elem];
  for (int q = threadIdx.x >> 5; q < box.nranks; q += blockDim.x >> 5) {
    volatile double* slot = stats_data(box, q) + (par * kMaxPeers + box.rank) * kStatsDoubles;
    slot[elem] = mine_v;
  }
  __threadfence_system();
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < box.nranks) {
    const int q = threadIdx.x;
    *reinterpret_cast<volatile unsigned long long*>(stats_flag(box, q) + par * kMaxPeers + box.rank) = seq;
    volatile unsigned long long* arrived = stats_flag(box, box.rank) + par * kMaxPeers + q;
    const long long t0 = clock64();
    while (*arrived < seq) {
      if (clock64() - t0 > box.spin_limit) {
        *s_ok = 0;
        break;
      }
    }
    __threadfence_system();
  }
  __syncthreads();
  const bool ok = *s_ok != 0;
  if (threadIdx.x < kStatsDoubles) {
    double acc;
    if (!ok) {
      acc = __longlong_as_double(0x7ff8000000000000LL);
    } else {
      const volatile double* base = stats_data(box, box.rank) + par * kMaxPeers * kStatsDoubles;
      const bool is_max = static_cast<int>(threadIdx.x) >= max_lo && static_cast<int>(threadIdx.x) < max_hi;
      acc = base[threadIdx.x];
      for (int q = 1; q < box.nranks; ++q) {
        const double v = base[q * kStatsDoubles + threadIdx.x];
        acc = is_max ? fmax(acc, v) : acc + v;
      }
    }
    global[threadIdx.x] = acc;
  }
  __syncthreads();
  return ok;
}

}  // namespace pgn
#endif
