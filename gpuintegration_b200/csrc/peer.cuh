// In-kernel all-reduce over NVLink peer memory (several ranks, one process per GPU).
//
// Every rank owns a small mailbox in its HBM that all peers can write (mapped with cudaIpc over
// NVLink 5 / NVSwitch).  A contribution travels as self-validating 8-byte PACKETS: the low word is
// 32 bits of payload, the high word the number of the exchange.  An aligned 8-byte store is
// single-copy atomic, so a receiver that sees the expected exchange number in a packet also sees
// its payload: no fence, no separate flag, no second round trip -- an exchange costs one NVLink
// one-way latency instead of a system-scope fence (which waits for the write acknowledgements)
// plus a flag store.  For exchange number s a rank stores its packets into slot [s & 1][own rank]
// of EVERY peer's mailbox, then polls the slots [s & 1][q] of its own mailbox and combines the
// contributions in rank order, so every rank obtains bit-identical totals without the host and
// without a separate collective launch.  Two slots per peer are enough: a rank can only start
// exchange s + 1 after it has read every contribution of exchange s, and a peer needs that rank's
// s + 1 before it can move on to s + 2.
//
// Two independent channels share one mailbox allocation:
//   search channel  (kBoxDoubles)    probes of the heuristic threshold search (threshold_search.cuh)
//   stats channel   (kStatsDoubles)  the iteration statistics, fused into the last block of the
//                                    classify / threshold kernels (stages.cuh, dev::finalize)
// The reference has no multi-GPU path; NCCL remains the transport for bulk region rebalancing.
#ifndef PGN_PEER_CUH
#define PGN_PEER_CUH

#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace pgn {

constexpr int kMaxPeers = 16;
constexpr int kBoxDoubles = 8;     // search channel, per slot
constexpr int kStatsDoubles = 32;  // stats channel, per slot (= kStatsLen of stages.cuh)

// byte offsets inside a mailbox: [2 parities][kMaxPeers senders][2 packets per double]
constexpr size_t kBoxOff = 0;
constexpr size_t kStatsOff = kBoxOff + sizeof(unsigned long long) * 2 * kMaxPeers * 2 * kBoxDoubles;
constexpr size_t kMailboxBytes = kStatsOff + sizeof(unsigned long long) * 2 * kMaxPeers * 2 * kStatsDoubles;

struct PeerBox {
  char* base[kMaxPeers];  // base[q]: peer q's mailbox (own rank: the local allocation)
  int nranks, rank;
  long long spin_limit;   // clock64() ticks before giving up (a lost peer must not hang the GPU for ever)
};

// exchange number -> packet tag (never 0: a zeroed mailbox matches no exchange)
__device__ __forceinline__ uint32_t
packet_tag(unsigned long long seq)
{
  return static_cast<uint32_t>(seq % 0xffffffffull) + 1u;
}

// All-reduce of W doubles over the ranks through the channel at byte offset `off` (kBoxOff / kStatsOff, whose
// slots hold `slot_doubles` doubles).  `mine`: this rank's contribution (shared or global memory, read by all
// threads); `s_recv`: shared scratch of kMaxPeers * 2 * W words.  Block-collective; afterwards
// s_recv[(q * W + j) * 2 + {0,1}] holds the two halves of rank q's j-th value.  Returns false (in every thread)
// when a peer did not answer in time.
template <int W>
__device__ __forceinline__ bool
peer_exchange(const PeerBox& box, size_t off, int slot_doubles, unsigned long long seq, const double* mine,
              uint32_t* s_recv, int* s_ok)
{
  const int par = static_cast<int>(seq & 1ull);
  const uint32_t tag = packet_tag(seq);
  if (threadIdx.x == 0)
    *s_ok = 1;
  __syncthreads();
  const int total = box.nranks * 2 * W;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int q = idx / (2 * W), k = idx - q * (2 * W);
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(mine[k >> 1]));
    const uint32_t half = (k & 1) ? static_cast<uint32_t>(bits >> 32) : static_cast<uint32_t>(bits);
    volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(box.base[q] + off) +
                                       (static_cast<size_t>(par) * kMaxPeers + box.rank) * 2 * slot_doubles + k;
    *dst = (static_cast<unsigned long long>(tag) << 32) | half;
  }
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int q = idx / (2 * W), k = idx - q * (2 * W);
    const volatile unsigned long long* src =
      reinterpret_cast<const volatile unsigned long long*>(box.base[box.rank] + off) +
      (static_cast<size_t>(par) * kMaxPeers + q) * 2 * slot_doubles + k;
    const long long t0 = clock64();
    unsigned long long pkt = *src;
    while (static_cast<uint32_t>(pkt >> 32) != tag) {
      if (clock64() - t0 > box.spin_limit) {
        *s_ok = 0;
        break;
      }
      pkt = *src;
    }
    s_recv[idx] = static_cast<uint32_t>(pkt);
  }
  __syncthreads();
  return *s_ok != 0;
}

__device__ __forceinline__ double
peer_value(const uint32_t* s_recv, int W, int q, int j)
{
  const uint32_t lo = s_recv[(q * W + j) * 2], hi = s_recv[(q * W + j) * 2 + 1];
  return __longlong_as_double(static_cast<long long>((static_cast<unsigned long long>(hi) << 32) | lo));
}

// Sum of W <= kBoxDoubles doubles over the ranks (rank order: identical totals everywhere).  Block-collective
// (any block size); `s_mine` [W] and `s_recv` [kMaxPeers * 2 * W] are shared scratch.  Returns false (in every
// thread) when a peer did not answer in time.
template <int W>
__device__ __forceinline__ bool
peer_allreduce_sum(const PeerBox& box, unsigned long long seq, double (&v)[W], double* s_mine, uint32_t* s_recv,
                   int* s_ok)
{
  static_assert(W <= kBoxDoubles, "search channel slot too small");
  __syncthreads();  // s_mine / s_recv of the previous exchange have been consumed
  if (threadIdx.x < W)
    s_mine[threadIdx.x] = v[threadIdx.x];
  __syncthreads();
  const bool ok = peer_exchange<W>(box, kBoxOff, kBoxDoubles, seq, s_mine, s_recv, s_ok);
#pragma unroll
  for (int j = 0; j < W; ++j) {
    double t = 0.0;
    if (ok)
      for (int q = 0; q < box.nranks; ++q)
        t += peer_value(s_recv, W, q, j);
    v[j] = t;
  }
  return ok;
}

// All-reduce of the kStatsDoubles-vector `local` into `global` (both in this rank's memory): SUM in
// rank order for every slot except [max_lo, max_hi), which take the maximum.  Block-collective, the
// block must have at least kStatsDoubles threads.  On a time-out the vector is filled with NaN (the
// host driver notices) and false is returned.
__device__ __forceinline__ bool
peer_allreduce_stats(const PeerBox& box, unsigned long long seq, const double* local, double* global,
                     int max_lo, int max_hi, int* s_ok)
{
  __shared__ uint32_t s_stats_recv[kMaxPeers * 2 * kStatsDoubles];
  __syncthreads();  // `local` was written by this block
  const bool ok = peer_exchange<kStatsDoubles>(box, kStatsOff, kStatsDoubles, seq, local, s_stats_recv, s_ok);
  if (threadIdx.x < kStatsDoubles) {
    double acc;
    if (!ok) {
      acc = __longlong_as_double(0x7ff8000000000000LL);
    } else {
      const int j = threadIdx.x;
      const bool is_max = j >= max_lo && j < max_hi;
      acc = peer_value(s_stats_recv, kStatsDoubles, 0, j);
      for (int q = 1; q < box.nranks; ++q) {
        const double v = peer_value(s_stats_recv, kStatsDoubles, q, j);
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
