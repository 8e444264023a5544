// HBM-bound stages of one PAGANI iteration, written for sm_100a: every stage is ONE pass over the
// SoA arrays with fully coalesced accesses, fixed-shape (hence run-to-run deterministic) reductions
// and no host round trips inside a stage.
//
//   k_uniform_split   create_uniform_split             (reference Sub_regions.cuh:11-36)
//   k_classify        RefineError + set_array_to_value + 2x thrust::reduce + 2x thrust::inner_product
//                     + the flag count                  (Phases.cuh:219-268, PaganiUtils.cuh:310-311,
//                                                        368-385, thrust_utils.cuh:16-47)
//   k_threshold       device_set_true_for_larger_than + reduce + inner_product of one probe of
//                     Heuristic_classifier::classify    (heuristic_classifier.cuh:60-72, 283-325)
//   k_compact         thrust::exclusive_scan + alignRegions [+ divideIntervalsGPU when fused]
//                                                        (Sub_region_filter.cuh:8-47, 61-139,
//                                                         Sub_region_splitter.cuh:8-46)
//   k_split           divideIntervalsGPU alone          (Sub_region_splitter.cuh:8-46)
//
// Regions are processed in tiles of kTile = 1024 (one 256-thread block, 4 rounds).  A stage that
// (re)writes the active flags also leaves the number of active regions of each tile and their
// exclusive prefix, so compaction never needs a separate scan over the flags.
#ifndef PGN_STAGES_CUH
#define PGN_STAGES_CUH

#include <cuda_runtime.h>

#include <cstdint>

#include "peer.cuh"

namespace pgn {

constexpr int kBlock = 256;
constexpr int kRounds = 4;
constexpr int kTile = kBlock * kRounds;

// slots of the per-stage statistics vector (device + pinned host mirror)
enum Stat {
  kSumEst = 0,
  kSumErr = 1,
  kSumActEst = 2,
  kSumActErr = 3,
  kNumActive = 4,
  kMinErr = 5,
  kMaxErr = 6,
  kStatCount = 8
};

struct TileScratch {
  double* partials;       // [tiles][kStatCount]
  uint32_t* tile_count;   // [tiles]   active regions per tile
  uint64_t* tile_offset;  // [tiles]   exclusive prefix of tile_count
  unsigned int* done;     // arrival counter (self-resetting)
  double* stats;          // [kStatsLen] final values of this rank, see below
  int rank, nranks;       // for the one-hot per-rank active count
  // several ranks with NVLink peer mailboxes: the last block also all-reduces `stats` into
  // `stats_glob` inside the kernel (peer.cuh); box.nranks <= 1 disables it
  double* stats_glob;
  unsigned long long seq;
  PeerBox box;
  // zero-copy publication: when set, the kernel's last block also stores `stats` (and `stats_glob`) into this
  // host-mapped buffer [2 * kStatsLen doubles + one 64-bit word] and then raises the word to host_seq, so the host
  // driver polls pinned memory instead of enqueueing a copy and synchronising the stream after every iteration
  double* host_out;
  unsigned long long host_seq;
};
// layout of `stats`: [0,8) the Stat slots; [8] = -min(err), [9] = max(err) (so one MAX all-reduce
// serves both); [16 + r] = active regions of rank r (zero except the own rank: a SUM all-reduce of
// the whole vector leaves every rank with every rank's count).
constexpr int kNegMinSlot = 8;
constexpr int kMaxSlot = 9;
constexpr int kRankSlot = 16;
constexpr int kMaxRanks = 16;
constexpr int kStatsLen = kRankSlot + kMaxRanks;
static_assert(kStatsLen == kStatsDoubles && kMaxRanks == kMaxPeers, "stats vector and mailbox slot must agree");

__global__ void
k_uniform_split(double* left, double* length, size_t n, int parts, int ndim, double edge)
{
  const size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n)
    return;
  size_t q = t;
  for (int d = 0; d < ndim; ++d) {
    const size_t id = q % parts;
    q /= parts;
    left[static_cast<size_t>(d) * n + t] = static_cast<double>(id) * edge;
    length[static_cast<size_t>(d) * n + t] = edge;
  }
}

namespace dev {

struct Acc {
  double v[5];  // the five sums
  double mn, mx;
};

__device__ __forceinline__ void
acc_zero(Acc& a)
{
#pragma unroll
  for (int i = 0; i < 5; ++i)
    a.v[i] = 0.0;
  a.mn = __longlong_as_double(0x7ff0000000000000LL);   // +inf
  a.mx = __longlong_as_double(0xfff0000000000000LL);   // -inf
}

__device__ __forceinline__ void
acc_merge(Acc& a, const Acc& b)
{
#pragma unroll
  for (int i = 0; i < 5; ++i)
    a.v[i] += b.v[i];
  a.mn = fmin(a.mn, b.mn);
  a.mx = fmax(a.mx, b.mx);
}

__device__ __forceinline__ Acc
acc_shfl_xor(const Acc& a, int off)
{
  Acc b;
#pragma unroll
  for (int i = 0; i < 5; ++i)
    b.v[i] = __shfl_xor_sync(0xffffffffu, a.v[i], off);
  b.mn = __shfl_xor_sync(0xffffffffu, a.mn, off);
  b.mx = __shfl_xor_sync(0xffffffffu, a.mx, off);
  return b;
}

// fixed-shape block reduction; result valid in thread 0
__device__ __forceinline__ Acc
block_reduce(Acc a, Acc* s_warp /*[kBlock/32]*/)
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    acc_merge(a, acc_shfl_xor(a, off));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0)
    s_warp[w] = a;
  __syncthreads();
  if (w == 0) {
    if (lane < kBlock / 32)
      a = s_warp[lane];
    else
      acc_zero(a);
#pragma unroll
    for (int off = (kBlock / 32) / 2; off > 0; off >>= 1)
      acc_merge(a, acc_shfl_xor(a, off));
  }
  __syncthreads();
  return a;
}

// Last step of a statistics kernel (one block): copy the final vectors into host-mapped memory, fence, raise the flag.
__device__ __forceinline__ void
publish_to_host(const TileScratch& ts)
{
  if (ts.host_out == nullptr)
    return;
  __syncthreads();  // ts.stats / ts.stats_glob were written by threads of this block
  if (threadIdx.x < kStatsDoubles) {
    ts.host_out[threadIdx.x] = ts.stats[threadIdx.x];
    ts.host_out[kStatsDoubles + threadIdx.x] = ts.box.nranks > 1 ? ts.stats_glob[threadIdx.x] : ts.stats[threadIdx.x];
    __threadfence_system();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    *reinterpret_cast<volatile unsigned long long*>(ts.host_out + 2 * kStatsDoubles) = ts.host_seq;
}

// Executed by every block after it stored its partial: the last block to arrive folds the
// partials in a fixed order, writes the final statistics and the exclusive prefix of the tile
// counts, and re-arms the counter.
__device__ __forceinline__ void
finalize(const TileScratch& ts, unsigned int tiles, bool with_offsets, Acc* s_warp, uint64_t* s_scan,
         unsigned int parts = 0 /* partial records and arriving blocks; 0: one per tile */)
{
  if (parts == 0)
    parts = tiles;
  __shared__ bool s_last;
  // only thread 0's stores (the tile partial and the tile count) are read by the last block: the fence
  // belongs to that thread alone, the other threads' region outputs need no ordering before kernel end
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = atomicAdd(ts.done, 1u) == parts - 1;
  }
  __syncthreads();
  if (!s_last)
    return;
  __threadfence();
  // This block runs alone at the end of the kernel: everything it reads comes from L2 at ~0.7 us per dependent
  // round trip, so the loads of several records / count vectors are issued before the first use (the order of the
  // merges, and with it every sum, is the same as in a plain loop).
  Acc a;
  acc_zero(a);
  {
    constexpr int kAhead = 4;
    unsigned int b = threadIdx.x;
    for (; b + (kAhead - 1) * kBlock < parts; b += kAhead * kBlock) {
      double2 r[kAhead][4];  // a 64-byte record as four 16-byte loads
#pragma unroll
      for (int k = 0; k < kAhead; ++k) {
        const double2* p = reinterpret_cast<const double2*>(ts.partials + static_cast<size_t>(b + k * kBlock) * kStatCount);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          r[k][q] = __ldcg(&p[q]);
      }
#pragma unroll
      for (int k = 0; k < kAhead; ++k) {
        Acc t;
        t.v[0] = r[k][0].x, t.v[1] = r[k][0].y, t.v[2] = r[k][1].x, t.v[3] = r[k][1].y, t.v[4] = r[k][2].x;
        t.mn = r[k][2].y;
        t.mx = r[k][3].x;
        acc_merge(a, t);
      }
    }
    for (; b < parts; b += kBlock) {
      const double* p = ts.partials + static_cast<size_t>(b) * kStatCount;
      Acc t;
#pragma unroll
      for (int i = 0; i < 5; ++i)
        t.v[i] = __ldcg(&p[i]);
      t.mn = __ldcg(&p[kMinErr]);
      t.mx = __ldcg(&p[kMaxErr]);
      acc_merge(a, t);
    }
  }
  a = block_reduce(a, s_warp);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < 5; ++i)
      ts.stats[i] = a.v[i];
    ts.stats[kMinErr] = a.mn;
    ts.stats[kMaxErr] = a.mx;
    ts.stats[kNegMinSlot] = -a.mn;
    ts.stats[kMaxSlot] = a.mx;
    for (int r = 0; r < ts.nranks; ++r)
      ts.stats[kRankSlot + r] = r == ts.rank ? a.v[kNumActive] : 0.0;
    *ts.done = 0u;
  }
  if (with_offsets) {
    // thread t owns the tiles [t * chunk, (t + 1) * chunk), chunk a multiple of four: counts travel as uint4
    const unsigned int chunk = (((tiles + kBlock - 1) / kBlock) + 3u) & ~3u;
    const unsigned int lo = threadIdx.x * chunk < tiles ? threadIdx.x * chunk : tiles;
    const unsigned int hi = lo + chunk < tiles ? lo + chunk : tiles;
    const unsigned int hi4 = lo + ((hi - lo) & ~3u);  // whole vectors
    uint64_t local = 0;
    {
      unsigned int b = lo;
#pragma unroll 8
      for (; b < hi4; b += 4) {
        const uint4 c = __ldcg(reinterpret_cast<const uint4*>(ts.tile_count + b));
        local += static_cast<uint64_t>(c.x) + c.y + c.z + c.w;
      }
      for (; b < hi; ++b)
        local += __ldcg(&ts.tile_count[b]);
    }
    // inclusive scan over the 256 chunk sums: shuffles inside a warp, the eight warp totals through shared memory
    uint64_t incl = local;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint64_t up = __shfl_up_sync(0xffffffffu, incl, off);
      if (lane >= off)
        incl += up;
    }
    if (lane == 31)
      s_scan[w] = incl;
    __syncthreads();
    uint64_t before = 0;
#pragma unroll
    for (int k = 0; k < kBlock / 32; ++k)
      before += k < w ? s_scan[k] : 0;
    uint64_t run = before + incl - local;
    {
      unsigned int b = lo;
#pragma unroll 4
      for (; b < hi4; b += 4) {
        const uint4 c = __ldcg(reinterpret_cast<const uint4*>(ts.tile_count + b));
        ulonglong2 o0, o1;
        o0.x = run;
        o0.y = o0.x + c.x;
        o1.x = o0.y + c.y;
        o1.y = o1.x + c.z;
        run = o1.y + c.w;
        reinterpret_cast<ulonglong2*>(ts.tile_offset + b)[0] = o0;
        reinterpret_cast<ulonglong2*>(ts.tile_offset + b)[1] = o1;
      }
      for (; b < hi; ++b) {
        ts.tile_offset[b] = run;
        run += __ldcg(&ts.tile_count[b]);
      }
    }
  }
  if (ts.box.nranks > 1) {
    // fused collective: sums in rank order, [-min, max] by maximum, straight over NVLink
    __shared__ int s_peer_ok;
    __syncthreads();
    peer_allreduce_stats(ts.box, ts.seq, ts.stats, ts.stats_glob, kNegMinSlot, kMaxSlot + 1, &s_peer_ok);
  }
  publish_to_host(ts);
}

}  // namespace dev

// One pass: (optional) two-level error refinement, relative-error flag, all iteration sums.
// REFINE=false is iteration 0 of the reference loop (no parents yet: hybrid.cuh:26-28 returns
// early, every region stays active and the raw errors are the errors).
//
// A block owns a SUPER-TILE of kSuper consecutive tiles: the per-tile active counts (what compaction
// needs) come from one __syncthreads_count per tile, the seven statistics are reduced ONCE per block
// and stored as one partial record per super-tile -- still a fixed shape that depends on n only, so
// the sums are run-to-run deterministic.  The reduction/arrival tail of a block, during which it has
// no loads in flight, is thereby paid once per 4096 regions instead of once per 1024.
constexpr int kSuper = 4;
__host__ __device__ inline unsigned int
super_tiles_for(unsigned int tiles)
{
  return (tiles + kSuper - 1) / kSuper;
}

// launch-bound / pipeline-depth knobs of k_classify (tools/gpu_sessions sweep them through alternative builds)
#ifndef PGN_CLASSIFY_MINBLOCKS
#define PGN_CLASSIFY_MINBLOCKS 3
#endif
#ifndef PGN_CLASSIFY_DEPTH
#define PGN_CLASSIFY_DEPTH 3
#endif
template <bool REFINE>
__global__ void __launch_bounds__(kBlock, PGN_CLASSIFY_MINBLOCKS)
k_classify(const double* __restrict__ est, const double* __restrict__ err_raw,
           const double* __restrict__ parent_est, double* __restrict__ err_out,
           uint8_t* __restrict__ active, size_t n, double epsrel, int relerr_classification,
           unsigned int tiles, TileScratch ts)
{
  __shared__ dev::Acc s_warp[kBlock / 32];
  __shared__ uint64_t s_scan[kBlock];
  const size_t half = n >> 1;
  // Blocks are dispatched in blockIdx order; consecutive blocks take a super-tile of the left half and the
  // one half a list further on, i.e. (nearly) each other's siblings, so the sibling reads of one are the
  // own reads of the other and hit L2 instead of going to HBM a second time (16.7 M regions: est/err are
  // 134 MB each).  Tile- and super-tile-indexed outputs do not depend on the processing order.
  unsigned int sup = blockIdx.x;
  if constexpr (REFINE) {
    const unsigned int first_right = (gridDim.x + 1u) >> 1;
    sup = (blockIdx.x & 1u) ? first_right + (blockIdx.x >> 1) : (blockIdx.x >> 1);
  }
  __shared__ unsigned int s_tile_active[kSuper];
  if (threadIdx.x < kSuper)
    s_tile_active[threadIdx.x] = 0u;
  __syncthreads();
  dev::Acc a;
  dev::acc_zero(a);
  // The 16 rounds of a super-tile (256 regions each) run as a software pipeline of depth kDepth: the five loads of
  // round r + kDepth are issued before the arithmetic of round r, so a thread always has 5 * kDepth loads in flight
  // while it computes (one correctly rounded FP64 division and seven accumulations per region: about as many FP64
  // pipe cycles as the 29 bytes cost in HBM time -- a load-everything-then-compute loop pays both in turn).
  constexpr int kSteps = kSuper * kRounds;
  constexpr int kDepth = PGN_CLASSIFY_DEPTH;
  const size_t sbase = static_cast<size_t>(sup) * kSuper * kTile;
  struct Stage {
    double est, err, sib_est, sib_err, par;
  };
  Stage st[kDepth];
  auto load = [&](int step, Stage& s) {
    size_t i = sbase + static_cast<size_t>(step) * kBlock + threadIdx.x;
    i = i < n ? i : n - 1;  // clamped: the loads are unconditional
    s.est = est[i];
    s.err = err_raw[i];
    if constexpr (REFINE) {
      const bool right = (i + i) >= n;
      const size_t sib = right ? i - half : i + half;
      const size_t par = right ? i - half : i;
      s.sib_est = est[sib];
      s.sib_err = err_raw[sib];
      s.par = parent_est[par];
    }
  };
#pragma unroll
  for (int d = 0; d < kDepth; ++d)
    load(d, st[d]);
  int my_active = 0;
#pragma unroll
  for (int step = 0; step < kSteps; ++step) {
    const Stage cur = st[step % kDepth];
    if (step + kDepth < kSteps)
      load(step + kDepth, st[step % kDepth]);
    const size_t i = sbase + static_cast<size_t>(step) * kBlock + threadIdx.x;
    if (i < n) {
      const double self_est = cur.est;
      double e = cur.err;
      double flag = 1.0;
      if constexpr (REFINE) {
        // Phases.cuh:251-261 with the rounding sequence of the reference's machine code
        const double diff = fabs(__dmul_rn(__dsub_rn(__dadd_rn(self_est, cur.sib_est), cur.par), 0.25));
        const double both = __dadd_rn(e, cur.sib_err);
        if (both > 0.0) {
          const double c = __dadd_rn(__ddiv_rn(__dadd_rn(diff, diff), both), 1.0);
          e = __dmul_rn(e, c);
        }
        e = __dadd_rn(diff, e);
        err_out[i] = e;
        if (relerr_classification) {
          const double tol = fmax(__dmul_rn(epsrel, fabs(self_est)), 1e-200);
          flag = (e < tol) ? 0.0 : 1.0;
        }
      }
      active[i] = flag != 0.0 ? 1 : 0;
      my_active += flag != 0.0 ? 1 : 0;
      a.v[kSumEst] += self_est;
      a.v[kSumErr] += e;
      a.v[kSumActEst] += flag * self_est;  // thrust::inner_product(flags, estimates)
      a.v[kSumActErr] += flag * e;
      a.v[kNumActive] += flag;
      a.mn = fmin(a.mn, e);
      a.mx = fmax(a.mx, e);
    }
    if ((step % kRounds) == kRounds - 1) {
      // active regions of the tile just finished: warp sums into a shared counter (integers: order does not matter)
      const int warp_active = __reduce_add_sync(0xffffffffu, my_active);
      if ((threadIdx.x & 31) == 0 && warp_active)
        atomicAdd(&s_tile_active[step / kRounds], static_cast<unsigned int>(warp_active));
      my_active = 0;
    }
  }
  __syncthreads();
  if (threadIdx.x < kSuper && sup * kSuper + threadIdx.x < tiles)
    ts.tile_count[sup * kSuper + threadIdx.x] = s_tile_active[threadIdx.x];
  a = dev::block_reduce(a, s_warp);
  if (threadIdx.x == 0) {
    double* p = ts.partials + static_cast<size_t>(sup) * kStatCount;
#pragma unroll
    for (int i = 0; i < 5; ++i)
      p[i] = a.v[i];
    p[kMinErr] = a.mn;
    p[kMaxErr] = a.mx;
  }
  dev::finalize(ts, tiles, true, s_warp, s_scan, gridDim.x);
}

// ---- classify, sibling-pair form with TMA-staged loads -------------------------------------------------------------
// The children of parent j sit at j (left half of the list) and j + n/2 (right half): the two refinements read the
// same five numbers -- both estimates, both raw errors, the parent estimate -- and share |est_l + est_r - parent| / 4,
// err_l + err_r and the correction factor (additions commute, so each twin obtains bit for bit what the one-region
// form computes).  A work item is therefore 1024 PAIRS: est/err of [b, b+1024) and of [b + n/2, b + n/2 + 1024) plus
// parent[b, b+1024) -- every input byte is fetched exactly once (20 B per region instead of 40 through L2) and a pair
// costs one correctly rounded division instead of two.
// The five 8 KB row segments of an item arrive by 1-D bulk copies (cp.async.bulk, the TMA engine) into one of two
// shared-memory stages, completion on an mbarrier: the copies of item k+1 (and k+2) are in flight while item k is
// computed, without holding a register, so the 128-register budget goes to the arithmetic and two resident blocks per
// SM keep ~160 KB of loads in flight (HBM needs ~45 KB per SM).  Persistent grid; one partial record per BLOCK (the
// item -> block map is static, so the sums stay run-to-run deterministic); per-tile active counts by integer atomics
// into tile_count, which the host zeroes before the launch (the right-hand rows straddle two tiles).
// A right-hand row starts at an odd element when n/2 is odd: the copy then starts one element earlier (16-byte
// aligned) and the consumers index one element further.  The last item, when shorter than 1024 pairs, is read with
// plain loads.
constexpr int kPairRow = kTile + 2;           // doubles per staged row (room for the alignment shift)
constexpr int kPairStage = 5 * kPairRow;      // est_l, err_l, parent, est_r, err_r
constexpr size_t kPairSmemBytes = sizeof(double) * 2 * kPairStage + 2 * sizeof(unsigned long long);
constexpr size_t kPairMinRegions = 128 * 1024;  // below this the one-region kernel is used

namespace dev {
__device__ __forceinline__ uint32_t
smem_addr(const void* p)
{
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void
mbar_init(unsigned long long* bar, unsigned int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void
mbar_expect_tx(unsigned long long* bar, unsigned int bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void
mbar_wait(unsigned long long* bar, unsigned int parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "PGN_WAIT:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra PGN_DONE;\n"
    "bra PGN_WAIT;\n"
    "PGN_DONE:\n"
    "}\n" ::"r"(smem_addr(bar)),
    "r"(parity)
    : "memory");
}
// global -> shared bulk copy (bytes and both addresses multiples of 16), completion counted on `bar`
__device__ __forceinline__ void
bulk_load(void* dst, const void* src, unsigned int bytes, unsigned long long* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smem_addr(dst)),
               "l"(src), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
}  // namespace dev

__global__ void __launch_bounds__(kBlock, 2)
k_classify_pairs(const double* __restrict__ est, const double* __restrict__ err_raw,
                 const double* __restrict__ parent_est, double* __restrict__ err_out,
                 uint8_t* __restrict__ active, size_t n, double epsrel, int relerr_classification,
                 unsigned int tiles, TileScratch ts)
{
  extern __shared__ __align__(128) unsigned char pair_smem[];
  double* stage = reinterpret_cast<double*>(pair_smem);
  unsigned long long* full = reinterpret_cast<unsigned long long*>(pair_smem + sizeof(double) * 2 * kPairStage);
  __shared__ dev::Acc s_warp[kBlock / 32];
  __shared__ uint64_t s_scan[kBlock];

  const size_t half = n >> 1;  // n is even (host-checked)
  const unsigned int items = static_cast<unsigned int>((half + kTile - 1) / kTile);
  const unsigned int shift = static_cast<unsigned int>(half & 1);  // right-hand rows start `shift` elements into their stage row
  auto item_len = [&](unsigned int it) -> unsigned int {
    const size_t b = static_cast<size_t>(it) * kTile;
    return static_cast<unsigned int>(half - b < kTile ? half - b : kTile);
  };
  // thread 0: arm the stage's barrier and start the five copies of item `it` (only whole items travel this way)
  auto issue = [&](unsigned int it, int s) {
    const size_t b = static_cast<size_t>(it) * kTile;
    double* dst = stage + s * kPairStage;
    const unsigned int row = sizeof(double) * kTile, rrow = sizeof(double) * (kTile + 2 * shift);
    dev::mbar_expect_tx(&full[s], 3 * row + 2 * rrow);
    dev::bulk_load(dst, est + b, row, &full[s]);
    dev::bulk_load(dst + kPairRow, err_raw + b, row, &full[s]);
    dev::bulk_load(dst + 2 * kPairRow, parent_est + b, row, &full[s]);
    dev::bulk_load(dst + 3 * kPairRow, est + b + half - shift, rrow, &full[s]);
    dev::bulk_load(dst + 4 * kPairRow, err_raw + b + half - shift, rrow, &full[s]);
  };
  if (threadIdx.x == 0) {
    dev::mbar_init(&full[0], 1);
    dev::mbar_init(&full[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int i0 = blockIdx.x, i1 = blockIdx.x + gridDim.x;
    if (i0 < items && item_len(i0) == kTile)
      issue(i0, 0);
    if (i1 < items && item_len(i1) == kTile)
      issue(i1, 1);
  }

  dev::Acc a;
  dev::acc_zero(a);
  int n_active = 0;
  const int lane = threadIdx.x & 31;
  unsigned int k = 0;
  for (unsigned int it = blockIdx.x; it < items; it += gridDim.x, ++k) {
    const int s = static_cast<int>(k & 1u);
    const size_t b = static_cast<size_t>(it) * kTile;
    const unsigned int len = item_len(it);
    const bool staged = len == kTile;
    const double* row = stage + s * kPairStage;
    if (staged)
      dev::mbar_wait(&full[s], (k >> 1) & 1u);
    int cnt_l = 0, cnt_r0 = 0, cnt_r1 = 0;
    const size_t r_base = b + half;
    const size_t r_split = ((r_base >> 10) + 1) << 10;  // first region of the right-hand row's second tile
#pragma unroll
    for (int q = 0; q < kRounds; ++q) {
      const unsigned int j = q * kBlock + threadIdx.x;
      if (j < len) {
        double est_l, err_l, par, est_r, err_r;
        if (staged) {
          est_l = row[j];
          err_l = row[kPairRow + j];
          par = row[2 * kPairRow + j];
          est_r = row[3 * kPairRow + shift + j];
          err_r = row[4 * kPairRow + shift + j];
        } else {
          est_l = est[b + j];
          err_l = err_raw[b + j];
          par = parent_est[b + j];
          est_r = est[r_base + j];
          err_r = err_raw[r_base + j];
        }
        // Phases.cuh:251-261 with the rounding sequence of the reference's machine code (see k_classify)
        const double diff = fabs(__dmul_rn(__dsub_rn(__dadd_rn(est_l, est_r), par), 0.25));
        const double both = __dadd_rn(err_l, err_r);
        double e_l = err_l, e_r = err_r;
        if (both > 0.0) {
          const double c = __dadd_rn(__ddiv_rn(__dadd_rn(diff, diff), both), 1.0);
          e_l = __dmul_rn(e_l, c);
          e_r = __dmul_rn(e_r, c);
        }
        e_l = __dadd_rn(diff, e_l);
        e_r = __dadd_rn(diff, e_r);
        err_out[b + j] = e_l;
        err_out[r_base + j] = e_r;
        double flag_l = 1.0, flag_r = 1.0;
        if (relerr_classification) {
          flag_l = (e_l < fmax(__dmul_rn(epsrel, fabs(est_l)), 1e-200)) ? 0.0 : 1.0;
          flag_r = (e_r < fmax(__dmul_rn(epsrel, fabs(est_r)), 1e-200)) ? 0.0 : 1.0;
        }
        active[b + j] = flag_l != 0.0 ? 1 : 0;
        active[r_base + j] = flag_r != 0.0 ? 1 : 0;
        cnt_l += flag_l != 0.0 ? 1 : 0;
        if (r_base + j < r_split)
          cnt_r0 += flag_r != 0.0 ? 1 : 0;
        else
          cnt_r1 += flag_r != 0.0 ? 1 : 0;
        a.v[kSumEst] += est_l;
        a.v[kSumErr] += e_l;
        a.v[kSumActEst] += flag_l * est_l;
        a.v[kSumActErr] += flag_l * e_l;
        a.mn = fmin(a.mn, e_l);
        a.mx = fmax(a.mx, e_l);
        a.v[kSumEst] += est_r;
        a.v[kSumErr] += e_r;
        a.v[kSumActEst] += flag_r * est_r;
        a.v[kSumActErr] += flag_r * e_r;
        a.mn = fmin(a.mn, e_r);
        a.mx = fmax(a.mx, e_r);
      }
    }
    n_active += cnt_l + cnt_r0 + cnt_r1;
    // active regions per tile of the list (what compaction needs): integer atomics, the order does not matter
    cnt_l = __reduce_add_sync(0xffffffffu, cnt_l);
    cnt_r0 = __reduce_add_sync(0xffffffffu, cnt_r0);
    cnt_r1 = __reduce_add_sync(0xffffffffu, cnt_r1);
    if (lane == 0) {
      if (cnt_l)
        atomicAdd(&ts.tile_count[b >> 10], static_cast<unsigned int>(cnt_l));
      if (cnt_r0)
        atomicAdd(&ts.tile_count[r_base >> 10], static_cast<unsigned int>(cnt_r0));
      if (cnt_r1)
        atomicAdd(&ts.tile_count[(r_base >> 10) + 1], static_cast<unsigned int>(cnt_r1));
    }
    __syncthreads();  // every thread has read stage s
    if (threadIdx.x == 0) {
      const unsigned int nxt = it + 2 * gridDim.x;
      if (nxt < items && item_len(nxt) == kTile)
        issue(nxt, s);
    }
  }
  a.v[kNumActive] = static_cast<double>(n_active);
  a = dev::block_reduce(a, s_warp);
  if (threadIdx.x == 0) {
    double* p = ts.partials + static_cast<size_t>(blockIdx.x) * kStatCount;
#pragma unroll
    for (int i = 0; i < 5; ++i)
      p[i] = a.v[i];
    p[kMinErr] = a.mn;
    p[kMaxErr] = a.mx;
  }
  dev::finalize(ts, tiles, true, s_warp, s_scan, gridDim.x);
}

// One probe of the heuristic threshold search: flag = err > threshold.
// MODE 0: count + sum(flag*err) only (nothing written);  MODE 1: also write the flags, the tile
// counts/prefix and sum(flag*est) (the accepted threshold);  MODE 2: re-activate everything
// (Workspace::fix_error_budget_overflow, Workspace.cuh:135-145).
template <int MODE>
__global__ void __launch_bounds__(kBlock)
k_threshold(const double* __restrict__ est, const double* __restrict__ err, uint8_t* __restrict__ active,
            size_t n, double threshold, TileScratch ts)
{
  __shared__ dev::Acc s_warp[kBlock / 32];
  __shared__ uint64_t s_scan[kBlock];
  const size_t base = static_cast<size_t>(blockIdx.x) * kTile;
  dev::Acc a;
  dev::acc_zero(a);
#pragma unroll
  for (int q = 0; q < kRounds; ++q) {
    const size_t i = base + static_cast<size_t>(q) * kBlock + threadIdx.x;
    if (i < n) {
      const double e = err[i];
      const double flag = (MODE == 2) ? 1.0 : ((e > threshold) ? 1.0 : 0.0);
      a.v[kSumActErr] += flag * e;
      a.v[kNumActive] += flag;
      a.mn = fmin(a.mn, e);  // the threshold search reads min/max of the errors after a MODE 2 pass
      a.mx = fmax(a.mx, e);
      if (MODE != 0) {
        const double v = est[i];
        a.v[kSumActEst] += flag * v;
        active[i] = flag != 0.0 ? 1 : 0;
      }
    }
  }
  a = dev::block_reduce(a, s_warp);
  if (threadIdx.x == 0) {
    double* p = ts.partials + static_cast<size_t>(blockIdx.x) * kStatCount;
#pragma unroll
    for (int i = 0; i < 5; ++i)
      p[i] = a.v[i];
    p[kMinErr] = a.mn;
    p[kMaxErr] = a.mx;
    if (MODE != 0)
      ts.tile_count[blockIdx.x] = static_cast<uint32_t>(a.v[kNumActive]);
  }
  dev::finalize(ts, gridDim.x, MODE != 0, s_warp, s_scan);
}

// Recount after the flags were overwritten from outside (pgn_set_active).
__global__ void __launch_bounds__(kBlock)
k_recount(const double* __restrict__ est, const double* __restrict__ err,
          const uint8_t* __restrict__ active, size_t n, TileScratch ts)
{
  __shared__ dev::Acc s_warp[kBlock / 32];
  __shared__ uint64_t s_scan[kBlock];
  const size_t base = static_cast<size_t>(blockIdx.x) * kTile;
  dev::Acc a;
  dev::acc_zero(a);
#pragma unroll
  for (int q = 0; q < kRounds; ++q) {
    const size_t i = base + static_cast<size_t>(q) * kBlock + threadIdx.x;
    if (i < n) {
      const double flag = active[i] ? 1.0 : 0.0;
      const double e = err[i], v = est[i];
      a.v[kSumEst] += v;
      a.v[kSumErr] += e;
      a.v[kSumActEst] += flag * v;
      a.v[kSumActErr] += flag * e;
      a.v[kNumActive] += flag;
      a.mn = fmin(a.mn, e);
      a.mx = fmax(a.mx, e);
    }
  }
  a = dev::block_reduce(a, s_warp);
  if (threadIdx.x == 0) {
    double* p = ts.partials + static_cast<size_t>(blockIdx.x) * kStatCount;
#pragma unroll
    for (int i = 0; i < 5; ++i)
      p[i] = a.v[i];
    p[kMinErr] = a.mn;
    p[kMaxErr] = a.mx;
    ts.tile_count[blockIdx.x] = static_cast<uint32_t>(a.v[kNumActive]);
  }
  dev::finalize(ts, gridDim.x, true, s_warp, s_scan);
}

// A rank whose region list is empty still takes part in the fused all-reduce.
__global__ void __launch_bounds__(kBlock)
k_publish_empty_stats(TileScratch ts)
{
  __shared__ int s_peer_ok;
  if (threadIdx.x < kStatsLen) {
    double v = 0.0;
    if (threadIdx.x == kMinErr)
      v = __longlong_as_double(0x7ff0000000000000LL);
    if (threadIdx.x == kMaxErr || threadIdx.x == kNegMinSlot || threadIdx.x == kMaxSlot)
      v = __longlong_as_double(0xfff0000000000000LL);
    ts.stats[threadIdx.x] = v;
  }
  __syncthreads();
  if (ts.box.nranks > 1)
    peer_allreduce_stats(ts.box, ts.seq, ts.stats, ts.stats_glob, kNegMinSlot, kMaxSlot + 1, &s_peer_ok);
  dev::publish_to_host(ts);
}

#ifndef PGN_COMPACT_AXES
#define PGN_COMPACT_AXES 8
#endif
struct CompactArgs {
  const double* left;       // [ndim][n]
  const double* length;     // [ndim][n]
  const double* est;        // [n]
  const double* err;        // [n]   (refined)
  const int32_t* axis;      // [n]
  const uint8_t* active;    // [n]
  const uint64_t* tile_offset;
  size_t n;
  size_t m;                 // number of active regions (known on the host before the launch)
  double* out_left;         // FUSED: [ndim][2m]   else [ndim][m]
  double* out_length;
  double* parent_est;       // [m]
  double* parent_err;       // [m]
  int32_t* out_axis;        // [m] (only when not fused)
  int ndim;
};

// Stable stream compaction of the active regions; FUSED additionally bisects each survivor along
// its split axis and writes both children at the reference's indices j (left half) and m + j
// (right half) of the next region list, so every surviving parent is read once and its children
// written once.
template <bool FUSED>
__global__ void __launch_bounds__(kBlock)  // (capping at 64 registers for a fourth resident block measured 2-4 % slower: session N)
k_compact(const CompactArgs A)
{
  __shared__ uint32_t s_wsum[kRounds][kBlock / 32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const size_t base = static_cast<size_t>(blockIdx.x) * kTile;
  const size_t out_stride = FUSED ? 2 * A.m : A.m;
  // all four rounds' flags first: one barrier per tile, and the loads of different rounds overlap
  bool on[kRounds];
  unsigned int ballot[kRounds];
#pragma unroll
  for (int q = 0; q < kRounds; ++q) {
    const size_t i = base + static_cast<size_t>(q) * kBlock + threadIdx.x;
    on[q] = i < A.n && A.active[i] != 0;
    ballot[q] = __ballot_sync(0xffffffffu, on[q]);
    if (lane == 0)
      s_wsum[q][w] = __popc(ballot[q]);
  }
  __syncthreads();
  uint64_t run = A.tile_offset[blockIdx.x];
  // A tile in which at least a quarter of the regions survive is read in FULL, dropped regions included: with
  // survivors scattered at random nearly every 32-byte sector holds one and reaches the SM anyway (at 48 % survival:
  // 2.58 GB read for 1.19 GB of survivors, ncu), while predicated-off lanes only take loads out of flight (4.7 TB/s of
  // DRAM traffic at 48 % survival against 6.3 TB/s when every region survives).  Sparse tiles keep predicated loads.
  uint32_t tile_survivors = 0;
#pragma unroll
  for (int q = 0; q < kRounds; ++q)
#pragma unroll
    for (int k = 0; k < kBlock / 32; ++k)
      tile_survivors += s_wsum[q][k];
  const bool dense = 4u * tile_survivors >= static_cast<uint32_t>(kTile);  // block-uniform
#pragma unroll
  for (int q = 0; q < kRounds; ++q) {
    uint32_t before = 0, total = 0;
#pragma unroll
    for (int k = 0; k < kBlock / 32; ++k) {
      const uint32_t c = s_wsum[q][k];
      before += k < w ? c : 0;
      total += c;
    }
    const size_t i0 = base + static_cast<size_t>(q) * kBlock + threadIdx.x;
    const bool fetch = dense ? i0 < A.n : on[q];
    if (fetch) {
      const size_t i = i0;
      const size_t j = run + before + __popc(ballot[q] & ((1u << lane) - 1u));
      const int ax = A.axis[i];
      const double pe = A.est[i], pr = A.err[i];
      if (on[q]) {
        A.parent_est[j] = pe;
        A.parent_err[j] = pr;
        if (!FUSED)
          A.out_axis[j] = ax;
      }
      // kAxes axes at a time: 2 * kAxes independent loads in flight per thread before the first store (the output
      // pointers may alias the inputs as far as the compiler knows).
      constexpr int kAxes = PGN_COMPACT_AXES;
      for (int d0 = 0; d0 < A.ndim; d0 += kAxes) {
        double lo[kAxes], ln[kAxes];
#pragma unroll
        for (int k = 0; k < kAxes; ++k) {
          const int d = d0 + k < A.ndim ? d0 + k : A.ndim - 1;
          lo[k] = __ldcs(&A.left[static_cast<size_t>(d) * A.n + i]);
          ln[k] = __ldcs(&A.length[static_cast<size_t>(d) * A.n + i]);
        }
        if (on[q]) {
#pragma unroll
          for (int k = 0; k < kAxes; ++k) {
            const int d = d0 + k;
            if (d < A.ndim) {
              double* ol = A.out_left + static_cast<size_t>(d) * out_stride;
              double* ow = A.out_length + static_cast<size_t>(d) * out_stride;
              if (FUSED) {
                const bool cut = d == ax;
                const double h = cut ? __ddiv_rn(ln[k], 2.0) : ln[k];  // Sub_region_splitter.cuh:33-35
                ol[j] = lo[k];
                ol[A.m + j] = cut ? __dadd_rn(lo[k], h) : lo[k];      // left + 1 * interval_length
                ow[j] = h;
                ow[A.m + j] = h;
              } else {
                ol[j] = lo[k];
                ow[j] = ln[k];
              }
            }
          }
        }
      }
    }
    run += total;
  }
}

// Bisection alone (stage-level API and the rebalancing iterations of the multi-GPU path).
__global__ void __launch_bounds__(kBlock)
k_split(const double* __restrict__ left, const double* __restrict__ length,
        const int32_t* __restrict__ axis, size_t m, int ndim, double* __restrict__ out_left,
        double* __restrict__ out_length)
{
  const size_t j = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (j >= m)
    return;
  const int ax = axis[j];
  for (int d = 0; d < ndim; ++d) {
    const double lo = left[static_cast<size_t>(d) * m + j];
    const double ln = length[static_cast<size_t>(d) * m + j];
    double* ol = out_left + static_cast<size_t>(d) * 2 * m;
    double* ow = out_length + static_cast<size_t>(d) * 2 * m;
    if (d == ax) {
      const double h = __ddiv_rn(ln, 2.0);
      ol[j] = lo;
      ol[m + j] = __dadd_rn(lo, h);
      ow[j] = h;
      ow[m + j] = h;
    } else {
      ol[j] = lo;
      ol[m + j] = lo;
      ow[j] = ln;
      ow[m + j] = ln;
    }
  }
}

// digest64(x) = sum_i bits(x[i]) * (2i+1)  (mod 2^64); modular sums commute, so atomics are exact.
template <typename T>
__global__ void
k_digest(const T* __restrict__ x, size_t n, unsigned long long* out)
{
  unsigned long long acc = 0;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    unsigned long long bits;
    if (sizeof(T) == 8)
      bits = *reinterpret_cast<const unsigned long long*>(&x[i]);
    else if (sizeof(T) == 4)
      bits = *reinterpret_cast<const unsigned int*>(&x[i]);
    else
      bits = *reinterpret_cast<const unsigned char*>(&x[i]);
    acc += bits * (2ull * i + 1ull);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
    acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if ((threadIdx.x & 31) == 0)
    atomicAdd(out, acc);
}

}  // namespace pgn
#endif
