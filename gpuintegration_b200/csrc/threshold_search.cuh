// Device-side threshold search of the heuristic (memory-pressure) classifier.
//
// Reference: Heuristic_classifier::classify (cuda/pagani/quad/GPUquad/heuristic_classifier.cuh:362-451)
// bisects an error threshold on the HOST; every probe is a flag kernel over all regions
// (device_set_true_for_larger_than, :60-72), a thrust::reduce (count) and a thrust::inner_product
// (sum of the errors above the threshold), each with a synchronous copy back -- up to several hundred
// probes per classification on a 16 M-region list.
//
// Here the error array is partitioned ONCE into value buckets (the bit pattern of a non-negative double
// is monotone in its value, so a bucket is a range of leading bits), and a probe becomes
//     count(err > t) = count of all higher buckets  +  a scan of the one bucket that contains t
// which a single thread block answers in a few microseconds.  The whole bisection (the reference's exact
// threshold sequence and decisions) then runs inside ONE kernel, k_search, with no host round trip.
// Several ranks: the bucket counts and integer sums are all-reduced once per search (k_tables_to_double ->
// communicator -> k_global_tables), so every rank holds the tables of the CONCATENATED list and walks through the
// same bisection; only probes those tables cannot decide are scanned and exchanged over the NVLink mailboxes
// (peer.cuh), with results bit for bit those of a single rank.  Without peer access the host drives the search
// with one k_query + all-reduce per probe.
//
// Sums are order independent by construction: inside a bucket every element has the same exponent, so
// the mantissas are added as integers (exactly), and bucket totals are combined in a fixed order.
// That makes the probe results -- and therefore the region trajectory -- run-to-run deterministic even
// though the scatter uses atomics.
#ifndef PGN_THRESHOLD_SEARCH_CUH
#define PGN_THRESHOLD_SEARCH_CUH

#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstring>

#include "peer.cuh"

namespace pgn {

constexpr int kSearchBlock = 1024;
constexpr int kMaxBuckets = 2048;
constexpr int kScatterPerThread = 8;
constexpr int kScatterChunk = kSearchBlock * kScatterPerThread;
constexpr int kProbeCache = 1024;  // memo table of probe results (power of two)
constexpr int kSearchRow = kMaxBuckets + 2;                             // one staged table of k_search (8-byte entries)

struct BucketPlan {
  int shift;     // key = bits >> shift
  int mbits;     // mantissa bits in the key (52 - shift)
  uint32_t klo;  // key of bucket 0; smaller keys are clamped into bucket 0
  uint32_t nb;   // buckets in use (<= kMaxBuckets)
};

// Device tables of one partition (all sized for kMaxBuckets).
struct BucketTables {
  unsigned long long* count;    // [nb]      elements per bucket
  unsigned long long* start;    // [nb + 1]  exclusive prefix of count
  unsigned long long* cursor;   // [nb]      scatter cursors
  double* bucket_sum;           // [nb]      exact sum of the bucket, rounded once
  double* suffix_sum;           // [nb + 1]  sum of buckets b..nb-1, accumulated from the top down
  unsigned long long* suffix_count;  // [nb + 1]
  unsigned long long* outside;  // [1] elements that are not finite non-negative numbers <= max
  unsigned long long* acc_lo;   // [nb] integer bucket sums (low / high 32-bit pieces), see k_bucket_sums
  unsigned long long* acc_hi;   // [nb]
  double* sorted;               // [n] the errors grouped by bucket
};

__host__ __device__ inline uint32_t
bucket_of(unsigned long long bits, const BucketPlan& p)
{
  const unsigned long long k = bits >> p.shift;  // shift >= 40: fits 24 bits
  const uint32_t kk = static_cast<uint32_t>(k);
  return kk < p.klo ? 0u : kk - p.klo;  // >= nb: not a finite non-negative value <= max
}

// Biased exponent (clamped to >= 1, i.e. subnormals use the scale of the smallest normal) of the
// upper edge of bucket b: the common scale of the bucket's integer sum.
__host__ __device__ inline int
bucket_exponent(uint32_t b, const BucketPlan& p)
{
  const int e = static_cast<int>(((p.klo + b) >> p.mbits) & 0x7ffu);
  return e < 1 ? 1 : e;
}

// value = mant * 2^(eref - 1075) (truncated when the element's own exponent is smaller, which only
// happens in the clamped bucket 0)
__device__ __forceinline__ unsigned long long
scaled_mantissa(unsigned long long bits, int eref)
{
  const int e = static_cast<int>((bits >> 52) & 0x7ffu);
  const unsigned long long m = bits & 0xfffffffffffffull;
  const unsigned long long mant = e ? (m | (1ull << 52)) : m;
  const int s = eref - (e < 1 ? 1 : e);
  return s <= 0 ? mant : (s < 64 ? mant >> s : 0ull);
}

// sum = (hi * 2^32 + lo) * 2^(eref - 1075)
__device__ __forceinline__ double
integer_sum_to_double(unsigned long long lo, unsigned long long hi, int eref)
{
  // lo collects 32-bit pieces and passes 2^53 on long lists: carry its upper half into hi in integers first, so that
  // both conversions are exact (hi < 2^46 for 2^24 elements, lo < 2^32) and the FMA is the only rounding
  hi += lo >> 32;
  lo &= 0xffffffffull;
  const double v = __fma_rn(static_cast<double>(hi), 4294967296.0, static_cast<double>(lo));
  return scalbn(v, eref - 1075);
}

// Host: choose the finest bucket width that covers [min_err, max_err] with <= kMaxBuckets buckets.
inline bool
make_bucket_plan(double min_err, double max_err, BucketPlan* p)
{
  if (!(min_err >= 0.0) || !(max_err >= min_err) || !std::isfinite(max_err))
    return false;
  unsigned long long lo, hi;
  std::memcpy(&lo, &min_err, 8);
  std::memcpy(&hi, &max_err, 8);
  lo &= 0x7fffffffffffffffull;  // -0.0
  for (int mb = 12; mb >= 0; --mb) {
    const int shift = 52 - mb;
    const unsigned long long span = (hi >> shift) - (lo >> shift) + 1;
    if (span <= static_cast<unsigned long long>(kMaxBuckets)) {
      p->shift = shift;
      p->mbits = mb;
      p->klo = static_cast<uint32_t>(lo >> shift);
      p->nb = static_cast<uint32_t>(span);
      return true;
    }
  }
  // more than 2048 binades cannot happen (11 exponent bits); more than kMaxBuckets at mb = 0 can:
  // keep the top kMaxBuckets binades and clamp everything smaller into bucket 0
  p->shift = 52;
  p->mbits = 0;
  const uint32_t kmax = static_cast<uint32_t>(hi >> 52);
  p->klo = kmax + 1 > static_cast<uint32_t>(kMaxBuckets) ? kmax + 1 - kMaxBuckets : 0u;
  p->nb = kmax - p->klo + 1;
  return true;
}

// ---- partition ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSearchBlock)
k_bucket_count(const double* __restrict__ err, size_t n, BucketPlan plan, BucketTables T)
{
  __shared__ unsigned int s_hist[kMaxBuckets];
  __shared__ unsigned int s_outside;
  for (int k = threadIdx.x; k < kMaxBuckets; k += kSearchBlock)
    s_hist[k] = 0u;
  if (threadIdx.x == 0)
    s_outside = 0u;
  __syncthreads();
  for (size_t i = static_cast<size_t>(blockIdx.x) * kSearchBlock + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * kSearchBlock) {
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(err[i]));
    const uint32_t b = bucket_of(bits, plan);
    if (b < plan.nb)
      atomicAdd(&s_hist[b], 1u);
    else
      atomicAdd(&s_outside, 1u);
  }
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < plan.nb; k += kSearchBlock) {
    const unsigned int c = s_hist[k];
    if (c)
      atomicAdd(&T.count[k], static_cast<unsigned long long>(c));
  }
  if (threadIdx.x == 0 && s_outside)
    atomicAdd(T.outside, static_cast<unsigned long long>(s_outside));
}

// one block: exclusive prefix of the bucket counts, cursors
__device__ __forceinline__ void
bucket_offsets_body(const BucketPlan& plan, const BucketTables& T, unsigned long long* s_scan /* [kSearchBlock] */)
{
  // kMaxBuckets = 2 * kSearchBlock: two buckets per thread
  const uint32_t b0 = 2u * threadIdx.x, b1 = b0 + 1u;
  const unsigned long long c0 = b0 < plan.nb ? T.count[b0] : 0ull;
  const unsigned long long c1 = b1 < plan.nb ? T.count[b1] : 0ull;
  s_scan[threadIdx.x] = c0 + c1;
  __syncthreads();
  for (int off = 1; off < kSearchBlock; off <<= 1) {
    const unsigned long long add = static_cast<int>(threadIdx.x) >= off ? s_scan[threadIdx.x - off] : 0ull;
    __syncthreads();
    s_scan[threadIdx.x] += add;
    __syncthreads();
  }
  const unsigned long long before = s_scan[threadIdx.x] - (c0 + c1);
  if (b0 < plan.nb) {
    T.start[b0] = before;
    T.cursor[b0] = before;
  }
  if (b1 < plan.nb) {
    T.start[b1] = before + c0;
    T.cursor[b1] = before + c0;
  }
  if (threadIdx.x == kSearchBlock - 1)
    T.start[plan.nb] = s_scan[threadIdx.x];
}

__global__ void __launch_bounds__(kSearchBlock)
k_bucket_offsets(BucketPlan plan, BucketTables T)
{
  __shared__ unsigned long long s_scan[kSearchBlock];
  bucket_offsets_body(plan, T, s_scan);
}

// Scatter into bucket order.  A block handles chunks of 8192 elements: ranks inside the chunk come
// from shared-memory atomics, one global atomic per (chunk, populated bucket) reserves the range.
__global__ void __launch_bounds__(kSearchBlock)
k_bucket_scatter(const double* __restrict__ err, size_t n, BucketPlan plan, BucketTables T)
{
  __shared__ unsigned int s_hist[kMaxBuckets];
  __shared__ unsigned long long s_base[kMaxBuckets];
  for (int k = threadIdx.x; k < kMaxBuckets; k += kSearchBlock)
    s_hist[k] = 0u;
  __syncthreads();
  const size_t chunks = (n + kScatterChunk - 1) / kScatterChunk;
  for (size_t c = blockIdx.x; c < chunks; c += gridDim.x) {
    const size_t base = c * kScatterChunk;
    double v[kScatterPerThread];
    uint32_t b[kScatterPerThread];
    unsigned int r[kScatterPerThread];
#pragma unroll
    for (int q = 0; q < kScatterPerThread; ++q) {
      const size_t i = base + static_cast<size_t>(q) * kSearchBlock + threadIdx.x;
      b[q] = 0xffffffffu;
      if (i < n) {
        v[q] = err[i];
        const uint32_t bb = bucket_of(static_cast<unsigned long long>(__double_as_longlong(v[q])), plan);
        if (bb < plan.nb) {
          b[q] = bb;
          r[q] = atomicAdd(&s_hist[bb], 1u);
        }
      }
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < plan.nb; k += kSearchBlock) {
      const unsigned int cnt = s_hist[k];
      if (cnt) {
        s_base[k] = atomicAdd(&T.cursor[k], static_cast<unsigned long long>(cnt));
        s_hist[k] = 0u;
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kScatterPerThread; ++q)
      if (b[q] != 0xffffffffu)
        T.sorted[s_base[b[q]] + r[q]] = v[q];
    __syncthreads();
  }
}

// block-wide exact reduction of (count, lo, hi); result valid in every thread
__device__ __forceinline__ void
block_sum3(unsigned long long& a, unsigned long long& b, unsigned long long& c,
           unsigned long long (*s_red)[3])
{
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
    c += __shfl_xor_sync(0xffffffffu, c, off);
  }
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  __syncthreads();  // protect s_red from the previous use
  if (lane == 0) {
    s_red[w][0] = a;
    s_red[w][1] = b;
    s_red[w][2] = c;
  }
  __syncthreads();
  const int nw = blockDim.x >> 5;
  a = lane < nw ? s_red[lane][0] : 0ull;
  b = lane < nw ? s_red[lane][1] : 0ull;
  c = lane < nw ? s_red[lane][2] : 0ull;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, off);
    b += __shfl_xor_sync(0xffffffffu, b, off);
    c += __shfl_xor_sync(0xffffffffu, c, off);
  }
}

// Exact integer sums of the buckets, parallel over ELEMENTS: the error distributions PAGANI produces put most
// of a list into a handful of buckets, so one block per bucket leaves the device idle.  A block takes chunks of
// kSumChunk consecutive elements of the bucket-sorted array; a chunk that lies inside one bucket (nearly all of
// them) is reduced by the block and added with one pair of integer atomics, a chunk that straddles bucket
// boundaries adds per-thread runs.  Integer additions commute: the totals do not depend on the order.
constexpr int kSumThreads = 256;
constexpr int kSumPerThread = 16;
constexpr int kSumChunk = kSumThreads * kSumPerThread;

__global__ void __launch_bounds__(kSumThreads)
k_bucket_sums(BucketPlan plan, BucketTables T)
{
  __shared__ unsigned long long s_red[32][3];
  const unsigned long long total = T.start[plan.nb];
  const unsigned long long chunks = (total + kSumChunk - 1) / kSumChunk;
  for (unsigned long long c = blockIdx.x; c < chunks; c += gridDim.x) {
    const unsigned long long base = c * kSumChunk;
    const unsigned long long end = base + kSumChunk < total ? base + kSumChunk : total;
    const uint32_t b_first = bucket_of(static_cast<unsigned long long>(__double_as_longlong(T.sorted[base])), plan);
    const uint32_t b_last = bucket_of(static_cast<unsigned long long>(__double_as_longlong(T.sorted[end - 1])), plan);
    if (b_first == b_last) {  // block-uniform
      const int eref = bucket_exponent(b_first, plan);
      unsigned long long cnt = 0, lo = 0, hi = 0;
#pragma unroll
      for (int q = 0; q < kSumPerThread; ++q) {
        const unsigned long long i = base + static_cast<unsigned long long>(q) * kSumThreads + threadIdx.x;
        if (i < end) {
          const unsigned long long m =
            scaled_mantissa(static_cast<unsigned long long>(__double_as_longlong(T.sorted[i])), eref);
          lo += m & 0xffffffffull;
          hi += m >> 32;
        }
      }
      block_sum3(cnt, lo, hi, s_red);
      if (threadIdx.x == 0) {
        atomicAdd(&T.acc_lo[b_first], lo);
        atomicAdd(&T.acc_hi[b_first], hi);
      }
    } else {
      // thread t owns elements base + t*16 .. +15: runs of equal bucket, one pair of atomics per run
      const unsigned long long i0 = base + static_cast<unsigned long long>(threadIdx.x) * kSumPerThread;
      uint32_t run_b = 0xffffffffu;
      int eref = 1;
      unsigned long long lo = 0, hi = 0;
      for (int q = 0; q < kSumPerThread; ++q) {
        const unsigned long long i = i0 + q;
        if (i >= end)
          break;
        const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(T.sorted[i]));
        const uint32_t bb = bucket_of(bits, plan);
        if (bb != run_b) {
          if (run_b != 0xffffffffu) {
            atomicAdd(&T.acc_lo[run_b], lo);
            atomicAdd(&T.acc_hi[run_b], hi);
          }
          run_b = bb;
          eref = bucket_exponent(bb, plan);
          lo = hi = 0;
        }
        const unsigned long long m = scaled_mantissa(bits, eref);
        lo += m & 0xffffffffull;
        hi += m >> 32;
      }
      if (run_b != 0xffffffffu) {
        atomicAdd(&T.acc_lo[run_b], lo);
        atomicAdd(&T.acc_hi[run_b], hi);
      }
    }
  }
}

// one block: suffix sums / counts, accumulated from the top bucket down in a fixed order
__device__ __forceinline__ void
bucket_suffix_body(const BucketPlan& plan, const BucketTables& T, double* s_sum /* [kSearchBlock] */)
{
  // thread t owns buckets nb-1-2t and nb-2-2t (descending): local sums, then an inclusive scan
  const int nb = static_cast<int>(plan.nb);
  const int k0 = nb - 1 - 2 * static_cast<int>(threadIdx.x), k1 = k0 - 1;
  // the integer totals of k_bucket_sums, rounded once each
  const double v0 = k0 >= 0 ? integer_sum_to_double(T.acc_lo[k0], T.acc_hi[k0], bucket_exponent(k0, plan)) : 0.0;
  const double v1 = k1 >= 0 ? integer_sum_to_double(T.acc_lo[k1], T.acc_hi[k1], bucket_exponent(k1, plan)) : 0.0;
  if (k0 >= 0)
    T.bucket_sum[k0] = v0;
  if (k1 >= 0)
    T.bucket_sum[k1] = v1;
  s_sum[threadIdx.x] = v0 + v1;
  __syncthreads();
  for (int off = 1; off < kSearchBlock; off <<= 1) {
    const double add = static_cast<int>(threadIdx.x) >= off ? s_sum[threadIdx.x - off] : 0.0;
    __syncthreads();
    s_sum[threadIdx.x] += add;
    __syncthreads();
  }
  const double before = threadIdx.x > 0 ? s_sum[threadIdx.x - 1] : 0.0;  // all buckets above k0
  const unsigned long long total = T.start[nb];
  if (k0 >= 0) {
    T.suffix_sum[k0] = before + v0;
    T.suffix_count[k0] = total - T.start[k0];
  }
  if (k1 >= 0) {
    T.suffix_sum[k1] = (before + v0) + v1;
    T.suffix_count[k1] = total - T.start[k1];
  }
  if (threadIdx.x == 0) {
    T.suffix_sum[nb] = 0.0;
    T.suffix_count[nb] = 0ull;
  }
}

__global__ void __launch_bounds__(kSearchBlock)
k_bucket_suffix(BucketPlan plan, BucketTables T)
{
  __shared__ double s_sum[kSearchBlock];
  bucket_suffix_body(plan, T, s_sum);
}

// ---- several ranks: the tables of the CONCATENATED list ---------------------------------------------------------
// Counts and integer bucket sums are all-reduced ONCE per search; the suffix tables built from the totals are then
// bit for bit those a single rank would build over the concatenated list (integer additions commute), every rank
// holds them, and the bisection needs a peer exchange only for the probes the tables cannot decide (k_search<true>).
// The totals travel as doubles through the communicator's sum all-reduce: after the carry each piece is an integer
// below 2^45 (hi) or 2^32 (lo) or 2^24 (count), so sums over <= 16 ranks are exact in any order.
// g: [nb] counts, [nb] low pieces, [nb] high pieces, [1] values outside the representable range
__global__ void __launch_bounds__(kSearchBlock)
k_tables_to_double(BucketPlan plan, BucketTables T, double* __restrict__ g)
{
  for (uint32_t b = threadIdx.x; b < plan.nb; b += kSearchBlock) {
    unsigned long long lo = T.acc_lo[b], hi = T.acc_hi[b];
    hi += lo >> 32;
    lo &= 0xffffffffull;
    g[b] = static_cast<double>(T.count[b]);
    g[plan.nb + b] = static_cast<double>(lo);
    g[2 * plan.nb + b] = static_cast<double>(hi);
  }
  if (threadIdx.x == 0)
    g[3 * plan.nb] = static_cast<double>(*T.outside);
}

// one block: the all-reduced totals back into integer tables, then offsets and suffix tables as for a local list
__global__ void __launch_bounds__(kSearchBlock)
k_global_tables(BucketPlan plan, const double* __restrict__ g, BucketTables G)
{
  __shared__ unsigned long long s_scan[kSearchBlock];
  __shared__ double s_sum[kSearchBlock];
  for (uint32_t b = threadIdx.x; b < plan.nb; b += kSearchBlock) {
    G.count[b] = static_cast<unsigned long long>(g[b]);
    G.acc_lo[b] = static_cast<unsigned long long>(g[plan.nb + b]);
    G.acc_hi[b] = static_cast<unsigned long long>(g[2 * plan.nb + b]);
  }
  if (threadIdx.x == 0)
    *G.outside = static_cast<unsigned long long>(g[3 * plan.nb]);
  __syncthreads();
  bucket_offsets_body(plan, G, s_scan);
  __syncthreads();
  bucket_suffix_body(plan, G, s_sum);
}

// ---- probe ----------------------------------------------------------------------------------------
// elements of bucket tb strictly above t: how many, and the exact integer sum of their mantissas at the bucket's
// scale (low / high 32-bit pieces); block-collective, results in every thread
__device__ __forceinline__ void
scan_bucket_above(double t, uint32_t tb, int eref, const BucketTables& T, unsigned long long (*s_red)[3],
                  unsigned long long* count, unsigned long long* sum_lo, unsigned long long* sum_hi)
{
  const unsigned long long lo_i = T.start[tb], hi_i = T.start[tb + 1];
  unsigned long long cnt = 0, lo = 0, hi = 0;
  auto take = [&](double e) {
    if (e > t) {
      const unsigned long long m = scaled_mantissa(static_cast<unsigned long long>(__double_as_longlong(e)), eref);
      lo += m & 0xffffffffull;
      hi += m >> 32;
      ++cnt;
    }
  };
  // one block reads the whole bucket: kAhead independent loads per thread in flight (the scan is latency bound)
  constexpr int kAhead = 8;
  unsigned long long i = lo_i + threadIdx.x;
  const unsigned long long step = blockDim.x;
  for (; i + (kAhead - 1) * step < hi_i; i += kAhead * step) {
    double e[kAhead];
#pragma unroll
    for (int k = 0; k < kAhead; ++k)
      e[k] = __ldg(&T.sorted[i + k * step]);
#pragma unroll
    for (int k = 0; k < kAhead; ++k)
      take(e[k]);
  }
  for (; i < hi_i; i += step)
    take(__ldg(&T.sorted[i]));
  block_sum3(cnt, lo, hi, s_red);
  *count = cnt;
  *sum_lo = lo;
  *sum_hi = hi;
}

// count and sum of the errors strictly above t (block-collective, result in every thread)
__device__ __forceinline__ void
probe_above(double t, const BucketPlan& plan, const BucketTables& T, unsigned long long (*s_red)[3],
            unsigned long long* count, double* sum)
{
  if (!(t == t)) {  // e > NaN is false for every e
    *count = 0ull;
    *sum = 0.0;
    return;
  }
  if (t < 0.0) {
    *count = T.suffix_count[0];
    *sum = T.suffix_sum[0];
    return;
  }
  const unsigned long long tb_bits = static_cast<unsigned long long>(__double_as_longlong(t)) & 0x7fffffffffffffffull;
  const uint32_t tb = bucket_of(tb_bits, plan);
  if (tb >= plan.nb) {  // t is above the largest error
    *count = 0ull;
    *sum = 0.0;
    return;
  }
  const int eref = bucket_exponent(tb, plan);
  unsigned long long cnt, lo, hi;
  scan_bucket_above(t, tb, eref, T, s_red, &cnt, &lo, &hi);
  *count = T.suffix_count[tb + 1] + cnt;
  *sum = T.suffix_sum[tb + 1] + integer_sum_to_double(lo, hi, eref);
}

// one probe for the host-driven (multi-rank) search: out[0] = count, out[1] = sum
__global__ void __launch_bounds__(kSearchBlock)
k_query(double t, BucketPlan plan, BucketTables T, double* out_count, double* out_sum)
{
  __shared__ unsigned long long s_red[32][3];
  unsigned long long c;
  double s;
  probe_above(t, plan, T, s_red, &c, &s);
  if (threadIdx.x == 0) {
    *out_count = static_cast<double>(c);
    *out_sum = s;
  }
}

// Inputs / outputs of the whole search.
struct SearchIO {
  // in
  double iter_errorest, finished_errorest, cum_errorest, target_error;
  double min_err, max_err;
  double n_regions;  // global, as a double (exact below 2^53)
  double max_percent_error_budget, max_active_regions_percentage;
  unsigned long long seq_base;  // probes exchanged so far on this mailbox (several ranks)
  // out
  double threshold, active_err_sum, extra_finished_errorest;
  double applied_threshold;  // threshold of the LAST PROBE: the flags the reference ends up with (see k_search)
  unsigned long long num_active;
  int pass_mem, pass_budget;
  int probes;   // -1: a peer did not answer; -2: some rank holds NaN / inf / negative errors
  int probes_computed;  // probes that were not answered from the memo table (several ranks: nor decided by the global tables)
  int exchanges;        // peer exchanges of this search (several ranks)
  // where the kernel's time went (SM clock cycles of thread 0; PGN_SEARCH_DEBUG prints them)
  long long cycles_total, cycles_scan, cycles_exchange;
  unsigned long long scanned;  // elements read by the bucket scans of this rank
};

// Heuristic_classifier::classify (heuristic_classifier.cuh:362-451), statement for statement, with
// every thread of the block running the same scalar control flow on the same values and the probe
// (apply_threshold + reduction + inner_product, :283-325) answered from the bucket tables.
//
// MULTI (several ranks, T: this rank's partition, G: the tables of the concatenated list, identical on every rank):
// a probe at threshold t needs the elements above t -- everything in the buckets above t's own bucket tb (known
// exactly from G) plus the part of bucket tb above t, which is spread over the ranks.  The search only COMPARES the
// probe's count and sum with limits, and both comparisons are monotone in them, so the bracket
//     G.suffix_count[tb+1] <= count <= G.suffix_count[tb],
//     G.suffix_sum[tb+1]   <= sum   <= G.suffix_sum[tb+1] + G.bucket_sum[tb]      (rounding is monotone)
// decides most probes without any communication; only a probe whose bracket straddles a limit -- and the one probe
// whose numbers the search returns -- is computed exactly: every rank scans its part of bucket tb, the integer
// partial sums (exact in any order) are all-reduced over the NVLink mailboxes, and the result is bit for bit what a
// single rank obtains on the concatenated list.  Every rank takes the same decisions from the same numbers.
// Execution: the bisection is SCALAR control flow (a few hundred probes, most of them answered from the memo table
// or the brackets), so ONE warp runs it; the other 31 warps of the block wait at a barrier and join only for the
// block-wide part of an exact probe -- the scan of a bucket and, with several ranks, the peer exchange.  (With all 32
// warps executing the control flow redundantly a memo-table probe cost 0.5 us of issue slots, 130 us per search.)
// The brackets are used on one rank too: there they save the scans.
//
// Tried and dropped (session L, profiles/r02_search_kernel_sessionL.txt): sub-bin tables of the bucket a bisection
// converges in (64 bins by the next 6 bits, built by one pass, all-reduced through a wide channel).  One level leaves
// the tail of every 20-step bisection undecided (its interval ends far narrower than a bin); refining recursively
// decides everything but needs as many passes as it saves scans, each dearer than a scan (2.3 us).
constexpr size_t kSearchCacheBytes = kProbeCache * (2 * sizeof(unsigned long long) + sizeof(double));
constexpr size_t kSearchSmemBytes = sizeof(double) * 5 * kSearchRow + kSearchCacheBytes;  // dynamic shared memory

struct ScanCommand {
  double t;
  uint32_t tb;
  int eref;
  int leave;
  unsigned long long seq;
};

template <bool MULTI>
__global__ void __launch_bounds__(kSearchBlock)
k_search(BucketPlan plan, BucketTables T, BucketTables G, SearchIO* io, PeerBox box)
{
  __shared__ unsigned long long s_red[32][3];
  __shared__ int s_ok;
  __shared__ double s_mine[3];
  __shared__ uint32_t s_recv[kMaxPeers * 2 * 3];
  __shared__ ScanCommand s_cmd;
  // dynamic: the tables a probe consults, staged once, and the memo table
  extern __shared__ __align__(16) unsigned char search_smem[];
  unsigned long long* s_start = reinterpret_cast<unsigned long long*>(search_smem);  // this rank's bucket bounds
  unsigned long long* s_sufc = s_start + kSearchRow;                                 // suffix counts (MULTI: global)
  double* s_sufs = reinterpret_cast<double*>(s_sufc + kSearchRow);                   // suffix sums   (MULTI: global)
  double* s_bsum = s_sufs + kSearchRow;                                              // bucket sums   (MULTI: global)
  unsigned long long* s_cnt = reinterpret_cast<unsigned long long*>(s_bsum + kSearchRow);  // counts  (MULTI: global)
  unsigned long long* c_key = s_cnt + kSearchRow;
  unsigned long long* c_cnt = c_key + kProbeCache;
  double* c_sum = reinterpret_cast<double*>(c_cnt + kProbeCache);
  for (uint32_t b = threadIdx.x; b <= plan.nb; b += blockDim.x) {
    s_start[b] = T.start[b];
    s_sufc[b] = G.suffix_count[b];
    s_sufs[b] = G.suffix_sum[b];
    if (b < plan.nb) {
      s_bsum[b] = G.bucket_sum[b];
      s_cnt[b] = G.count[b];
    }
  }
  const unsigned long long outside_global = *G.outside;
  T.start = s_start;
  for (int i = threadIdx.x; i < kProbeCache; i += blockDim.x)
    c_key[i] = ~0ull;  // a NaN pattern: never cached
  __syncthreads();

  long long c_xchg = 0;
  // block-wide (every thread, after the command barrier): elements of bucket cmd.tb above cmd.t -- count and integer
  // mantissa sum, over all ranks when MULTI; results in every thread
  auto collective_scan = [&](const ScanCommand cmd, unsigned long long* c, unsigned long long* lo,
                             unsigned long long* hi) -> bool {
    scan_bucket_above(cmd.t, cmd.tb, cmd.eref, T, s_red, c, lo, hi);
    *hi += *lo >> 32;
    *lo &= 0xffffffffull;
    if (!MULTI)
      return true;
    const long long c0 = clock64();
    double v[3] = {static_cast<double>(*c), static_cast<double>(*lo), static_cast<double>(*hi)};
    const bool ok = peer_allreduce_sum<3>(box, cmd.seq, v, s_mine, s_recv, &s_ok);
    *c = static_cast<unsigned long long>(v[0]);
    *lo = static_cast<unsigned long long>(v[1]);
    *hi = static_cast<unsigned long long>(v[2]);
    c_xchg += clock64() - c0;
    return ok;
  };

  if (threadIdx.x >= 32) {
    // worker warps: wait for a command, join the block-wide operation, until told to leave
    for (;;) {
      __syncthreads();
      const ScanCommand cmd = s_cmd;
      if (cmd.leave)
        return;
      unsigned long long c, lo, hi;
      collective_scan(cmd, &c, &lo, &hi);
    }
  }

  // ---- warp 0: the search ----------------------------------------------------------------------------------------
  const int lane = threadIdx.x;
  int computed = 0, exchanges = 0;
  const long long c_begin = clock64();
  long long c_scan = 0;
  unsigned long long scanned = 0;
  const SearchIO in = *io;
  const double n_d = in.n_regions;
  double range_lo = in.min_err, range_hi = in.max_err;
  double threshold = in.iter_errorest / n_d;
  double max_budget = in.max_percent_error_budget;
  double max_active = in.max_active_regions_percentage;
  bool pass_mem = false, pass_budget = false;
  unsigned long long num_active = 0;
  double active_err_sum = 0.0, finished_errorest = 0.0;
  unsigned int increases = 0, decreases = 0;
  int probes = 0;
  // The reference adopts the flag array of its last apply_threshold call (Workspace.cuh:104-106), while
  // the exhausted-attempts branch below may still reset `threshold` after that probe succeeded (when the
  // 20th decrease has just been used up): the accepted flags then belong to the last PROBED threshold,
  // not to the value `threshold` ends with.  Keep both.
  double applied = threshold;
  unsigned long long seq = in.seq_base;
  bool failed = false;

  // state of the last probe: exact numbers (num_active, active_err_sum) or a bracket
  bool exact = true;
  unsigned long long cnt_lo = 0, cnt_hi = 0;
  double sum_lo = 0.0, sum_hi = 0.0;
  // threshold whose sum produced finished_errorest, and whether that sum was exact
  double fin_thr = threshold;
  bool fin_exact = true;

  auto post = [&](bool leave, double t, uint32_t tb, int eref, bool exchange) {
    if (lane == 0) {
      s_cmd.t = t;
      s_cmd.tb = tb;
      s_cmd.eref = eref;
      s_cmd.leave = leave ? 1 : 0;
      s_cmd.seq = exchange ? ++seq : seq;
    }
    seq = __shfl_sync(0xffffffffu, seq, 0);
    __syncthreads();  // the command barrier: the worker warps pick the command up
  };

  if (MULTI && outside_global != 0ull) {  // some rank holds NaN / inf / negative errors: every rank sees it here
    if (lane == 0) {
      io->probes = -2;
      io->seq_base = seq;
    }
    post(true, 0.0, 0u, 0, false);
    return;
  }

  auto cache_slot = [&](unsigned long long tbits) {
    return static_cast<unsigned int>((tbits ^ (tbits >> 19) ^ (tbits >> 41)) & (kProbeCache - 1));
  };
  // exact numbers of the probe at t: from the memo table, trivially, or by a block-wide scan (MULTI: + one exchange)
  auto probe_exact = [&](double t) {
    const unsigned long long tbits = static_cast<unsigned long long>(__double_as_longlong(t));
    const unsigned int slot = cache_slot(tbits);
    if (c_key[slot] == tbits) {
      num_active = c_cnt[slot];
      active_err_sum = c_sum[slot];
      return;
    }
    const uint32_t tb = bucket_of(tbits & 0x7fffffffffffffffull, plan);
    if (!(t == t) || (!(t < 0.0) && tb >= plan.nb)) {  // e > NaN is false for every e; t above the largest error
      num_active = 0ull;
      active_err_sum = 0.0;
    } else if (t < 0.0) {
      num_active = s_sufc[0];
      active_err_sum = s_sufs[0];
    } else {
      const int eref = bucket_exponent(tb, plan);
      const long long c0 = clock64();
      post(false, t, tb, eref, MULTI);
      const ScanCommand cmd = s_cmd;
      unsigned long long c, lo, hi;
      const bool ok = collective_scan(cmd, &c, &lo, &hi);
      c_scan += clock64() - c0;
      scanned += s_start[tb + 1] - s_start[tb];
      if (MULTI)
        ++exchanges;
      if (!ok) {
        failed = true;
        return;
      }
      num_active = s_sufc[tb + 1] + c;
      active_err_sum = s_sufs[tb + 1] + integer_sum_to_double(lo, hi, eref);
    }
    __syncwarp();  // every lane has read the old entry (only this warp uses the memo table)
    if (lane == 0 && t == t) {
      c_key[slot] = tbits;
      c_cnt[slot] = num_active;
      c_sum[slot] = active_err_sum;
    }
    __syncwarp();
    ++computed;
  };
  // bracket of the probe at t from the bucket tables; exact when the memo table has it or when t lies outside every
  // populated bucket
  auto probe_bracket = [&](double t) {
    const unsigned long long tbits = static_cast<unsigned long long>(__double_as_longlong(t));
    const unsigned int slot = cache_slot(tbits);
    exact = true;
    if (c_key[slot] == tbits) {
      num_active = c_cnt[slot];
      active_err_sum = c_sum[slot];
      return;
    }
    const uint32_t tb = bucket_of(tbits & 0x7fffffffffffffffull, plan);
    if (!(t == t) || (!(t < 0.0) && tb >= plan.nb)) {
      num_active = 0ull;
      active_err_sum = 0.0;
    } else if (t < 0.0) {
      num_active = s_sufc[0];
      active_err_sum = s_sufs[0];
    } else if (s_cnt[tb] == 0ull) {
      num_active = s_sufc[tb + 1];
      active_err_sum = s_sufs[tb + 1];  // + integer_sum_to_double(0, 0, .) = + 0.0
    } else {
      exact = false;
      cnt_lo = s_sufc[tb + 1];
      cnt_hi = s_sufc[tb];
      sum_lo = s_sufs[tb + 1];
      sum_hi = sum_lo + s_bsum[tb];
    }
  };
  // pct <= max_active with the reference's int casts and division (heuristic_classifier.cuh:394-396).  Monotone in
  // the count, so the largest passing count is found once per value of max_active (31 divisions) and a probe then
  // costs an integer comparison instead of a correctly rounded division on the scalar critical path.
  const bool small_n = n_d < 2147483648.0;
  auto mem_ok_div = [&](unsigned long long c) {
    const double pct = static_cast<double>(static_cast<int>(c)) / static_cast<double>(static_cast<int>(n_d));
    return pct <= max_active;
  };
  double limit_for = -1.0;  // max_active the cached limit belongs to
  long long limit = -1;     // largest count that passes (-1: none)
  auto mem_ok = [&](unsigned long long c) {
    if (!small_n || c > static_cast<unsigned long long>(static_cast<long long>(n_d)))
      return mem_ok_div(c);
    if (limit_for != max_active) {
      long long a = -1, b = static_cast<long long>(n_d);  // a passes (or is -1), b fails (or is n and untested)
      if (mem_ok_div(static_cast<unsigned long long>(b))) {
        a = b;
      } else {
        while (b - a > 1) {
          const long long mid = a + (b - a) / 2;
          if (mem_ok_div(static_cast<unsigned long long>(mid)))
            a = mid;
          else
            b = mid;
        }
      }
      limit = a;
      limit_for = max_active;
    }
    return static_cast<long long>(c) <= limit;
  };
  auto budget_ok = [&](double sum, double budget) {
    return in.iter_errorest - sum - in.finished_errorest <= max_budget * budget;
  };

  do {
    if (!pass_mem && increases <= 20) {
      pass_mem = false;
      unsigned int attempts = 0;
      while (!pass_mem && attempts < 20) {
        // The reference restarts its bisection from the same threshold every time it relaxes a budget,
        // so most probes of a long search repeat earlier ones: (global) results are memoised by the
        // bit pattern of the threshold.  Every rank holds the same table, so a hit needs no exchange.
        ++probes;
        applied = threshold;
        probe_bracket(threshold);
        if (!exact && mem_ok(cnt_lo) != mem_ok(cnt_hi)) {
          probe_exact(threshold);
          exact = true;
        }
        if (failed)
          break;
        pass_mem = exact ? mem_ok(num_active) : mem_ok(cnt_hi);
        if (!pass_mem) {
          range_lo = threshold;
          threshold += fabs(range_hi - threshold) * .5;
        }
        attempts++;
      }
      increases++;
    }
    if (failed)
      break;
    if (pass_mem) {
      const double budget = in.target_error - in.cum_errorest;
      if (!exact) {
        // extra is monotone (decreasing) in the sum: the bracket's two ends give its two extremes
        const bool a = budget_ok(sum_hi, budget), b = budget_ok(sum_lo, budget);
        if (a != b) {
          probe_exact(applied);
          exact = true;
          if (failed)
            break;
        } else {
          pass_budget = a;
          finished_errorest = in.iter_errorest - sum_lo - in.finished_errorest;  // placeholder, see fin_exact
          fin_thr = applied;
          fin_exact = false;
        }
      }
      if (exact) {
        const double extra = in.iter_errorest - active_err_sum - in.finished_errorest;
        pass_budget = extra <= max_budget * budget;
        finished_errorest = extra;
        fin_thr = applied;
        fin_exact = true;
      }
      if (!pass_budget && decreases <= 20) {
        range_hi = threshold;
        threshold -= fabs(range_lo - threshold) * .5;
        pass_mem = false;
        decreases++;
      }
    }
    const bool exhausted = decreases >= 20 || increases >= 20;
    if (exhausted && max_budget < .7) {
      max_budget += 0.1;
      decreases = 0;
      increases = 0;
      range_lo = in.min_err;
      range_hi = in.max_err;
      threshold = in.iter_errorest / n_d;
    } else if (exhausted && max_budget >= .7 && max_active <= .7) {
      max_active += .1;
    } else if (exhausted) {
      break;
    }
  } while (!pass_mem || !pass_budget);

  if (!failed) {
    // the numbers the search returns are exact ones: those of the last probe, and the sum behind finished_errorest
    if (!fin_exact) {
      probe_exact(fin_thr);
      if (!failed)
        finished_errorest = in.iter_errorest - active_err_sum - in.finished_errorest;
    }
    if (!failed && (!exact || !fin_exact))
      probe_exact(applied);  // (a memo hit when it is the threshold just computed)
  }
  post(true, 0.0, 0u, 0, false);
  if (lane == 0) {
    if (failed) {
      io->probes = -1;
    } else {
      io->threshold = threshold;
      io->applied_threshold = applied;
      io->active_err_sum = active_err_sum;
      io->extra_finished_errorest = finished_errorest;
      io->num_active = num_active;
      io->pass_mem = pass_mem ? 1 : 0;
      io->pass_budget = pass_budget ? 1 : 0;
      io->probes = probes;
      io->probes_computed = computed;
      io->exchanges = exchanges;
      io->cycles_total = clock64() - c_begin;
      io->cycles_scan = c_scan;
      io->cycles_exchange = c_xchg;
      io->scanned = scanned;
    }
    io->seq_base = seq;
  }
}

}  // namespace pgn
#endif
