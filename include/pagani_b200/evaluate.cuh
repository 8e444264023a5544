// Region-evaluation kernel of the B200-native PAGANI engine (header template: it is compiled with
// the caller's integrand functor inlined, everything else lives behind the C ABI in pagani_b200.h).
//
// Replaces quad::INTEGRATE_GPU_PHASE1 -> INIT_REGION_POOL -> SampleRegionBlock -> computePermutation
// (reference cuda/pagani/quad/GPUquad/Phases.cuh:301-404, Sample.cuh:49-290) and is BIT-IDENTICAL
// to it per region (estimate, error, split axis), by construction:
//
//  * The reference runs one 64-thread block per region; thread t accumulates points t, t+64, ...
//    into 5 rule sums with fma(fun, w, sum), reduces each warp with a shuffle-down tree and adds
//    the two warp results.  Here ONE WARP owns a region and every lane plays reference threads
//    `lane` (accumulator set A) and `lane+32` (set B): the same per-thread chains, the same tree
//    (A and B are reduced in the two half-warps after one xor-16 exchange), the same final A+B.
//    No __syncthreads, no thread-0 serial sections.
//  * The affine map region->integration volume is evaluated with the reference's exact operation
//    sequence (x = fma(range, fma(.5+g, lo, (.5-g)*hi), low); checked against its PTX) but only
//    once per (axis, generator value): a coordinate can take just 11 values per axis, so each warp
//    builds an 11 x ndim table in shared memory per region and a point becomes ndim table look-ups.
//  * Fourth differences (split axis) and the 27 null-rule terms are computed lane-parallel; max and
//    first-wins argmax are order independent, so this does not change results.
//
// Two variants of the point loop exist and must agree bit for bit (tests/test_gpu_parity.py):
//   evaluate_kernel     look-ups addressed with shifts/masks from a 4-bit-per-axis code word;
//   evaluate_kernel_bp  ("byte permute", the default for ndim <= 12): the code word of a point holds
//                       one BYTE per axis = 8 x its coordinate code, the per-warp table and the weight
//                       table sit on 256-byte boundaries, and a shared-memory address is formed by ONE
//                       PRMT that splices that byte under the table base; the axis offset is the
//                       immediate of the LDS.  1 ALU + 1 LDS per coordinate instead of 3 ALU + 1 LDS:
//                       for cheap integrands the kernel is issue-bound, not FP64-bound, so integer
//                       instructions are what there is to save.
//
// Persistent launch: grid = SMs x resident blocks, warps stride over the region list.
#ifndef PAGANI_B200_EVALUATE_CUH
#define PAGANI_B200_EVALUATE_CUH

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <atomic>
#include <cstring>
#include <new>
#include <type_traits>
#include <utility>

#include "../pagani_b200.h"

namespace pagani_b200 {

#ifndef PGN_EVAL_WARPS
#define PGN_EVAL_WARPS 4
#endif
constexpr int kWarpsPerBlock = PGN_EVAL_WARPS;
#ifndef PGN_EVAL_MIN_BLOCKS
#define PGN_EVAL_MIN_BLOCKS 4
#endif
// 1: two chunks (four points per lane) per trip of the point loop inside a run of single-set chunks.
// Measured on B200 (profiles/r01_cos_sweep2.jsonl): no gain over two-point trips -- the loop is bound by
// issue slots (an FP64 instruction holds the port for two cycles), not by dependent-issue latency.
#ifndef PGN_EVAL_QUAD
#define PGN_EVAL_QUAD 0
#endif
// 1: double-buffer the coordinates of the point loop (loads of trip t+1 issued before the arithmetic of trip t)
#ifndef PGN_EVAL_PIPELINE
#define PGN_EVAL_PIPELINE 0
#endif

// 1: the middle trips run as compile-time segments (TripPlan: one copy of the trip body per run of trips that share
// a generator set, weights as constant-bank operands, no per-trip set lookup); 0: one loop over all middle trips.
// Measured on B200 (profiles/r02_evalbench_h.jsonl): the segments save 14 of 134 instructions per trip but multiply
// the code size by the number of runs -- +4 % for the 10-D C0 family (41 trips, long runs, small body), -2 % for
// F_1_8D, -27 % for the Gaussian (its pow(x, 2.) body is ~1300 instructions: seven copies no longer fit the
// instruction cache).  Off by default; a functor opts in through EvalTuning<F>::kSegments.
#ifndef PGN_EVAL_SEGMENTS
#define PGN_EVAL_SEGMENTS 0
#endif

// 1: a last trip none of whose chain-B points exists (8-D: 17 points) evaluates chain A only.  Off: the extra code
// path costs registers, and with them 13 of the cosine literals ptxas otherwise keeps in registers across the trip
// loop of F_1_8D are re-loaded every trip (LDC 17 -> 30 per trip, 176 -> 189 instructions): 314 instead of 289 us on
// 65 536 regions (ncu, session I), bench 2.45 instead of 2.73e11 evaluations/s.
#ifndef PGN_EVAL_TAIL_A_ONLY
#define PGN_EVAL_TAIL_A_ONLY 0
#endif
#ifndef PGN_EVAL_PREFETCH_BOUNDS
#define PGN_EVAL_PREFETCH_BOUNDS 1
#endif
// 1: the PER-LANE rule weights (first and last trip, trips that straddle a generator-set boundary: 6 of 17 at 8-D)
// come from a shared-memory copy of the weight table instead of the constant bank: a constant-bank read whose
// address differs between the lanes is replayed once per distinct address, a shared-memory read of two distinct rows
// is one wavefront.  +1.1 % on F_1_8D (2.815 -> 2.845e11 evaluations/s, profiles/r02_evalbench_s.jsonl), same bits;
// within +-0.3 % for the other families except the 10-D discontinuous one (-2.6 %: its gather already loads the
// shared-memory pipe).  Off by default; a functor opts in through EvalTuning<F>::kLaneWeightsSmem.
// (Loading the code words of trip t+1 during trip t, also tried there, is 0.7 % slower.)
#ifndef PGN_EVAL_LANE_WEIGHTS_SMEM
#define PGN_EVAL_LANE_WEIGHTS_SMEM 0
#endif
#ifndef PGN_EVAL_UNIT_JACOBIAN
#define PGN_EVAL_UNIT_JACOBIAN 0
#endif
// warp index as a broadcast from lane 0 (warp-uniform for the compiler)
#ifndef PGN_EVAL_UNIFORM_WARP
#define PGN_EVAL_UNIFORM_WARP 1
#endif
// kernel instantiation without the integrand's range checks where the functor promises them (fast_in_box)
#ifndef PGN_EVAL_NOCHECK
#define PGN_EVAL_NOCHECK 1
#endif

// Launch-bound tuning knob of the evaluation kernel: resident 128-thread blocks per SM the register
// allocation must allow (4 -> up to 128 registers per thread, 7 -> 72, 8 -> 64).  Cheap integrands
// want registers (the point loop is latency bound), pow()-heavy ones want warps.  Specialise for a
// functor type to override the default; the sweep behind the Genz values is in profiles/.
template <typename F>
struct EvalTuning {
  static constexpr int kMinBlocks = PGN_EVAL_MIN_BLOCKS;
  static constexpr bool kSegments = PGN_EVAL_SEGMENTS != 0;
  static constexpr bool kLaneWeightsSmem = PGN_EVAL_LANE_WEIGHTS_SMEM != 0;
};
// kSegments of a specialisation that does not declare it: the default
template <typename F, class = void>
struct eval_segments : std::integral_constant<bool, PGN_EVAL_SEGMENTS != 0> {};
template <typename F>
struct eval_segments<F, std::void_t<decltype(EvalTuning<F>::kSegments)>>
  : std::integral_constant<bool, EvalTuning<F>::kSegments> {};
// kUnitJacobianVariant: a second instantiation of the kernel for a unit Jacobian (integration over the unit cube),
// without the multiplication of every function value by it (f * 1.0 = f exactly) and the load of the factor
template <typename F, class = void>
struct eval_unit_jacobian_variant : std::integral_constant<bool, PGN_EVAL_UNIT_JACOBIAN != 0> {};
template <typename F>
struct eval_unit_jacobian_variant<F, std::void_t<decltype(EvalTuning<F>::kUnitJacobianVariant)>>
  : std::integral_constant<bool, EvalTuning<F>::kUnitJacobianVariant> {};
// likewise kUniformWarp
template <typename F, class = void>
struct eval_uniform_warp : std::integral_constant<bool, PGN_EVAL_UNIFORM_WARP != 0> {};
template <typename F>
struct eval_uniform_warp<F, std::void_t<decltype(EvalTuning<F>::kUniformWarp)>>
  : std::integral_constant<bool, EvalTuning<F>::kUniformWarp> {};
// likewise kLaneWeightsSmem
template <typename F, class = void>
struct eval_lane_weights_smem : std::integral_constant<bool, PGN_EVAL_LANE_WEIGHTS_SMEM != 0> {};
template <typename F>
struct eval_lane_weights_smem<F, std::void_t<decltype(EvalTuning<F>::kLaneWeightsSmem)>>
  : std::integral_constant<bool, EvalTuning<F>::kLaneWeightsSmem> {};
constexpr int kWeightRow = 6;  // 5 rule weights padded to 48 B so a row is three 16-byte loads

// values of pgn_region_view::reserved
enum EvalVariant { kVariantDefault = 0, kVariantShift = 1, kVariantBytePermute = 2 };

namespace detail {

template <class F, size_t N, size_t... I>
__device__ __forceinline__ double
apply_impl(F& f, const double (&x)[N], std::index_sequence<I...>)
{
  return f(x[I]...);
}

// call f(x[0], ..., x[N-1])   (reference: gpu::apply, common/cuda/cudaApply.cuh:18-26)
template <class F, size_t N>
__device__ __forceinline__ double
apply(F& f, const double (&x)[N])
{
  return apply_impl(f, x, std::make_index_sequence<N>());
}

// Optional functor protocol: a functor may expose
//     template <int M> void batch(const double (&x)[M][N], double (&f)[M]) const
// returning f[m] == operator()(x[m][0], ..., x[m][N-1]) exactly.  The point loop evaluates the M = 2 or 4
// points a lane owns in one or two chunks through it, which lets an integrand share its rare slow paths
// between them (see genz::Oscillatory) so that the common paths are one straight-line block the compiler
// interleaves: the loop is bound by dependent-issue latency, not by any pipe.
template <class F, int N, int M, class = void>
struct has_batch : std::false_type {};
template <class F, int N, int M>
struct has_batch<F, N, M,
                 std::void_t<decltype(std::declval<const F&>().template batch<M>(
                   std::declval<const double (&)[M][N]>(), std::declval<double (&)[M]>()))>>
  : std::true_type {};

// f[m] = functor(x[m]...) for m < M: through the batch protocol when the functor has it
template <class F, int N, int M, bool HAS = has_batch<F, N, M>::value>
struct BatchCall {
  __device__ __forceinline__ static void
  run(F* f, const double (&x)[M][N], double (&out)[M])
  {
#pragma unroll
    for (int m = 0; m < M; ++m)
      out[m] = apply<F, N>(*f, x[m]);
  }
};
template <class F, int N, int M>
struct BatchCall<F, N, M, true> {
  __device__ __forceinline__ static void
  run(F* f, const double (&x)[M][N], double (&out)[M])
  {
    static_cast<const F*>(f)->template batch<M>(x, out);
  }
};

// Optional functor protocol, second form:
//     template <int M> bool batch_fast(const double (&x)[M][N], double (&f)[M]) const
// straight-line code with NO function call and no slow path: returns true when every f[m] equals
// operator()(x[m]...) exactly, false when some argument needs the functor's slow path (f is then
// meaningless).  The point loop calls it instead of batch/operator() and, in the rare case that a lane
// reported false, re-evaluates the whole region afterwards with operator() (evaluate_region_exact, not
// inlined).  Keeping the library's slow paths (calls into libdevice, their uniform-register clobbers and the
// re-convergence points around them) out of the loop body lets the compiler keep every constant of the fast
// path resident: for the oscillatory family that is 11 fewer instructions per trip.
template <class F, int N, int M, class = void>
struct has_batch_fast : std::false_type {};
template <class F, int N, int M>
struct has_batch_fast<F, N, M,
                      std::void_t<decltype(std::declval<const F&>().template batch_fast<M>(
                        std::declval<const double (&)[M][N]>(), std::declval<double (&)[M]>()))>>
  : std::true_type {};

// Optional launch-time promise on top of batch_fast: `bool fast_in_box(const double* lo, const double* hi) const`
// (host) says that every point of the integration volume takes batch_fast's fast path (its arguments are in range
// whatever the region), and `batch_nocheck<M>(x, f)` is batch_fast without the per-point range checks (the same
// values).  The launcher then picks a kernel instantiation without the checks, the vote and the fallback.
// The promise is asked for the integration VOLUME: it covers every region because regions are expressed in unit-cube
// coordinates, the invariant of the reference's region store (Sub_regions.cuh:11-36); a caller that imports regions
// outside [0, 1]^ndim must not offer it (fast_in_box should then answer for the box those regions map to).
template <class F, int N, int M, class = void>
struct has_batch_nocheck : std::false_type {};
template <class F, int N, int M>
struct has_batch_nocheck<F, N, M,
                         std::void_t<decltype(std::declval<const F&>().template batch_nocheck<M>(
                                       std::declval<const double (&)[M][N]>(), std::declval<double (&)[M]>())),
                                     decltype(std::declval<const F&>().fast_in_box(
                                       std::declval<const double*>(), std::declval<const double*>()))>>
  : std::true_type {};

struct EvalParams {
  const double* left;
  const double* length;
  size_t n;
  double* estimate;
  double* errorest;
  int32_t* axis;
  const uint64_t* point_code;
  const double* weights;
  const double* scale;
  const double* norm;
  double lows[PGN_MAX_NDIM + 1];
  double ranges[PGN_MAX_NDIM + 1];
  double half_plus_g[PGN_NCODES];
  double half_minus_g[PGN_NCODES];
  double wc[PGN_NSETS * PGN_NRULES];  // rule weights [set][rule] (byte-permute variant: constant bank)
  uint8_t half_set[256];              // generator set of points 32j..32j+31 if they share one, else 0xFF
  uint8_t chunk_set[128];             // the same for points 64k..64k+63
  double jacobian;
  double ratio;
  int feval;
};

// first point of generator set s (sets follow each other in point order, Rule.cuh:438-505)
__host__ __device__ constexpr int
set_begin(int n, int s)
{
  const int count[PGN_NSETS] = {1,
                                2 * n,
                                2 * n,
                                2 * n,
                                2 * n,
                                2 * n * (n - 1),
                                4 * n * (n - 1),
                                4 * n * (n - 1) * (n - 2) / 3,
                                1 << n};
  int b = 0;
  for (int k = 0; k < s; ++k)
    b += count[k];
  return b;
}

// generator set of point p: number of set boundaries at or below p
template <int NDIM>
__device__ __forceinline__ int
set_of_point(int p)
{
  int s = 0;
#pragma unroll
  for (int k = 1; k < PGN_NSETS; ++k)
    s += p >= set_begin(NDIM, k) ? 1 : 0;
  return s;
}

// How the functor reaches the kernel: small trivially-copyable functors travel by value in the
// kernel's constant parameter space (state reads cost nothing); anything else is read through a
// copy-constructed object in managed memory, as the reference does for every functor
// (cuda_copy_to_managed, common/cuda/cudaMemoryUtil.h:181-196).
template <typename F>
struct pass_by_value
  : std::integral_constant<bool, std::is_trivially_copyable<F>::value && (sizeof(F) <= 2048)> {};

template <typename F, bool BY_VALUE>
struct FunctorArg;
template <typename F>
struct FunctorArg<F, true> {
  F f;
  // the reference's functors have non-const call operators; they do not write to themselves
  __device__ __forceinline__ F* get() const { return const_cast<F*>(&f); }
};
template <typename F>
struct FunctorArg<F, false> {
  F* p;
  __device__ __forceinline__ F* get() const { return p; }
};

__host__ __device__ constexpr int
feval_of(int n)
{
  return 1 + 8 * n + 6 * n * (n - 1) + 4 * n * (n - 1) * (n - 2) / 3 + (1 << n);
}

// 64-bit shuffles as two 32-bit ones, LOW word first (the toolkit's double overloads shuffle the high word first):
// the two results then land in an aligned register pair in issue order
#ifndef PGN_EVAL_SHFL_LO_FIRST
#define PGN_EVAL_SHFL_LO_FIRST 1
#endif
__device__ __forceinline__ double
shfl_xor_f64(double v, int mask, int width = 32)
{
#if PGN_EVAL_SHFL_LO_FIRST
  const int lo = __shfl_xor_sync(0xffffffffu, __double2loint(v), mask, width);
  const int hi = __shfl_xor_sync(0xffffffffu, __double2hiint(v), mask, width);
  return __hiloint2double(hi, lo);
#else
  return __shfl_xor_sync(0xffffffffu, v, mask, width);
#endif
}
__device__ __forceinline__ double
shfl_down_f64(double v, unsigned delta, int width = 32)
{
#if PGN_EVAL_SHFL_LO_FIRST
  const int lo = __shfl_down_sync(0xffffffffu, __double2loint(v), delta, width);
  const int hi = __shfl_down_sync(0xffffffffu, __double2hiint(v), delta, width);
  return __hiloint2double(hi, lo);
#else
  return __shfl_down_sync(0xffffffffu, v, delta, width);
#endif
}
__device__ __forceinline__ double
shfl_idx_f64(double v, int src, int width = 32)
{
#if PGN_EVAL_SHFL_LO_FIRST
  const int lo = __shfl_sync(0xffffffffu, __double2loint(v), src, width);
  const int hi = __shfl_sync(0xffffffffu, __double2hiint(v), src, width);
  return __hiloint2double(hi, lo);
#else
  return __shfl_sync(0xffffffffu, v, src, width);
#endif
}

// first-wins arg-max over lanes (value v >= 0, index idx); ties go to the smaller index
template <int WIDTH>
__device__ __forceinline__ void
warp_argmax_first(double& v, int& idx)
{
#pragma unroll
  for (int off = WIDTH / 2; off > 0; off >>= 1) {
    const double ov = shfl_xor_f64(v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, idx, off);
    if (ov > v || (ov == v && oi < idx)) {
      v = ov;
      idx = oi;
    }
  }
}

// ---- per-region prologue: bounds, volume, default axis, coordinate table -------------------------
// (INIT_REGION_POOL, Phases.cuh:323-349, and the affine map of computePermutation, Sample.cuh:124-131)
// bounds of region `reg` on axis `lane` (lanes < NDIM): the only global loads of a region
template <int NDIM>
__device__ __forceinline__ void
load_bounds(const EvalParams& P, size_t reg, int lane, double& lo, double& len)
{
  if (lane < NDIM) {
    lo = P.left[static_cast<size_t>(lane) * P.n + reg];
    len = P.length[static_cast<size_t>(lane) * P.n + reg];
  }
}

template <int NDIM>
__device__ __forceinline__ void
region_prologue(const EvalParams& P, double lo, double len, int lane, const double* s_tp, const double* s_tm,
                double* s_lo, double* s_up, double* s_ext, double* s_x, double& vol, int& max_dim)
{
  if (lane < NDIM) {
    const double up = __dadd_rn(lo, len);
    s_lo[lane] = lo;
    s_up[lane] = up;
    s_ext[lane] = __dsub_rn(up, lo);
  }
  __syncwarp();
  // volume and longest axis from the NDIM extents, which every lane reads anyway (broadcast loads): the product in
  // axis order (Phases.cuh:331-333) and the FIRST axis with the largest positive extent (:344-347), in registers --
  // no shuffle rounds
  {
    double ext[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d)
      ext[d] = s_ext[d];
    vol = ext[0];
    double best = 0.0;
    max_dim = 0;
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
      if (d > 0)
        vol = __dmul_rn(vol, ext[d]);
      if (ext[d] > best) {
        best = ext[d];
        max_dim = d;
      }
    }
  }
  // coordinate table: lane -> (axis = lane / 4, codes lane % 4 + {0, 4, 8}), eight axes per pass; no index division,
  // one load of the axis' bounds per lane
#pragma unroll
  for (int d0 = 0; d0 < NDIM; d0 += 8) {
    const int d = d0 + (lane >> 2);
    if (d < NDIM) {
      const double lo_d = s_lo[d], up_d = s_up[d];
      const double range = P.ranges[d], low = P.lows[d];
#pragma unroll
      for (int c = lane & 3; c < PGN_NCODES; c += 4) {
        const double inner = __fma_rn(s_tp[c], lo_d, __dmul_rn(s_tm[c], up_d));
        s_x[d * PGN_NCODES + c] = __fma_rn(range, inner, low);
      }
    }
  }
  __syncwarp();
}

// ---- per-region epilogue: split axis, block reduction replay, null-rule error, stores ------------
template <int NDIM>
__device__ __forceinline__ void
region_epilogue(const EvalParams& P, size_t reg, int lane, const double* s_f, const double* s_scale,
                const double* s_norm, const double (&sa)[PGN_NRULES], const double (&sb)[PGN_NRULES],
                double vol, int max_dim, double ratio, double one_minus_ratio)
{
  constexpr int kArgWidth = NDIM <= 16 ? 16 : 32;
  const bool lo_half = lane < 16;
  // split axis: first axis with the largest fourth difference (Sample.cuh:194-218)
  int bisect;
  {
    double fd = 0.0;
    int idx = lane;
    if (lane < NDIM) {
      const double f0 = s_f[0];
      const double fp = s_f[1 + 2 * lane], fm = s_f[2 + 2 * lane];
      const double gp = s_f[1 + 2 * lane + 2 * NDIM], gm = s_f[2 + 2 * lane + 2 * NDIM];
      // `base + ratio * (fp + fm) - (gp + gm)` with base = 2 f0 (1 - ratio): which of the two products the
      // reference's compiled code fuses depends on ndim (nvcc 12.9, read off the SASS of INTEGRATE_GPU_PHASE1 for
      // ndim 2..11).  Up to 4 axes the loop over the axes is unrolled before contraction and the base product goes
      // into the FMA; from 5 axes on base is a rounded loop invariant.  For smooth integrands the difference is far
      // below the fourth differences, for (bi)linear ones they ARE rounding noise and this decides the axis
      // (golden xy_2_1e-06, region 9).
      double t;
      if constexpr (NDIM <= 4) {
        t = __fma_rn(__dadd_rn(f0, f0), one_minus_ratio, __dmul_rn(ratio, __dadd_rn(fp, fm)));
      } else {
        const double base = __dmul_rn(__dadd_rn(f0, f0), one_minus_ratio);
        t = __fma_rn(ratio, __dadd_rn(fp, fm), base);
      }
      const double v = fabs(__dsub_rn(t, __dadd_rn(gp, gm)));
      fd = v > 0.0 ? v : 0.0;
    }
    warp_argmax_first<kArgWidth>(fd, idx);
    bisect = fd > 0.0 ? idx : max_dim;
  }
  // the reference's block reduction (Sample.cuh:49-83) replayed inside one warp
  double tot[PGN_NRULES];
#pragma unroll
  for (int r = 0; r < PGN_NRULES; ++r) {
    const double keep = lo_half ? sa[r] : sb[r];
    const double send = lo_half ? sb[r] : sa[r];
    const double recv = shfl_xor_f64(send, 16);
    double v = __dadd_rn(keep, recv);  // lanes 0-15: A[i]+A[i+16]; lanes 16-31: B[i]+B[i+16]
    v = __dadd_rn(v, shfl_down_f64(v, 8, 16));
    v = __dadd_rn(v, shfl_down_f64(v, 4, 16));
    v = __dadd_rn(v, shfl_down_f64(v, 2, 16));
    v = __dadd_rn(v, shfl_down_f64(v, 1, 16));
    tot[r] = __dadd_rn(shfl_idx_f64(v, 0), shfl_idx_f64(v, 16));
  }
  // null-rule error (Sample.cuh:261-289): lane = 3*set + (rule-1)
  double term = 0.0;
  if (lane < 3 * PGN_NSETS) {
    const int s = lane / 3;
    const int k = lane - 3 * s;  // rule-1
    const double a = k == 0 ? tot[1] : (k == 1 ? tot[2] : tot[3]);
    const double b = k == 0 ? tot[2] : (k == 1 ? tot[3] : tot[4]);
    term = __dmul_rn(fabs(__fma_rn(s_scale[s * PGN_NRULES + k + 1], a, b)),
                     s_norm[s * PGN_NRULES + k + 1]);
    term = fmax(0.0, term);
  }
  term = fmax(term, shfl_down_f64(term, 12));
  term = fmax(term, shfl_down_f64(term, 6));
  {
    const double t3 = shfl_down_f64(term, 3);
    const double t6 = shfl_down_f64(term, 6);
    term = fmax(fmax(term, t3), t6);
  }
  const double e1 = shfl_idx_f64(term, 0);
  const double e2 = shfl_idx_f64(term, 1);
  const double e3 = shfl_idx_f64(term, 2);
  if (lane == 0) {
    const bool benign = (__dmul_rn(5.0, e1) <= e2) && (__dmul_rn(5.0, e2) <= e3);
    const double inner = benign ? e1 : __dmul_rn(5.0, fmax(fmax(e1, e2), e3));
    P.estimate[reg] = __dmul_rn(vol, tot[0]);
    P.errorest[reg] = __dmul_rn(vol, inner);
    P.axis[reg] = bisect;
  }
}

// =================================================================================================
// variant 1: shift/mask addressing from the 4-bit code words
// =================================================================================================
template <int NDIM>
__host__ __device__ constexpr int
per_warp_doubles()
{
  // coordinate table, first-chunk function values, lower, upper, extent
  return NDIM * PGN_NCODES + 64 + 3 * NDIM;
}

template <int NDIM>
__host__ __device__ constexpr size_t
smem_bytes(bool codes_in_smem)
{
  return sizeof(double) * (PGN_NSETS * kWeightRow + 2 * PGN_NSETS * PGN_NRULES + 2 * PGN_NCODES +
                           kWarpsPerBlock * per_warp_doubles<NDIM>()) +
         (codes_in_smem ? sizeof(uint64_t) * static_cast<size_t>(feval_of(NDIM)) : 0);
}

template <typename F, int NDIM, bool BY_VALUE, bool CODES_IN_SMEM>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
evaluate_kernel(const __grid_constant__ FunctorArg<F, BY_VALUE> farg, const __grid_constant__ EvalParams P)
{
  static_assert(NDIM >= 2 && NDIM <= PGN_MAX_NDIM, "4*ndim+1 must fit the 64 emulated threads");
  extern __shared__ __align__(16) unsigned char smem_raw[];

  double* s_w = reinterpret_cast<double*>(smem_raw);         // [9][6]
  double* s_scale = s_w + PGN_NSETS * kWeightRow;            // [9][5]
  double* s_norm = s_scale + PGN_NSETS * PGN_NRULES;         // [9][5]
  double* s_tp = s_norm + PGN_NSETS * PGN_NRULES;            // [11]
  double* s_tm = s_tp + PGN_NCODES;                          // [11]
  double* s_warp0 = s_tm + PGN_NCODES;
  const uint64_t* codes = P.point_code;
  if constexpr (CODES_IN_SMEM) {
    uint64_t* s_codes =
      reinterpret_cast<uint64_t*>(s_warp0 + kWarpsPerBlock * per_warp_doubles<NDIM>());
    for (int i = threadIdx.x; i < P.feval; i += blockDim.x)
      s_codes[i] = __ldg(&P.point_code[i]);
    codes = s_codes;
  }
  for (int i = threadIdx.x; i < PGN_NSETS * kWeightRow; i += blockDim.x) {
    const int s = i / kWeightRow, r = i - s * kWeightRow;
    s_w[i] = r < PGN_NRULES ? __ldg(&P.weights[s * PGN_NRULES + r]) : 0.0;
  }
  for (int i = threadIdx.x; i < PGN_NSETS * PGN_NRULES; i += blockDim.x) {
    s_scale[i] = __ldg(&P.scale[i]);
    s_norm[i] = __ldg(&P.norm[i]);
  }
  if (threadIdx.x < PGN_NCODES) {
    s_tp[threadIdx.x] = P.half_plus_g[threadIdx.x];
    s_tm[threadIdx.x] = P.half_minus_g[threadIdx.x];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  double* s_x = s_warp0 + warp * per_warp_doubles<NDIM>();  // [NDIM][11]
  double* s_f = s_x + NDIM * PGN_NCODES;                    // [64]
  double* s_lo = s_f + 64;                                  // [NDIM]
  double* s_up = s_lo + NDIM;                               // [NDIM]
  double* s_ext = s_up + NDIM;                              // [NDIM]

  // the functor: by value for empty/trivial ones, through a managed-memory copy otherwise
  // (reference: cuda_copy_to_managed / make_gpu_integrand, cudaMemoryUtil.h:181-196, 345-353)
  F* fptr = farg.get();

  const int feval = P.feval;
  const int full_chunks = feval >> 6;  // chunks of 64 points in which every emulated thread works
  const double jac = P.jacobian;
  const double ratio = P.ratio;
  const double one_minus_ratio = __dsub_rn(1.0, ratio);

  const size_t n = P.n;
  const size_t stride = static_cast<size_t>(gridDim.x) * kWarpsPerBlock;
  for (size_t reg = static_cast<size_t>(blockIdx.x) * kWarpsPerBlock + warp; reg < n; reg += stride) {
    __syncwarp();
    double vol;
    int max_dim;
    {
      double lo = 0.0, len = 0.0;
      load_bounds<NDIM>(P, reg, lane, lo, len);
      region_prologue<NDIM>(P, lo, len, lane, s_tp, s_tm, s_lo, s_up, s_ext, s_x, vol, max_dim);
    }

    double sa[PGN_NRULES], sb[PGN_NRULES];
#pragma unroll
    for (int r = 0; r < PGN_NRULES; ++r)
      sa[r] = sb[r] = 0.0;

    auto eval_point = [&](int p, double (&sum)[PGN_NRULES], bool keep_f, int slot) {
      const uint64_t code = codes[p];
      double x[NDIM];
#pragma unroll
      for (int d = 0; d < NDIM; ++d)
        x[d] = s_x[d * PGN_NCODES + static_cast<int>((code >> (4 * d)) & 0xF)];
      const double fv = apply<F, NDIM>(*fptr, x);
      const double fun = __dmul_rn(fv, jac);
      if (keep_f)
        s_f[slot] = fun;
      const double* w = s_w + static_cast<int>(code >> 60) * kWeightRow;
      const double2 w01 = *reinterpret_cast<const double2*>(w);
      const double2 w23 = *reinterpret_cast<const double2*>(w + 2);
      const double w4 = w[4];
      sum[0] = __fma_rn(fun, w01.x, sum[0]);
      sum[1] = __fma_rn(fun, w01.y, sum[1]);
      sum[2] = __fma_rn(fun, w23.x, sum[2]);
      sum[3] = __fma_rn(fun, w23.y, sum[3]);
      sum[4] = __fma_rn(fun, w4, sum[4]);
    };

    // chunk 0 keeps its function values for the fourth differences (Sample.cuh:176-219)
    {
      eval_point(lane, sa, true, lane);
      if (lane + 32 < feval)
        eval_point(lane + 32, sb, true, lane + 32);
    }
    // full chunks: both emulated threads of every lane are busy
    for (int k = 1; k < full_chunks; ++k) {
      const int p = (k << 6) + lane;
      eval_point(p, sa, false, 0);
      eval_point(p + 32, sb, false, 0);
    }
    // ragged tail (Sample.cuh:238-252)
    if (full_chunks >= 1) {
      const int p = (full_chunks << 6) + lane;
      if (p < feval)
        eval_point(p, sa, false, 0);
      if (p + 32 < feval)
        eval_point(p + 32, sb, false, 0);
    }
    __syncwarp();
    region_epilogue<NDIM>(P, reg, lane, s_f, s_scale, s_norm, sa, sb, vol, max_dim, ratio,
                          one_minus_ratio);
  }
}

// =================================================================================================
// variant 2: byte-permute addressing
// =================================================================================================
template <int NDIM>
struct BpLayout {
  static constexpr int kCodeWords = NDIM <= 8 ? 2 : 4;  // 32-bit words per point, one byte per axis
  static constexpr int kXBytes = ((NDIM * PGN_NCODES * 8 + 255) / 256) * 256;
  static constexpr int kWarpBytes = ((kXBytes + (64 + 3 * NDIM) * 8 + 255) / 256) * 256;
  static constexpr int kMiscDoubles = 3 * PGN_NSETS * PGN_NRULES + 2 * PGN_NCODES;  // scale, norm, weights; tp, tm
  static constexpr int kCodesOffset = kWarpsPerBlock * kWarpBytes + kMiscDoubles * 8;
  // the code table is padded to whole trips of 64 points (the padding holds the centre point's code)
  static constexpr int kPaddedPoints = ((feval_of(NDIM) + 63) / 64) * 64;
  static constexpr size_t
  bytes()
  {
    return 256 /* alignment slack */ + kCodesOffset + 16 + static_cast<size_t>(kPaddedPoints) * kCodeWords * 4 +
           feval_of(NDIM);
  }
};

// Compile-time plan of the trips 1 .. kTrips-2 of a region (all 64 points exist): maximal runs of consecutive trips
// whose points belong to ONE generator set (the rule weights are then compile-time constant-bank operands of the
// accumulating FMAs and the loop over the run carries no set bookkeeping at all), separated by runs of trips that
// straddle a set boundary (per-lane weights).  8-D: mixed {1,2}, set 6 x3, mixed {6}, set 7 x6, mixed {13}, set 8 x3.
template <int NDIM>
struct TripPlan {
  static constexpr int kFeval = feval_of(NDIM);
  static constexpr int kTrips = (kFeval + 63) / 64;
  static constexpr int kCap = kTrips > 2 ? kTrips : 2;
  int nseg = 0;
  int first[kCap] = {};
  int count[kCap] = {};
  int set[kCap] = {};  // generator set of the run, -1: mixed
  static constexpr int
  set_of(int p)
  {
    int st = 0;
    for (int k = 1; k < PGN_NSETS; ++k)
      st += p >= set_begin(NDIM, k) ? 1 : 0;
    return st;
  }
  constexpr TripPlan()
  {
    for (int t = 1; t <= kTrips - 2; ++t) {
      const int s0 = set_of(64 * t), s1 = set_of(64 * t + 63);
      const int st = s0 == s1 ? s0 : -1;
      if (nseg > 0 && set[nseg - 1] == st) {
        ++count[nseg - 1];
      } else {
        first[nseg] = t;
        count[nseg] = 1;
        set[nseg] = st;
        ++nseg;
      }
    }
  }
};

template <int NDIM>
inline constexpr TripPlan<NDIM> kTripPlan{};

// fn(std::integral_constant<int, 0>{}), fn(std::integral_constant<int, 1>{}), ...
template <class Fn, int... I>
__device__ __forceinline__ void
for_constexpr(Fn&& fn, std::integer_sequence<int, I...>)
{
  (fn(std::integral_constant<int, I>{}), ...);
}

template <int IMM>
__device__ __forceinline__ double
lds_f64(uint32_t addr)
{
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(addr), "n"(IMM) : "memory");
  return v;
}

// code words of a point: 8 bytes (up to 8 axes) or 16 bytes at a 32-bit shared-memory address
template <int IMM>
__device__ __forceinline__ uint2
lds_u32x2(uint32_t addr)
{
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0, %1}, [%2+%3];" : "=r"(v.x), "=r"(v.y) : "r"(addr), "n"(IMM) : "memory");
  return v;
}
template <int IMM>
__device__ __forceinline__ uint4
lds_u32x4(uint32_t addr)
{
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4+%5];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "r"(addr), "n"(IMM)
               : "memory");
  return v;
}
__device__ __forceinline__ int
lds_u8(uint32_t addr)
{
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return static_cast<int>(v);
}

template <int NDIM, int D>
struct BpGather {
  // x[d] = table[d][code byte d], for d = D .. NDIM-1
  template <typename Words>
  __device__ __forceinline__ static void
  run(const Words& w, uint32_t xbase, double (&x)[NDIM])
  {
    const uint32_t word = D < 4 ? w.x : (D < 8 ? w.y : (D < 12 ? w.z : w.w));
    x[D] = lds_f64<D * PGN_NCODES * 8>(__byte_perm(word, xbase, 0x7650 | (D & 3)));
    BpGather<NDIM, D + 1>::run(w, xbase, x);
  }
};
template <int NDIM>
struct BpGather<NDIM, NDIM> {
  template <typename Words>
  __device__ __forceinline__ static void
  run(const Words&, uint32_t, double (&)[NDIM])
  {}
};

struct Words4 {
  uint32_t x, y, z, w;
};

// Cold path of evaluate_kernel_bp for functors with batch_fast: the whole region after the prologue, every point
// through operator() (the functor's complete code, slow paths included).  Not inlined: its calls into the math
// library must not constrain the register allocation of the hot loop.
template <typename F, int NDIM>
__device__ __noinline__ void
evaluate_region_exact(F* fptr, const EvalParams& P, size_t reg, int lane, uint32_t lane_codes, uint32_t set_addr,
                      uint32_t xbase, double* s_f, const double* s_scale, const double* s_norm, double vol,
                      int max_dim, double jac, double ratio, double one_minus_ratio)
{
  using L = BpLayout<NDIM>;
  constexpr int kFeval = feval_of(NDIM);
  constexpr int kTrips = (kFeval + 63) / 64;
  constexpr int kPointBytes = L::kCodeWords * 4;
  double sa[PGN_NRULES], sb[PGN_NRULES];
#pragma unroll
  for (int r = 0; r < PGN_NRULES; ++r)
    sa[r] = sb[r] = 0.0;
#pragma unroll 1
  for (int t = 0; t < kTrips; ++t) {
#pragma unroll 1
    for (int m = 0; m < 2; ++m) {
      const int p = (t << 6) + 32 * m + lane;
      const uint32_t a = lane_codes + static_cast<uint32_t>(t) * (64u * kPointBytes) + m * (32 * kPointBytes);
      Words4 cw;
      if constexpr (L::kCodeWords == 2) {
        const uint2 v = lds_u32x2<0>(a);
        cw = Words4{v.x, v.y, 0u, 0u};
      } else {
        const uint4 v = lds_u32x4<0>(a);
        cw = Words4{v.x, v.y, v.z, v.w};
      }
      double x[NDIM];
      BpGather<NDIM, 0>::run(cw, xbase, x);
      const double fun = __dmul_rn(apply<F, NDIM>(*fptr, x), jac);
      if (t == 0)
        s_f[32 * m + lane] = fun;
      if (p < kFeval) {
        const double* w = P.wc + lds_u8(set_addr + static_cast<uint32_t>(p)) * PGN_NRULES;
        double* sum = m == 0 ? sa : sb;
        sum[0] = __fma_rn(fun, w[0], sum[0]);
        sum[1] = __fma_rn(fun, w[1], sum[1]);
        sum[2] = __fma_rn(fun, w[2], sum[2]);
        sum[3] = __fma_rn(fun, w[3], sum[3]);
        sum[4] = __fma_rn(fun, w[4], sum[4]);
      }
    }
  }
  __syncwarp();
  region_epilogue<NDIM>(P, reg, lane, s_f, s_scale, s_norm, sa, sb, vol, max_dim, ratio, one_minus_ratio);
}

// The point loop is bound by shared-memory wavefronts (a 64-bit LDS of a warp costs two, a 128-bit
// one four, whatever the addresses), so everything that is not a genuinely per-lane gather stays
// out of shared memory:
//   * code word of a point: one byte per axis, 8 bytes up to 8-D (one LDS.64);
//   * rule weights: 32 consecutive points almost always belong to ONE generator set, whose five
//     weights are then read from the kernel's constant parameter bank with a warp-uniform index
//     (no LSU traffic); only the few half-chunks that straddle a set boundary look the set up per
//     lane.
// VAR: bit 0 = unit Jacobian (no f * 1.0), bit 1 = the functor's arguments are in range over the whole volume
// (batch_nocheck instead of batch_fast: no range checks, no vote, no fallback)
template <typename F, int NDIM, bool BY_VALUE, int VAR = 0>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, EvalTuning<F>::kMinBlocks)
evaluate_kernel_bp(const __grid_constant__ FunctorArg<F, BY_VALUE> farg, const __grid_constant__ EvalParams P)
{
  static_assert(NDIM >= 2 && NDIM <= 12, "byte-permute variant keeps its code table in shared memory");
  using L = BpLayout<NDIM>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t raw_addr = static_cast<uint32_t>(__cvta_generic_to_shared(smem_raw));
  const uint32_t pad = (256u - (raw_addr & 255u)) & 255u;
  unsigned char* sm = smem_raw + pad;      // 256-byte aligned, still a shared-space pointer
  const uint32_t sm_addr = raw_addr + pad;

  double* s_scale = reinterpret_cast<double*>(sm + kWarpsPerBlock * L::kWarpBytes);
  double* s_norm = s_scale + PGN_NSETS * PGN_NRULES;
  double* s_tp = s_norm + PGN_NSETS * PGN_NRULES;
  double* s_tm = s_tp + PGN_NCODES;
  double* s_wc = s_tm + PGN_NCODES;  // [9][5] rule weights (EvalTuning<F>::kLaneWeightsSmem)
  uint32_t* s_codes = reinterpret_cast<uint32_t*>(sm + ((L::kCodesOffset + 15) / 16) * 16);
  unsigned char* s_set = reinterpret_cast<unsigned char*>(s_codes + L::kPaddedPoints * L::kCodeWords);

  // stage the tables: 4-bit codes widened to one pre-scaled byte per axis
  for (int i = threadIdx.x; i < PGN_NSETS * PGN_NRULES; i += blockDim.x) {
    s_scale[i] = __ldg(&P.scale[i]);
    s_norm[i] = __ldg(&P.norm[i]);
    s_wc[i] = P.wc[i];
  }
  if (threadIdx.x < PGN_NCODES) {
    s_tp[threadIdx.x] = P.half_plus_g[threadIdx.x];
    s_tm[threadIdx.x] = P.half_minus_g[threadIdx.x];
  }
  constexpr int kFeval = feval_of(NDIM);
  for (int i = threadIdx.x; i < kFeval; i += blockDim.x) {
    const uint64_t c4 = __ldg(&P.point_code[i]);
    uint32_t words[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int d = 0; d < NDIM; ++d)
      words[d >> 2] |= static_cast<uint32_t>(((c4 >> (4 * d)) & 0xF) * 8u) << (8 * (d & 3));
#pragma unroll
    for (int k = 0; k < L::kCodeWords; ++k)
      s_codes[i * L::kCodeWords + k] = words[k];
    s_set[i] = static_cast<unsigned char>(c4 >> 60);
  }
  for (int i = kFeval * L::kCodeWords + threadIdx.x; i < L::kPaddedPoints * L::kCodeWords; i += blockDim.x)
    s_codes[i] = 0u;  // code 0 on every axis: the centre
  __syncthreads();

  const int lane = threadIdx.x & 31;
  // a broadcast from lane 0 is warp-uniform for the compiler's divergence analysis: the region index, the loop over
  // the regions and the warp's shared-memory addresses move to the uniform datapath, and the shuffles of the epilogue
  // are known to be convergent (no divergent twin of every shuffle group, no register moves where the twins rejoin:
  // the kernel shrinks from 2190 to 1290 instructions).  +4 % F_1_8D, +9 % F_1_5D, +8 % F_3_8D, +2..4 % for the 10-D
  // families and F_2_6D, -1.7 % F_4_8D (profiles/r02_evalbench_aa.jsonl): per-functor opt-out kUniformWarp.
  int warp = threadIdx.x >> 5;
  if constexpr (eval_uniform_warp<F>::value)
    warp = __shfl_sync(0xffffffffu, warp, 0);
  unsigned char* warp_area = sm + warp * L::kWarpBytes;
  const uint32_t xbase = sm_addr + warp * L::kWarpBytes;  // 256-aligned
  double* s_x = reinterpret_cast<double*>(warp_area);
  double* s_f = reinterpret_cast<double*>(warp_area + L::kXBytes);
  double* s_lo = s_f + 64;
  double* s_up = s_lo + NDIM;
  double* s_ext = s_up + NDIM;

  F* fptr = farg.get();
  // 32-bit shared-memory addresses of the lane's code word in trip 0 and of the per-point set table
  uint32_t lane_codes =
    sm_addr + static_cast<uint32_t>(((L::kCodesOffset + 15) / 16) * 16) + static_cast<uint32_t>(lane) * (L::kCodeWords * 4);
  uint32_t set_addr =
    sm_addr + static_cast<uint32_t>(((L::kCodesOffset + 15) / 16) * 16) + static_cast<uint32_t>(L::kPaddedPoints * L::kCodeWords * 4);
  // opaque to the compiler: under register pressure it would otherwise re-derive these addresses (shared window
  // base, alignment padding: eight instructions) inside every trip of the point loop instead of keeping a register
  asm volatile("" : "+r"(lane_codes), "+r"(set_addr));

  // Trips: trip t covers points 64t .. 64t+63, the lane's two points are 64t+lane (chain A = reference thread
  // `lane`) and 64t+32+lane (chain B = thread lane+32).  The code table is padded to a whole number of trips
  // (with the centre point), so a gather never needs a bounds check; results of non-existent points are dropped
  // by predication.
  constexpr int kTrips = (kFeval + 63) / 64;
  constexpr int kMiddle = kTrips >= 2 ? kTrips - 2 : 0;       // trips 1 .. kTrips-2: all 64 points exist
  constexpr bool UNIT_JAC = (VAR & 1) != 0;
  constexpr bool kNoCheck = (VAR & 2) != 0;                   // straight-line values, in range by the functor's promise
  constexpr bool kFast = !kNoCheck && has_batch_fast<F, NDIM, 2>::value;  // ... + exact fallback per region
  const double jac = P.jacobian;
  const double ratio = P.ratio;
  const double one_minus_ratio = __dsub_rn(1.0, ratio);

  const size_t n = P.n;
  const size_t stride = static_cast<size_t>(gridDim.x) * kWarpsPerBlock;
  // The bounds of the NEXT region of this warp are loaded while the current one is evaluated: the two global loads
  // (HBM / L2 latency, several hundred cycles) are the only long-latency operations of a region, and with four to
  // five warps per scheduler a warp waiting for them at the top of every region costs issue slots.
  // (PGN_EVAL_PREFETCH_BOUNDS=0 loads them at the top of the region instead: four registers fewer across the loop.)
  double next_lo = 0.0, next_len = 0.0;
  if constexpr (PGN_EVAL_PREFETCH_BOUNDS) {
    const size_t first = static_cast<size_t>(blockIdx.x) * kWarpsPerBlock + warp;
    if (first < n)
      load_bounds<NDIM>(P, first, lane, next_lo, next_len);
  }
  for (size_t reg = static_cast<size_t>(blockIdx.x) * kWarpsPerBlock + warp; reg < n; reg += stride) {
    __syncwarp();
    double vol;
    int max_dim;
    if constexpr (!PGN_EVAL_PREFETCH_BOUNDS)
      load_bounds<NDIM>(P, reg, lane, next_lo, next_len);
    region_prologue<NDIM>(P, next_lo, next_len, lane, s_tp, s_tm, s_lo, s_up, s_ext, s_x, vol, max_dim);
    if constexpr (PGN_EVAL_PREFETCH_BOUNDS) {
      if (reg + stride < n)
        load_bounds<NDIM>(P, reg + stride, lane, next_lo, next_len);
    }

    double sa[PGN_NRULES], sb[PGN_NRULES];
#pragma unroll
    for (int r = 0; r < PGN_NRULES; ++r)
      sa[r] = sb[r] = 0.0;
    bool exact = true;  // kFast: no lane needed the functor's slow path so far

    // coordinates of the lane's two points of trip t (one address computation for both code words)
    auto gather = [&](double (&x)[2][NDIM], int t) {
      constexpr int kPointBytes = L::kCodeWords * 4;
      const uint32_t a = lane_codes + static_cast<uint32_t>(t) * (64u * kPointBytes);
      Words4 ca, cb;
      if constexpr (L::kCodeWords == 2) {
        const uint2 va = lds_u32x2<0>(a), vb = lds_u32x2<32 * kPointBytes>(a);
        ca = Words4{va.x, va.y, 0u, 0u};
        cb = Words4{vb.x, vb.y, 0u, 0u};
      } else {
        const uint4 va = lds_u32x4<0>(a), vb = lds_u32x4<32 * kPointBytes>(a);
        ca = Words4{va.x, va.y, va.z, va.w};
        cb = Words4{vb.x, vb.y, vb.z, vb.w};
      }
      BpGather<NDIM, 0>::run(ca, xbase, x[0]);
      BpGather<NDIM, 0>::run(cb, xbase, x[1]);
    };
    // coordinates of the lane's chain-A point of trip t only
    auto gather_a = [&](double (&xa)[NDIM], int t) {
      constexpr int kPointBytes = L::kCodeWords * 4;
      const uint32_t a = lane_codes + static_cast<uint32_t>(t) * (64u * kPointBytes);
      Words4 ca;
      if constexpr (L::kCodeWords == 2) {
        const uint2 va = lds_u32x2<0>(a);
        ca = Words4{va.x, va.y, 0u, 0u};
      } else {
        const uint4 va = lds_u32x4<0>(a);
        ca = Words4{va.x, va.y, va.z, va.w};
      }
      BpGather<NDIM, 0>::run(ca, xbase, xa);
    };
    // f(x) * Jacobian of both points
    auto values = [&](const double (&x)[2][NDIM], double (&f)[2]) {
      double fv[2];
      if constexpr (kNoCheck)
        static_cast<const F*>(fptr)->template batch_nocheck<2>(x, fv);
      else if constexpr (kFast)
        exact = static_cast<const F*>(fptr)->template batch_fast<2>(x, fv) && exact;
      else
        BatchCall<F, NDIM, 2>::run(fptr, x, fv);
      if constexpr (UNIT_JAC) {
        f[0] = fv[0];
        f[1] = fv[1];
      } else {
        f[0] = __dmul_rn(fv[0], jac);
        f[1] = __dmul_rn(fv[1], jac);
      }
    };
    auto accumulate = [&](double (&sum)[PGN_NRULES], double fun, const double* w) {
      sum[0] = __fma_rn(fun, w[0], sum[0]);
      sum[1] = __fma_rn(fun, w[1], sum[1]);
      sum[2] = __fma_rn(fun, w[2], sum[2]);
      sum[3] = __fma_rn(fun, w[3], sum[3]);
      sum[4] = __fma_rn(fun, w[4], sum[4]);
    };
    // per-lane rule weights of point p in the half-chunk starting at q (uniform for most half-chunks)
    auto lane_weights = [&](int q, int p) -> const double* {
      int set = P.half_set[q >> 5];
      if (set == 0xFF)
        set = lds_u8(set_addr + static_cast<uint32_t>(p < kFeval ? p : 0));
      if constexpr (eval_lane_weights_smem<F>::value)
        return static_cast<const double*>(s_wc) + set * PGN_NRULES;
      else
        return P.wc + set * PGN_NRULES;
    };
    // a trip whose points may belong to different generator sets and may not all exist (first, last)
    auto ragged_trip = [&](const double (&x)[2][NDIM], int t, bool keep_f) {
      const int pa = (t << 6) + lane, pb = pa + 32;
      double f[2];
      values(x, f);
      if (keep_f) {
        s_f[lane] = f[0];
        s_f[lane + 32] = f[1];
      }
      if (pa < kFeval)
        accumulate(sa, f[0], lane_weights(t << 6, pa));
      if (pb < kFeval)
        accumulate(sb, f[1], lane_weights((t << 6) + 32, pb));
    };
    // rule weights of the current uniform run (registers), reloaded when the generator set changes
    double w[PGN_NRULES] = {0.0, 0.0, 0.0, 0.0, 0.0};
    int cur_set = -1;
    auto middle_trip = [&](const double (&x)[2][NDIM], int t) {
      double f[2];
      values(x, f);
      const int st = P.chunk_set[t];
      if (st != 0xFF) {
        if (st != cur_set) {
          cur_set = st;
#pragma unroll
          for (int r = 0; r < PGN_NRULES; ++r)
            w[r] = P.wc[st * PGN_NRULES + r];
        }
        accumulate(sa, f[0], w);
        accumulate(sb, f[1], w);
      } else {
        const int pa = (t << 6) + lane;
        accumulate(sa, f[0], lane_weights(t << 6, pa));
        accumulate(sb, f[1], lane_weights((t << 6) + 32, pa + 32));
      }
    };

#if PGN_EVAL_PIPELINE
    // Software pipeline: the 2 + 2*NDIM shared-memory loads of trip t+1 are issued before the arithmetic of
    // trip t, into a second coordinate buffer.  Measured on B200 (profiles/r02_evalbench_pipeline.jsonl): no gain
    // for the oscillatory family, 5-19 % slower for the register-hungry pow families -- with 4-5 warps per
    // scheduler the loads of one warp already overlap the arithmetic of the others.  Kept as an option.
    double x0[2][NDIM], x1[2][NDIM];
    gather(x0, 0);
    if constexpr (kTrips >= 2)
      gather(x1, 1);
    ragged_trip(x0, 0, true);
    if constexpr (kTrips >= 2) {
      int t = 1;
#pragma unroll 1
      for (; t + 1 <= kMiddle; t += 2) {
        gather(x0, t + 1);
        middle_trip(x1, t);
        gather(x1, t + 2);
        middle_trip(x0, t + 1);
      }
      if constexpr (kMiddle & 1) {  // t == kMiddle, odd: its coordinates are in x1
        gather(x0, kMiddle + 1);
        middle_trip(x1, kMiddle);
      }
      if constexpr (((kTrips - 1) & 1) != 0)
        ragged_trip(x1, kTrips - 1, false);
      else
        ragged_trip(x0, kTrips - 1, false);
    }
#else
    {
      double x[2][NDIM];
      // trip 0 keeps its function values for the fourth differences (Sample.cuh:176-219)
      gather(x, 0);
      ragged_trip(x, 0, true);
      if constexpr (kTrips >= 2) {
        if constexpr (eval_segments<F>::value) {
        // trips 1 .. kTrips-2 as compile-time runs (TripPlan): inside a single-set run the five weights are
        // constant-bank operands and a trip is gather + values + 10 FMAs, nothing else
        auto segment = [&](auto sc) {
          constexpr int S = decltype(sc)::value;
          constexpr int kFirst = kTripPlan<NDIM>.first[S], kCount = kTripPlan<NDIM>.count[S];
          constexpr int kSet = kTripPlan<NDIM>.set[S];
#pragma unroll 1
          for (int t = kFirst; t < kFirst + kCount; ++t) {
            gather(x, t);
            double f[2];
            values(x, f);
            if constexpr (kSet >= 0) {
              accumulate(sa, f[0], P.wc + kSet * PGN_NRULES);
              accumulate(sb, f[1], P.wc + kSet * PGN_NRULES);
            } else {
              const int pa = (t << 6) + lane;
              accumulate(sa, f[0], lane_weights(t << 6, pa));
              accumulate(sb, f[1], lane_weights((t << 6) + 32, pa + 32));
            }
          }
        };
        for_constexpr(segment, std::make_integer_sequence<int, kTripPlan<NDIM>.nseg>{});
        } else {
#pragma unroll 1
          for (int t = 1; t <= kMiddle; ++t) {
            gather(x, t);
            middle_trip(x, t);
          }
        }
        // last trip (Sample.cuh:238-252): only some of its points exist
        constexpr int kTail = kFeval - 64 * (kTrips - 1);
        if constexpr (PGN_EVAL_TAIL_A_ONLY && kTail <= 32) {
          // none of the points 64 (kTrips-1) + 32 + lane exists: chain B is complete, only chain A has a last term
          double xa[1][NDIM], fv[1];
          gather_a(xa[0], kTrips - 1);
          if constexpr (kNoCheck)
            static_cast<const F*>(fptr)->template batch_nocheck<1>(xa, fv);
          else if constexpr (has_batch_fast<F, NDIM, 1>::value)
            exact = static_cast<const F*>(fptr)->template batch_fast<1>(xa, fv) && exact;
          else
            BatchCall<F, NDIM, 1>::run(fptr, xa, fv);
          const int pa = ((kTrips - 1) << 6) + lane;
          if (pa < kFeval)
            accumulate(sa, UNIT_JAC ? fv[0] : __dmul_rn(fv[0], jac), lane_weights((kTrips - 1) << 6, pa));
        } else {
          gather(x, kTrips - 1);
          ragged_trip(x, kTrips - 1, false);
        }
      }
    }
#endif
    __syncwarp();
    if constexpr (kFast) {
      if (!__all_sync(0xffffffffu, exact)) {
        // some lane's argument needs the functor's slow path: redo the region with operator() (cold)
        evaluate_region_exact<F, NDIM>(fptr, P, reg, lane, lane_codes, set_addr, xbase, s_f, s_scale, s_norm, vol,
                                       max_dim, jac, ratio, one_minus_ratio);
        continue;
      }
    }
    region_epilogue<NDIM>(P, reg, lane, s_f, s_scale, s_norm, sa, sb, vol, max_dim, ratio,
                          one_minus_ratio);
  }
}

template <auto Kernel>
struct KernelTag {
  static constexpr auto value = Kernel;
};

// One instantiation's launch logic (called through the pgn_eval_fn trampoline).  `functor` is the
// host object when the functor travels by value (nullptr: default-construct it, which is what the
// pre-instantiated stateless integrands do) and the managed-memory copy otherwise.
template <typename F, int NDIM>
int
launch_evaluate(F* functor, const pgn_region_view* v)
{
  constexpr bool kByValue = pass_by_value<F>::value;
  if (v == nullptr || v->abi_version != PGN_ABI_VERSION || v->view_bytes != sizeof(pgn_region_view)) {
    std::fprintf(stderr, "pagani_b200: evaluator and libpagani_b200.so were built from different headers\n");
    return PGN_ERR_INVALID;
  }
  if (v->rule.ndim != NDIM || v->rule.feval != feval_of(NDIM))
    return PGN_ERR_INVALID;
  if (v->n == 0)
    return PGN_OK;
  EvalParams P{};
  P.left = v->left;
  P.length = v->length;
  P.n = v->n;
  P.estimate = v->estimate;
  P.errorest = v->errorest;
  P.axis = v->axis;
  P.point_code = v->rule.point_code;
  P.weights = v->rule.weights;
  P.scale = v->rule.scale;
  P.norm = v->rule.norm;
  for (int d = 0; d <= PGN_MAX_NDIM; ++d) {
    P.lows[d] = v->lows[d];
    P.ranges[d] = v->ranges[d];
  }
  for (int c = 0; c < PGN_NCODES; ++c) {
    P.half_plus_g[c] = v->rule.half_plus_g[c];
    P.half_minus_g[c] = v->rule.half_minus_g[c];
  }
  for (int i = 0; i < PGN_NSETS * PGN_NRULES; ++i)
    P.wc[i] = v->rule.weights_host[i];
  if constexpr (NDIM <= 12) {
    // generator set of every 32-point half-chunk / 64-point chunk (0xFF: mixed): a function of NDIM only, built once
    struct SetTables {
      uint8_t half_set[256], chunk_set[128];
      SetTables()
      {
        constexpr int kFeval = feval_of(NDIM);
        static_assert((kFeval + 31) / 32 <= 256, "half_set table too small");
        auto host_set_of = [](int p) {
          int st = 0;
          for (int k = 1; k < PGN_NSETS; ++k)
            st += p >= set_begin(NDIM, k) ? 1 : 0;
          return st;
        };
        for (int j = 0; j < 256; ++j) {
          const int first = 32 * j, last = first + 31 < kFeval - 1 ? first + 31 : kFeval - 1;
          half_set[j] = (first < kFeval && host_set_of(first) == host_set_of(last))
                          ? static_cast<uint8_t>(host_set_of(first))
                          : static_cast<uint8_t>(0xFF);
        }
        for (int k = 0; k < 128; ++k)
          chunk_set[k] = (half_set[2 * k] != 0xFF && half_set[2 * k] == half_set[2 * k + 1]) ? half_set[2 * k]
                                                                                             : static_cast<uint8_t>(0xFF);
      }
    };
    static const SetTables tables;
    std::memcpy(P.half_set, tables.half_set, sizeof P.half_set);
    std::memcpy(P.chunk_set, tables.chunk_set, sizeof P.chunk_set);
  }
  P.jacobian = v->jacobian;
  P.ratio = v->rule.ratio;
  P.feval = v->rule.feval;
  cudaStream_t stream = static_cast<cudaStream_t>(v->stream);

  FunctorArg<F, kByValue> farg = [&]() {
    if constexpr (kByValue) {
      if constexpr (std::is_default_constructible<F>::value) {
        if (functor == nullptr)
          return FunctorArg<F, true>{F{}};
      }
      return FunctorArg<F, true>{*functor};
    } else {
      return FunctorArg<F, false>{functor};
    }
  }();

  // the kernel is a template ARGUMENT of the tag: the occupancy / attribute cache below is per kernel
  // (the two point-loop variants have the same function type, so `auto kernel` alone would share it)
  auto launch = [&](auto tag, size_t smem) -> int {
    constexpr auto kernel = decltype(tag)::value;
    // the shared-memory attribute is per DEVICE: one cache slot per device ordinal (0: not set up yet); atomics,
    // because hosts driving different engines may arrive here concurrently (the set-up itself is idempotent)
    constexpr int kMaxDevices = 64;
    static std::atomic<int> cache[kMaxDevices];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices)
      return PGN_ERR_CUDA;
    int blocks_per_sm = cache[dev].load(std::memory_order_acquire);
    if (blocks_per_sm == 0) {
      if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(smem)) != cudaSuccess)
        return PGN_ERR_CUDA;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, kWarpsPerBlock * 32,
                                                        smem) != cudaSuccess ||
          blocks_per_sm < 1)
        return PGN_ERR_CUDA;
      cache[dev].store(blocks_per_sm, std::memory_order_release);
    }
    const size_t want = (v->n + kWarpsPerBlock - 1) / kWarpsPerBlock;
    const size_t cap = static_cast<size_t>(v->sm_count > 0 ? v->sm_count : 148) * blocks_per_sm;
    const unsigned grid = static_cast<unsigned>(want < cap ? want : cap);
    kernel<<<grid, kWarpsPerBlock * 32, smem, stream>>>(farg, P);
    // peek: the engine reads (and clears) the error to report its text
    return cudaPeekAtLastError() == cudaSuccess ? PGN_OK : PGN_ERR_CUDA;
  };

  constexpr bool kCodesFit = smem_bytes<NDIM>(true) <= 100 * 1024;
  if constexpr (NDIM <= 12) {
    if (v->reserved != kVariantShift) {
      if constexpr (eval_unit_jacobian_variant<F>::value) {
        if (v->jacobian == 1.0) {
          if constexpr (PGN_EVAL_NOCHECK != 0 && has_batch_nocheck<F, NDIM, 2>::value) {
            // the functor's promise for the whole integration volume (every region lies inside it)
            double lo[NDIM], hi[NDIM];
            for (int d = 0; d < NDIM; ++d) {
              lo[d] = v->lows[d];
              hi[d] = v->lows[d] + v->ranges[d];
            }
            bool in_range = false;
            if constexpr (std::is_default_constructible<F>::value) {
              in_range = functor != nullptr ? functor->fast_in_box(lo, hi) : F{}.fast_in_box(lo, hi);
            } else {
              in_range = functor != nullptr && functor->fast_in_box(lo, hi);
            }
            if (in_range)
              return launch(KernelTag<evaluate_kernel_bp<F, NDIM, kByValue, 3>>{}, BpLayout<NDIM>::bytes());
          }
          return launch(KernelTag<evaluate_kernel_bp<F, NDIM, kByValue, 1>>{}, BpLayout<NDIM>::bytes());
        }
      }
      return launch(KernelTag<evaluate_kernel_bp<F, NDIM, kByValue>>{}, BpLayout<NDIM>::bytes());
    }
  }
  return launch(KernelTag<evaluate_kernel<F, NDIM, kByValue, kCodesFit>>{}, smem_bytes<NDIM>(kCodesFit));
}

template <typename F, int NDIM>
int
trampoline(void* user, const pgn_region_view* v)
{
  return launch_evaluate<F, NDIM>(static_cast<F*>(user), v);
}

}  // namespace detail

// Owns the device-visible copy of a functor for the duration of one integrate call.
// Mirrors quad::cuda_copy_to_managed (copy-construct into managed memory) and the explicit
// destructor call + cudaFree at the end of Workspace::integrate (Workspace.cuh:451-452).
template <typename F, int NDIM>
class Evaluator {
public:
  explicit Evaluator(const F& f)
  {
    if constexpr (detail::pass_by_value<F>::value) {
      copy_ = static_cast<F*>(::operator new(sizeof(F)));  // host copy, handed to the launch by value
      new (copy_) F(f);
    } else {
      void* mem = nullptr;
      if (cudaMallocManaged(&mem, sizeof(F)) != cudaSuccess)
        throw std::bad_alloc();
      copy_ = new (mem) F(f);
      int dev = 0;
      cudaGetDevice(&dev);
      cudaMemPrefetchAsync(copy_, sizeof(F), dev, 0);
    }
  }
  ~Evaluator()
  {
    if (copy_ == nullptr)
      return;
    if constexpr (detail::pass_by_value<F>::value) {
      copy_->~F();
      ::operator delete(copy_);
    } else {
      cudaDeviceSynchronize();
      copy_->~F();
      cudaFree(copy_);
    }
  }
  Evaluator(const Evaluator&) = delete;
  Evaluator& operator=(const Evaluator&) = delete;

  pgn_eval_fn fn() const { return &detail::trampoline<F, NDIM>; }
  void* user() const { return copy_; }

private:
  F* copy_ = nullptr;
};

}  // namespace pagani_b200
#endif
