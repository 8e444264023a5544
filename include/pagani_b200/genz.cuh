// The six Genz test-integrand families in the parameterisation the reference benchmarks use
// (common/cuda/integrands.cuh:1303-1932: F_1 oscillatory, F_2 product peak, F_3 corner peak,
// F_4 Gaussian, F_5 C0, F_6 discontinuous at 5..8 dimensions) plus the 10-dimensional F_5/F_6
// this project adds for BASELINE config 4.
//
// Every functor is a stateless callable taking ndim doubles.  The floating-point expression trees
// (association order, literal constants, pow(x, 2.) kept as a genuine pow call) are those of the
// mathematical definitions as the reference writes them, because region-trajectory parity needs
// f(x) bit-identical: symmetric integrands tie their fourth differences and rounding decides the
// split axis.  tests/ pin this against the reference binary on the GPU.
//
// The same header is compiled into the reference oracle harness (oracle/ref/ref_pagani.cu) for the
// 10-D functors the reference does not define.
#ifndef PAGANI_B200_GENZ_CUH
#define PAGANI_B200_GENZ_CUH

#include <cmath>

#include "fastmath.cuh"

// PGN_GENZ_LIBRARY_MATH: call the CUDA math library directly (what the reference's functors do; used by
// the machine-code parity test); default: the bit-identical, load-free re-expressions of fastmath.cuh
#ifdef PGN_GENZ_LIBRARY_MATH
#define PGN_GENZ_COS(x) cos(x)
#else
#define PGN_GENZ_COS(x) ::pagani_b200::fastmath::cos_pinned(x)
#endif

// PGN_GENZ_FAST_POW (default 1): the squares of the product-peak and Gaussian families through
// fastmath::pow2_pinned_n -- the math library's pow(x, 2.) operation sequence as straight-line code, bit-identical
// to pow(x, 2.) (tests/cpp/test_fastmath.cu) and 21-26 % faster in the evaluation kernel (profiles/r01_pow_sweep.jsonl).
#ifndef PGN_GENZ_FAST_POW
#define PGN_GENZ_FAST_POW 1
#endif
// PGN_GENZ_FAST_EXP (default 1): exp() of the Gaussian, C0 and discontinuous families through fastmath::exp_pinned
// (bit-identical to exp() on the GPU, tests/cpp/test_fastmath.cu; the C0 / discontinuous / Gaussian trajectories were
// re-run against the reference's goldens with it in round 2: profiles/r02_pytest_gpu_sessionC.txt; F_5_10D +8 %)
#ifndef PGN_GENZ_FAST_EXP
#define PGN_GENZ_FAST_EXP 1
#endif
#if PGN_GENZ_FAST_EXP && !defined(PGN_GENZ_LIBRARY_MATH)
#define PGN_GENZ_EXP(x) ::pagani_b200::fastmath::exp_pinned(x)
#else
#define PGN_GENZ_EXP(x) exp(x)
#endif

#ifdef __CUDACC__
#define PGN_HD __host__ __device__
#else
#define PGN_HD
#endif

namespace pagani_b200 {
namespace genz {

// explicitly rounded product / fused multiply-add, immune to the compiler's contraction choices
PGN_HD inline double
mul_rn(double a, double b)
{
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
PGN_HD inline double
fma_rn(double a, double b, double c)
{
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}

// F_1: cos(x0 + 2 x1 + 3 x2 + ...)
template <int N>
struct Oscillatory {
  static constexpr int ndim = N;
  template <class... T>
  PGN_HD double
  operator()(T... xs) const
  {
    static_assert(sizeof...(T) == N, "arity");
    const double x[N] = {xs...};
    double s = x[0];
#pragma unroll
    for (int k = 1; k < N; ++k)
      s = s + static_cast<double>(k + 1) * x[k];
    return PGN_GENZ_COS(s);
  }
#if defined(__CUDACC__) && !defined(PGN_GENZ_LIBRARY_MATH) && !defined(PGN_GENZ_NO_BATCH)
  // M points at once (evaluate.cuh: has_batch): the same values as M calls, one shared cold branch
  template <int M>
  __device__ void
  batch(const double (&x)[M][N], double (&f)[M]) const
  {
    double s[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      s[m] = x[m][0];
#pragma unroll
      for (int k = 1; k < N; ++k)
        s[m] = s[m] + static_cast<double>(k + 1) * x[m][k];
    }
    ::pagani_b200::fastmath::cos_pinned_n<M>(s, f);
  }
  // straight-line form (evaluate.cuh: has_batch_fast): false when an argument needs cos()'s slow path
  template <int M>
  __device__ bool
  batch_fast(const double (&x)[M][N], double (&f)[M]) const
  {
    bool ok = true;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      double s = x[m][0];
#pragma unroll
      for (int k = 1; k < N; ++k)
        s = s + static_cast<double>(k + 1) * x[m][k];
      bool okc;
      f[m] = ::pagani_b200::fastmath::cos_fast(s, &okc);
      ok = ok && okc;
    }
    return ok;
  }
  // launch-time promise (evaluate.cuh: has_batch_nocheck): over the box [lo, hi] the cosine's argument stays far
  // inside the fast path's range |s| < 2^31 (the bound leaves room for the rounding of the sum)
  PGN_HD bool
  fast_in_box(const double* lo, const double* hi) const
  {
    double bound = 0.;
    for (int k = 0; k < N; ++k) {
      const double a = lo[k] < 0. ? -lo[k] : lo[k], b = hi[k] < 0. ? -hi[k] : hi[k];
      bound += static_cast<double>(k + 1) * (a > b ? a : b);
    }
    return bound < 1073741824.;  // false for NaN bounds as well
  }
  template <int M>
  __device__ void
  batch_nocheck(const double (&x)[M][N], double (&f)[M]) const
  {
#pragma unroll
    for (int m = 0; m < M; ++m) {
      double s = x[m][0];
#pragma unroll
      for (int k = 1; k < N; ++k)
        s = s + static_cast<double>(k + 1) * x[m][k];
      f[m] = ::pagani_b200::fastmath::cos_in_range(s);
    }
  }
#endif
};

// F_2: prod_k 1 / (a^-2 + (x_k - b)^2),  a = 50, b = .5
template <int N>
struct ProductPeak {
  static constexpr int ndim = N;
  template <class... T>
  PGN_HD double
  operator()(T... xs) const
  {
    static_assert(sizeof...(T) == N, "arity");
    const double x[N] = {xs...};
    const double a = 50.;
    const double b = .5;
#if PGN_GENZ_FAST_POW && !defined(PGN_GENZ_LIBRARY_MATH)
    double d[N], sq[N];
#pragma unroll
    for (int k = 0; k < N; ++k)
      d[k] = x[k] - b;
    ::pagani_b200::fastmath::pow2_pinned_n<N>(d, sq);
    double val = 1. / ((1. / pow(a, 2.)) + sq[0]);
#pragma unroll
    for (int k = 1; k < N; ++k) {
      const double term = 1. / ((1. / pow(a, 2.)) + sq[k]);
      val = val * term;
    }
#else
    double val = 1. / ((1. / pow(a, 2.)) + pow(x[0] - b, 2.));
#pragma unroll
    for (int k = 1; k < N; ++k) {
      const double term = 1. / ((1. / pow(a, 2.)) + pow(x[k] - b, 2.));
      val = val * term;
    }
#endif
    return val;
  }
};

// F_3: (1 + N x_{N-1} + ... + 4 x_3 + 3 x_0 + 2 x_1 + x_2)^-(N+1)
template <int N>
struct CornerPeak {
  static constexpr int ndim = N;
  template <class... T>
  PGN_HD double
  operator()(T... xs) const
  {
    static_assert(sizeof...(T) == N && N >= 4, "arity");
    const double x[N] = {xs...};
    double s = 1.;
#pragma unroll
    for (int c = N; c >= 4; --c)
      s = s + static_cast<double>(c) * x[c - 1];
    s = s + 3. * x[0];
    s = s + 2. * x[1];
    s = s + x[2];
    return pow(s, -static_cast<double>(N + 1));
  }
};

// F_4: exp(-sum_k 25^2 (x_k - .5)^2)
template <int N>
struct Gaussian {
  static constexpr int ndim = N;
  template <class... T>
  PGN_HD double
  operator()(T... xs) const
  {
    static_assert(sizeof...(T) == N, "arity");
    const double x[N] = {xs...};
    const double beta = .5;
#if PGN_GENZ_FAST_POW && !defined(PGN_GENZ_LIBRARY_MATH)
    double d[N], sq[N];
#pragma unroll
    for (int k = 0; k < N; ++k)
      d[k] = x[k] - beta;
    ::pagani_b200::fastmath::pow2_pinned_n<N>(d, sq);
    double s = pow(25., 2.) * sq[0];
#pragma unroll
    for (int k = 1; k < N; ++k)
      s = s + pow(25., 2.) * sq[k];
#else
    double s = pow(25., 2.) * pow(x[0] - beta, 2.);
#pragma unroll
    for (int k = 1; k < N; ++k)
      s = s + pow(25., 2.) * pow(x[k] - beta, 2.);
#endif
    return PGN_GENZ_EXP(-1.0 * s);
  }
  // (no batch_fast form: interleaving the 2 x N pow kernels of two points without a branch between them measured
  // 10-14 % SLOWER than operator() per point -- register pressure -- profiles/r02_evalbench_d.jsonl)
};

// F_5: exp(-10 sum_k |x_k - .5|)
template <int N>
struct CZero {
  static constexpr int ndim = N;
  template <class... T>
  PGN_HD double
  operator()(T... xs) const
  {
    static_assert(sizeof...(T) == N, "arity");
    const double x[N] = {xs...};
    const double beta = .5;
    // one expression in the reference: -10|x0-b| - 10|x1-b| - ... ; nvcc (12.9, sm_100a) rounds the
    // SECOND product and fuses the first, then chains FMAs (checked against the reference's SASS)
    double t1 = mul_rn(-10., fabs(x[1] - beta));
    t1 = fma_rn(-10., fabs(x[0] - beta), t1);
#pragma unroll
    for (int k = 2; k < N; ++k)
      t1 = fma_rn(-10., fabs(x[k] - beta), t1);
    return PGN_GENZ_EXP(t1);
  }
#if defined(__CUDACC__) && PGN_GENZ_FAST_EXP && !defined(PGN_GENZ_LIBRARY_MATH) && !defined(PGN_GENZ_NO_BATCH)
  // M points at once (evaluate.cuh: has_batch): the same values as M calls, one shared cold branch
  template <int M>
  __device__ void
  batch(const double (&x)[M][N], double (&f)[M]) const
  {
    const double beta = .5;
    double t[M];
#pragma unroll
    for (int m = 0; m < M; ++m) {
      t[m] = mul_rn(-10., fabs(x[m][1] - beta));
      t[m] = fma_rn(-10., fabs(x[m][0] - beta), t[m]);
#pragma unroll
      for (int k = 2; k < N; ++k)
        t[m] = fma_rn(-10., fabs(x[m][k] - beta), t[m]);
    }
    ::pagani_b200::fastmath::exp_pinned_n<M>(t, f);
  }
#endif
#if defined(__CUDACC__) && !defined(PGN_GENZ_LIBRARY_MATH) && !defined(PGN_GENZ_NO_BATCH)
  // branch-free form (evaluate.cuh: has_batch_fast): false when exp() would have left its direct range (never on the
  // unit cube: the exponent is >= -5 N); +9 % at 10-D over the batch form with its cold branch
  template <int M>
  __device__ bool
  batch_fast(const double (&x)[M][N], double (&f)[M]) const
  {
    const double beta = .5;
    bool ok = true;
#pragma unroll
    for (int m = 0; m < M; ++m) {
      double t = mul_rn(-10., fabs(x[m][1] - beta));
      t = fma_rn(-10., fabs(x[m][0] - beta), t);
#pragma unroll
      for (int k = 2; k < N; ++k)
        t = fma_rn(-10., fabs(x[m][k] - beta), t);
      bool oke;
      f[m] = ::pagani_b200::fastmath::exp_fast(t, &oke);
      ok = ok && oke;
    }
    return ok;
  }
  // (a batch_nocheck form without the range test of exp -- two integer instructions per point -- measured 0.5-1.5 %
  // SLOWER at 10-D, profiles/r02_evalbench_ab.jsonl: not offered)
#endif
};

// F_6: 0 if any x_k > beta_k, else exp(sum alpha_k x_k).  The reference orders its arguments
// differently at every dimension; the tables below give, per argument, its threshold and its
// coefficient, and `Descending` says whether the exponent is summed from the last argument down
// (8-D, 10-D) or from the first up (5..7-D).
template <int N>
struct DiscontinuousSpec;
template <>
struct DiscontinuousSpec<5> {
  static constexpr bool descending = false;
  PGN_HD static double beta(int i) { constexpr double b[5] = {.8, .7, .6, .5, .4}; return b[i]; }
  PGN_HD static double alpha(int i) { constexpr double a[5] = {9., 8., 7., 6., 5.}; return a[i]; }
};
template <>
struct DiscontinuousSpec<6> {
  static constexpr bool descending = false;
  PGN_HD static double beta(int i) { constexpr double b[6] = {.9, .8, .7, .6, .5, .4}; return b[i]; }
  PGN_HD static double alpha(int i) { constexpr double a[6] = {10., 9., 8., 7., 6., 5.}; return a[i]; }
};
template <>
struct DiscontinuousSpec<7> {
  static constexpr bool descending = false;
  PGN_HD static double beta(int i) { constexpr double b[7] = {.9, .8, .7, .6, .5, .4, .3}; return b[i]; }
  PGN_HD static double alpha(int i) { constexpr double a[7] = {10., 9., 8., 7., 6., 5., 4.}; return a[i]; }
};
template <>
struct DiscontinuousSpec<8> {
  static constexpr bool descending = true;
  PGN_HD static double beta(int i) { constexpr double b[8] = {.2, .3, .4, .5, .6, .7, .8, .9}; return b[i]; }
  PGN_HD static double alpha(int i) { constexpr double a[8] = {3., 4., 5., 6., 7., 8., 9., 10.}; return a[i]; }
};
// our 10-D extension: the literal pattern would reach beta = 0, so the ends are
// squeezed to .15 and .95 instead.
template <>
struct DiscontinuousSpec<10> {
  static constexpr bool descending = true;
  PGN_HD static double beta(int i)
  {
    constexpr double b[10] = {.15, .2, .3, .4, .5, .6, .7, .8, .9, .95};
    return b[i];
  }
  PGN_HD static double alpha(int i)
  {
    constexpr double a[10] = {2., 3., 4., 5., 6., 7., 8., 9., 10., 11.};
    return a[i];
  }
};

template <int N>
struct Discontinuous {
  static constexpr int ndim = N;
  using Spec = DiscontinuousSpec<N>;
  template <class... T>
  PGN_HD double
  operator()(T... xs) const
  {
    static_assert(sizeof...(T) == N, "arity");
    const double x[N] = {xs...};
    bool outside = false;
#pragma unroll
    for (int k = 0; k < N; ++k)
      outside = outside || (x[k] > Spec::beta(k));
    if (outside)
      return 0.;
    // The reference writes the exponent as ONE expression a*x + b*y + ... ; which of the first two
    // products nvcc rounds and which it fuses into the FMA chain is the compiler's choice (12.9,
    // sm_100a: the second product is rounded at 5..7-D, the first at 8-D -- read off the reference
    // kernel's SASS and re-checked by tests/test_integrand_sass.py), so it is pinned explicitly.
    double s;
    if (Spec::descending) {
      s = mul_rn(Spec::alpha(N - 1), x[N - 1]);
#pragma unroll
      for (int k = N - 2; k >= 0; --k)
        s = fma_rn(Spec::alpha(k), x[k], s);
    } else {
      s = mul_rn(Spec::alpha(1), x[1]);
      s = fma_rn(Spec::alpha(0), x[0], s);
#pragma unroll
      for (int k = 2; k < N; ++k)
        s = fma_rn(Spec::alpha(k), x[k], s);
    }
    return PGN_GENZ_EXP(s);
  }
  // (no batch form: most points of this family lie outside the support, and operator() skips their exp entirely --
  // measured 3.4e11 evaluations/s at 10-D against 2.4e11 with an unconditional batched exp, profiles/r02_evalbench_c.jsonl)
};

}  // namespace genz

// Tiny integrands the reference's unit tests use (exposed as "families" 0, 7, 8 of the C ABI).
namespace testfn {

// family 0: RegionSampling.cu:18-46 / Easy_Integrals.cu:14-21 -- a constant
template <int N>
struct Constant {
  static constexpr int ndim = N;
  template <class... T>
  PGN_HD double
  operator()(T...) const
  {
    return 15.37;
  }
};

// family 7: do_integration_from_c.cu:11-17 -- x*y (generalised: product of all coordinates)
template <int N>
struct Product {
  static constexpr int ndim = N;
  template <class... T>
  PGN_HD double
  operator()(T... xs) const
  {
    const double x[N] = {xs...};
    double p = x[0];
#pragma unroll
    for (int k = 1; k < N; ++k)
      p = p * x[k];
    return p;
  }
};

// family 8: NanIntegral.cu:15-23
template <int N>
struct NotANumber {
  static constexpr int ndim = N;
  template <class... T>
  PGN_HD double
  operator()(T...) const
  {
    return NAN;
  }
};

}  // namespace testfn

// Launch-bound tuning of the evaluation kernel for the Genz families (B200 sweep of
// tools/evalbench, profiles/r01_evalbench_*.jsonl, r01_cos_sweep2.jsonl, r01_pow_sweep.jsonl): with the pinned,
// straight-line cos and pow(x, 2.) the oscillatory, Gaussian and product-peak families are fastest with registers
// to spare (4 / 3 / 4 resident blocks), the corner peak (general pow) with 8.
template <typename F>
struct EvalTuning;
#ifndef PGN_EVAL_NO_GENZ_TUNING  // tools/evalbench sweeps the knob through the primary template instead
template <int N>
struct EvalTuning<genz::Oscillatory<N>> {
  // launch bound 5 (at most 102 registers; the unit-cube kernel takes 80, six resident blocks).  With the
  // warp-uniform warp index of evaluate.cuh the cosine literals live in uniform registers whatever the bound, and the
  // smaller allocation wins: 3.007e11 vs 2.98e11 evaluations/s at a bound of 4 (profiles/r02_evalbench_z.jsonl;
  // before that change 4 was ahead, r02_evalbench_j.jsonl)
  static constexpr int kMinBlocks = 5;
  // per-lane rule weights from shared memory: +1.1 % at 8-D, neutral at 5-D (profiles/r02_evalbench_s.jsonl)
  static constexpr bool kLaneWeightsSmem = true;
  // second kernel instantiation for the unit cube, without f * 1.0: +1.6 % (profiles/r02_evalbench_u.jsonl); a third
  // one without the |s| < 2^31 checks when fast_in_box() promises them for the whole volume: +2.7 % (r02_evalbench_v)
  static constexpr bool kUnitJacobianVariant = true;
};
template <int N>
struct EvalTuning<genz::ProductPeak<N>> {
  static constexpr int kMinBlocks = PGN_GENZ_FAST_POW ? 4 : 8;
};
template <int N>
struct EvalTuning<genz::CZero<N>> {
  static constexpr int kMinBlocks = 4;     // the primary template's default
  static constexpr bool kSegments = true;  // +4 % at 10-D (r02_evalbench_h.jsonl)
  // unit-cube kernel without f * 1.0: +3.8 % at 10-D together with the warp-uniform warp index (-5 % without it;
  // profiles/r02_evalbench_aa.jsonl)
  static constexpr bool kUnitJacobianVariant = true;
};
template <int N>
struct EvalTuning<genz::Discontinuous<N>> {
  static constexpr int kMinBlocks = 4;
  static constexpr bool kUnitJacobianVariant = true;  // +0.7 % at 10-D (profiles/r02_evalbench_aa.jsonl)
};
template <int N>
struct EvalTuning<genz::CornerPeak<N>> {
  static constexpr int kMinBlocks = 8;
};
template <int N>
struct EvalTuning<genz::Gaussian<N>> {
  static constexpr int kMinBlocks = PGN_GENZ_FAST_POW ? 3 : 8;
  // warp-uniform warp index: +1.4 % at 5-D, -1.7 % at 8-D (profiles/r02_evalbench_aa.jsonl)
  static constexpr bool kUniformWarp = N < 8;
};
#endif
}  // namespace pagani_b200
#endif
