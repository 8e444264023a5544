// Drop-in test: the reference's own PAGANI tests, re-typed against include/pagani_b200/Workspace.cuh.
// Each case cites the reference test it mirrors.  The integrand functors below are written by the
// caller, exactly as a user of the reference would write them (non-const call operators, data
// members, __device__-only operators ...), which is what the header template has to accept.
//
// Built by __graft_entry__.build() into tests/cpp/_build/test_dropin; run by tests/test_gpu_dropin.py.
// Prints one line per case: "PASS name" / "FAIL name ...", exit code = number of failures.
#include <cmath>
#include <cstdio>
#include <vector>

#include "pagani_b200/Workspace.cuh"

extern "C" int do_integration_from_c(double* res);  // tests/cpp/dropin_c_linkage.cu

static int failures = 0;
#define CHECK(name, cond, ...)                                                                     \
  do {                                                                                             \
    if (cond) {                                                                                    \
      std::printf("PASS %s\n", name);                                                              \
    } else {                                                                                       \
      std::printf("FAIL %s ", name);                                                               \
      std::printf(__VA_ARGS__);                                                                    \
      std::printf("\n");                                                                           \
      ++failures;                                                                                  \
    }                                                                                              \
  } while (0)

// ---- test/cuda/pagani/Easy_Integrals.cu:14-21, 54-66 ---------------------------------------------
struct Fcn {
  double const normalization;
  Fcn(double xlow, double xhigh, double ylow, double yhigh)
    : normalization(1. / ((xhigh - xlow) * (yhigh - ylow)))
  {}
  __device__ double
  operator()(double, double) const
  {
    return normalization * 15.37;
  }
};

// ---- integrands written the way common/cuda/integrands.cuh writes them (a data member, non-const
// operator()): the 6-D product peak and the 8-D C0 function of Easy_Integrals.cu:68-148
class PeakUser6 {
public:
  __device__ __host__ double
  operator()(double x, double y, double z, double w, double v, double u)
  {
    const double a = 50.;
    const double b = .5;
    const double t1 = 1. / ((1. / pow(a, 2.)) + pow(x - b, 2.));
    const double t2 = 1. / ((1. / pow(a, 2.)) + pow(y - b, 2.));
    const double t3 = 1. / ((1. / pow(a, 2.)) + pow(z - b, 2.));
    const double t4 = 1. / ((1. / pow(a, 2.)) + pow(w - b, 2.));
    const double t5 = 1. / ((1. / pow(a, 2.)) + pow(v - b, 2.));
    const double t6 = 1. / ((1. / pow(a, 2.)) + pow(u - b, 2.));
    return t1 * t2 * t3 * t4 * t5 * t6;
  }
  double true_value = 1.286889807581113e+13;
};

class CZeroUser8 {
public:
  __device__ __host__ double
  operator()(double x, double y, double z, double k, double m, double n, double p, double q)
  {
    double beta = .5;
    double t1 = -10. * fabs(x - beta) - 10. * fabs(y - beta) - 10. * fabs(z - beta) -
                10. * fabs(k - beta) - 10. * fabs(m - beta) - 10. * fabs(n - beta) -
                10. * fabs(p - beta) - 10. * fabs(q - beta);
    return exp(t1);
  }
  double true_value = 2.425217625641885e-06;
};

// ---- test/cuda/pagani/BoxIntegrals.cu:16-31 (functor of demos/function.cuh:113-132) -----------------
class BoxIntegral8_15 {
public:
  __device__ __host__ double
  operator()(double x, double y, double z, double k, double l, double m, double n, double o)
  {
    double s = 15;
    double sum = 0;
    sum = pow(x, 2) + pow(y, 2) + pow(z, 2) + pow(k, 2) + pow(l, 2) + pow(m, 2) + pow(n, 2) + pow(o, 2);
    return pow(sum, s / 2);
  }
};

// ---- test/cuda/pagani/accuracy_improves_with_epsrel.cu:16-24, 38-67 ---------------------------------
struct Fun6 {
  __device__ __host__ double
  operator()(double u, double v, double w, double x, double y, double z)
  {
    return (12.0 / (7.0 - 6 * log(2.0) * log(2.0) + log(64.0))) *
           (u * v + (pow(w, y) * x * y) / (1 + u) + z * z);
  }
};

// ---- test/cuda/pagani/NanIntegral.cu:15-37 ----------------------------------------------------------
class NaN_Integral {
public:
  __device__ double
  operator()(double, double)
  {
    return NAN;
  }
};

// ---- test/cuda/common/StatefulIntegrand.cu: a functor that owns state --------------------------------
// (a) trivially copyable state in an array member; (b) a non-trivially-copyable functor whose copy
// constructor deep-copies into managed memory, like gpu::cudaDynamicArray -- it must go through the
// managed-memory path (cuda_copy_to_managed semantics).
struct ArrayState {
  double coeff[5];
  __device__ double
  operator()(double x, double y) const
  {
    return coeff[0] + coeff[1] * x + coeff[2] * y + coeff[3] * x * y + coeff[4] * x * x;
  }
};

struct HeapState {
  double* data = nullptr;
  size_t n = 0;
  HeapState(const double* src, size_t count) : n(count)
  {
    cudaMallocManaged(&data, sizeof(double) * n);
    for (size_t i = 0; i < n; ++i)
      data[i] = src[i];
  }
  HeapState(const HeapState& o) : n(o.n)
  {
    cudaMallocManaged(&data, sizeof(double) * n);
    for (size_t i = 0; i < n; ++i)
      data[i] = o.data[i];
  }
  HeapState& operator=(const HeapState&) = delete;
  ~HeapState() { cudaFree(data); }
  __device__ double
  operator()(double x, double y) const
  {
    return data[0] + data[1] * x + data[2] * y + data[3] * x * y + data[4] * x * x;
  }
};

int
main()
{
  {  // Easy_Integrals.cu "Constant Positive Value Function"
    constexpr int ndim = 2;
    Fcn integrand(0, 1, 0, 1);
    Workspace<double, ndim> pagani;
    quad::Volume<double, ndim> vol;
    auto res = pagani.integrate<Fcn>(integrand, 1.0e-3, 1.0e-12, vol);
    CHECK("easy_constant_2d", res.status == 0 && std::fabs(res.estimate - 15.37) <= 1e-3 * 15.37,
          "estimate=%.17g status=%d", res.estimate, res.status);
    CHECK("easy_constant_2d_neval", res.neval > 0 && res.nregions > 0, "neval=%zu", res.neval);
  }
  {  // Easy_Integrals.cu "F_2_6D" and "F_5_8D" shapes with user-written functors
    PeakUser6 f2;
    Workspace<double, 6> w6;
    quad::Volume<double, 6> v6;
    auto r2 = w6.integrate<PeakUser6>(f2, 1.0e-3, 1.0e-12, v6);
    CHECK("easy_product_peak_6d", r2.status == 0 && std::fabs(r2.estimate - f2.true_value) <= 1e-3 * f2.true_value,
          "estimate=%.17g status=%d", r2.estimate, r2.status);
    CZeroUser8 f5;
    Workspace<double, 8> w8;
    quad::Volume<double, 8> v8;
    auto r5 = w8.integrate<CZeroUser8>(f5, 1.0e-3, 1.0e-12, v8);
    CHECK("easy_c0_8d", r5.status == 0 && std::fabs(r5.estimate - f5.true_value) <= 1e-3 * f5.true_value,
          "estimate=%.17g status=%d", r5.estimate, r5.status);
  }
  {  // BoxIntegrals.cu "BoxIntegral8_15"
    BoxIntegral8_15 f;
    Workspace<double, 8> w;
    quad::Volume<double, 8> vol;
    auto r = w.integrate<BoxIntegral8_15>(f, 1.0e-3, 1.0e-12, vol);
    const double truth = 8879.851175413485;
    CHECK("box_integral_8_15", r.status == 0 && std::fabs(r.estimate - truth) / truth <= 1e-3,
          "estimate=%.17g status=%d", r.estimate, r.status);
  }
  {  // accuracy_improves_with_epsrel.cu
    Fun6 f;
    Workspace<double, 6> w;
    quad::Volume<double, 6> vol;
    double epsrel = 1.0e-3, previous = 1.0e-2;
    bool ok = true;
    while (epsrel > 1.0e-6) {
      auto r = w.integrate<Fun6>(f, epsrel, 1.0e-12, vol);
      ok = ok && r.status == 0 && r.errorest <= previous;
      previous = r.errorest;
      epsrel /= 2.0;
    }
    CHECK("accuracy_improves_with_epsrel", ok, "last errorest=%.3g", previous);
  }
  {  // NanIntegral.cu
    NaN_Integral f;
    Workspace<double, 2> w;
    quad::Volume<double, 2> vol;
    auto r = w.integrate<NaN_Integral>(f, 1.0e-3, 1e-12, vol);
    CHECK("nan_integrand", std::isnan(r.estimate), "estimate=%.17g", r.estimate);
  }
  {  // StatefulIntegrand.cu: by-value state and managed-memory (deep copy) state must agree
    const double c[5] = {1.0, 2.0, 3.0, 4.0, 5.0};
    ArrayState a;
    for (int i = 0; i < 5; ++i)
      a.coeff[i] = c[i];
    HeapState h(c, 5);
    Workspace<double, 2> w;
    quad::Volume<double, 2> vol;
    auto ra = w.integrate<ArrayState>(a, 1.0e-8, 1e-40, vol);
    auto rh = w.integrate<HeapState>(h, 1.0e-8, 1e-40, vol);
    const double truth = 1.0 + 1.0 + 1.5 + 1.0 + 5.0 / 3.0;
    CHECK("stateful_by_value", ra.status == 0 && std::fabs(ra.estimate - truth) <= 1e-8 * truth,
          "estimate=%.17g", ra.estimate);
    CHECK("stateful_managed_copy", rh.status == 0 && rh.estimate == ra.estimate && rh.nregions == ra.nregions,
          "managed=%.17g value=%.17g", rh.estimate, ra.estimate);
  }
  {  // non-unit volume (quad::Volume with explicit bounds, BoxIntegrals-style callers)
    Fcn integrand(0, 2, 1, 4);
    Workspace<double, 2> w;
    double lows[2] = {0., 1.}, highs[2] = {2., 4.};
    quad::Volume<double, 2> vol(lows, highs);
    auto r = w.integrate<Fcn>(integrand, 1.0e-6, 1e-40, vol);
    CHECK("non_unit_volume", r.status == 0 && std::fabs(r.estimate - 15.37) <= 1e-6 * 15.37,
          "estimate=%.17g", r.estimate);
  }
  {  // the Sub_regions& overload (Workspace.cuh:148-358): caller-supplied uniform grid
    PeakUser6 f2;
    Workspace<double, 6> w;
    quad::Volume<double, 6> vol;
    Sub_regions<double, 6> regs(2);
    auto r1 = w.integrate<PeakUser6>(f2, regs, 1.0e-3, 1.0e-12, vol);
    auto r2 = w.integrate<PeakUser6>(f2, 1.0e-3, 1.0e-12, vol);
    CHECK("subregions_overload", r1.status == 0 && r1.estimate == r2.estimate && r1.nregions == r2.nregions &&
                                   r1.iters == r2.iters + 1,
          "est %.17g vs %.17g, iters %zu vs %zu", r1.estimate, r2.estimate, r1.iters, r2.iters);
  }
  {  // pagani_loop_test.cu:17-32 (many integrate calls on one Workspace; shortened)
    Fcn integrand(0, 1, 0, 1);
    Workspace<double, 2> w;
    quad::Volume<double, 2> vol;
    bool ok = true;
    double first = 0;
    for (int i = 0; i < 200; ++i) {
      auto r = w.integrate<Fcn>(integrand, 1.0e-3, 1.0e-12, vol);
      if (i == 0)
        first = r.estimate;
      ok = ok && r.status == 0 && r.estimate == first;
    }
    CHECK("workspace_reuse_loop", ok, "first=%.17g", first);
  }
  {  // test_integration_from_c.c:5-23
    double result = 0.0;
    const int rc = do_integration_from_c(&result);
    CHECK("c_linkage", rc == 0 && std::fabs(result - 0.25) / 0.25 <= 1.0e-6, "rc=%d result=%.17g", rc, result);
    CHECK("c_linkage_null", do_integration_from_c(nullptr) == 1, "null pointer must return 1");
  }
  std::printf("%d failure(s)\n", failures);
  return failures;
}
