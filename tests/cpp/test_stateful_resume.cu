// TEST: SURVEY 8(f-2) stateful integrands and 8(f-3) snapshot / resume, through the drop-in Workspace::integrate.
//
//  * functors whose state lives in managed memory behind a copy constructor -- the reference's own
//    gpu::cudaArray / gpu::cudaDynamicArray (common/cuda/cudaArray.cuh, used as in
//    test/cuda/common/StatefulIntegrand.cu:34-98) and its 1-D interpolator (common/cuda/Interp1D.cuh, the shape of the
//    cosmology integrands) -- integrated with the B200-native engine;
//  * Sub_regions::take_snapshot / load_snapshot (Sub_regions.cuh:219-239) and the Sub_regions& overload
//    (Workspace.cuh:148-358, :437-441): a run restarted in a NEW Workspace from a snapshot of an intermediate region
//    list ends with the same list and the same result as the run it was taken from.
// Compiled by tests/cpp/build_reference_tests.sh (needs the reference's utility headers), run by
// tests/test_gpu_reference_tests.py.
#include <array>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <string>
#include <vector>

#include "cuda/pagani/quad/GPUquad/Workspace.cuh"
#include "common/cuda/cudaArray.cuh"
#include "common/cuda/Interp1D.cuh"

static int failures = 0;
#define CHECK(name, cond, ...)                                                                     \
  do {                                                                                             \
    const bool ok_ = (cond);                                                                       \
    std::printf("[%s] %s  ", ok_ ? "ok" : "FAIL", name);                                           \
    std::printf(__VA_ARGS__);                                                                      \
    std::printf("\n");                                                                             \
    failures += ok_ ? 0 : 1;                                                                       \
  } while (0)

// StatefulIntegrand.cu:34-47, 66-80: d0 + x d1 + y d2 + x y d3
struct ToyModelWithHeapArray {
  ToyModelWithHeapArray(int const* d, size_t s) : data(d, s) {}
  ToyModelWithHeapArray() = default;
  __host__ __device__ double
  operator()(double x, double y) const
  {
    return data[0] + x * data[1] + y * data[2] + x * y * data[3];
  }
  gpu::cudaDynamicArray<int> data;
};
struct ToyIntegrandWithHeapModel {
  explicit ToyIntegrandWithHeapModel(int* const modeldata, size_t s) : model(modeldata, s) {}
  __host__ __device__ double
  operator()(double x, double y) const
  {
    return model(x, y);
  }
  ToyModelWithHeapArray model;
};
struct ToyModel {
  explicit ToyModel(int const* d) { data.Initialize(d); }
  ToyModel() = default;
  __host__ __device__ double
  operator()(double x, double y) const
  {
    return data[0] + x * data[1] + y * data[2] + x * y * data[3];
  }
  gpu::cudaArray<int, 5> data;
};
struct ToyIntegrand {
  explicit ToyIntegrand(std::array<int, 5> const& modeldata) : model(modeldata.data()) {}
  __host__ __device__ double
  operator()(double x, double y) const
  {
    return model(x, y);
  }
  ToyModel model;
};
// an interpolated 1-D profile times a polynomial: the shape of the cosmology integrands (managed-memory tables)
struct InterpIntegrand {
  InterpIntegrand(std::array<double, 9> const& xs, std::array<double, 9> const& ys) : table(xs, ys) {}
  __device__ double
  operator()(double x, double y, double z) const
  {
    return table(x) * (1. + y * z);
  }
  quad::Interp1D table;
};

struct Peak6 {
  __device__ double
  operator()(double a, double b, double c, double d, double e, double f) const
  {
    const double s = (a - .3) * (a - .3) + (b - .4) * (b - .4) + (c - .5) * (c - .5) + (d - .6) * (d - .6) +
                     (e - .7) * (e - .7) + (f - .2) * (f - .2);
    return 1. / (.01 + s);
  }
};

template <size_t ndim>
static std::vector<double>
host_copy(const Sub_regions<double, ndim>& r)
{
  std::vector<double> v(2 * r.size * ndim);
  cudaMemcpy(v.data(), r.dLeftCoord, sizeof(double) * r.size * ndim, cudaMemcpyDeviceToHost);
  cudaMemcpy(v.data() + r.size * ndim, r.dLength, sizeof(double) * r.size * ndim, cudaMemcpyDeviceToHost);
  return v;
}

// (a function of its own: nvcc 12.9's front end prints quad::Volume<double, 6> inside main() as
// "Volume< double, unsigned long((6))>", which the host compiler rejects)
using Vol6 = quad::Volume<double, 6>;
static void
recorder_case()
{
  // debug recorders (Workspace.cuh:171-242): cuda_pagani_iters.csv in the reference's format, one row per iteration
  Peak6 f;
  Vol6 vol;
  Workspace<double, 6, 1> w;  // debug = 1
  auto r = w.integrate<Peak6>(f, 1e-4, 1e-40, vol);
  std::ifstream in("cuda_pagani_iters.csv");
  std::string header, line, last;
  std::getline(in, header);
  size_t rows = 0;
  while (std::getline(in, line))
    if (!line.empty()) {
      last = line;
      ++rows;
    }
  double it = -1, est = 0, err = 0;
  std::sscanf(last.c_str(), "%lf,%lf,%lf", &it, &est, &err);
  // the last row is printed before the convergence test of the last iteration: cumulative + that iteration
  CHECK("iters_csv", header == "it, estimate, errorest, festimate, ferrorest, nregions, nfregions" && rows >= 2 &&
                       static_cast<size_t>(it) + 1 == rows && std::fabs(est - r.estimate) <= 1e-5 * std::fabs(r.estimate) &&
                       std::fabs(err - r.errorest) <= 1e-5 * r.errorest,
        "rows %zu, last row: %s, result %.9g +- %.3g", rows, last.c_str(), r.estimate, r.errorest);
  std::ifstream tb("cuda_pagani_time_breakdown.csv");
  std::getline(tb, header);
  CHECK("time_breakdown_csv", header == "id, ndim, epsrel, it, name, time", "header: %s", header.c_str());
}

int
main()
{
  const double exact = 5. + 4. / 2 + 3. / 2 + 2. / 4;  // integral of d0 + x d1 + y d2 + x y d3 over the unit square
  {
    std::array<int, 5> attributes{5, 4, 3, 2, 1};
    ToyIntegrand integrand(attributes);
    Workspace<double, 2> w;
    quad::Volume<double, 2> vol;
    auto r = w.integrate<ToyIntegrand>(integrand, 1e-8, 1e-40, vol);
    CHECK("cudaArray_state", r.status == 0 && std::fabs(r.estimate - exact) <= 1e-8 * exact, "estimate %.17g", r.estimate);
  }
  {
    std::array<int, 5> attributes{5, 4, 3, 2, 1};
    ToyIntegrandWithHeapModel integrand(attributes.data(), attributes.size());
    Workspace<double, 2> w;
    quad::Volume<double, 2> vol;
    auto r = w.integrate<ToyIntegrandWithHeapModel>(integrand, 1e-8, 1e-40, vol);
    auto r2 = w.integrate<ToyIntegrandWithHeapModel>(integrand, 1e-8, 1e-40, vol);  // the functor is copied per call
    CHECK("cudaDynamicArray_state", r.status == 0 && std::fabs(r.estimate - exact) <= 1e-8 * exact && r2.estimate == r.estimate,
          "estimate %.17g", r.estimate);
  }
  {
    // y = 2 x on a 9-point table (linear interpolation is exact): integral of 2x (1 + y z) over the unit cube = 1.25
    std::array<double, 9> xs, ys;
    for (int i = 0; i < 9; ++i) {
      xs[i] = i / 8.;
      ys[i] = 2. * xs[i];
    }
    InterpIntegrand integrand(xs, ys);
    Workspace<double, 3> w;
    quad::Volume<double, 3> vol;
    auto r = w.integrate<InterpIntegrand>(integrand, 1e-6, 1e-40, vol);
    CHECK("interp1d_state", r.status == 0 && std::fabs(r.estimate - 1.25) <= 1e-6 * 1.25, "estimate %.17g", r.estimate);
  }
  {
    // snapshot / resume.  Run A: Sub_regions& overload from the uniform grid to convergence.  Run B: the same, but
    // stopped after a few iterations by a loose tolerance... the reference has no iteration cap in its interface, so
    // the intermediate list is produced the way its tests do it: evaluate / filter / split by hand with the stage
    // objects, snapshot, and hand that list to integrate() -- once directly, once after a round trip through
    // take_snapshot / load_snapshot into a fresh Sub_regions in a NEW Workspace.
    constexpr size_t nd = 6;
    Peak6 f;
    Peak6* d_f = quad::make_gpu_integrand(f);
    quad::Volume<double, nd> vol;
    Cubature_rules<double, nd> rules;
    rules.set_device_volume(vol.lows, vol.highs);
    Sub_regions<double, nd> regs(2);
    for (int it = 0; it < 3; ++it) {  // three all-active iterations: 64 -> 512 regions
      const size_t n = regs.size;
      Region_characteristics<nd> chars(n);
      Region_estimates<double, nd> ests(n), parents(n / 2 > 0 ? n / 2 : 1);
      rules.apply_cubature_integration_rules<Peak6>(d_f, it, regs, ests, chars, true);
      Sub_regions_filter<double, nd> filter(n);
      const size_t m = filter.filter(regs, chars, ests, parents);
      Sub_region_splitter<double, nd> splitter(m);
      splitter.split(regs, chars);
    }
    CHECK("stage_objects_grow_list", regs.size == 512, "size %zu", regs.size);
    regs.take_snapshot();
    Sub_regions<double, nd> direct(regs);  // deep copy of the intermediate list
    Workspace<double, nd> wa;
    auto ra = wa.integrate<Peak6>(f, direct, 1e-4, 1e-40, vol);
    const std::vector<double> final_a = host_copy(direct);

    regs.uniform_split(1);  // clobber the live list, then restore it from the snapshot
    regs.load_snapshot();
    CHECK("load_snapshot_restores", regs.size == 512, "size %zu", regs.size);
    Workspace<double, nd> wb;  // a NEW workspace resumes from the snapshot
    auto rb = wb.integrate<Peak6>(f, regs, 1e-4, 1e-40, vol);
    const std::vector<double> final_b = host_copy(regs);
    CHECK("resume_equals_uninterrupted",
          ra.status == 0 && rb.status == ra.status && rb.estimate == ra.estimate && rb.errorest == ra.errorest &&
            rb.nregions == ra.nregions && rb.iters == ra.iters && final_a == final_b,
          "estimate %.17g vs %.17g, nregions %zu vs %zu, final list %zu vs %zu doubles", ra.estimate, rb.estimate,
          ra.nregions, rb.nregions, final_a.size(), final_b.size());
    cudaFree(d_f);
  }
  recorder_case();
  std::printf("%d failure(s)\n", failures);
  return failures;
}
