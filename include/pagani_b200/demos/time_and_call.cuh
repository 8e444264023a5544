// Benchmark/CSV driver with the reference's names and output columns, on the B200-native engine:
//   clean_time_and_call<F, ndim, use_custom, debug, runs_per_epsrel, collect_mult_runs>(id, epsrel, outfile)
//   print_header(), get_partitions_per_axis(ndim)
// (reference cuda/pagani/demos/new_time_and_call.cuh:115-126, 190-253; the demo programs such as
// cuda/pagani/demos/genz_integrals.cu:9-82 are loops over this call).  One CSV row per timed run:
//   id, ndim, library|custom, true_value, epsrel, epsabs, estimate, errorest, nregions,
//   nFinishedRegions, iters, status, milliseconds
// in the reference's number format (scientific, 15 digits), the first run of every (integrand, epsrel)
// being an unreported warm-up exactly as in the reference, so its analysis scripts read the file as is.
// Timer placement is the reference's: the clock starts before the initial uniform split and stops when
// integrate() returns; Workspace construction is outside.
//
// F follows the reference's integrand contract: default-constructible, `void set_true_value()` and a
// public `double true_value` (common/cuda/integrands.cuh), plus the device call operator.
#ifndef PAGANI_B200_DEMOS_TIME_AND_CALL_CUH
#define PAGANI_B200_DEMOS_TIME_AND_CALL_CUH

#include <chrono>
#include <cmath>
#include <iomanip>
#include <iostream>
#include <ostream>
#include <string>

#include "../Workspace.cuh"

// initial grid: 2 cuts per axis from 5-D up, finer below (new_time_and_call.cuh:115-126)
inline size_t
get_partitions_per_axis(int ndim)
{
  if (ndim < 5)
    return 4;
  if (ndim <= 10)
    return 2;
  return 1;
}

inline void
print_header(std::ostream& os = std::cout)
{
  os << "id, ndim, use_custom, integral, epsrel, epsabs, estimate, errorest, nregions, nFinishedRegions, "
        "iters, status, time\n";
}

template <typename F, int ndim, bool use_custom = false, int debug = 0, int runs_per_esprel = 10,
          bool collect_mult_runs = false>
bool
clean_time_and_call(std::string id, double epsrel, std::ostream& outfile, bool echo = true)
{
  using MilliSeconds = std::chrono::duration<double, std::chrono::milliseconds::period>;
  double constexpr epsabs = 1.0e-40;
  bool good = false;
  bool relerr_classification = true;
  Workspace<double, ndim, debug, use_custom, collect_mult_runs> workspace;
  F integrand;
  integrand.set_true_value();
  auto write_row = [&](std::ostream& os, numint::integration_result const& result, double ms) {
    os << std::fixed << std::scientific << std::setprecision(15) << id << "," << ndim << ","
       << (use_custom ? "custom" : "library") << "," << integrand.true_value << "," << epsrel << ","
       << epsabs << "," << result.estimate << "," << result.errorest << "," << result.nregions << ","
       << result.nFinishedRegions << "," << result.iters << "," << result.status << "," << ms << std::endl;
  };
  for (int i = 0; i < runs_per_esprel; i++) {
    quad::Volume<double, ndim> vol;
    auto const t0 = std::chrono::high_resolution_clock::now();
    Sub_regions<double, ndim> sub_regions(get_partitions_per_axis(ndim));
    numint::integration_result result = workspace.template integrate<F, false, false>(
      integrand, sub_regions, epsrel, epsabs, vol, relerr_classification, id);
    MilliSeconds dt = std::chrono::high_resolution_clock::now() - t0;
    if (result.status == 0)
      good = true;
    if (i != 0) {
      if (echo)
        write_row(std::cout, result, dt.count());
      write_row(outfile, result, dt.count());
    }
  }
  return good;
}

#endif
