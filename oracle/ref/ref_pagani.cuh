// TEST INFRASTRUCTURE (oracle side) -- not part of the product, never linked into it.
//
// Drives the UNMODIFIED reference PAGANI (marcpaterno/gpuintegration, CUDA back-end, headers used
// where they lie under /root/reference) and records what it does, so that the B200-native engine
// can be compared with it region by region.  Only include paths are supplied; no reference file is
// copied or edited.
//
//  * mode "integrate": calls Workspace<double,ndim>::integrate (Workspace.cuh:361-494 or, with
//    --subregions, the Sub_regions& overload :148-358) and prints the integration_result and the
//    wall time of each repeat, following the demos' timing protocol (new_time_and_call.cuh:187-247).
//  * mode "trace": runs the same loop re-assembled from the reference's own public stage objects
//    (Cubature_rules, two_level_errorest_and_relerr_classify, reduction, accuracy_reached,
//    compute_finished_estimates, Sub_regions_filter, Sub_region_splitter -- the way
//    pagani_clone, PaganiUtils.cuh:405-515, does) plus Workspace's two private helpers
//    (heuristic_classify, fix_error_budget_overflow) reached through explicit template
//    instantiation, and after every stage copies the device arrays back.  Per iteration it writes one
//    JSON line (scalars + 64-bit digests of every array) and, for iterations with at most
//    --full regions, the raw arrays.  The final result is cross-checked against a plain
//    Workspace::integrate call in the same process.
#ifndef ORACLE_REF_PAGANI_CUH
#define ORACLE_REF_PAGANI_CUH

#include <array>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "cuda/pagani/quad/GPUquad/Workspace.cuh"

struct RefOptions {
  std::string mode = "integrate";
  std::string integrand;
  double epsrel = 1e-3;
  double epsabs = 1e-40;
  int repeats = 3;
  bool subregions_overload = true;  // what the reference demos time
  std::string out_prefix = "ref";
  size_t full_max = 0;  // dump raw arrays for iterations with <= this many regions
  int max_iters = 700;
};

typedef int (*RefRunner)(const RefOptions&);
std::map<std::string, RefRunner>& ref_registry();
struct RefRegistrar {
  RefRegistrar(const char* name, RefRunner r) { ref_registry()[name] = r; }
};

// ---- access to Workspace's private helpers without touching the reference -----------------------
template <typename Tag, typename Tag::type M>
struct RefRob {
  friend typename Tag::type ref_get(Tag) { return M; }
};
template <size_t ndim>
struct RefTagHC {
  typedef bool (Workspace<double, ndim>::*type)(Heuristic_classifier<double, ndim>&,
                                                 Region_characteristics<ndim>&,
                                                 const Region_estimates<double, ndim>&,
                                                 numint::integration_result&,
                                                 const numint::integration_result&,
                                                 const numint::integration_result&);
  friend type ref_get(RefTagHC);
};
template <size_t ndim>
struct RefTagFB {
  typedef void (Workspace<double, ndim>::*type)(Region_characteristics<ndim>&,
                                                 const numint::integration_result&,
                                                 const numint::integration_result&,
                                                 numint::integration_result&, const double);
  friend type ref_get(RefTagFB);
};
#define REF_EXPOSE_WORKSPACE(ndim)                                                                 \
  template struct RefRob<RefTagHC<ndim>, &Workspace<double, ndim>::heuristic_classify>;           \
  template struct RefRob<RefTagFB<ndim>, &Workspace<double, ndim>::fix_error_budget_overflow>;

inline uint64_t
ref_digest(const void* data, size_t elem, size_t n)
{
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint64_t acc = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t bits = 0;
    memcpy(&bits, p + i * elem, elem);
    acc += bits * (2ull * i + 1ull);
  }
  return acc;
}

template <typename T>
std::vector<T>
ref_to_host(const T* dev, size_t n)
{
  std::vector<T> h(n);
  if (n)
    cudaMemcpy(h.data(), dev, sizeof(T) * n, cudaMemcpyDeviceToHost);
  return h;
}

inline void
ref_write(FILE* f, const void* p, size_t bytes)
{
  if (bytes && fwrite(p, 1, bytes, f) != bytes) {
    perror("fwrite");
    exit(2);
  }
}

template <typename F, int ndim>
int
ref_integrate(const RefOptions& o)
{
  using MilliSeconds = std::chrono::duration<double, std::chrono::milliseconds::period>;
  Workspace<double, ndim> workspace;
  F integrand;
  for (int i = 0; i < o.repeats; ++i) {
    quad::Volume<double, ndim> vol;
    cudaDeviceSynchronize();
    auto const t0 = std::chrono::high_resolution_clock::now();
    numint::integration_result r;
    if (o.subregions_overload) {
      size_t parts = ndim < 5 ? 4 : (ndim <= 10 ? 2 : 1);
      Sub_regions<double, ndim> sub_regions(parts);
      r = workspace.template integrate<F>(integrand, sub_regions, o.epsrel, o.epsabs, vol, true,
                                          o.integrand);
    } else {
      r = workspace.template integrate<F>(integrand, o.epsrel, o.epsabs, vol, true);
    }
    MilliSeconds dt = std::chrono::high_resolution_clock::now() - t0;
    printf("{\"kind\":\"integrate\",\"integrand\":\"%s\",\"ndim\":%d,\"epsrel\":%.17g,\"epsabs\":%.17g,"
           "\"repeat\":%d,\"estimate\":%.17g,\"errorest\":%.17g,\"nregions\":%zu,\"iters\":%zu,"
           "\"status\":%d,\"ms\":%.6f,\"overload\":\"%s\"}\n",
           o.integrand.c_str(), ndim, o.epsrel, o.epsabs, i, r.estimate, r.errorest, r.nregions,
           r.iters, r.status, dt.count(), o.subregions_overload ? "subregions" : "plain");
    fflush(stdout);
  }
  return 0;
}

template <typename F, int ndim>
int
ref_trace(const RefOptions& o)
{
  using Estimates = Region_estimates<double, ndim>;
  using Chars = Region_characteristics<ndim>;
  using Classifier = Heuristic_classifier<double, ndim>;
  using Res = numint::integration_result;
  const double epsrel = o.epsrel, epsabs = o.epsabs;

  std::string jl = o.out_prefix + ".jsonl";
  FILE* js = fopen(jl.c_str(), "w");
  if (!js) {
    perror(jl.c_str());
    return 2;
  }

  Workspace<double, ndim> helper;  // only for its two private member functions
  auto hc = ref_get(RefTagHC<ndim>());
  auto fb = ref_get(RefTagFB<ndim>());

  Cubature_rules<double, ndim> rules;
  quad::Volume<double, ndim> vol;
  rules.set_device_volume(vol.lows, vol.highs);
  Estimates prev_iter_estimates;
  Res cummulative;
  size_t partitions_per_axis = ndim < 5 ? 4 : (ndim <= 10 ? 2 : 1);
  Sub_regions<double, ndim> subregions(partitions_per_axis);
  Classifier classifier(epsrel, epsabs);
  cummulative.status = 1;
  F integrand;
  F* d_integrand = quad::cuda_copy_to_managed(integrand);
  const size_t feval = pagani::CuhreFuncEvalsPerRegion<ndim>();
  bool finished_loop = false;

  for (size_t it = 0; it < (size_t)o.max_iters && subregions.size > 0; it++) {
    const size_t n = subregions.size;
    Chars characteristics(n);
    Estimates estimates(n);
    const bool full = n <= o.full_max;

    std::vector<double> h_left = ref_to_host(subregions.dLeftCoord, n * ndim);
    std::vector<double> h_len = ref_to_host(subregions.dLength, n * ndim);

    Res iter = rules.template apply_cubature_integration_rules<F>(d_integrand, it, subregions,
                                                                  estimates, characteristics, false);
    std::vector<double> h_est = ref_to_host(estimates.integral_estimates, n);
    std::vector<double> h_err_raw = ref_to_host(estimates.error_estimates, n);
    std::vector<int> h_axis = ref_to_host(characteristics.sub_dividing_dim, n);

    two_level_errorest_and_relerr_classify<double, ndim>(estimates, prev_iter_estimates,
                                                          characteristics, epsrel, true);
    iter.errorest = reduction<double, false>(estimates.error_estimates, n);
    std::vector<double> h_err = ref_to_host(estimates.error_estimates, n);
    std::vector<double> h_act0 = ref_to_host(characteristics.active_regions, n);

    auto emit = [&](const char* outcome, const Res& finished, size_t nactive, int overflow,
                    int hran, const std::vector<double>& act_final) {
      std::vector<uint8_t> a0(n), a1(n);
      for (size_t i = 0; i < n; ++i) {
        a0[i] = h_act0[i] != 0.0;
        a1[i] = act_final.empty() ? a0[i] : (act_final[i] != 0.0);
      }
      fprintf(js,
              "{\"it\":%zu,\"n\":%zu,\"outcome\":\"%s\",\"iter_estimate\":%.17g,\"iter_errorest\":%.17g,"
              "\"finished_estimate\":%.17g,\"finished_errorest\":%.17g,\"cum_estimate\":%.17g,"
              "\"cum_errorest\":%.17g,\"cum_nregions\":%zu,\"nactive\":%zu,\"budget_overflow\":%d,"
              "\"heuristic_ran\":%d,\"dig_left\":\"%llu\",\"dig_length\":\"%llu\",\"dig_estimate\":\"%llu\","
              "\"dig_errorest_raw\":\"%llu\",\"dig_errorest\":\"%llu\",\"dig_axis\":\"%llu\","
              "\"dig_active_relerr\":\"%llu\",\"dig_active\":\"%llu\",\"full\":%d}\n",
              it, n, outcome, iter.estimate, iter.errorest, finished.estimate, finished.errorest,
              cummulative.estimate, cummulative.errorest, cummulative.nregions, nactive, overflow, hran,
              (unsigned long long)ref_digest(h_left.data(), 8, h_left.size()),
              (unsigned long long)ref_digest(h_len.data(), 8, h_len.size()),
              (unsigned long long)ref_digest(h_est.data(), 8, n),
              (unsigned long long)ref_digest(h_err_raw.data(), 8, n),
              (unsigned long long)ref_digest(h_err.data(), 8, n),
              (unsigned long long)ref_digest(h_axis.data(), 4, n),
              (unsigned long long)ref_digest(a0.data(), 1, n),
              (unsigned long long)ref_digest(a1.data(), 1, n), full ? 1 : 0);
      fflush(js);
      if (full) {
        std::string bn = o.out_prefix + ".it" + std::to_string(it) + ".bin";
        FILE* b = fopen(bn.c_str(), "wb");
        if (!b) {
          perror(bn.c_str());
          exit(2);
        }
        uint64_t hdr[4] = {0x50474e5245463031ull, (uint64_t)ndim, (uint64_t)n, (uint64_t)it};
        ref_write(b, hdr, sizeof hdr);
        ref_write(b, h_left.data(), 8 * h_left.size());
        ref_write(b, h_len.data(), 8 * h_len.size());
        ref_write(b, h_est.data(), 8 * n);
        ref_write(b, h_err_raw.data(), 8 * n);
        ref_write(b, h_err.data(), 8 * n);
        ref_write(b, h_axis.data(), 4 * n);
        ref_write(b, a0.data(), n);
        ref_write(b, a1.data(), n);
        fclose(b);
      }
    };

    if (accuracy_reached(epsrel, epsabs, std::abs(cummulative.estimate + iter.estimate),
                         cummulative.errorest + iter.errorest)) {
      cummulative.estimate += iter.estimate;
      cummulative.errorest += iter.errorest;
      cummulative.status = 0;
      cummulative.nregions += n;
      Res none;
      emit("converged", none, 0, 0, 0, std::vector<double>());
      finished_loop = true;
      break;
    }
    classifier.store_estimate(cummulative.estimate + iter.estimate);
    Res finished = compute_finished_estimates<double, ndim, false>(estimates, characteristics, iter);
    const double fin_before = finished.errorest;
    (helper.*fb)(characteristics, cummulative, iter, finished, epsrel);
    const int overflow = (finished.errorest == 0. && finished.estimate == 0. && fin_before != 0.) ? 1 : 0;
    const int hran = classifier.classification_criteria_met(n) ? 1 : 0;
    const bool must_terminate =
      (helper.*hc)(classifier, characteristics, estimates, finished, iter, cummulative);
    std::vector<double> h_act1 = ref_to_host(characteristics.active_regions, n);
    size_t nactive = 0;
    for (double a : h_act1)
      nactive += a != 0.0;
    if (must_terminate) {
      cummulative.estimate += iter.estimate;
      cummulative.errorest += iter.errorest;
      cummulative.nregions += n;
      emit("terminated", finished, nactive, overflow, hran, h_act1);
      finished_loop = true;
      break;
    }
    cummulative.estimate += finished.estimate;
    cummulative.errorest += finished.errorest;
    Sub_regions_filter<double, ndim, false> filter_obj(n);
    size_t num_active_regions =
      filter_obj.filter(subregions, characteristics, estimates, prev_iter_estimates);
    cummulative.nregions += n - num_active_regions;
    subregions.size = num_active_regions;
    emit("split", finished, num_active_regions, overflow, hran, h_act1);
    Sub_region_splitter<double, ndim> splitter(subregions.size);
    splitter.split(subregions, characteristics);
    cummulative.iters++;
  }
  if (!finished_loop)
    cummulative.nregions += subregions.size;
  d_integrand->~F();
  cudaFree(d_integrand);
  fclose(js);

  // cross-check the re-assembled loop against the reference's own integrate()
  Workspace<double, ndim> ws;
  numint::integration_result real = ws.template integrate<F>(integrand, epsrel, epsabs, vol, true);
  const bool same = real.estimate == cummulative.estimate && real.errorest == cummulative.errorest &&
                    real.nregions == cummulative.nregions && real.status == cummulative.status &&
                    real.iters == cummulative.iters;
  printf("{\"kind\":\"trace\",\"integrand\":\"%s\",\"ndim\":%d,\"feval\":%zu,\"epsrel\":%.17g,"
         "\"epsabs\":%.17g,\"estimate\":%.17g,\"errorest\":%.17g,\"nregions\":%zu,\"iters\":%zu,"
         "\"status\":%d,\"workspace_estimate\":%.17g,\"workspace_errorest\":%.17g,"
         "\"workspace_nregions\":%zu,\"workspace_iters\":%zu,\"workspace_status\":%d,"
         "\"matches_workspace\":%s}\n",
         o.integrand.c_str(), ndim, feval, epsrel, epsabs, cummulative.estimate, cummulative.errorest,
         cummulative.nregions, cummulative.iters, cummulative.status, real.estimate, real.errorest,
         real.nregions, real.iters, real.status, same ? "true" : "false");
  fflush(stdout);
  return same ? 0 : 3;
}

template <typename F, int ndim>
int
ref_run(const RefOptions& o)
{
  if (o.mode == "trace")
    return ref_trace<F, ndim>(o);
  return ref_integrate<F, ndim>(o);
}

#define REF_REGISTER(name, F, ndim) static RefRegistrar reg_##name(#name, &ref_run<F, ndim>);

#endif
