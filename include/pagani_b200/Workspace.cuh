// Drop-in C++ front end: the reference's PAGANI interface (same names, argument meaning, result
// type and error behaviour) on top of the B200-native engine.  A caller of the reference replaces
//     #include "cuda/pagani/quad/GPUquad/Workspace.cuh"
// by
//     #include "pagani_b200/Workspace.cuh"          (and links libpagani_b200.so)
//
// Mirrors:  Workspace<T,ndim,debug,use_custom,collect_mult_runs>::integrate, both overloads
//           (reference cuda/pagani/quad/GPUquad/Workspace.cuh:23-71, 148-358, 361-494)
//           numint::integration_result   (common/integration_result.hh:11-21)
//           quad::Volume<T,NDIM>         (common/cuda/Volume.cuh:13-52)
//           Sub_regions<T,ndim>          (GPUquad/Sub_regions.cuh:38-258; host-visible subset)
// The integrand is any copyable functor with a __device__ double operator()(double x ndim); it is
// compiled into the evaluation kernel of evaluate.cuh, everything else runs behind the C ABI.
// Errors: CUDA failures abort with a message like the reference's CudaCheckError
// (common/cuda/cudaDebugUtil.h:118-151); allocation failure throws std::bad_alloc (cudaMemoryUtil.h:198-208).
#ifndef PAGANI_B200_WORKSPACE_CUH
#define PAGANI_B200_WORKSPACE_CUH

#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <ostream>
#include <string>
#include <vector>

#include "../pagani_b200.h"
#include "evaluate.cuh"

namespace numint {
  struct integration_result {
    double estimate = 0.;
    double errorest = 0.;
    size_t neval = 0;
    size_t nregions = 0;
    size_t nFinishedRegions = 0;
    int status = -1;
    int lastPhase = -1;
    double chi_sq = 0.;
    size_t iters = 0;
  };
  inline std::ostream&
  operator<<(std::ostream& os, integration_result const& res)
  {
    os << res.estimate << "," << res.errorest << "," << res.nregions << "," << res.chi_sq << ","
       << res.status;
    return os;
  }
}

namespace quad {
  template <typename T, int NDIM>
  struct Volume {
    T lows[NDIM] = {0.0};
    T highs[NDIM];
    Volume()
    {
      for (T& x : highs)
        x = 1.0;
    }
    Volume(std::array<T, NDIM> l, std::array<T, NDIM> h)
    {
      std::memcpy(lows, l.data(), NDIM * sizeof(T));
      std::memcpy(highs, h.data(), NDIM * sizeof(T));
    }
    Volume(T const* l, T const* h)
    {
      std::memcpy(lows, l, NDIM * sizeof(T));
      std::memcpy(highs, h, NDIM * sizeof(T));
    }
    Volume(T l, T h)
    {
      for (int d = 0; d < NDIM; ++d) {
        lows[d] = l;
        highs[d] = h;
      }
    }
  };
}

namespace pagani_b200 {
  inline void
  check(int rc, const char* what)
  {
    if (rc == PGN_OK)
      return;
    if (rc == PGN_ERR_NOMEM)
      throw std::bad_alloc();
    std::fprintf(stderr, "pagani_b200: %s failed (%d): %s\n", what, rc, pgn_last_error());
    std::abort();
  }
}

// The caller-visible region list.  The device copy lives inside the Workspace's engine context;
// this object carries what the caller specified (a uniform grid or explicit host arrays) and, after
// integrate(), the size of the list the run ended with.
template <typename T, size_t ndim>
struct Sub_regions {
  Sub_regions() {}
  explicit Sub_regions(const size_t partitions_per_axis) { uniform_split(partitions_per_axis); }

  void
  uniform_split(size_t numOfDivisionPerRegionPerDimension)
  {
    partitions = numOfDivisionPerRegionPerDimension;
    size = 1;
    for (size_t d = 0; d < ndim; ++d)
      size *= partitions;
    LeftCoord.clear();
    Length.clear();
  }

  // explicit list, SoA [dim * n + i] exactly like dLeftCoord / dLength of the reference
  void
  set_host(const T* left, const T* length, size_t n)
  {
    partitions = 0;
    size = n;
    LeftCoord.assign(left, left + n * ndim);
    Length.assign(length, length + n * ndim);
  }

  T
  compute_region_volume(size_t const regionID) const
  {
    T reg_vol = 1.;
    for (size_t dim = 0; dim < ndim; dim++)
      reg_vol *= Length[size * dim + regionID];
    return reg_vol;
  }

  std::vector<T> LeftCoord, Length;  // host copies, filled by Workspace::refresh_host
  size_t partitions = 0;
  size_t size = 0;
};

template <typename T, size_t ndim, int debug = 0, bool use_custom = false, bool collect_mult_runs = false>
class Workspace {
  static_assert(sizeof(T) == sizeof(double), "the engine is FP64 only, like every use in the reference");

public:
  Workspace()
  {
    pgn_opts o;
    std::memset(&o, 0, sizeof o);
    o.device = -1;
    o.trace = debug > 0;
    pagani_b200::check(pgn_create(&ctx_, static_cast<int>(ndim), &o), "pgn_create");
  }
  ~Workspace() { pgn_destroy(ctx_); }
  Workspace(const Workspace&) = delete;
  Workspace& operator=(const Workspace&) = delete;

  template <typename IntegT, bool predict_split = false, bool collect_iters = false>
  numint::integration_result
  integrate(const IntegT& integrand, Sub_regions<T, ndim>& subregions, T epsrel, T epsabs,
            quad::Volume<T, ndim> const& vol, bool relerr_classification = true,
            const std::string& optional = "default")
  {
    (void)optional;
    if (subregions.partitions > 0)
      pagani_b200::check(pgn_regions_init_uniform(ctx_, static_cast<int>(subregions.partitions)),
                         "pgn_regions_init_uniform");
    else
      pagani_b200::check(pgn_regions_import(ctx_, subregions.LeftCoord.data(), subregions.Length.data(),
                                            subregions.size),
                         "pgn_regions_import");
    numint::integration_result res = run(integrand, epsrel, epsabs, vol, relerr_classification, true);
    size_t n = 0;
    pgn_regions_count(ctx_, &n);
    subregions.size = n;  // the reference leaves the caller's list replaced by the final one
    subregions.partitions = 0;
    subregions.LeftCoord.clear();
    subregions.Length.clear();
    return res;
  }

  template <typename IntegT, bool predict_split = false, bool collect_iters = false>
  numint::integration_result
  integrate(const IntegT& integrand, T epsrel, T epsabs, quad::Volume<T, ndim> const& vol,
            bool relerr_classification = true)
  {
    return run(integrand, epsrel, epsabs, vol, relerr_classification, false);
  }

  // Sub_regions::refresh_host_device (Sub_regions.cuh:82-87): copy the engine's current list back
  void
  refresh_host(Sub_regions<T, ndim>& subregions)
  {
    size_t n = 0;
    pgn_regions_count(ctx_, &n);
    subregions.LeftCoord.resize(n * ndim);
    subregions.Length.resize(n * ndim);
    pagani_b200::check(
      pgn_regions_export(ctx_, subregions.LeftCoord.data(), subregions.Length.data(), n, &n),
      "pgn_regions_export");
    subregions.size = n;
  }

  pgn_ctx* context() { return ctx_; }

private:
  template <typename IntegT>
  numint::integration_result
  run(const IntegT& integrand, T epsrel, T epsabs, quad::Volume<T, ndim> const& vol,
      bool relerr_classification, bool caller_regions)
  {
    pagani_b200::check(pgn_set_volume(ctx_, vol.lows, vol.highs), "pgn_set_volume");
    pagani_b200::Evaluator<IntegT, static_cast<int>(ndim)> evaluator(integrand);
    pgn_integrate_args a;
    std::memset(&a, 0, sizeof a);
    a.epsrel = epsrel;
    a.epsabs = epsabs;
    a.relerr_classification = relerr_classification ? 1 : 0;
    a.use_current_regions = caller_regions ? 1 : 0;
    a.iters_mode = caller_regions ? PGN_ITERS_BEFORE_CHECK : PGN_ITERS_AFTER_SPLIT;
    pgn_result_t r;
    pagani_b200::check(pgn_integrate(ctx_, evaluator.fn(), evaluator.user(), &a, &r), "pgn_integrate");
    numint::integration_result out;
    out.estimate = r.estimate;
    out.errorest = r.errorest;
    out.neval = r.neval;
    out.nregions = r.nregions;
    out.nFinishedRegions = r.nFinishedRegions;
    out.status = r.status;
    out.lastPhase = r.lastPhase;
    out.chi_sq = r.chi_sq;
    out.iters = r.iters;
    return out;
  }

  pgn_ctx* ctx_ = nullptr;
};

#endif
