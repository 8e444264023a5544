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
//           Sub_regions<T,ndim> and the other stage-level classes: stages.cuh
// The integrand is any copyable functor with a __device__ double operator()(double x ndim); it is
// compiled into the evaluation kernel of evaluate.cuh, everything else runs behind the C ABI.
// Errors: CUDA failures abort with a message like the reference's CudaCheckError
// (common/cuda/cudaDebugUtil.h:118-151); allocation failure throws std::bad_alloc (cudaMemoryUtil.h:198-208).
#ifndef PAGANI_B200_WORKSPACE_CUH
#define PAGANI_B200_WORKSPACE_CUH

#include <fstream>
#include <string>

#include "stages.cuh"

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
    // the caller's device list is copied into the engine's arena; afterwards the caller's object is re-pointed at a
    // copy of the list the run ended with (the reference frees and replaces its arrays every iteration,
    // Workspace.cuh:148-358)
    pagani_b200::check(pgn_regions_import_device(ctx_, subregions.dLeftCoord, subregions.dLength, subregions.size),
                       "pgn_regions_import_device");
    numint::integration_result res = run(integrand, epsrel, epsabs, vol, relerr_classification, true, optional);
    size_t n = 0;
    pgn_regions_count(ctx_, &n);
    subregions.device_init(n);
    pagani_b200::check(pgn_regions_export_device(ctx_, subregions.dLeftCoord, subregions.dLength, n, &n),
                       "pgn_regions_export_device");
    return res;
  }

  template <typename IntegT, bool predict_split = false, bool collect_iters = false>
  numint::integration_result
  integrate(const IntegT& integrand, T epsrel, T epsabs, quad::Volume<T, ndim> const& vol,
            bool relerr_classification = true)
  {
    return run(integrand, epsrel, epsabs, vol, relerr_classification, false);
  }

  pgn_ctx* context() { return ctx_; }

private:
  template <typename IntegT>
  numint::integration_result
  run(const IntegT& integrand, T epsrel, T epsabs, quad::Volume<T, ndim> const& vol,
      bool relerr_classification, bool caller_regions, const std::string& id = "default")
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
    if constexpr (debug > 0)
      write_recorders(epsrel, id);
    return out;
  }

  // debug > 0: the reference's two recorder files (Workspace.cuh:171-242; Recorder, common/cuda/cudaDebugUtil.h:18-37),
  // same names, header lines and columns, written from the engine's per-iteration trace once the call has ended
  // (truncated per call, appended with collect_mult_runs):
  //   cuda_pagani_iters.csv           it, estimate, errorest, festimate, ferrorest, nregions, nfregions
  //   cuda_pagani_time_breakdown.csv  id, ndim, epsrel, it, name, time   (device milliseconds per stage)
  void
  write_recorders(T epsrel, const std::string& id)
  {
    const auto mode = collect_mult_runs ? std::ios_base::app : std::ios_base::out;
    std::ofstream iters("cuda_pagani_iters.csv", mode), times("cuda_pagani_time_breakdown.csv", mode);
    if (!collect_mult_runs) {
      iters << "it, estimate, errorest, festimate, ferrorest, nregions, nfregions" << std::endl;
      times << "id, ndim, epsrel, it, name, time" << std::endl;
    }
    double cum_est = 0., cum_err = 0.;
    size_t finished_regions = 0;
    const int n = pgn_trace_count(ctx_);
    for (int i = 0; i < n; ++i) {
      pgn_iter_stats st;
      if (pgn_trace_get(ctx_, i, &st) != PGN_OK)
        break;
      // the reference prints before this iteration's finished regions are added (Workspace.cuh:230-241)
      iters << st.iteration << "," << cum_est + st.iter_estimate << "," << cum_err + st.iter_errorest << "," << cum_est
            << "," << cum_err << "," << st.nregions_global << "," << finished_regions << std::endl;
      const char* names[4] = {"apply_cubature_rules", "two_level_errorest", "heuristic_classify", "filter_and_split"};
      const double ms[4] = {st.eval_ms, st.classify_ms, st.heuristic_ms, st.split_ms};
      for (int k = 0; k < 4; ++k)
        times << id << "," << ndim << "," << epsrel << "," << st.iteration << "," << names[k] << "," << ms[k]
              << std::endl;
      cum_est = st.cum_estimate;
      cum_err = st.cum_errorest;
      finished_regions += st.nregions_global - st.nactive;
    }
  }

  pgn_ctx* ctx_ = nullptr;
};

#endif
