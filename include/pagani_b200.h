/*
 * pagani_b200.h -- C ABI of the B200-native PAGANI adaptive-cubature engine.
 *
 * Everything the reference (marcpaterno/gpuintegration, CUDA back-end) does inside
 * Workspace<T,ndim>::integrate (cuda/pagani/quad/GPUquad/Workspace.cuh:148-494) except the
 * integrand-templated evaluation kernel lives behind these entry points; the evaluation kernel
 * is a header template (include/pagani_b200/evaluate.cuh) compiled with the caller's functor and
 * handed to the engine as a plain function pointer (pgn_eval_fn).  No C++ types, no exceptions and
 * no torch types cross this boundary.  All functions return PGN_OK (0) or a negative pgn_err.
 *
 * Threading: one pgn_ctx per host thread and per device (the reference is single threaded on the
 * legacy stream; we use one private non-blocking stream per context).
 */
#ifndef PAGANI_B200_H
#define PAGANI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGN_MAX_NDIM 15 /* 4*ndim+1 <= 64: same ceiling as the reference (Sample.cuh:151) */
#define PGN_NSETS 9
#define PGN_NRULES 5
#define PGN_NCODES 11 /* distinct generator coordinates: 0, +-l1 .. +-l5 */

typedef enum pgn_err {
  PGN_OK = 0,
  PGN_ERR_INVALID = -1,   /* bad argument */
  PGN_ERR_CUDA = -2,      /* a CUDA call failed (pgn_last_error has the text) */
  PGN_ERR_NOMEM = -3,     /* device allocation failed (reference: std::bad_alloc) */
  PGN_ERR_NO_DEVICE = -4, /* no sm_100 device / extension cannot run: there is NO CPU fallback */
  PGN_ERR_NCCL = -5,
  PGN_ERR_STATE = -6 /* call order violated */
} pgn_err;

/* Memory model of the heuristic (memory-pressure) classifier.
 * PARITY reproduces the reference's hard-coded 16e9-byte device
 * (heuristic_classifier.cuh:89-94) so the region trajectory is identical;
 * NATIVE uses the real free HBM of the device (times the number of ranks). */
typedef enum pgn_mem_mode { PGN_MEM_PARITY = 0, PGN_MEM_NATIVE = 1 } pgn_mem_mode;

/* Which reference overload's bookkeeping to mirror for integration_result::iters:
 * Workspace.cuh:488 increments after the split (value 0), :250 before the convergence test (1). */
typedef enum pgn_iters_mode { PGN_ITERS_AFTER_SPLIT = 0, PGN_ITERS_BEFORE_CHECK = 1 } pgn_iters_mode;

typedef struct pgn_opts {
  int device;            /* CUDA device ordinal, -1 = current */
  int mem_mode;          /* pgn_mem_mode */
  int max_iters;         /* reference: 700 (Workspace.cuh:182,399); 0 selects 700 */
  int trace;             /* 1: keep per-iteration pgn_iter_stats (pgn_trace_*) */
  int trace_arrays;      /* 1: additionally keep 64-bit digests of every per-region array */
  double native_mem_fraction; /* NATIVE mode: fraction of free HBM the region store may use (0 -> 0.85) */
} pgn_opts;

/* POD mirror of numint::integration_result (common/integration_result.hh:11-21).
 * neval is filled (sum over iterations of regions x points-per-region); the reference leaves it 0. */
typedef struct pgn_result_t {
  double estimate;
  double errorest;
  uint64_t neval;
  uint64_t nregions;
  uint64_t nFinishedRegions;
  int32_t status; /* 0 tolerance met, 1 stopped (iteration cap / memory heuristic / no active regions) */
  int32_t lastPhase;
  double chi_sq;
  uint64_t iters;
} pgn_result_t;

/* Device-side view of the rule tables (built once per context; mirrors what
 * quad::Rule<T>::Init + ComputeGenerators leave on the device, Rule.cuh:311-510, Phases.cuh:157-217,
 * re-laid-out: one packed 4-bit code per (point,axis) instead of a dense FP64 generator matrix). */
typedef struct pgn_rule_view {
  int32_t ndim;
  int32_t feval;                /* points per region, quad.h:68-75 */
  const uint64_t* point_code;   /* [feval] bits 4d..4d+3: coordinate code on axis d; bits 60..63: generator set */
  const double* weights;        /* [PGN_NSETS][PGN_NRULES]  cRuleWt */
  const double* scale;          /* [PGN_NSETS][PGN_NRULES]  GPUScale (rules 1..3 used) */
  const double* norm;           /* [PGN_NSETS][PGN_NRULES]  GPUNorm  (rules 1..3 used) */
  double half_plus_g[PGN_NCODES];  /* 0.5 + g(code) */
  double half_minus_g[PGN_NCODES]; /* 0.5 - g(code) */
  double ratio;                 /* (g[set2]/g[set1])^2, Sample.cuh:195-196 */
  double two_one_minus_ratio_unused;
  double weights_host[PGN_NSETS * PGN_NRULES]; /* host copy of `weights`: evaluators pass it in kernel parameters */
} pgn_rule_view;

/* What one evaluation launch sees: the SoA region store of this iteration and its outputs. */
#define PGN_ABI_VERSION 2u
typedef struct pgn_region_view {
  uint32_t abi_version; /* PGN_ABI_VERSION of the library that filled the view */
  uint32_t view_bytes;  /* sizeof(pgn_region_view) there: an evaluator compiled against other headers refuses */
  const double* left;   /* [ndim][n]  lower corner in unit-cube coordinates (Sub_regions::dLeftCoord) */
  const double* length; /* [ndim][n]  edge lengths                       (Sub_regions::dLength)    */
  size_t n;             /* regions in this launch */
  double* estimate;     /* [n] out  Region_estimates::integral_estimates */
  double* errorest;     /* [n] out  Region_estimates::error_estimates (raw, before two-level refinement) */
  int32_t* axis;        /* [n] out  Region_characteristics::sub_dividing_dim */
  double lows[PGN_MAX_NDIM + 1];   /* integration volume (quad::Volume) */
  double ranges[PGN_MAX_NDIM + 1]; /* highs - lows */
  double jacobian;      /* product of ranges, multiplied in axis order */
  pgn_rule_view rule;
  void* stream;         /* cudaStream_t the launch must be enqueued on */
  int32_t sm_count;
  int32_t reserved;
} pgn_region_view;

/* The integrand-specific part: enqueue the evaluation of every region in `view` on view->stream.
 * Generated by pagani_b200::make_evaluator<F,ndim>() for a user functor, or internal for Genz. */
typedef int (*pgn_eval_fn)(void* user, const pgn_region_view* view);

/* Per-iteration statistics (what the reference's debug recorder prints, Workspace.cuh:230-241, and more). */
typedef struct pgn_iter_stats {
  uint64_t iteration;
  uint64_t nregions;        /* regions evaluated this iteration (this rank) */
  uint64_t nregions_global; /* over all ranks */
  uint64_t nactive;         /* regions kept for splitting (global) */
  double iter_estimate, iter_errorest;       /* global sums */
  double finished_estimate, finished_errorest;
  double cum_estimate, cum_errorest;         /* after this iteration's finished contribution */
  int32_t heuristic_ran;    /* 1 if Heuristic_classifier::classify was entered */
  int32_t heuristic_probes; /* threshold probes it took */
  int32_t budget_overflow;  /* 1 if fix_error_budget_overflow re-activated everything */
  int32_t rebalanced;       /* 1 if regions moved between ranks after this iteration */
  double threshold;         /* final heuristic threshold (0 if none) */
  /* digests (opts.trace_arrays): sum over i of bits(x[i]) * (2i+1) mod 2^64, see pgn_digest64 */
  uint64_t dig_left, dig_length, dig_estimate, dig_errorest_raw, dig_errorest, dig_axis, dig_active;
  double eval_ms, classify_ms, heuristic_ms, split_ms; /* device times (CUDA events) when tracing */
} pgn_iter_stats;

typedef struct pgn_ctx pgn_ctx;

/* ---- lifetime ------------------------------------------------------------------------------- */
int pgn_create(pgn_ctx** out, int ndim, const pgn_opts* opts); /* replaces Workspace()/Cubature_rules() ctor, PaganiUtils.cuh:52-96 */
void pgn_destroy(pgn_ctx* ctx);                                /* ~Cubature_rules, PaganiUtils.cuh:114-128 */
const char* pgn_last_error(void);
const char* pgn_version(void);
int pgn_device_info(pgn_ctx* ctx, int* sm_count, size_t* free_bytes, size_t* total_bytes);
/* Which point-loop variant the pre-instantiated evaluators launch (they must agree bit for bit):
 * 0 default, 1 shift/mask addressing, 2 byte-permute addressing.  Also env PGN_EVAL_VARIANT. */
int pgn_set_eval_variant(pgn_ctx* ctx, int variant);
/* The context's private cudaStream_t (for event timing by the caller) and the number of kernels of
 * this library it has enqueued so far. */
int pgn_get_stream(pgn_ctx* ctx, void** stream);
int pgn_launch_count(pgn_ctx* ctx, uint64_t* launches);

/* ---- rule tables (golden-vector access; Rule.cuh) ------------------------------------------- */
int pgn_feval(int ndim); /* pagani::CuhreFuncEvalsPerRegion, quad.h:68-75 */
int pgn_rule_tables_host(int ndim, double* generators /*[ndim][feval]*/, double* weights /*[9][5]*/,
                         double* scale /*[9][5]*/, double* norm /*[9][5]*/, int32_t* set_of_point /*[feval]*/,
                         uint64_t* point_code /*[feval]*/);
int pgn_rule_view_get(pgn_ctx* ctx, pgn_rule_view* out);

/* ---- region store (Sub_regions.cuh) ---------------------------------------------------------- */
int pgn_regions_init_uniform(pgn_ctx* ctx, int parts_per_axis); /* Sub_regions::uniform_split :164-185 */
int pgn_regions_import(pgn_ctx* ctx, const double* left, const double* length, size_t n); /* host SoA [ndim][n] */
int pgn_regions_export(pgn_ctx* ctx, double* left, double* length, size_t cap, size_t* n); /* refresh_host_device :82-87 */
int pgn_regions_count(pgn_ctx* ctx, size_t* n);
int pgn_set_volume(pgn_ctx* ctx, const double* lows, const double* highs); /* Cubature_rules::set_device_volume, PaganiUtils.cuh:98-112; NULL,NULL = unit cube */

/* ---- stage-level seams used by the reference's unit tests ------------------------------------ */
/* Cubature_rules::apply_cubature_integration_rules (PaganiUtils.cuh:291-343): evaluate the current
 * region list; sums are returned, per-region arrays stay on the device (pgn_fetch_*). */
int pgn_evaluate(pgn_ctx* ctx, pgn_eval_fn fn, void* user, double* sum_estimate, double* sum_errorest);
/* two_level_errorest_and_relerr_classify (hybrid.cuh:11-45) + the reductions the driver needs. */
int pgn_classify(pgn_ctx* ctx, double epsrel, int relerr_classification, pgn_iter_stats* out);
/* Overwrite the active flags (host bytes, 0/1) -- what RegionFiltering.cu does by hand. */
int pgn_set_active(pgn_ctx* ctx, const uint8_t* flags, size_t n);
/* Upload per-region estimate / error / split axis by hand (RegionFiltering.cu:60-120 fills them so). */
int pgn_set_estimates(pgn_ctx* ctx, const double* estimate, const double* errorest, const int32_t* axis, size_t n);
/* Sub_regions_filter::filter (Sub_region_filter.cuh:85-139): stable compaction; returns survivors. */
int pgn_filter(pgn_ctx* ctx, size_t* num_active);
/* Sub_region_splitter::split (Sub_region_splitter.cuh:55-87). */
int pgn_split(pgn_ctx* ctx);
/* Heuristic_classifier::classify (heuristic_classifier.cuh:362-451) on the CURRENT error list, regardless of the
 * memory-pressure criteria (the reference's classifier test drives it directly,
 * test/kokkos/pagani/test_heuristic_classifier.cpp:32-103).  Needs pgn_classify / pgn_set_active first (statistics).
 * last_estimate is what Heuristic_classifier::store_estimate saw last; on success the active flags are replaced by
 * the ones the search adopted (those of its last probe). */
typedef struct pgn_heuristic_result {
  int32_t success;        /* pass_mem && pass_errorest_budget */
  int32_t must_terminate; /* Workspace::heuristic_classify's return value had the criteria been met */
  int32_t probes;
  int32_t reserved;
  uint64_t num_active;
  double threshold;        /* of the last probe */
  double finished_errorest;
} pgn_heuristic_result;
int pgn_heuristic_classify(pgn_ctx* ctx, double epsrel, double last_estimate, double iter_errorest,
                           double iter_finished_errorest, double total_finished_errorest, pgn_heuristic_result* out);
/* copies of the per-region arrays of the current iteration (any pointer may be NULL) */
int pgn_fetch_estimates(pgn_ctx* ctx, double* estimate, double* errorest, int32_t* axis, uint8_t* active, size_t cap);
int pgn_fetch_parents(pgn_ctx* ctx, double* estimate, double* errorest, int32_t* axis, size_t cap, size_t* n);

/* ---- the same stages on CALLER-OWNED device arrays ----------------------------------------------
 * What the reference's stage objects own and hand to each other: bounds as [ndim][n] doubles
 * (Sub_regions::dLeftCoord / dLength, Sub_regions.cuh:38-258), estimates (Region_estimates.cuh:8-38), split axes as
 * int and active flags as DOUBLES 0./1. (Region_characteristics.cuh:7-27).  include/pagani_b200/compat/ builds the
 * reference's classes (Cubature_rules, Sub_regions_filter, Sub_region_splitter ...) on these, so that code written
 * against the reference's stage interface -- its own unit tests -- compiles with only the include path changed.
 * Every call is synchronous (the reference's stages end in cudaDeviceSynchronize). */
int pgn_dev_uniform_split(pgn_ctx* ctx, int parts_per_axis, double* d_left, double* d_length,
                          size_t n /* = parts^ndim */); /* create_uniform_split, Sub_regions.cuh:11-36 */
int pgn_dev_evaluate(pgn_ctx* ctx, pgn_eval_fn fn, void* user, const double* d_left, const double* d_length, size_t n,
                     double* d_estimate, double* d_errorest, int32_t* d_axis, double* d_active /* set to 1., may be NULL */,
                     double* sum_estimate, double* sum_errorest); /* apply_cubature_integration_rules, PaganiUtils.cuh:291-343 */
int pgn_dev_count_active(pgn_ctx* ctx, const double* d_active, size_t n,
                         size_t* count); /* Sub_regions_filter::get_num_active_regions, Sub_region_filter.cuh:61-82 */
int pgn_dev_filter(pgn_ctx* ctx, const double* d_left, const double* d_length, const double* d_estimate,
                   const double* d_errorest, const int32_t* d_axis, const double* d_active, size_t n, size_t m /* survivors */,
                   double* out_left, double* out_length, double* out_estimate, double* out_errorest,
                   int32_t* out_axis); /* alignRegions, Sub_region_filter.cuh:8-47: stable compaction into [..][m] arrays */
int pgn_dev_split(pgn_ctx* ctx, const double* d_left, const double* d_length, const int32_t* d_axis, size_t m,
                  double* out_left, double* out_length); /* divideIntervalsGPU, Sub_region_splitter.cuh:8-46: [..][2m] */
int pgn_regions_import_device(pgn_ctx* ctx, const double* d_left, const double* d_length, size_t n);
int pgn_regions_export_device(pgn_ctx* ctx, double* d_left, double* d_length, size_t cap, size_t* n);

/* ---- the whole loop (Workspace::integrate, Workspace.cuh:361-494 / :148-358) ------------------ */
typedef struct pgn_integrate_args {
  double epsrel, epsabs;
  int relerr_classification; /* default 1 */
  int use_current_regions;   /* 0: uniform initial split like the 4-argument overload; 1: Sub_regions& overload */
  int iters_mode;            /* pgn_iters_mode */
  int reserved;
} pgn_integrate_args;
int pgn_integrate(pgn_ctx* ctx, pgn_eval_fn fn, void* user, const pgn_integrate_args* args, pgn_result_t* out);

/* Pre-instantiated Genz families (common/cuda/integrands.cuh:1303-1932 and the 10-D extensions of
 * SURVEY.md 8d).  family 1..6, ndim must be one this build instantiated (pgn_genz_available). */
int pgn_genz_available(int family, int ndim);
int pgn_genz_evaluator(int family, int ndim, pgn_eval_fn* fn, void** user);
int pgn_integrate_genz(pgn_ctx* ctx, int family, double epsrel, double epsabs, const double* lows,
                       const double* highs, pgn_result_t* out);
double pgn_genz_true_value(int family, int ndim);
double pgn_genz_flops_per_eval(int family, int ndim); /* algorithmic figure of SURVEY.md 8d */

/* ---- trace ------------------------------------------------------------------------------------ */
int pgn_trace_count(pgn_ctx* ctx);
int pgn_trace_get(pgn_ctx* ctx, int i, pgn_iter_stats* out);
uint64_t pgn_digest64(const void* data, size_t elem_bytes, size_t n); /* host helper, same formula as the device digests */

/* ---- multi-GPU (one process per GPU; NCCL over NVLink) ---------------------------------------- */
#define PGN_NCCL_ID_BYTES 128
int pgn_nccl_unique_id(uint8_t id[PGN_NCCL_ID_BYTES]);               /* rank 0, then broadcast by the launcher */
int pgn_comm_init(pgn_ctx* ctx, const uint8_t id[PGN_NCCL_ID_BYTES], int nranks, int rank);
int pgn_comm_size(pgn_ctx* ctx, int* nranks, int* rank);
int pgn_set_rebalance(pgn_ctx* ctx, double max_over_mean, size_t min_regions); /* default 1.10, 4096 */
/* test hook: replace the NCCL collectives by host callbacks (gloo in the CPU tests) */
typedef int (*pgn_allreduce_fn)(void* user, double* buf, int n, int op /*0 sum,1 max*/);
typedef int (*pgn_alltoallv_fn)(void* user, const double* send, const uint64_t* send_counts,
                                const uint64_t* send_offsets, double* recv, const uint64_t* recv_counts,
                                const uint64_t* recv_offsets); /* counts/offsets in doubles, one entry per rank */
int pgn_comm_set_host_collective(pgn_ctx* ctx, int nranks, int rank, pgn_allreduce_fn allreduce,
                                 pgn_alltoallv_fn alltoallv, void* user);
/* host-only planner of the order-preserving repartition (unit-testable without a GPU):
 * counts[nranks] -> new_counts[nranks] and up to 2*nranks transfers {src,dst,src_off,dst_off,count} */
int pgn_plan_rebalance(const uint64_t* counts, int nranks, uint64_t* new_counts, uint64_t* transfers /*[5*max]*/,
                       int max_transfers, int* ntransfers);

/* ---- micro-benchmarks used by bench.py for the roofline denominators -------------------------- */
int pgn_bench_fp64_peak(int device, int iters, double* tflops);       /* dependent-free DFMA stream */
int pgn_bench_hbm_copy(int device, size_t bytes, int iters, double* gbs);

#ifdef __cplusplus
}
#endif
#endif /* PAGANI_B200_H */
