// Host driver of the B200-native PAGANI engine: region store (ping-pong SoA arena), the iteration
// loop of Workspace::integrate, the heuristic classifier and the multi-GPU partitioner.
// Reference: cuda/pagani/quad/GPUquad/Workspace.cuh:74-494, heuristic_classifier.cuh:147-452.
#ifndef PGN_ENGINE_H
#define PGN_ENGINE_H

#include <cuda_runtime.h>

#include <array>
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/pagani_b200.h"
#include "comm.h"
#include "rule.h"

namespace pgn {

void set_last_error(const std::string& msg);
const char* last_error();

// A device allocation that only ever grows (no per-iteration cudaMalloc/cudaFree churn, unlike the
// reference, which re-allocates ~25 arrays per iteration).
struct DeviceBuf {
  void* ptr = nullptr;
  size_t bytes = 0;
  // contents are NOT preserved
  int reserve(size_t need);
  void release();
  template <typename T>
  T* as() const
  {
    return static_cast<T*>(ptr);
  }
};

// Restatement of Heuristic_classifier's host state (heuristic_classifier.cuh:147-225).
struct ConvergenceTracker {
  double epsrel = 0, epsabs = 0;
  int required_digits = 0;
  double last[3] = {0.0, 0.0, 0.0};  // the reference leaves these uninitialised (:153)
  size_t collected = 0;
  void reset(double rel, double abs_);
  void store(double estimate);
  bool sig_digits_same() const;
  bool converged() const;
};

class Engine {
public:
  Engine(int ndim, const pgn_opts& opts);
  ~Engine();
  int init();  // device selection, rule upload; PGN_ERR_NO_DEVICE when there is no usable GPU

  int ndim() const { return ndim_; }
  const RuleTables& rule() const { return rule_; }
  int rule_view(pgn_rule_view* out) const;
  int device_info(int* sm_count, size_t* free_b, size_t* total_b);

  int set_volume(const double* lows, const double* highs);
  int regions_init_uniform(int parts);
  int regions_import(const double* left, const double* length, size_t n);
  int regions_export(double* left, double* length, size_t cap, size_t* n);
  size_t regions_count() const { return n_; }

  int evaluate(pgn_eval_fn fn, void* user, double* sum_est, double* sum_err);
  int classify(double epsrel, int relerr, pgn_iter_stats* out);
  int set_active(const uint8_t* flags, size_t n);
  int set_estimates(const double* est, const double* err, const int32_t* axis, size_t n);
  int filter(size_t* num_active);
  int heuristic_classify_stage(double epsrel, double last_estimate, double iter_errorest,
                               double iter_finished_errorest, double total_finished_errorest,
                               pgn_heuristic_result* out);
  int split();
  int fetch_estimates(double* est, double* err, int32_t* axis, uint8_t* active, size_t cap);
  int fetch_parents(double* est, double* err, int32_t* axis, size_t cap, size_t* n);

  int integrate(pgn_eval_fn fn, void* user, const pgn_integrate_args& args, pgn_result_t* out);

  // ---- the same stages on CALLER-OWNED device arrays in the reference's layouts (what its stage objects
  // Cubature_rules / Sub_regions_filter / Sub_region_splitter bind to; include/pagani_b200/compat)
  int dev_uniform_split(int parts, double* d_left, double* d_length, size_t n);
  int dev_evaluate(pgn_eval_fn fn, void* user, const double* d_left, const double* d_length, size_t n,
                   double* d_est, double* d_err, int32_t* d_axis, double* d_active_f64, double* sum_est,
                   double* sum_err);
  int dev_count_active(const double* d_active_f64, size_t n, size_t* count);
  int dev_filter(const double* d_left, const double* d_length, const double* d_est, const double* d_err,
                 const int32_t* d_axis, const double* d_active_f64, size_t n, size_t m, double* out_left,
                 double* out_length, double* out_est, double* out_err, int32_t* out_axis);
  int dev_split(const double* d_left, const double* d_length, const int32_t* d_axis, size_t m, double* out_left,
                double* out_length);
  int regions_import_device(const double* d_left, const double* d_length, size_t n);
  int regions_export_device(double* d_left, double* d_length, size_t cap, size_t* n);

  int comm_init_nccl(const uint8_t* id, int nranks, int rank);
  int comm_init_host(int nranks, int rank, pgn_allreduce_fn ar, pgn_alltoallv_fn a2a, void* user);
  int comm_size(int* nranks, int* rank) const;
  void set_eval_variant(int v) { eval_variant_ = v; }
  void set_rebalance(double max_over_mean, size_t min_regions)
  {
    rebalance_ratio_ = max_over_mean;
    rebalance_min_ = min_regions;
  }

  uint64_t launches() const { return launches_; }
  const std::vector<pgn_iter_stats>& trace() const { return trace_; }
  cudaStream_t stream() const { return stream_; }

private:
  // ---- stage helpers (all enqueue on stream_)
  int launch_evaluate(pgn_eval_fn fn, void* user);
  int launch_evaluate_on(pgn_eval_fn fn, void* user, const double* left, const double* length, size_t n,
                         double* est, double* err, int32_t* axis);
  // reduce + (all-reduce) + read back; fills host_local_/host_global_
  int run_classify(double epsrel, int relerr);
  int run_threshold(int mode, double threshold);
  int run_recount();
  int sync_stats(bool need_minmax);
  int compact_and_split(size_t m_local, size_t m_global, bool* rebalanced);
  int rebalance_parents(size_t* m_local);
  int ensure_iteration_buffers(size_t n);
  int tiles_for(size_t n) const;
  void digest_arrays(pgn_iter_stats* st, bool after_classify);
  uint64_t digest_device(const void* p, size_t elem, size_t n);

  // heuristic classifier (Workspace::heuristic_classify + Heuristic_classifier::classify)
  struct HeuristicOutcome {
    bool ran = false;
    bool success = false;
    bool must_terminate = false;
    size_t num_active = 0;
    double finished_errorest = 0;
    double finished_estimate = 0;
    double threshold = 0;
    int probes = 0;
  };
  int heuristic_classify(size_t n_global, double iter_est, double iter_err, double fin_err,
                         double cum_err, HeuristicOutcome* out, bool force = false);
  // bucket partition of the current errors for the device-side threshold search
  int prepare_search(const struct BucketPlan& plan, struct BucketTables* tables, struct BucketTables* global);
  double mem_capacity_bytes() const;  // this rank's figure
  int agree_mem_capacity();            // -> mem_capacity_, identical on every rank
  double mem_capacity_ = 16e9;

  int ndim_;
  pgn_opts opts_;
  int device_ = 0;
  int sm_count_ = 148;
  int search_attr_device_ = -1; // device on which k_search's shared-memory attribute was set by this engine
  int pairs_attr_device_ = -1;  // device on which k_classify_pairs' shared-memory attribute was last set by this engine
  int eval_variant_ = 0;  // pgn_region_view::reserved: 0 default, 1 shift/mask, 2 byte-permute
  cudaStream_t stream_ = nullptr;
  bool ready_ = false;

  RuleTables rule_;
  DeviceBuf d_codes_, d_weights_, d_scale_, d_norm_;

  double lows_[PGN_MAX_NDIM + 1];
  double ranges_[PGN_MAX_NDIM + 1];
  double jacobian_ = 1.0;

  // region store: current list and the list being built for the next iteration
  DeviceBuf left_[2], length_[2];
  int cur_ = 0;
  size_t n_ = 0;  // local regions in the current list
  // per-region results of the current iteration
  DeviceBuf est_, err_raw_, err_ref_, axis_, active_;
  bool evaluated_ = false, classified_ = false;
  const double* err_cur_ = nullptr;  // refined errors when parents exist, raw otherwise
  // parents (previous iteration's surviving estimates), double buffered
  DeviceBuf par_est_[2], par_err_[2], par_axis_, par_axis_next_;
  int par_cur_ = 0;
  size_t parents_n_ = 0;  // 0: no parents (iteration 0)
  bool filtered_pending_ = false;  // filter() done, split() not yet: cur list = compacted parents
  // reductions
  DeviceBuf partials_, tile_count_, tile_offset_, done_, stats_dev_, stats_glob_, digest_dev_;
  double* host_local_ = nullptr;   // pinned [kHostStats]
  double* host_global_ = nullptr;  // pinned [kHostStats]
  // zero-copy statistics: host-mapped [2 * kHostStats doubles + 1 word], written by the statistics kernels
  double* host_pub_ = nullptr;
  double* host_pub_dev_ = nullptr;       // device address of host_pub_
  unsigned long long pub_seq_ = 0;       // number of the last publishing kernel enqueued
  bool stats_published_ = false;         // the last statistics kernel publishes (sync_stats polls instead of copying)
  int wait_published();
  std::vector<size_t> rank_counts_;  // active regions per rank after the last flag-changing stage
  // rebalance staging
  DeviceBuf pack_send_, pack_recv_;
  // heuristic threshold search: errors grouped by value bucket, bucket tables, search in/out block
  DeviceBuf sorted_err_, search_tab_, search_io_;
  DeviceBuf search_gtab_, search_gd_;  // several ranks: tables of the concatenated list, their all-reduce buffer
  void* host_search_ = nullptr;  // pinned mirror of search_io_ (SearchIO + the `outside` counter)
  bool device_search_ = true;    // env PGN_HOST_SEARCH=1 forces the probe-per-pass search
  // NVLink peer mailboxes for the in-kernel all-reduce of the multi-rank search (threshold_search.cuh)
  int setup_peer_mailboxes();
  void release_peer_mailboxes();
  DeviceBuf mailbox_;
  std::vector<void*> peer_mapped_;       // cudaIpcOpenMemHandle results (nullptr for the own rank)
  bool peer_ready_ = false;
  unsigned long long peer_seq_ = 0;      // probes exchanged so far (identical on every rank)
  unsigned long long stats_seq_ = 0;     // fused statistics all-reduces so far
  bool stats_fused_ = false;             // the last statistics kernel already produced the global vector
  struct TileScratch make_scratch(bool allow_fused = true);  // false: rank-local statistics (stage API)
  void fill_peer_box(struct PeerBox* box, long long spin_limit) const;

  ConvergenceTracker tracker_;
  double max_percent_error_budget_ = .25;
  double max_active_regions_percentage_ = .5;

  Comm* comm_ = nullptr;
  double rebalance_ratio_ = 1.10;
  size_t rebalance_min_ = 4096;

  uint64_t launches_ = 0;  // kernels of this library enqueued so far (bench.py's gpu_launches)
  std::vector<pgn_iter_stats> trace_;
  std::vector<cudaEvent_t> ev_pool_;  // stage-timing events of a traced integrate(): six per iteration
};

}  // namespace pgn
#endif
