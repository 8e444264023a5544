#include "engine.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>

#include "stages.cuh"
#include "threshold_search.cuh"

namespace pgn {

namespace {
thread_local std::string g_last_error;

#define PGN_CUDA(expr)                                                                             \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      set_last_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                         \
      return _e == cudaErrorMemoryAllocation ? PGN_ERR_NOMEM : PGN_ERR_CUDA;                       \
    }                                                                                              \
  } while (0)

#define PGN_TRY(expr)                                                                              \
  do {                                                                                             \
    int _rc = (expr);                                                                              \
    if (_rc != PGN_OK)                                                                             \
      return _rc;                                                                                  \
  } while (0)

constexpr int kHostStats = kStatsLen;

// pack / unpack of the columns that travel with a region during a rebalance:
// 2*ndim bounds columns, parent estimate, parent error, split axis (as a double)
__global__ void
k_pack(const double* __restrict__ left, const double* __restrict__ length,
       const double* __restrict__ pest, const double* __restrict__ perr,
       const int32_t* __restrict__ axis, size_t m, int ndim, size_t first, size_t count,
       double* __restrict__ out)
{
  const size_t cols = 2 * static_cast<size_t>(ndim) + 3;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < cols * count;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t c = t / count, j = t - c * count, i = first + j;
    double v;
    if (c < static_cast<size_t>(ndim))
      v = left[c * m + i];
    else if (c < 2 * static_cast<size_t>(ndim))
      v = length[(c - ndim) * m + i];
    else if (c == 2 * static_cast<size_t>(ndim))
      v = pest[i];
    else if (c == 2 * static_cast<size_t>(ndim) + 1)
      v = perr[i];
    else
      v = static_cast<double>(axis[i]);
    out[t] = v;
  }
}

__global__ void
k_unpack(const double* __restrict__ in, size_t count, size_t first, size_t m_new, int ndim,
         double* __restrict__ left, double* __restrict__ length, double* __restrict__ pest,
         double* __restrict__ perr, int32_t* __restrict__ axis)
{
  const size_t cols = 2 * static_cast<size_t>(ndim) + 3;
  for (size_t t = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; t < cols * count;
       t += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t c = t / count, j = t - c * count, i = first + j;
    const double v = in[t];
    if (c < static_cast<size_t>(ndim))
      left[c * m_new + i] = v;
    else if (c < 2 * static_cast<size_t>(ndim))
      length[(c - ndim) * m_new + i] = v;
    else if (c == 2 * static_cast<size_t>(ndim))
      pest[i] = v;
    else if (c == 2 * static_cast<size_t>(ndim) + 1)
      perr[i] = v;
    else
      axis[i] = static_cast<int32_t>(v);
  }
}

unsigned
blocks_for(size_t work, int block)
{
  size_t b = (work + block - 1) / block;
  if (b < 1)
    b = 1;
  if (b > 148u * 32u)
    b = 148u * 32u;
  return static_cast<unsigned>(b);
}

}  // namespace

void
set_last_error(const std::string& msg)
{
  g_last_error = msg;
}
const char*
last_error()
{
  return g_last_error.c_str();
}

// ------------------------------------------------------------------------------------------------
int
DeviceBuf::reserve(size_t need)
{
  if (need <= bytes)
    return PGN_OK;
  // grow with head-room: the region list can double every iteration
  size_t want = need + need / 4;
  if (want < 4096)
    want = 4096;
  if (ptr) {
    cudaFree(ptr);
    ptr = nullptr;
    bytes = 0;
  }
  cudaError_t e = cudaMalloc(&ptr, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    e = cudaMalloc(&ptr, need);
    want = need;
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    ptr = nullptr;
    set_last_error("device allocation of " + std::to_string(need) + " bytes failed");
    return PGN_ERR_NOMEM;
  }
  bytes = want;
  return PGN_OK;
}

void
DeviceBuf::release()
{
  if (ptr)
    cudaFree(ptr);
  ptr = nullptr;
  bytes = 0;
}

// ------------------------------------------------------------------------------------------------
void
ConvergenceTracker::reset(double rel, double abs_)
{
  epsrel = rel;
  epsabs = abs_;
  required_digits = static_cast<int>(std::ceil(std::log10(1 / epsrel)));
  last[0] = last[1] = last[2] = 0.0;
  collected = 0;
}

void
ConvergenceTracker::store(double estimate)
{
  last[0] = last[1];
  last[1] = last[2];
  last[2] = estimate;
  collected++;
}

// heuristic_classifier.cuh:169-203: the last three estimates, scaled up to >= 1, printed with 15
// decimals, must share their first `required_digits` significant characters.
bool
ConvergenceTracker::sig_digits_same() const
{
  auto norm = [](double v) {
    v = std::fabs(v);
    while (v != 0. && v < 1.)
      v *= 10;
    return v;
  };
  auto fixed15 = [](double v) {
    char buf[512];
    snprintf(buf, sizeof buf, "%.15f", v);
    return std::string(buf);
  };
  const std::string second_to_last = fixed15(norm(last[0]));
  const std::string prev = fixed15(norm(last[1]));
  const std::string current = fixed15(norm(last[2]));
  auto at = [](const std::string& s, int i) {
    return i < static_cast<int>(s.size()) ? s[i] : '\0';
  };
  bool verdict = true;
  int sig = 0;
  for (int i = 0; i < required_digits + 1 && sig < required_digits && verdict; ++i) {
    verdict = at(current, i) == at(prev, i) && at(prev, i) == at(second_to_last, i);
    sig += (verdict && at(current, i) != '.') ? 1 : 0;
  }
  return verdict;
}

bool
ConvergenceTracker::converged() const
{
  // heuristic_classifier.cuh:205-216 (min_iters_for_convergence = 1, unsigned arithmetic)
  if (collected - 1 < 1 || !sig_digits_same())
    return false;
  return true;
}

// ------------------------------------------------------------------------------------------------
Engine::Engine(int ndim, const pgn_opts& opts) : ndim_(ndim), opts_(opts)
{
  for (int d = 0; d <= PGN_MAX_NDIM; ++d) {
    lows_[d] = 0.0;
    ranges_[d] = 1.0;
  }
}

Engine::~Engine()
{
  if (ready_) {
    cudaSetDevice(device_);
    cudaStreamSynchronize(stream_);
  }
  release_peer_mailboxes();
  delete comm_;
  DeviceBuf* all[] = {&d_codes_,   &d_weights_, &d_scale_,    &d_norm_,     &left_[0],    &left_[1],
                      &length_[0], &length_[1], &est_,        &err_raw_,    &err_ref_,    &axis_,
                      &active_,    &par_est_[0], &par_est_[1], &par_err_[0], &par_err_[1], &par_axis_,
                      &partials_,  &tile_count_, &tile_offset_, &done_,      &stats_dev_,  &stats_glob_,
                      &digest_dev_, &pack_send_, &pack_recv_, &sorted_err_, &search_tab_, &search_io_, &mailbox_,
                      &search_gtab_, &search_gd_, &par_axis_next_};
  for (DeviceBuf* b : all)
    b->release();
  if (host_local_)
    cudaFreeHost(host_local_);
  if (host_global_)
    cudaFreeHost(host_global_);
  if (host_pub_)
    cudaFreeHost(host_pub_);
  if (host_search_)
    cudaFreeHost(host_search_);
  for (cudaEvent_t e : ev_pool_)
    cudaEventDestroy(e);
  if (stream_)
    cudaStreamDestroy(stream_);
}

int
Engine::init()
{
  if (!build_rule(ndim_, &rule_)) {
    set_last_error("ndim must be in [2, 15]");
    return PGN_ERR_INVALID;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    cudaGetLastError();
    set_last_error("no CUDA device: pagani_b200 has no CPU fallback");
    return PGN_ERR_NO_DEVICE;
  }
  if (opts_.device >= 0)
    device_ = opts_.device;
  else
    PGN_CUDA(cudaGetDevice(&device_));
  if (device_ >= count) {
    set_last_error("device ordinal out of range");
    return PGN_ERR_INVALID;
  }
  PGN_CUDA(cudaSetDevice(device_));
  cudaDeviceProp prop;
  PGN_CUDA(cudaGetDeviceProperties(&prop, device_));
  if (prop.major < 10) {
    set_last_error(std::string("device ") + prop.name +
                   " is not sm_100 or newer: this build carries sm_100a code only");
    return PGN_ERR_NO_DEVICE;
  }
  sm_count_ = prop.multiProcessorCount;
  PGN_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));

  PGN_TRY(d_codes_.reserve(sizeof(uint64_t) * rule_.feval));
  PGN_TRY(d_weights_.reserve(sizeof(double) * kSets * kRules));
  PGN_TRY(d_scale_.reserve(sizeof(double) * kSets * kRules));
  PGN_TRY(d_norm_.reserve(sizeof(double) * kSets * kRules));
  PGN_CUDA(cudaMemcpy(d_codes_.ptr, rule_.point_code.data(), sizeof(uint64_t) * rule_.feval,
                      cudaMemcpyHostToDevice));
  PGN_CUDA(cudaMemcpy(d_weights_.ptr, rule_.weights, sizeof(double) * kSets * kRules,
                      cudaMemcpyHostToDevice));
  PGN_CUDA(cudaMemcpy(d_scale_.ptr, rule_.scale, sizeof(double) * kSets * kRules,
                      cudaMemcpyHostToDevice));
  PGN_CUDA(cudaMemcpy(d_norm_.ptr, rule_.norm, sizeof(double) * kSets * kRules,
                      cudaMemcpyHostToDevice));

  PGN_TRY(done_.reserve(sizeof(unsigned int)));
  PGN_CUDA(cudaMemset(done_.ptr, 0, sizeof(unsigned int)));
  PGN_TRY(stats_dev_.reserve(sizeof(double) * kStatsLen));
  PGN_TRY(stats_glob_.reserve(sizeof(double) * kStatsLen));
  PGN_TRY(digest_dev_.reserve(sizeof(unsigned long long) * 8));
  PGN_CUDA(cudaMallocHost(&host_local_, sizeof(double) * kHostStats));
  PGN_CUDA(cudaMallocHost(&host_global_, sizeof(double) * kHostStats));
  if (!std::getenv("PGN_NO_ZERO_COPY_STATS")) {
    void* pub = nullptr;
    if (cudaHostAlloc(&pub, sizeof(double) * (2 * kHostStats + 1), cudaHostAllocMapped) == cudaSuccess) {
      std::memset(pub, 0, sizeof(double) * (2 * kHostStats + 1));
      void* dev = nullptr;
      if (cudaHostGetDevicePointer(&dev, pub, 0) == cudaSuccess) {
        host_pub_ = static_cast<double*>(pub);
        host_pub_dev_ = static_cast<double*>(dev);
      } else {
        cudaGetLastError();
        cudaFreeHost(pub);
      }
    } else {
      cudaGetLastError();
    }
  }
  rank_counts_.assign(1, 0);
  // bucket tables: count, start(+1), cursor, bucket_sum, suffix_sum(+1), suffix_count(+1), outside, acc_lo, acc_hi
  PGN_TRY(search_tab_.reserve(8 * (8 * static_cast<size_t>(kMaxBuckets) + 8)));
  PGN_TRY(search_io_.reserve(sizeof(SearchIO) + 16));
  PGN_CUDA(cudaMallocHost(&host_search_, sizeof(SearchIO) + 16));
  if (const char* hs = std::getenv("PGN_HOST_SEARCH"))
    device_search_ = std::atoi(hs) == 0;
  if (const char* ev = std::getenv("PGN_EVAL_VARIANT"))
    eval_variant_ = std::atoi(ev);
  ready_ = true;
  return PGN_OK;
}

int
Engine::rule_view(pgn_rule_view* out) const
{
  if (!out)
    return PGN_ERR_INVALID;
  out->ndim = ndim_;
  out->feval = rule_.feval;
  out->point_code = d_codes_.as<uint64_t>();
  out->weights = d_weights_.as<double>();
  out->scale = d_scale_.as<double>();
  out->norm = d_norm_.as<double>();
  for (int c = 0; c < kCodes; ++c) {
    out->half_plus_g[c] = rule_.half_plus_g[c];
    out->half_minus_g[c] = rule_.half_minus_g[c];
  }
  out->ratio = rule_.ratio;
  out->two_one_minus_ratio_unused = 0.0;
  for (int i = 0; i < kSets * kRules; ++i)
    out->weights_host[i] = rule_.weights[i];
  return PGN_OK;
}

int
Engine::device_info(int* sm_count, size_t* free_b, size_t* total_b)
{
  PGN_CUDA(cudaSetDevice(device_));
  size_t f = 0, t = 0;
  PGN_CUDA(cudaMemGetInfo(&f, &t));
  if (sm_count)
    *sm_count = sm_count_;
  if (free_b)
    *free_b = f;
  if (total_b)
    *total_b = t;
  return PGN_OK;
}

int
Engine::set_volume(const double* lows, const double* highs)
{
  // PaganiUtils.cuh:98-112 + the per-block prologue of INIT_REGION_POOL (Phases.cuh:338-343):
  // ranges and their running product (the Jacobian) are launch constants, computed here once.
  double jac = 1.0;
  for (int d = 0; d < ndim_; ++d) {
    const double lo = lows ? lows[d] : 0.0;
    const double hi = highs ? highs[d] : 1.0;
    lows_[d] = lo;
    ranges_[d] = hi - lo;
    jac = jac * ranges_[d];
  }
  jacobian_ = jac;
  return PGN_OK;
}

int
Engine::tiles_for(size_t n) const
{
  return static_cast<int>((n + kTile - 1) / kTile);
}

int
Engine::ensure_iteration_buffers(size_t n)
{
  const size_t nn = n > 0 ? n : 1;
  PGN_TRY(est_.reserve(sizeof(double) * nn));
  PGN_TRY(err_raw_.reserve(sizeof(double) * nn));
  PGN_TRY(err_ref_.reserve(sizeof(double) * nn));
  PGN_TRY(axis_.reserve(sizeof(int32_t) * nn));
  PGN_TRY(active_.reserve(nn));
  const size_t tiles = static_cast<size_t>(tiles_for(nn));
  PGN_TRY(partials_.reserve(sizeof(double) * kStatCount * tiles));
  PGN_TRY(tile_count_.reserve(sizeof(uint32_t) * tiles));
  PGN_TRY(tile_offset_.reserve(sizeof(uint64_t) * tiles));
  return PGN_OK;
}

int
Engine::regions_init_uniform(int parts)
{
  if (parts < 1)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  // Sub_regions::uniform_split (Sub_regions.cuh:164-185); with several ranks every rank builds its
  // contiguous slice of the global grid.
  double total_d = 1.0;
  for (int d = 0; d < ndim_; ++d)
    total_d *= parts;
  if (total_d > 4e9) {
    set_last_error("initial grid too large");
    return PGN_ERR_INVALID;
  }
  const size_t total = static_cast<size_t>(total_d + 0.5);
  const int w = comm_ ? comm_->nranks : 1, r = comm_ ? comm_->rank : 0;
  const size_t lo = total * r / w, hi = total * (r + 1) / w;
  const size_t n = hi - lo;
  const double edge = 1. / static_cast<double>(parts);
  PGN_TRY(left_[cur_].reserve(sizeof(double) * std::max<size_t>(n, 1) * ndim_));
  PGN_TRY(length_[cur_].reserve(sizeof(double) * std::max<size_t>(n, 1) * ndim_));
  if (w == 1) {
    if (n > 0)
      ++launches_, k_uniform_split<<<blocks_for(n, 512), 512, 0, stream_>>>(
        left_[cur_].as<double>(), length_[cur_].as<double>(), n, parts, ndim_, edge);
    PGN_CUDA(cudaGetLastError());
  } else {
    // small: build the slice on the host with the same arithmetic
    std::vector<double> hl(n * ndim_), hw(n * ndim_);
    for (size_t t = lo; t < hi; ++t) {
      size_t q = t;
      for (int d = 0; d < ndim_; ++d) {
        const size_t id = q % parts;
        q /= parts;
        hl[static_cast<size_t>(d) * n + (t - lo)] = static_cast<double>(id) * edge;
        hw[static_cast<size_t>(d) * n + (t - lo)] = edge;
      }
    }
    if (n > 0) {
      PGN_CUDA(cudaMemcpyAsync(left_[cur_].ptr, hl.data(), sizeof(double) * hl.size(),
                               cudaMemcpyHostToDevice, stream_));
      PGN_CUDA(cudaMemcpyAsync(length_[cur_].ptr, hw.data(), sizeof(double) * hw.size(),
                               cudaMemcpyHostToDevice, stream_));
      PGN_CUDA(cudaStreamSynchronize(stream_));
    }
  }
  n_ = n;
  parents_n_ = 0;
  evaluated_ = classified_ = filtered_pending_ = false;
  return PGN_OK;
}

int
Engine::regions_import(const double* left, const double* length, size_t n)
{
  if ((!left || !length) && n > 0)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  PGN_TRY(left_[cur_].reserve(sizeof(double) * std::max<size_t>(n, 1) * ndim_));
  PGN_TRY(length_[cur_].reserve(sizeof(double) * std::max<size_t>(n, 1) * ndim_));
  if (n > 0) {
    PGN_CUDA(cudaMemcpyAsync(left_[cur_].ptr, left, sizeof(double) * n * ndim_,
                             cudaMemcpyHostToDevice, stream_));
    PGN_CUDA(cudaMemcpyAsync(length_[cur_].ptr, length, sizeof(double) * n * ndim_,
                             cudaMemcpyHostToDevice, stream_));
    PGN_CUDA(cudaStreamSynchronize(stream_));
  }
  n_ = n;
  parents_n_ = 0;
  evaluated_ = classified_ = filtered_pending_ = false;
  return PGN_OK;
}

int
Engine::regions_export(double* left, double* length, size_t cap, size_t* n)
{
  if (n)
    *n = n_;
  if (cap < n_)
    return left || length ? PGN_ERR_INVALID : PGN_OK;
  PGN_CUDA(cudaSetDevice(device_));
  if (n_ > 0 && left)
    PGN_CUDA(cudaMemcpyAsync(left, left_[cur_].ptr, sizeof(double) * n_ * ndim_,
                             cudaMemcpyDeviceToHost, stream_));
  if (n_ > 0 && length)
    PGN_CUDA(cudaMemcpyAsync(length, length_[cur_].ptr, sizeof(double) * n_ * ndim_,
                             cudaMemcpyDeviceToHost, stream_));
  PGN_CUDA(cudaStreamSynchronize(stream_));
  return PGN_OK;
}

// ------------------------------------------------------------------------------------------------
int
Engine::launch_evaluate(pgn_eval_fn fn, void* user)
{
  PGN_TRY(ensure_iteration_buffers(n_));
  if (n_ == 0)
    return PGN_OK;
  return launch_evaluate_on(fn, user, left_[cur_].as<double>(), length_[cur_].as<double>(), n_, est_.as<double>(),
                            err_raw_.as<double>(), axis_.as<int32_t>());
}

int
Engine::launch_evaluate_on(pgn_eval_fn fn, void* user, const double* left, const double* length, size_t n,
                           double* est, double* err, int32_t* axis)
{
  pgn_region_view v;
  std::memset(&v, 0, sizeof v);
  v.abi_version = PGN_ABI_VERSION;
  v.view_bytes = static_cast<uint32_t>(sizeof v);
  v.left = left;
  v.length = length;
  v.n = n;
  v.estimate = est;
  v.errorest = err;
  v.axis = axis;
  for (int d = 0; d <= PGN_MAX_NDIM; ++d) {
    v.lows[d] = lows_[d];
    v.ranges[d] = ranges_[d];
  }
  v.jacobian = jacobian_;
  rule_view(&v.rule);
  v.stream = stream_;
  v.sm_count = sm_count_;
  v.reserved = eval_variant_;
  ++launches_;
  const int rc = fn(user, &v);
  if (rc != PGN_OK) {
    cudaError_t e = cudaGetLastError();
    set_last_error(std::string("evaluation launch failed: ") + cudaGetErrorString(e));
  }
  return rc;
}

// Wait until the statistics kernel number pub_seq_ has raised its flag in host-mapped memory.  The stream is queried
// now and then so that a failed launch or a sticky error ends the wait.
int
Engine::wait_published()
{
  volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(host_pub_ + 2 * kHostStats);
  unsigned long spins = 0;
  while (*flag != pub_seq_) {
    if ((++spins & 0x3fffu) == 0) {
      const cudaError_t q = cudaStreamQuery(stream_);
      if (q == cudaSuccess) {
        if (*flag == pub_seq_)
          break;
        set_last_error("statistics kernel finished without publishing its result");
        return PGN_ERR_CUDA;
      }
      if (q != cudaErrorNotReady) {
        set_last_error(std::string("statistics kernel failed: ") + cudaGetErrorString(q));
        return PGN_ERR_CUDA;
      }
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  return PGN_OK;
}

int
Engine::sync_stats(bool need_minmax)
{
  const bool multi = comm_ && comm_->nranks > 1;
  if (stats_published_ && (!multi || stats_fused_)) {
    // zero-copy path: both vectors are already (or will shortly be) in pinned host memory
    PGN_TRY(wait_published());
    std::memcpy(host_local_, host_pub_, sizeof(double) * kStatsLen);
    std::memcpy(host_global_, host_pub_ + kStatsLen, sizeof(double) * kStatsLen);
    stats_published_ = false;
    if (multi && std::isnan(host_global_[kNumActive])) {
      set_last_error("fused statistics all-reduce: a peer did not answer within the spin limit");
      return PGN_ERR_NCCL;
    }
    const int w = comm_ ? comm_->nranks : 1;
    rank_counts_.resize(w);
    for (int r = 0; r < w; ++r)
      rank_counts_[r] = static_cast<size_t>(host_global_[kRankSlot + r]);
    return PGN_OK;
  }
  stats_published_ = false;
  if (multi && stats_fused_) {
    // the statistics kernel all-reduced its vector itself (sums and [-min, max]) over NVLink
    PGN_CUDA(cudaMemcpyAsync(host_global_, stats_glob_.ptr, sizeof(double) * kStatsLen,
                             cudaMemcpyDeviceToHost, stream_));
    need_minmax = true;
  } else if (multi) {
    PGN_CUDA(cudaMemcpyAsync(stats_glob_.ptr, stats_dev_.ptr, sizeof(double) * kStatsLen,
                             cudaMemcpyDeviceToDevice, stream_));
    PGN_TRY(comm_->allreduce(stats_glob_.as<double>(), kStatsLen, kOpSum, stream_));
    if (need_minmax) {
      PGN_CUDA(cudaMemcpyAsync(stats_glob_.as<double>() + kNegMinSlot,
                               stats_dev_.as<double>() + kNegMinSlot, sizeof(double) * 2,
                               cudaMemcpyDeviceToDevice, stream_));
      PGN_TRY(comm_->allreduce(stats_glob_.as<double>() + kNegMinSlot, 2, kOpMax, stream_));
    }
    PGN_CUDA(cudaMemcpyAsync(host_global_, stats_glob_.ptr, sizeof(double) * kStatsLen,
                             cudaMemcpyDeviceToHost, stream_));
  }
  PGN_CUDA(cudaMemcpyAsync(host_local_, stats_dev_.ptr, sizeof(double) * kStatsLen,
                           cudaMemcpyDeviceToHost, stream_));
  PGN_CUDA(cudaStreamSynchronize(stream_));
  if (!multi)
    std::memcpy(host_global_, host_local_, sizeof(double) * kStatsLen);
  else if (!need_minmax) {
    host_global_[kNegMinSlot] = std::numeric_limits<double>::quiet_NaN();
    host_global_[kMaxSlot] = std::numeric_limits<double>::quiet_NaN();
  }
  if (multi && stats_fused_ && std::isnan(host_global_[kNumActive])) {
    set_last_error("fused statistics all-reduce: a peer did not answer within the spin limit");
    return PGN_ERR_NCCL;
  }
  const int w = comm_ ? comm_->nranks : 1;
  rank_counts_.resize(w);
  for (int r = 0; r < w; ++r)
    rank_counts_[r] = static_cast<size_t>(host_global_[kRankSlot + r]);
  return PGN_OK;
}

namespace {
// statistics of an empty local list (a rank that ran out of regions still joins the collectives)
int
write_empty_stats(double* dev, cudaStream_t s)
{
  double h[kStatsLen];
  for (int i = 0; i < kStatsLen; ++i)
    h[i] = 0.0;
  h[kMinErr] = std::numeric_limits<double>::infinity();
  h[kMaxErr] = -std::numeric_limits<double>::infinity();
  h[kNegMinSlot] = -std::numeric_limits<double>::infinity();
  h[kMaxSlot] = -std::numeric_limits<double>::infinity();
  // pageable source: the copy is staged before the call returns
  return cudaMemcpyAsync(dev, h, sizeof h, cudaMemcpyHostToDevice, s) == cudaSuccess ? PGN_OK
                                                                                       : PGN_ERR_CUDA;
}
}  // namespace

void
Engine::fill_peer_box(PeerBox* box, long long spin_limit) const
{
  std::memset(box, 0, sizeof *box);
  box->nranks = comm_->nranks;
  box->rank = comm_->rank;
  box->spin_limit = spin_limit;
  for (int q = 0; q < comm_->nranks; ++q)
    box->base[q] = static_cast<char*>(q == comm_->rank ? mailbox_.ptr : peer_mapped_[q]);
}

// Scratch of one statistics kernel.  With peer mailboxes the kernel's last block all-reduces the
// statistics itself; every rank launches the same sequence of statistics kernels, so the exchange
// numbers agree.
TileScratch
Engine::make_scratch(bool allow_fused)
{
  TileScratch ts;
  std::memset(&ts, 0, sizeof ts);
  ts.partials = partials_.as<double>();
  ts.tile_count = tile_count_.as<uint32_t>();
  ts.tile_offset = tile_offset_.as<uint64_t>();
  ts.done = done_.as<unsigned int>();
  ts.stats = stats_dev_.as<double>();
  ts.rank = comm_ ? comm_->rank : 0;
  ts.nranks = comm_ ? comm_->nranks : 1;
  ts.stats_glob = stats_glob_.as<double>();
  ts.host_out = host_pub_dev_;
  ts.host_seq = host_pub_dev_ ? ++pub_seq_ : 0;
  stats_published_ = host_pub_dev_ != nullptr;
  stats_fused_ = false;
  if (allow_fused && comm_ && comm_->nranks > 1 && peer_ready_) {
    // A peer may still be inside a long evaluation kernel (seconds on a list of 10^8 regions), so the wait is generous:
    // 6e10 SM clocks, about half a minute, after which the kernel gives up, the statistics come back as NaN and the
    // call fails with PGN_ERR_NCCL instead of hanging on a lost peer.  PGN_PEER_SPIN_TICKS overrides it.
    static const long long spin_ticks = [] {
      const char* e = std::getenv("PGN_PEER_SPIN_TICKS");
      const long long v = e ? std::atoll(e) : 0;
      return v > 0 ? v : 60000000000ll;
    }();
    fill_peer_box(&ts.box, spin_ticks);
    ts.seq = ++stats_seq_;
    stats_fused_ = true;
  }
  return ts;
}

int
Engine::run_classify(double epsrel, int relerr)
{
  TileScratch ts = make_scratch();
  if (n_ == 0) {
    if (stats_fused_ || stats_published_)
      ++launches_, k_publish_empty_stats<<<1, kBlock, 0, stream_>>>(ts);
    else
      PGN_TRY(write_empty_stats(stats_dev_.as<double>(), stream_));
    err_cur_ = err_raw_.as<double>();
  } else if (parents_n_ > 0) {
    static const bool no_pairs = std::getenv("PGN_CLASSIFY_NO_PAIRS") != nullptr;
    if (!no_pairs && n_ == 2 * parents_n_ && n_ >= kPairMinRegions) {
      // sibling-pair kernel with TMA-staged loads (stages.cuh); its per-tile counts are atomics on a zeroed array
      if (pairs_attr_device_ != device_) {  // the attribute is per device: once per engine is enough
        PGN_CUDA(cudaFuncSetAttribute(k_classify_pairs, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(kPairSmemBytes)));
        pairs_attr_device_ = device_;
      }
      const unsigned int tiles = static_cast<unsigned int>(tiles_for(n_));
      PGN_CUDA(cudaMemsetAsync(tile_count_.ptr, 0, sizeof(uint32_t) * tiles, stream_));
      const size_t items = (parents_n_ + kTile - 1) / kTile;
      const size_t cap = static_cast<size_t>(sm_count_) * 2;
      ++launches_, k_classify_pairs<<<static_cast<unsigned int>(items < cap ? items : cap), kBlock, kPairSmemBytes,
                                      stream_>>>(est_.as<double>(), err_raw_.as<double>(),
                                                 par_est_[par_cur_].as<double>(), err_ref_.as<double>(),
                                                 active_.as<uint8_t>(), n_, epsrel, relerr, tiles, ts);
    } else {
      ++launches_, k_classify<true><<<super_tiles_for(tiles_for(n_)), kBlock, 0, stream_>>>(
        est_.as<double>(), err_raw_.as<double>(), par_est_[par_cur_].as<double>(), err_ref_.as<double>(),
        active_.as<uint8_t>(), n_, epsrel, relerr, static_cast<unsigned int>(tiles_for(n_)), ts);
    }
    err_cur_ = err_ref_.as<double>();
  } else {
    ++launches_, k_classify<false><<<super_tiles_for(tiles_for(n_)), kBlock, 0, stream_>>>(
      est_.as<double>(), err_raw_.as<double>(), nullptr, nullptr, active_.as<uint8_t>(), n_, epsrel,
      relerr, static_cast<unsigned int>(tiles_for(n_)), ts);
    err_cur_ = err_raw_.as<double>();
  }
  PGN_CUDA(cudaGetLastError());
  classified_ = true;
  return PGN_OK;
}

int
Engine::run_threshold(int mode, double threshold)
{
  TileScratch ts = make_scratch();
  if (n_ == 0) {
    if (stats_fused_ || stats_published_) {
      ++launches_, k_publish_empty_stats<<<1, kBlock, 0, stream_>>>(ts);
      return cudaGetLastError() == cudaSuccess ? PGN_OK : PGN_ERR_CUDA;
    }
    return write_empty_stats(stats_dev_.as<double>(), stream_);
  }
  const int tiles = tiles_for(n_);
  if (mode == 0)
    ++launches_, k_threshold<0><<<tiles, kBlock, 0, stream_>>>(est_.as<double>(), err_cur_, active_.as<uint8_t>(),
                                                  n_, threshold, ts);
  else if (mode == 1)
    ++launches_, k_threshold<1><<<tiles, kBlock, 0, stream_>>>(est_.as<double>(), err_cur_, active_.as<uint8_t>(),
                                                  n_, threshold, ts);
  else
    ++launches_, k_threshold<2><<<tiles, kBlock, 0, stream_>>>(est_.as<double>(), err_cur_, active_.as<uint8_t>(),
                                                  n_, threshold, ts);
  PGN_CUDA(cudaGetLastError());
  return PGN_OK;
}

int
Engine::run_recount()
{
  TileScratch ts = make_scratch();
  if (n_ == 0) {
    if (stats_fused_ || stats_published_) {
      ++launches_, k_publish_empty_stats<<<1, kBlock, 0, stream_>>>(ts);
      return cudaGetLastError() == cudaSuccess ? PGN_OK : PGN_ERR_CUDA;
    }
    return write_empty_stats(stats_dev_.as<double>(), stream_);
  }
  ++launches_, k_recount<<<tiles_for(n_), kBlock, 0, stream_>>>(est_.as<double>(), err_cur_,
                                                   active_.as<uint8_t>(), n_, ts);
  PGN_CUDA(cudaGetLastError());
  return PGN_OK;
}

uint64_t
Engine::digest_device(const void* p, size_t elem, size_t n)
{
  unsigned long long* out = digest_dev_.as<unsigned long long>();
  cudaMemsetAsync(out, 0, sizeof(unsigned long long), stream_);
  if (n > 0) {
    const unsigned g = blocks_for(n, 256);
    if (elem == 8)
      ++launches_, k_digest<double><<<g, 256, 0, stream_>>>(static_cast<const double*>(p), n, out);
    else if (elem == 4)
      ++launches_, k_digest<int32_t><<<g, 256, 0, stream_>>>(static_cast<const int32_t*>(p), n, out);
    else
      ++launches_, k_digest<uint8_t><<<g, 256, 0, stream_>>>(static_cast<const uint8_t*>(p), n, out);
  }
  unsigned long long h = 0;
  cudaMemcpyAsync(&h, out, sizeof h, cudaMemcpyDeviceToHost, stream_);
  cudaStreamSynchronize(stream_);
  return h;
}

void
Engine::digest_arrays(pgn_iter_stats* st, bool after_classify)
{
  if (!after_classify) {
    st->dig_left = digest_device(left_[cur_].ptr, 8, n_ * ndim_);
    st->dig_length = digest_device(length_[cur_].ptr, 8, n_ * ndim_);
    st->dig_estimate = digest_device(est_.ptr, 8, n_);
    st->dig_errorest_raw = digest_device(err_raw_.ptr, 8, n_);
    st->dig_axis = digest_device(axis_.ptr, 4, n_);
  } else {
    st->dig_errorest = digest_device(err_cur_, 8, n_);
    st->dig_active = digest_device(active_.ptr, 1, n_);
  }
}

// ------------------------------------------------------------------------------------------------
// stage-level API
// ------------------------------------------------------------------------------------------------
int
Engine::evaluate(pgn_eval_fn fn, void* user, double* sum_est, double* sum_err)
{
  if (!fn)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  if (filtered_pending_) {
    set_last_error("evaluate() between filter() and split()");
    return PGN_ERR_STATE;
  }
  PGN_TRY(launch_evaluate(fn, user));
  evaluated_ = true;
  // apply_cubature_integration_rules also resets every flag to active and returns the two sums
  // (PaganiUtils.cuh:310-311, 336-341); parents are left untouched.
  const size_t keep_parents = parents_n_;
  parents_n_ = 0;
  int rc = run_classify(0.0, 0);
  parents_n_ = keep_parents;
  PGN_TRY(rc);
  classified_ = false;
  PGN_TRY(sync_stats(false));
  if (sum_est)
    *sum_est = host_global_[kSumEst];
  if (sum_err)
    *sum_err = host_global_[kSumErr];
  return PGN_OK;
}

int
Engine::classify(double epsrel, int relerr, pgn_iter_stats* out)
{
  PGN_CUDA(cudaSetDevice(device_));
  if (!evaluated_) {
    set_last_error("classify() before evaluate()");
    return PGN_ERR_STATE;
  }
  PGN_TRY(run_classify(epsrel, relerr));
  PGN_TRY(sync_stats(false));
  if (out) {
    std::memset(out, 0, sizeof *out);
    out->nregions = n_;
    out->iter_estimate = host_global_[kSumEst];
    out->iter_errorest = host_global_[kSumErr];
    out->finished_estimate = host_global_[kSumEst] - host_global_[kSumActEst];
    out->finished_errorest = host_global_[kSumErr] - host_global_[kSumActErr];
    out->nactive = static_cast<uint64_t>(host_global_[kNumActive]);
  }
  return PGN_OK;
}

// Stage-level Heuristic_classifier::classify on the current error list (pgn_heuristic_classify).
int
Engine::heuristic_classify_stage(double epsrel, double last_estimate, double iter_errorest,
                                 double iter_finished_errorest, double total_finished_errorest,
                                 pgn_heuristic_result* out)
{
  if (!out)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  if (!classified_ || n_ == 0) {
    set_last_error("heuristic_classify() needs a classified, non-empty region list");
    return PGN_ERR_STATE;
  }
  tracker_.reset(epsrel, 0.0);
  for (int k = 0; k < 3; ++k)
    tracker_.store(last_estimate);
  PGN_TRY(agree_mem_capacity());
  HeuristicOutcome hs;
  PGN_TRY(heuristic_classify(n_, 0.0, iter_errorest, iter_finished_errorest, total_finished_errorest, &hs, true));
  if (!hs.success) {
    // The flags are still those of classify(), but the host-driven search leaves the count of its last PROBE in the
    // statistics slot filter() sizes its output from: recount from the flags.
    PGN_TRY(run_recount());
    PGN_TRY(sync_stats(false));
  }
  std::memset(out, 0, sizeof *out);
  out->success = hs.success ? 1 : 0;
  out->must_terminate = hs.must_terminate ? 1 : 0;
  out->probes = hs.probes;
  out->num_active = hs.num_active;
  out->threshold = hs.threshold;
  out->finished_errorest = hs.finished_errorest;
  return PGN_OK;
}

int
Engine::set_active(const uint8_t* flags, size_t n)
{
  if (n != n_ || (!flags && n > 0))
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  PGN_TRY(ensure_iteration_buffers(n_));
  if (!err_cur_)
    err_cur_ = err_raw_.as<double>();
  if (n > 0)
    PGN_CUDA(cudaMemcpyAsync(active_.ptr, flags, n, cudaMemcpyHostToDevice, stream_));
  PGN_TRY(run_recount());
  PGN_TRY(sync_stats(false));
  classified_ = true;
  return PGN_OK;
}

int
Engine::set_estimates(const double* est, const double* err, const int32_t* axis, size_t n)
{
  if (n != n_)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  PGN_TRY(ensure_iteration_buffers(n_));
  if (n > 0) {
    if (est)
      PGN_CUDA(cudaMemcpyAsync(est_.ptr, est, sizeof(double) * n, cudaMemcpyHostToDevice, stream_));
    if (err)
      PGN_CUDA(
        cudaMemcpyAsync(err_raw_.ptr, err, sizeof(double) * n, cudaMemcpyHostToDevice, stream_));
    if (axis)
      PGN_CUDA(
        cudaMemcpyAsync(axis_.ptr, axis, sizeof(int32_t) * n, cudaMemcpyHostToDevice, stream_));
  }
  PGN_CUDA(cudaStreamSynchronize(stream_));
  err_cur_ = err_raw_.as<double>();
  evaluated_ = true;
  return PGN_OK;
}

int
Engine::fetch_estimates(double* est, double* err, int32_t* axis, uint8_t* active, size_t cap)
{
  if (cap < n_)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  if (n_ > 0) {
    if (est)
      PGN_CUDA(cudaMemcpyAsync(est, est_.ptr, sizeof(double) * n_, cudaMemcpyDeviceToHost, stream_));
    if (err)
      PGN_CUDA(cudaMemcpyAsync(err, err_cur_ ? err_cur_ : err_raw_.as<double>(), sizeof(double) * n_,
                               cudaMemcpyDeviceToHost, stream_));
    if (axis)
      PGN_CUDA(
        cudaMemcpyAsync(axis, axis_.ptr, sizeof(int32_t) * n_, cudaMemcpyDeviceToHost, stream_));
    if (active)
      PGN_CUDA(cudaMemcpyAsync(active, active_.ptr, n_, cudaMemcpyDeviceToHost, stream_));
  }
  PGN_CUDA(cudaStreamSynchronize(stream_));
  return PGN_OK;
}

int
Engine::fetch_parents(double* est, double* err, int32_t* axis, size_t cap, size_t* n)
{
  if (n)
    *n = parents_n_;
  if (cap < parents_n_)
    return (est || err || axis) ? PGN_ERR_INVALID : PGN_OK;
  PGN_CUDA(cudaSetDevice(device_));
  if (parents_n_ > 0) {
    if (est)
      PGN_CUDA(cudaMemcpyAsync(est, par_est_[par_cur_].ptr, sizeof(double) * parents_n_,
                               cudaMemcpyDeviceToHost, stream_));
    if (err)
      PGN_CUDA(cudaMemcpyAsync(err, par_err_[par_cur_].ptr, sizeof(double) * parents_n_,
                               cudaMemcpyDeviceToHost, stream_));
    if (axis && filtered_pending_)
      PGN_CUDA(cudaMemcpyAsync(axis, par_axis_.ptr, sizeof(int32_t) * parents_n_,
                               cudaMemcpyDeviceToHost, stream_));
  }
  PGN_CUDA(cudaStreamSynchronize(stream_));
  return PGN_OK;
}

// Sub_regions_filter::filter (Sub_region_filter.cuh:85-139): compaction only.
int
Engine::filter(size_t* num_active)
{
  PGN_CUDA(cudaSetDevice(device_));
  if (!classified_) {
    set_last_error("filter() needs classify() or set_active() first");
    return PGN_ERR_STATE;
  }
  const size_t m = static_cast<size_t>(host_local_[kNumActive]);
  if (num_active)
    *num_active = m;
  if (m == 0) {
    // the reference returns early and leaves the list untouched; the caller sets size = 0
    n_ = 0;
    parents_n_ = 0;
    evaluated_ = classified_ = false;
    return PGN_OK;
  }
  const int nxt = 1 - cur_, pn = 1 - par_cur_;
  PGN_TRY(left_[nxt].reserve(sizeof(double) * m * ndim_));
  PGN_TRY(length_[nxt].reserve(sizeof(double) * m * ndim_));
  PGN_TRY(par_est_[pn].reserve(sizeof(double) * m));
  PGN_TRY(par_err_[pn].reserve(sizeof(double) * m));
  PGN_TRY(par_axis_.reserve(sizeof(int32_t) * m));
  CompactArgs A;
  A.left = left_[cur_].as<double>();
  A.length = length_[cur_].as<double>();
  A.est = est_.as<double>();
  A.err = err_cur_;
  A.axis = axis_.as<int32_t>();
  A.active = active_.as<uint8_t>();
  A.tile_offset = tile_offset_.as<uint64_t>();
  A.n = n_;
  A.m = m;
  A.out_left = left_[nxt].as<double>();
  A.out_length = length_[nxt].as<double>();
  A.parent_est = par_est_[pn].as<double>();
  A.parent_err = par_err_[pn].as<double>();
  A.out_axis = par_axis_.as<int32_t>();
  A.ndim = ndim_;
  ++launches_, k_compact<false><<<tiles_for(n_), kBlock, 0, stream_>>>(A);
  PGN_CUDA(cudaGetLastError());
  cur_ = nxt;
  par_cur_ = pn;
  n_ = m;
  parents_n_ = m;
  filtered_pending_ = true;
  evaluated_ = classified_ = false;
  return PGN_OK;
}

// Sub_region_splitter::split (Sub_region_splitter.cuh:55-87)
int
Engine::split()
{
  PGN_CUDA(cudaSetDevice(device_));
  if (n_ == 0)
    return PGN_OK;
  // after filter() the compacted axes are used; without it (RegionSplitting.cu sets the axes by
  // hand and splits every region) the per-region axes of the current list
  if (!filtered_pending_ && !evaluated_) {
    set_last_error("split() needs split axes: evaluate(), set_estimates() or filter() first");
    return PGN_ERR_STATE;
  }
  const int32_t* axes = filtered_pending_ ? par_axis_.as<int32_t>() : axis_.as<int32_t>();
  const size_t m = n_;
  const int nxt = 1 - cur_;
  PGN_TRY(left_[nxt].reserve(sizeof(double) * 2 * m * ndim_));
  PGN_TRY(length_[nxt].reserve(sizeof(double) * 2 * m * ndim_));
  ++launches_, k_split<<<blocks_for(m, kBlock), kBlock, 0, stream_>>>(
    left_[cur_].as<double>(), length_[cur_].as<double>(), axes, m, ndim_,
    left_[nxt].as<double>(), length_[nxt].as<double>());
  if (!filtered_pending_)
    parents_n_ = 0;
  evaluated_ = classified_ = false;
  PGN_CUDA(cudaGetLastError());
  cur_ = nxt;
  n_ = 2 * m;
  filtered_pending_ = false;
  return PGN_OK;
}


// ------------------------------------------------------------------------------------------------
// stage API on caller-owned device arrays (reference layouts: bounds [ndim][n], flags as doubles)
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void
k_fill_f64(double* x, size_t n, double v)
{
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    x[i] = v;
}
// Region_characteristics::active_regions is an array of doubles in the reference (0. / 1.)
__global__ void
k_flags_from_f64(const double* __restrict__ f, uint8_t* __restrict__ out, size_t n)
{
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    out[i] = f[i] != 0.0 ? 1 : 0;
}
}  // namespace

int
Engine::dev_uniform_split(int parts, double* d_left, double* d_length, size_t n)
{
  if (parts < 1 || (!d_left || !d_length))
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  double total = 1.0;
  for (int d = 0; d < ndim_; ++d)
    total *= parts;
  if (static_cast<double>(n) != total)
    return PGN_ERR_INVALID;
  ++launches_, k_uniform_split<<<blocks_for(n, 512), 512, 0, stream_>>>(d_left, d_length, n, parts, ndim_,
                                                                         1. / static_cast<double>(parts));
  PGN_CUDA(cudaGetLastError());
  PGN_CUDA(cudaStreamSynchronize(stream_));
  return PGN_OK;
}

int
Engine::dev_evaluate(pgn_eval_fn fn, void* user, const double* d_left, const double* d_length, size_t n,
                     double* d_est, double* d_err, int32_t* d_axis, double* d_active_f64, double* sum_est,
                     double* sum_err)
{
  if (!fn || ((!d_left || !d_length || !d_est || !d_err || !d_axis) && n > 0))
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  PGN_TRY(ensure_iteration_buffers(n));
  if (n > 0) {
    PGN_TRY(launch_evaluate_on(fn, user, d_left, d_length, n, d_est, d_err, d_axis));
    // apply_cubature_integration_rules resets every flag to active (PaganiUtils.cuh:310-311)
    if (d_active_f64)
      ++launches_, k_fill_f64<<<blocks_for(n, 256), 256, 0, stream_>>>(d_active_f64, n, 1.0);
    // the two sums it returns (:336-341): one pass of the classify kernel without refinement
    TileScratch ts = make_scratch(false);
    stats_published_ = false;  // read back below with an explicit copy
    ++launches_, k_classify<false><<<super_tiles_for(tiles_for(n)), kBlock, 0, stream_>>>(
      d_est, d_err, nullptr, nullptr, active_.as<uint8_t>(), n, 0.0, 0, static_cast<unsigned int>(tiles_for(n)), ts);
    PGN_CUDA(cudaGetLastError());
    PGN_CUDA(cudaMemcpyAsync(host_local_, stats_dev_.ptr, sizeof(double) * kStatsLen, cudaMemcpyDeviceToHost,
                             stream_));
  }
  PGN_CUDA(cudaStreamSynchronize(stream_));
  if (sum_est)
    *sum_est = n > 0 ? host_local_[kSumEst] : 0.0;
  if (sum_err)
    *sum_err = n > 0 ? host_local_[kSumErr] : 0.0;
  return PGN_OK;
}

// Sub_regions_filter::get_num_active_regions (Sub_region_filter.cuh:61-82); leaves the tile prefix for dev_filter
int
Engine::dev_count_active(const double* d_active_f64, size_t n, size_t* count)
{
  if (!count || (!d_active_f64 && n > 0))
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  *count = 0;
  if (n == 0)
    return PGN_OK;
  PGN_TRY(ensure_iteration_buffers(n));
  ++launches_, k_flags_from_f64<<<blocks_for(n, 256), 256, 0, stream_>>>(d_active_f64, active_.as<uint8_t>(), n);
  TileScratch ts = make_scratch(false);
  stats_published_ = false;  // read back below with an explicit copy
  // the flag array doubles as the (unused) value arrays of the counting kernel
  ++launches_, k_recount<<<tiles_for(n), kBlock, 0, stream_>>>(d_active_f64, d_active_f64, active_.as<uint8_t>(), n,
                                                                ts);
  PGN_CUDA(cudaGetLastError());
  PGN_CUDA(cudaMemcpyAsync(host_local_, stats_dev_.ptr, sizeof(double) * kStatsLen, cudaMemcpyDeviceToHost,
                           stream_));
  PGN_CUDA(cudaStreamSynchronize(stream_));
  *count = static_cast<size_t>(host_local_[kNumActive]);
  return PGN_OK;
}

// Sub_regions_filter::filter (Sub_region_filter.cuh:85-139) into caller-allocated arrays of m survivors
int
Engine::dev_filter(const double* d_left, const double* d_length, const double* d_est, const double* d_err,
                   const int32_t* d_axis, const double* d_active_f64, size_t n, size_t m, double* out_left,
                   double* out_length, double* out_est, double* out_err, int32_t* out_axis)
{
  if (n == 0 || m == 0)
    return PGN_OK;
  if (!d_left || !d_length || !d_est || !d_err || !d_axis || !d_active_f64 || !out_left || !out_length ||
      !out_est || !out_err || !out_axis)
    return PGN_ERR_INVALID;
  size_t counted = 0;
  PGN_TRY(dev_count_active(d_active_f64, n, &counted));
  if (counted != m) {
    set_last_error("dev_filter: the output arrays are sized for a different number of survivors");
    return PGN_ERR_INVALID;
  }
  CompactArgs A;
  A.left = d_left;
  A.length = d_length;
  A.est = d_est;
  A.err = d_err;
  A.axis = d_axis;
  A.active = active_.as<uint8_t>();
  A.tile_offset = tile_offset_.as<uint64_t>();
  A.n = n;
  A.m = m;
  A.out_left = out_left;
  A.out_length = out_length;
  A.parent_est = out_est;
  A.parent_err = out_err;
  A.out_axis = out_axis;
  A.ndim = ndim_;
  ++launches_, k_compact<false><<<tiles_for(n), kBlock, 0, stream_>>>(A);
  PGN_CUDA(cudaGetLastError());
  PGN_CUDA(cudaStreamSynchronize(stream_));
  return PGN_OK;
}

// Sub_region_splitter::split (Sub_region_splitter.cuh:55-87): m parents -> 2m children at j and m + j
int
Engine::dev_split(const double* d_left, const double* d_length, const int32_t* d_axis, size_t m, double* out_left,
                  double* out_length)
{
  if (m == 0)
    return PGN_OK;
  if (!d_left || !d_length || !d_axis || !out_left || !out_length)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  ++launches_, k_split<<<blocks_for(m, kBlock), kBlock, 0, stream_>>>(d_left, d_length, d_axis, m, ndim_, out_left,
                                                                       out_length);
  PGN_CUDA(cudaGetLastError());
  PGN_CUDA(cudaStreamSynchronize(stream_));
  return PGN_OK;
}

int
Engine::regions_import_device(const double* d_left, const double* d_length, size_t n)
{
  if ((!d_left || !d_length) && n > 0)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  PGN_TRY(left_[cur_].reserve(sizeof(double) * std::max<size_t>(n, 1) * ndim_));
  PGN_TRY(length_[cur_].reserve(sizeof(double) * std::max<size_t>(n, 1) * ndim_));
  if (n > 0) {
    PGN_CUDA(cudaMemcpyAsync(left_[cur_].ptr, d_left, sizeof(double) * n * ndim_, cudaMemcpyDeviceToDevice, stream_));
    PGN_CUDA(cudaMemcpyAsync(length_[cur_].ptr, d_length, sizeof(double) * n * ndim_, cudaMemcpyDeviceToDevice,
                             stream_));
    PGN_CUDA(cudaStreamSynchronize(stream_));
  }
  n_ = n;
  parents_n_ = 0;
  evaluated_ = classified_ = filtered_pending_ = false;
  return PGN_OK;
}

int
Engine::regions_export_device(double* d_left, double* d_length, size_t cap, size_t* n)
{
  if (n)
    *n = n_;
  if (cap < n_)
    return d_left || d_length ? PGN_ERR_INVALID : PGN_OK;
  PGN_CUDA(cudaSetDevice(device_));
  if (n_ > 0 && d_left)
    PGN_CUDA(cudaMemcpyAsync(d_left, left_[cur_].ptr, sizeof(double) * n_ * ndim_, cudaMemcpyDeviceToDevice, stream_));
  if (n_ > 0 && d_length)
    PGN_CUDA(cudaMemcpyAsync(d_length, length_[cur_].ptr, sizeof(double) * n_ * ndim_, cudaMemcpyDeviceToDevice,
                             stream_));
  PGN_CUDA(cudaStreamSynchronize(stream_));
  return PGN_OK;
}

// ------------------------------------------------------------------------------------------------
// multi-GPU: move contiguous runs of compacted parents so every rank holds ~the same number
// ------------------------------------------------------------------------------------------------
int
Engine::rebalance_parents(size_t* m_local)
{
  std::vector<size_t> new_counts;
  std::vector<Transfer> plan;
  plan_rebalance(rank_counts_, &new_counts, &plan);
  const int me = comm_->rank;
  const size_t m_old = *m_local, m_new = new_counts[me];
  const size_t cols = 2 * static_cast<size_t>(ndim_) + 3;
  // staging offsets (in doubles) of every transfer that touches this rank
  std::vector<size_t> soff(plan.size(), 0), roff(plan.size(), 0);
  size_t stot = 0, rtot = 0;
  for (size_t k = 0; k < plan.size(); ++k) {
    if (plan[k].src == me) {
      soff[k] = stot;
      stot += plan[k].count * cols;
    }
    if (plan[k].dst == me) {
      roff[k] = rtot;
      rtot += plan[k].count * cols;
    }
  }
  PGN_TRY(pack_send_.reserve(sizeof(double) * std::max<size_t>(stot, 1)));
  PGN_TRY(pack_recv_.reserve(sizeof(double) * std::max<size_t>(rtot, 1)));
  const int pc = par_cur_;
  for (size_t k = 0; k < plan.size(); ++k) {
    const Transfer& t = plan[k];
    if (t.src != me)
      continue;
    // self transfers are packed straight into the receive staging area
    double* dst = t.dst == me ? pack_recv_.as<double>() + roff[k] : pack_send_.as<double>() + soff[k];
    ++launches_, k_pack<<<blocks_for(t.count * cols, 256), 256, 0, stream_>>>(
      left_[cur_].as<double>(), length_[cur_].as<double>(), par_est_[pc].as<double>(),
      par_err_[pc].as<double>(), par_axis_.as<int32_t>(), m_old, ndim_, t.src_offset, t.count, dst);
  }
  PGN_CUDA(cudaGetLastError());
  PGN_TRY(comm_->exchange(plan, pack_send_.as<double>(), soff, pack_recv_.as<double>(), roff, cols,
                          stream_));
  // unpack into the other buffers (the pre-compaction list is dead by now)
  const int nxt = 1 - cur_, pn = 1 - par_cur_;
  PGN_TRY(left_[nxt].reserve(sizeof(double) * std::max<size_t>(m_new, 1) * ndim_));
  PGN_TRY(length_[nxt].reserve(sizeof(double) * std::max<size_t>(m_new, 1) * ndim_));
  PGN_TRY(par_est_[pn].reserve(sizeof(double) * std::max<size_t>(m_new, 1)));
  PGN_TRY(par_err_[pn].reserve(sizeof(double) * std::max<size_t>(m_new, 1)));
  // the split axes are unpacked in place: every byte of par_axis_ was packed already
  PGN_CUDA(cudaStreamSynchronize(stream_));
  DeviceBuf& new_axis = par_axis_next_;  // grow-only twin of par_axis_: no allocation per rebalance, nothing to leak
  PGN_TRY(new_axis.reserve(sizeof(int32_t) * std::max<size_t>(m_new, 1)));
  for (size_t k = 0; k < plan.size(); ++k) {
    const Transfer& t = plan[k];
    if (t.dst != me)
      continue;
    ++launches_, k_unpack<<<blocks_for(t.count * cols, 256), 256, 0, stream_>>>(
      pack_recv_.as<double>() + roff[k], t.count, t.dst_offset, m_new, ndim_,
      left_[nxt].as<double>(), length_[nxt].as<double>(), par_est_[pn].as<double>(),
      par_err_[pn].as<double>(), new_axis.as<int32_t>());
  }
  PGN_CUDA(cudaGetLastError());
  PGN_CUDA(cudaStreamSynchronize(stream_));
  std::swap(par_axis_, par_axis_next_);
  cur_ = nxt;
  par_cur_ = pn;
  *m_local = m_new;
  rank_counts_ = new_counts;
  return PGN_OK;
}

int
Engine::compact_and_split(size_t m_local, size_t m_global, bool* rebalanced)
{
  *rebalanced = false;
  bool want_rebalance = false;
  if (comm_ && comm_->nranks > 1 && m_global >= rebalance_min_) {
    size_t mx = 0;
    for (size_t c : rank_counts_)
      mx = std::max(mx, c);
    const double mean = static_cast<double>(m_global) / comm_->nranks;
    want_rebalance = static_cast<double>(mx) > rebalance_ratio_ * mean;
  }
  CompactArgs A;
  A.left = left_[cur_].as<double>();
  A.length = length_[cur_].as<double>();
  A.est = est_.as<double>();
  A.err = err_cur_;
  A.axis = axis_.as<int32_t>();
  A.active = active_.as<uint8_t>();
  A.tile_offset = tile_offset_.as<uint64_t>();
  A.n = n_;
  A.m = m_local;
  A.ndim = ndim_;
  const int nxt = 1 - cur_, pn = 1 - par_cur_;
  const size_t m1 = std::max<size_t>(m_local, 1);
  PGN_TRY(par_est_[pn].reserve(sizeof(double) * m1));
  PGN_TRY(par_err_[pn].reserve(sizeof(double) * m1));
  A.parent_est = par_est_[pn].as<double>();
  A.parent_err = par_err_[pn].as<double>();
  if (!want_rebalance) {
    PGN_TRY(left_[nxt].reserve(sizeof(double) * 2 * m1 * ndim_));
    PGN_TRY(length_[nxt].reserve(sizeof(double) * 2 * m1 * ndim_));
    A.out_left = left_[nxt].as<double>();
    A.out_length = length_[nxt].as<double>();
    A.out_axis = nullptr;
    if (n_ > 0 && m_local > 0)
      ++launches_, k_compact<true><<<tiles_for(n_), kBlock, 0, stream_>>>(A);
    PGN_CUDA(cudaGetLastError());
    cur_ = nxt;
    par_cur_ = pn;
    n_ = 2 * m_local;
    parents_n_ = m_local;
    return PGN_OK;
  }
  // compaction only, exchange, then bisect
  PGN_TRY(left_[nxt].reserve(sizeof(double) * m1 * ndim_));
  PGN_TRY(length_[nxt].reserve(sizeof(double) * m1 * ndim_));
  PGN_TRY(par_axis_.reserve(sizeof(int32_t) * m1));
  A.out_left = left_[nxt].as<double>();
  A.out_length = length_[nxt].as<double>();
  A.out_axis = par_axis_.as<int32_t>();
  if (n_ > 0 && m_local > 0)
    ++launches_, k_compact<false><<<tiles_for(n_), kBlock, 0, stream_>>>(A);
  PGN_CUDA(cudaGetLastError());
  cur_ = nxt;
  par_cur_ = pn;
  size_t m = m_local;
  PGN_TRY(rebalance_parents(&m));
  *rebalanced = true;
  const int nx2 = 1 - cur_;
  const size_t m2 = std::max<size_t>(m, 1);
  PGN_TRY(left_[nx2].reserve(sizeof(double) * 2 * m2 * ndim_));
  PGN_TRY(length_[nx2].reserve(sizeof(double) * 2 * m2 * ndim_));
  if (m > 0)
    ++launches_, k_split<<<blocks_for(m, kBlock), kBlock, 0, stream_>>>(
      left_[cur_].as<double>(), length_[cur_].as<double>(), par_axis_.as<int32_t>(), m, ndim_,
      left_[nx2].as<double>(), length_[nx2].as<double>());
  PGN_CUDA(cudaGetLastError());
  cur_ = nx2;
  n_ = 2 * m;
  parents_n_ = m;
  return PGN_OK;
}

// ------------------------------------------------------------------------------------------------
// bucket partition for the device-side threshold search (threshold_search.cuh)
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void
k_outside_to_double(const unsigned long long* outside, double* out)
{
  *out = static_cast<double>(*outside);
}
}  // namespace

namespace {
void
lay_out_tables(unsigned long long* base, BucketTables* T)
{
  const size_t nbm = kMaxBuckets;
  T->count = base;
  T->start = base + nbm;               // nbm + 1
  T->cursor = base + 2 * nbm + 1;
  T->bucket_sum = reinterpret_cast<double*>(base + 3 * nbm + 1);
  T->suffix_sum = reinterpret_cast<double*>(base + 4 * nbm + 1);    // nbm + 1
  T->suffix_count = base + 5 * nbm + 2;                              // nbm + 1
  T->outside = base + 6 * nbm + 3;
  T->acc_lo = base + 6 * nbm + 8;
  T->acc_hi = base + 7 * nbm + 8;
  T->sorted = nullptr;
}
}  // namespace

// T: this rank's partition.  G (several ranks with peer mailboxes, else null): the tables of the concatenated list,
// from ONE all-reduce of the counts and integer bucket sums (threshold_search.cuh, k_tables_to_double).
int
Engine::prepare_search(const BucketPlan& plan, BucketTables* T, BucketTables* G)
{
  const size_t nbm = kMaxBuckets;
  lay_out_tables(search_tab_.as<unsigned long long>(), T);
  PGN_TRY(sorted_err_.reserve(sizeof(double) * std::max<size_t>(n_, 1)));
  T->sorted = sorted_err_.as<double>();
  PGN_CUDA(cudaMemsetAsync(search_tab_.ptr, 0, 8 * (8 * nbm + 8), stream_));
  const unsigned grid = static_cast<unsigned>(std::min<size_t>((n_ + kScatterChunk - 1) / kScatterChunk,
                                                               static_cast<size_t>(sm_count_) * 2));
  if (n_ > 0)
    ++launches_, k_bucket_count<<<grid, kSearchBlock, 0, stream_>>>(err_cur_, n_, plan, *T);
  ++launches_, k_bucket_offsets<<<1, kSearchBlock, 0, stream_>>>(plan, *T);
  if (n_ > 0) {
    ++launches_, k_bucket_scatter<<<grid, kSearchBlock, 0, stream_>>>(err_cur_, n_, plan, *T);
    const unsigned sum_grid = static_cast<unsigned>(std::min<size_t>((n_ + kSumChunk - 1) / kSumChunk,
                                                                     static_cast<size_t>(sm_count_) * 8));
    ++launches_, k_bucket_sums<<<sum_grid, kSumThreads, 0, stream_>>>(plan, *T);
  }
  if (G) {
    PGN_TRY(search_gtab_.reserve(8 * (8 * nbm + 8)));
    PGN_TRY(search_gd_.reserve(sizeof(double) * (3 * nbm + 8)));
    lay_out_tables(search_gtab_.as<unsigned long long>(), G);
    ++launches_, k_tables_to_double<<<1, kSearchBlock, 0, stream_>>>(plan, *T, search_gd_.as<double>());
    PGN_CUDA(cudaGetLastError());
    PGN_TRY(comm_->allreduce(search_gd_.as<double>(), static_cast<int>(3 * plan.nb + 1), kOpSum, stream_));
    ++launches_, k_global_tables<<<1, kSearchBlock, 0, stream_>>>(plan, search_gd_.as<double>(), *G);
  } else {
    ++launches_, k_bucket_suffix<<<1, kSearchBlock, 0, stream_>>>(plan, *T);
  }
  PGN_CUDA(cudaGetLastError());
  return PGN_OK;
}

// ------------------------------------------------------------------------------------------------
// heuristic (memory-pressure) classifier
// ------------------------------------------------------------------------------------------------
double
Engine::mem_capacity_bytes() const
{
  if (opts_.mem_mode == PGN_MEM_NATIVE) {
    size_t f = 0, t = 0;
    cudaMemGetInfo(&f, &t);
    const double frac = opts_.native_mem_fraction > 0 ? opts_.native_mem_fraction : 0.85;
    // what is free now plus what this engine already holds
    const DeviceBuf* mine[] = {&left_[0], &left_[1], &length_[0], &length_[1], &est_, &err_raw_,
                               &err_ref_, &axis_, &active_, &par_est_[0], &par_est_[1],
                               &par_err_[0], &par_err_[1]};
    size_t held = 0;
    for (const DeviceBuf* b : mine)
      held += b->bytes;
    return frac * static_cast<double>(f + held);
  }
  return 16e9;  // heuristic_classifier.cuh:89-94
}

// The capacity every rank bases its heuristic decisions on.  PGN_MEM_PARITY: the reference's constant.
// PGN_MEM_NATIVE: ranks x the SMALLEST per-rank figure, agreed once per call through a MIN all-reduce -- ranks that
// derived `criteria` / `must_terminate` from their own free memory could disagree on entering the (collective)
// heuristic path and dead-lock in it.
int
Engine::agree_mem_capacity()
{
  double mine = mem_capacity_bytes();
  const int w = comm_ ? comm_->nranks : 1;
  if (opts_.mem_mode == PGN_MEM_NATIVE && w > 1) {
    double neg = -mine;
    PGN_CUDA(cudaMemcpyAsync(stats_glob_.ptr, &neg, sizeof neg, cudaMemcpyHostToDevice, stream_));
    PGN_TRY(comm_->allreduce(stats_glob_.as<double>(), 1, kOpMax, stream_));
    PGN_CUDA(cudaMemcpyAsync(&neg, stats_glob_.ptr, sizeof neg, cudaMemcpyDeviceToHost, stream_));
    PGN_CUDA(cudaStreamSynchronize(stream_));
    mine = -neg * w;
  }
  mem_capacity_ = mine;
  return PGN_OK;
}

int
Engine::heuristic_classify(size_t n_global, double iter_est, double iter_err, double fin_err,
                           double cum_err, HeuristicOutcome* out, bool force)
{
  (void)iter_est;
  // heuristic_classifier.cuh:105-145, 227-264: bytes a full split of n regions would need
  const size_t nd = static_cast<size_t>(ndim_);
  const size_t doubles = n_global * (10 * nd + 8);
  const size_t ints = 4 * n_global;
  const size_t required = 8 * doubles + 4 * ints;
  const size_t total = static_cast<size_t>(mem_capacity_);
  const size_t free_mem = total > required ? total - required : 1;
  const double ratio = static_cast<double>(required) / static_cast<double>(free_mem);
  const bool necessary = ratio > 1.;
  const bool criteria = ratio > 1. || (ratio > .1 && tracker_.converged());
  if (!criteria && !force) {
    out->must_terminate = necessary;
    return PGN_OK;
  }
  out->ran = true;

  // Heuristic_classifier::classify (heuristic_classifier.cuh:362-451)
  if (comm_ && comm_->nranks > 1 && !stats_fused_) {
    // min/max of the errors over all ranks (the local ones are still in stats_dev_)
    PGN_CUDA(cudaMemcpyAsync(stats_glob_.as<double>() + kNegMinSlot,
                             stats_dev_.as<double>() + kNegMinSlot, sizeof(double) * 2,
                             cudaMemcpyDeviceToDevice, stream_));
    PGN_TRY(comm_->allreduce(stats_glob_.as<double>() + kNegMinSlot, 2, kOpMax, stream_));
    PGN_CUDA(cudaMemcpyAsync(host_global_ + kNegMinSlot, stats_glob_.as<double>() + kNegMinSlot,
                             sizeof(double) * 2, cudaMemcpyDeviceToHost, stream_));
    PGN_CUDA(cudaStreamSynchronize(stream_));
  }
  const double min_err = -host_global_[kNegMinSlot];
  const double max_err = host_global_[kMaxSlot];
  const bool multi = comm_ && comm_->nranks > 1;
  static const bool search_debug = std::getenv("PGN_SEARCH_DEBUG") != nullptr;

  double range_lo = min_err, range_hi = max_err;
  double threshold = iter_err / static_cast<double>(n_global);
  const double target_error = std::fabs(tracker_.last[2]) * tracker_.epsrel;
  bool pass_mem = false, pass_budget = false;
  size_t num_active = 0;
  double finished_errorest = 0.;
  double active_err_sum = 0.;
  size_t increases = 0, decreases = 0;
  // threshold of the last probe: the flags the reference adopts are those of its last apply_threshold
  // call, even when the exhausted-attempts branch resets `threshold` after that probe succeeded
  // (heuristic_classifier.cuh:404-437, Workspace.cuh:104-106)
  double applied_threshold = threshold;

  // Device-side search: partition the errors into value buckets once, then either run the whole
  // bisection in one kernel (single rank) or answer every probe from the tables (several ranks).
  BucketPlan plan;
  BucketTables tables, global_tables;
  bool fast = device_search_ && make_bucket_plan(min_err, max_err, &plan);
  if (fast)
    PGN_TRY(prepare_search(plan, &tables, multi && peer_ready_ ? &global_tables : nullptr));
  bool searched = false;
  if (fast && (!multi || peer_ready_)) {
    SearchIO* h = static_cast<SearchIO*>(host_search_);
    std::memset(h, 0, sizeof *h);
    h->iter_errorest = iter_err;
    h->finished_errorest = fin_err;
    h->cum_errorest = cum_err;
    h->target_error = target_error;
    h->min_err = min_err;
    h->max_err = max_err;
    h->n_regions = static_cast<double>(n_global);
    h->max_percent_error_budget = max_percent_error_budget_;
    h->max_active_regions_percentage = max_active_regions_percentage_;
    h->seq_base = peer_seq_;
    unsigned long long* h_outside =
      reinterpret_cast<unsigned long long*>(static_cast<char*>(host_search_) + sizeof(SearchIO));
    PeerBox box;
    std::memset(&box, 0, sizeof box);
    if (multi)
      fill_peer_box(&box, 4000000000ll);  // ~2 s of SM clocks: the ranks run this kernel in lock step
    PGN_CUDA(cudaMemcpyAsync(search_io_.ptr, h, sizeof *h, cudaMemcpyHostToDevice, stream_));
    if (search_attr_device_ != device_) {  // per device: once per engine is enough
      PGN_CUDA(cudaFuncSetAttribute(k_search<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kSearchSmemBytes)));
      PGN_CUDA(cudaFuncSetAttribute(k_search<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    static_cast<int>(kSearchSmemBytes)));
      search_attr_device_ = device_;
    }
    if (multi)
      ++launches_, k_search<true><<<1, kSearchBlock, kSearchSmemBytes, stream_>>>(plan, tables, global_tables,
                                                                                     search_io_.as<SearchIO>(), box);
    else
      ++launches_, k_search<false><<<1, kSearchBlock, kSearchSmemBytes, stream_>>>(plan, tables, tables,
                                                                                      search_io_.as<SearchIO>(), box);
    PGN_CUDA(cudaGetLastError());
    PGN_CUDA(cudaMemcpyAsync(h, search_io_.ptr, sizeof *h, cudaMemcpyDeviceToHost, stream_));
    PGN_CUDA(cudaMemcpyAsync(h_outside, multi ? global_tables.outside : tables.outside, sizeof *h_outside,
                             cudaMemcpyDeviceToHost, stream_));
    PGN_CUDA(cudaStreamSynchronize(stream_));
    if (multi)
      peer_seq_ = h->seq_base;
    if (h->probes == -1) {
      set_last_error("multi-GPU threshold search: a peer did not answer within the spin limit");
      return PGN_ERR_NCCL;
    }
    if (h->probes >= 0 && (multi || *h_outside == 0)) {
      searched = true;
      threshold = h->threshold;
      applied_threshold = h->applied_threshold;
      active_err_sum = h->active_err_sum;
      finished_errorest = h->extra_finished_errorest;
      num_active = static_cast<size_t>(h->num_active);
      pass_mem = h->pass_mem != 0;
      pass_budget = h->pass_budget != 0;
      out->probes += h->probes;
      if (search_debug)
        std::fprintf(stderr, "[pgn device search] n_local=%zu probes=%d computed=%d peer exchanges=%d cycles: kernel=%lld "
                             "scans=%lld exchanges=%lld elements scanned=%llu\n", n_, h->probes,
                     h->probes_computed, h->exchanges, h->cycles_total, h->cycles_scan, h->cycles_exchange, h->scanned);
    } else {
      fast = false;  // NaN / inf / negative errors: keep the reference's flag*err semantics exactly
    }
  }

  bool first_probe = true;
  auto probe = [&]() -> int {
    applied_threshold = threshold;
    if (fast) {
      ++launches_, k_query<<<1, kSearchBlock, 0, stream_>>>(threshold, plan, tables,
                                                   stats_dev_.as<double>() + kNumActive,
                                                   stats_dev_.as<double>() + kSumActErr);
      PGN_CUDA(cudaGetLastError());
      stats_fused_ = false;  // k_query is not a statistics kernel: its two numbers go through the communicator
      stats_published_ = false;
      if (first_probe) {
        // the number of values the buckets cannot represent travels in an otherwise unused slot
        ++launches_, k_outside_to_double<<<1, 1, 0, stream_>>>(tables.outside,
                                                      stats_dev_.as<double>() + kSumEst);
        PGN_CUDA(cudaGetLastError());
      }
      PGN_TRY(sync_stats(false));
      if (first_probe && host_global_[kSumEst] != 0.0)
        fast = false;  // some rank saw NaN / inf: redo this probe with the full pass
      first_probe = false;
    }
    if (!fast) {
      PGN_TRY(run_threshold(0, threshold));
      PGN_TRY(sync_stats(false));
    }
    out->probes++;
    num_active = static_cast<size_t>(host_global_[kNumActive]);
    active_err_sum = host_global_[kSumActErr];
    if (search_debug)
      std::fprintf(stderr, "[pgn probe] thr=%.17g active=%zu active_err=%.17g extra=%.17g quarter_budget=%.17g\n", threshold,
                   num_active, active_err_sum, iter_err - active_err_sum - fin_err,
                   max_percent_error_budget_ * (target_error - cum_err));
    const double pct =
      static_cast<double>(static_cast<int>(num_active)) / static_cast<double>(static_cast<int>(n_global));
    pass_mem = pct <= max_active_regions_percentage_;
    return PGN_OK;
  };

  if (!searched) {
    do {
      if (!pass_mem && increases <= 20) {
        pass_mem = false;
        size_t attempts = 0;
        while (!pass_mem && attempts < 20) {
          PGN_TRY(probe());
          if (!pass_mem) {
            range_lo = threshold;
            threshold += std::fabs(range_hi - threshold) * .5;
          }
          attempts++;
        }
        increases++;
      }
      if (pass_mem) {
        const double extra = iter_err - active_err_sum - fin_err;
        const double budget = target_error - cum_err;
        pass_budget = extra <= max_percent_error_budget_ * budget;
        finished_errorest = extra;
        if (!pass_budget && decreases <= 20) {
          range_hi = threshold;
          threshold -= std::fabs(range_lo - threshold) * .5;
          pass_mem = false;
          decreases++;
        }
      }
      const bool exhausted = decreases >= 20 || increases >= 20;
      if (exhausted && max_percent_error_budget_ < .7) {
        max_percent_error_budget_ += 0.1;
        decreases = 0;
        increases = 0;
        range_lo = min_err;
        range_hi = max_err;
        threshold = iter_err / static_cast<double>(n_global);
      } else if (exhausted && max_percent_error_budget_ >= .7 &&
                 max_active_regions_percentage_ <= .7) {
        max_active_regions_percentage_ += .1;
      } else if (exhausted) {
        break;
      }
    } while (!pass_mem || !pass_budget);
  }

  max_active_regions_percentage_ = .5;
  max_percent_error_budget_ = .25;

  if (search_debug) {
    std::fprintf(stderr, "[pgn search] %s n=%zu iter_err=%.17g fin_err=%.17g cum_err=%.17g target=%.17g min=%.17g max=%.17g"
                         " -> thr=%.17g applied=%.17g active=%zu active_err=%.17g extra=%.17g pass_mem=%d pass_budget=%d probes=%d\n",
                 searched ? "device" : "host", n_global, iter_err, fin_err, cum_err, target_error, min_err, max_err,
                 threshold, applied_threshold, num_active, active_err_sum, finished_errorest, int(pass_mem), int(pass_budget), out->probes);
  }

  out->success = pass_mem && pass_budget;
  out->num_active = num_active;
  out->finished_errorest = finished_errorest;
  out->threshold = applied_threshold;
  if (out->success) {
    // adopt the flags of the last probe (Workspace.cuh:104-112)
    PGN_TRY(run_threshold(1, applied_threshold));
    PGN_TRY(sync_stats(false));
    out->num_active = static_cast<size_t>(host_global_[kNumActive]);
  }
  out->must_terminate = (!out->success && necessary) || out->num_active == 0;
  return PGN_OK;
}

// ------------------------------------------------------------------------------------------------
// the loop (Workspace::integrate)
// ------------------------------------------------------------------------------------------------
int
Engine::integrate(pgn_eval_fn fn, void* user, const pgn_integrate_args& args, pgn_result_t* out)
{
  if (!fn || !out)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  const double epsrel = args.epsrel, epsabs = args.epsabs;
  const int relerr = args.relerr_classification;
  const int max_iters = opts_.max_iters > 0 ? opts_.max_iters : 700;
  const bool multi = comm_ && comm_->nranks > 1;
  const bool tracing = opts_.trace != 0;

  std::memset(out, 0, sizeof *out);
  out->status = 1;
  out->lastPhase = -1;
  trace_.clear();
  tracker_.reset(epsrel, epsabs);
  max_percent_error_budget_ = .25;
  max_active_regions_percentage_ = .5;

  if (!args.use_current_regions) {
    const int parts = ndim_ < 5 ? 4 : (ndim_ <= 10 ? 2 : 1);  // Workspace.cuh:379-385
    PGN_TRY(regions_init_uniform(parts));
  }
  if (filtered_pending_) {
    set_last_error("integrate() between filter() and split()");
    return PGN_ERR_STATE;
  }
  parents_n_ = 0;  // the reference starts every call without parent estimates (:165, :375)
  PGN_TRY(agree_mem_capacity());

  size_t n_global = n_;
  if (multi) {
    double h = static_cast<double>(n_);
    PGN_CUDA(cudaMemcpyAsync(stats_glob_.ptr, &h, sizeof h, cudaMemcpyHostToDevice, stream_));
    PGN_TRY(comm_->allreduce(stats_glob_.as<double>(), 1, kOpSum, stream_));
    PGN_CUDA(cudaMemcpyAsync(&h, stats_glob_.ptr, sizeof h, cudaMemcpyDeviceToHost, stream_));
    PGN_CUDA(cudaStreamSynchronize(stream_));
    n_global = static_cast<size_t>(h);
  }

  double cum_est = 0., cum_err = 0.;
  uint64_t nregions = 0, iters = 0, neval = 0;
  int status = 1;
  bool done = false;
  // Stage timing: six events per iteration from a grow-only pool, resolved once after the loop (an event
  // synchronisation per iteration would add a host round trip to every iteration of a traced run).
  enum { kEvStart = 0, kEvEvaluated, kEvClassified, kEvHeuristicBegin, kEvHeuristicEnd, kEvSplit, kEvPerIter };
  std::vector<unsigned char> ev_recorded;
  auto mark = [&](int it, int k) {
    if (!tracing)
      return;
    const size_t idx = static_cast<size_t>(it) * kEvPerIter + k;
    while (ev_pool_.size() <= idx) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess)
        return;
      ev_pool_.push_back(e);
    }
    if (ev_recorded.size() <= idx)
      ev_recorded.resize(idx + 1, 0);
    if (cudaEventRecord(ev_pool_[idx], stream_) == cudaSuccess)
      ev_recorded[idx] = 1;
  };

  // PGN_HOST_TIMING=1: where the host spends an integrate() call (stderr, one line per call)
  static const bool host_timing = std::getenv("PGN_HOST_TIMING") != nullptr;
  using Clock = std::chrono::steady_clock;
  double t_launch = 0, t_sync = 0, t_heur = 0, t_split = 0;
  const Clock::time_point t_call = Clock::now();
  auto since = [](Clock::time_point t0) { return std::chrono::duration<double, std::micro>(Clock::now() - t0).count(); };
  int n_iters_run = 0;

  for (int it = 0; it < max_iters && n_global > 0 && !done; ++it) {
    ++n_iters_run;
    pgn_iter_stats st;
    std::memset(&st, 0, sizeof st);
    st.iteration = it;
    st.nregions = n_;
    st.nregions_global = n_global;

    Clock::time_point t0 = Clock::now();
    if (it == 0)
      mark(it, kEvStart);  // later iterations start where the previous one's split ended
    PGN_TRY(launch_evaluate(fn, user));
    evaluated_ = true;
    mark(it, kEvEvaluated);
    PGN_TRY(run_classify(epsrel, relerr));
    mark(it, kEvClassified);
    t_launch += since(t0);
    t0 = Clock::now();
    PGN_TRY(sync_stats(false));
    t_sync += since(t0);
    neval += static_cast<uint64_t>(n_global) * rule_.feval;

    const double iter_est = host_global_[kSumEst];
    const double iter_err = host_global_[kSumErr];
    if (opts_.trace_arrays) {
      digest_arrays(&st, false);
      digest_arrays(&st, true);
    }
    st.iter_estimate = iter_est;
    st.iter_errorest = iter_err;

    if (args.iters_mode == PGN_ITERS_BEFORE_CHECK)
      iters++;

    // accuracy_reached (PaganiUtils.cuh:387-394)
    {
      const double e = std::fabs(cum_est + iter_est), r = cum_err + iter_err;
      if (r / e <= epsrel || r <= epsabs) {
        cum_est += iter_est;
        cum_err += iter_err;
        status = 0;
        nregions += n_global;
        st.cum_estimate = cum_est;
        st.cum_errorest = cum_err;
        if (tracing)
          trace_.push_back(st);
        done = true;
        break;
      }
    }

    tracker_.store(cum_est + iter_est);
    double fin_est = iter_est - host_global_[kSumActEst];  // compute_finished_estimates
    double fin_err = iter_err - host_global_[kSumActErr];
    size_t m_global = static_cast<size_t>(host_global_[kNumActive]);
    // the probes of the heuristic search overwrite the statistics: remember the accepted flags'
    size_t m_local = static_cast<size_t>(host_local_[kNumActive]);
    std::vector<size_t> counts = rank_counts_;

    // fix_error_budget_overflow (Workspace.cuh:121-146)
    if (cum_err + fin_err > std::fabs(cum_est + iter_est) * epsrel) {
      PGN_TRY(run_threshold(2, 0.0));
      PGN_TRY(sync_stats(false));
      fin_est = 0.;
      fin_err = 0.;
      m_global = static_cast<size_t>(host_global_[kNumActive]);
      m_local = static_cast<size_t>(host_local_[kNumActive]);
      counts = rank_counts_;
      st.budget_overflow = 1;
      if (opts_.trace_arrays)
        st.dig_active = digest_device(active_.ptr, 1, n_);
    }

    HeuristicOutcome hs;
    t0 = Clock::now();
    if (st.budget_overflow)
      mark(it, kEvHeuristicBegin);  // otherwise nothing ran since kEvClassified
    PGN_TRY(heuristic_classify(n_global, iter_est, iter_err, fin_err, cum_err, &hs));
    mark(it, kEvHeuristicEnd);
    t_heur += since(t0);
    st.heuristic_ran = hs.ran;
    st.heuristic_probes = hs.probes;
    st.threshold = hs.threshold;
    if (hs.success) {
      fin_est = iter_est - host_global_[kSumActEst];
      fin_err = hs.finished_errorest;
      m_global = hs.num_active;
      m_local = static_cast<size_t>(host_local_[kNumActive]);
      counts = rank_counts_;
      if (opts_.trace_arrays)
        st.dig_active = digest_device(active_.ptr, 1, n_);
    }
    if (hs.must_terminate) {
      cum_est += iter_est;
      cum_err += iter_err;
      nregions += n_global;
      st.cum_estimate = cum_est;
      st.cum_errorest = cum_err;
      st.nactive = m_global;
      if (tracing)
        trace_.push_back(st);
      done = true;
      break;
    }
    cum_est += fin_est;
    cum_err += fin_err;
    st.finished_estimate = fin_est;
    st.finished_errorest = fin_err;
    st.cum_estimate = cum_est;
    st.cum_errorest = cum_err;
    st.nactive = m_global;

    // filter + split
    nregions += n_global - m_global;
    if (m_global == 0) {
      n_ = 0;
      n_global = 0;
      parents_n_ = 0;
    } else {
      rank_counts_ = counts;
      bool rebalanced = false;
      t0 = Clock::now();
      PGN_TRY(compact_and_split(m_local, m_global, &rebalanced));
      t_split += since(t0);
      st.rebalanced = rebalanced;
      n_global = 2 * m_global;
    }
    mark(it, kEvSplit);
    if (tracing)
      trace_.push_back(st);
    evaluated_ = classified_ = false;
    if (args.iters_mode == PGN_ITERS_AFTER_SPLIT)
      iters++;
  }
  if (!done)
    nregions += n_global;
  PGN_CUDA(cudaStreamSynchronize(stream_));
  if (host_timing)
    std::fprintf(stderr, "[pgn host] %d iterations, %.0f us in the call: launching evaluate+classify %.0f, waiting for the "
                         "statistics %.0f, heuristic %.0f, compact/split launch %.0f\n",
                 n_iters_run, since(t_call), t_launch, t_sync, t_heur, t_split);
  if (tracing) {
    auto elapsed = [&](int it, int a, int b) -> double {
      size_t ia = static_cast<size_t>(it) * kEvPerIter + a;
      const size_t ib = static_cast<size_t>(it) * kEvPerIter + b;
      // events that were not recorded because they coincide with an earlier one
      if (a == kEvStart && it > 0)
        ia = static_cast<size_t>(it - 1) * kEvPerIter + kEvSplit;
      if (a == kEvHeuristicBegin && (ia >= ev_recorded.size() || !ev_recorded[ia]))
        ia = static_cast<size_t>(it) * kEvPerIter + kEvClassified;
      if (ia >= ev_recorded.size() || ib >= ev_recorded.size() || !ev_recorded[ia] || !ev_recorded[ib])
        return 0.0;
      float ms = 0;
      return cudaEventElapsedTime(&ms, ev_pool_[ia], ev_pool_[ib]) == cudaSuccess ? ms : 0.0;
    };
    for (pgn_iter_stats& t : trace_) {
      const int it = static_cast<int>(t.iteration);
      t.eval_ms = elapsed(it, kEvStart, kEvEvaluated);
      t.classify_ms = elapsed(it, kEvEvaluated, kEvClassified);
      t.heuristic_ms = elapsed(it, kEvHeuristicBegin, kEvHeuristicEnd);
      t.split_ms = elapsed(it, kEvHeuristicEnd, kEvSplit);
    }
  }

  out->estimate = cum_est;
  out->errorest = cum_err;
  out->neval = neval;
  out->nregions = nregions;
  out->nFinishedRegions = 0;
  out->status = status;
  out->lastPhase = -1;
  out->chi_sq = 0.;
  out->iters = iters;
  return PGN_OK;
}

// ------------------------------------------------------------------------------------------------
int
Engine::comm_init_nccl(const uint8_t* id, int nranks, int rank)
{
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
    return PGN_ERR_INVALID;
  PGN_CUDA(cudaSetDevice(device_));
  delete comm_;
  comm_ = nullptr;
  std::string err;
  comm_ = make_nccl_comm(id, nranks, rank, &err);
  if (!comm_) {
    set_last_error("NCCL: " + err);
    return PGN_ERR_NCCL;
  }
  rank_counts_.assign(nranks, 0);
  if (nranks > 1 && !std::getenv("PGN_NO_PEER_SEARCH")) {
    const int rc = setup_peer_mailboxes();
    if (rc != PGN_OK)
      return rc;
  }
  return PGN_OK;
}

// Every rank allocates a mailbox, publishes its cudaIpc handle (the 64 bytes travel as 64 doubles
// through one SUM all-reduce of a one-hot table) and maps the mailboxes of all peers.
int
Engine::setup_peer_mailboxes()
{
  release_peer_mailboxes();
  const int w = comm_->nranks, me = comm_->rank;
  // A rank that fails locally (allocation, IPC export) must still take part in both all-reduces below -- its peers
  // would otherwise wait in them for ever -- and simply votes against the in-kernel path.
  bool local_ok = mailbox_.reserve(kMailboxBytes) == PGN_OK;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof mine);
  if (local_ok && (cudaMemset(mailbox_.ptr, 0, kMailboxBytes) != cudaSuccess ||
                   cudaIpcGetMemHandle(&mine, mailbox_.ptr) != cudaSuccess)) {
    cudaGetLastError();
    local_ok = false;
  }
  constexpr int kH = static_cast<int>(sizeof(cudaIpcMemHandle_t));
  std::vector<double> table(static_cast<size_t>(w) * kH, 0.0);
  const unsigned char* bytes = reinterpret_cast<const unsigned char*>(&mine);
  for (int i = 0; i < kH; ++i)
    table[static_cast<size_t>(me) * kH + i] = static_cast<double>(bytes[i]);
  // scratch: the (grow-only) rebalance staging buffer, idle here
  PGN_TRY(pack_send_.reserve(sizeof(double) * table.size()));
  double* xchg = pack_send_.as<double>();
  PGN_CUDA(cudaMemcpyAsync(xchg, table.data(), sizeof(double) * table.size(), cudaMemcpyHostToDevice, stream_));
  PGN_TRY(comm_->allreduce(xchg, static_cast<int>(table.size()), kOpSum, stream_));
  PGN_CUDA(cudaMemcpyAsync(table.data(), xchg, sizeof(double) * table.size(), cudaMemcpyDeviceToHost, stream_));
  PGN_CUDA(cudaStreamSynchronize(stream_));
  peer_mapped_.assign(w, nullptr);
  bool ok = local_ok;
  for (int q = 0; q < w && ok; ++q) {
    if (q == me)
      continue;
    cudaIpcMemHandle_t h;
    unsigned char* hb = reinterpret_cast<unsigned char*>(&h);
    for (int i = 0; i < kH; ++i)
      hb[i] = static_cast<unsigned char>(table[static_cast<size_t>(q) * kH + i]);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
      cudaGetLastError();
      ok = false;  // no peer access between these devices: the host-driven search is used instead
    } else {
      peer_mapped_[q] = p;
    }
  }
  // every rank must agree on whether the in-kernel path is usable
  double agree = ok ? 0.0 : 1.0;
  PGN_CUDA(cudaMemcpyAsync(xchg, &agree, sizeof agree, cudaMemcpyHostToDevice, stream_));
  PGN_TRY(comm_->allreduce(xchg, 1, kOpSum, stream_));
  PGN_CUDA(cudaMemcpyAsync(&agree, xchg, sizeof agree, cudaMemcpyDeviceToHost, stream_));
  PGN_CUDA(cudaStreamSynchronize(stream_));
  peer_ready_ = agree == 0.0;
  peer_seq_ = 0;
  if (!peer_ready_)
    release_peer_mailboxes();
  return PGN_OK;
}

void
Engine::release_peer_mailboxes()
{
  for (void* p : peer_mapped_)
    if (p)
      cudaIpcCloseMemHandle(p);
  peer_mapped_.clear();
  peer_ready_ = false;
}

int
Engine::comm_init_host(int nranks, int rank, pgn_allreduce_fn ar, pgn_alltoallv_fn a2a, void* user)
{
  if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks || !ar || !a2a)
    return PGN_ERR_INVALID;
  delete comm_;
  comm_ = make_host_comm(nranks, rank, ar, a2a, user);
  rank_counts_.assign(nranks, 0);
  return PGN_OK;
}

int
Engine::comm_size(int* nranks, int* rank) const
{
  if (nranks)
    *nranks = comm_ ? comm_->nranks : 1;
  if (rank)
    *rank = comm_ ? comm_->rank : 0;
  return PGN_OK;
}

}  // namespace pgn
