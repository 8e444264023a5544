// TEST INFRASTRUCTURE -- the oracle.  Nothing in the product may include, link or call this file;
// only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg use it, as the checker.
//
// Plain sequential C++ restatement of the reference's PAGANI (marcpaterno/gpuintegration, CUDA
// back-end).  Each function cites the reference lines it follows.  The rule tables are NOT built
// here: callers pass the dense tables, which the tests take from tests/golden/rule_tables.npz, i.e.
// from the reference's own host code (oracle/ref/ref_rule_dump.cu).
//
// Pinning: tests/test_oracle_golden.py checks this file against per-iteration trajectories recorded
// from the reference binary on a B200 (tests/golden/trajectories/*.json, made by
// tools/make_golden_trajectories.py) and against the reference tests' known answers, among them the
// heuristic classifier's own (test/kokkos/pagani/test_heuristic_classifier.cpp:32-103) through
// orc_classify_threshold.
// Host libm (cos/exp/pow) is not bit-identical to CUDA's libdevice, so per-region values agree to
// ~1e-15 relative rather than bit-for-bit; the golden comparison states its tolerances.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <vector>

#include "pagani_b200/genz.cuh"

namespace {

constexpr int kRules = 5, kSets = 9, kBlock = 64;

struct Rule {
  int ndim, feval;
  const double* gens;     // [ndim][feval]  (ComputeGenerators output, Phases.cuh:157-217)
  const double* weights;  // [9][5]
  const double* scale;    // [9][5]
  const double* norm;     // [9][5]
  const int32_t* set_of_point;
  const double* G;        // [9][ndim] generator prototypes (only G[1][0], G[2][0] are used)
};

using Integrand = double (*)(const double*);

template <typename F, size_t... I>
double
call(const double* x, std::index_sequence<I...>)
{
  F f;
  return f(x[I]...);
}
template <typename F, int N>
double
thunk(const double* x)
{
  return call<F>(x, std::make_index_sequence<N>());
}

Integrand
lookup(int family, int ndim)
{
  using namespace pagani_b200::genz;
#define ORC_CASE(fam, T, n)                                                                        \
  if (family == fam && ndim == n)                                                                  \
    return &thunk<T<n>, n>;
  ORC_CASE(1, Oscillatory, 5) ORC_CASE(1, Oscillatory, 6) ORC_CASE(1, Oscillatory, 7)
  ORC_CASE(1, Oscillatory, 8) ORC_CASE(2, ProductPeak, 5) ORC_CASE(2, ProductPeak, 6)
  ORC_CASE(2, ProductPeak, 7) ORC_CASE(2, ProductPeak, 8) ORC_CASE(3, CornerPeak, 5)
  ORC_CASE(3, CornerPeak, 6) ORC_CASE(3, CornerPeak, 7) ORC_CASE(3, CornerPeak, 8)
  ORC_CASE(4, Gaussian, 5) ORC_CASE(4, Gaussian, 6) ORC_CASE(4, Gaussian, 7) ORC_CASE(4, Gaussian, 8)
  ORC_CASE(5, CZero, 5) ORC_CASE(5, CZero, 6) ORC_CASE(5, CZero, 7) ORC_CASE(5, CZero, 8)
  ORC_CASE(5, CZero, 10) ORC_CASE(6, Discontinuous, 5) ORC_CASE(6, Discontinuous, 6)
  ORC_CASE(6, Discontinuous, 7) ORC_CASE(6, Discontinuous, 8) ORC_CASE(6, Discontinuous, 10)
#undef ORC_CASE
  return nullptr;
}

// family 0: the constant 15.37 of RegionSampling.cu / Easy_Integrals.cu; family -1: product of
// the coordinates (do_integration_from_c.cu uses x*y)
double
constant_fn(const double*)
{
  return 15.37;
}
int g_prod_ndim = 2;
double
product_fn(const double* x)
{
  double p = x[0];
  for (int d = 1; d < g_prod_ndim; ++d)
    p = p * x[d];
  return p;
}

// One region: INIT_REGION_POOL + SampleRegionBlock + computePermutation
// (Phases.cuh:301-404, Sample.cuh:103-290), with the 64 threads of the block emulated one by one.
void
evaluate_region(const Rule& R, Integrand f, const double* left, const double* length, size_t n,
                size_t reg, const double* lows, const double* highs, double* est, double* err,
                int32_t* axis)
{
  const int nd = R.ndim;
  double lower[16], upper[16], range[16];
  double jacobian = 1., vol = 1., max_range = 0;
  int max_dim = 0;
  for (int d = 0; d < nd; ++d) {  // Phases.cuh:330-348
    lower[d] = left[(size_t)d * n + reg];
    upper[d] = lower[d] + length[(size_t)d * n + reg];
    vol *= upper[d] - lower[d];
    range[d] = highs[d] - lows[d];
    const double ext = upper[d] - lower[d];
    jacobian = jacobian * range[d];
    if (ext > max_range) {
      max_dim = d;
      max_range = ext;
    }
  }
  double sum[kBlock][kRules];
  std::memset(sum, 0, sizeof sum);
  double sdata[kBlock];
  std::memset(sdata, 0, sizeof sdata);
  double x[16];
  for (int p = 0; p < R.feval; ++p) {  // thread p%64 visits p, p+64, ... in this order
    const int t = p % kBlock;
    for (int d = 0; d < nd; ++d) {  // Sample.cuh:124-131 (contraction as in the reference's PTX)
      const double g = R.gens[(size_t)d * R.feval + p];
      const double inner = std::fma(.5 + g, lower[d], (.5 - g) * upper[d]);
      x[d] = std::fma(range[d], inner, lows[d]);
    }
    const double fun = f(x) * jacobian;  // :133
    if (p < kBlock)
      sdata[t] = fun;  // :134 (only the first pass is read back)
    const int s = R.set_of_point[p];
    for (int r = 0; r < kRules; ++r)  // :146-148
      sum[t][r] = std::fma(fun, R.weights[s * kRules + r], sum[t][r]);
  }
  // fourth differences, Sample.cuh:194-218
  int bisect = max_dim;
  {
    const double q = R.G[2 * nd] / R.G[1 * nd];
    const double ratio = q * q;
    const double base = (sdata[0] + sdata[0]) * (1 - ratio);
    double maxdiff = 0;
    for (int d = 0; d < nd; ++d) {
      const double fp = sdata[1 + 2 * d], fm = sdata[2 + 2 * d];
      const double gp = sdata[1 + 2 * d + 2 * nd], gm = sdata[2 + 2 * d + 2 * nd];
      // `base + ratio * (fp + fm) - ...`: which product nvcc 12.9 fuses depends on ndim.  Up to 4 axes the loop is
      // unrolled before contraction and the base product goes into the FMA (the other one is rounded); from 5 axes
      // on base is a rounded loop invariant (read off the reference kernels' SASS for ndim 2..11).
      const double t = nd <= 4 ? std::fma(sdata[0] + sdata[0], 1 - ratio, ratio * (fp + fm)) : std::fma(ratio, fp + fm, base);
      const double fourth = std::fabs(t - (gp + gm));
      if (fourth > maxdiff) {
        maxdiff = fourth;
        bisect = d;
      }
    }
  }
  // blockReduceSum, Sample.cuh:49-83: shuffle-down tree inside each warp, then warp0 + warp1
  double tot[kRules];
  for (int r = 0; r < kRules; ++r) {
    double w[2];
    for (int wp = 0; wp < 2; ++wp) {
      double v[32];
      for (int i = 0; i < 32; ++i)
        v[i] = sum[32 * wp + i][r];
      for (int off = 16; off > 0; off >>= 1)
        for (int i = 0; i < off; ++i)
          v[i] = v[i] + v[i + off];
      w[wp] = v[0];
    }
    tot[r] = w[0] + w[1];
  }
  // null-rule error, Sample.cuh:261-289
  double e[kRules];
  for (int r = 0; r < kRules; ++r)
    e[r] = tot[r];
  for (int rul = 1; rul < kRules - 1; ++rul) {
    double maxerr = 0.;
    for (int s = 0; s < kSets; ++s)
      maxerr = std::fmax(maxerr, std::fabs(std::fma(R.scale[s * kRules + rul], tot[rul], tot[rul + 1])) *
                                   R.norm[s * kRules + rul]);
    e[rul] = maxerr;
  }
  est[reg] = vol * tot[0];
  err[reg] = vol * ((5. * e[1] <= e[2] && 5. * e[2] <= e[3]) ? e[1] : 5. * std::fmax(std::fmax(e[1], e[2]), e[3]));
  axis[reg] = bisect;
}

uint64_t
digest(const void* data, size_t elem, size_t n)
{
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint64_t acc = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t bits = 0;
    std::memcpy(&bits, p + i * elem, elem);
    acc += bits * (2ull * i + 1ull);
  }
  return acc;
}

// Heuristic_classifier host state (heuristic_classifier.cuh:147-225)
struct Tracker {
  double epsrel, epsabs;
  int required_digits;
  double last[3] = {0, 0, 0};
  size_t collected = 0;
  void
  store(double v)
  {
    last[0] = last[1];
    last[1] = last[2];
    last[2] = v;
    collected++;
  }
  bool
  sig_digits_same() const
  {  // :169-203
    auto scale = [](double v) {
      v = std::fabs(v);
      while (v != 0. && v < 1.)
        v *= 10;
      return v;
    };
    char a[512], b[512], c[512];
    snprintf(a, sizeof a, "%.15f", scale(last[0]));
    snprintf(b, sizeof b, "%.15f", scale(last[1]));
    snprintf(c, sizeof c, "%.15f", scale(last[2]));
    const size_t la = strlen(a), lb = strlen(b), lc = strlen(c);
    bool verdict = true;
    int sig = 0;
    for (int i = 0; i < required_digits + 1 && sig < required_digits && verdict; ++i) {
      const char ca = (size_t)i < la ? a[i] : 0, cb = (size_t)i < lb ? b[i] : 0,
                 cc = (size_t)i < lc ? c[i] : 0;
      verdict = cc == cb && cb == ca;
      sig += (verdict && cc != '.') ? 1 : 0;
    }
    return verdict;
  }
  bool
  converged() const
  {  // :205-216
    return !(collected - 1 < 1 || !sig_digits_same());
  }
};


// Heuristic_classifier::classify (heuristic_classifier.cuh:362-451) on a host array.  The flags left in
// `flags` are those of the LAST apply_threshold call -- what the reference adopts (Workspace.cuh:104-106) --
// also when the exhausted-attempts branch has reset `thr` after that probe succeeded.
struct SearchOut {
  bool pass_mem = false, pass_budget = false;
  size_t num_active = 0;
  double finished_errorest = 0., threshold = 0., budget_pct = .25, active_pct = .5;
  int probes = 0;
};

SearchOut
classify_threshold(const double* err, size_t n, double iter_err, double fin_err, double cum_err, double target,
                   std::vector<uint8_t>& flags)
{
  SearchOut o;
  double mn = std::numeric_limits<double>::infinity(), mx = -mn;
  for (size_t i = 0; i < n; ++i) {
    mn = std::fmin(mn, err[i]);
    mx = std::fmax(mx, err[i]);
  }
  double lo = mn, hi = mx, thr = iter_err / (double)n;
  double budget_pct = .25, active_pct = .5;
  bool pass_mem = false, pass_budget = false;
  size_t cnt = 0, inc = 0, dec = 0;
  double act_sum = 0, hs_fin = 0;
  flags.assign(n, 0);
  auto apply = [&]() {
    cnt = 0;
    act_sum = 0;
    for (size_t i = 0; i < n; ++i) {
      flags[i] = err[i] > thr;
      cnt += flags[i];
      act_sum += flags[i] * err[i];
    }
    pass_mem = (double)(int)cnt / (double)(int)n <= active_pct;
    o.probes++;
  };
  do {
    if (!pass_mem && inc <= 20) {
      pass_mem = false;
      size_t attempts = 0;
      while (!pass_mem && attempts < 20) {
        apply();
        if (!pass_mem) {
          lo = thr;
          thr += std::fabs(hi - thr) * .5;
        }
        attempts++;
      }
      inc++;
    }
    if (pass_mem) {
      const double extra = iter_err - act_sum - fin_err;
      pass_budget = extra <= budget_pct * (target - cum_err);
      hs_fin = extra;
      if (!pass_budget && dec <= 20) {
        hi = thr;
        thr -= std::fabs(lo - thr) * .5;
        pass_mem = false;
        dec++;
      }
    }
    const bool exhausted = dec >= 20 || inc >= 20;
    if (exhausted && budget_pct < .7) {
      budget_pct += 0.1;
      dec = inc = 0;
      lo = mn;
      hi = mx;
      thr = iter_err / (double)n;
    } else if (exhausted && budget_pct >= .7 && active_pct <= .7) {
      active_pct += .1;
    } else if (exhausted) {
      break;
    }
  } while (!pass_mem || !pass_budget);
  o.pass_mem = pass_mem;
  o.pass_budget = pass_budget;
  o.num_active = cnt;
  o.finished_errorest = hs_fin;
  o.threshold = thr;
  o.budget_pct = budget_pct;
  o.active_pct = active_pct;
  return o;
}

}  // namespace

extern "C" {

struct orc_iter {
  uint64_t n, nactive;
  double iter_estimate, iter_errorest, finished_estimate, finished_errorest, cum_estimate, cum_errorest;
  int32_t heuristic_ran, budget_overflow;
  uint64_t dig_left, dig_length, dig_estimate, dig_errorest_raw, dig_errorest, dig_axis, dig_active;
};

struct orc_result {
  double estimate, errorest;
  uint64_t neval, nregions, iters;
  int32_t status, niter_records;
};

// Cubature_rules::apply_cubature_integration_rules on a caller-supplied region list
// (PaganiUtils.cuh:291-343): fills est/err/axis for n regions.
int
orc_evaluate(int family, int ndim, int feval, const double* gens, const double* weights,
             const double* scale, const double* norm, const int32_t* set_of_point, const double* G,
             const double* left, const double* length, uint64_t n, const double* lows,
             const double* highs, double* est, double* err, int32_t* axis)
{
  Integrand f = family == 0 ? &constant_fn : (family == -1 ? &product_fn : lookup(family, ndim));
  if (!f)
    return -1;
  g_prod_ndim = ndim;
  Rule R{ndim, feval, gens, weights, scale, norm, set_of_point, G};
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t r = 0; r < (int64_t)n; ++r)
    evaluate_region(R, f, left, length, n, (size_t)r, lows, highs, est, err, axis);
  return 0;
}

// quad::RefineError (Phases.cuh:219-268)
void
orc_refine(const double* est, const double* err_raw, const double* parent_est, uint64_t n,
           double epsrel, int relerr_classification, double* err_out, uint8_t* active)
{
  const uint64_t half = n / 2;
  for (uint64_t i = 0; i < n; ++i) {
    double self_err = err_raw[i];
    const double self_res = est[i];
    const bool right = 2 * i >= n;
    const uint64_t sib = right ? i - half : i + half;
    const uint64_t par = right ? i - half : i;
    double diff = est[sib] + self_res - parent_est[par];
    diff = std::fabs(.25 * diff);
    const double both = self_err + err_raw[sib];
    if (both > 0.0) {
      const double c = 1 + 2 * diff / both;
      self_err *= c;
    }
    self_err += diff;
    err_out[i] = self_err;
    const bool pass = relerr_classification && self_err < std::fmax(epsrel * std::fabs(self_res), 1e-200);
    active[i] = !pass;
  }
}

// Workspace::integrate (Workspace.cuh:361-494), sequential.  `records` (may be null) receives up to
// max_records per-iteration summaries.  mem_bytes is the device size the heuristic assumes
// (16e9 in the reference, heuristic_classifier.cuh:89-94).
// Heuristic_classifier::classify on a caller-supplied error list (the reference's classifier KAT,
// test/kokkos/pagani/test_heuristic_classifier.cpp:32-103).  last_estimate is what store_estimate() saw last.
// out[0..6] = pass_mem, pass_errorest_budget, num_active, finished_errorest, max_budget_perc_to_cover,
// max_active_perc, probes, threshold (the variable's final value); flags[n] receives the adopted flags.
void
orc_classify_threshold(const double* err, uint64_t n, double iter_errorest, double iter_finished_errorest,
                       double total_finished_errorest, double last_estimate, double epsrel, uint8_t* flags,
                       double* out)
{
  std::vector<uint8_t> f;
  const SearchOut so = classify_threshold(err, n, iter_errorest, iter_finished_errorest, total_finished_errorest,
                                          std::fabs(last_estimate) * epsrel, f);
  for (uint64_t i = 0; i < n; ++i)
    flags[i] = f[i];
  out[0] = so.pass_mem;
  out[1] = so.pass_budget;
  out[2] = static_cast<double>(so.num_active);
  out[3] = so.finished_errorest;
  out[4] = so.budget_pct;
  out[5] = so.active_pct;
  out[6] = so.probes;
  out[7] = so.threshold;
}

int
orc_integrate(int family, int ndim, int feval, const double* gens, const double* weights,
              const double* scale, const double* norm, const int32_t* set_of_point, const double* G,
              double epsrel, double epsabs, const double* lows, const double* highs,
              int relerr_classification, double mem_bytes, int max_iters, uint64_t max_regions,
              orc_result* out, orc_iter* records, int max_records)
{
  Integrand f = family == 0 ? &constant_fn : (family == -1 ? &product_fn : lookup(family, ndim));
  if (!f || !out)
    return -1;
  g_prod_ndim = ndim;
  Rule R{ndim, feval, gens, weights, scale, norm, set_of_point, G};
  const size_t nd = (size_t)ndim;

  // Sub_regions::uniform_split (Sub_regions.cuh:11-36, 164-185)
  const int parts = ndim < 5 ? 4 : (ndim <= 10 ? 2 : 1);
  size_t n = 1;
  for (int d = 0; d < ndim; ++d)
    n *= parts;
  std::vector<double> left(n * nd), length(n * nd);
  const double edge = 1. / (double)parts;
  for (size_t t = 0; t < n; ++t) {
    size_t q = t;
    for (int d = 0; d < ndim; ++d) {
      left[(size_t)d * n + t] = (double)(q % parts) * edge;
      length[(size_t)d * n + t] = edge;
      q /= parts;
    }
  }

  Tracker tracker;
  tracker.epsrel = epsrel;
  tracker.epsabs = epsabs;
  tracker.required_digits = (int)std::ceil(std::log10(1 / epsrel));

  std::vector<double> parent_est;
  double cum_est = 0, cum_err = 0;
  uint64_t nregions = 0, iters = 0, neval = 0;
  int status = 1, nrec = 0;
  bool done = false;

  for (int it = 0; it < max_iters && n > 0 && !done; ++it) {
    if (n > max_regions)
      return -2;  // the caller bounds oracle work
    std::vector<double> est(n), err_raw(n), err(n);
    std::vector<int32_t> axis(n);
    std::vector<uint8_t> active(n, 1);
#pragma omp parallel for schedule(dynamic, 16)
    for (int64_t r = 0; r < (int64_t)n; ++r)
      evaluate_region(R, f, left.data(), length.data(), n, (size_t)r, lows, highs, est.data(),
                      err_raw.data(), axis.data());
    neval += (uint64_t)n * feval;
    double iter_est = 0;
    for (size_t i = 0; i < n; ++i)
      iter_est += est[i];
    if (!parent_est.empty())  // hybrid.cuh:26-28
      orc_refine(est.data(), err_raw.data(), parent_est.data(), n, epsrel, relerr_classification,
                 err.data(), active.data());
    else
      err = err_raw;
    double iter_err = 0;
    for (size_t i = 0; i < n; ++i)
      iter_err += err[i];

    orc_iter rec;
    std::memset(&rec, 0, sizeof rec);
    rec.n = n;
    rec.iter_estimate = iter_est;
    rec.iter_errorest = iter_err;
    rec.dig_left = digest(left.data(), 8, left.size());
    rec.dig_length = digest(length.data(), 8, length.size());
    rec.dig_estimate = digest(est.data(), 8, n);
    rec.dig_errorest_raw = digest(err_raw.data(), 8, n);
    rec.dig_errorest = digest(err.data(), 8, n);
    rec.dig_axis = digest(axis.data(), 4, n);
    auto push = [&]() {
      rec.dig_active = digest(active.data(), 1, n);
      rec.cum_estimate = cum_est;
      rec.cum_errorest = cum_err;
      if (records && nrec < max_records)
        records[nrec] = rec;
      nrec++;
    };

    {  // accuracy_reached, PaganiUtils.cuh:387-394
      const double e = std::fabs(cum_est + iter_est), r = cum_err + iter_err;
      if (r / e <= epsrel || r <= epsabs) {
        cum_est += iter_est;
        cum_err += iter_err;
        status = 0;
        nregions += n;
        push();
        done = true;
        break;
      }
    }
    tracker.store(cum_est + iter_est);
    // compute_finished_estimates, PaganiUtils.cuh:368-385
    double act_est = 0, act_err = 0;
    size_t nactive = 0;
    for (size_t i = 0; i < n; ++i) {
      act_est += active[i] * est[i];
      act_err += active[i] * err[i];
      nactive += active[i];
    }
    double fin_est = iter_est - act_est, fin_err = iter_err - act_err;
    // fix_error_budget_overflow, Workspace.cuh:121-146
    if (cum_err + fin_err > std::fabs(cum_est + iter_est) * epsrel) {
      std::fill(active.begin(), active.end(), 1);
      nactive = n;
      fin_est = fin_err = 0.;
      rec.budget_overflow = 1;
    }
    // heuristic_classify, Workspace.cuh:74-118
    bool must_terminate = false;
    {
      const size_t required = 8 * (n * (10 * nd + 8)) + 4 * (4 * n);  // heuristic_classifier.cuh:105-133
      const size_t total = (size_t)mem_bytes;
      const size_t free_mem = total > required ? total - required : 1;
      const double ratio = (double)required / (double)free_mem;
      const bool necessary = ratio > 1.;
      const bool criteria = ratio > 1. || (ratio > .1 && tracker.converged());
      if (!criteria) {
        must_terminate = necessary;
      } else {
        rec.heuristic_ran = 1;
        // Heuristic_classifier::classify, heuristic_classifier.cuh:362-451
        std::vector<uint8_t> flags;
        const SearchOut so = classify_threshold(err.data(), n, iter_err, fin_err, cum_err,
                                                std::fabs(tracker.last[2]) * epsrel, flags);
        const bool pass_mem = so.pass_mem, pass_budget = so.pass_budget;
        const size_t cnt = so.num_active;
        const double hs_fin = so.finished_errorest;
        const bool success = pass_mem && pass_budget;
        if (success) {
          active = flags;
          double dot = 0;
          for (size_t i = 0; i < n; ++i)
            dot += active[i] * est[i];
          fin_est = iter_est - dot;
          fin_err = hs_fin;
          nactive = cnt;
        }
        must_terminate = (!success && necessary) || cnt == 0;
      }
    }
    rec.nactive = nactive;
    rec.finished_estimate = fin_est;
    rec.finished_errorest = fin_err;
    if (must_terminate) {
      cum_est += iter_est;
      cum_err += iter_err;
      nregions += n;
      push();
      done = true;
      break;
    }
    cum_est += fin_est;
    cum_err += fin_err;
    push();
    // Sub_regions_filter::filter + Sub_region_splitter::split
    // (Sub_region_filter.cuh:8-47, Sub_region_splitter.cuh:8-46)
    const size_t m = nactive;
    nregions += n - m;
    std::vector<double> nl(2 * m * nd), nw(2 * m * nd);
    parent_est.assign(m, 0.);
    size_t j = 0;
    for (size_t i = 0; i < n; ++i) {
      if (!active[i])
        continue;
      parent_est[j] = est[i];
      for (size_t d = 0; d < nd; ++d) {
        const double lo_ = left[d * n + i], ln = length[d * n + i];
        if ((int)d == axis[i]) {
          const double h = ln / 2;
          nl[d * 2 * m + j] = lo_ + 0 * h;
          nl[d * 2 * m + m + j] = lo_ + 1 * h;
          nw[d * 2 * m + j] = h;
          nw[d * 2 * m + m + j] = h;
        } else {
          nl[d * 2 * m + j] = lo_;
          nl[d * 2 * m + m + j] = lo_;
          nw[d * 2 * m + j] = ln;
          nw[d * 2 * m + m + j] = ln;
        }
      }
      ++j;
    }
    left.swap(nl);
    length.swap(nw);
    n = 2 * m;
    iters++;
  }
  if (!done)
    nregions += n;
  out->estimate = cum_est;
  out->errorest = cum_err;
  out->neval = neval;
  out->nregions = nregions;
  out->iters = iters;
  out->status = status;
  out->niter_records = nrec;
  return 0;
}

}  // extern "C"
