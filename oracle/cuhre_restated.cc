// TEST / BASELINE INFRASTRUCTURE -- "Cuhre (restated, not libcuba)".
//
// The reference's CPU comparison is cubacpp::Cuhre, a thin wrapper over llCuhre of CUBA 4.2
// (cubacpp/cubacpp/cuhre.hh:71-89; CUBA pinned only by cubacpp/scripts/build_sharedlib.sh:14-24).
// CUBA is not vendored in the reference tree and cannot be fetched here, so its published
// algorithm is restated: the classic sequential globally-adaptive loop (pop the region with the
// largest error, bisect it along the axis with the largest fourth difference, evaluate both halves
// with the degree-9 rule -- key 9 for ndim >= 4, cuhre.hh:16-35 --, apply the two-level error
// refinement, stop when sum(err) <= max(epsabs, epsrel*|sum(est)|) or maxeval is reached).
// The per-region arithmetic is the one the reference's GPU code documents (rule tables
// Rule.cuh:24-270, error Sample.cuh:261-289, axis Sample.cuh:194-218, refinement Phases.cuh:251-261).
// PARITY UNPINNED: the reference's own Cuhre tests (cubacpp/test/integration.test.cc:95-153) pin
// accuracy only, not evaluation counts, so this is a reported baseline, nothing more.
#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <queue>
#include <vector>

#include "pagani_b200/genz.cuh"

namespace {

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
#define CUH_CASE(fam, T, n)                                                                        \
  if (family == fam && ndim == n)                                                                  \
    return &thunk<T<n>, n>;
  CUH_CASE(1, Oscillatory, 5) CUH_CASE(1, Oscillatory, 6) CUH_CASE(1, Oscillatory, 7)
  CUH_CASE(1, Oscillatory, 8) CUH_CASE(2, ProductPeak, 5) CUH_CASE(2, ProductPeak, 6)
  CUH_CASE(2, ProductPeak, 7) CUH_CASE(2, ProductPeak, 8) CUH_CASE(3, CornerPeak, 5)
  CUH_CASE(3, CornerPeak, 6) CUH_CASE(3, CornerPeak, 7) CUH_CASE(3, CornerPeak, 8)
  CUH_CASE(4, Gaussian, 5) CUH_CASE(4, Gaussian, 6) CUH_CASE(4, Gaussian, 7) CUH_CASE(4, Gaussian, 8)
  CUH_CASE(5, CZero, 5) CUH_CASE(5, CZero, 6) CUH_CASE(5, CZero, 7) CUH_CASE(5, CZero, 8)
  CUH_CASE(5, CZero, 10) CUH_CASE(6, Discontinuous, 5) CUH_CASE(6, Discontinuous, 6)
  CUH_CASE(6, Discontinuous, 7) CUH_CASE(6, Discontinuous, 8) CUH_CASE(6, Discontinuous, 10)
#undef CUH_CASE
  return nullptr;
}

struct Region {
  double avg, err;
  int bisect;
  std::vector<double> lo, len;
};

struct Tables {
  int ndim, feval;
  const double *gens, *weights, *scale, *norm, *G;
  const int32_t* set_of_point;
};

// degree-9 rule on one box; `fvals` is scratch of feval doubles
void
sample(const Tables& T, Integrand f, Region& r, std::vector<double>& fvals, int nthreads)
{
  const int nd = T.ndim, fe = T.feval;
  double vol = 1.;
  for (int d = 0; d < nd; ++d)
    vol *= r.len[d];
#pragma omp parallel for schedule(static) num_threads(nthreads) if (nthreads > 1)
  for (int p = 0; p < fe; ++p) {
    double x[16];
    for (int d = 0; d < nd; ++d) {
      const double g = T.gens[(size_t)d * fe + p];
      x[d] = (.5 + g) * r.lo[d] + (.5 - g) * (r.lo[d] + r.len[d]);
    }
    fvals[p] = f(x);
  }
  double sum[5] = {0, 0, 0, 0, 0};
  for (int p = 0; p < fe; ++p) {
    const double* w = &T.weights[T.set_of_point[p] * 5];
    for (int k = 0; k < 5; ++k)
      sum[k] += fvals[p] * w[k];
  }
  // split axis
  const double q = T.G[2 * nd] / T.G[nd];
  const double ratio = q * q;
  const double base = 2 * fvals[0] * (1 - ratio);
  double maxdiff = 0, maxlen = 0;
  int bisect = 0;
  for (int d = 0; d < nd; ++d)
    if (r.len[d] > maxlen) {
      maxlen = r.len[d];
      bisect = d;
    }
  for (int d = 0; d < nd; ++d) {
    const double fourth = std::fabs(base + ratio * (fvals[1 + 2 * d] + fvals[2 + 2 * d]) -
                                    (fvals[1 + 2 * d + 2 * nd] + fvals[2 + 2 * d + 2 * nd]));
    if (fourth > maxdiff) {
      maxdiff = fourth;
      bisect = d;
    }
  }
  double e[5];
  for (int rul = 1; rul < 4; ++rul) {
    double maxerr = 0;
    for (int s = 0; s < 9; ++s)
      maxerr = std::fmax(maxerr, std::fabs(sum[rul + 1] + T.scale[s * 5 + rul] * sum[rul]) * T.norm[s * 5 + rul]);
    e[rul] = maxerr;
  }
  r.avg = vol * sum[0];
  r.err = vol * ((5 * e[1] <= e[2] && 5 * e[2] <= e[3]) ? e[1] : 5 * std::fmax(std::fmax(e[1], e[2]), e[3]));
  r.bisect = bisect;
}

}  // namespace

extern "C" {

struct cuhre_result {
  double estimate, errorest;
  uint64_t neval, nregions;
  int32_t status;  // 0 converged, 1 maxeval / time cap
  int32_t threads;
  double seconds;
};

int
cuhre_restated(int family, int ndim, int feval, const double* gens, const double* weights,
               const double* scale, const double* norm, const int32_t* set_of_point, const double* G,
               double epsrel, double epsabs, uint64_t maxeval, int nthreads, double time_cap_s,
               cuhre_result* out)
{
  Integrand f = lookup(family, ndim);
  if (!f || !out)
    return -1;
  if (nthreads < 1)
    nthreads = omp_get_max_threads();
  Tables T{ndim, feval, gens, weights, scale, norm, G, set_of_point};
  const auto t0 = std::chrono::steady_clock::now();
  auto elapsed = [&]() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  };

  std::vector<Region> pool;
  pool.reserve(1 << 16);
  auto cmp = [&](int a, int b) { return pool[a].err < pool[b].err; };
  std::priority_queue<int, std::vector<int>, decltype(cmp)> heap(cmp);
  std::vector<double> fvals(feval);

  Region root;
  root.lo.assign(ndim, 0.);
  root.len.assign(ndim, 1.);
  sample(T, f, root, fvals, nthreads);
  pool.push_back(root);
  heap.push(0);
  double tot_avg = root.avg, tot_err = root.err;
  uint64_t neval = feval;
  int status = 1;

  while (true) {
    if (tot_err <= std::fmax(epsrel * std::fabs(tot_avg), epsabs)) {
      status = 0;
      break;
    }
    if (neval >= maxeval || (time_cap_s > 0 && elapsed() > time_cap_s))
      break;
    const int top = heap.top();
    heap.pop();
    const Region parent = pool[top];
    Region L = parent, R = parent;
    const int ax = parent.bisect;
    L.len[ax] = R.len[ax] = parent.len[ax] * .5;
    R.lo[ax] = parent.lo[ax] + L.len[ax];
    sample(T, f, L, fvals, nthreads);
    sample(T, f, R, fvals, nthreads);
    neval += 2ull * feval;
    // two-level refinement
    const double diff = std::fabs(.25 * (L.avg + R.avg - parent.avg));
    const double both = L.err + R.err;
    if (both > 0) {
      const double c = 1 + 2 * diff / both;
      L.err *= c;
      R.err *= c;
    }
    L.err += diff;
    R.err += diff;
    tot_avg += L.avg + R.avg - parent.avg;
    tot_err += L.err + R.err - parent.err;
    pool[top] = L;
    heap.push(top);
    pool.push_back(R);
    heap.push((int)pool.size() - 1);
  }
  out->estimate = tot_avg;
  out->errorest = tot_err;
  out->neval = neval;
  out->nregions = pool.size();
  out->status = status;
  out->threads = nthreads;
  out->seconds = elapsed();
  return 0;
}

}  // extern "C"
