// extern "C" surface of libpagani_b200.so (declared in include/pagani_b200.h) and the
// pre-instantiated Genz evaluators.
#include <cmath>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/pagani_b200.h"
#include "../../include/pagani_b200/evaluate.cuh"
#include "../../include/pagani_b200/genz.cuh"
#include "engine.h"

struct pgn_ctx {
  pgn::Engine engine;
  pgn_ctx(int ndim, const pgn_opts& o) : engine(ndim, o) {}
};

namespace {

using namespace pagani_b200;

struct GenzEntry {
  int family, ndim;
  pgn_eval_fn fn;
};

template <typename F>
constexpr GenzEntry
entry(int family)
{
  return GenzEntry{family, F::ndim, &detail::trampoline<F, F::ndim>};
}

// what this build instantiates: the reference's 5..8-D sweep (genz_integrals.cu:9-82) and the 10-D
// C0 / discontinuous integrands of BASELINE config 4
const GenzEntry kGenz[] = {
  entry<genz::Oscillatory<5>>(1),   entry<genz::Oscillatory<6>>(1),   entry<genz::Oscillatory<7>>(1),
  entry<genz::Oscillatory<8>>(1),   entry<genz::ProductPeak<5>>(2),   entry<genz::ProductPeak<6>>(2),
  entry<genz::ProductPeak<7>>(2),   entry<genz::ProductPeak<8>>(2),   entry<genz::CornerPeak<5>>(3),
  entry<genz::CornerPeak<6>>(3),    entry<genz::CornerPeak<7>>(3),    entry<genz::CornerPeak<8>>(3),
  entry<genz::Gaussian<5>>(4),      entry<genz::Gaussian<6>>(4),      entry<genz::Gaussian<7>>(4),
  entry<genz::Gaussian<8>>(4),      entry<genz::CZero<5>>(5),         entry<genz::CZero<6>>(5),
  entry<genz::CZero<7>>(5),         entry<genz::CZero<8>>(5),         entry<genz::CZero<10>>(5),
  entry<genz::Discontinuous<5>>(6), entry<genz::Discontinuous<6>>(6), entry<genz::Discontinuous<7>>(6),
  entry<genz::Discontinuous<8>>(6), entry<genz::Discontinuous<10>>(6),
  // the reference unit tests' integrands: family 0 constant, 7 product of coordinates, 8 NaN
  entry<testfn::Constant<2>>(0),    entry<testfn::Constant<3>>(0),    entry<testfn::Constant<5>>(0),
  entry<testfn::Constant<8>>(0),    entry<testfn::Product<2>>(7),     entry<testfn::NotANumber<2>>(8),
  entry<testfn::Product<3>>(7),     entry<testfn::Product<4>>(7),
};

const GenzEntry*
find_genz(int family, int ndim)
{
  for (const GenzEntry& e : kGenz)
    if (e.family == family && e.ndim == ndim)
      return &e;
  return nullptr;
}

int
fail(int code, const std::string& msg)
{
  pgn::set_last_error(msg);
  return code;
}

}  // namespace

extern "C" {

const char*
pgn_version(void)
{
  return "pagani_b200 0.1 (sm_100a)";
}

const char*
pgn_last_error(void)
{
  return pgn::last_error();
}

int
pgn_create(pgn_ctx** out, int ndim, const pgn_opts* opts)
{
  if (!out)
    return PGN_ERR_INVALID;
  *out = nullptr;
  pgn_opts o;
  std::memset(&o, 0, sizeof o);
  o.device = -1;
  if (opts)
    o = *opts;
  if (ndim < 2 || ndim > PGN_MAX_NDIM)
    return fail(PGN_ERR_INVALID, "ndim must be in [2, 15]");
  pgn_ctx* c = new (std::nothrow) pgn_ctx(ndim, o);
  if (!c)
    return PGN_ERR_NOMEM;
  const int rc = c->engine.init();
  if (rc != PGN_OK) {
    delete c;
    return rc;
  }
  *out = c;
  return PGN_OK;
}

void
pgn_destroy(pgn_ctx* ctx)
{
  delete ctx;
}

int
pgn_device_info(pgn_ctx* ctx, int* sm_count, size_t* free_bytes, size_t* total_bytes)
{
  return ctx ? ctx->engine.device_info(sm_count, free_bytes, total_bytes) : PGN_ERR_INVALID;
}

int
pgn_set_eval_variant(pgn_ctx* ctx, int variant)
{
  if (!ctx || variant < 0 || variant > 2)
    return PGN_ERR_INVALID;
  ctx->engine.set_eval_variant(variant);
  return PGN_OK;
}

int
pgn_get_stream(pgn_ctx* ctx, void** stream)
{
  if (!ctx || !stream)
    return PGN_ERR_INVALID;
  *stream = ctx->engine.stream();
  return PGN_OK;
}

int
pgn_launch_count(pgn_ctx* ctx, uint64_t* launches)
{
  if (!ctx || !launches)
    return PGN_ERR_INVALID;
  *launches = ctx->engine.launches();
  return PGN_OK;
}

int
pgn_feval(int ndim)
{
  return (ndim < 2 || ndim > PGN_MAX_NDIM) ? PGN_ERR_INVALID : pgn::feval_of(ndim);
}

int
pgn_rule_tables_host(int ndim, double* generators, double* weights, double* scale, double* norm,
                     int32_t* set_of_point, uint64_t* point_code)
{
  pgn::RuleTables r;
  if (!pgn::build_rule(ndim, &r))
    return fail(PGN_ERR_INVALID, "ndim must be in [2, 15]");
  if (generators) {
    const std::vector<double> g = r.dense();
    std::memcpy(generators, g.data(), sizeof(double) * g.size());
  }
  if (weights)
    std::memcpy(weights, r.weights, sizeof r.weights);
  if (scale)
    std::memcpy(scale, r.scale, sizeof r.scale);
  if (norm)
    std::memcpy(norm, r.norm, sizeof r.norm);
  if (set_of_point)
    std::memcpy(set_of_point, r.set_of_point.data(), sizeof(int32_t) * r.feval);
  if (point_code)
    std::memcpy(point_code, r.point_code.data(), sizeof(uint64_t) * r.feval);
  return PGN_OK;
}

int
pgn_rule_view_get(pgn_ctx* ctx, pgn_rule_view* out)
{
  return ctx ? ctx->engine.rule_view(out) : PGN_ERR_INVALID;
}

int
pgn_regions_init_uniform(pgn_ctx* ctx, int parts_per_axis)
{
  return ctx ? ctx->engine.regions_init_uniform(parts_per_axis) : PGN_ERR_INVALID;
}

int
pgn_regions_import(pgn_ctx* ctx, const double* left, const double* length, size_t n)
{
  return ctx ? ctx->engine.regions_import(left, length, n) : PGN_ERR_INVALID;
}

int
pgn_regions_export(pgn_ctx* ctx, double* left, double* length, size_t cap, size_t* n)
{
  return ctx ? ctx->engine.regions_export(left, length, cap, n) : PGN_ERR_INVALID;
}

int
pgn_regions_count(pgn_ctx* ctx, size_t* n)
{
  if (!ctx || !n)
    return PGN_ERR_INVALID;
  *n = ctx->engine.regions_count();
  return PGN_OK;
}

int
pgn_set_volume(pgn_ctx* ctx, const double* lows, const double* highs)
{
  return ctx ? ctx->engine.set_volume(lows, highs) : PGN_ERR_INVALID;
}

int
pgn_evaluate(pgn_ctx* ctx, pgn_eval_fn fn, void* user, double* sum_estimate, double* sum_errorest)
{
  return ctx ? ctx->engine.evaluate(fn, user, sum_estimate, sum_errorest) : PGN_ERR_INVALID;
}

int
pgn_classify(pgn_ctx* ctx, double epsrel, int relerr_classification, pgn_iter_stats* out)
{
  return ctx ? ctx->engine.classify(epsrel, relerr_classification, out) : PGN_ERR_INVALID;
}

int
pgn_set_active(pgn_ctx* ctx, const uint8_t* flags, size_t n)
{
  return ctx ? ctx->engine.set_active(flags, n) : PGN_ERR_INVALID;
}

int
pgn_set_estimates(pgn_ctx* ctx, const double* estimate, const double* errorest, const int32_t* axis,
                  size_t n)
{
  return ctx ? ctx->engine.set_estimates(estimate, errorest, axis, n) : PGN_ERR_INVALID;
}

int
pgn_filter(pgn_ctx* ctx, size_t* num_active)
{
  return ctx ? ctx->engine.filter(num_active) : PGN_ERR_INVALID;
}

int
pgn_heuristic_classify(pgn_ctx* ctx, double epsrel, double last_estimate, double iter_errorest,
                       double iter_finished_errorest, double total_finished_errorest, pgn_heuristic_result* out)
{
  return ctx ? ctx->engine.heuristic_classify_stage(epsrel, last_estimate, iter_errorest, iter_finished_errorest,
                                                    total_finished_errorest, out)
             : PGN_ERR_INVALID;
}

int
pgn_split(pgn_ctx* ctx)
{
  return ctx ? ctx->engine.split() : PGN_ERR_INVALID;
}

int
pgn_dev_uniform_split(pgn_ctx* ctx, int parts_per_axis, double* d_left, double* d_length, size_t n)
{
  return ctx ? ctx->engine.dev_uniform_split(parts_per_axis, d_left, d_length, n) : PGN_ERR_INVALID;
}

int
pgn_dev_evaluate(pgn_ctx* ctx, pgn_eval_fn fn, void* user, const double* d_left, const double* d_length, size_t n,
                 double* d_estimate, double* d_errorest, int32_t* d_axis, double* d_active, double* sum_estimate,
                 double* sum_errorest)
{
  return ctx ? ctx->engine.dev_evaluate(fn, user, d_left, d_length, n, d_estimate, d_errorest, d_axis, d_active,
                                        sum_estimate, sum_errorest)
             : PGN_ERR_INVALID;
}

int
pgn_dev_count_active(pgn_ctx* ctx, const double* d_active, size_t n, size_t* count)
{
  return ctx ? ctx->engine.dev_count_active(d_active, n, count) : PGN_ERR_INVALID;
}

int
pgn_dev_filter(pgn_ctx* ctx, const double* d_left, const double* d_length, const double* d_estimate,
               const double* d_errorest, const int32_t* d_axis, const double* d_active, size_t n, size_t m,
               double* out_left, double* out_length, double* out_estimate, double* out_errorest, int32_t* out_axis)
{
  return ctx ? ctx->engine.dev_filter(d_left, d_length, d_estimate, d_errorest, d_axis, d_active, n, m, out_left,
                                      out_length, out_estimate, out_errorest, out_axis)
             : PGN_ERR_INVALID;
}

int
pgn_dev_split(pgn_ctx* ctx, const double* d_left, const double* d_length, const int32_t* d_axis, size_t m,
              double* out_left, double* out_length)
{
  return ctx ? ctx->engine.dev_split(d_left, d_length, d_axis, m, out_left, out_length) : PGN_ERR_INVALID;
}

int
pgn_regions_import_device(pgn_ctx* ctx, const double* d_left, const double* d_length, size_t n)
{
  return ctx ? ctx->engine.regions_import_device(d_left, d_length, n) : PGN_ERR_INVALID;
}

int
pgn_regions_export_device(pgn_ctx* ctx, double* d_left, double* d_length, size_t cap, size_t* n)
{
  return ctx ? ctx->engine.regions_export_device(d_left, d_length, cap, n) : PGN_ERR_INVALID;
}

int
pgn_fetch_estimates(pgn_ctx* ctx, double* estimate, double* errorest, int32_t* axis, uint8_t* active,
                    size_t cap)
{
  return ctx ? ctx->engine.fetch_estimates(estimate, errorest, axis, active, cap) : PGN_ERR_INVALID;
}

int
pgn_fetch_parents(pgn_ctx* ctx, double* estimate, double* errorest, int32_t* axis, size_t cap,
                  size_t* n)
{
  return ctx ? ctx->engine.fetch_parents(estimate, errorest, axis, cap, n) : PGN_ERR_INVALID;
}

int
pgn_integrate(pgn_ctx* ctx, pgn_eval_fn fn, void* user, const pgn_integrate_args* args,
              pgn_result_t* out)
{
  if (!ctx || !args)
    return PGN_ERR_INVALID;
  return ctx->engine.integrate(fn, user, *args, out);
}

int
pgn_genz_available(int family, int ndim)
{
  return find_genz(family, ndim) != nullptr;
}

int
pgn_genz_evaluator(int family, int ndim, pgn_eval_fn* fn, void** user)
{
  const GenzEntry* e = find_genz(family, ndim);
  if (!e || !fn)
    return fail(PGN_ERR_INVALID, "Genz family/ndim not instantiated in this build");
  *fn = e->fn;
  if (user)
    *user = nullptr;
  return PGN_OK;
}

int
pgn_integrate_genz(pgn_ctx* ctx, int family, double epsrel, double epsabs, const double* lows,
                   const double* highs, pgn_result_t* out)
{
  if (!ctx)
    return PGN_ERR_INVALID;
  const GenzEntry* e = find_genz(family, ctx->engine.ndim());
  if (!e)
    return fail(PGN_ERR_INVALID, "Genz family/ndim not instantiated in this build");
  int rc = ctx->engine.set_volume(lows, highs);
  if (rc != PGN_OK)
    return rc;
  pgn_integrate_args a;
  std::memset(&a, 0, sizeof a);
  a.epsrel = epsrel;
  a.epsabs = epsabs;
  a.relerr_classification = 1;
  a.use_current_regions = 0;
  a.iters_mode = PGN_ITERS_AFTER_SPLIT;
  return ctx->engine.integrate(e->fn, nullptr, a, out);
}

double
pgn_genz_true_value(int family, int ndim)
{
  // closed forms of demos/compute_genz_integrals.cuh:243-340 and the constants the reference
  // hard-codes where it has no closed form (integrands.cuh:1322, 1381, ...)
  const double pi = 3.14159265358979323846;
  double v = 1.0;
  switch (family) {
    case 0:
      return 15.37;
    case 7:
      return std::pow(.5, ndim);
    case 1: {
      // Re prod_k (e^{i k} - 1) / (i k)
      double re = 1.0, im = 0.0;
      for (int k = 1; k <= ndim; ++k) {
        const double a = std::sin(static_cast<double>(k)) / k;
        const double b = (1.0 - std::cos(static_cast<double>(k))) / k;
        const double nr = re * a - im * b, ni = re * b + im * a;
        re = nr;
        im = ni;
      }
      return re;
    }
    case 2:
      for (int j = 0; j < ndim; ++j)
        v = v * 50. * (std::atan((1.0 - .5) * 50.) + std::atan(.5 * 50.));
      return v;
    case 3:
      switch (ndim) {
        case 5: return 0.00002602538279621613;
        case 6: return 7.1790160638199853886e-7;
        case 7: return 1.459429e-08;
        case 8: return 2.2751965817917756076e-10;
        default: return std::nan("");
      }
    case 4:
      for (int j = 0; j < ndim; ++j)
        v = v * (std::sqrt(pi) / 25.) * .5 * (std::erf(25. * (1.0 - .5)) + std::erf(25. * .5));
      return v;
    case 5:
      for (int j = 0; j < ndim; ++j)
        v = v * (2.0 - std::exp(-5.) - std::exp(5. - 10.)) / 10.;
      return v;
    case 6: {
      double alpha[10], beta[10];
      auto fill = [&](auto spec, int n) {
        for (int i = 0; i < n; ++i) {
          alpha[i] = decltype(spec)::alpha(i);
          beta[i] = decltype(spec)::beta(i);
        }
      };
      switch (ndim) {
        case 5: fill(genz::DiscontinuousSpec<5>{}, 5); break;
        case 6: fill(genz::DiscontinuousSpec<6>{}, 6); break;
        case 7: fill(genz::DiscontinuousSpec<7>{}, 7); break;
        case 8: fill(genz::DiscontinuousSpec<8>{}, 8); break;
        case 10: fill(genz::DiscontinuousSpec<10>{}, 10); break;
        default: return std::nan("");
      }
      for (int j = 0; j < ndim; ++j)
        v = v * (std::exp(alpha[j] * beta[j]) - 1.0) / alpha[j];
      return v;
    }
    default:
      return std::nan("");
  }
}

double
pgn_genz_flops_per_eval(int family, int ndim)
{
  // 7*ndim (affine map) + 1 (Jacobian) + 10 (five rule FMAs) + integrand cost, with
  // libdevice costs cos 26, exp 31, pow 125, divide 20 (static sm_100a SASS counts)
  const double common = 7.0 * ndim + 11.0;
  switch (family) {
    case 1: return common + 2.0 * (ndim - 1) + 26.0;
    case 2: return common + ndim * (125. + 1. + 1. + 20.) + (ndim - 1);
    case 3: return common + 2.0 * ndim + 125.0;
    case 4: return common + ndim * 125. + ndim + 2.0 * ndim + 1. + 31.;
    case 5: return common + 3.0 * ndim + 31.0;
    case 6: return common + ndim + 2.0 * ndim + 31.0;
    default: return std::nan("");
  }
}

int
pgn_trace_count(pgn_ctx* ctx)
{
  return ctx ? static_cast<int>(ctx->engine.trace().size()) : PGN_ERR_INVALID;
}

int
pgn_trace_get(pgn_ctx* ctx, int i, pgn_iter_stats* out)
{
  if (!ctx || !out || i < 0 || i >= static_cast<int>(ctx->engine.trace().size()))
    return PGN_ERR_INVALID;
  *out = ctx->engine.trace()[i];
  return PGN_OK;
}

uint64_t
pgn_digest64(const void* data, size_t elem_bytes, size_t n)
{
  const unsigned char* p = static_cast<const unsigned char*>(data);
  uint64_t acc = 0;
  for (size_t i = 0; i < n; ++i) {
    uint64_t bits = 0;
    std::memcpy(&bits, p + i * elem_bytes, elem_bytes < 8 ? elem_bytes : 8);
    acc += bits * (2ull * i + 1ull);
  }
  return acc;
}

int
pgn_nccl_unique_id(uint8_t id[PGN_NCCL_ID_BYTES])
{
  std::string err;
  const int rc = pgn::nccl_unique_id(id, &err);
  if (rc != PGN_OK)
    pgn::set_last_error("NCCL: " + err);
  return rc;
}

int
pgn_comm_init(pgn_ctx* ctx, const uint8_t id[PGN_NCCL_ID_BYTES], int nranks, int rank)
{
  return ctx ? ctx->engine.comm_init_nccl(id, nranks, rank) : PGN_ERR_INVALID;
}

int
pgn_comm_size(pgn_ctx* ctx, int* nranks, int* rank)
{
  return ctx ? ctx->engine.comm_size(nranks, rank) : PGN_ERR_INVALID;
}

int
pgn_set_rebalance(pgn_ctx* ctx, double max_over_mean, size_t min_regions)
{
  if (!ctx || !(max_over_mean >= 1.0))
    return PGN_ERR_INVALID;
  ctx->engine.set_rebalance(max_over_mean, min_regions);
  return PGN_OK;
}

int
pgn_comm_set_host_collective(pgn_ctx* ctx, int nranks, int rank, pgn_allreduce_fn allreduce,
                             pgn_alltoallv_fn alltoallv, void* user)
{
  return ctx ? ctx->engine.comm_init_host(nranks, rank, allreduce, alltoallv, user) : PGN_ERR_INVALID;
}

int
pgn_plan_rebalance(const uint64_t* counts, int nranks, uint64_t* new_counts, uint64_t* transfers,
                   int max_transfers, int* ntransfers)
{
  if (!counts || nranks < 1 || !new_counts || !ntransfers)
    return PGN_ERR_INVALID;
  std::vector<size_t> c(counts, counts + nranks), nc;
  std::vector<pgn::Transfer> plan;
  pgn::plan_rebalance(c, &nc, &plan);
  for (int r = 0; r < nranks; ++r)
    new_counts[r] = nc[r];
  *ntransfers = static_cast<int>(plan.size());
  if (static_cast<int>(plan.size()) > max_transfers)
    return transfers ? PGN_ERR_INVALID : PGN_OK;
  for (size_t k = 0; k < plan.size() && transfers; ++k) {
    transfers[5 * k + 0] = plan[k].src;
    transfers[5 * k + 1] = plan[k].dst;
    transfers[5 * k + 2] = plan[k].src_offset;
    transfers[5 * k + 3] = plan[k].dst_offset;
    transfers[5 * k + 4] = plan[k].count;
  }
  return PGN_OK;
}

}  // extern "C"
