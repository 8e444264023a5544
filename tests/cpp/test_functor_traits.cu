// Host-only checks of the evaluation kernel's per-functor protocol (no GPU needed; built and run by
// tests/test_functor_traits.py): which optional members the Genz functors offer, what the launch-time range promise
// answers for a few integration volumes, and the tuning traits the launcher reads.
#include <cmath>
#include <cstdio>
#include <limits>

#include "pagani_b200/evaluate.cuh"
#include "pagani_b200/genz.cuh"

using namespace pagani_b200;
namespace d = pagani_b200::detail;

static_assert(d::has_batch_fast<genz::Oscillatory<8>, 8, 2>::value, "F_1 offers batch_fast");
static_assert(d::has_batch_nocheck<genz::Oscillatory<8>, 8, 2>::value, "F_1 offers the range promise");
static_assert(d::has_batch_nocheck<genz::Oscillatory<5>, 5, 1>::value, "... for one point per lane as well");
static_assert(!d::has_batch_nocheck<genz::Gaussian<8>, 8, 2>::value, "F_4: x - .5 == 0 is data, not a volume property");
static_assert(!d::has_batch_nocheck<genz::CZero<10>, 10, 2>::value, "F_5: measured slower, not offered");
static_assert(!d::has_batch_nocheck<genz::Discontinuous<10>, 10, 2>::value, "F_6 has no batch form at all");
static_assert(eval_unit_jacobian_variant<genz::Oscillatory<8>>::value, "F_1 unit-cube kernel");
static_assert(eval_unit_jacobian_variant<genz::CZero<10>>::value, "F_5 unit-cube kernel");
static_assert(!eval_unit_jacobian_variant<genz::Gaussian<8>>::value, "F_4: none");
static_assert(eval_uniform_warp<genz::Oscillatory<8>>::value, "warp-uniform index: default on");
static_assert(eval_uniform_warp<genz::Gaussian<5>>::value && !eval_uniform_warp<genz::Gaussian<8>>::value,
              "F_4 opts out at 8-D");
static_assert(EvalTuning<genz::Oscillatory<8>>::kMinBlocks == 5, "launch bound of the oscillatory family");

template <int N>
static bool
promise(double lo, double hi)
{
  double l[N], h[N];
  for (int k = 0; k < N; ++k) {
    l[k] = lo;
    h[k] = hi;
  }
  return genz::Oscillatory<N>{}.fast_in_box(l, h);
}

int
main()
{
  int bad = 0;
  auto expect = [&](bool got, bool want, const char* what) {
    if (got != want) {
      std::printf("FAIL %s: got %d\n", what, int(got));
      ++bad;
    }
  };
  expect(promise<8>(0., 1.), true, "unit cube");
  expect(promise<8>(100., 101.), true, "shifted unit cube");
  expect(promise<8>(-1e6, 1e6), true, "|s| <= 3.6e7");
  expect(promise<8>(0., 1e9), false, "|s| up to 3.6e10 (Payne-Hanek range)");
  expect(promise<8>(-3e7, 0.), false, "36 * 3e7 = 1.08e9 >= 2^30, by the lower bounds");
  expect(promise<8>(0., 2.9e7), true, "36 * 2.9e7 = 1.044e9 < 2^30");
  expect(promise<5>(0., std::numeric_limits<double>::quiet_NaN()), false, "NaN bound");
  expect(promise<5>(0., std::numeric_limits<double>::infinity()), false, "infinite bound");
  std::printf(bad ? "%d failures\n" : "ok\n", bad);
  return bad ? 1 : 0;
}
