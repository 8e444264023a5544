// Kernel-level benchmark of the evaluation kernel through the drop-in path: the evaluator is compiled
// HERE with the functor (as a user of include/pagani_b200/Workspace.cuh would), everything else comes
// from libpagani_b200.so.  Protocol of the reference's kernel-only profile
// (cuda/pagani/demos/new_time_and_call.cuh:30-70): uniform grid of parts^ndim regions, repeated.
//   nvcc ... -DBENCH_FAMILY=Oscillatory -DBENCH_NDIM=8 [-DPGN_EVAL_MIN_BLOCKS=8] eval_bench.cu -lpagani_b200
//   ./eval_bench parts reps [variant]
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "pagani_b200.h"
#include "pagani_b200/evaluate.cuh"
#include "pagani_b200/genz.cuh"

#ifndef BENCH_FAMILY
#define BENCH_FAMILY Oscillatory
#endif
#ifndef BENCH_NDIM
#define BENCH_NDIM 8
#endif
#define STR2(x) #x
#define STR(x) STR2(x)

int
main(int argc, char** argv)
{
  const int parts = argc > 1 ? atoi(argv[1]) : 5;
  const int reps = argc > 2 ? atoi(argv[2]) : 5;
  const int variant = argc > 3 ? atoi(argv[3]) : 0;
#ifdef BENCH_TESTFN
  using F = pagani_b200::testfn::BENCH_FAMILY<BENCH_NDIM>;
#else
  using F = pagani_b200::genz::BENCH_FAMILY<BENCH_NDIM>;
#endif
  pgn_ctx* ctx = nullptr;
  if (pgn_create(&ctx, BENCH_NDIM, nullptr) != PGN_OK) {
    fprintf(stderr, "pgn_create: %s\n", pgn_last_error());
    return 1;
  }
  pgn_set_eval_variant(ctx, variant);
  F f;
  pagani_b200::Evaluator<F, BENCH_NDIM> ev(f);
  void* sv = nullptr;
  pgn_get_stream(ctx, &sv);
  cudaStream_t stream = static_cast<cudaStream_t>(sv);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double est = 0, err = 0, best = 1e30;
  size_t n = 0;
  for (int r = 0; r < reps + 1; ++r) {
    if (pgn_regions_init_uniform(ctx, parts) != PGN_OK) {
      fprintf(stderr, "init: %s\n", pgn_last_error());
      return 1;
    }
    pgn_regions_count(ctx, &n);
    cudaEventRecord(e0, stream);
    if (pgn_evaluate(ctx, ev.fn(), ev.user(), &est, &err) != PGN_OK) {
      fprintf(stderr, "evaluate: %s\n", pgn_last_error());
      return 1;
    }
    cudaEventRecord(e1, stream);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best)
      best = ms;
  }
  const double evals = static_cast<double>(n) * pgn_feval(BENCH_NDIM);
  printf("{\"functor\": \"%s<%d>\", \"min_blocks\": %d, \"variant\": %d, \"regions\": %zu, \"ms\": %.4f, "
         "\"evals_per_s\": %.4e, \"estimate\": %.17g, \"errorest\": %.17g}\n",
         STR(BENCH_FAMILY), BENCH_NDIM, PGN_EVAL_MIN_BLOCKS, variant, n, best, evals / (best * 1e-3), est, err);
  pgn_destroy(ctx);
  return 0;
}
