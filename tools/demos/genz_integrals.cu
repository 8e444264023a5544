// The reference's Genz sweep (cuda/pagani/demos/genz_integrals.cu:9-82) on the B200-native engine:
// F_1..F_6 at 8, 7, 6 and 5 dimensions, epsrel 1e-3 ... 1e-9, three runs each (first one unreported),
// CSV columns of clean_time_and_call.  The functors are include/pagani_b200/genz.cuh's (bit-identical to
// the reference's F_k_nD, tests/test_integrand_sass.py) wrapped in the reference's integrand contract
// (set_true_value / true_value).
//   genz_integrals [out.csv] [max_epsrel_exponent=9] [ndims=8,7,6,5]
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>

#include "pagani_b200/demos/time_and_call.cuh"
#include "pagani_b200/genz.cuh"

template <template <int> class G, int FAMILY, int N>
struct Wrapped : G<N> {
  double true_value = 0.;
  void set_true_value() { true_value = pgn_genz_true_value(FAMILY, N); }
};

template <int N>
void
sweep(const std::vector<double>& epsrels, std::ostream& out)
{
  using namespace pagani_b200::genz;
  constexpr bool use_custom = false;
  constexpr int debug = 0, num_runs = 3;
  constexpr bool collect = true;
  for (double epsrel : epsrels) {
    clean_time_and_call<Wrapped<Oscillatory, 1, N>, N, use_custom, debug, num_runs, collect>("f1", epsrel, out);
    clean_time_and_call<Wrapped<ProductPeak, 2, N>, N, use_custom, debug, num_runs, collect>("f2", epsrel, out);
    clean_time_and_call<Wrapped<CornerPeak, 3, N>, N, use_custom, debug, num_runs, collect>("f3", epsrel, out);
    clean_time_and_call<Wrapped<Gaussian, 4, N>, N, use_custom, debug, num_runs, collect>("f4", epsrel, out);
    clean_time_and_call<Wrapped<CZero, 5, N>, N, use_custom, debug, num_runs, collect>("f5", epsrel, out);
    clean_time_and_call<Wrapped<Discontinuous, 6, N>, N, use_custom, debug, num_runs, collect>("f6", epsrel, out);
  }
}

int
main(int argc, char** argv)
{
  const char* path = argc > 1 ? argv[1] : "pagani_genz_integrals.csv";
  const int max_exp = argc > 2 ? std::atoi(argv[2]) : 9;
  std::vector<int> dims = {8, 7, 6, 5};
  if (argc > 3) {
    dims.clear();
    std::stringstream ss(argv[3]);
    for (std::string tok; std::getline(ss, tok, ',');)
      dims.push_back(std::atoi(tok.c_str()));
  }
  std::vector<double> epsrels;
  for (int e = 3; e <= max_exp; ++e)
    epsrels.push_back(std::pow(10., -e));
  std::ofstream outfile(path);
  print_header(outfile);
  print_header();
  for (int n : dims) {
    if (n == 8) sweep<8>(epsrels, outfile);
    if (n == 7) sweep<7>(epsrels, outfile);
    if (n == 6) sweep<6>(epsrels, outfile);
    if (n == 5) sweep<5>(epsrels, outfile);
  }
  return 0;
}
