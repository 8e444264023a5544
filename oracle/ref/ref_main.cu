// TEST INFRASTRUCTURE (oracle side).  Command line front end of the reference harness; see
// ref_pagani.cuh.  usage:
//   ref_pagani list
//   ref_pagani integrate <integrand> <epsrel> [--epsabs x] [--repeats n] [--plain]
//   ref_pagani trace     <integrand> <epsrel> [--epsabs x] [--out prefix] [--full maxregions]
#include "ref_pagani.cuh"

std::map<std::string, RefRunner>&
ref_registry()
{
  static std::map<std::string, RefRunner> r;
  return r;
}

int
main(int argc, char** argv)
{
  if (argc >= 2 && std::string(argv[1]) == "list") {
    for (auto& kv : ref_registry())
      printf("%s\n", kv.first.c_str());
    return 0;
  }
  if (argc < 4) {
    fprintf(stderr, "usage: %s integrate|trace <integrand> <epsrel> [options]\n", argv[0]);
    return 1;
  }
  RefOptions o;
  o.mode = argv[1];
  o.integrand = argv[2];
  o.epsrel = atof(argv[3]);
  for (int i = 4; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--epsabs" && i + 1 < argc)
      o.epsabs = atof(argv[++i]);
    else if (a == "--repeats" && i + 1 < argc)
      o.repeats = atoi(argv[++i]);
    else if (a == "--plain")
      o.subregions_overload = false;
    else if (a == "--out" && i + 1 < argc)
      o.out_prefix = argv[++i];
    else if (a == "--full" && i + 1 < argc)
      o.full_max = strtoull(argv[++i], nullptr, 10);
    else if (a == "--max-iters" && i + 1 < argc)
      o.max_iters = atoi(argv[++i]);
    else {
      fprintf(stderr, "unknown option %s\n", a.c_str());
      return 1;
    }
  }
  auto it = ref_registry().find(o.integrand);
  if (it == ref_registry().end()) {
    fprintf(stderr, "unknown integrand %s (try: list)\n", o.integrand.c_str());
    return 1;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    fprintf(stderr, "no CUDA device\n");
    return 4;
  }
  return it->second(o);
}
