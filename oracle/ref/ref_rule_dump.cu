// TEST INFRASTRUCTURE (oracle side) -- not part of the product.
//
// Runs the *reference's own* host-side rule construction
// (/root/reference/cuda/pagani/quad/GPUquad/Rule.cuh:370-510, Rule9Generate :24-270)
// without a GPU (its cudaMalloc/cudaMemcpy calls only return error codes, cudaDebugUtil.h:66-94)
// and writes every table raw so that tests/golden/rule_tables.npz can pin our own
// rule construction (gpuintegration_b200/csrc/rule.cc) bit-for-bit.
//
// The dense generator table is what quad::ComputeGenerators/ActualCompute
// (Phases.cuh:157-217) would leave on the device; that expansion is restated here on the host
// because the kernel itself cannot run in the GPU-less build container.
//
// Build (see oracle/Makefile): nvcc -std=c++17 -I/root/reference ... -o oracle/_ref/ref_rule_dump
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include "cuda/pagani/quad/GPUquad/Rule.cuh"

// quad::Rule keeps its host tables private; explicit template instantiation may name
// private members, which lets this tool read them without touching the reference.
template <typename Tag, typename Tag::type M>
struct Rob {
  friend typename Tag::type get(Tag) { return M; }
};
#define ROB(name, T)                                                           \
  struct Tag_##name {                                                          \
    typedef T quad::Rule<double>::*type;                                       \
    friend type get(Tag_##name);                                               \
  };                                                                           \
  template struct Rob<Tag_##name, &quad::Rule<double>::name>;
ROB(NSETS, int)
ROB(PERMUTATIONS_POS_ARRAY_SIZE, int)
ROB(cpuG, double*)
ROB(CPURuleWt, double*)
ROB(CPUScale, double*)
ROB(CPUNorm, double*)
ROB(cpuGenPermGIndex, int*)
ROB(cpuGenPermVarStart, int*)
ROB(genPtr, int*)
#define RULE(obj, name) ((obj)->*get(Tag_##name()))

static void
put(FILE* f, const void* p, size_t bytes)
{
  if (fwrite(p, 1, bytes, f) != bytes) {
    perror("fwrite");
    exit(2);
  }
}

int
main(int argc, char** argv)
{
  if (argc < 3) {
    fprintf(stderr, "usage: %s <outdir> <ndim> [<ndim> ...]\n", argv[0]);
    return 1;
  }
  std::string outdir = argv[1];
  for (int a = 2; a < argc; ++a) {
    const int ndim = atoi(argv[a]);
    const int feval = 1 + 8 * ndim + 6 * ndim * (ndim - 1) +
                      4 * ndim * (ndim - 1) * (ndim - 2) / 3 + (1 << ndim);
    Structures<double> constMem;
    quad::Rule<double>* rule = new quad::Rule<double>;
    rule->Init(ndim, feval, 0, 0, &constMem);
    cudaGetLastError();

    const int nsets = RULE(rule, NSETS);
    const int npos = RULE(rule, PERMUTATIONS_POS_ARRAY_SIZE);
    // dense expansion (value of generator coordinate `dim` for point `p`)
    std::vector<double> dense((size_t)ndim * feval, 0.0);
    for (int p = 0; p < feval; ++p) {
      const int start = RULE(rule, cpuGenPermVarStart)[p];
      const int cnt = RULE(rule, cpuGenPermVarStart)[p + 1] - start;
      const int set = RULE(rule, cpuGenPermGIndex)[p];
      for (int k = 0; k < cnt; ++k) {
        const int pos = RULE(rule, genPtr)[start + k];
        const int ap = abs(pos);
        const double v = RULE(rule, cpuG)[set * ndim + k];
        dense[(size_t)(ap - 1) * feval + p] = (pos == ap) ? v : -v;
      }
    }

    std::string fn = outdir + "/rule_ndim" + std::to_string(ndim) + ".bin";
    FILE* f = fopen(fn.c_str(), "wb");
    if (!f) {
      perror(fn.c_str());
      return 2;
    }
    int hdr[4] = {ndim, feval, nsets, npos};
    put(f, hdr, sizeof(hdr));
    put(f, RULE(rule, cpuG), sizeof(double) * ndim * nsets);
    put(f, RULE(rule, CPURuleWt), sizeof(double) * nsets * 5);
    put(f, RULE(rule, CPUScale), sizeof(double) * nsets * 5);
    put(f, RULE(rule, CPUNorm), sizeof(double) * nsets * 5);
    put(f, RULE(rule, cpuGenPermGIndex), sizeof(int) * feval);
    put(f, RULE(rule, cpuGenPermVarStart), sizeof(int) * (feval + 1));
    put(f, RULE(rule, genPtr), sizeof(int) * npos);
    put(f, dense.data(), sizeof(double) * dense.size());
    fclose(f);
    printf("ndim=%d feval=%d nsets=%d npos=%d -> %s\n", ndim, feval, nsets, npos, fn.c_str());
    // the reference under-allocates cpuGenPermVarStart (Rule.cuh:420-421); leak the
    // object rather than run its destructor on that heap block.
  }
  return 0;
}
