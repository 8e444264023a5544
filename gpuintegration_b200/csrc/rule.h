// Host-side construction of the degree-9 Genz-Malik / CUBA "rule 9" tables PAGANI evaluates.
// Restates quad::Rule<T>::Init + Rule9Generate (reference cuda/pagani/quad/GPUquad/Rule.cuh:24-270,
// 370-510) and the dense expansion of quad::ComputeGenerators (Phases.cuh:157-217); pinned
// bit-for-bit against tests/golden/rule_tables.npz (made from the reference's own host code).
#ifndef PGN_RULE_H
#define PGN_RULE_H

#include <cstdint>
#include <vector>

namespace pgn {

constexpr int kSets = 9;
constexpr int kRules = 5;
constexpr int kCodes = 11;
constexpr int kMaxDim = 15;

// points per region: 1 + 4*(2n) + 2n(n-1) + 4n(n-1) + 4n(n-1)(n-2)/3 + 2^n   (quad.h:68-75)
inline int
feval_of(int n)
{
  return 1 + 8 * n + 6 * n * (n - 1) + 4 * n * (n - 1) * (n - 2) / 3 + (1 << n);
}

struct RuleTables {
  int ndim = 0;
  int feval = 0;
  double weights[kSets * kRules];  // cRuleWt[set*5 + rule]
  double scale[kSets * kRules];    // GPUScale[set*5 + rule], rules 1..3
  double norm[kSets * kRules];     // GPUNorm[set*5 + rule],  rules 1..3
  double gen_value[kCodes];        // signed generator coordinate of each 4-bit code
  double half_plus_g[kCodes];      // 0.5 + g
  double half_minus_g[kCodes];     // 0.5 - g
  double ratio;                    // (lambda2 / lambda1)^2, Sample.cuh:195-196
  int set_count[kSets];            // points in each generator set
  std::vector<int32_t> set_of_point;   // [feval]   gpuGenPermGIndex
  std::vector<uint64_t> point_code;    // [feval]   4 bits per axis + set in bits 60..63
  // dense [ndim][feval] generator matrix, only materialised on request (golden comparison)
  std::vector<double> dense() const;
};

// Throws nothing; returns false for ndim outside [2, kMaxDim].
bool build_rule(int ndim, RuleTables* out);

}  // namespace pgn
#endif
