// See rule.h.  Compiled with -ffp-contract=off: every product/sum below must round exactly once,
// in the order written, to reproduce the reference's host arithmetic (Rule.cuh:36-147, 256-269).
#include "rule.h"

#include <algorithm>
#include <cmath>

namespace pgn {
namespace {

// The five generator abscissae of the rule (CUBA rule 9; Rule.cuh:201-205).
constexpr double kLambda[5] = {.4779536579022695061928604197171830064732,
                               .2030285873691198677998034402373279133258,
                               .4476273546261781288207704806530998539285,
                               .125,
                               .3430378987808781457001426145164678603407};

// Polynomial coefficient sets of the dimension-dependent weights.
constexpr double kA[4] = {-.002361170967785511788400941242259231309691,
                          .1141539002385732526821323741697655347686,
                          -.6383392007670238909386026193674701393074,
                          .7484998850468520800423030047583803945205};
constexpr double kB[4] = {-.001432401703339912514196154599769007103671,
                          .05747150786448972594860897296200006759892,
                          -.1422510457143424323449521620935950679394,
                          -.06287502873828697998942424881040490136987};
constexpr double kF[3] = {.003541756451678267682601411863388846964536,
                          -.07260936739589367960492815865074633743652,
                          .1055749162521899101218622863269817454540};
constexpr double kG[3] = {.002148602555009868771294231899653510655506,
                          -.03226856389295394999786630399875134318006,
                          .01063678399023121748083624225818915724455};

// volatile-free helpers; -ffp-contract=off keeps each operation separately rounded.
inline double
cubic_head(double n, const double* c)  // n*(n*(n*c0 + c1) + c2)
{
  double t = n * c[0] + c[1];
  t = n * t + c[2];
  return n * t;
}
inline double
quad_head(double n, const double* c)  // n*(n*c0 + c1)
{
  const double t = n * c[0] + c[1];
  return n * t;
}

void
fill_weights(int ndim, double* w)
{
  const double n = static_cast<double>(ndim);
  const double two_n = static_cast<double>(1 << ndim);
  auto W = [&](int set, int rule) -> double& { return w[set * kRules + rule]; };

  // set 0: the centre
  const double pa = cubic_head(n, kA) + kA[3];
  const double pb = cubic_head(n, kB) - kB[3];
  W(0, 0) = pa;
  W(0, 1) = pb;
  W(0, 2) = n * .2545911332489590890011611142429070613156 - pb;
  {
    const double d[2] = {-1.207328566678236261002219995185143356737,
                         .8956736576416067650809467826488567200939};
    double t = quad_head(n, d) - 1.0;
    t = t + cubic_head(n, kA);
    W(0, 3) = t + kA[3];
  }
  W(0, 4) = (n * -.3647935698604914666100134551377381205297 + 1.0) - pa;

  // set 1: (l1, 0, ...)
  const double qf = quad_head(n, kF) + kF[2];
  const double qg = quad_head(n, kG) + kG[2];
  W(1, 0) = qf;
  W(1, 1) = qg;
  W(1, 2) = .01468910249614349017540783437728097691502 - qg;
  {
    double t = n * .5113470834646759143109387357149329909126 +
               .4597644812080634464633352781605214342691;
    t = t + quad_head(n, kF);
    W(1, 3) = t + kF[2];
  }
  W(1, 4) = .1823967849302457333050067275688690602649 - qf;

  // set 2: (l2, 0, ...)
  const double l1 = n * -.04508628929435784075980562738240804429658 +
                    .2141588352435279340097929526588394300172;
  const double l2 = n * -.02735154652654564472203690086290223507436 +
                    .05494106704871123410060080562462135546101;
  W(2, 0) = l1;
  W(2, 1) = l2;
  W(2, 2) = .1193759620257077529708962121565290178730 - l2;
  W(2, 3) = n * .6508951939192025059314756320878023215278 +
            .1474493982943446016775696826942585013243;
  W(2, 4) = -l1;

  // set 3: (l3, 0, ...)
  W(3, 0) = .05769338449097348357291272840392627722165;
  W(3, 1) = .03499962660214358382244159694487155861542;
  W(3, 2) = -.05769338449097348357291272840392627722165;
  W(3, 3) = -1.386862771927828143599782668709014266770;
  W(3, 4) = -.05769338449097348357291272840392627722165;

  // set 4: (l4, 0, ...)
  W(4, 0) = 0.0;
  W(4, 1) = 0.0;
  W(4, 2) = -.2386668732575008878964134721962088068396;
  W(4, 3) = 0.0;
  W(4, 4) = 0.0;

  // set 5: (l1, l1, 0, ...)
  const double m1 = .01553241727660705326386197156586357005224 - n * kF[0];
  const double m2 = .003532809960709087023561817517751309380604 - n * kG[0];
  W(5, 0) = m1;
  W(5, 1) = m2;
  W(5, 2) = -m2;
  W(5, 3) = (.09231719987444221619017126187763868745587 +
             .01553241727660705326386197156586357005224) -
            n * kF[0];
  W(5, 4) = -m1;

  // set 6: (l1, l2, 0, ...)
  W(6, 0) = .02254314464717892037990281369120402214829;
  W(6, 1) = .01367577326327282236101845043145111753718;
  W(6, 2) = -.01367577326327282236101845043145111753718;
  W(6, 3) = -.3254475969596012529657378160439011607639;
  W(6, 4) = -.02254314464717892037990281369120402214829;

  // set 7: (l1, l1, l1, 0, ...)
  W(7, 0) = .001770878225839133841300705931694423482268;
  W(7, 1) = .001074301277504934385647115949826755327753;
  W(7, 2) = -.001074301277504934385647115949826755327753;
  W(7, 3) = .001770878225839133841300705931694423482268;
  W(7, 4) = -.001770878225839133841300705931694423482268;

  // set 8: the 2^n corners (l5, ..., l5)
  const double p = .2515001149531479199576969952416196054795 / two_n;
  const double q = -.06287502873828697998942424881040490136987 / two_n;
  W(8, 0) = p;
  W(8, 1) = q;
  W(8, 2) = -q;
  W(8, 3) = p;
  W(8, 4) = -p;
}

// value pattern (as lambda indices, -1 = zero) of each set's generator, in axis order
struct SetShape {
  int nnz;      // non-zero coordinates (ndim for the corner set)
  int lam[3];   // lambda index of the first three non-zero coordinates
};
constexpr SetShape kShape[kSets] = {{0, {-1, -1, -1}}, {1, {0, -1, -1}}, {1, {1, -1, -1}},
                                    {1, {2, -1, -1}},  {1, {3, -1, -1}}, {2, {0, 0, -1}},
                                    {2, {0, 1, -1}},   {3, {0, 0, 0}},   {-1, {4, 4, 4}}};

inline uint64_t
code_of(int lam, bool negative)
{
  return lam < 0 ? 0u : static_cast<uint64_t>(1 + 2 * lam + (negative ? 1 : 0));
}

}  // namespace

bool
build_rule(int ndim, RuleTables* out)
{
  if (ndim < 2 || ndim > kMaxDim || out == nullptr)
    return false;
  RuleTables& r = *out;
  r.ndim = ndim;
  r.feval = feval_of(ndim);

  fill_weights(ndim, r.weights);

  const int count[kSets] = {1,
                            2 * ndim,
                            2 * ndim,
                            2 * ndim,
                            2 * ndim,
                            2 * ndim * (ndim - 1),
                            4 * ndim * (ndim - 1),
                            4 * ndim * (ndim - 1) * (ndim - 2) / 3,
                            1 << ndim};
  for (int s = 0; s < kSets; ++s)
    r.set_count[s] = count[s];

  // null-rule scaling (Rule.cuh:256-269): for every set and rules 1..3
  for (int i = 0; i < kSets * kRules; ++i)
    r.scale[i] = r.norm[i] = 0.0;
  for (int s = 0; s < kSets; ++s) {
    const double* ws = &r.weights[s * kRules];
    for (int k = 1; k < kRules - 1; ++k) {
      const double sc = (ws[k] == 0) ? 100.0 : -ws[k + 1] / ws[k];
      double sum = 0.0;
      for (int x = 0; x < kSets; ++x) {
        const double* wx = &r.weights[x * kRules];
        sum += static_cast<double>(count[x]) * std::fabs(wx[k + 1] + sc * wx[k]);
      }
      r.scale[s * kRules + k] = sc;
      r.norm[s * kRules + k] = 1.0 / sum;
    }
  }

  // coordinate codes: 0 -> 0, 1+2i -> +lambda_i, 2+2i -> -lambda_i
  r.gen_value[0] = 0.0;
  for (int i = 0; i < 5; ++i) {
    r.gen_value[1 + 2 * i] = kLambda[i];
    r.gen_value[2 + 2 * i] = -kLambda[i];
  }
  for (int c = 0; c < kCodes; ++c) {
    r.half_plus_g[c] = .5 + r.gen_value[c];
    r.half_minus_g[c] = .5 - r.gen_value[c];
  }
  {
    const double q = kLambda[1] / kLambda[0];
    r.ratio = q * q;
  }

  // Point order (Rule.cuh:438-505): sets in order; inside a set every distinct placement of the
  // generator's non-zero values over the axes, visited so that the axis-reversed vector runs through
  // its permutations in increasing lexicographic order; inside a placement all sign patterns, the
  // lowest axis toggling fastest.
  r.set_of_point.assign(r.feval, 0);
  r.point_code.assign(r.feval, 0);
  int p = 0;
  std::vector<int> rev(ndim);  // rev[j] = lambda index (+1; 0 = zero) on axis ndim-1-j
  for (int s = 0; s < kSets; ++s) {
    if (count[s] == 0)
      continue;
    const int nnz = kShape[s].nnz < 0 ? ndim : kShape[s].nnz;
    // rank lambdas by value so that std::next_permutation orders placements numerically
    auto rank_of = [&](int lam) {
      int rk = 1;
      for (int j = 0; j < 5; ++j)
        if (kLambda[j] < kLambda[lam])
          ++rk;
      return rk;
    };
    int lam_of_rank[8] = {-1, -1, -1, -1, -1, -1, -1, -1};
    std::fill(rev.begin(), rev.end(), 0);
    for (int k = 0; k < nnz; ++k) {
      const int lam = kShape[s].lam[k < 3 ? k : 2];
      const int rk = rank_of(lam);
      lam_of_rank[rk] = lam;
      rev[ndim - 1 - k] = rk;  // axis k holds the k-th generator value
    }
    int emitted = 0;
    do {
      // axes carrying a non-zero value, in axis order
      int axes[kMaxDim];
      int na = 0;
      for (int d = 0; d < ndim; ++d)
        if (rev[ndim - 1 - d] != 0)
          axes[na++] = d;
      for (uint32_t signs = 0; signs < (1u << na); ++signs) {
        uint64_t code = static_cast<uint64_t>(s) << 60;
        for (int k = 0; k < na; ++k) {
          const int d = axes[k];
          const int lam = lam_of_rank[rev[ndim - 1 - d]];
          code |= code_of(lam, (signs >> k) & 1u) << (4 * d);
        }
        r.set_of_point[p] = s;
        r.point_code[p] = code;
        ++p;
        ++emitted;
      }
    } while (std::next_permutation(rev.begin(), rev.end()));
    if (emitted != count[s])
      return false;
  }
  return p == r.feval;
}

std::vector<double>
RuleTables::dense() const
{
  std::vector<double> g(static_cast<size_t>(ndim) * feval, 0.0);
  for (int p = 0; p < feval; ++p)
    for (int d = 0; d < ndim; ++d)
      g[static_cast<size_t>(d) * feval + p] = gen_value[(point_code[p] >> (4 * d)) & 0xF];
  return g;
}

}  // namespace pgn
