// CPU ORACLE (test infrastructure, see kc_oracle.h) of the Coffee tree search: a restatement of the reference's
// single-threaded playout loop for one game,
//   Search::playoutDescend                         cpp/search/search.cpp:935-1160
//   Search::selectBestChildToDescend               cpp/search/searchexplorehelpers.cpp:323-451
//   getExploreScaling / getExploreSelectionValue   cpp/search/searchexplorehelpers.cpp:9-45
//   getFpuValueForChildrenAssumeVisited            cpp/search/searchexplorehelpers.cpp:248-320
//   addLeafValue / recomputeNodeStats              cpp/search/searchupdatehelpers.cpp:12-76, 151-326
// under SearchParams() defaults (cpp/search/searchparams.cpp:8-90) with valueWeightExponent = 0, cpuctExplorationLog = 0,
// no graph search, no noise, one thread: every visit has weight 1, so a node's utilityAvg is the mean of the leaf
// utilities below it (kept as a sum W and a count N), a child's weight is its visit count.  PARITY UNPINNED: the
// reference's search does not compile against Coffee positions and has no Coffee test; this file is the
// definition the CUDA search (katacoffee_b200/csrc/search.cu) is compared with, bit for bit.
//
// Compiled with -ffp-contract=off: all arithmetic is plain IEEE double, the same operations in the same order as the
// device code (which uses explicit round-to-nearest intrinsics).  The only order-sensitive sum (the policy mass of the
// visited children) is accumulated exactly as a 32-lane warp does it: lane = index mod 32, then an xor butterfly.
#include <cmath>
#include <cstring>
#include <vector>

#include "kc_oracle.h"

namespace {

constexpr uint64_t PHI = 0x9E3779B97F4A7C15ULL;
constexpr uint64_t CHOOSE_SALT = 0xC0FFEE5EA4C4ULL;

struct Node {
  int N = 0, numChildren = 0, nextPla = 0;
  double W = 0.0;
  std::vector<double> edgeW;
  std::vector<float> policy;
  std::vector<int> child;     // -1 none, >= 0 node, -2 draw, -3 black win, -4 white win
  std::vector<int> edgeN;
  std::vector<uint8_t> order;
  explicit Node(int P) : edgeW(P, 0.0), policy(P, -1.0f), child(P, -1), edgeN(P, 0), order(P, 0) {}
};

double terminalValue(int winner) { return winner == 2 ? 1.0 : winner == 1 ? -1.0 : 0.0; }

double butterfly(double part[32]) {
  double cur[32], nxt[32];
  memcpy(cur, part, sizeof(cur));
  for(int o = 16; o > 0; o >>= 1) {
    for(int l = 0; l < 32; l++) nxt[l] = cur[l] + cur[l ^ o];
    memcpy(cur, nxt, sizeof(cur));
  }
  return cur[0];
}

constexpr uint64_t SYM_SALT = 0x5A11E7C0FFEEULL;
constexpr uint64_t ROOTSYM_SALT = 0x7007575E5A17ULL;
struct Evaluator {
  const ko_model* model;   // null: integer-hash evaluator
  int W, H, P, LW;
  bool randomSym = false;  // nnRandomize (nneval.cpp:515-524): symmetry drawn from the position's sit-hash and the seed
  uint64_t seed = 0;
  // fills policy [P] (-1 illegal), (whiteWin, whiteLoss) and, if asked, shorttermWinlossError for the position in g (player to move =
  // next player).  forcedSym >= 0: evaluate under exactly that symmetry (the root's rootNumSymmetriesToSample evaluations); the
  // integer-hash evaluator then mixes the symmetry into its hash, so a wrong symmetry choice shows up as a different result.
  void eval(const ko_game* g, float* policy, float winLoss[2], int forcedSym = -1, float* shortErr = nullptr) const {
    std::vector<uint32_t> legal(LW);
    const int pla = ko_game_next_pla(g);
    ko_game_legal_mask(g, pla, legal.data());
    if(!model) {
      uint64_t h[2];
      ko_game_sit_hash(g, pla, h);
      if(forcedSym >= 0) { h[0] ^= ko_splitmix64(ROOTSYM_SALT + (uint64_t)forcedSym); h[1] ^= ko_splitmix64(ROOTSYM_SALT * 3 + (uint64_t)forcedSym); }
      int sum = 0;
      for(int pos = 0; pos < P; pos++)
        if((legal[pos >> 5] >> (pos & 31)) & 1u) sum += 1 + (int)((ko_splitmix64(h[0] ^ ((uint64_t)(pos + 1) * PHI)) >> 20) & 255);
      for(int pos = 0; pos < P; pos++) {
        const bool ok = (legal[pos >> 5] >> (pos & 31)) & 1u;
        const int w = 1 + (int)((ko_splitmix64(h[0] ^ ((uint64_t)(pos + 1) * PHI)) >> 20) & 255);
        policy[pos] = ok ? (float)w / (float)sum : -1.0f;
      }
      const uint64_t r = ko_splitmix64(h[1]);
      winLoss[0] = (float)(r & 0xFFFF) * (1.0f / 131072.0f);
      winLoss[1] = (float)((r >> 16) & 0xFFFF) * (1.0f / 131072.0f);
      if(shortErr) *shortErr = (float)((r >> 32) & 0xFFFF) * (1.0f / 131072.0f);   // [0, 0.5)
      return;
    }
    std::vector<float> planes((size_t)15 * W * H), own((size_t)W * H);
    float glob = 0.f, value[2], misc[2];
    ko_game_fill_row_v1(g, pla, W, H, 0, planes.data(), &glob);
    int8_t sym = 0;
    if(forcedSym >= 0) sym = (int8_t)forcedSym;
    else if(randomSym) {
      uint64_t h[2];
      ko_game_sit_hash(g, pla, h);
      sym = (int8_t)(ko_splitmix64(seed ^ h[0] ^ SYM_SALT) & 7);
    }
    ko_model_forward(model, 1, W, H, 0, planes.data(), &glob, (randomSym || forcedSym >= 0) ? &sym : nullptr, policy, value, misc, own.data(), 0, 1);
    ko_postprocess(policy, P, legal.data(), 1.0f, value, misc, pla);
    winLoss[0] = value[0]; winLoss[1] = value[1];
    if(shortErr) *shortErr = misc[1];
  }
};

}  // namespace

extern "C" {

static void searchRunOnTree(std::vector<Node>& nodes, const ko_game* rootGame, int x_size, int y_size, const ko_search_params* p,
                            const ko_model* modelOrNull, int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits,
                            double* edgeUtilitySum, float* policyOut, uint8_t* orderOut, uint64_t counters[3]) {
  const int P = 4 * x_size * y_size;
  Evaluator ev{modelOrNull, x_size, y_size, P, (P + 31) / 32};
  ev.randomSym = p->nnRandomize != 0; ev.seed = p->noiseSeed;
  nodes.reserve(p->maxVisits);
  uint64_t cVisits = 0, cEvals = 0, cTerminal = 0;
  ko_game* g = ko_game_create(x_size, y_size, 4);
  std::vector<float> pol(P);
  float wl[2];
  for(int it = 0; it < p->maxVisits && !ko_game_finished(rootGame); it++) {
    ko_game_copy(g, rootGame);
    std::vector<std::pair<int, int>> path;   // (node, pos)
    int kind = 0;
    double v = 0.0;
    if(nodes.empty()) kind = 4;
    else if(nodes[0].N >= p->maxVisits) break;
    else {
      int node = 0, depth = 0;
      while(true) {
        Node& nd = nodes[node];
        const int pla = nd.nextPla;
        const double parentUtility = nd.W / (double)nd.N;
        double partT[32] = {0}, partM[32] = {0};
        for(int pos = 0; pos < P; pos++)
          if(nd.child[pos] != -1) { partT[pos & 31] = partT[pos & 31] + (double)nd.edgeN[pos]; partM[pos & 31] = partM[pos & 31] + (double)nd.policy[pos]; }
        const double total = butterfly(partT), mass = butterfly(partM);
        const double red = (depth == 0 ? p->rootFpuReductionMax : p->fpuReductionMax) * std::sqrt(mass);
        const double fpu = pla == 2 ? parentUtility - red : parentUtility + red;
        const double scale = p->cpuctExploration * std::sqrt(total + 0.01);
        double bestVal = 0.0; int bestOrd = 1 << 20, bestPos = -1;
        float newP = -1.0f; int newPos = -1;
        for(int pos = 0; pos < P; pos++) {
          const float pr = nd.policy[pos];
          if(nd.child[pos] != -1) {
            const double n = (double)nd.edgeN[pos];
            const double q = nd.edgeW[pos] / n;
            const double val = (scale * (double)pr) / (1.0 + n) + (pla == 2 ? q : -q);
            const int o = nd.order[pos];
            if(bestPos < 0 || val > bestVal || (val == bestVal && o < bestOrd)) { bestVal = val; bestOrd = o; bestPos = pos; }
          } else if(pr >= 0.0f) {
            if(pr > newP) { newP = pr; newPos = pos; }
          }
        }
        bool takeNew = false;
        if(newPos >= 0) {
          const double valNew = (scale * (double)newP) / 1.0 + (pla == 2 ? fpu : -fpu);
          takeNew = bestPos < 0 || valNew > bestVal;
        }
        const int pos = takeNew ? newPos : bestPos;
        if(pos < 0) { kind = 0; break; }
        path.push_back({node, pos});
        depth++;
        if(!takeNew) {
          const int cc = nd.child[pos];
          if(cc <= -2) { kind = 3; v = terminalValue(-2 - cc); break; }
          ko_game_play(g, pos);
          node = cc;
          continue;
        }
        ko_game_play(g, pos);
        if(ko_game_finished(g)) { kind = 2; v = terminalValue(ko_game_winner(g)); }
        else kind = 1;
        break;
      }
    }
    if(kind == 0) continue;
    int newIdx = -1;
    if(kind == 1 || kind == 4) {
      ev.eval(g, pol.data(), wl);
      v = (double)wl[0] - (double)wl[1];
      newIdx = (int)nodes.size();
      nodes.emplace_back(P);
      Node& nn = nodes.back();
      nn.N = 1; nn.W = v; nn.nextPla = ko_game_next_pla(g);
      for(int pos = 0; pos < P; pos++) nn.policy[pos] = pol[pos];
    }
    for(size_t d = 0; d < path.size(); d++) {
      Node& nd = nodes[path[d].first];
      const int pos = path[d].second;
      if(d + 1 == path.size() && (kind == 1 || kind == 2)) {
        nd.child[pos] = kind == 1 ? newIdx : (v > 0.0 ? -4 : v < 0.0 ? -3 : -2);
        nd.order[pos] = (uint8_t)nd.numChildren;
        nd.numChildren++;
      }
      nd.N += 1;
      nd.W = nd.W + v;
      nd.edgeN[pos] += 1;
      nd.edgeW[pos] = nd.edgeW[pos] + v;
    }
    cVisits++;
    if(kind == 1 || kind == 4) cEvals++; else cTerminal++;
  }
  ko_game_destroy(g);
  const bool have = !nodes.empty();
  if(rootVisits) *rootVisits = have ? nodes[0].N : 0;
  if(rootUtilitySum) *rootUtilitySum = have ? nodes[0].W : 0.0;
  for(int pos = 0; pos < P; pos++) {
    const bool ex = have && nodes[0].child[pos] != -1;
    if(edgeVisits) edgeVisits[pos] = ex ? nodes[0].edgeN[pos] : 0;
    if(edgeUtilitySum) edgeUtilitySum[pos] = ex ? nodes[0].edgeW[pos] : 0.0;
    if(policyOut) policyOut[pos] = have ? nodes[0].policy[pos] : 0.f;
    if(orderOut) orderOut[pos] = ex ? nodes[0].order[pos] : 255;
  }
  if(counters) { counters[0] += cVisits; counters[1] += cEvals; counters[2] += cTerminal; }
}

void ko_search_run(const ko_game* rootGame, int x_size, int y_size, const ko_search_params* p, const ko_model* modelOrNull,
                   int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits, double* edgeUtilitySum, float* policyOut,
                   uint8_t* orderOut, uint64_t counters[3]) {
  std::vector<Node> nodes;
  searchRunOnTree(nodes, rootGame, x_size, y_size, p, modelOrNull, rootVisits, rootUtilitySum, edgeVisits, edgeUtilitySum, policyOut, orderOut, counters);
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------------------------
// Graph search + subtree value bias (BASELINE config 4: useGraphSearch, subtreeValueBiasFactor).
// Restates, for one game and one thread:
//   Search::allocateOrFindNode (transposition lookup)        cpp/search/search.cpp:704-757
//   Search::playoutDescend / maybeCatchUpEdgeVisits          cpp/search/search.cpp:935-1207
//   selection on child NODE statistics, getChildWeight       cpp/search/searchexplorehelpers.cpp:92-127, 323-451, searchnode.h:59-65
//   addLeafValue / recomputeNodeStats incl. the bias table   cpp/search/searchupdatehelpers.cpp:12-76, 151-326
//   SubtreeValueBiasTable::get                               cpp/search/subtreevaluebiastable.cpp:61-78
// Canonical choices (ledger rows L and M):
//  * transposition key = getSitHash(next player) mixed with the last move (cell + direction): the state that decides
//    legality.  The literal GraphHash chains the previous hash after every non-pass move (graphhash.cpp:14-29) and would
//    never transpose in a game without passes.  Finished positions are not shared (terminal children stay per edge;
//    equivalent, their statistics are constants).
//  * a node keeps (visits, weightSum, utilityAvg); an edge keeps its visit count; a child's weight seen from a parent is
//    weightSum * edgeVisits / max(visits, 1).  valueWeightExponent 0, no noise pruning, unit evaluation weights.
//  * bias entry key = (player who moved, previous move, move, colours of the 5x5 window around the move on the board
//    before it); Go's atari and ko terms do not exist.  The root has no entry; the table is per search.
//  * origTotalChildWeight^exponent: sqrt for 0.5, identity for 1, otherwise detPow (below) -- a pow built from IEEE
//    basic operations only, so that the device computes the same bits.
//  * sums over children are accumulated per lane (policy index mod 32) and combined by an xor butterfly, like the warp does.
// PARITY UNPINNED, like the rest of the search.
#include <map>
namespace {

double detLog(double x) {
  uint64_t b; memcpy(&b, &x, 8);
  int k = (int)((b >> 52) & 0x7ff) - 1022;                       // x = m * 2^k, m in [0.5, 1)
  b = (b & 0x800fffffffffffffULL) | 0x3fe0000000000000ULL;
  double m; memcpy(&m, &b, 8);
  if(m < 0.70710678118654752) { m = m * 2.0; k -= 1; }
  const double z = (m - 1.0) / (m + 1.0), z2 = z * z;
  double s = 1.0 / 27.0;
  for(int n = 25; n >= 1; n -= 2) s = s * z2 + 1.0 / (double)n;
  return (double)k * 0.69314718055994531 + (2.0 * z) * s;
}
double detExp(double y) {
  if(y < -700.0) return 0.0;   // below the normal range of the 2^n scaling
  const double n = std::nearbyint(y * 1.4426950408889634);
  const double r = (y - n * 0.693147180369123816490) - n * 1.90821492927058770002e-10;
  double s = 1.0 / 6227020800.0;                                 // 1/13!
  static const double inv[13] = {1.0, 1.0, 1.0 / 2.0, 1.0 / 6.0, 1.0 / 24.0, 1.0 / 120.0, 1.0 / 720.0, 1.0 / 5040.0, 1.0 / 40320.0,
                                 1.0 / 362880.0, 1.0 / 3628800.0, 1.0 / 39916800.0, 1.0 / 479001600.0};
  for(int i = 12; i >= 0; i--) s = s * r + inv[i];
  const uint64_t eb = (uint64_t)((int)n + 1023) << 52;           // 2^n, n well inside the normal range here
  double sc; memcpy(&sc, &eb, 8);
  return s * sc;
}
double biasPow(double x, double e) {
  if(e == 0.5) return std::sqrt(x);
  if(e == 1.0) return x;
  return detExp(e * detLog(x));
}

typedef std::pair<uint64_t, uint64_t> Key;
// Root policy temperature and shaped Dirichlet noise (Search::maybeAddPolicyNoiseAndTemp / addDirichletNoise /
// computeDirichletAlphaDistribution, cpp/search/searchhelpers.cpp:51-221; Rand::nextGamma / nextGaussian / nextDouble,
// cpp/core/rand.cpp:335-363, rand.h:235-288) with the two things that cannot be shared with a GPU replaced: the generator is
// a counter-based splitmix64 stream keyed by (seed, game id, ply), and log / exp / pow are detLog / detExp.
constexpr uint64_t NOISE_SALT = 0xD1A1C4137E5EEDULL;
struct DetRng {
  uint64_t s; bool hasG = false; double g = 0.0;
  DetRng(uint64_t seed, uint64_t gameId, int ply) : s(ko_splitmix64(seed ^ (gameId * PHI) ^ (uint64_t)ply ^ NOISE_SALT)) {}
  uint64_t nextU64() { s += PHI; return ko_splitmix64(s); }
  double nextDouble() { return (double)(nextU64() & ((1ULL << 53) - 1ULL)) * (1.0 / 9007199254740992.0); }
  double nextGaussian() {
    if(hasG) { hasG = false; return g; }
    double v1, v2, q;
    do {
      v1 = nextDouble() * 2.0 - 1.0;
      v2 = nextDouble() * 2.0 - 1.0;
      q = v1 * v1 + v2 * v2;
    } while(q >= 1.0 || q == 0.0);
    const double mult = std::sqrt((-2.0 * detLog(q)) / q);
    g = v2 * mult; hasG = true;
    return v1 * mult;
  }
  double nextGamma(double a) {
    if(a <= 1.0) {
      const double r = nextGamma(a + 1.0);
      const double inva = 1.0 / a;
      const double u = nextDouble();
      const double scale = u == 0.0 ? 0.0 : detExp(inva * detLog(u));
      return r * scale;
    }
    const double d = a - 1.0 / 3.0;
    const double c = (1.0 / 3.0) / std::sqrt(d);
    while(true) {
      const double x = nextGaussian();
      const double vtmp = 1.0 + c * x;
      if(vtmp <= 0.0) continue;
      const double v = (vtmp * vtmp) * vtmp;
      const double u = nextDouble();
      const double xx = x * x;
      if(u < 1.0 - (0.0331 * xx) * xx) return d * v;
      if(u == 0.0 || detLog(u) < 0.5 * xx + d * ((1.0 - v) + detLog(v))) return d * v;
    }
  }
};

// in place on the root's policy (illegal = negative entries stay as they are)
void rootPolicyNoiseAndTemp(const ko_search_params* p, int P, int boardArea, int turn, uint64_t seed, uint64_t gameId, float* pol) {
  const double tEarly = p->rootPolicyTemperatureEarly > 0.0 ? p->rootPolicyTemperatureEarly : 1.0;
  const double tLate = p->rootPolicyTemperature > 0.0 ? p->rootPolicyTemperature : 1.0;
  if(tEarly != 1.0 || tLate != 1.0) {
    const double halflife = p->chosenMoveTemperatureHalflife > 0.0 ? p->chosenMoveTemperatureHalflife : 19.0;
    const double halflives = (((double)turn / halflife) * 19.0) / std::sqrt((double)boardArea);     // interpolateEarly :463-467
    const double T = tLate + (tEarly - tLate) * detExp(halflives * detLog(0.5));
    double maxValue = 0.0;
    for(int i = 0; i < P; i++) if((double)pol[i] > maxValue) maxValue = (double)pol[i];
    if(maxValue > 0.0) {
      const double logMax = detLog(maxValue), invTemp = 1.0 / T;
      double sum = 0.0;
      for(int i = 0; i < P; i++)
        if(pol[i] > 0) { const float q = (float)detExp((detLog((double)pol[i]) - logMax) * invTemp); pol[i] = q; sum += (double)q; }
      for(int i = 0; i < P; i++) if(pol[i] >= 0) pol[i] = (float)((double)pol[i] / sum);
    }
  }
  if(p->rootNoiseEnabled) {
    std::vector<double> r(P, 0.0);
    int legalCount = 0;
    for(int i = 0; i < P; i++) if(pol[i] >= 0) legalCount++;
    if(legalCount == 0) return;
    double logSum = 0.0;
    for(int i = 0; i < P; i++) if(pol[i] >= 0) { r[i] = detLog(std::min(0.01, (double)pol[i]) + 1e-20); logSum += r[i]; }
    const double logMean = logSum / (double)legalCount;
    double alphaPropSum = 0.0;
    for(int i = 0; i < P; i++) if(pol[i] >= 0) { r[i] = std::max(0.0, r[i] - logMean); alphaPropSum += r[i]; }
    const double uniformProb = 1.0 / (double)legalCount;
    for(int i = 0; i < P; i++)
      if(pol[i] >= 0) r[i] = alphaPropSum <= 0.0 ? uniformProb : 0.5 * (r[i] / alphaPropSum + uniformProb);
    DetRng rng(seed, gameId, turn);
    double rSum = 0.0;
    for(int i = 0; i < P; i++) {
      if(pol[i] >= 0) { r[i] = rng.nextGamma(r[i] * p->rootDirichletNoiseTotalConcentration); rSum += r[i]; }
      else r[i] = 0.0;
    }
    const double w = p->rootDirichletNoiseWeight;
    for(int i = 0; i < P; i++)
      if(pol[i] >= 0) pol[i] = (float)((r[i] / rSum) * w + (double)pol[i] * (1.0 - w));
  }
}
bool wantsRootPolicyChange(const ko_search_params* p) {
  return p->rootNoiseEnabled || (p->rootPolicyTemperature > 0.0 && p->rootPolicyTemperature != 1.0) ||
         (p->rootPolicyTemperatureEarly > 0.0 && p->rootPolicyTemperatureEarly != 1.0);
}

// valueWeightExponent (downweightBadChildrenAndNormalizeWeight, cpp/search/searchupdatehelpers.cpp:330-417): a child whose
// utility lies z standard errors below the weighted mean of its siblings keeps weight * cdf(z)^exponent, cdf = Student t with
// 3 degrees of freedom tabulated at 2,000 points on [-50, 50] and interpolated (DistributionTable, search.cpp:111-116).
// Canonical: the table is filled from the closed form of the df = 3 cdf (the reference goes through an incomplete beta
// function); p^exponent is sqrt / sqrt(sqrt) for 0.5 / 0.25, detPow otherwise.
struct TCdfTable {
  double cdf[2000];
  TCdfTable() {
    const double PI = 3.14159265358979323846, s3 = std::sqrt(3.0);
    for(int i = 0; i < 2000; i++) {
      if(i == 0) cdf[i] = 0.0;
      else if(i == 1999) cdf[i] = 1.0;
      else {
        const double z = -50.0 + (double)i * 100.0 / 1999.0, x = z / s3;
        cdf[i] = 0.5 + (x / (1.0 + x * x) + std::atan(x)) / PI;
      }
    }
  }
  double get(double z) const {   // DistributionTable::getCdf
    const double d = (1999.0 * (z - -50.0)) / 100.0;
    if(d <= 0) return 0.0;
    const int idx = (int)d;
    if(idx >= 1999) return 1.0;
    const double lambda = d - (double)idx;
    return cdf[idx] + lambda * (cdf[idx + 1] - cdf[idx]);
  }
};
const TCdfTable& tcdf() { static TCdfTable t; return t; }
double valueWeightPow(double x, double e) {
  if(e == 0.5) return std::sqrt(x);
  if(e == 0.25) return std::sqrt(std::sqrt(x));
  if(e == 1.0) return x;
  return detExp(e * detLog(x));
}

thread_local std::vector<double> g_lastPlaySelection;

struct GNode {
  bool noised = false;
  int visits = 0, numChildren = 0, nextPla = 0, biasEntry = -1, depth = 0;
  double weightSum = 0.0, utilityAvg = 0.0, nnUtility = 0.0, lastDelta = 0.0, lastWeight = 0.0;
  double utilitySqAvg = 0.0, weightSqSum = 0.0;   // NodeStats::utilitySqAvg / weightSqSum (searchnode.h:44-48): what LCB reads
  double nnWeight = 1.0;                          // computeWeightFromNNOutput (searchupdatehelpers.cpp:91-113): 1 without useUncertainty
  Key key = Key(0, 0);
  std::vector<float> policy;
  std::vector<int> child, edgeN;
  std::vector<uint8_t> order;
  explicit GNode(int P) : policy(P, -1.0f), child(P, -1), edgeN(P, 0), order(P, 0) {}
};
struct BiasEntry { double deltaSum = 0.0, weightSum = 0.0; Key key = Key(0, 0); };

struct GraphSearch {
  int W, H, P;
  const ko_search_params* p;
  std::vector<GNode> nodes;
  std::map<Key, int> table;       // transposition key -> node
  std::map<Key, int> biasIndex;   // bias key -> entry
  std::vector<BiasEntry> bias;
  uint64_t seed = 0, gameId = 0;  // key the root noise stream

  void maybeNoiseRoot(const ko_game* rootGame) {
    if(nodes.empty() || nodes[0].noised) return;
    nodes[0].noised = true;
    if(wantsRootPolicyChange(p)) rootPolicyNoiseAndTemp(p, P, W * H, ko_game_num_turns(rootGame), seed, gameId, nodes[0].policy.data());
  }

  static Key stateKey(const ko_game* g) {
    uint64_t h[2];
    ko_game_sit_hash(g, ko_game_next_pla(g), h);
    const uint64_t lm = (uint64_t)(ko_game_recent_move_pos(g, 0) + 1);
    return Key(h[0] ^ ko_splitmix64(lm), h[1] ^ ko_splitmix64(lm * PHI));
  }
  // key of the bias entry of the node reached by `movePos` from `before` (the position before the move)
  Key biasKey(const ko_game* before, int movePos) const {
    const int HW = W * H, cell = movePos % HW, cx = cell % W, cy = cell / W;
    uint64_t win = 0;
    for(int dy = -2; dy <= 2; dy++)
      for(int dx = -2; dx <= 2; dx++) {
        const int x = cx + dx, y = cy + dy;
        const uint64_t code = (x < 0 || y < 0 || x >= W || y >= H) ? 3 : (uint64_t)ko_game_color_at(before, x, y);
        win |= code << (2 * ((dy + 2) * 5 + (dx + 2)));
      }
    const uint64_t mover = (uint64_t)ko_game_next_pla(before);
    return Key(win | (mover << 50) | ((uint64_t)movePos << 52), (uint64_t)(ko_game_recent_move_pos(before, 0) + 1));
  }
  void childStats(const GNode& nd, int pos, int& cv, double& cw, double& cu) const {
    const int c = nd.child[pos];
    if(c >= 0) { cv = nodes[c].visits; cw = nodes[c].weightSum; cu = nodes[c].utilityAvg; }
    else { cv = nd.edgeN[pos]; cw = (double)nd.edgeN[pos]; cu = terminalValue(-2 - c); }
  }
  // the same with utilitySqAvg and weightSqSum; a terminal child is a node that got addLeafValue(result, weight 1) on every visit
  void childStatsSq(const GNode& nd, int pos, int& cv, double& cw, double& cu, double& cusq, double& cwsq) const {
    const int c = nd.child[pos];
    if(c >= 0) { cv = nodes[c].visits; cw = nodes[c].weightSum; cu = nodes[c].utilityAvg; cusq = nodes[c].utilitySqAvg; cwsq = nodes[c].weightSqSum; }
    else { cv = nd.edgeN[pos]; cw = (double)nd.edgeN[pos]; cu = terminalValue(-2 - c); cusq = cu * cu; cwsq = (double)nd.edgeN[pos]; }
  }
  // computeWeightFromNNOutput (searchupdatehelpers.cpp:91-113) with Coffee's outputs: no score term
  double nnWeightOf(float shorttermWinlossError) const {
    if(!p->useUncertainty) return 1.0;
    const double unc = (double)shorttermWinlossError;   // winLossUtilityFactor 1
    const double powered = p->uncertaintyExponent == 1.0 ? unc : p->uncertaintyExponent == 0.5 ? std::sqrt(unc) : (unc <= 0.0 ? 0.0 : detExp(p->uncertaintyExponent * detLog(unc)));
    const double baseline = p->uncertaintyCoeff / p->uncertaintyMaxWeight;
    return p->uncertaintyCoeff / (powered + baseline);
  }
  static double childWeight(double cw, int e, int cv) { return cw * ((double)e / (double)std::max(cv, 1)); }
  // the node's children in creation order (policy indices): the order the reference's children array has.  Order-sensitive sums
  // run over it the way a warp does: lane = creation index mod 32, then the xor butterfly.
  std::vector<int> kids(const GNode& nd) const {
    std::vector<int> k(nd.numChildren, -1);
    for(int pos = 0; pos < P; pos++) if(nd.child[pos] != -1) k[nd.order[pos]] = pos;
    return k;
  }

  // recomputeNodeStats (searchupdatehelpers.cpp:151-326).  Canonical order of the order-dependent sums: lane = policy index mod 32,
  // then the xor butterfly (what a warp does); the reference walks the children in creation order.
  void recompute(GNode& nd, int inc = 1, bool isRoot = false) {
    double partW[32] = {0}, partWU[32] = {0};
    double maxW = 0.0;
    const std::vector<int> ks = kids(nd);
    const int nk = (int)ks.size();
    std::vector<double> w(nk, 0.0);
    for(int k = 0; k < nk; k++) {
      const int pos = ks[k];
      int cv; double cw, cu;
      childStats(nd, pos, cv, cw, cu);
      const int e = nd.edgeN[pos];
      if(cv <= 0 || cw <= 0.0 || e <= 0) continue;
      w[k] = childWeight(cw, e, cv);
      if(w[k] > maxW) maxW = w[k];
      partW[k & 31] = partW[k & 31] + w[k];
      partWU[k & 31] = partWU[k & 31] + w[k] * cu;
    }
    const double origW = butterfly(partW);   // origTotalChildWeight: what the subtree value bias weighs with
    double sumW = origW;
    double sumWU = butterfly(partWU);
    if(p->useNoisePruning) {
      // pruneNoiseWeight (searchupdatehelpers.cpp:422-470), children in creation order: a child whose utility lies below the weighted
      // average of the children before it keeps at most twice its raw-policy share of their weight, the excess shrinking with the gap
      int good = 0;
      for(int k = 0; k < nk; k++) good += w[k] != 0.0;
      if(good > 1 && sumW > 0.00001) {
        double uSum = 0.0, wSum = 0.0, pSum = 0.0;
        for(int k = 0; k < nk; k++) {
          if(w[k] == 0.0) continue;
          int cv; double cw, cu;
          childStats(nd, ks[k], cv, cw, cu);
          const double utility = nd.nextPla == 2 ? cu : -cu;
          const double rawPolicy = std::max(1e-30, (double)nd.policy[ks[k]]);
          double nwk = w[k];
          if(wSum > 0 && pSum > 0) {
            const double gap = uSum / wSum - utility;
            if(gap > 0) {
              const double lenient = 2.0 * ((wSum * rawPolicy) / pSum);
              if(w[k] > lenient) {
                double toSub = (w[k] - lenient) * (1.0 - detExp(-(gap / p->noisePruneUtilityScale)));
                if(toSub > p->noisePruningCap) toSub = p->noisePruningCap;
                nwk = w[k] - toSub;
              }
            }
          }
          w[k] = nwk;
          uSum = uSum + utility * nwk; wSum = wSum + nwk; pSum = pSum + rawPolicy;
        }
        sumW = wSum;
        double partP[32] = {0};
        for(int k = 0; k < nk; k++) if(w[k] != 0.0) { int cv; double cw, cu; childStats(nd, ks[k], cv, cw, cu); partP[k & 31] = partP[k & 31] + w[k] * cu; }
        sumWU = butterfly(partP);
      }
    }
    // at a noised root the children the move choice would prune / reduce lose the same weight here (:196-206)
    double amountToSubtract = 0.0, amountToPrune = 0.0;
    if(isRoot && p->rootNoiseEnabled && !p->useNoisePruning) {
      amountToSubtract = std::min(p->chosenMoveSubtract, maxW / 64.0);
      amountToPrune = std::min(p->chosenMovePrune, maxW / 64.0);
    }
    const bool reweigh = sumW > 0.0 && (p->valueWeightExponent != 0.0 || amountToSubtract > 0.0 || amountToPrune > 0.0);
    std::vector<double> nw(w);
    if(reweigh) {
      // downweightBadChildrenAndNormalizeWeight (:330-417): prune / subtract, then (valueWeightExponent) a child keeps
      // weight * cdf(z)^exponent, z = its utility's distance from the siblings' weighted mean in standard errors; the total stays sumW
      const double simpleValue = sumWU / sumW;   // selfUtility is +-utility, handled through the sign below
      double partN[32] = {0};
      for(int k = 0; k < nk; k++) {
        if(w[k] == 0.0) continue;
        const int pos = ks[k];
        double x = w[k];
        if(x < amountToPrune) x = 0.0;
        else { x = x - amountToSubtract; if(x <= 0.0) x = 0.0; }
        if(x > 0.0 && p->valueWeightExponent != 0.0) {
          int cv; double cw, cu;
          childStats(nd, pos, cv, cw, cu);
          const double stdev = std::sqrt(0.00000001 + 1.0 / (1.5 * std::sqrt(w[k])));
          const double diff = nd.nextPla == 2 ? cu - simpleValue : simpleValue - cu;   // selfUtility - simpleValue (own view)
          const double pr = tcdf().get(diff / stdev) + 0.0001;
          x = x * valueWeightPow(pr, p->valueWeightExponent);
        }
        nw[k] = x;
        partN[k & 31] = partN[k & 31] + x;
      }
      const double factor = sumW / butterfly(partN);
      for(int k = 0; k < nk; k++) nw[k] = nw[k] * factor;
    }
    double partU[32] = {0}, partUSq[32] = {0}, partWSq[32] = {0};
    for(int k = 0; k < nk; k++)
      if(nw[k] != 0.0) {
        int cv; double cw, cu, cusq, cwsq;
        childStatsSq(nd, ks[k], cv, cw, cu, cusq, cwsq);
        const double scaling = nw[k] / cw;
        partU[k & 31] = partU[k & 31] + nw[k] * cu;
        partUSq[k & 31] = partUSq[k & 31] + nw[k] * cusq;
        partWSq[k & 31] = partWSq[k & 31] + (scaling * scaling) * cwsq;
      }
    if(reweigh) sumWU = butterfly(partU);
    const double sumWUSq = butterfly(partUSq), sumWSq = butterfly(partWSq);
    double utility = nd.nnUtility;
    if(p->subtreeValueBiasFactor != 0.0 && nd.biasEntry >= 0) {
      BiasEntry& E = bias[nd.biasEntry];
      if(sumW > 1e-10) {
        const double uc = sumWU / sumW;
        const double bw = biasPow(origW, p->subtreeValueBiasWeightExponent);
        const double ds = (uc - nd.nnUtility) * bw;
        E.deltaSum = E.deltaSum + (ds - nd.lastDelta);
        E.weightSum = E.weightSum + (bw - nd.lastWeight);
        nd.lastDelta = ds; nd.lastWeight = bw;
      }
      if(E.weightSum > 0.001) utility = utility + (p->subtreeValueBiasFactor * E.deltaSum) / E.weightSum;
    }
    const double w0 = nd.nnWeight;
    const double weightSum = sumW + w0;
    nd.utilityAvg = (sumWU + utility * w0) / weightSum;
    nd.utilitySqAvg = (sumWUSq + (utility * utility) * w0) / weightSum;
    nd.weightSqSum = sumWSq + w0 * w0;
    nd.weightSum = weightSum;
    nd.visits += inc;
  }

  // Search::getPlaySelectionValues at the root (cpp/search/searchresults.cpp:66-231): the value of a move is its child's weight; the
  // children other than the most stably explored one are cut down to the weight the final explore-selection value of that one
  // would have asked for (getReducedPlaySelectionWeight, searchexplorehelpers.cpp:209-243, rounded up); with useLcbForSelection
  // the child with the best lower confidence bound (getSelfUtilityLCBAndRadius, searchhelpers.cpp:469-522) is raised above every
  // child it beats.  Canonical: children are walked in policy-index order with creation order breaking ties (the reference walks
  // creation order with strict comparisons -- the same choice); cpuctUtilityStdevScale 0, no pass, no score utility.
  void playSelectionValues(double* psv) const {
    const GNode& nd = nodes[0];
    std::vector<double> lcb(P, 0.0), radius(P, 0.0);
    double total = 0.0;
    int n = 0;
    for(int pos = 0; pos < P; pos++) {
      psv[pos] = 0.0;
      if(nd.child[pos] == -1) continue;
      int cv; double cw, cu;
      childStats(nd, pos, cv, cw, cu);
      psv[pos] = childWeight(cw, nd.edgeN[pos], cv);
      total = total + psv[pos];
      n++;
    }
    if(n == 0) return;
    int best = -1, bestOrd = 1 << 20;
    double bestWeight = -1e30, maxGoodness = -1e30;
    for(int pos = 0; pos < P; pos++) {
      if(nd.child[pos] == -1) continue;
      const double e = (double)nd.edgeN[pos];
      const double g = (psv[pos] * std::max(0.0, e - 1.0)) / std::max(1.0, e) + 2.0 * (double)nd.policy[pos];
      if(g > maxGoodness || (g == maxGoodness && nd.order[pos] < bestOrd)) { maxGoodness = g; bestWeight = psv[pos]; best = pos; bestOrd = nd.order[pos]; }
    }
    const int pla = nd.nextPla;
    {
      const double scaling = p->cpuctExploration * std::sqrt(total + 0.01);
      int cv; double cw, cu;
      childStats(nd, best, cv, cw, cu);
      const double bestValue = (scaling * (double)nd.policy[best]) / (1.0 + psv[best]) + (pla == 2 ? cu : -cu);
      for(int pos = 0; pos < P; pos++) {
        if(nd.child[pos] == -1 || pos == best) continue;
        childStats(nd, pos, cv, cw, cu);
        double reduced = 0.0;
        if(cv > 0 && psv[pos] > 0.0) {
          double wanted = 0.0;
          if(nd.policy[pos] >= 0) {
            const double exploreComponent = bestValue - (pla == 2 ? cu : -cu);
            if(exploreComponent <= 0) wanted = 1e100;
            else { wanted = (scaling * (double)nd.policy[pos]) / exploreComponent - 1.0; if(wanted < 0) wanted = 0.0; }
          }
          reduced = psv[pos] > wanted ? wanted : psv[pos];
        }
        psv[pos] = std::ceil(reduced);
      }
    }
    if(!p->useLcbForSelection) return;
    double bestLcb = -1e10;
    int bestLcbPos = -1, bestLcbOrd = 1 << 20;
    for(int pos = 0; pos < P; pos++) {
      if(nd.child[pos] == -1) continue;
      int cv; double cw, cu, cusq, cwsq;
      childStatsSq(nd, pos, cv, cw, cu, cusq, cwsq);
      const double ratio = (double)nd.edgeN[pos] / (double)std::max(cv, 1);
      double weightSum = cw * ratio, weightSqSum = cwsq * ratio;
      radius[pos] = (2.0 * 1.0) * p->lcbStdevs;   // utilityRangeRadius = winLossUtilityFactor = 1
      lcb[pos] = -radius[pos];
      if(!(cv <= 0 || weightSum <= 0.0 || weightSqSum <= 0.0)) {
        double ess = (weightSum * weightSum) / weightSqSum;
        const double priorWeight = weightSum / ((ess * ess) * ess);
        double usq = std::max(cusq, cu * cu + 1e-8);
        usq = (usq * weightSum + (usq + 1.0) * priorWeight) / (weightSum + priorWeight);
        weightSum = weightSum + priorWeight;
        weightSqSum = weightSqSum + priorWeight * priorWeight;
        ess = (weightSum * weightSum) / weightSqSum;
        const double selfUtility = pla == 2 ? cu : -cu;
        const double variance = usq - cu * cu;
        const double r = std::sqrt(variance / ess) * p->lcbStdevs;
        lcb[pos] = selfUtility - r;
        radius[pos] = r;
      }
      if(psv[pos] > 0 && psv[pos] >= p->minVisitPropForLCB * bestWeight)
        if(lcb[pos] > bestLcb || (lcb[pos] == bestLcb && nd.order[pos] < bestLcbOrd)) { bestLcb = lcb[pos]; bestLcbPos = pos; bestLcbOrd = nd.order[pos]; }
    }
    // useNonBuggyLcb false: the historical bug that never promotes the first-created child
    if(bestLcbPos < 0 || (!p->useNonBuggyLcb && nd.order[bestLcbPos] == 0)) return;
    double adjusted = psv[bestLcbPos];
    for(int pos = 0; pos < P; pos++) {
      if(nd.child[pos] == -1 || pos == bestLcbPos) continue;
      const double excess = bestLcb - lcb[pos];
      if(excess < 0) continue;
      const double factor = (radius[pos] + excess) / (radius[pos] + 0.20 * excess);
      const double lbound = (factor * factor) * psv[pos];
      if(lbound > adjusted) adjusted = lbound;
    }
    psv[bestLcbPos] = adjusted;
  }

  // Search::makeMove with tree re-use in graph mode (search.cpp:262-331) followed by the next beginSearch's
  // recursivelyRecomputeStats (search.cpp:670-690, 834-910).  Canonical order of the order-dependent steps:
  //  1. the subgraph reachable from the played child is kept, renumbered breadth first (children in policy-index order); the new
  //     root is a COPY of the child without bias entry (SearchNode copy constructor, searchnode.cpp:149-188) and is not in the table;
  //  2. every other node -- the old copy of the child included -- is deleted in old-index order and gives
  //     subtreeValueBiasFreeProp of its last contribution back (removeSubtreeValueBias, search.cpp:773-786);
  //  3. bias entries no kept node refers to are erased (clearUnusedSynchronous);
  //  4. if the bias is on, every kept node is re-computed children first (deepest position first, ties by new index) without
  //     adding a visit; a node without children gets its plain evaluation back (search.cpp:877-897).
  void advance(int movePos) {
    int c0 = -1;
    if(!nodes.empty() && movePos >= 0) c0 = nodes[0].child[movePos];
    if(c0 < 0) { nodes.clear(); table.clear(); biasIndex.clear(); bias.clear(); return; }
    std::vector<int> remap(nodes.size(), -1), queue{c0};
    remap[c0] = 0;
    std::vector<GNode> out;
    for(size_t i = 0; i < queue.size(); i++) {
      GNode n = nodes[queue[i]];
      for(int pos = 0; pos < P; pos++)
        if(n.child[pos] >= 0) {
          if(remap[n.child[pos]] < 0) { remap[n.child[pos]] = (int)queue.size(); queue.push_back(n.child[pos]); }
          n.child[pos] = remap[n.child[pos]];
        }
      out.push_back(std::move(n));
    }
    for(size_t i = 0; i < nodes.size(); i++)
      if((remap[i] < 0 || (int)i == c0) && nodes[i].biasEntry >= 0) {
        BiasEntry& E = bias[nodes[i].biasEntry];
        E.deltaSum = E.deltaSum - nodes[i].lastDelta * p->subtreeValueBiasFreeProp;
        E.weightSum = E.weightSum - nodes[i].lastWeight * p->subtreeValueBiasFreeProp;
      }
    out[0].biasEntry = -1; out[0].lastDelta = 0.0; out[0].lastWeight = 0.0; out[0].noised = false;
    std::vector<BiasEntry> nbias;
    std::map<Key, int> nindex, ntable;
    for(size_t i = 1; i < out.size(); i++) {
      ntable[out[i].key] = (int)i;
      if(out[i].biasEntry >= 0) {
        const BiasEntry& E = bias[out[i].biasEntry];
        auto f = nindex.find(E.key);
        if(f == nindex.end()) { f = nindex.insert({E.key, (int)nbias.size()}).first; nbias.push_back(E); }
        out[i].biasEntry = f->second;
      }
    }
    nodes.swap(out); table.swap(ntable); biasIndex.swap(nindex); bias.swap(nbias);
    if(p->subtreeValueBiasFactor != 0.0) {
      int lo = 1 << 20, hi = -1;
      for(const GNode& n : nodes) { lo = std::min(lo, n.depth); hi = std::max(hi, n.depth); }
      for(int d = hi; d >= lo; d--)
        for(GNode& n : nodes)
          if(n.depth == d) {
            if(n.numChildren == 0) { n.utilityAvg = n.nnUtility; n.utilitySqAvg = n.nnUtility * n.nnUtility; }
            else recompute(n, 0, &n == &nodes[0]);
          }
    }
  }
};

}  // namespace

extern "C" {

// One graph search from rootGame (no tree re-use).  Outputs as ko_search_run, with rootUtilitySum = utilityAvg * visits and
// edgeUtilitySum[pos] = child's utilityAvg * edgeVisits (one rounding each); counters += visits, evaluations, terminal visits,
// transposition hits (new edge to an existing node), catch-up visits (edge behind its child: no descent); digest (may be NULL)
// receives a hash over every node of the graph in creation order.
static uint64_t graphDigest(const GraphSearch& S) {
  const int P = S.P;
  uint64_t h = 0;
  for(size_t i = 0; i < S.nodes.size(); i++) {
    const GNode& nd = S.nodes[i];
    uint64_t wb, ub, qb, sb;
    memcpy(&wb, &nd.weightSum, 8); memcpy(&ub, &nd.utilityAvg, 8); memcpy(&qb, &nd.utilitySqAvg, 8); memcpy(&sb, &nd.weightSqSum, 8);
    uint64_t nh = ko_splitmix64((uint64_t)nd.visits ^ ((uint64_t)nd.numChildren << 32)) ^ ko_splitmix64(wb ^ PHI) ^ ko_splitmix64(ub + PHI) ^
                  ko_splitmix64(qb ^ (PHI << 1)) ^ ko_splitmix64(sb + (PHI << 1));
    for(int pos = 0; pos < P; pos++)
      if(nd.child[pos] != -1)
        nh ^= ko_splitmix64(((uint64_t)(uint32_t)nd.child[pos] << 32 | (uint64_t)(uint32_t)nd.edgeN[pos]) + (uint64_t)(pos + 1) * PHI + nd.order[pos]);
    h ^= ko_splitmix64(nh + (uint64_t)(i + 1) * PHI);
  }
  return h;
}

static void graphRun(GraphSearch& S, const ko_game* rootGame, int x_size, int y_size, const ko_search_params* p, const ko_model* modelOrNull,
                     int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits, double* edgeUtilitySum, float* policyOut,
                     uint8_t* orderOut, uint64_t counters[5], uint64_t* digest) {
  const int P = 4 * x_size * y_size;
  Evaluator ev{modelOrNull, x_size, y_size, P, (P + 31) / 32};
  ev.randomSym = p->nnRandomize != 0; ev.seed = S.seed;
  S.p = p;
  const size_t maxNodes = (size_t)p->maxVisits + (p->reuseTree ? (size_t)p->maxVisits / 4 : 0);   // the device's node pool
  S.nodes.reserve(maxNodes);
  uint64_t cnt[5] = {0, 0, 0, 0, 0};
  ko_game* g = ko_game_create(x_size, y_size, 4);
  ko_game* before = ko_game_create(x_size, y_size, 4);
  std::vector<float> pol(P);
  float wl[2];
  float shortErr = 0.f;
  if(!ko_game_finished(rootGame) && p->rootNumSymmetriesToSample > 1 && (S.nodes.empty() || !S.nodes[0].noised)) {
    // initNodeNNOutput at the root (searchnnhelpers.cpp:67-83; for a root kept by tree re-use maybeRecomputeExistingNNOutput :133-174):
    // the root is evaluated under rootNumSymmetriesToSample distinct symmetries -- a partial Fisher-Yates shuffle of 0..7 -- and
    // the outputs are averaged (NNOutput's averaging constructor, nninputs.cpp:95-170: float sums in order, then / n).  Canonical:
    // the draws come from a counter stream keyed by (seed, game id, ply) instead of the search thread's generator.
    const int N = std::min(8, p->rootNumSymmetriesToSample);
    int idx[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    uint64_t st = ko_splitmix64(S.seed ^ (S.gameId * PHI) ^ (uint64_t)ko_game_num_turns(rootGame) ^ ROOTSYM_SALT);
    std::vector<float> acc(P, 0.0f), one(P);
    float aw = 0.f, al = 0.f, ae = 0.f;
    for(int i = 0; i < N; i++) {
      st += PHI;
      const int j = i + (int)(ko_splitmix64(st) % (uint64_t)(8 - i));
      std::swap(idx[i], idx[j]);
      float w2[2], e1 = 0.f;
      ev.eval(rootGame, one.data(), w2, idx[i], &e1);
      for(int pos = 0; pos < P; pos++) acc[pos] += one[pos];
      aw += w2[0]; al += w2[1]; ae += e1;
    }
    const float fl = (float)N;
    for(int pos = 0; pos < P; pos++) acc[pos] /= fl;
    aw /= fl; al /= fl; ae /= fl;
    const double v = (double)aw - (double)al;
    if(S.nodes.empty()) {   // a fresh root: this is its evaluation and its first visit
      S.nodes.emplace_back(P);
      GNode& nn = S.nodes.back();
      nn.visits = 1; nn.nnUtility = v; nn.nextPla = ko_game_next_pla(rootGame); nn.depth = ko_game_num_turns(rootGame);
      nn.nnWeight = S.nnWeightOf(ae);
      nn.weightSum = nn.nnWeight; nn.weightSqSum = nn.nnWeight * nn.nnWeight; nn.utilityAvg = v; nn.utilitySqAvg = v * v;
      cnt[0]++;
    } else {                // a kept root: new policy and evaluation, statistics untouched until the next re-computation (isReInit)
      S.nodes[0].nnUtility = v; S.nodes[0].nnWeight = S.nnWeightOf(ae);
    }
    for(int pos = 0; pos < P; pos++) S.nodes[0].policy[pos] = acc[pos];
    cnt[1] += (uint64_t)N;
  }
  if(!ko_game_finished(rootGame)) S.maybeNoiseRoot(rootGame);
  for(int it = 0; it < p->maxVisits && !ko_game_finished(rootGame); it++) {
    ko_game_copy(g, rootGame);
    std::vector<std::pair<int, int>> path;
    int kind = 0, target = -1;   // 1 new node, 2 new terminal child, 3 terminal revisit, 4 root evaluation, 5 catch-up, 6 transposition
    double v = 0.0;
    Key leafKey(0, 0), leafBias(0, 0);
    bool haveBias = false;
    if(S.nodes.empty()) kind = 4;
    else if(S.nodes[0].visits >= p->maxVisits) break;
    else {
      int node = 0, depth = 0;
      while(true) {
        GNode& nd = S.nodes[node];
        const int pla = nd.nextPla;
        double partT[32] = {0}, partM[32] = {0};
        const std::vector<int> ks = S.kids(nd);
        for(int k = 0; k < (int)ks.size(); k++) {
          const int pos = ks[k];
          int cv; double cw, cu;
          S.childStats(nd, pos, cv, cw, cu);
          partT[k & 31] = partT[k & 31] + GraphSearch::childWeight(cw, nd.edgeN[pos], cv);
          partM[k & 31] = partM[k & 31] + (double)nd.policy[pos];
        }
        const double total = butterfly(partT), mass = butterfly(partM);
        const double parentUtility = nd.utilityAvg;
        double parentUtilityForFPU = parentUtility;
        if(p->fpuParentWeightByVisitedPolicy) {   // searchexplorehelpers.cpp:279-282
          const double pw = p->fpuParentWeightByVisitedPolicyPow;
          const double raised = mass <= 0.0 ? 0.0 : pw == 1.0 ? mass : pw == 2.0 ? mass * mass : detExp(pw * detLog(mass));
          const double avgWeight = std::min(1.0, raised);
          parentUtilityForFPU = avgWeight * parentUtility + (1.0 - avgWeight) * nd.nnUtility;
        }
        const double red = (depth == 0 ? p->rootFpuReductionMax : p->fpuReductionMax) * std::sqrt(mass);
        const double fpu = pla == 2 ? parentUtilityForFPU - red : parentUtilityForFPU + red;
        const double scale = p->cpuctExploration * std::sqrt(total + 0.01);
        double bestVal = 0.0; int bestOrd = 1 << 20, bestPos = -1;
        float newP = -1.0f; int newPos = -1;
        for(int pos = 0; pos < P; pos++) {
          const float pr = nd.policy[pos];
          if(nd.child[pos] != -1) {
            int cv; double cw, cu;
            S.childStats(nd, pos, cv, cw, cu);
            const double w = GraphSearch::childWeight(cw, nd.edgeN[pos], cv);
            double val = (scale * (double)pr) / (1.0 + w) + (pla == 2 ? cu : -cu);
            // rootDesiredPerChildVisitsCoeff: funnel visits into under-visited root children (searchexplorehelpers.cpp:150-155)
            if(depth == 0 && p->rootDesiredPerChildVisitsCoeff > 0.0 && pr > 0 &&
               w < std::sqrt(((double)pr * total) * p->rootDesiredPerChildVisitsCoeff)) val = 1e20;
            const int o = nd.order[pos];
            if(bestPos < 0 || val > bestVal || (val == bestVal && o < bestOrd)) { bestVal = val; bestOrd = o; bestPos = pos; }
          } else if(pr >= 0.0f) {
            if(pr > newP) { newP = pr; newPos = pos; }
          }
        }
        bool takeNew = false;
        if(newPos >= 0) {
          const double valNew = (scale * (double)newP) / 1.0 + (pla == 2 ? fpu : -fpu);
          takeNew = bestPos < 0 || valNew > bestVal;
        }
        const int pos = takeNew ? newPos : bestPos;
        if(pos < 0) { kind = 0; break; }
        path.push_back({node, pos});
        depth++;
        if(!takeNew) {
          const int cc = nd.child[pos];
          if(cc <= -2) { kind = 3; v = terminalValue(-2 - cc); break; }
          if(nd.edgeN[pos] < S.nodes[cc].visits) { kind = 5; break; }   // maybeCatchUpEdgeVisits: no descent
          ko_game_play(g, pos);
          node = cc;
          continue;
        }
        ko_game_copy(before, g);
        ko_game_play(g, pos);
        if(ko_game_finished(g)) { kind = 2; v = terminalValue(ko_game_winner(g)); break; }
        leafKey = GraphSearch::stateKey(g);
        if(p->useGraphSearch) {
          auto f = S.table.find(leafKey);
          if(f != S.table.end()) { kind = 6; target = f->second; break; }
        }
        kind = 1;
        // moveHistory.size() >= 2 after the move (search.cpp:740): the position before it has a last move
        if(p->subtreeValueBiasFactor != 0.0 && ko_game_recent_move_pos(before, 0) >= 0) { haveBias = true; leafBias = S.biasKey(before, pos); }
        break;
      }
    }
    if(kind == 0) continue;
    if((kind == 1 || kind == 4) && S.nodes.size() >= maxNodes) continue;   // node pool exhausted (re-use in graph mode only): the visit is dropped
    int newIdx = -1;
    if(kind == 1 || kind == 4) {
      ev.eval(g, pol.data(), wl, -1, &shortErr);
      v = (double)wl[0] - (double)wl[1];
      newIdx = (int)S.nodes.size();
      S.nodes.emplace_back(P);
      GNode& nn = S.nodes.back();
      nn.nnWeight = S.nnWeightOf(shortErr);
      nn.visits = 1; nn.weightSum = nn.nnWeight; nn.weightSqSum = nn.nnWeight * nn.nnWeight; nn.nnUtility = v; nn.nextPla = ko_game_next_pla(g);
      nn.key = kind == 1 ? leafKey : Key(0, 0); nn.depth = ko_game_num_turns(g);
      for(int pos = 0; pos < P; pos++) nn.policy[pos] = pol[pos];
      double utility = v;
      if(haveBias) {
        auto f = S.biasIndex.find(leafBias);
        if(f == S.biasIndex.end()) { f = S.biasIndex.insert({leafBias, (int)S.bias.size()}).first; S.bias.emplace_back(); S.bias.back().key = leafBias; }
        nn.biasEntry = f->second;
        const BiasEntry& E = S.bias[nn.biasEntry];
        if(E.weightSum > 0.001) utility = utility + (p->subtreeValueBiasFactor * E.deltaSum) / E.weightSum;   // addLeafValue :27-37
      }
      nn.utilityAvg = utility; nn.utilitySqAvg = utility * utility;
      if(kind == 1 && p->useGraphSearch) S.table[leafKey] = newIdx;
      if(kind == 4) S.maybeNoiseRoot(rootGame);
    }
    for(int d = (int)path.size() - 1; d >= 0; d--) {
      GNode& nd = S.nodes[path[d].first];
      const int pos = path[d].second;
      if(d + 1 == (int)path.size() && (kind == 1 || kind == 2 || kind == 6)) {
        nd.child[pos] = kind == 1 ? newIdx : kind == 6 ? target : (v > 0.0 ? -4 : v < 0.0 ? -3 : -2);
        nd.order[pos] = (uint8_t)nd.numChildren;
        nd.numChildren++;
      }
      nd.edgeN[pos] += 1;
      S.recompute(nd, 1, d == 0);
    }
    cnt[0]++;
    if(kind == 1 || kind == 4) cnt[1]++;
    else if(kind == 2 || kind == 3) cnt[2]++;
    else if(kind == 6) cnt[3]++;
    else cnt[4]++;
  }
  ko_game_destroy(g); ko_game_destroy(before);
  const bool have = !S.nodes.empty();
  if(rootVisits) *rootVisits = have ? S.nodes[0].visits : 0;
  if(rootUtilitySum) *rootUtilitySum = have ? S.nodes[0].utilityAvg * (double)S.nodes[0].visits : 0.0;
  for(int pos = 0; pos < P; pos++) {
    const bool ex = have && S.nodes[0].child[pos] != -1;
    int cv = 0; double cw = 0.0, cu = 0.0;
    if(ex) S.childStats(S.nodes[0], pos, cv, cw, cu);
    if(edgeVisits) edgeVisits[pos] = ex ? S.nodes[0].edgeN[pos] : 0;
    if(edgeUtilitySum) edgeUtilitySum[pos] = ex ? cu * (double)S.nodes[0].edgeN[pos] : 0.0;
    if(policyOut) policyOut[pos] = have ? S.nodes[0].policy[pos] : 0.f;
    if(orderOut) orderOut[pos] = ex ? S.nodes[0].order[pos] : 255;
  }
  if(counters) for(int i = 0; i < 5; i++) counters[i] += cnt[i];
  if(digest) *digest = graphDigest(S);
  g_lastPlaySelection.assign(P, 0.0);
  if(have && !ko_game_finished(rootGame)) S.playSelectionValues(g_lastPlaySelection.data());
}

// play-selection values of the root after the last graph search on this thread ([P], 0 for moves without a child)
void ko_search_last_play_selection(double* out, int P) {
  for(int pos = 0; pos < P; pos++) out[pos] = pos < (int)g_lastPlaySelection.size() ? g_lastPlaySelection[pos] : 0.0;
}

void ko_search_run_graph(const ko_game* rootGame, int x_size, int y_size, const ko_search_params* p, const ko_model* modelOrNull,
                         int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits, double* edgeUtilitySum, float* policyOut,
                         uint8_t* orderOut, uint64_t counters[5], uint64_t* digest) {
  GraphSearch S{x_size, y_size, 4 * x_size * y_size, p, {}, {}, {}, {}};
  S.seed = p->noiseSeed; S.gameId = p->noiseGameId;
  graphRun(S, rootGame, x_size, y_size, p, modelOrNull, rootVisits, rootUtilitySum, edgeVisits, edgeUtilitySum, policyOut, orderOut, counters, digest);
}

// Persistent graph search with re-use between moves: continue() searches the current graph on until the root has maxVisits
// visits, advance() re-roots it at the move played (GraphSearch::advance).
struct ko_graph_search { GraphSearch S; ko_search_params params; };
ko_graph_search* ko_graph_search_create(int x_size, int y_size, const ko_search_params* p) {
  ko_graph_search* s = new ko_graph_search{GraphSearch{x_size, y_size, 4 * x_size * y_size, nullptr, {}, {}, {}, {}}, *p};
  s->S.p = &s->params;
  s->S.seed = p->noiseSeed; s->S.gameId = p->noiseGameId;
  return s;
}
void ko_graph_search_destroy(ko_graph_search* s) { delete s; }
void ko_graph_search_continue(ko_graph_search* s, const ko_game* rootGame, const ko_model* modelOrNull, int32_t* rootVisits, double* rootUtilitySum,
                              int32_t* edgeVisits, double* edgeUtilitySum, float* policyOut, uint8_t* orderOut, uint64_t counters[5], uint64_t* digest) {
  graphRun(s->S, rootGame, s->S.W, s->S.H, &s->params, modelOrNull, rootVisits, rootUtilitySum, edgeVisits, edgeUtilitySum, policyOut, orderOut, counters, digest);
}
void ko_graph_search_advance(ko_graph_search* s, int movePos) { s->S.p = &s->params; s->S.advance(movePos); }
uint64_t ko_graph_search_digest(const ko_graph_search* s) { return graphDigest(s->S); }
int ko_graph_search_num_nodes(const ko_graph_search* s) { return (int)s->S.nodes.size(); }

}  // extern "C"

extern "C" {

// Persistent tree with re-use between moves (Search::makeMove keeps the chosen child's subtree): run() continues the
// search of the current tree until the root has maxVisits visits, advance() re-roots at the child reached by movePos
// (an unexpanded or terminal child drops the tree).
struct ko_search { std::vector<Node> nodes; };
ko_search* ko_search_create(void) { return new ko_search(); }
void ko_search_destroy(ko_search* s) { delete s; }
void ko_search_clear(ko_search* s) { s->nodes.clear(); }
void ko_search_continue(ko_search* s, const ko_game* rootGame, int x_size, int y_size, const ko_search_params* p, const ko_model* modelOrNull,
                        int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits, double* edgeUtilitySum, float* policyOut,
                        uint8_t* orderOut, uint64_t counters[3]) {
  searchRunOnTree(s->nodes, rootGame, x_size, y_size, p, modelOrNull, rootVisits, rootUtilitySum, edgeVisits, edgeUtilitySum, policyOut, orderOut, counters);
}
void ko_search_advance(ko_search* s, int movePos) {
  if(s->nodes.empty() || movePos < 0) { s->nodes.clear(); return; }
  const int c = s->nodes[0].child[movePos];
  if(c < 0) { s->nodes.clear(); return; }
  // breadth-first copy of the subtree (the device re-roots in the same order, although node indices carry no meaning)
  std::vector<Node> out;
  std::vector<int> queue{c};
  for(size_t i = 0; i < queue.size(); i++) {
    Node n = s->nodes[queue[i]];
    for(size_t pos = 0; pos < n.child.size(); pos++)
      if(n.child[pos] >= 0) { queue.push_back(n.child[pos]); n.child[pos] = (int)queue.size() - 1; }
    out.push_back(std::move(n));
  }
  s->nodes.swap(out);
}

// The move played after a search: proportional to the root children's visits for ply < temperaturePlies (counter RNG),
// otherwise the most visited child, ties to the earliest created.  Returns -1 if the root has no child.
int ko_search_choose(const int32_t* edgeVisits, const uint8_t* order, int P, int ply, int temperaturePlies, uint64_t seed, uint64_t gameId) {
  long long total = 0;
  int bestN = -1, bestOrd = 1 << 20, bestPos = -1;
  for(int pos = 0; pos < P; pos++)
    if(order[pos] != 255) {
      total += edgeVisits[pos];
      if(edgeVisits[pos] > bestN || (edgeVisits[pos] == bestN && order[pos] < bestOrd)) { bestN = edgeVisits[pos]; bestOrd = order[pos]; bestPos = pos; }
    }
  if(ply < temperaturePlies && total > 0) {
    const uint64_t r = ko_splitmix64(seed ^ (gameId * PHI) ^ (uint64_t)ply ^ CHOOSE_SALT);
    long long k = (long long)(r % (uint64_t)total);
    for(int pos = 0; pos < P; pos++)
      if(order[pos] != 255) { if(k < edgeVisits[pos]) return pos; k -= edgeVisits[pos]; }
  }
  return bestPos;
}

// The move played after a search under the reference's temperature schedule (Search::getChosenMoveLoc -> getPlaySelectionValues'
// subtract / prune step, cpp/search/searchresults.cpp:287-298; chooseIndexWithTemperature, cpp/search/searchhelpers.cpp:12-49;
// interpolateEarly :463-467).  Canonical: the play-selection value of a move is its edge visit count (no LCB, no reduced
// weights), candidates are walked in policy-index order, the uniform draw comes from the counter stream of ko_search_choose,
// log / exp are detLog / detExp.  T <= 1e-4 picks the most visited move (ties: earliest created).
int ko_search_choose_values(const double* values, const uint8_t* order, int P, int boardArea, int ply, double tempEarly, double tempLate,
                            double halflife, double subtract, double prune, uint64_t seed, uint64_t gameId);
int ko_search_choose_temperature(const int32_t* edgeVisits, const uint8_t* order, int P, int boardArea, int ply, double tempEarly, double tempLate,
                                 double halflife, double subtract, double prune, uint64_t seed, uint64_t gameId) {
  std::vector<double> v(P);
  for(int pos = 0; pos < P; pos++) v[pos] = (double)edgeVisits[pos];
  return ko_search_choose_values(v.data(), order, P, boardArea, ply, tempEarly, tempLate, halflife, subtract, prune, seed, gameId);
}
// the same on arbitrary play-selection values (the full getPlaySelectionValues of GraphSearch::playSelectionValues)
int ko_search_choose_values(const double* values, const uint8_t* order, int P, int boardArea, int ply, double tempEarly, double tempLate,
                            double halflife, double subtract, double prune, uint64_t seed, uint64_t gameId) {
  const double* edgeVisits = values;
  double maxValue = 0.0;
  for(int pos = 0; pos < P; pos++) if(order[pos] != 255 && (double)edgeVisits[pos] > maxValue) maxValue = (double)edgeVisits[pos];
  if(maxValue <= 0.0) return -1;
  const double amountToSubtract = std::min(subtract, maxValue / 64.0), amountToPrune = std::min(prune, maxValue / 64.0);
  std::vector<double> v(P, 0.0);
  double newMax = 0.0;
  int bestPos = -1, bestOrd = 1 << 20;
  for(int pos = 0; pos < P; pos++) {
    if(order[pos] == 255) continue;
    double x = (double)edgeVisits[pos];
    if(x < amountToPrune) x = 0.0;
    else { x = x - amountToSubtract; if(x <= 0.0) x = 0.0; }
    v[pos] = x;
    if(x > newMax || (x == newMax && x > 0.0 && order[pos] < bestOrd)) { newMax = x; bestPos = pos; bestOrd = order[pos]; }
  }
  const double hl = halflife > 0.0 ? halflife : 19.0;
  const double halflives = (((double)ply / hl) * 19.0) / std::sqrt((double)boardArea);
  const double T = tempLate + (tempEarly - tempLate) * detExp(halflives * detLog(0.5));
  if(T <= 1.0e-4 || newMax <= 0.0) return bestPos;
  const double logMax = detLog(newMax);
  double sum = 0.0;
  for(int pos = 0; pos < P; pos++) { v[pos] = v[pos] <= 0.0 ? 0.0 : detExp((detLog(v[pos]) - logMax) / T); sum += v[pos]; }
  const uint64_t r = ko_splitmix64(seed ^ (gameId * PHI) ^ (uint64_t)ply ^ CHOOSE_SALT);
  const double d = ((double)(r & ((1ULL << 53) - 1ULL)) * (1.0 / 9007199254740992.0)) * sum;
  double acc = 0.0;
  int last = -1;
  for(int pos = 0; pos < P; pos++) {
    if(order[pos] == 255) continue;
    last = pos;
    acc += v[pos];
    if(acc > d) return pos;
  }
  return last;
}

// Training rows of one finished game (TrainingWriteBuffers::addRow, cpp/dataio/trainingwrite.cpp:316-566, restated with the
// canonical choices listed at kc_search_read_training_rows in include/katacoffee_b200.h): the game is replayed from the
// empty board with `movePos`, turn i carries the root visits / utility sum / visit counts of the search that chose move i.
void ko_training_rows(int x_size, int y_size, int win_len, int R, const int32_t* movePos, const int32_t* rootN, const double* rootW,
                      const int16_t* visits, uint64_t gameId, uint8_t* bin, float* globalIn, int16_t* policy, float* globalT, int8_t* value) {
  const int HW = x_size * y_size, P = 4 * HW, PB = (HW + 7) / 8;
  std::vector<ko_game*> pos(R + 1);
  pos[0] = ko_game_create(x_size, y_size, win_len);
  for(int i = 0; i < R; i++) {
    pos[i + 1] = ko_game_create(x_size, y_size, win_len);
    ko_game_copy(pos[i + 1], pos[i]);
    ko_game_play(pos[i + 1], movePos[i]);
  }
  const ko_game* fin = pos[R];
  const int winner = ko_game_winner(fin);
  const double finalWin = winner == 2 ? 1.0 : winner == 1 ? 0.0 : 0.5;
  const double area = (double)HW;
  const double nowFactors[5] = {0.0, 1.0 / (1.0 + area * 0.176), 1.0 / (1.0 + area * 0.056), 1.0 / (1.0 + area * 0.016), 1.0};
  const uint64_t gh0 = ko_splitmix64(gameId), gh1 = ko_splitmix64(gameId ^ PHI);
  // longest same-colour run through each final stone, over the four line directions
  std::vector<int> maxRun(HW, 0);
  const int DX[4] = {0, -1, -1, 1}, DY[4] = {-1, 0, -1, -1};
  for(int y = 0; y < y_size; y++)
    for(int x = 0; x < x_size; x++) {
      const int col = ko_game_color_at(fin, x, y);
      if(col == 0) continue;
      for(int d = 0; d < 4; d++) {
        int len = 1;
        for(int sgn = -1; sgn <= 1; sgn += 2)
          for(int k = 1;; k++) {
            const int xx = x + sgn * k * DX[d], yy = y + sgn * k * DY[d];
            if(xx < 0 || yy < 0 || xx >= x_size || yy >= y_size || ko_game_color_at(fin, xx, yy) != col) break;
            len++;
          }
        if(len > maxRun[y * x_size + x]) maxRun[y * x_size + x] = len;
      }
    }
  std::vector<float> planes((size_t)15 * HW);
  for(int i = 0; i < R; i++) {
    const ko_game* g = pos[i];
    const int pla = ko_game_next_pla(g);
    float glob = 0.f;
    ko_game_fill_row_v1(g, pla, x_size, y_size, 0, planes.data(), &glob);
    for(int c = 0; c < 15; c++)
      for(int byte = 0; byte < PB; byte++) {
        uint8_t b = 0;
        for(int k = 0; k < 8 && byte * 8 + k < HW; k++) b |= (uint8_t)((uint8_t)planes[(size_t)c * HW + byte * 8 + k] << (7 - k));
        bin[((size_t)i * 15 + c) * PB + byte] = b;
      }
    globalIn[i] = glob;
    for(int p = 0; p < P; p++) {
      policy[((size_t)i * 2 + 0) * P + p] = visits[(size_t)i * P + p];
      policy[((size_t)i * 2 + 1) * P + p] = (i + 1 < R) ? visits[(size_t)(i + 1) * P + p] : (int16_t)1;
    }
    const ko_game* g2 = pos[std::min(i + 2, R)];
    const ko_game* g3 = pos[std::min(i + 6, R)];
    auto rel = [&](const ko_game* gg, int x, int y) -> int8_t {
      const int col = ko_game_color_at(gg, x, y);
      return (int8_t)(col == 0 ? 0 : (col == pla ? 1 : -1));
    };
    int8_t* v = value + (size_t)i * 5 * HW;
    for(int y = 0; y < y_size; y++)
      for(int x = 0; x < x_size; x++) {
        const int cell = y * x_size + x;
        v[cell] = rel(fin, x, y);
        v[HW + cell] = 0;
        v[2 * HW + cell] = rel(g2, x, y);
        v[3 * HW + cell] = rel(g3, x, y);
        v[4 * HW + cell] = (int8_t)maxRun[cell];
      }
    float* gt = globalT + (size_t)i * 64;
    for(int k = 0; k < 64; k++) gt[k] = 0.f;
    for(int f = 0; f < 5; f++) {
      const double nowFactor = nowFactors[f];
      double winV = 0.0, lossV = 0.0, weightLeft = 1.0;
      for(int j = i; j <= R; j++) {
        double weightNow;
        if(j == R) { weightNow = weightLeft; weightLeft = 0.0; }
        else { weightNow = weightLeft * nowFactor; weightLeft = weightLeft * (1.0 - nowFactor); }
        double tw, tl;
        if(j == R) { tw = finalWin; tl = 1.0 - finalWin; }
        else { const double u = rootW[j] / (double)rootN[j]; tw = (1.0 + u) * 0.5; tl = (1.0 - u) * 0.5; }
        winV = winV + weightNow * (pla == 2 ? tw : tl);
        lossV = lossV + weightNow * (pla == 2 ? tl : tw);
      }
      gt[2 * f] = (float)winV; gt[2 * f + 1] = (float)lossV;
    }
    gt[25] = 1.0f; gt[26] = 1.0f; gt[27] = 1.0f; gt[28] = (i + 1 < R) ? 1.0f : 0.0f; gt[33] = 1.0f;
    for(int k = 36; k <= 40; k++) gt[k] = 1.0f;
    gt[41] = (float)(gh0 & 0x3FFFFF); gt[42] = (float)((gh0 >> 22) & 0x3FFFFF); gt[43] = (float)((gh0 >> 44) & 0xFFFFF);
    gt[44] = (float)(gh1 & 0x3FFFFF); gt[45] = (float)((gh1 >> 22) & 0x3FFFFF); gt[46] = (float)((gh1 >> 44) & 0xFFFFF);
    gt[51] = (float)ko_game_num_turns(g);
    gt[60] = (float)rootN[i];
    gt[63] = 1.0f;
  }
  for(ko_game* g : pos) ko_game_destroy(g);
}

}  // extern "C"
