// Batched Coffee tree search on the device: G games, one PUCT tree each, advanced in lock step -- every iteration each
// game descends its tree to one leaf, the G leaves go through the hot path (V1 planes -> net forward -> post-processing)
// as ONE batch, and the results are expanded into the trees and backed up.  This is the caller of the leaf-evaluation
// path in the reference: Search::playoutDescend (cpp/search/search.cpp:935-1160),
// selectBestChildToDescend / getExploreSelectionValue / getFpuValueForChildrenAssumeVisited
// (cpp/search/searchexplorehelpers.cpp:9-45, 248-451), addLeafValue / recomputeNodeStats
// (cpp/search/searchupdatehelpers.cpp:12-76, 151-326), restated for Coffee positions.
//
// Canonical semantics (what the CPU checker under tests and this file both implement, bit for bit):
//  * SearchParams() defaults (cpp/search/searchparams.cpp:8-90) with valueWeightExponent = 0, i.e. no utility-based
//    re-weighting of children (searchupdatehelpers.cpp:340-366): winLossUtilityFactor 1, cpuctExploration c,
//    cpuctExplorationLog 0, cpuctUtilityStdevScale 0, fpuReductionMax f, rootFpuReductionMax f_root, fpuLossProp 0, no
//    uncertainty weighting (every visit has weight 1), no graph search, no noise, no virtual losses (one descent per
//    tree per iteration).  With unit weights recomputeNodeStats' weighted average over children plus the node's own
//    evaluation equals the running mean of all leaf utilities below the node, which is what is accumulated (W, N).
//  * selection value of a child (white-positive utilities, pla = player to move at the parent):
//      exploreScaling = c * sqrt(totalChildWeight + 0.01)                     (searchexplorehelpers.cpp:17-25)
//      value = exploreScaling * prior / (1 + childVisits) + (pla == WHITE ? Q : -Q)          (:27-45)
//      unvisited: Q = fpu = parentUtility -/+ fpuReduction * sqrt(policy mass of visited children)   (:248-320)
//    existing children are scanned in creation order and only a strictly larger value replaces the best; the single
//    best unvisited move (largest prior, lowest policy index on ties) then competes, again strictly (:385-449).
//  * a move that ends the game makes a terminal child without a net evaluation; each visit to it adds the result
//    (white win +1, black win -1, draw 0; search.cpp:943-953 with ledger C for the draw).
//  * the move played after the search is sampled in proportion to the root children's visits for the first
//    `temperaturePlies` plies (chosenMoveTemperature 1) with the counter RNG of SURVEY.md 8(d), afterwards the most
//    visited child (ties: earliest created).
//  All search arithmetic is IEEE double without fused multiply-add (explicit round-to-nearest intrinsics here,
//  -ffp-contract=off in the oracle), so trees agree exactly whenever the leaf evaluations agree exactly.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "games.h"
#include "net.h"

namespace kc {

constexpr int MAX_PATH = 52;   // a 7x7 game has at most 49 plies
constexpr uint64_t PHI = 0x9E3779B97F4A7C15ULL;
constexpr uint64_t CHOOSE_SALT = 0xC0FFEE5EA4C4ULL;
constexpr uint64_t SYM_SALT = 0x5A11E7C0FFEEULL;

struct SearchCfg {
  int P;             // policy size 4*H*W
  int LW;            // legal words
  int nodeStride;    // bytes
  int maxNodes;
  int maxVisits;
  int temperaturePlies;
  int numGames;
  int autoRefill;
  int compact;       // leaves that need the net are packed into a dense batch (slot = arrival order); 0: slot = game
  int reuseTree;     // keep the chosen child's subtree for the next search (Search::makeMove)
  int randomSym;     // NNEvaluator's nnRandomize: every leaf is evaluated under a symmetry drawn from its position (nneval.cpp:515-524)
  int gOff, gCnt, half;   // the games a per-iteration kernel launch covers: [gOff, gOff + gCnt); half = index of its batch counter
  int graph;         // node-centric statistics (graph search and / or subtree value bias), see "graph mode" below
  int useTable;      // graph mode: look new positions up in the transposition table (SearchParams::useGraphSearch)
  int polOff;        // byte offset of the per-move arrays inside a node
  int tableCap;      // graph mode: slots of the per-game transposition and bias tables (power of two, >= 2 * maxNodes)
  double cpuct, fpuRed, rootFpuRed;
  double biasFactor, biasExp, biasFreeProp;   // SearchParams::subtreeValueBiasFactor / WeightExponent / FreeProp
  // graph mode options (SearchParams of the same names; 0 / 1.0 = off)
  int rootNoise, fpuPW;
  double noiseConc, noiseWeight, rootTemp, rootTempEarly, tempHalflife, fpuPWPow, rootDesired, vwExp;
  double moveTemp, moveTempEarly, moveSubtract, movePrune;   // chosenMoveTemperature / Early / Subtract / Prune (either temperature > 0: schedule)
  int boardArea;
  // the rest of selfplay1.cfg / the GTP defaults: LCB move selection, root symmetry averaging, uncertainty weighting
  int useLcb, nonBuggyLcb, rootSyms, noisePruning, useUncertainty;
  double lcbStdevs, minVisitPropLcb, uncCoeff, uncExp, uncMaxWeight, noisePruneScale, noisePruneCap;
  int iterStamp;           // NN cache: number of the current iteration (entries hit in it are pinned against replacement until it is over)
  int fullPlaySelection;   // the move choice runs Search::getPlaySelectionValues (child weights, reduced weights, LCB) instead of plain edge visits
  uint64_t seed;
};

struct TreeMem {
  uint8_t* nodes;         // [G][maxNodes][nodeStride]
  int* nodeCount;         // [G]
  int* pathNode;          // [G][MAX_PATH]
  uint8_t* pathPos;       // [G][MAX_PATH]
  int* pathLen;           // [G]
  int* leafKind;          // [G] 0 idle, 1 new node (needs the net), 2 new terminal child, 3 revisit of a terminal child, 4 root evaluation
  double* leafValue;      // [G] result of a terminal leaf (white-positive)
  int* leafNextPla;       // [G] player to move at a new node
  int* leafSlot;          // [G] row of the evaluation batch holding this game's leaf
  int* evalCount;         // [1] rows of the current evaluation batch (compact mode)
  uint8_t* nodesAlt;      // second tree buffer: re-rooting copies the kept subtree there, then the two swap
  int* rerootQueue;       // [G][maxNodes] breadth-first order of the kept subtree (old node indices)
  int* active;            // [1] set by k_select when any game still had a visit to make
  unsigned long long* stats;  // 0 visits, 1 net evaluations, 2 terminal visits, 3 moves played, 4 games finished, 5 black, 6 white, 7 draws,
                              // 8 transposition hits, 9 catch-up visits (graph mode)
  // graph mode
  uint64_t* tblKeys;      // [G][tableCap][2] transposition keys
  int* tblVals;           // [G][tableCap] node index + 1 (0 = empty slot)
  uint64_t* biasKeys;     // [G][tableCap][2] subtree-value-bias keys (key0 == 0: empty slot)
  double* biasVals;       // [G][tableCap][2] deltaUtilitySum, weightSum
  uint64_t* leafKey;      // [G][2] transposition key of a new node
  uint64_t* leafBiasKey;  // [G][2] bias key of a new node (key0 == 0: none)
  int* leafTarget;        // [G] existing node a new edge transposes to
  // graph mode with tree re-use: re-rooting rebuilds the tables into a second set, then the two swap (like nodes / nodesAlt)
  uint64_t* tblKeysAlt; int* tblValsAlt; uint64_t* biasKeysAlt; double* biasValsAlt;
  int* remap;             // [G][maxNodes] old node index -> new index (-1: dropped)
  // NN cache shared by all games of the search (NNCacheTable, cpp/neuralnet/nneval.cpp:874-932; selfplay1.cfg:121 nnCacheSizePowerOfTwo = 21):
  // direct-mapped on the low bits of the key, entries hold the post-processed outputs.  The key is the whole identity of the net's inputs
  // (sit-hash of the player to move, the last five moves with their players, the last direction, min(numTurns, 5)), so a hit returns
  // bit for bit what the evaluation would have returned and the search result does not depend on the cache.
  uint64_t* cacheKeys;    // [N][2] (0, 0 = empty)
  int* cacheState;        // [N][2]: lock (1 while an entry is being written), iteration stamp of the last hit
  float* cachePolicy;     // [N][P]
  float* cacheScalars;    // [N][4] whiteWin, whiteLoss, varTimeLeft, shorttermWinlossError
  unsigned cacheMask;     // N - 1; 0 = no cache
  int* leafCache;         // [G] entry the leaf's evaluation comes from (hit); -1: evaluated in the batch (then inserted); -3: another game evaluates the
                          //     same position in this very batch (leafOwner) -- in lock step the duplicates of an iteration arrive together
  unsigned long long* claimWord;   // [N] (iteration stamp << 32) | (owner game + 1), 0xffffffff in the low half while the key is being written
  uint64_t* claimKey;     // [N][2] key the slot is claimed for in that iteration
  int* leafOwner;         // [G] the game whose batch row a -3 leaf reads
  uint64_t* leafCKey;     // [G][2] cache key of a leaf that missed
  const double* tcdf;     // [2000] Student-t (3 degrees of freedom) cdf on [-50, 50] for valueWeightExponent
  const double* stdevTab; // [1025] sqrt(1e-8 + 1 / (1.5 sqrt(w))) for integer child weights w (the common case): same bits as computing it
};
constexpr int NUM_STATS = 16;

// Training-row capture (reference cpp/dataio/trainingwrite.cpp:316-566 TrainingWriteBuffers::addRow,
// cpp/program/play.cpp:1428-1460).  Every move played by kc_search_play is recorded per game (position before the move,
// root visits / utility, visit counts); when the game ends the rows are finished (result-dependent targets) and appended to
// an output buffer.  Canonical row (the reference's literal 64-channel layout, entries it derives from Go-only or
// random state fixed): see the comment of kc_search_read_training_rows in the header.
struct TrainMem {
  int enabled, maxPlies, maxRows, pad_;
  uint64_t* recBlack; uint64_t* recWhite; uint64_t* recMisc;   // [G][maxPlies] position before the move
  int* recN; double* recW;                                     // [G][maxPlies] root visits, root utility sum (white-positive)
  int16_t* recVisits;                                          // [G][maxPlies][P]
  int* recCount;                                               // [G]
  uint64_t* recGameId;                                         // [G] id of the game being recorded
  int* rowCount;                                               // [2]: rows written, rows dropped (buffer full)
  uint8_t* outBin; float* outGlobalIn; int16_t* outPolicy; float* outGlobalT; int8_t* outValue;
  double nowFactor[5];                                         // fillValueTDTargets factors (trainingwrite.cpp:403-414)
};

// node layout, tree mode:  header { int N; int numChildren; int nextPla; int pad; double W; double pad } | edgeW[P] f64 |
//                           policy[P] f32 | child[P] i32 | edgeN[P] i32 | order[P] u8            (polOff = 32 + 8 P)
//              graph mode: header { int visits; int numChildren; int nextPla; int biasEntry; double weightSum, utilityAvg,
//                           nnUtility, lastBiasDeltaSum, lastBiasWeight; int depth (stones on the board), noised;
//                           uint64 key[2] (transposition key); double utilitySqAvg, weightSqSum, nnWeight } | policy | child | edgeN | order | list u8[P]  (polOff = 104)
struct NodeRef {
  uint8_t* base; int P; int polOff;
  __device__ __forceinline__ int& N() const { return *reinterpret_cast<int*>(base); }
  __device__ __forceinline__ int& numChildren() const { return *reinterpret_cast<int*>(base + 4); }
  __device__ __forceinline__ int& nextPla() const { return *reinterpret_cast<int*>(base + 8); }
  __device__ __forceinline__ double& W() const { return *reinterpret_cast<double*>(base + 16); }
  __device__ __forceinline__ double* edgeW() const { return reinterpret_cast<double*>(base + 32); }
  __device__ __forceinline__ float* policy() const { return reinterpret_cast<float*>(base + polOff); }
  __device__ __forceinline__ int* child() const { return reinterpret_cast<int*>(base + polOff + 4 * P); }
  __device__ __forceinline__ int* edgeN() const { return reinterpret_cast<int*>(base + polOff + 8 * P); }
  __device__ __forceinline__ uint8_t* order() const { return base + polOff + 12 * P; }
  __device__ __forceinline__ uint8_t* list() const { return base + polOff + 13 * P; }   // graph mode: policy index of the k-th created child
  // graph mode header
  __device__ __forceinline__ int& biasEntry() const { return *reinterpret_cast<int*>(base + 12); }
  __device__ __forceinline__ double& weightSum() const { return *reinterpret_cast<double*>(base + 16); }
  __device__ __forceinline__ double& utilityAvg() const { return *reinterpret_cast<double*>(base + 24); }
  __device__ __forceinline__ double& nnUtility() const { return *reinterpret_cast<double*>(base + 32); }
  __device__ __forceinline__ double& lastDelta() const { return *reinterpret_cast<double*>(base + 40); }
  __device__ __forceinline__ double& lastWeight() const { return *reinterpret_cast<double*>(base + 48); }
  __device__ __forceinline__ int& depth() const { return *reinterpret_cast<int*>(base + 56); }
  __device__ __forceinline__ int& noised() const { return *reinterpret_cast<int*>(base + 60); }   // root policy already noised / tempered
  __device__ __forceinline__ uint64_t* key() const { return reinterpret_cast<uint64_t*>(base + 64); }
  __device__ __forceinline__ double& utilitySqAvg() const { return *reinterpret_cast<double*>(base + 80); }   // NodeStats::utilitySqAvg / weightSqSum
  __device__ __forceinline__ double& weightSqSum() const { return *reinterpret_cast<double*>(base + 88); }    // (searchnode.h:44-48): what LCB reads
  __device__ __forceinline__ double& nnWeight() const { return *reinterpret_cast<double*>(base + 96); }       // computeWeightFromNNOutput: 1 without useUncertainty
};
constexpr int GRAPH_POL_OFF = 104;
// child codes: -1 none, >= 0 node index, -2 terminal draw, -3 terminal black win, -4 terminal white win
__device__ __forceinline__ double terminalValue(int winner) { return winner == 2 ? 1.0 : winner == 1 ? -1.0 : 0.0; }

template <class D>
__device__ __forceinline__ void applyMoveLight(const D& dm, GameRegs<typename D::BB>& s, int pos, const uint64_t* __restrict__ zob) {
  using BB = typename D::BB;
  const int fl = flagsOf(s.misc), pla = (fl >> 3) & 3;
  const int dir = pos / dm.HW(), cell = pos % dm.HW();
  const BB bit = (BB)1 << padOf(dm, cell);
  if(pla == 1) s.black |= bit; else s.white |= bit;
  const uint64_t* z = zob + ((size_t)cell * 2 + (pla - 1)) * 2;
  s.h0 ^= z[0]; s.h1 ^= z[1];
  const uint64_t hist = ((s.misc & 0xffffffffULL) << 8) | (uint64_t)(cell | (pla << 6));
  const int nt = numTurnsOf(s.misc) + 1;
  s.misc = (hist & 0xffffffffffULL) | ((uint64_t)dir << 40) | ((uint64_t)(nt & 0xff) << 48) | ((uint64_t)((pla ^ 3) << 3) << 56);
}

__device__ __forceinline__ double warpSumD(double v) {   // butterfly: every lane ends with the same, order-defined sum
  for(int o = 16; o > 0; o >>= 1) v = __dadd_rn(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// NN cache key of a position: everything the planes and the legal mask read (see TreeMem::cacheKeys)
__device__ __forceinline__ void nnCacheKey(const Geom& g, uint64_t h0, uint64_t h1, uint64_t misc, uint64_t& k0, uint64_t& k1) {
  const int np = (flagsOf(misc) >> 3) & 3;
  const uint64_t hist = (misc & 0xffffffffffffULL) | ((uint64_t)min(numTurnsOf(misc), 5) << 48);   // five moves, last direction, capped turn count
  const uint64_t m0 = splitmix64(hist ^ 0x6b4c1e2d9a5f3c71ULL), m1 = splitmix64(m0 ^ hist);
  k0 = h0 ^ g.playerHash[np][0] ^ m0; k1 = h1 ^ g.playerHash[np][1] ^ m1;
  if((k0 | k1) == 0) k0 = 1;   // (0, 0) marks an empty entry
}
// lane 0 of the selecting warp: true (and *entry) if the leaf's evaluation is in the cache; the entry is pinned for this iteration
__device__ __forceinline__ bool nnCacheLookup(const SearchCfg& c, const TreeMem& t, int gi, uint64_t k0, uint64_t k1) {
  const unsigned idx = (unsigned)k0 & t.cacheMask;
  if(t.cacheKeys[2 * (size_t)idx] == k0 && t.cacheKeys[2 * (size_t)idx + 1] == k1) {
    t.cacheState[2 * (size_t)idx + 1] = c.iterStamp;   // racing writers store the same value
    t.leafCache[gi] = (int)idx;
    return true;
  }
  t.leafCKey[2 * (size_t)gi] = k0; t.leafCKey[2 * (size_t)gi + 1] = k1;
  // a miss: claim the slot for this key and this iteration, or follow the game that has
  const unsigned long long stampHi = (unsigned long long)(unsigned)c.iterStamp << 32;
  unsigned long long w = t.claimWord[idx];
  if((w >> 32) != (unsigned)c.iterStamp) {
    if(atomicCAS(&t.claimWord[idx], w, stampHi | 0xffffffffULL) == w) {   // ours: key first, then the owner becomes visible
      t.claimKey[2 * (size_t)idx] = k0; t.claimKey[2 * (size_t)idx + 1] = k1;
      __threadfence();
      atomicExch(&t.claimWord[idx], stampHi | (unsigned long long)(gi + 1));
      t.leafCache[gi] = -1;
      return false;
    }
    w = t.claimWord[idx];
  }
  if((w >> 32) == (unsigned)c.iterStamp && (unsigned)w != 0xffffffffu && (unsigned)w != 0u) {
    __threadfence();
    if(t.claimKey[2 * (size_t)idx] == k0 && t.claimKey[2 * (size_t)idx + 1] == k1) {
      t.leafCache[gi] = -3;
      t.leafOwner[gi] = (int)(unsigned)w - 1;
      return true;
    }
  }
  t.leafCache[gi] = -2;   // evaluated in the batch, not inserted (the slot belongs to another key this iteration)
  return false;
}
// whole warp, after an evaluation that missed: store it unless the slot was hit in this iteration or another warp is writing it
__device__ __forceinline__ void nnCacheInsert(const SearchCfg& c, const TreeMem& t, int gi, int lane, const float* __restrict__ pol, const float* __restrict__ wl,
                                              const float* __restrict__ misc) {
  const uint64_t k0 = t.leafCKey[2 * (size_t)gi], k1 = t.leafCKey[2 * (size_t)gi + 1];
  const unsigned idx = (unsigned)k0 & t.cacheMask;
  int ok = 0;
  if(lane == 0) ok = t.cacheState[2 * (size_t)idx + 1] != c.iterStamp && atomicCAS(&t.cacheState[2 * (size_t)idx], 0, 1) == 0;
  ok = __shfl_sync(0xffffffffu, ok, 0);
  if(!ok) return;
  if(lane == 0) { t.cacheKeys[2 * (size_t)idx] = 0; t.cacheKeys[2 * (size_t)idx + 1] = 0; }   // no reader runs beside this kernel; kept tidy anyway
  for(int pos = lane; pos < c.P; pos += 32) t.cachePolicy[(size_t)idx * c.P + pos] = pol[pos];
  if(lane == 0) {
    float* sc = t.cacheScalars + (size_t)idx * 4;
    sc[0] = wl[0]; sc[1] = wl[1]; sc[2] = misc ? misc[0] : 0.f; sc[3] = misc ? misc[1] : 0.f;
  }
  __syncwarp();
  if(lane == 0) {
    t.cacheKeys[2 * (size_t)idx] = k0; t.cacheKeys[2 * (size_t)idx + 1] = k1;
    __threadfence();
    atomicExch(&t.cacheState[2 * (size_t)idx], 0);
  }
}

// ---------------------------------------------------------------------------------------------
// select: one warp per game descends from the root to a leaf
// ---------------------------------------------------------------------------------------------
template <class D>
__global__ void __launch_bounds__(128, 10) k_select(const Geom g, const SearchCfg c, State root, State leaf, TreeMem t, const uint64_t* __restrict__ zob, int8_t* __restrict__ leafSym) {
  const D dm(g);
  using BB = typename D::BB;
  const int li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(li >= c.gCnt) return;
  const int gi = c.gOff + li;
  GameRegs<BB> s;
  s.black = (BB)root.black[gi]; s.white = (BB)root.white[gi]; s.h0 = root.hash0[gi]; s.h1 = root.hash1[gi];
  s.id = root.gameId[gi]; s.misc = root.misc[gi];
  uint8_t* treeBase = t.nodes + (size_t)gi * c.maxNodes * c.nodeStride;
  const int count = t.nodeCount[gi];
  int kind = 0, depth = 0;
  double leafVal = 0.0;
  const bool rootFinished = flagsOf(s.misc) & 1;
  if(rootFinished) {
    kind = 0;
  } else if(count == 0) {
    kind = 4;
  } else if(NodeRef{treeBase, c.P, c.polOff}.N() >= c.maxVisits) {
    kind = 0;
  } else {
    int node = 0;
    while(true) {
      NodeRef nd{treeBase + (size_t)node * c.nodeStride, c.P, c.polOff};
      const int pla = nd.nextPla();
      const double parentUtility = __ddiv_rn(nd.W(), (double)nd.N());
      const double* eW = nd.edgeW(); const float* pol = nd.policy(); const int* ch = nd.child(); const int* eN = nd.edgeN(); const uint8_t* ord = nd.order();
      double total = 0.0, mass = 0.0;
      for(int pos = lane; pos < c.P; pos += 32)
        if(ch[pos] != -1) { total = __dadd_rn(total, (double)eN[pos]); mass = __dadd_rn(mass, (double)pol[pos]); }
      total = warpSumD(total);
      mass = warpSumD(mass);
      const double red = __dmul_rn(depth == 0 ? c.rootFpuRed : c.fpuRed, __dsqrt_rn(mass));
      const double fpu = pla == 2 ? __dsub_rn(parentUtility, red) : __dadd_rn(parentUtility, red);
      const double scale = __dmul_rn(c.cpuct, __dsqrt_rn(__dadd_rn(total, 0.01)));
      // best existing child (value, then earliest created) and best unvisited move (prior, then lowest index)
      double bestVal = 0.0; int bestOrd = 1 << 20, bestPos = -1;
      float newP = -1.0f; int newPos = -1;
      for(int pos = lane; pos < c.P; pos += 32) {
        const float p = pol[pos];
        if(ch[pos] != -1) {
          const double n = (double)eN[pos];
          const double q = __ddiv_rn(eW[pos], n);
          const double val = __dadd_rn(__ddiv_rn(__dmul_rn(scale, (double)p), __dadd_rn(1.0, n)), pla == 2 ? q : -q);
          const int o = ord[pos];
          if(bestPos < 0 || val > bestVal || (val == bestVal && o < bestOrd)) { bestVal = val; bestOrd = o; bestPos = pos; }
        } else if(p >= 0.0f) {
          if(p > newP) { newP = p; newPos = pos; }   // ascending pos within a lane: ties keep the lowest index
        }
      }
      for(int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, bestVal, o);
        const int o2 = __shfl_xor_sync(0xffffffffu, bestOrd, o), p2 = __shfl_xor_sync(0xffffffffu, bestPos, o);
        if(p2 >= 0 && (bestPos < 0 || v2 > bestVal || (v2 == bestVal && o2 < bestOrd))) { bestVal = v2; bestOrd = o2; bestPos = p2; }
        const float np2 = __shfl_xor_sync(0xffffffffu, newP, o);
        const int npos2 = __shfl_xor_sync(0xffffffffu, newPos, o);
        if(npos2 >= 0 && (newPos < 0 || np2 > newP || (np2 == newP && npos2 < newPos))) { newP = np2; newPos = npos2; }
      }
      bool takeNew = false;
      if(newPos >= 0) {
        const double valNew = __dadd_rn(__ddiv_rn(__dmul_rn(scale, (double)newP), 1.0), pla == 2 ? fpu : -fpu);
        takeNew = bestPos < 0 || valNew > bestVal;
      }
      const int pos = takeNew ? newPos : bestPos;
      if(pos < 0) { kind = 0; break; }   // no legal move recorded: cannot happen for a non-terminal node
      if(lane == 0) { t.pathNode[(size_t)gi * MAX_PATH + depth] = node; t.pathPos[(size_t)gi * MAX_PATH + depth] = (uint8_t)pos; }
      depth++;
      if(!takeNew) {
        const int cc = ch[pos];
        if(cc <= -2) { kind = 3; leafVal = terminalValue(-2 - cc); break; }
        applyMoveLight(dm, s, pos, zob);
        node = cc;
        continue;
      }
      BB L[4]; bool illegal;
      stepGame(dm, g, s, pos, true, zob, L, illegal);
      const int fl = flagsOf(s.misc);
      if(fl & 1) { kind = 2; leafVal = terminalValue((fl >> 1) & 3); }
      else kind = 1;
      break;
    }
  }
  if(lane == 0) {
    if(kind != 0) t.active[c.half] = 1;
    t.leafKind[gi] = kind; t.pathLen[gi] = depth; t.leafValue[gi] = leafVal;
    t.leafNextPla[gi] = (flagsOf(s.misc) >> 3) & 3;
    bool needsNet = kind == 1 || kind == 4;
    if(t.cacheMask) {
      t.leafCache[gi] = -2;
      if(needsNet) {
        uint64_t ck0, ck1;
        nnCacheKey(g, s.h0, s.h1, s.misc, ck0, ck1);
        if(nnCacheLookup(c, t, gi, ck0, ck1) && c.compact) needsNet = false;   // the evaluation is already there: no row
      }
    }
    // the position handed to the evaluator.  Compact mode: only leaves that need the net take a row, in arrival order
    // (an evaluation does not depend on its row); otherwise row = game and idle rows are evaluated and ignored.
    if(needsNet || !c.compact) {
      const int slot = c.compact ? atomicAdd(t.evalCount + c.half, 1) : li;
      t.leafSlot[gi] = slot;
      leaf.black[slot] = (uint64_t)s.black; leaf.white[slot] = (uint64_t)s.white; leaf.hash0[slot] = s.h0; leaf.hash1[slot] = s.h1;
      leaf.gameId[slot] = s.id; leaf.misc[slot] = s.misc;
      if(c.randomSym) {   // position-keyed, so a position is evaluated under the same symmetry whichever game or path reaches it
        const int np = (flagsOf(s.misc) >> 3) & 3;
        leafSym[slot] = (int8_t)(splitmix64(c.seed ^ s.h0 ^ g.playerHash[np][0] ^ SYM_SALT) & 7);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// deterministic stand-in evaluator (integer hash of the situation -> exact fp32 policy / value), used to test the
// search logic bit for bit against the oracle; the product path uses the net
// ---------------------------------------------------------------------------------------------
constexpr uint64_t ROOTSYM_SALT = 0x7007575E5A17ULL;
__device__ __forceinline__ double nnWeightOf(const SearchCfg& c, float shorttermWinlossError);   // defined with the graph-mode helpers below
__global__ void k_hash_eval(int n, int P, int LW, const uint32_t* __restrict__ legal, const uint64_t* __restrict__ sitHash,
                            float* __restrict__ policy, float* __restrict__ winLoss, float* __restrict__ misc, const int8_t* __restrict__ forcedSym) {
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(gi >= n) return;
  uint64_t h0 = sitHash[2 * (size_t)gi], h1 = sitHash[2 * (size_t)gi + 1];
  if(forcedSym) {   // the root's rootNumSymmetriesToSample evaluations: the symmetry is part of the hash, so a wrong choice shows
    h0 ^= splitmix64(ROOTSYM_SALT + (uint64_t)forcedSym[gi]); h1 ^= splitmix64(ROOTSYM_SALT * 3 + (uint64_t)forcedSym[gi]);
  }
  int sum = 0;
  for(int pos = lane; pos < P; pos += 32) {
    const bool ok = (legal[(size_t)gi * LW + (pos >> 5)] >> (pos & 31)) & 1u;
    if(ok) sum += 1 + (int)((splitmix64(h0 ^ ((uint64_t)(pos + 1) * PHI)) >> 20) & 255);
  }
  for(int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for(int pos = lane; pos < P; pos += 32) {
    const bool ok = (legal[(size_t)gi * LW + (pos >> 5)] >> (pos & 31)) & 1u;
    const int w = 1 + (int)((splitmix64(h0 ^ ((uint64_t)(pos + 1) * PHI)) >> 20) & 255);
    policy[(size_t)gi * P + pos] = ok ? __fdiv_rn((float)w, (float)sum) : -1.0f;
  }
  if(lane == 0) {
    const uint64_t r = splitmix64(h1);
    winLoss[2 * (size_t)gi] = (float)(r & 0xFFFF) * (1.0f / 131072.0f);
    winLoss[2 * (size_t)gi + 1] = (float)((r >> 16) & 0xFFFF) * (1.0f / 131072.0f);
    if(misc) { misc[2 * (size_t)gi] = 0.f; misc[2 * (size_t)gi + 1] = (float)((r >> 32) & 0xFFFF) * (1.0f / 131072.0f); }   // shorttermWinlossError in [0, 0.5)
  }
}

// ---------------------------------------------------------------------------------------------
// rootNumSymmetriesToSample (searchnnhelpers.cpp:67-83; for a root kept by tree re-use maybeRecomputeExistingNNOutput :133-174): at
// the start of a search every root is evaluated under N distinct symmetries -- a partial Fisher-Yates shuffle of 0..7 drawn from a
// counter stream keyed by (seed, game id, ply) -- and the post-processed outputs are averaged (NNOutput's averaging constructor,
// nninputs.cpp:95-170: float sums in order, then / N).  Pass i: k_rootsym_prepare writes every unfinished root with its i-th
// symmetry into the evaluation batch (row = game), the batch is evaluated, k_rootsym_accumulate adds the outputs up and, after
// the last pass, installs the average: a fresh root becomes node 0 with its first visit, a kept root gets the new priors and
// evaluation with its statistics untouched until the next re-computation (isReInit).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int rootSymOf(const SearchCfg& c, uint64_t gameId, int ply, int pass) {
  int idx[8] = {0, 1, 2, 3, 4, 5, 6, 7};
  uint64_t st = splitmix64(c.seed ^ (gameId * PHI) ^ (uint64_t)ply ^ ROOTSYM_SALT);
  for(int i = 0; i <= pass; i++) {
    st += PHI;
    const int j = i + (int)(splitmix64(st) % (uint64_t)(8 - i));
    const int tmp = idx[i]; idx[i] = idx[j]; idx[j] = tmp;
  }
  return idx[pass];
}
__global__ void __launch_bounds__(128) k_rootsym_prepare(const Geom g, const SearchCfg c, State root, State leaf, TreeMem t, int8_t* __restrict__ leafSym, int pass) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if(li >= c.gCnt) return;
  const int gi = c.gOff + li;
  // row li = the root of game gi (finished games too: their row is evaluated and ignored)
  leaf.black[li] = root.black[gi]; leaf.white[li] = root.white[gi]; leaf.hash0[li] = root.hash0[gi]; leaf.hash1[li] = root.hash1[gi];
  leaf.gameId[li] = root.gameId[gi]; leaf.misc[li] = root.misc[gi];
  leafSym[li] = (int8_t)rootSymOf(c, root.gameId[gi], numTurnsOf(root.misc[gi]), pass);
  if(pass == 0 && c.compact) { if(li == 0) t.evalCount[c.half] = c.gCnt; }   // a full batch
}
__global__ void __launch_bounds__(128) k_rootsym_accumulate(const SearchCfg c, TreeMem t, State root, const float* __restrict__ policy, const float* __restrict__ winLoss,
                                                            const float* __restrict__ misc, float* __restrict__ accPolicy, float* __restrict__ accScalars, int pass) {
  const int li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(li >= c.gCnt) return;
  const int gi = c.gOff + li;
  if(flagsOf(root.misc[gi]) & 1) return;
  if(t.nodeCount[gi] > 0 && NodeRef{t.nodes + (size_t)gi * c.maxNodes * c.nodeStride, c.P, c.polOff}.noised()) return;   // this root has had its evaluations
  const size_t row = (size_t)li;
  float* ap = accPolicy + (size_t)gi * c.P;
  float* as = accScalars + (size_t)gi * 4;
  const bool last = pass == c.rootSyms - 1;
  const float fl = (float)c.rootSyms;
  for(int pos = lane; pos < c.P; pos += 32) {
    float v = __fadd_rn(pass == 0 ? 0.0f : ap[pos], policy[row * c.P + pos]);
    if(last) v = __fdiv_rn(v, fl);
    ap[pos] = v;
  }
  float aw = 0.f, al = 0.f, ae = 0.f;
  if(lane == 0) {
    aw = __fadd_rn(pass == 0 ? 0.0f : as[0], winLoss[2 * row]); al = __fadd_rn(pass == 0 ? 0.0f : as[1], winLoss[2 * row + 1]);
    ae = __fadd_rn(pass == 0 ? 0.0f : as[2], misc ? misc[2 * row + 1] : 0.0f);
    if(last) { aw = __fdiv_rn(aw, fl); al = __fdiv_rn(al, fl); ae = __fdiv_rn(ae, fl); }
    as[0] = aw; as[1] = al; as[2] = ae;
    atomicAdd(&t.stats[1], 1ULL);
  }
  if(!last) return;
  __syncwarp();
  NodeRef nd{t.nodes + (size_t)gi * c.maxNodes * c.nodeStride, c.P, c.polOff};
  const bool fresh = t.nodeCount[gi] == 0;
  for(int pos = lane; pos < c.P; pos += 32) {
    nd.policy()[pos] = ap[pos];
    if(fresh) { nd.child()[pos] = -1; nd.edgeN()[pos] = 0; nd.order()[pos] = 0; }
  }
  if(lane == 0) {
    const double v = __dsub_rn((double)aw, (double)al);
    const double w0 = nnWeightOf(c, ae);
    if(fresh) {
      nd.N() = 1; nd.numChildren() = 0; nd.nextPla() = (flagsOf(root.misc[gi]) >> 3) & 3; nd.biasEntry() = -1;
      nd.weightSum() = w0; nd.utilityAvg() = v; nd.nnUtility() = v; nd.lastDelta() = 0.0; nd.lastWeight() = 0.0;
      nd.utilitySqAvg() = __dmul_rn(v, v); nd.weightSqSum() = __dmul_rn(w0, w0); nd.nnWeight() = w0;
      nd.depth() = numTurnsOf(root.misc[gi]); nd.key()[0] = 0; nd.key()[1] = 0;
      t.nodeCount[gi] = 1;
      atomicAdd(&t.stats[0], 1ULL);
    } else { nd.nnUtility() = v; nd.nnWeight() = w0; }
    nd.noised() = 0;
  }
}

// ---------------------------------------------------------------------------------------------
// expand + backup: one warp per game
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_expand_backup(const SearchCfg c, TreeMem t, const float* __restrict__ policy, const float* __restrict__ winLoss) {
  const int li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(li >= c.gCnt) return;
  const int gi = c.gOff + li;
  const int kind = t.leafKind[gi];
  if(kind == 0) return;
  bool cached = false;
  uint8_t* treeBase = t.nodes + (size_t)gi * c.maxNodes * c.nodeStride;
  const int depth = t.pathLen[gi];
  double v = t.leafValue[gi];
  int newIdx = -1;
  if(kind == 1 || kind == 4) {
    const int ce = t.cacheMask ? t.leafCache[gi] : -2;                     // >= 0: the evaluation comes from the NN cache
    const size_t row = (size_t)t.leafSlot[ce == -3 ? t.leafOwner[gi] : gi];   // -3: the row of the game that evaluates the same position
    const float* polSrc = ce >= 0 ? t.cachePolicy + (size_t)ce * c.P : policy + row * c.P;
    const float* wlSrc = ce >= 0 ? t.cacheScalars + (size_t)ce * 4 : winLoss + 2 * row;
    cached = ce >= 0 || ce == -3;
    v = __dsub_rn((double)wlSrc[0], (double)wlSrc[1]);   // white-positive utility of the evaluation
    newIdx = t.nodeCount[gi];
    if(newIdx >= c.maxNodes) return;   // cannot happen: one new node per visit, maxNodes == maxVisits
    NodeRef nd{treeBase + (size_t)newIdx * c.nodeStride, c.P, c.polOff};
    for(int pos = lane; pos < c.P; pos += 32) {
      nd.edgeW()[pos] = 0.0; nd.policy()[pos] = polSrc[pos]; nd.child()[pos] = -1; nd.edgeN()[pos] = 0; nd.order()[pos] = 0;
    }
    if(ce == -1) nnCacheInsert(c, t, gi, lane, polSrc, wlSrc, nullptr);
    if(lane == 0) { nd.N() = 1; nd.numChildren() = 0; nd.nextPla() = t.leafNextPla[gi]; nd.W() = v; t.nodeCount[gi] = newIdx + 1; }
  }
  if(lane != 0) return;
  for(int d = 0; d < depth; d++) {
    NodeRef nd{treeBase + (size_t)t.pathNode[(size_t)gi * MAX_PATH + d] * c.nodeStride, c.P, c.polOff};
    const int pos = t.pathPos[(size_t)gi * MAX_PATH + d];
    if(d == depth - 1 && (kind == 1 || kind == 2)) {
      nd.child()[pos] = kind == 1 ? newIdx : (v > 0.0 ? -4 : v < 0.0 ? -3 : -2);
      nd.order()[pos] = (uint8_t)nd.numChildren();
      nd.numChildren() = nd.numChildren() + 1;
    }
    nd.N() = nd.N() + 1;
    nd.W() = __dadd_rn(nd.W(), v);
    nd.edgeN()[pos] = nd.edgeN()[pos] + 1;
    nd.edgeW()[pos] = __dadd_rn(nd.edgeW()[pos], v);
  }
  atomicAdd(&t.stats[0], 1ULL);
  if(kind == 1 || kind == 4) atomicAdd(&t.stats[cached ? 10 : 1], 1ULL); else atomicAdd(&t.stats[2], 1ULL);
}

// =============================================================================================
// graph mode: graph search (transpositions) and subtree value bias -- BASELINE config 4
// =============================================================================================
// Reference: Search::allocateOrFindNode (cpp/search/search.cpp:704-757), playoutDescend / maybeCatchUpEdgeVisits (:935-1207),
// selection on the children's NODE statistics with getChildWeight (cpp/search/searchexplorehelpers.cpp:92-127, 323-451,
// cpp/search/searchnode.h:59-65), addLeafValue / recomputeNodeStats with the bias table (cpp/search/searchupdatehelpers.cpp:12-76,
// 151-326), SubtreeValueBiasTable::get (cpp/search/subtreevaluebiastable.cpp:61-78).  Canonical semantics on top of the tree
// mode's (DESIGN.md ledger rows L, M):
//  * a node keeps (visits, weightSum, utilityAvg, its own evaluation nnUtility); an edge keeps its visit count; the weight of
//    a child seen from a parent is weightSum * (edgeVisits / max(visits, 1)); a visit re-computes every node on the path,
//    bottom up, from its children: utilityAvg = (sum_i w_i u_i + own) / (sum_i w_i + 1) (valueWeightExponent 0, unit weights).
//  * transposition key = getSitHash(next player) ^ mix(last move): the state that decides legality (the literal GraphHash
//    chains the previous hash after every non-pass move, graphhash.cpp:14-29, and would never transpose in a game without
//    passes).  Finished positions are not shared: terminal children stay per edge (equivalent, their statistics are constants).
//  * choosing a child whose edge has fewer visits than the node behind it adds the edge visit without descending
//    (maybeCatchUpEdgeVisits); a new edge that transposes to an existing node does exactly that on its first visit.
//  * bias entry of a node = per-search table keyed by (player who moved, previous move, move, colours of the 5x5 window
//    around the move on the board before it); on every re-computation the node moves its contribution
//    (childrenUtility - nnUtility) * W^exponent to the entry and adds factor * deltaSum / weightSum to its own evaluation.
//    The root has no entry.  W^exponent: sqrt for 0.5, identity for 1, otherwise detPow (IEEE basic operations only, the
//    oracle computes the same bits).

__device__ __forceinline__ double detLog(double x) {
  long long b = __double_as_longlong(x);
  int k = (int)((b >> 52) & 0x7ff) - 1022;
  b = (b & 0x800fffffffffffffLL) | 0x3fe0000000000000LL;
  double m = __longlong_as_double(b);
  if(m < 0.70710678118654752) { m = __dmul_rn(m, 2.0); k -= 1; }
  const double z = __ddiv_rn(__dsub_rn(m, 1.0), __dadd_rn(m, 1.0)), z2 = __dmul_rn(z, z);
  double s = __ddiv_rn(1.0, 27.0);
  for(int n = 25; n >= 1; n -= 2) s = __dadd_rn(__dmul_rn(s, z2), __ddiv_rn(1.0, (double)n));
  return __dadd_rn(__dmul_rn((double)k, 0.69314718055994531), __dmul_rn(__dmul_rn(2.0, z), s));
}
__device__ __forceinline__ double detExp(double y) {
  if(y < -700.0) return 0.0;   // below the normal range of the 2^n scaling
  const double n = rint(__dmul_rn(y, 1.4426950408889634));
  const double r = __dsub_rn(__dsub_rn(y, __dmul_rn(n, 0.693147180369123816490)), __dmul_rn(n, 1.90821492927058770002e-10));
  double s = __ddiv_rn(1.0, 6227020800.0);
  const double inv[13] = {1.0, 1.0, __ddiv_rn(1.0, 2.0), __ddiv_rn(1.0, 6.0), __ddiv_rn(1.0, 24.0), __ddiv_rn(1.0, 120.0), __ddiv_rn(1.0, 720.0),
                          __ddiv_rn(1.0, 5040.0), __ddiv_rn(1.0, 40320.0), __ddiv_rn(1.0, 362880.0), __ddiv_rn(1.0, 3628800.0),
                          __ddiv_rn(1.0, 39916800.0), __ddiv_rn(1.0, 479001600.0)};
#pragma unroll
  for(int i = 12; i >= 0; i--) s = __dadd_rn(__dmul_rn(s, r), inv[i]);
  const double sc = __longlong_as_double((long long)((int)n + 1023) << 52);
  return __dmul_rn(s, sc);
}
__device__ __forceinline__ double biasPow(double x, double e) {
  if(e == 0.5) return __dsqrt_rn(x);
  if(e == 1.0) return x;
  return detExp(__dmul_rn(e, detLog(x)));
}

// statistics of the child behind edge `pos` of `nd`: (visits, weightSum, utilityAvg); terminal children are constants
__device__ __forceinline__ void childStats(const SearchCfg& c, uint8_t* treeBase, int cc, int e, int& cv, double& cw, double& cu) {
  if(cc >= 0) {
    NodeRef ch{treeBase + (size_t)cc * c.nodeStride, c.P, c.polOff};
    cv = ch.N(); cw = ch.weightSum(); cu = ch.utilityAvg();
  } else { cv = e; cw = (double)e; cu = terminalValue(-2 - cc); }
}
__device__ __forceinline__ void childStatsSq(const SearchCfg& c, uint8_t* treeBase, int cc, int e, int& cv, double& cw, double& cu, double& cusq, double& cwsq) {
  if(cc >= 0) {
    NodeRef ch{treeBase + (size_t)cc * c.nodeStride, c.P, c.polOff};
    cv = ch.N(); cw = ch.weightSum(); cu = ch.utilityAvg(); cusq = ch.utilitySqAvg(); cwsq = ch.weightSqSum();
  } else { cv = e; cw = (double)e; cu = terminalValue(-2 - cc); cusq = __dmul_rn(cu, cu); cwsq = (double)e; }
}
// computeWeightFromNNOutput (searchupdatehelpers.cpp:91-113) with Coffee's outputs (no score term)
__device__ __forceinline__ double nnWeightOf(const SearchCfg& c, float shorttermWinlossError);
// weightSum * edgeVisits / max(visits, 1) (searchnode.h:59-62).  An edge that carries all of its child's visits -- every edge of a
// tree -- has ratio exactly 1.0 and cw * 1.0 == cw, so the division is skipped without changing a bit.
__device__ __forceinline__ double childWeightOf(double cw, int e, int cv) { return e == max(cv, 1) ? cw : __dmul_rn(cw, __ddiv_rn((double)e, (double)max(cv, 1))); }
// A per-position array held by a warp as reg[m] = array[lane + 32 m] (m < 4): the value at position `pos`, for any pos per lane
template <class T>
__device__ __forceinline__ T gather4(const T (&reg)[4], int pos) {
  const int src = pos & 31, slot = pos >> 5;
  const T a = __shfl_sync(0xffffffffu, reg[0], src), b = __shfl_sync(0xffffffffu, reg[1], src);
  const T c = __shfl_sync(0xffffffffu, reg[2], src), d = __shfl_sync(0xffffffffu, reg[3], src);
  return slot == 0 ? a : slot == 1 ? b : slot == 2 ? c : d;
}
template <class D> struct PolicySlots { static constexpr int N = (4 * KC_MAX_DEVICE_LEN * KC_MAX_DEVICE_LEN + 31) / 32; };   // children a lane can hold
template <> struct PolicySlots<StaticDims<5, 5, 4>> { static constexpr int N = 4; };

template <class D, bool PRE>
__global__ void __launch_bounds__(128, 8) k_select_graph(const Geom g, const SearchCfg c, State root, State leaf, TreeMem t, const uint64_t* __restrict__ zob, int8_t* __restrict__ leafSym) {
  const D dm(g);
  using BB = typename D::BB;
  const int li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(li >= c.gCnt) return;
  const int gi = c.gOff + li;
  GameRegs<BB> s;
  s.black = (BB)root.black[gi]; s.white = (BB)root.white[gi]; s.h0 = root.hash0[gi]; s.h1 = root.hash1[gi];
  s.id = root.gameId[gi]; s.misc = root.misc[gi];
  uint8_t* treeBase = t.nodes + (size_t)gi * c.maxNodes * c.nodeStride;
  const int count = t.nodeCount[gi];
  int kind = 0, depth = 0, target = -1;
  double leafVal = 0.0;
  uint64_t key0 = 0, key1 = 0, bk0 = 0, bk1 = 0;
  if(flagsOf(s.misc) & 1) kind = 0;
  else if(count == 0) kind = 4;
  else if(NodeRef{treeBase, c.P, c.polOff}.N() >= c.maxVisits) kind = 0;
  else {
    int node = 0;
    while(true) {
      NodeRef nd{treeBase + (size_t)node * c.nodeStride, c.P, c.polOff};
      const int pla = nd.nextPla();
      const double parentUtility = nd.utilityAvg();
      const float* pol = nd.policy(); const int* ch = nd.child(); const int* eN = nd.edgeN(); const uint8_t* lst = nd.list();
      // the children, k-th created child on lane k mod 32: statistics read once and kept in registers for both passes
      constexpr int SL = PolicySlots<D>::N;
      const int nc = nd.numChildren();
      double wv[SL], cuv[SL]; float pv[SL]; int posv[SL];
      double total = 0.0, mass = 0.0;
      // SL == 4 (policy size <= 128): the per-move arrays are fetched whole, coalesced and independent of the child list, so a level costs
      // two dependent memory round trips (node, then the children's headers) instead of three; a child's entries come from the lane
      // that holds its position (shuffles)
      float polR[4]; int chR[4], eNR[4];
      if constexpr(SL == 4 && PRE) {
#pragma unroll
        for(int m = 0; m < 4; m++) {
          const int pos = lane + 32 * m;
          const bool in = pos < c.P;
          polR[m] = in ? pol[pos] : -1.0f; chR[m] = in ? ch[pos] : -1; eNR[m] = in ? eN[pos] : 0;
        }
      }
#pragma unroll
      for(int m = 0; m < SL; m++) {
        const int k = lane + 32 * m;
        posv[m] = -1;
        const bool anyK = 32 * m < nc;   // warp-uniform
        int pos = 0, e = 0, cc = -1; float p = 0.f;
        if constexpr(SL == 4 && PRE) {
          if(anyK) {
            pos = k < nc ? lst[k] : 0;
            cc = gather4(chR, pos); e = gather4(eNR, pos); p = gather4(polR, pos);
          }
        } else if(k < nc) { pos = lst[k]; cc = ch[pos]; e = eN[pos]; p = pol[pos]; }
        if(k < nc) {
          int cv; double cw, cu;
          childStats(c, treeBase, cc, e, cv, cw, cu);
          wv[m] = childWeightOf(cw, e, cv); cuv[m] = cu; pv[m] = p; posv[m] = pos;
          total = __dadd_rn(total, wv[m]);
          mass = __dadd_rn(mass, (double)pv[m]);
        }
      }
      total = warpSumD(total);
      mass = warpSumD(mass);
      double parentUtilityForFPU = parentUtility;
      if(c.fpuPW) {   // fpuParentWeightByVisitedPolicy (searchexplorehelpers.cpp:279-282)
        const double raised = mass <= 0.0 ? 0.0 : c.fpuPWPow == 1.0 ? mass : c.fpuPWPow == 2.0 ? __dmul_rn(mass, mass) : detExp(__dmul_rn(c.fpuPWPow, detLog(mass)));
        const double avgWeight = fmin(1.0, raised);
        parentUtilityForFPU = __dadd_rn(__dmul_rn(avgWeight, parentUtility), __dmul_rn(__dsub_rn(1.0, avgWeight), nd.nnUtility()));
      }
      const double red = __dmul_rn(depth == 0 ? c.rootFpuRed : c.fpuRed, __dsqrt_rn(mass));
      const double fpu = pla == 2 ? __dsub_rn(parentUtilityForFPU, red) : __dadd_rn(parentUtilityForFPU, red);
      const double scale = __dmul_rn(c.cpuct, __dsqrt_rn(__dadd_rn(total, 0.01)));
      double bestVal = 0.0; int bestOrd = 1 << 20, bestPos = -1;
      float newP = -1.0f; int newPos = -1;
#pragma unroll
      for(int m = 0; m < SL; m++) {
        if(posv[m] < 0) continue;
        const double w = wv[m];
        const float p = pv[m];
        double val = __dadd_rn(__ddiv_rn(__dmul_rn(scale, (double)p), __dadd_rn(1.0, w)), pla == 2 ? cuv[m] : -cuv[m]);
        // rootDesiredPerChildVisitsCoeff (searchexplorehelpers.cpp:150-155)
        if(depth == 0 && c.rootDesired > 0.0 && p > 0.0f && w < __dsqrt_rn(__dmul_rn(__dmul_rn((double)p, total), c.rootDesired))) val = 1e20;
        const int o = lane + 32 * m;   // creation order
        if(bestPos < 0 || val > bestVal || (val == bestVal && o < bestOrd)) { bestVal = val; bestOrd = o; bestPos = posv[m]; }
      }
      if constexpr(SL == 4 && PRE) {
#pragma unroll
        for(int m = 0; m < 4; m++)
          if(polR[m] >= 0.0f && chR[m] == -1 && polR[m] > newP) { newP = polR[m]; newPos = lane + 32 * m; }
      } else
      for(int pos = lane; pos < c.P; pos += 32) {
        const float p = pol[pos];
        if(p >= 0.0f && ch[pos] == -1 && p > newP) { newP = p; newPos = pos; }   // ascending pos within a lane: ties keep the lowest index
      }
      for(int o = 16; o > 0; o >>= 1) {
        const double v2 = __shfl_xor_sync(0xffffffffu, bestVal, o);
        const int o2 = __shfl_xor_sync(0xffffffffu, bestOrd, o), p2 = __shfl_xor_sync(0xffffffffu, bestPos, o);
        if(p2 >= 0 && (bestPos < 0 || v2 > bestVal || (v2 == bestVal && o2 < bestOrd))) { bestVal = v2; bestOrd = o2; bestPos = p2; }
        const float np2 = __shfl_xor_sync(0xffffffffu, newP, o);
        const int npos2 = __shfl_xor_sync(0xffffffffu, newPos, o);
        if(npos2 >= 0 && (newPos < 0 || np2 > newP || (np2 == newP && npos2 < newPos))) { newP = np2; newPos = npos2; }
      }
      bool takeNew = false;
      if(newPos >= 0) {
        const double valNew = __dadd_rn(__ddiv_rn(__dmul_rn(scale, (double)newP), 1.0), pla == 2 ? fpu : -fpu);
        takeNew = bestPos < 0 || valNew > bestVal;
      }
      const int pos = takeNew ? newPos : bestPos;
      if(pos < 0) { kind = 0; break; }
      if(lane == 0) { t.pathNode[(size_t)gi * MAX_PATH + depth] = node; t.pathPos[(size_t)gi * MAX_PATH + depth] = (uint8_t)pos; }
      depth++;
      if(!takeNew) {
        const int cc = ch[pos];
        if(cc <= -2) { kind = 3; leafVal = terminalValue(-2 - cc); break; }
        if(eN[pos] < NodeRef{treeBase + (size_t)cc * c.nodeStride, c.P, c.polOff}.N()) { kind = 5; break; }   // catch up, no descent
        applyMoveLight(dm, s, pos, zob);
        node = cc;
        continue;
      }
      const BB beforeB = s.black, beforeW = s.white;
      const uint64_t beforeMisc = s.misc;
      BB L[4]; bool illegal;
      stepGame(dm, g, s, pos, true, zob, L, illegal);
      const int fl = flagsOf(s.misc);
      if(fl & 1) { kind = 2; leafVal = terminalValue((fl >> 1) & 3); break; }
      const int nextPla = (fl >> 3) & 3;
      const uint64_t lm = (uint64_t)(pos + 1);
      key0 = s.h0 ^ g.playerHash[nextPla][0] ^ splitmix64(lm);
      key1 = s.h1 ^ g.playerHash[nextPla][1] ^ splitmix64(lm * PHI);
      if(c.useTable) {
        int found = -1;
        if(lane == 0) {
          const uint64_t* keys = t.tblKeys + (size_t)gi * c.tableCap * 2;
          const int* vals = t.tblVals + (size_t)gi * c.tableCap;
          int slot = (int)(key0 & (uint64_t)(c.tableCap - 1));
          while(true) {
            const int v = vals[slot];
            if(v == 0) break;
            if(keys[2 * slot] == key0 && keys[2 * slot + 1] == key1) { found = v - 1; break; }
            slot = (slot + 1) & (c.tableCap - 1);
          }
        }
        found = __shfl_sync(0xffffffffu, found, 0);
        if(found >= 0) { kind = 6; target = found; break; }
      }
      kind = 1;
      if(c.biasFactor != 0.0 && histPla(beforeMisc, 0) != 0) {
        const int HW = dm.HW(), cell = pos % HW, cx = cell % dm.W(), cy = cell / dm.W();
        uint64_t win = 0;
        for(int dy = -2; dy <= 2; dy++)
          for(int dx = -2; dx <= 2; dx++) {
            const int x = cx + dx, y = cy + dy;
            uint64_t code = 3;
            if(x >= 0 && y >= 0 && x < dm.W() && y < dm.H()) {
              const int bit = y * dm.stride() + x;
              code = ((beforeB >> bit) & 1) ? 1 : ((beforeW >> bit) & 1) ? 2 : 0;
            }
            win |= code << (2 * ((dy + 2) * 5 + (dx + 2)));
          }
        const uint64_t mover = (uint64_t)((flagsOf(beforeMisc) >> 3) & 3);
        bk0 = win | (mover << 50) | ((uint64_t)pos << 52);
        bk1 = (uint64_t)(lastDirOf(beforeMisc) * HW + histCell(beforeMisc, 0) + 1);
      }
      break;
    }
  }
  if(lane == 0) {
    if(kind != 0) t.active[c.half] = 1;
    t.leafKind[gi] = kind; t.pathLen[gi] = depth; t.leafValue[gi] = leafVal;
    t.leafNextPla[gi] = ((flagsOf(s.misc) >> 3) & 3) | (numTurnsOf(s.misc) << 8);   // + the position's depth (stones on the board)
    t.leafKey[2 * (size_t)gi] = key0; t.leafKey[2 * (size_t)gi + 1] = key1;
    t.leafBiasKey[2 * (size_t)gi] = bk0; t.leafBiasKey[2 * (size_t)gi + 1] = bk1;
    t.leafTarget[gi] = target;
    bool needsNet = kind == 1 || kind == 4;
    if(t.cacheMask) {
      t.leafCache[gi] = -2;
      if(needsNet) {
        uint64_t ck0, ck1;
        nnCacheKey(g, s.h0, s.h1, s.misc, ck0, ck1);
        if(nnCacheLookup(c, t, gi, ck0, ck1) && c.compact) needsNet = false;   // the evaluation is already there: no row
      }
    }
    if(needsNet || !c.compact) {
      const int slot = c.compact ? atomicAdd(t.evalCount + c.half, 1) : li;
      t.leafSlot[gi] = slot;
      leaf.black[slot] = (uint64_t)s.black; leaf.white[slot] = (uint64_t)s.white; leaf.hash0[slot] = s.h0; leaf.hash1[slot] = s.h1;
      leaf.gameId[slot] = s.id; leaf.misc[slot] = s.misc;
      if(c.randomSym) {   // position-keyed, so a position is evaluated under the same symmetry whichever game or path reaches it
        const int np = (flagsOf(s.misc) >> 3) & 3;
        leafSym[slot] = (int8_t)(splitmix64(c.seed ^ s.h0 ^ g.playerHash[np][0] ^ SYM_SALT) & 7);
      }
    }
  }
}

__device__ __forceinline__ double nnWeightOf(const SearchCfg& c, float shorttermWinlossError) {
  if(!c.useUncertainty) return 1.0;
  const double unc = (double)shorttermWinlossError;
  const double powered = c.uncExp == 1.0 ? unc : c.uncExp == 0.5 ? __dsqrt_rn(unc) : (unc <= 0.0 ? 0.0 : detExp(__dmul_rn(c.uncExp, detLog(unc))));
  const double baseline = __ddiv_rn(c.uncCoeff, c.uncMaxWeight);
  return __ddiv_rn(c.uncCoeff, __dadd_rn(powered, baseline));
}

// recomputeNodeStats (searchupdatehelpers.cpp:151-326) for one node by one warp; `inc` visits are added
// One warp; the k-th created child sits on lane k mod 32 (a node with at most 32 children -- nearly all -- runs the per-child code once),
// its statistics are read once and kept in registers.  SL = children a lane can hold.
template <int SL, bool PRE>
__device__ __forceinline__ void recomputeNode(const SearchCfg& c, double* biasValsOfGame, const double* __restrict__ tcdfTable, const double* __restrict__ stdevTab,
                                              uint8_t* treeBase, NodeRef nd, int lane, int inc, bool isRoot) {
  const int* ch = nd.child(); const int* eN = nd.edgeN(); const uint8_t* lst = nd.list();
  const int nc = nd.numChildren();
  double sumW = 0.0, sumWU = 0.0, maxW = 0.0;
  double nwl[SL], cuv[SL], cwv[SL];   // this lane's children: weight (then desired weight), utility, raw weightSum
  int ccv[SL], ev[SL];
  int chR[4], eNR[4];
  if constexpr(SL == 4 && PRE) {   // the per-move arrays whole and coalesced, independent of the child list (see k_select_graph)
#pragma unroll
    for(int m = 0; m < 4; m++) {
      const int pos = lane + 32 * m;
      const bool in = pos < c.P;
      chR[m] = in ? ch[pos] : -1; eNR[m] = in ? eN[pos] : 0;
    }
  }
#pragma unroll
  for(int m = 0; m < SL; m++) {
    const int k = lane + 32 * m;
    nwl[m] = 0.0; cuv[m] = 0.0; cwv[m] = 1.0; ccv[m] = -1; ev[m] = 0;
    int e = 0, cc = -1;
    if constexpr(SL == 4 && PRE) {
      if(32 * m < nc) {   // warp-uniform
        const int pos = k < nc ? lst[k] : 0;
        cc = gather4(chR, pos); e = gather4(eNR, pos);
      }
    } else if(k < nc) { const int pos = lst[k]; e = eN[pos]; cc = ch[pos]; }
    if(k < nc) {
      int cv; double cw, cu;
      childStats(c, treeBase, cc, e, cv, cw, cu);
      if(cv <= 0 || cw <= 0.0 || e <= 0) continue;
      const double w = childWeightOf(cw, e, cv);
      nwl[m] = w; cuv[m] = cu; cwv[m] = cw; ccv[m] = cc; ev[m] = e;
      maxW = fmax(maxW, w);
      sumW = __dadd_rn(sumW, w);
      sumWU = __dadd_rn(sumWU, __dmul_rn(w, cu));
    }
  }
  sumW = warpSumD(sumW);
  sumWU = warpSumD(sumWU);
  const double origW = sumW;   // origTotalChildWeight: what the subtree value bias weighs with
  if(c.noisePruning) {
    // pruneNoiseWeight (searchupdatehelpers.cpp:422-470): a sequential pass over the children in creation order.  Child k lives on lane
    // k mod 32; every lane follows the same recurrence on broadcast values and the owner keeps the new weight.
    int good = 0;
#pragma unroll
    for(int m = 0; m < SL; m++) good += nwl[m] != 0.0;
    for(int o = 16; o > 0; o >>= 1) good += __shfl_xor_sync(0xffffffffu, good, o);
    if(good > 1 && sumW > 0.00001) {
      const float* pol = nd.policy();
      const int pla = nd.nextPla();
      double uSum = 0.0, wSum = 0.0, pSum = 0.0;
#pragma unroll
      for(int m = 0; m < SL; m++) {
        const int kk = lane + 32 * m;
        const double myPol = (kk < nc && nwl[m] != 0.0) ? fmax(1e-30, (double)pol[lst[kk]]) : 0.0;
        const int cnt = min(32, nc - 32 * m);   // warp-uniform
        for(int j = 0; j < cnt; j++) {
          const double wk = __shfl_sync(0xffffffffu, nwl[m], j);
          if(wk == 0.0) continue;   // uniform
          const double cu = __shfl_sync(0xffffffffu, cuv[m], j);
          const double rawPolicy = __shfl_sync(0xffffffffu, myPol, j);
          const double utility = pla == 2 ? cu : -cu;
          double nwk = wk;
          if(wSum > 0 && pSum > 0) {
            const double gap = __dsub_rn(__ddiv_rn(uSum, wSum), utility);
            if(gap > 0) {
              const double lenient = __dmul_rn(2.0, __ddiv_rn(__dmul_rn(wSum, rawPolicy), pSum));
              if(wk > lenient) {
                double toSub = __dmul_rn(__dsub_rn(wk, lenient), __dsub_rn(1.0, detExp(-__ddiv_rn(gap, c.noisePruneScale))));
                if(toSub > c.noisePruneCap) toSub = c.noisePruneCap;
                nwk = __dsub_rn(wk, toSub);
              }
            }
          }
          if(lane == j) nwl[m] = nwk;
          uSum = __dadd_rn(uSum, __dmul_rn(utility, nwk)); wSum = __dadd_rn(wSum, nwk); pSum = __dadd_rn(pSum, rawPolicy);
        }
      }
      sumW = wSum;
      double part = 0.0;
#pragma unroll
      for(int m = 0; m < SL; m++) if(nwl[m] != 0.0) part = __dadd_rn(part, __dmul_rn(nwl[m], cuv[m]));
      sumWU = warpSumD(part);
    }
  }
  // at a noised root the children the move choice would prune / reduce lose the same weight here (:196-206)
  double amountToSubtract = 0.0, amountToPrune = 0.0;
  if(isRoot && c.rootNoise && !c.noisePruning) {
    for(int o = 16; o > 0; o >>= 1) maxW = fmax(maxW, __shfl_xor_sync(0xffffffffu, maxW, o));
    amountToSubtract = fmin(c.moveSubtract, __ddiv_rn(maxW, 64.0));
    amountToPrune = fmin(c.movePrune, __ddiv_rn(maxW, 64.0));
  }
  const bool reweigh = sumW > 0.0 && (c.vwExp != 0.0 || amountToSubtract > 0.0 || amountToPrune > 0.0);
  if(reweigh) {
    // downweightBadChildrenAndNormalizeWeight (searchupdatehelpers.cpp:330-417): prune / subtract, then (valueWeightExponent) a child
    // keeps weight * cdf(z)^exponent, z = its utility's distance from the siblings' weighted mean in standard errors, cdf = Student t
    // (3 degrees of freedom) from the interpolated 2,000-point table; the total child weight stays what it was
    const double simpleValue = __ddiv_rn(sumWU, sumW);
    const int pla = nd.nextPla();
    double totalNew = 0.0;
#pragma unroll
    for(int m = 0; m < SL; m++) {
      const double w = nwl[m];
      if(w == 0.0) continue;
      double x = w;
      if(x < amountToPrune) x = 0.0;
      else { x = __dsub_rn(x, amountToSubtract); if(x <= 0.0) x = 0.0; }
      if(x > 0.0 && c.vwExp != 0.0) {
        const double cu = cuv[m];
        // stdev = sqrt(1e-8 + 1 / (1.5 sqrt(w))): from the table when w is a whole number (the bits are the same: every operation is
        // correctly rounded on the host too)
        double stdev;
        const int wi = (int)w;
        if(w <= 1024.0 && (double)wi == w) stdev = stdevTab[wi];
        else stdev = __dsqrt_rn(__dadd_rn(0.00000001, __ddiv_rn(1.0, __dmul_rn(1.5, __dsqrt_rn(w)))));
        const double diff = pla == 2 ? __dsub_rn(cu, simpleValue) : __dsub_rn(simpleValue, cu);
        const double z = __ddiv_rn(diff, stdev);
        const double d = __ddiv_rn(__dmul_rn(1999.0, __dsub_rn(z, -50.0)), 100.0);
        double cdf;
        if(d <= 0) cdf = 0.0;
        else {
          const int idx = (int)d;
          if(idx >= 1999) cdf = 1.0;
          else cdf = __dadd_rn(tcdfTable[idx], __dmul_rn(__dsub_rn(d, (double)idx), __dsub_rn(tcdfTable[idx + 1], tcdfTable[idx])));
        }
        const double pr = __dadd_rn(cdf, 0.0001);
        const double raised = c.vwExp == 0.5 ? __dsqrt_rn(pr) : c.vwExp == 0.25 ? __dsqrt_rn(__dsqrt_rn(pr)) : c.vwExp == 1.0 ? pr : detExp(__dmul_rn(c.vwExp, detLog(pr)));
        x = __dmul_rn(x, raised);
      }
      nwl[m] = x;
      totalNew = __dadd_rn(totalNew, x);
    }
    totalNew = warpSumD(totalNew);
    const double factor = __ddiv_rn(sumW, totalNew);
#pragma unroll
    for(int m = 0; m < SL; m++) nwl[m] = __dmul_rn(nwl[m], factor);
  }
  double partU = 0.0, partUSq = 0.0, partWSq = 0.0;
#pragma unroll
  for(int m = 0; m < SL; m++)
    if(nwl[m] != 0.0) {
      double cusq, cwsq;
      const int cc = ccv[m];
      if(cc >= 0) {
        NodeRef chn{treeBase + (size_t)cc * c.nodeStride, c.P, c.polOff};
        cusq = chn.utilitySqAvg(); cwsq = chn.weightSqSum();
      } else { cusq = __dmul_rn(cuv[m], cuv[m]); cwsq = (double)ev[m]; }
      const double scaling = __ddiv_rn(nwl[m], cwv[m]);
      partU = __dadd_rn(partU, __dmul_rn(nwl[m], cuv[m]));
      partUSq = __dadd_rn(partUSq, __dmul_rn(nwl[m], cusq));
      partWSq = __dadd_rn(partWSq, __dmul_rn(__dmul_rn(scaling, scaling), cwsq));
    }
  partU = warpSumD(partU);
  if(reweigh) sumWU = partU;
  const double sumWUSq = warpSumD(partUSq), sumWSq = warpSumD(partWSq);
  if(lane == 0) {
    double utility = nd.nnUtility();
    const int be = nd.biasEntry();
    if(c.biasFactor != 0.0 && be >= 0) {
      double* E = biasValsOfGame + (size_t)be * 2;
      double ed = E[0], ew = E[1];
      if(sumW > 1e-10) {
        const double uc = __ddiv_rn(sumWU, sumW);
        const double bw = biasPow(origW, c.biasExp);
        const double ds = __dmul_rn(__dsub_rn(uc, nd.nnUtility()), bw);
        ed = __dadd_rn(ed, __dsub_rn(ds, nd.lastDelta()));
        ew = __dadd_rn(ew, __dsub_rn(bw, nd.lastWeight()));
        E[0] = ed; E[1] = ew;
        nd.lastDelta() = ds; nd.lastWeight() = bw;
      }
      if(ew > 0.001) utility = __dadd_rn(utility, __ddiv_rn(__dmul_rn(c.biasFactor, ed), ew));
    }
    const double w0 = nd.nnWeight();
    const double weightSum = __dadd_rn(sumW, w0);
    nd.utilityAvg() = __ddiv_rn(__dadd_rn(sumWU, __dmul_rn(utility, w0)), weightSum);
    nd.utilitySqAvg() = __ddiv_rn(__dadd_rn(sumWUSq, __dmul_rn(__dmul_rn(utility, utility), w0)), weightSum);
    nd.weightSqSum() = __dadd_rn(sumWSq, __dmul_rn(w0, w0));
    nd.weightSum() = weightSum;
    nd.N() = nd.N() + inc;
  }
  __syncwarp();
}

constexpr int MAX_POLICY_SLOTS_PER_LANE = (4 * KC_MAX_DEVICE_LEN * KC_MAX_DEVICE_LEN + 31) / 32;
template <int SL, bool PRE>
__global__ void __launch_bounds__(128, 8) k_expand_backup_graph(const SearchCfg c, TreeMem t, const float* __restrict__ policy, const float* __restrict__ winLoss,
                                                                const float* __restrict__ misc) {
  const int li = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(li >= c.gCnt) return;
  const int gi = c.gOff + li;
  const int kind = t.leafKind[gi];
  if(kind == 0) return;
  bool cached = false;
  float shortErr = 0.0f;
  uint8_t* treeBase = t.nodes + (size_t)gi * c.maxNodes * c.nodeStride;
  const int depth = t.pathLen[gi];
  double v = t.leafValue[gi];
  int newIdx = -1;
  if(kind == 1 || kind == 4) {
    const int ce = t.cacheMask ? t.leafCache[gi] : -2;                     // >= 0: the evaluation comes from the NN cache
    const size_t row = (size_t)t.leafSlot[ce == -3 ? t.leafOwner[gi] : gi];   // -3: the row of the game that evaluates the same position
    const float* polSrc = ce >= 0 ? t.cachePolicy + (size_t)ce * c.P : policy + row * c.P;
    const float* wlSrc = ce >= 0 ? t.cacheScalars + (size_t)ce * 4 : winLoss + 2 * row;
    const float* miscSrc = ce >= 0 ? t.cacheScalars + (size_t)ce * 4 + 2 : (misc ? misc + 2 * row : nullptr);
    cached = ce >= 0 || ce == -3;
    shortErr = miscSrc ? miscSrc[1] : 0.0f;
    v = __dsub_rn((double)wlSrc[0], (double)wlSrc[1]);
    newIdx = t.nodeCount[gi];
    if(newIdx >= c.maxNodes) return;   // cannot happen: at most one new node per visit
    NodeRef nd{treeBase + (size_t)newIdx * c.nodeStride, c.P, c.polOff};
    for(int pos = lane; pos < c.P; pos += 32) {
      nd.policy()[pos] = polSrc[pos]; nd.child()[pos] = -1; nd.edgeN()[pos] = 0; nd.order()[pos] = 0;
    }
    if(ce == -1) nnCacheInsert(c, t, gi, lane, polSrc, wlSrc, miscSrc);
    if(lane == 0) {
      int be = -1;
      double utility = v;
      const uint64_t bk0 = t.leafBiasKey[2 * (size_t)gi], bk1 = t.leafBiasKey[2 * (size_t)gi + 1];
      if(kind == 1 && bk0 != 0) {   // SubtreeValueBiasTable::get: find or create the entry
        uint64_t* keys = t.biasKeys + (size_t)gi * c.tableCap * 2;
        int slot = (int)(splitmix64(bk0 ^ splitmix64(bk1)) & (uint64_t)(c.tableCap - 1));
        while(true) {
          if(keys[2 * slot] == 0) { keys[2 * slot] = bk0; keys[2 * slot + 1] = bk1; break; }
          if(keys[2 * slot] == bk0 && keys[2 * slot + 1] == bk1) break;
          slot = (slot + 1) & (c.tableCap - 1);
        }
        be = slot;
        const double* E = t.biasVals + ((size_t)gi * c.tableCap + be) * 2;
        if(E[1] > 0.001) utility = __dadd_rn(utility, __ddiv_rn(__dmul_rn(c.biasFactor, E[0]), E[1]));   // addLeafValue :27-37
      }
      const double w0 = nnWeightOf(c, shortErr);
      nd.N() = 1; nd.numChildren() = 0; nd.nextPla() = t.leafNextPla[gi] & 0xff; nd.biasEntry() = be;
      nd.weightSum() = w0; nd.utilityAvg() = utility; nd.nnUtility() = v; nd.lastDelta() = 0.0; nd.lastWeight() = 0.0;
      nd.utilitySqAvg() = __dmul_rn(utility, utility); nd.weightSqSum() = __dmul_rn(w0, w0); nd.nnWeight() = w0;
      nd.depth() = t.leafNextPla[gi] >> 8; nd.noised() = 0;
      nd.key()[0] = kind == 1 ? t.leafKey[2 * (size_t)gi] : 0; nd.key()[1] = kind == 1 ? t.leafKey[2 * (size_t)gi + 1] : 0;
      t.nodeCount[gi] = newIdx + 1;
      if(kind == 1 && c.useTable) {
        uint64_t* keys = t.tblKeys + (size_t)gi * c.tableCap * 2;
        int* vals = t.tblVals + (size_t)gi * c.tableCap;
        const uint64_t key0 = t.leafKey[2 * (size_t)gi], key1 = t.leafKey[2 * (size_t)gi + 1];
        int slot = (int)(key0 & (uint64_t)(c.tableCap - 1));
        while(vals[slot] != 0) slot = (slot + 1) & (c.tableCap - 1);
        keys[2 * slot] = key0; keys[2 * slot + 1] = key1; vals[slot] = newIdx + 1;
      }
    }
  }
  __syncwarp();
  for(int d = depth - 1; d >= 0; d--) {
    NodeRef nd{treeBase + (size_t)t.pathNode[(size_t)gi * MAX_PATH + d] * c.nodeStride, c.P, c.polOff};
    const int pos = t.pathPos[(size_t)gi * MAX_PATH + d];
    if(lane == 0) {
      if(d == depth - 1 && (kind == 1 || kind == 2 || kind == 6)) {
        nd.child()[pos] = kind == 1 ? newIdx : kind == 6 ? t.leafTarget[gi] : (v > 0.0 ? -4 : v < 0.0 ? -3 : -2);
        nd.order()[pos] = (uint8_t)nd.numChildren();
        nd.list()[nd.numChildren()] = (uint8_t)pos;
        nd.numChildren() = nd.numChildren() + 1;
      }
      nd.edgeN()[pos] = nd.edgeN()[pos] + 1;
    }
    __syncwarp();
    recomputeNode<SL, PRE>(c, t.biasVals + (size_t)gi * c.tableCap * 2, t.tcdf, t.stdevTab, treeBase, nd, lane, 1, d == 0);
  }
  if(lane != 0) return;
  atomicAdd(&t.stats[0], 1ULL);
  if(kind == 1 || kind == 4) atomicAdd(&t.stats[cached ? 10 : 1], 1ULL);
  else if(kind == 2 || kind == 3) atomicAdd(&t.stats[2], 1ULL);
  else if(kind == 6) atomicAdd(&t.stats[8], 1ULL);
  else atomicAdd(&t.stats[9], 1ULL);
}

// Root policy temperature and shaped Dirichlet noise: Search::maybeAddPolicyNoiseAndTemp / addDirichletNoise /
// computeDirichletAlphaDistribution (cpp/search/searchhelpers.cpp:51-221) with Rand::nextGamma / nextGaussian / nextDouble
// (cpp/core/rand.cpp:335-363, rand.h:235-288).  Two things cannot be shared between a CPU and a GPU and are replaced: the
// generator is a counter-based splitmix64 stream keyed by (seed, game id, ply), and log / exp / pow are detLog / detExp; the
// oracle computes the same bits.  One thread per game, once per search, in place on the root's priors (the reference keeps
// a separate noisedPolicyProbs; nothing else reads the root's raw policy afterwards).
constexpr uint64_t NOISE_SALT = 0xD1A1C4137E5EEDULL;
struct DetRng {
  uint64_t s; bool hasG; double g;
  __device__ DetRng(uint64_t seed, uint64_t gameId, int ply) : s(splitmix64(seed ^ (gameId * PHI) ^ (uint64_t)ply ^ NOISE_SALT)), hasG(false), g(0.0) {}
  __device__ uint64_t nextU64() { s += PHI; return splitmix64(s); }
  __device__ double nextDouble() { return __dmul_rn((double)(nextU64() & ((1ULL << 53) - 1ULL)), 1.0 / 9007199254740992.0); }
  __device__ double nextGaussian() {
    if(hasG) { hasG = false; return g; }
    double v1, v2, q;
    do {
      v1 = __dsub_rn(__dmul_rn(nextDouble(), 2.0), 1.0);
      v2 = __dsub_rn(__dmul_rn(nextDouble(), 2.0), 1.0);
      q = __dadd_rn(__dmul_rn(v1, v1), __dmul_rn(v2, v2));
    } while(q >= 1.0 || q == 0.0);
    const double mult = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, detLog(q)), q));
    g = __dmul_rn(v2, mult); hasG = true;
    return __dmul_rn(v1, mult);
  }
  __device__ double gammaAbove1(double a) {   // Marsaglia-Tsang, a > 1
    const double d = __dsub_rn(a, 1.0 / 3.0);
    const double c = __ddiv_rn(1.0 / 3.0, __dsqrt_rn(d));
    while(true) {
      const double x = nextGaussian();
      const double vtmp = __dadd_rn(1.0, __dmul_rn(c, x));
      if(vtmp <= 0.0) continue;
      const double v = __dmul_rn(__dmul_rn(vtmp, vtmp), vtmp);
      const double u = nextDouble();
      const double xx = __dmul_rn(x, x);
      if(u < __dsub_rn(1.0, __dmul_rn(__dmul_rn(0.0331, xx), xx))) return __dmul_rn(d, v);
      if(u == 0.0 || detLog(u) < __dadd_rn(__dmul_rn(0.5, xx), __dmul_rn(d, __dadd_rn(__dsub_rn(1.0, v), detLog(v))))) return __dmul_rn(d, v);
    }
  }
  __device__ double nextGamma(double a) {
    if(a > 1.0) return gammaAbove1(a);
    const double r = gammaAbove1(__dadd_rn(a, 1.0));
    const double inva = __ddiv_rn(1.0, a);
    const double u = nextDouble();
    const double scale = u == 0.0 ? 0.0 : detExp(__dmul_rn(inva, detLog(u)));
    return __dmul_rn(r, scale);
  }
};
constexpr int MAX_POLICY = 4 * KC_MAX_DEVICE_LEN * KC_MAX_DEVICE_LEN;

__global__ void __launch_bounds__(64) k_root_noise(const SearchCfg c, TreeMem t, State root, int boardArea) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if(li >= c.gCnt) return;
  const int gi = c.gOff + li;
  if(t.nodeCount[gi] <= 0 || (flagsOf(root.misc[gi]) & 1)) return;
  NodeRef nd{t.nodes + (size_t)gi * c.maxNodes * c.nodeStride, c.P, c.polOff};
  if(nd.noised()) return;
  nd.noised() = 1;
  float* pol = nd.policy();
  const int P = c.P, turn = numTurnsOf(root.misc[gi]);
  const double tEarly = c.rootTempEarly > 0.0 ? c.rootTempEarly : 1.0, tLate = c.rootTemp > 0.0 ? c.rootTemp : 1.0;
  if(tEarly != 1.0 || tLate != 1.0) {
    const double halflife = c.tempHalflife > 0.0 ? c.tempHalflife : 19.0;
    const double halflives = __ddiv_rn(__dmul_rn(__ddiv_rn((double)turn, halflife), 19.0), __dsqrt_rn((double)boardArea));
    const double T = __dadd_rn(tLate, __dmul_rn(__dsub_rn(tEarly, tLate), detExp(__dmul_rn(halflives, detLog(0.5)))));
    double maxValue = 0.0;
    for(int i = 0; i < P; i++) if((double)pol[i] > maxValue) maxValue = (double)pol[i];
    if(maxValue > 0.0) {
      const double logMax = detLog(maxValue), invTemp = __ddiv_rn(1.0, T);
      double sum = 0.0;
      for(int i = 0; i < P; i++)
        if(pol[i] > 0.0f) {
          const float q = __double2float_rn(detExp(__dmul_rn(__dsub_rn(detLog((double)pol[i]), logMax), invTemp)));
          pol[i] = q; sum = __dadd_rn(sum, (double)q);
        }
      for(int i = 0; i < P; i++) if(pol[i] >= 0.0f) pol[i] = __double2float_rn(__ddiv_rn((double)pol[i], sum));
    }
  }
  if(c.rootNoise) {
    double r[MAX_POLICY];
    int legalCount = 0;
    for(int i = 0; i < P; i++) if(pol[i] >= 0.0f) legalCount++;
    if(legalCount == 0) return;
    double logSum = 0.0;
    for(int i = 0; i < P; i++) if(pol[i] >= 0.0f) { r[i] = detLog(__dadd_rn(fmin(0.01, (double)pol[i]), 1e-20)); logSum = __dadd_rn(logSum, r[i]); }
    const double logMean = __ddiv_rn(logSum, (double)legalCount);
    double alphaPropSum = 0.0;
    for(int i = 0; i < P; i++) if(pol[i] >= 0.0f) { r[i] = fmax(0.0, __dsub_rn(r[i], logMean)); alphaPropSum = __dadd_rn(alphaPropSum, r[i]); }
    const double uniformProb = __ddiv_rn(1.0, (double)legalCount);
    for(int i = 0; i < P; i++)
      if(pol[i] >= 0.0f) r[i] = alphaPropSum <= 0.0 ? uniformProb : __dmul_rn(0.5, __dadd_rn(__ddiv_rn(r[i], alphaPropSum), uniformProb));
    DetRng rng(c.seed, root.gameId[gi], turn);
    double rSum = 0.0;
    for(int i = 0; i < P; i++) {
      if(pol[i] >= 0.0f) { r[i] = rng.nextGamma(__dmul_rn(r[i], c.noiseConc)); rSum = __dadd_rn(rSum, r[i]); }
      else r[i] = 0.0;
    }
    const double w = c.noiseWeight;
    for(int i = 0; i < P; i++)
      if(pol[i] >= 0.0f) pol[i] = __double2float_rn(__dadd_rn(__dmul_rn(__ddiv_rn(r[i], rSum), w), __dmul_rn((double)pol[i], __dsub_rn(1.0, w))));
  }
}

// Tree re-use in graph mode: Search::makeMove (cpp/search/search.cpp:262-331) followed by the next beginSearch's
// recursivelyRecomputeStats (:670-690, 834-910), one warp per game.  Canonical order of the order-dependent steps:
//  1. the subgraph reachable from the played child is copied breadth first into the other node buffer (children in policy-index
//     order get the next free indices); the new root is a COPY of the child without bias entry (searchnode.cpp:149-188) and is
//     not in the transposition table;
//  2. every other node -- the old copy of the child included -- is dropped, in old-index order, and gives
//     subtreeValueBiasFreeProp of its last contribution back to its entry (removeSubtreeValueBias, search.cpp:773-786);
//  3. both tables are rebuilt for the kept nodes; bias entries without a kept node disappear (clearUnusedSynchronous);
//  4. with the bias on, every kept node is re-computed children first (deepest position first, ties by new index) without
//     adding a visit; a node without children gets its plain evaluation back (search.cpp:877-897).
__global__ void __launch_bounds__(128) k_reroot_graph(const SearchCfg c, TreeMem t, State root, const int16_t* __restrict__ chosen) {
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(gi >= c.numGames) return;
  uint8_t* src = t.nodes + (size_t)gi * c.maxNodes * c.nodeStride;
  uint8_t* dst = t.nodesAlt + (size_t)gi * c.maxNodes * c.nodeStride;
  int* queue = t.rerootQueue + (size_t)gi * c.maxNodes;
  int* remap = t.remap + (size_t)gi * c.maxNodes;
  const size_t tb = (size_t)gi * c.tableCap;
  const uint64_t* oldBiasKeys = t.biasKeys + tb * 2;
  double* oldBiasVals = t.biasVals + tb * 2;
  uint64_t* newTblKeys = t.tblKeysAlt + tb * 2; int* newTblVals = t.tblValsAlt + tb;
  uint64_t* newBiasKeys = t.biasKeysAlt + tb * 2; double* newBiasVals = t.biasValsAlt + tb * 2;
  for(int k = lane; k < c.tableCap; k += 32) {
    newTblVals[k] = 0; newBiasKeys[2 * k] = 0; newBiasKeys[2 * k + 1] = 0; newBiasVals[2 * k] = 0.0; newBiasVals[2 * k + 1] = 0.0;
  }
  const int oldCount = t.nodeCount[gi];
  const int move = chosen[gi];
  int c0 = -1;
  if(move >= 0 && oldCount > 0 && !(flagsOf(root.misc[gi]) & 1)) c0 = NodeRef{src, c.P, c.polOff}.child()[move];
  int count = 0;
  if(c0 >= 0) {
    for(int k = lane; k < oldCount; k += 32) remap[k] = -1;
    __syncwarp();
    if(lane == 0) { queue[0] = c0; remap[c0] = 0; }
    count = 1;
    __syncwarp();
    for(int i = 0; i < count; i++) {
      const int old = queue[i];
      const uint4* s4 = reinterpret_cast<const uint4*>(src + (size_t)old * c.nodeStride);
      uint4* d4 = reinterpret_cast<uint4*>(dst + (size_t)i * c.nodeStride);
      for(int k = lane; k < c.nodeStride / 16; k += 32) d4[k] = s4[k];
      __syncwarp();
      int* dch = NodeRef{dst + (size_t)i * c.nodeStride, c.P, c.polOff}.child();
      for(int p0 = 0; p0 < c.P; p0 += 32) {
        const int pos = p0 + lane;
        const int ch = pos < c.P ? dch[pos] : -1;
        const int r = ch >= 0 ? remap[ch] : 0;
        const bool isNew = ch >= 0 && r < 0;
        const unsigned m = __ballot_sync(0xffffffffu, isNew);
        if(isNew) {
          const int ni = count + __popc(m & ((1u << lane) - 1));
          queue[ni] = ch; remap[ch] = ni; dch[pos] = ni;
        } else if(ch >= 0) dch[pos] = r;
        count += __popc(m);
        __syncwarp();
      }
    }
    if(lane == 0) {
      // 2. dropped nodes release their bias contribution
      if(c.biasFactor != 0.0)
        for(int k = 0; k < oldCount; k++)
          if(remap[k] < 0 || k == c0) {
            NodeRef nd{src + (size_t)k * c.nodeStride, c.P, c.polOff};
            const int be = nd.biasEntry();
            if(be >= 0) {
              oldBiasVals[2 * be] = __dsub_rn(oldBiasVals[2 * be], __dmul_rn(nd.lastDelta(), c.biasFreeProp));
              oldBiasVals[2 * be + 1] = __dsub_rn(oldBiasVals[2 * be + 1], __dmul_rn(nd.lastWeight(), c.biasFreeProp));
            }
          }
      NodeRef nr{dst, c.P, c.polOff};
      nr.biasEntry() = -1; nr.lastDelta() = 0.0; nr.lastWeight() = 0.0;
      // 3. tables of the kept nodes
      for(int i = 1; i < count; i++) {
        NodeRef nd{dst + (size_t)i * c.nodeStride, c.P, c.polOff};
        if(c.useTable) {
          const uint64_t k0 = nd.key()[0], k1 = nd.key()[1];
          int slot = (int)(k0 & (uint64_t)(c.tableCap - 1));
          while(newTblVals[slot] != 0) slot = (slot + 1) & (c.tableCap - 1);
          newTblKeys[2 * slot] = k0; newTblKeys[2 * slot + 1] = k1; newTblVals[slot] = i + 1;
        }
        const int be = nd.biasEntry();
        if(be >= 0) {
          const uint64_t bk0 = oldBiasKeys[2 * be], bk1 = oldBiasKeys[2 * be + 1];
          int slot = (int)(splitmix64(bk0 ^ splitmix64(bk1)) & (uint64_t)(c.tableCap - 1));
          while(true) {
            if(newBiasKeys[2 * slot] == 0) {
              newBiasKeys[2 * slot] = bk0; newBiasKeys[2 * slot + 1] = bk1;
              newBiasVals[2 * slot] = oldBiasVals[2 * be]; newBiasVals[2 * slot + 1] = oldBiasVals[2 * be + 1];
              break;
            }
            if(newBiasKeys[2 * slot] == bk0 && newBiasKeys[2 * slot + 1] == bk1) break;
            slot = (slot + 1) & (c.tableCap - 1);
          }
          nd.biasEntry() = slot;
        }
      }
    }
    __syncwarp();
    // 4. children-first re-computation
    if(c.biasFactor != 0.0) {
      int lo = 1 << 20, hi = -1;
      for(int i = lane; i < count; i += 32) {
        const int d = NodeRef{dst + (size_t)i * c.nodeStride, c.P, c.polOff}.depth();
        lo = min(lo, d); hi = max(hi, d);
      }
      for(int o = 16; o > 0; o >>= 1) { lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
      for(int d = hi; d >= lo; d--)
        for(int base = 0; base < count; base += 32) {
          const int i = base + lane;
          const bool match = i < count && NodeRef{dst + (size_t)i * c.nodeStride, c.P, c.polOff}.depth() == d;
          unsigned m = __ballot_sync(0xffffffffu, match);
          while(m) {
            const int j = base + __ffs(m) - 1;
            m &= m - 1;
            NodeRef nd{dst + (size_t)j * c.nodeStride, c.P, c.polOff};
            if(nd.numChildren() == 0) {
              if(lane == 0) { nd.utilityAvg() = nd.nnUtility(); nd.utilitySqAvg() = __dmul_rn(nd.nnUtility(), nd.nnUtility()); }
              __syncwarp();
            } else recomputeNode<MAX_POLICY_SLOTS_PER_LANE, false>(c, newBiasVals, t.tcdf, t.stdevTab, dst, nd, lane, 0, j == 0);
          }
        }
    }
  }
  if(lane == 0) t.nodeCount[gi] = count;
}

// hash over every node of every game's graph (creation order = node index), compared with the oracle's
__global__ void __launch_bounds__(128) k_tree_digest(const SearchCfg c, TreeMem t, uint64_t* __restrict__ out) {
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(gi >= c.numGames) return;
  uint8_t* treeBase = t.nodes + (size_t)gi * c.maxNodes * c.nodeStride;
  const int count = t.nodeCount[gi];
  uint64_t h = 0;
  for(int i = lane; i < count; i += 32) {
    NodeRef nd{treeBase + (size_t)i * c.nodeStride, c.P, c.polOff};
    const uint64_t wb = (uint64_t)__double_as_longlong(nd.weightSum()), ub = (uint64_t)__double_as_longlong(nd.utilityAvg());
    const uint64_t qb = (uint64_t)__double_as_longlong(nd.utilitySqAvg()), sb = (uint64_t)__double_as_longlong(nd.weightSqSum());
    uint64_t nh = splitmix64((uint64_t)nd.N() ^ ((uint64_t)nd.numChildren() << 32)) ^ splitmix64(wb ^ PHI) ^ splitmix64(ub + PHI) ^
                  splitmix64(qb ^ (PHI << 1)) ^ splitmix64(sb + (PHI << 1));
    for(int pos = 0; pos < c.P; pos++) {
      const int cc = nd.child()[pos];
      if(cc != -1)
        nh ^= splitmix64((((uint64_t)(uint32_t)cc << 32) | (uint64_t)(uint32_t)nd.edgeN()[pos]) + (uint64_t)(pos + 1) * PHI + nd.order()[pos]);
    }
    h ^= splitmix64(nh + (uint64_t)(i + 1) * PHI);
  }
  for(int o = 16; o > 0; o >>= 1) h ^= __shfl_xor_sync(0xffffffffu, h, o);
  if(lane == 0) out[gi] = h;
}

// Search::getPlaySelectionValues at the root (cpp/search/searchresults.cpp:66-231), graph mode, one thread: the value of a move is
// its child's weight; children other than the most stably explored one are cut down to the weight the final explore-selection
// value of that one would have asked for (getReducedPlaySelectionWeight, searchexplorehelpers.cpp:209-243, rounded up); with
// useLcbForSelection the child with the best lower confidence bound (getSelfUtilityLCBAndRadius, searchhelpers.cpp:469-522) is
// raised above every child it beats.  psv [P] (0 for moves without a child); lcb / radius are scratch [P].
__device__ void playSelectionValues(const SearchCfg& c, uint8_t* treeBase, NodeRef nd, double* psv, double* lcb, double* radius) {
  const int* ch = nd.child(); const int* eN = nd.edgeN(); const uint8_t* ord = nd.order(); const float* pol = nd.policy();
  double total = 0.0;
  int n = 0;
  for(int pos = 0; pos < c.P; pos++) {
    psv[pos] = 0.0;
    if(ch[pos] == -1) continue;
    int cv; double cw, cu;
    childStats(c, treeBase, ch[pos], eN[pos], cv, cw, cu);
    psv[pos] = childWeightOf(cw, eN[pos], cv);
    total = __dadd_rn(total, psv[pos]);
    n++;
  }
  if(n == 0) return;
  int best = -1, bestOrd = 1 << 20;
  double bestWeight = -1e30, maxGoodness = -1e30;
  for(int pos = 0; pos < c.P; pos++) {
    if(ch[pos] == -1) continue;
    const double e = (double)eN[pos];
    const double g = __dadd_rn(__ddiv_rn(__dmul_rn(psv[pos], fmax(0.0, __dsub_rn(e, 1.0))), fmax(1.0, e)), __dmul_rn(2.0, (double)pol[pos]));
    if(g > maxGoodness || (g == maxGoodness && ord[pos] < bestOrd)) { maxGoodness = g; bestWeight = psv[pos]; best = pos; bestOrd = ord[pos]; }
  }
  const int pla = nd.nextPla();
  {
    const double scaling = __dmul_rn(c.cpuct, __dsqrt_rn(__dadd_rn(total, 0.01)));
    int cv; double cw, cu;
    childStats(c, treeBase, ch[best], eN[best], cv, cw, cu);
    const double bestValue = __dadd_rn(__ddiv_rn(__dmul_rn(scaling, (double)pol[best]), __dadd_rn(1.0, psv[best])), pla == 2 ? cu : -cu);
    for(int pos = 0; pos < c.P; pos++) {
      if(ch[pos] == -1 || pos == best) continue;
      childStats(c, treeBase, ch[pos], eN[pos], cv, cw, cu);
      double reduced = 0.0;
      if(cv > 0 && psv[pos] > 0.0) {
        double wanted = 0.0;
        if(pol[pos] >= 0.0f) {
          const double exploreComponent = __dsub_rn(bestValue, pla == 2 ? cu : -cu);
          if(exploreComponent <= 0) wanted = 1e100;
          else { wanted = __dsub_rn(__ddiv_rn(__dmul_rn(scaling, (double)pol[pos]), exploreComponent), 1.0); if(wanted < 0) wanted = 0.0; }
        }
        reduced = psv[pos] > wanted ? wanted : psv[pos];
      }
      psv[pos] = ceil(reduced);
    }
  }
  if(!c.useLcb) return;
  double bestLcb = -1e10;
  int bestLcbPos = -1, bestLcbOrd = 1 << 20;
  for(int pos = 0; pos < c.P; pos++) {
    if(ch[pos] == -1) continue;
    int cv; double cw, cu, cusq, cwsq;
    childStatsSq(c, treeBase, ch[pos], eN[pos], cv, cw, cu, cusq, cwsq);
    const double ratio = __ddiv_rn((double)eN[pos], (double)max(cv, 1));
    double weightSum = __dmul_rn(cw, ratio), weightSqSum = __dmul_rn(cwsq, ratio);
    radius[pos] = __dmul_rn(2.0, c.lcbStdevs);   // utilityRangeRadius = winLossUtilityFactor = 1
    lcb[pos] = -radius[pos];
    if(!(cv <= 0 || weightSum <= 0.0 || weightSqSum <= 0.0)) {
      double ess = __ddiv_rn(__dmul_rn(weightSum, weightSum), weightSqSum);
      const double priorWeight = __ddiv_rn(weightSum, __dmul_rn(__dmul_rn(ess, ess), ess));
      double usq = fmax(cusq, __dadd_rn(__dmul_rn(cu, cu), 1e-8));
      usq = __ddiv_rn(__dadd_rn(__dmul_rn(usq, weightSum), __dmul_rn(__dadd_rn(usq, 1.0), priorWeight)), __dadd_rn(weightSum, priorWeight));
      weightSum = __dadd_rn(weightSum, priorWeight);
      weightSqSum = __dadd_rn(weightSqSum, __dmul_rn(priorWeight, priorWeight));
      ess = __ddiv_rn(__dmul_rn(weightSum, weightSum), weightSqSum);
      const double selfUtility = pla == 2 ? cu : -cu;
      const double variance = __dsub_rn(usq, __dmul_rn(cu, cu));
      const double r = __dmul_rn(__dsqrt_rn(__ddiv_rn(variance, ess)), c.lcbStdevs);
      lcb[pos] = __dsub_rn(selfUtility, r);
      radius[pos] = r;
    }
    if(psv[pos] > 0 && psv[pos] >= __dmul_rn(c.minVisitPropLcb, bestWeight))
      if(lcb[pos] > bestLcb || (lcb[pos] == bestLcb && ord[pos] < bestLcbOrd)) { bestLcb = lcb[pos]; bestLcbPos = pos; bestLcbOrd = ord[pos]; }
  }
  if(bestLcbPos < 0 || (!c.nonBuggyLcb && ord[bestLcbPos] == 0)) return;   // useNonBuggyLcb false: the first-created child is never promoted
  double adjusted = psv[bestLcbPos];
  for(int pos = 0; pos < c.P; pos++) {
    if(ch[pos] == -1 || pos == bestLcbPos) continue;
    const double excess = __dsub_rn(bestLcb, lcb[pos]);
    if(excess < 0) continue;
    const double factor = __ddiv_rn(__dadd_rn(radius[pos], excess), __dadd_rn(radius[pos], __dmul_rn(0.20, excess)));
    const double lbound = __dmul_rn(__dmul_rn(factor, factor), psv[pos]);
    if(lbound > adjusted) adjusted = lbound;
  }
  psv[bestLcbPos] = adjusted;
}

// ---------------------------------------------------------------------------------------------
// choose the move, play it on the root game, drop the tree; refill finished games
// ---------------------------------------------------------------------------------------------
template <class D>
__global__ void k_choose_play(const Geom g, const SearchCfg c, State root, TreeMem t, TrainMem tr, const uint64_t* __restrict__ zob, int16_t* __restrict__ chosen,
                              double* __restrict__ psvOut) {
  const D dm(g);
  using BB = typename D::BB;
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if(gi >= c.numGames) return;
  GameRegs<BB> s;
  s.black = (BB)root.black[gi]; s.white = (BB)root.white[gi]; s.h0 = root.hash0[gi]; s.h1 = root.hash1[gi];
  s.id = root.gameId[gi]; s.misc = root.misc[gi];
  int move = -1;
  if(!(flagsOf(s.misc) & 1) && t.nodeCount[gi] > 0) {
    NodeRef nd{t.nodes + (size_t)gi * c.maxNodes * c.nodeStride, c.P, c.polOff};
    const int* ch = nd.child(); const int* eN = nd.edgeN(); const uint8_t* ord = nd.order();
    long long total = 0;
    int bestN = -1, bestOrd = 1 << 20, bestPos = -1;
    for(int pos = 0; pos < c.P; pos++)
      if(ch[pos] != -1) {
        total += eN[pos];
        if(eN[pos] > bestN || (eN[pos] == bestN && ord[pos] < bestOrd)) { bestN = eN[pos]; bestOrd = ord[pos]; bestPos = pos; }
      }
    const int ply = numTurnsOf(s.misc);
    double psv[MAX_POLICY];
    if(c.fullPlaySelection) {
      double lcb[MAX_POLICY], radius[MAX_POLICY];
      playSelectionValues(c, t.nodes + (size_t)gi * c.maxNodes * c.nodeStride, nd, psv, lcb, radius);
    } else
      for(int pos = 0; pos < c.P; pos++) psv[pos] = ch[pos] != -1 ? (double)eN[pos] : 0.0;
    if(psvOut) for(int pos = 0; pos < c.P; pos++) psvOut[(size_t)gi * c.P + pos] = psv[pos];
    if((c.moveTemp > 0.0 || c.moveTempEarly > 0.0 || c.fullPlaySelection) && bestN > 0) {
      // the reference's temperature schedule (searchresults.cpp:287-298, searchhelpers.cpp:12-49, 463-467): play-selection value = edge
      // visits, pruned / reduced, raised to 1/T with T interpolated from the early to the late temperature by 0.5^(turn/halflife...)
      double maxValue = 0.0;
      for(int pos = 0; pos < c.P; pos++) if(ch[pos] != -1 && psv[pos] > maxValue) maxValue = psv[pos];
      const double amountToSubtract = fmin(c.moveSubtract, __ddiv_rn(maxValue, 64.0)), amountToPrune = fmin(c.movePrune, __ddiv_rn(maxValue, 64.0));
      double newMax = 0.0;
      int bPos = -1, bOrd = 1 << 20;
      for(int pos = 0; pos < c.P; pos++) {
        if(ch[pos] == -1) continue;
        double x = psv[pos];
        if(x < amountToPrune) x = 0.0;
        else { x = __dsub_rn(x, amountToSubtract); if(x <= 0.0) x = 0.0; }
        if(x > newMax || (x == newMax && x > 0.0 && ord[pos] < bOrd)) { newMax = x; bPos = pos; bOrd = ord[pos]; }
      }
      const double hl = c.tempHalflife > 0.0 ? c.tempHalflife : 19.0;
      const double halflives = __ddiv_rn(__dmul_rn(__ddiv_rn((double)ply, hl), 19.0), __dsqrt_rn((double)c.boardArea));
      const double T = __dadd_rn(c.moveTemp, __dmul_rn(__dsub_rn(c.moveTempEarly, c.moveTemp), detExp(__dmul_rn(halflives, detLog(0.5)))));
      if(T <= 1.0e-4 || newMax <= 0.0) move = bPos;
      else {
        const double logMax = detLog(newMax);
        auto weightOf = [&](int pos) -> double {
          double x = psv[pos];
          if(x < amountToPrune) x = 0.0;
          else { x = __dsub_rn(x, amountToSubtract); if(x <= 0.0) x = 0.0; }
          return x <= 0.0 ? 0.0 : detExp(__ddiv_rn(__dsub_rn(detLog(x), logMax), T));
        };
        double sum = 0.0;
        for(int pos = 0; pos < c.P; pos++) if(ch[pos] != -1) sum = __dadd_rn(sum, weightOf(pos));
        const uint64_t r = splitmix64(c.seed ^ (s.id * PHI) ^ (uint64_t)ply ^ CHOOSE_SALT);
        const double d = __dmul_rn(__dmul_rn((double)(r & ((1ULL << 53) - 1ULL)), 1.0 / 9007199254740992.0), sum);
        double acc = 0.0;
        int last = -1;
        for(int pos = 0; pos < c.P; pos++) {
          if(ch[pos] == -1) continue;
          last = pos;
          acc = __dadd_rn(acc, weightOf(pos));
          if(acc > d) { move = pos; break; }
        }
        if(move < 0) move = last;
      }
    } else if(ply < c.temperaturePlies && total > 0) {
      const uint64_t r = splitmix64(c.seed ^ (s.id * PHI) ^ (uint64_t)ply ^ CHOOSE_SALT);
      long long k = (long long)(r % (uint64_t)total);
      for(int pos = 0; pos < c.P; pos++)
        if(ch[pos] != -1) { if(k < eN[pos]) { move = pos; break; } k -= eN[pos]; }
    } else move = bestPos;
    if(move >= 0) {
      if(tr.enabled) {
        const int k = tr.recCount[gi];
        if(k < tr.maxPlies) {
          const size_t r = (size_t)gi * tr.maxPlies + k;
          if(k == 0) tr.recGameId[gi] = s.id;
          tr.recBlack[r] = (uint64_t)s.black; tr.recWhite[r] = (uint64_t)s.white; tr.recMisc[r] = s.misc;
          tr.recN[r] = nd.N(); tr.recW[r] = c.graph ? __dmul_rn(nd.utilityAvg(), (double)nd.N()) : nd.W();
          for(int pos = 0; pos < c.P; pos++) tr.recVisits[r * c.P + pos] = (int16_t)(ch[pos] != -1 ? min(eN[pos], 32767) : 0);
          tr.recCount[gi] = k + 1;
        }
      }
      BB L[4]; bool illegal;
      Geom g2 = g; g2.autoRefill = 0;
      stepGame(dm, g2, s, move, true, zob, L, illegal);
      root.black[gi] = (uint64_t)s.black; root.white[gi] = (uint64_t)s.white; root.hash0[gi] = s.h0; root.hash1[gi] = s.h1; root.misc[gi] = s.misc;
      atomicAdd(&t.stats[3], 1ULL);
      const int fl = flagsOf(s.misc);
      if(fl & 1) {
        atomicAdd(&t.stats[4], 1ULL);
        const int w = (fl >> 1) & 3;
        atomicAdd(&t.stats[w == 1 ? 5 : w == 2 ? 6 : 7], 1ULL);
      }
    }
  }
  if(!c.reuseTree) t.nodeCount[gi] = 0;   // with tree re-use k_reroot decides what survives
  if(chosen) chosen[gi] = (int16_t)move;
}

// Tree re-use (Search::makeMove): one warp per game copies the subtree under the move just played, breadth first, into the
// other tree buffer (new root = node 0) and remaps the child indices; no subtree (unexpanded or terminal child, finished
// game, no move) leaves an empty tree.
__global__ void __launch_bounds__(128) k_reroot(const SearchCfg c, TreeMem t, State root, const int16_t* __restrict__ chosen) {
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(gi >= c.numGames) return;
  const uint8_t* src = t.nodes + (size_t)gi * c.maxNodes * c.nodeStride;
  uint8_t* dst = t.nodesAlt + (size_t)gi * c.maxNodes * c.nodeStride;
  int* queue = t.rerootQueue + (size_t)gi * c.maxNodes;
  const int move = chosen[gi];
  int count = 0;
  if(move >= 0 && t.nodeCount[gi] > 0 && !(flagsOf(root.misc[gi]) & 1)) {
    const int first = NodeRef{const_cast<uint8_t*>(src), c.P, c.polOff}.child()[move];
    if(first >= 0) {
      if(lane == 0) queue[0] = first;
      count = 1;
      __syncwarp();
      for(int i = 0; i < count; i++) {
        const int old = queue[i];
        const uint4* s4 = reinterpret_cast<const uint4*>(src + (size_t)old * c.nodeStride);
        uint4* d4 = reinterpret_cast<uint4*>(dst + (size_t)i * c.nodeStride);
        for(int k = lane; k < c.nodeStride / 16; k += 32) d4[k] = s4[k];
        __syncwarp();
        // children in policy-index order get the next free indices
        int* dch = NodeRef{dst + (size_t)i * c.nodeStride, c.P, c.polOff}.child();
        for(int p0 = 0; p0 < c.P; p0 += 32) {
          const int pos = p0 + lane;
          const int ch = pos < c.P ? dch[pos] : -1;
          const unsigned m = __ballot_sync(0xffffffffu, ch >= 0);
          if(ch >= 0) {
            const int ni = count + __popc(m & ((1u << lane) - 1));
            queue[ni] = ch;
            dch[pos] = ni;
          }
          count += __popc(m);
        }
        __syncwarp();
      }
    }
  }
  if(lane == 0) t.nodeCount[gi] = count;
}

// One warp per finished game: turn its records into training rows.
template <class D>
__global__ void __launch_bounds__(128) k_emit_rows(const Geom g, const SearchCfg c, State root, TrainMem tr) {
  const D dm(g);
  using BB = typename D::BB;
  const int gi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(gi >= c.numGames) return;
  const uint64_t fmisc = root.misc[gi];
  const int R = tr.recCount[gi];
  if(!(flagsOf(fmisc) & 1) || R == 0) return;
  const int HW = dm.HW(), P = c.P, PB = (HW + 7) / 8;
  int base = 0;
  if(lane == 0) {
    base = atomicAdd(&tr.rowCount[0], R);
    if(base + R > tr.maxRows) { atomicAdd(&tr.rowCount[0], -R); atomicAdd(&tr.rowCount[1], R); base = -1; }
  }
  base = __shfl_sync(0xffffffffu, base, 0);
  __syncwarp();
  if(base >= 0) {
    const int winner = (flagsOf(fmisc) >> 1) & 3;
    const double finalWin = winner == 2 ? 1.0 : winner == 1 ? 0.0 : 0.5;   // ScoreValue::whiteWinsOfWinner, draw = 0.5 (ledger C)
    const BB finalB = (BB)root.black[gi], finalW = (BB)root.white[gi];
    // longest same-colour run through every stone at the end (recordMaxConsecutives, canonical: stones only, 4 line directions)
    BB runAtLeast[2][8];   // [colour][len-1] cells of that colour lying in a run of length >= len (len 1..8)
    const int maxLen = max(dm.W(), dm.H());   // no run is longer (also keeps every shift inside the bitboard width)
    for(int col = 0; col < 2; col++) {
      const BB m = col ? finalW : finalB;
      for(int len = 1; len <= 8; len++) {
        BB acc = 0;
        if(len <= maxLen)
          for(int d = 0; d < 4; d++) acc |= coverAtLeast<D>(m, shiftOf(dm, d), len);
        runAtLeast[col][len - 1] = len == 1 ? m : acc;
      }
    }
    const uint64_t gid = tr.recGameId[gi];
    const uint64_t gh0 = splitmix64(gid), gh1 = splitmix64(gid ^ PHI);
    for(int i = 0; i < R; i++) {
      const size_t r = (size_t)gi * tr.maxPlies + i;
      const size_t row = (size_t)base + i;
      GameRegs<BB> s;
      s.black = (BB)tr.recBlack[r]; s.white = (BB)tr.recWhite[r]; s.misc = tr.recMisc[r]; s.h0 = s.h1 = 0; s.id = gid;
      const int pla = (flagsOf(s.misc) >> 3) & 3;
      // binaryInputNCHWPacked: the V1 planes of the position, bits big-endian within each byte (packBits, trainingwrite.cpp:218-233)
      BB empty = (BB)g.all & ~(s.black | s.white);
      int lastCell = histPla(s.misc, 0) ? histCell(s.misc, 0) : -1;
      BB L[4];
      legalMasks(dm, g, empty, lastCell, lastDirOf(s.misc), L);
      uint64_t planes[15];
      v1Planes(dm, g, s, L, planes, 1);
      for(int j = lane; j < 15 * PB; j += 32) {
        const int ch = j / PB, byte = j - ch * PB;
        const uint64_t dense = toDense(dm, (BB)planes[ch]);
        tr.outBin[row * 15 * PB + j] = (uint8_t)(__brev((uint32_t)((dense >> (8 * byte)) & 0xff)) >> 24);
      }
      // policy targets: this turn's visits, next turn's visits (uniform ones with weight 0 on the last turn)
      for(int pos = lane; pos < P; pos += 32) {
        tr.outPolicy[(row * 2 + 0) * P + pos] = tr.recVisits[r * P + pos];
        tr.outPolicy[(row * 2 + 1) * P + pos] = (i + 1 < R) ? tr.recVisits[(r + 1) * P + pos] : (int16_t)1;
      }
      // spatial value targets [5][HW]: final ownership, 0, board 2 plies ahead, board 6 plies ahead, final longest run
      auto boardAt = [&](int k, BB& b, BB& w) {   // position before move k; k == R: the final position
        if(k >= R) { b = finalB; w = finalW; } else { b = (BB)tr.recBlack[(size_t)gi * tr.maxPlies + k]; w = (BB)tr.recWhite[(size_t)gi * tr.maxPlies + k]; }
      };
      BB b2, w2, b3, w3;
      boardAt(min(i + 2, R), b2, w2);
      boardAt(min(i + 6, R), b3, w3);
      for(int cell = lane; cell < HW; cell += 32) {
        const int bit = padOf(dm, cell);
        auto rel = [&](BB b, BB w) -> int8_t {
          const int col = ((b >> bit) & 1) ? 1 : ((w >> bit) & 1) ? 2 : 0;
          return (int8_t)(col == 0 ? 0 : (col == pla ? 1 : -1));
        };
        int8_t* v = tr.outValue + row * 5 * HW;
        v[cell] = rel(finalB, finalW);
        v[HW + cell] = 0;
        v[2 * HW + cell] = rel(b2, w2);
        v[3 * HW + cell] = rel(b3, w3);
        int len = 0;
        for(int col = 0; col < 2; col++)
          for(int l = 1; l <= 8; l++) if((runAtLeast[col][l - 1] >> bit) & 1) len = max(len, l);
        v[4 * HW + cell] = (int8_t)len;
      }
      if(lane == 0) {
        tr.outGlobalIn[row] = (float)dm.K();
        float* gt = tr.outGlobalT + row * 64;
        for(int k = 0; k < 64; k++) gt[k] = 0.f;
        // td-like value targets from the player to move's view (fillValueTDTargets, trainingwrite.cpp:286-314): white win/loss
        // targets of turn j are the search's root estimate ((1 +- utility)/2), of the end the game result
        for(int f = 0; f < 5; f++) {
          const double nowFactor = tr.nowFactor[f];
          double winV = 0.0, lossV = 0.0, weightLeft = 1.0;
          for(int j = i; j <= R; j++) {
            double weightNow;
            if(j == R) { weightNow = weightLeft; weightLeft = 0.0; }
            else { weightNow = __dmul_rn(weightLeft, nowFactor); weightLeft = __dmul_rn(weightLeft, __dsub_rn(1.0, nowFactor)); }
            double tw, tl;
            if(j == R) { tw = finalWin; tl = __dsub_rn(1.0, finalWin); }
            else {
              const double u = __ddiv_rn(tr.recW[(size_t)gi * tr.maxPlies + j], (double)tr.recN[(size_t)gi * tr.maxPlies + j]);
              tw = __dmul_rn(__dadd_rn(1.0, u), 0.5); tl = __dmul_rn(__dsub_rn(1.0, u), 0.5);
            }
            winV = __dadd_rn(winV, __dmul_rn(weightNow, pla == 2 ? tw : tl));
            lossV = __dadd_rn(lossV, __dmul_rn(weightNow, pla == 2 ? tl : tw));
          }
          gt[2 * f] = (float)winV; gt[2 * f + 1] = (float)lossV;
        }
        gt[25] = 1.0f;                          // row weight
        gt[26] = 1.0f;                          // policy target weight
        gt[27] = 1.0f;                          // final ownership weight
        gt[28] = (i + 1 < R) ? 1.0f : 0.0f;     // next-turn policy target weight
        gt[33] = 1.0f;                          // future position weight
        for(int k = 36; k <= 40; k++) gt[k] = 1.0f;   // history masks (the reference draws them at random with p = 0.98 each)
        gt[41] = (float)(gh0 & 0x3FFFFF); gt[42] = (float)((gh0 >> 22) & 0x3FFFFF); gt[43] = (float)((gh0 >> 44) & 0xFFFFF);
        gt[44] = (float)(gh1 & 0x3FFFFF); gt[45] = (float)((gh1 >> 22) & 0x3FFFFF); gt[46] = (float)((gh1 >> 44) & 0xFFFFF);
        gt[51] = (float)numTurnsOf(s.misc);     // turn idx
        gt[60] = (float)tr.recN[r];             // visits of the search behind the row
        gt[63] = 1.0f;                          // data format version
      }
    }
  }
  __syncwarp();
  if(lane == 0) tr.recCount[gi] = 0;
}

template <class BB>
__global__ void k_refill(const Geom g, const SearchCfg c, State root) {
  const int gi = blockIdx.x * blockDim.x + threadIdx.x;
  if(gi >= c.numGames) return;
  const uint64_t misc = root.misc[gi];
  if(!(flagsOf(misc) & 1)) return;
  GameRegs<BB> s;
  resetGame(g, s, root.gameId[gi] + (uint64_t)c.numGames);
  root.black[gi] = 0; root.white[gi] = 0; root.hash0[gi] = s.h0; root.hash1[gi] = s.h1; root.gameId[gi] = s.id; root.misc[gi] = s.misc;
}

}  // namespace kc

// =============================================================================================
// host
// =============================================================================================
struct kc_search {
  kc_ctx* ctx = nullptr;
  kc_handle* handle = nullptr;   // null: hash evaluator
  kc_games* root = nullptr;      // the games being played (owned)
  kc_games* leaf = nullptr;      // the leaf positions of the current iteration (owned)
  kc::SearchCfg cfg;
  kc::TreeMem tree;
  float* d_policy = nullptr; float* d_winLoss = nullptr; float* d_misc = nullptr; uint64_t* d_nnHash = nullptr;
  int iterCounter = 0;            // NN cache: stamps of the iterations run so far
  int16_t* d_chosen = nullptr;
  double* d_psv = nullptr;                                   // [G][P] play-selection values of the last move choice
  float* d_rootAccPolicy = nullptr; float* d_rootAccScalars = nullptr;   // rootNumSymmetriesToSample: [G][P] and [G][4] running sums
  kc::TrainMem train = {};
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;
  // half batches (see runVisits): leaf positions of games [halfOff[h], halfOff[h] + halfCnt[h]) on their own stream
  bool pipelined = false;
  kc_games* leafHalf[2] = {nullptr, nullptr};
  int halfOff[2] = {0, 0}, halfCnt[2] = {0, 0};
  cudaEvent_t evFork = nullptr, evJoin[2] = {nullptr, nullptr};
};

using namespace kc;

namespace {
bool isStatic5(const Geom& g) { return g.W == 5 && g.H == 5 && g.K == 4; }

// graph mode: a search starts with empty transposition and bias tables (no tree re-use: every node of the previous search is
// gone, so every bias entry is unreferenced -- SubtreeValueBiasTable::clearUnusedSynchronous, search.cpp:689-690)
int clearTables(kc_search* S, cudaStream_t st) {
  const SearchCfg& c = S->cfg;
  if(!c.graph) return 0;
  const size_t slots = (size_t)c.numGames * c.tableCap;
  KC_CUDA(cudaMemsetAsync(S->tree.tblVals, 0, slots * 4, st));
  KC_CUDA(cudaMemsetAsync(S->tree.biasKeys, 0, slots * 16, st));
  KC_CUDA(cudaMemsetAsync(S->tree.biasVals, 0, slots * 16, st));
  return 0;
}

int runVisits(kc_search* S) {
  const SearchCfg& c = S->cfg;
  kc_games* R = S->root;
  cudaStream_t st = S->leaf->stream;
  const bool rootPolicyChange = c.graph && (c.rootNoise || (c.rootTemp > 0.0 && c.rootTemp != 1.0) || (c.rootTempEarly > 0.0 && c.rootTempEarly != 1.0));
  // One batch (every game, on the search's stream), or two half batches on their own streams: the small kernels of one half
  // (post-processing, expand / backup, select, root noise) then run while the trunk kernel of the other half has the tensor
  // cores -- the trunk's launch leaves 16 k registers per SM free for them (setmaxnreg, net_bf16.cu).  Trees of different games
  // never interact, so the split changes no result.
  const int nh = S->pipelined ? 2 : 1;
  // KC_SEARCH_PRELOAD=0: per-child loads instead of whole per-move arrays + shuffles in the graph-mode kernels (diagnostic)
  static const bool preload = [] { const char* e = getenv("KC_SEARCH_PRELOAD"); return !e || atoi(e) != 0; }();
  struct HalfRun { kc_games* leaf; SearchCfg c; cudaStream_t st; float* policy; float* winLoss; float* misc; uint64_t* nnHash; int rowOff; };
  HalfRun H[2];
  for(int h = 0; h < nh; h++) {
    H[h].leaf = S->pipelined ? S->leafHalf[h] : S->leaf;
    H[h].c = c;
    H[h].c.gOff = S->pipelined ? S->halfOff[h] : 0;
    H[h].c.gCnt = S->pipelined ? S->halfCnt[h] : c.numGames;
    H[h].c.half = h;
    H[h].st = H[h].leaf->stream;
    H[h].rowOff = H[h].c.gOff;
    H[h].policy = S->d_policy + (size_t)H[h].rowOff * c.P; H[h].winLoss = S->d_winLoss + (size_t)H[h].rowOff * 2;
    H[h].misc = S->d_misc + (size_t)H[h].rowOff * 2; H[h].nnHash = S->d_nnHash + (size_t)H[h].rowOff * 2;
  }
  if(S->pipelined) {
    KC_CUDA(cudaEventRecord(S->evFork, st));
    for(int h = 0; h < nh; h++) KC_CUDA(cudaStreamWaitEvent(H[h].st, S->evFork, 0));
  }
  if(c.rootSyms > 1) {
    for(int pass = 0; pass < c.rootSyms; pass++)
      for(int h = 0; h < nh; h++) {
        const SearchCfg& ch = H[h].c;
        kc_games* Lf = H[h].leaf;
        cudaStream_t hs = H[h].st;
        const int warpBlocks = (ch.gCnt * 32 + 127) / 128;
        k_rootsym_prepare<<<(ch.gCnt + 127) / 128, 128, 0, hs>>>(R->geom, ch, R->st, Lf->st, S->tree, Lf->d_sym, pass);
        S->launches++;
        if(S->handle) {
          if(kc::gamesEval(Lf, S->handle, nullptr, c.compact ? S->tree.evalCount + h : nullptr, H[h].rowOff, S->pipelined, /*symOnDevice=*/true)) return 1;
          kc::launchPostprocess(S->handle, ch.gCnt, c.LW, Lf->d_legal, Lf->d_status, Lf->d_sitHash, 1.0f, H[h].policy, H[h].winLoss, H[h].misc, H[h].nnHash, hs,
                                H[h].rowOff);
          S->launches += 3;
        } else {
          if(kc::gamesRefreshOutputs(Lf)) return 1;
          k_hash_eval<<<warpBlocks, 128, 0, hs>>>(ch.gCnt, c.P, c.LW, Lf->d_legal, Lf->d_sitHash, H[h].policy, H[h].winLoss, H[h].misc, Lf->d_sym);
          S->launches += 2;
        }
        k_rootsym_accumulate<<<warpBlocks, 128, 0, hs>>>(ch, S->tree, R->st, H[h].policy, H[h].winLoss, H[h].misc, S->d_rootAccPolicy, S->d_rootAccScalars, pass);
        S->launches++;
      }
  }
  for(int it = 0; it < c.maxVisits; it++) {
    if(c.reuseTree && it > 0 && (it & 31) == 0) {
      // games that kept a subtree finish their visit budget early: stop once no game had a visit left to make
      int active[2] = {0, 0};
      for(int h = 0; h < nh; h++) KC_CUDA(cudaMemcpyAsync(&active[h], S->tree.active + h, 4, cudaMemcpyDeviceToHost, H[h].st));
      for(int h = 0; h < nh; h++) KC_CUDA(cudaStreamSynchronize(H[h].st));
      if(!active[0] && !active[1]) break;
      for(int h = 0; h < nh; h++) KC_CUDA(cudaMemsetAsync(S->tree.active + h, 0, 4, H[h].st));
    }
    S->iterCounter++;
    for(int h = 0; h < nh; h++) {
      H[h].c.iterStamp = S->iterCounter;
      const SearchCfg& ch = H[h].c;
      kc_games* Lf = H[h].leaf;
      cudaStream_t hs = H[h].st;
      const int warpBlocks = (ch.gCnt * 32 + 127) / 128;
      // root noise / temperature: once per search, on roots kept by tree re-use (before the first descent) and on roots created
      // by the first iteration (before the second)
      if((rootPolicyChange || c.rootSyms > 1) && it <= 1) { k_root_noise<<<(ch.gCnt + 63) / 64, 64, 0, hs>>>(ch, S->tree, R->st, R->geom.HW); S->launches++; }
      if(c.compact) KC_CUDA(cudaMemsetAsync(S->tree.evalCount + h, 0, 4, hs));
      if(c.graph) {
        if(isStatic5(R->geom) && preload) k_select_graph<StaticDims<5, 5, 4>, true><<<warpBlocks, 128, 0, hs>>>(R->geom, ch, R->st, Lf->st, S->tree, R->d_zob, Lf->d_sym);
        else if(isStatic5(R->geom)) k_select_graph<StaticDims<5, 5, 4>, false><<<warpBlocks, 128, 0, hs>>>(R->geom, ch, R->st, Lf->st, S->tree, R->d_zob, Lf->d_sym);
        else k_select_graph<DynDims, false><<<warpBlocks, 128, 0, hs>>>(R->geom, ch, R->st, Lf->st, S->tree, R->d_zob, Lf->d_sym);
      } else if(isStatic5(R->geom)) k_select<StaticDims<5, 5, 4>><<<warpBlocks, 128, 0, hs>>>(R->geom, ch, R->st, Lf->st, S->tree, R->d_zob, Lf->d_sym);
      else k_select<DynDims><<<warpBlocks, 128, 0, hs>>>(R->geom, ch, R->st, Lf->st, S->tree, R->d_zob, Lf->d_sym);
      S->launches++;
      if(S->handle) {
        if(kc::gamesEval(Lf, S->handle, nullptr, c.compact ? S->tree.evalCount + h : nullptr, H[h].rowOff, S->pipelined, c.randomSym != 0)) return 1;
        kc::launchPostprocess(S->handle, ch.gCnt, c.LW, Lf->d_legal, Lf->d_status, Lf->d_sitHash, 1.0f, H[h].policy, H[h].winLoss, H[h].misc, H[h].nnHash, hs,
                              H[h].rowOff);
        S->launches += 3;
      } else {
        if(kc::gamesRefreshOutputs(Lf)) return 1;
        k_hash_eval<<<warpBlocks, 128, 0, hs>>>(ch.gCnt, c.P, c.LW, Lf->d_legal, Lf->d_sitHash, H[h].policy, H[h].winLoss, H[h].misc, nullptr);
        S->launches += 2;
      }
      if(c.graph) {
        if(c.P <= 128 && preload) k_expand_backup_graph<4, true><<<warpBlocks, 128, 0, hs>>>(ch, S->tree, H[h].policy, H[h].winLoss, H[h].misc);
        else if(c.P <= 128) k_expand_backup_graph<4, false><<<warpBlocks, 128, 0, hs>>>(ch, S->tree, H[h].policy, H[h].winLoss, H[h].misc);
        else k_expand_backup_graph<MAX_POLICY_SLOTS_PER_LANE, false><<<warpBlocks, 128, 0, hs>>>(ch, S->tree, H[h].policy, H[h].winLoss, H[h].misc);
      }
      else k_expand_backup<<<warpBlocks, 128, 0, hs>>>(ch, S->tree, H[h].policy, H[h].winLoss);
      S->launches++;
    }
  }
  if(S->pipelined)
    for(int h = 0; h < nh; h++) {
      KC_CUDA(cudaEventRecord(S->evJoin[h], H[h].st));
      KC_CUDA(cudaStreamWaitEvent(st, S->evJoin[h], 0));
    }
  KC_CUDA(cudaGetLastError());
  return 0;
}
}  // namespace

extern "C" {

int kc_search_create(kc_ctx* ctx, kc_handle* handleOrNull, int numGames, int xSize, int ySize, int winLen, const kc_search_params* p, kc_search** out) {
  KC_CHECK(ctx && p && out, "kc_search_create: null argument");
  KC_CHECK(p->maxVisits >= 1 && p->maxVisits <= 65536, "kc_search_create: maxVisits must be within 1..65536");
  KC_CHECK(p->cpuctExploration > 0 && p->fpuReductionMax >= 0 && p->rootFpuReductionMax >= 0, "kc_search_create: bad exploration parameters");
  KC_CHECK(p->subtreeValueBiasFreeProp >= 0.0 && p->subtreeValueBiasFreeProp <= 1.0, "kc_search_create: subtreeValueBiasFreeProp must be within 0..1");
  KC_CHECK(!p->rootNoiseEnabled || (p->rootDirichletNoiseTotalConcentration > 0.0 && p->rootDirichletNoiseWeight >= 0.0 && p->rootDirichletNoiseWeight <= 1.0),
           "kc_search_create: root noise needs rootDirichletNoiseTotalConcentration > 0 and a weight within 0..1");
  KC_CHECK(p->chosenMoveTemperature >= 0.0 && p->chosenMoveTemperatureEarly >= 0.0 && p->chosenMoveSubtract >= 0.0 && p->chosenMovePrune >= 0.0,
           "kc_search_create: negative move-choice option");
  KC_CHECK(p->valueWeightExponent >= 0.0 && p->valueWeightExponent <= 1.0, "kc_search_create: valueWeightExponent must be within 0..1");
  KC_CHECK(p->rootPolicyTemperature >= 0.0 && p->rootPolicyTemperatureEarly >= 0.0 && p->rootDesiredPerChildVisitsCoeff >= 0.0, "kc_search_create: negative root option");
  KC_CHECK(p->subtreeValueBiasFactor == 0.0 || p->subtreeValueBiasWeightExponent > 0.0, "kc_search_create: subtreeValueBiasWeightExponent must be positive");
  KC_CHECK(!p->useLcbForSelection || (p->lcbStdevs > 0.0 && p->minVisitPropForLCB >= 0.0), "kc_search_create: useLcbForSelection needs lcbStdevs > 0 and minVisitPropForLCB >= 0");
  KC_CHECK(p->rootNumSymmetriesToSample >= 0 && p->rootNumSymmetriesToSample <= 8, "kc_search_create: rootNumSymmetriesToSample must be within 0..8");
  KC_CHECK(!p->useUncertainty || (p->uncertaintyCoeff > 0.0 && p->uncertaintyExponent >= 0.0 && p->uncertaintyMaxWeight >= 1.0),
           "kc_search_create: useUncertainty needs uncertaintyCoeff > 0, uncertaintyExponent >= 0 and uncertaintyMaxWeight >= 1");
  KC_CHECK(p->nnCacheSizePowerOfTwo >= 0 && p->nnCacheSizePowerOfTwo <= 26, "kc_search_create: nnCacheSizePowerOfTwo must be within 0..26");
  KC_CHECK(!p->useNoisePruning || (p->noisePruneUtilityScale > 0.0 && p->noisePruningCap >= 0.0), "kc_search_create: useNoisePruning needs noisePruneUtilityScale > 0 and noisePruningCap >= 0");
  KC_CHECK(xSize <= KC_MAX_DEVICE_LEN && ySize <= KC_MAX_DEVICE_LEN,
           "kc_search_create: the device search runs on boards up to 7x7 (its rules and node layout are the 64-bit ones); kc_games_* take up to 10x10");
  KC_CUDA(cudaSetDevice(ctx->device));
  kc_search* S = new kc_search();
  S->ctx = ctx; S->handle = handleOrNull;
  if(kc_games_create(ctx, numGames, xSize, ySize, winLen, &S->root) || kc_games_create(ctx, numGames, xSize, ySize, winLen, &S->leaf)) { delete S; return 1; }
  if(handleOrNull && kc::handleCheckGeometry(handleOrNull, xSize, ySize, numGames)) { kc_games_destroy(S->root); kc_games_destroy(S->leaf); delete S; return 1; }
  SearchCfg& c = S->cfg;
  c.P = 4 * xSize * ySize; c.LW = (c.P + 31) / 32;
  c.rootNoise = p->rootNoiseEnabled ? 1 : 0; c.fpuPW = p->fpuParentWeightByVisitedPolicy ? 1 : 0;
  c.noiseConc = p->rootDirichletNoiseTotalConcentration; c.noiseWeight = p->rootDirichletNoiseWeight;
  c.rootTemp = p->rootPolicyTemperature; c.rootTempEarly = p->rootPolicyTemperatureEarly; c.tempHalflife = p->chosenMoveTemperatureHalflife;
  c.fpuPWPow = p->fpuParentWeightByVisitedPolicyPow > 0.0 ? p->fpuParentWeightByVisitedPolicyPow : 1.0; c.rootDesired = p->rootDesiredPerChildVisitsCoeff;
  c.vwExp = p->valueWeightExponent;
  c.randomSym = (p->nnRandomize && handleOrNull) ? 1 : 0;
  c.moveTemp = p->chosenMoveTemperature; c.moveTempEarly = p->chosenMoveTemperatureEarly; c.moveSubtract = p->chosenMoveSubtract; c.movePrune = p->chosenMovePrune;
  c.boardArea = xSize * ySize;
  const bool rootPolicyChange = c.rootNoise || (c.rootTemp > 0.0 && c.rootTemp != 1.0) || (c.rootTempEarly > 0.0 && c.rootTempEarly != 1.0);
  c.useLcb = p->useLcbForSelection ? 1 : 0; c.nonBuggyLcb = p->useNonBuggyLcb ? 1 : 0; c.lcbStdevs = p->lcbStdevs; c.minVisitPropLcb = p->minVisitPropForLCB;
  c.rootSyms = p->rootNumSymmetriesToSample > 1 ? std::min(8, p->rootNumSymmetriesToSample) : 1;
  c.noisePruning = p->useNoisePruning ? 1 : 0; c.noisePruneScale = p->noisePruneUtilityScale; c.noisePruneCap = p->noisePruningCap;
  c.useUncertainty = p->useUncertainty ? 1 : 0; c.uncCoeff = p->uncertaintyCoeff; c.uncExp = p->uncertaintyExponent; c.uncMaxWeight = p->uncertaintyMaxWeight;
  // every option beyond plain PUCT runs on the node-centric statistics of graph mode
  c.graph = (p->useGraphSearch || p->subtreeValueBiasFactor != 0.0 || rootPolicyChange || c.fpuPW || c.rootDesired > 0.0 || c.vwExp != 0.0 || c.useLcb || c.rootSyms > 1 ||
             c.useUncertainty || c.noisePruning) ? 1 : 0;
  // graph mode under the reference's move-choice schedule (or LCB): the move is chosen from the full getPlaySelectionValues
  c.fullPlaySelection = (c.graph && (c.useLcb || c.moveTemp > 0.0 || c.moveTempEarly > 0.0)) ? 1 : 0;
  c.useTable = p->useGraphSearch ? 1 : 0;
  c.biasFactor = p->subtreeValueBiasFactor; c.biasExp = p->subtreeValueBiasWeightExponent; c.biasFreeProp = p->subtreeValueBiasFreeProp;
  c.polOff = c.graph ? GRAPH_POL_OFF : 32 + 8 * c.P;
  c.nodeStride = (c.polOff + (c.graph ? 14 : 13) * c.P + 15) / 16 * 16;
  // graph mode with re-use: a kept subgraph can hold nodes whose creating visits went through another root child, so the
  // pool is a quarter larger than maxVisits; a visit that finds it empty is dropped (the oracle does the same)
  c.maxNodes = p->maxVisits + ((c.graph && p->reuseTree) ? p->maxVisits / 4 : 0);
  c.tableCap = 16;
  while(c.tableCap < 2 * c.maxNodes) c.tableCap *= 2;
  c.maxVisits = p->maxVisits; c.temperaturePlies = p->temperaturePlies;
  c.numGames = numGames; c.autoRefill = p->autoRefill ? 1 : 0;
  c.compact = (handleOrNull && kc::handleIsBf16(handleOrNull) && !p->noCompaction) ? 1 : 0;
  c.reuseTree = p->reuseTree ? 1 : 0;
  c.cpuct = p->cpuctExploration; c.fpuRed = p->fpuReductionMax; c.rootFpuRed = p->rootFpuReductionMax;
  c.seed = 0;
  const size_t n = (size_t)numGames;
  const size_t treeBytes = n * c.maxNodes * c.nodeStride * (c.reuseTree ? 2 : 1);
  size_t freeB = 0, totalB = 0;
  KC_CUDA(cudaMemGetInfo(&freeB, &totalB));
  KC_CHECK(treeBytes + (1ULL << 30) < freeB, "kc_search_create: the trees (numGames x maxVisits x " + std::to_string(c.nodeStride) + " B) do not fit in free device memory");
  KC_CUDA(cudaMalloc(&S->tree.nodes, treeBytes / (c.reuseTree ? 2 : 1)));
  if(c.reuseTree) {
    KC_CUDA(cudaMalloc(&S->tree.nodesAlt, treeBytes / 2));
    KC_CUDA(cudaMalloc(&S->tree.rerootQueue, n * c.maxNodes * 4));
  }
  KC_CUDA(cudaMalloc(&S->tree.active, 8)); KC_CUDA(cudaMemset(S->tree.active, 0, 8));
  KC_CUDA(cudaMalloc(&S->tree.nodeCount, n * 4)); KC_CUDA(cudaMemset(S->tree.nodeCount, 0, n * 4));
  KC_CUDA(cudaMalloc(&S->tree.pathNode, n * MAX_PATH * 4)); KC_CUDA(cudaMalloc(&S->tree.pathPos, n * MAX_PATH));
  KC_CUDA(cudaMalloc(&S->tree.pathLen, n * 4)); KC_CUDA(cudaMalloc(&S->tree.leafKind, n * 4));
  KC_CUDA(cudaMalloc(&S->tree.leafValue, n * 8)); KC_CUDA(cudaMalloc(&S->tree.leafNextPla, n * 4));
  KC_CUDA(cudaMalloc(&S->tree.leafSlot, n * 4)); KC_CUDA(cudaMemset(S->tree.leafSlot, 0, n * 4));
  KC_CUDA(cudaMalloc(&S->tree.evalCount, 8)); KC_CUDA(cudaMemset(S->tree.evalCount, 0, 8));
  KC_CUDA(cudaMalloc(&S->tree.stats, NUM_STATS * 8)); KC_CUDA(cudaMemset(S->tree.stats, 0, NUM_STATS * 8));
  if(c.graph) {
    {   // DistributionTable of the Student-t cdf, 3 degrees of freedom, closed form (search.cpp:111-116, distributiontable.cpp)
      std::vector<double> tab(2000);
      const double PI = 3.14159265358979323846, s3 = std::sqrt(3.0);
      for(int i = 0; i < 2000; i++) {
        const double z = -50.0 + (double)i * 100.0 / 1999.0, x = z / s3;
        tab[i] = i == 0 ? 0.0 : i == 1999 ? 1.0 : 0.5 + (x / (1.0 + x * x) + std::atan(x)) / PI;
      }
      double* d = nullptr;
      KC_CUDA(cudaMalloc(&d, 2000 * 8));
      KC_CUDA(cudaMemcpy(d, tab.data(), 2000 * 8, cudaMemcpyHostToDevice));
      S->tree.tcdf = d;
      std::vector<double> sd(1025, 0.0);
      for(int w = 1; w <= 1024; w++) sd[w] = std::sqrt(0.00000001 + 1.0 / (1.5 * std::sqrt((double)w)));
      double* d2 = nullptr;
      KC_CUDA(cudaMalloc(&d2, 1025 * 8));
      KC_CUDA(cudaMemcpy(d2, sd.data(), 1025 * 8, cudaMemcpyHostToDevice));
      S->tree.stdevTab = d2;
    }
    const size_t slots = n * c.tableCap;
    KC_CUDA(cudaMalloc(&S->tree.tblKeys, slots * 16)); KC_CUDA(cudaMalloc(&S->tree.tblVals, slots * 4));
    KC_CUDA(cudaMalloc(&S->tree.biasKeys, slots * 16)); KC_CUDA(cudaMalloc(&S->tree.biasVals, slots * 16));
    KC_CUDA(cudaMalloc(&S->tree.leafKey, n * 16)); KC_CUDA(cudaMalloc(&S->tree.leafBiasKey, n * 16)); KC_CUDA(cudaMalloc(&S->tree.leafTarget, n * 4));
    KC_CUDA(cudaMemset(S->tree.tblVals, 0, slots * 4)); KC_CUDA(cudaMemset(S->tree.biasKeys, 0, slots * 16)); KC_CUDA(cudaMemset(S->tree.biasVals, 0, slots * 16));
    if(c.reuseTree) {
      KC_CUDA(cudaMalloc(&S->tree.tblKeysAlt, slots * 16)); KC_CUDA(cudaMalloc(&S->tree.tblValsAlt, slots * 4));
      KC_CUDA(cudaMalloc(&S->tree.biasKeysAlt, slots * 16)); KC_CUDA(cudaMalloc(&S->tree.biasValsAlt, slots * 16));
      KC_CUDA(cudaMalloc(&S->tree.remap, n * c.maxNodes * 4));
    }
  }
  KC_CUDA(cudaMalloc(&S->d_policy, n * c.P * 4)); KC_CUDA(cudaMalloc(&S->d_winLoss, n * 8));
  KC_CUDA(cudaMalloc(&S->d_misc, n * 8)); KC_CUDA(cudaMalloc(&S->d_nnHash, n * 16));
  if(p->nnCacheSizePowerOfTwo > 0) {
    const size_t N = (size_t)1 << p->nnCacheSizePowerOfTwo;
    c.iterStamp = 0;
    S->tree.cacheMask = (unsigned)(N - 1);
    KC_CUDA(cudaMalloc(&S->tree.cacheKeys, N * 16)); KC_CUDA(cudaMemset(S->tree.cacheKeys, 0, N * 16));
    KC_CUDA(cudaMalloc(&S->tree.cacheState, N * 8)); KC_CUDA(cudaMemset(S->tree.cacheState, 0, N * 8));
    KC_CUDA(cudaMalloc(&S->tree.cachePolicy, N * c.P * 4)); KC_CUDA(cudaMalloc(&S->tree.cacheScalars, N * 16));
    KC_CUDA(cudaMalloc(&S->tree.leafCache, n * 4)); KC_CUDA(cudaMalloc(&S->tree.leafCKey, n * 16));
    KC_CUDA(cudaMalloc(&S->tree.claimWord, N * 8)); KC_CUDA(cudaMemset(S->tree.claimWord, 0, N * 8));
    KC_CUDA(cudaMalloc(&S->tree.claimKey, N * 16)); KC_CUDA(cudaMalloc(&S->tree.leafOwner, n * 4));
  }
  KC_CUDA(cudaMalloc(&S->d_chosen, n * 2));
  KC_CUDA(cudaMalloc(&S->d_psv, n * c.P * 8)); KC_CUDA(cudaMemset(S->d_psv, 0, n * c.P * 8));
  if(c.rootSyms > 1) { KC_CUDA(cudaMalloc(&S->d_rootAccPolicy, n * c.P * 4)); KC_CUDA(cudaMalloc(&S->d_rootAccScalars, n * 16)); }
  KC_CUDA(cudaEventCreate(&S->ev0)); KC_CUDA(cudaEventCreate(&S->ev1));
  {
    // two half batches when each still gives every SM pair a work item of the trunk kernel (bf16 path with device-side batching);
    // KC_SEARCH_PIPELINE=0 keeps one batch
    // Measured after the select / expand kernels got their compact child lists (b10c128, 18,944 games, full selfplay1.cfg, 4 moves):
    // one batch 5.97 s, two half batches 6.13 s -- beside a resident trunk CTA the search kernels get 8 warps per SM instead of 32 and
    // take about as long as the half trunk they hide under, while every half launch pays its own trunk prologue and wave tail.  So one
    // batch is the default; KC_SEARCH_PIPELINE=1 selects the half-batch pipeline.
    static const bool allow = [] { const char* e = getenv("KC_SEARCH_PIPELINE"); return e && atoi(e) != 0; }();
    const int item = 2 * kc::boardsPerTile(xSize, ySize);                 // boards per CTA work item
    const int half = ((numGames + 1) / 2 + item - 1) / item * item;
    if((p->noPipeline == 2 || (allow && p->noPipeline == 0)) && p->nnCacheSizePowerOfTwo <= 0 && c.compact && kc::handleCanLeaveRegisters(handleOrNull) && half >= item * ctx->smCount && numGames - half > 0) {
      S->pipelined = true;
      kc::handleLeaveRegisters(handleOrNull, true);
      S->halfOff[0] = 0; S->halfCnt[0] = half; S->halfOff[1] = half; S->halfCnt[1] = numGames - half;
      for(int h = 0; h < 2; h++) if(kc_games_create(ctx, S->halfCnt[h], xSize, ySize, winLen, &S->leafHalf[h])) return 1;
      KC_CUDA(cudaEventCreateWithFlags(&S->evFork, cudaEventDisableTiming));
      for(int h = 0; h < 2; h++) KC_CUDA(cudaEventCreateWithFlags(&S->evJoin[h], cudaEventDisableTiming));
    }
  }
  *out = S;
  return 0;
}

int kc_search_destroy(kc_search* S) {
  if(!S) return 0;
  cudaSetDevice(S->ctx->device);
  cudaStreamSynchronize(S->leaf->stream);
  cudaFree(S->tree.nodes); cudaFree(S->tree.nodeCount); cudaFree(S->tree.pathNode); cudaFree(S->tree.pathPos); cudaFree(S->tree.pathLen);
  cudaFree(S->tree.leafKind); cudaFree(S->tree.leafValue); cudaFree(S->tree.leafNextPla); cudaFree(S->tree.stats);
  cudaFree(S->tree.leafSlot); cudaFree(S->tree.evalCount);
  cudaFree(S->tree.nodesAlt); cudaFree(S->tree.rerootQueue); cudaFree(S->tree.active);
  cudaFree(S->tree.tblKeys); cudaFree(S->tree.tblVals); cudaFree(S->tree.biasKeys); cudaFree(S->tree.biasVals);
  cudaFree(S->tree.leafKey); cudaFree(S->tree.leafBiasKey); cudaFree(S->tree.leafTarget); cudaFree(const_cast<double*>(S->tree.tcdf)); cudaFree(const_cast<double*>(S->tree.stdevTab));
  cudaFree(S->tree.tblKeysAlt); cudaFree(S->tree.tblValsAlt); cudaFree(S->tree.biasKeysAlt); cudaFree(S->tree.biasValsAlt); cudaFree(S->tree.remap);
  { kc::TrainMem& t = S->train;
    cudaFree(t.recBlack); cudaFree(t.recWhite); cudaFree(t.recMisc); cudaFree(t.recN); cudaFree(t.recW); cudaFree(t.recVisits); cudaFree(t.recCount);
    cudaFree(t.recGameId); cudaFree(t.rowCount); cudaFree(t.outBin); cudaFree(t.outGlobalIn); cudaFree(t.outPolicy); cudaFree(t.outGlobalT); cudaFree(t.outValue); }
  cudaFree(S->d_policy); cudaFree(S->d_winLoss); cudaFree(S->d_misc); cudaFree(S->d_nnHash); cudaFree(S->d_chosen);
  cudaFree(S->d_psv); cudaFree(S->d_rootAccPolicy); cudaFree(S->d_rootAccScalars);
  cudaFree(S->tree.cacheKeys); cudaFree(S->tree.cacheState); cudaFree(S->tree.cachePolicy); cudaFree(S->tree.cacheScalars); cudaFree(S->tree.leafCache); cudaFree(S->tree.leafCKey); cudaFree(S->tree.claimWord); cudaFree(S->tree.claimKey); cudaFree(S->tree.leafOwner);
  cudaEventDestroy(S->ev0); cudaEventDestroy(S->ev1);
  for(int h = 0; h < 2; h++) { if(S->leafHalf[h]) { cudaStreamSynchronize(S->leafHalf[h]->stream); kc_games_destroy(S->leafHalf[h]); } if(S->evJoin[h]) cudaEventDestroy(S->evJoin[h]); }
  if(S->evFork) cudaEventDestroy(S->evFork);
  kc_games_destroy(S->root); kc_games_destroy(S->leaf);
  delete S;
  return 0;
}

kc_games* kc_search_games(kc_search* S) { return S ? S->root : nullptr; }

int kc_search_reset(kc_search* S, uint64_t seed, uint64_t firstGameId) {
  KC_CHECK(S, "kc_search_reset: null search");
  KC_CUDA(cudaSetDevice(S->ctx->device));
  S->cfg.seed = seed;
  if(kc_games_reset(S->root, seed, firstGameId, 0)) return 1;
  KC_CUDA(cudaMemset(S->tree.nodeCount, 0, (size_t)S->cfg.numGames * 4));
  KC_CUDA(cudaMemset(S->tree.stats, 0, NUM_STATS * 8));
  if(clearTables(S, S->leaf->stream)) return 1;
  if(S->tree.cacheMask) {   // evaluations depend on the seed through nnRandomize's symmetry: a new run starts with an empty NN cache
    KC_CUDA(cudaMemsetAsync(S->tree.cacheKeys, 0, ((size_t)S->tree.cacheMask + 1) * 16, S->leaf->stream));
    KC_CUDA(cudaMemsetAsync(S->tree.cacheState, 0, ((size_t)S->tree.cacheMask + 1) * 8, S->leaf->stream));
    KC_CUDA(cudaMemsetAsync(S->tree.claimWord, 0, ((size_t)S->tree.cacheMask + 1) * 8, S->leaf->stream));
  }
  KC_CUDA(cudaStreamSynchronize(S->leaf->stream));
  return 0;
}

int kc_search_run_visits(kc_search* S) {
  KC_CHECK(S, "kc_search_run_visits: null search");
  KC_CUDA(cudaSetDevice(S->ctx->device));
  KC_CUDA(cudaMemsetAsync(S->tree.nodeCount, 0, (size_t)S->cfg.numGames * 4, S->leaf->stream));
  if(clearTables(S, S->leaf->stream)) return 1;
  if(runVisits(S)) return 1;
  KC_CUDA(cudaStreamSynchronize(S->leaf->stream));
  if(S->handle && kc::handleCheckAbort(S->handle)) return 1;
  return 0;
}

int kc_search_read_root(kc_search* S, int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits, double* edgeUtilitySum, float* policy, uint8_t* order) {
  KC_CHECK(S, "kc_search_read_root: null search");
  KC_CUDA(cudaSetDevice(S->ctx->device));
  const SearchCfg& c = S->cfg;
  std::vector<uint8_t> node(c.nodeStride);
  std::vector<int> counts(c.numGames);
  KC_CUDA(cudaMemcpy(counts.data(), S->tree.nodeCount, (size_t)c.numGames * 4, cudaMemcpyDeviceToHost));
  for(int gi = 0; gi < c.numGames; gi++) {
    const bool have = counts[gi] > 0;
    if(have) KC_CUDA(cudaMemcpy(node.data(), S->tree.nodes + (size_t)gi * c.maxNodes * c.nodeStride, c.nodeStride, cudaMemcpyDeviceToHost));
    else std::fill(node.begin(), node.end(), 0);
    const uint8_t* b = node.data();
    const int* child = reinterpret_cast<const int*>(b + c.polOff + 4 * c.P);
    const int* edgeN = reinterpret_cast<const int*>(b + c.polOff + 8 * c.P);
    const int visits = have ? *reinterpret_cast<const int*>(b) : 0;
    if(rootVisits) rootVisits[gi] = visits;
    // graph mode reports utilityAvg * visits (and per edge the child's utilityAvg * edgeVisits): one rounding, as the oracle does
    if(rootUtilitySum) rootUtilitySum[gi] = !have ? 0.0 : c.graph ? *reinterpret_cast<const double*>(b + 24) * (double)visits : *reinterpret_cast<const double*>(b + 16);
    for(int pos = 0; pos < c.P; pos++) {
      const bool ex = have && child[pos] != -1;
      if(edgeVisits) edgeVisits[(size_t)gi * c.P + pos] = ex ? edgeN[pos] : 0;
      if(edgeUtilitySum) {
        double v = 0.0;
        if(ex && !c.graph) v = reinterpret_cast<const double*>(b + 32)[pos];
        else if(ex) {
          double cu = child[pos] == -4 ? 1.0 : child[pos] == -3 ? -1.0 : 0.0;
          if(child[pos] >= 0)
            KC_CUDA(cudaMemcpy(&cu, S->tree.nodes + ((size_t)gi * c.maxNodes + child[pos]) * c.nodeStride + 24, 8, cudaMemcpyDeviceToHost));
          v = cu * (double)edgeN[pos];
        }
        edgeUtilitySum[(size_t)gi * c.P + pos] = v;
      }
      if(policy) policy[(size_t)gi * c.P + pos] = have ? reinterpret_cast<const float*>(b + c.polOff)[pos] : 0.f;
      if(order) order[(size_t)gi * c.P + pos] = ex ? (b + c.polOff + 12 * c.P)[pos] : 255;
    }
  }
  return 0;
}

int kc_search_play(kc_search* S, int moves, int16_t* chosenLast, kc_search_stats* acc, float* msTotal) {
  KC_CHECK(S && moves > 0, "kc_search_play: bad argument");
  KC_CUDA(cudaSetDevice(S->ctx->device));
  const SearchCfg& c = S->cfg;
  kc_games* R = S->root;
  cudaStream_t st = S->leaf->stream;
  const int blocks = (c.numGames + 127) / 128;
  KC_CUDA(cudaMemsetAsync(S->tree.stats, 0, NUM_STATS * 8, st));
  KC_CUDA(cudaEventRecord(S->ev0, st));
  for(int m = 0; m < moves; m++) {
    if(c.autoRefill) {
      if(isStatic5(R->geom)) k_refill<uint32_t><<<blocks, 128, 0, st>>>(R->geom, c, R->st);
      else k_refill<uint64_t><<<blocks, 128, 0, st>>>(R->geom, c, R->st);
      S->launches++;
    }
    if(!c.reuseTree) KC_CUDA(cudaMemsetAsync(S->tree.nodeCount, 0, (size_t)c.numGames * 4, st));
    KC_CUDA(cudaMemsetAsync(S->tree.active, 0, 8, st));
    if(!c.reuseTree && clearTables(S, st)) return 1;   // with re-use the re-rooting keeps the tables in step with the graph
    if(runVisits(S)) return 1;
    if(isStatic5(R->geom)) k_choose_play<StaticDims<5, 5, 4>><<<blocks, 128, 0, st>>>(R->geom, c, R->st, S->tree, S->train, R->d_zob, S->d_chosen, S->d_psv);
    else k_choose_play<DynDims><<<blocks, 128, 0, st>>>(R->geom, c, R->st, S->tree, S->train, R->d_zob, S->d_chosen, S->d_psv);
    S->launches++;
    if(c.reuseTree) {
      if(c.graph) {
        k_reroot_graph<<<(c.numGames * 32 + 127) / 128, 128, 0, st>>>(c, S->tree, R->st, S->d_chosen);
        std::swap(S->tree.tblKeys, S->tree.tblKeysAlt); std::swap(S->tree.tblVals, S->tree.tblValsAlt);
        std::swap(S->tree.biasKeys, S->tree.biasKeysAlt); std::swap(S->tree.biasVals, S->tree.biasValsAlt);
      } else k_reroot<<<(c.numGames * 32 + 127) / 128, 128, 0, st>>>(c, S->tree, R->st, S->d_chosen);
      std::swap(S->tree.nodes, S->tree.nodesAlt);
      S->launches++;
    }
    if(S->train.enabled) {
      const int wb = (c.numGames * 32 + 127) / 128;
      if(isStatic5(R->geom)) k_emit_rows<StaticDims<5, 5, 4>><<<wb, 128, 0, st>>>(R->geom, c, R->st, S->train);
      else k_emit_rows<DynDims><<<wb, 128, 0, st>>>(R->geom, c, R->st, S->train);
      S->launches++;
    }
  }
  KC_CUDA(cudaEventRecord(S->ev1, st));
  KC_CUDA(cudaGetLastError());
  unsigned long long hs[NUM_STATS];
  KC_CUDA(cudaMemcpyAsync(hs, S->tree.stats, NUM_STATS * 8, cudaMemcpyDeviceToHost, st));
  if(chosenLast) KC_CUDA(cudaMemcpyAsync(chosenLast, S->d_chosen, (size_t)c.numGames * 2, cudaMemcpyDeviceToHost, st));
  KC_CUDA(cudaStreamSynchronize(st));
  if(S->handle && kc::handleCheckAbort(S->handle)) return 1;
  if(msTotal) KC_CUDA(cudaEventElapsedTime(msTotal, S->ev0, S->ev1));
  if(acc) {
    acc->visits += hs[0]; acc->netEvals += hs[1]; acc->terminalVisits += hs[2]; acc->movesPlayed += hs[3];
    acc->gamesFinished += hs[4]; acc->blackWins += hs[5]; acc->whiteWins += hs[6]; acc->draws += hs[7];
    acc->batchRows += c.compact ? hs[1] : (uint64_t)c.numGames * c.maxVisits * moves;
    acc->transpositionHits += hs[8]; acc->catchUpVisits += hs[9]; acc->nnCacheHits += hs[10];
  }
  return 0;
}

int kc_search_enable_training_rows(kc_search* S, int maxRows) {
  KC_CHECK(S && maxRows > 0, "kc_search_enable_training_rows: bad argument");
  KC_CHECK(!S->train.enabled, "kc_search_enable_training_rows: already enabled");
  KC_CUDA(cudaSetDevice(S->ctx->device));
  const SearchCfg& c = S->cfg;
  const Geom& g = S->root->geom;
  TrainMem& t = S->train;
  const size_t n = (size_t)c.numGames, mp = (size_t)g.HW, PB = (size_t)(g.HW + 7) / 8;
  t.maxPlies = g.HW; t.maxRows = maxRows;
  KC_CUDA(cudaMalloc(&t.recBlack, n * mp * 8)); KC_CUDA(cudaMalloc(&t.recWhite, n * mp * 8)); KC_CUDA(cudaMalloc(&t.recMisc, n * mp * 8));
  KC_CUDA(cudaMalloc(&t.recN, n * mp * 4)); KC_CUDA(cudaMalloc(&t.recW, n * mp * 8));
  KC_CUDA(cudaMalloc(&t.recVisits, n * mp * c.P * 2));
  KC_CUDA(cudaMalloc(&t.recCount, n * 4)); KC_CUDA(cudaMemset(t.recCount, 0, n * 4));
  KC_CUDA(cudaMalloc(&t.recGameId, n * 8));
  KC_CUDA(cudaMalloc(&t.rowCount, 8)); KC_CUDA(cudaMemset(t.rowCount, 0, 8));
  KC_CUDA(cudaMalloc(&t.outBin, (size_t)maxRows * 15 * PB)); KC_CUDA(cudaMalloc(&t.outGlobalIn, (size_t)maxRows * 4));
  KC_CUDA(cudaMalloc(&t.outPolicy, (size_t)maxRows * 2 * c.P * 2)); KC_CUDA(cudaMalloc(&t.outGlobalT, (size_t)maxRows * 64 * 4));
  KC_CUDA(cudaMalloc(&t.outValue, (size_t)maxRows * 5 * g.HW));
  const double area = (double)g.HW;
  t.nowFactor[0] = 0.0; t.nowFactor[1] = 1.0 / (1.0 + area * 0.176); t.nowFactor[2] = 1.0 / (1.0 + area * 0.056);
  t.nowFactor[3] = 1.0 / (1.0 + area * 0.016); t.nowFactor[4] = 1.0;
  t.enabled = 1;
  return 0;
}

int kc_search_read_training_rows(kc_search* S, int* numRows, int* numDropped, uint8_t* binaryInputNCHWPacked, float* globalInputNC,
                                 int16_t* policyTargetsNCMove, float* globalTargetsNC, int8_t* valueTargetsNCHW, int clear) {
  KC_CHECK(S && numRows, "kc_search_read_training_rows: null argument");
  KC_CHECK(S->train.enabled, "kc_search_read_training_rows: call kc_search_enable_training_rows first");
  KC_CUDA(cudaSetDevice(S->ctx->device));
  KC_CUDA(cudaStreamSynchronize(S->leaf->stream));
  const SearchCfg& c = S->cfg;
  const Geom& g = S->root->geom;
  const TrainMem& t = S->train;
  int cnt[2];
  KC_CUDA(cudaMemcpy(cnt, t.rowCount, 8, cudaMemcpyDeviceToHost));
  const size_t rows = (size_t)cnt[0], PB = (size_t)(g.HW + 7) / 8;
  *numRows = cnt[0];
  if(numDropped) *numDropped = cnt[1];
  if(rows) {
    if(binaryInputNCHWPacked) KC_CUDA(cudaMemcpy(binaryInputNCHWPacked, t.outBin, rows * 15 * PB, cudaMemcpyDeviceToHost));
    if(globalInputNC) KC_CUDA(cudaMemcpy(globalInputNC, t.outGlobalIn, rows * 4, cudaMemcpyDeviceToHost));
    if(policyTargetsNCMove) KC_CUDA(cudaMemcpy(policyTargetsNCMove, t.outPolicy, rows * 2 * c.P * 2, cudaMemcpyDeviceToHost));
    if(globalTargetsNC) KC_CUDA(cudaMemcpy(globalTargetsNC, t.outGlobalT, rows * 64 * 4, cudaMemcpyDeviceToHost));
    if(valueTargetsNCHW) KC_CUDA(cudaMemcpy(valueTargetsNCHW, t.outValue, rows * 5 * g.HW, cudaMemcpyDeviceToHost));
  }
  if(clear) KC_CUDA(cudaMemset(t.rowCount, 0, 8));
  return 0;
}

int kc_search_read_play_selection(kc_search* S, double* playSelection) {
  KC_CHECK(S && playSelection, "kc_search_read_play_selection: null argument");
  KC_CUDA(cudaSetDevice(S->ctx->device));
  KC_CUDA(cudaStreamSynchronize(S->leaf->stream));
  KC_CUDA(cudaMemcpy(playSelection, S->d_psv, (size_t)S->cfg.numGames * S->cfg.P * 8, cudaMemcpyDeviceToHost));
  return 0;
}

int kc_search_tree_digest(kc_search* S, uint64_t* digest) {
  KC_CHECK(S && digest, "kc_search_tree_digest: null argument");
  KC_CHECK(S->cfg.graph, "kc_search_tree_digest: only for searches created with useGraphSearch or a subtree value bias");
  KC_CUDA(cudaSetDevice(S->ctx->device));
  const SearchCfg& c = S->cfg;
  uint64_t* d = nullptr;
  KC_CUDA(cudaMalloc(&d, (size_t)c.numGames * 8));
  k_tree_digest<<<(c.numGames * 32 + 127) / 128, 128, 0, S->leaf->stream>>>(c, S->tree, d);
  const cudaError_t e = cudaMemcpyAsync(digest, d, (size_t)c.numGames * 8, cudaMemcpyDeviceToHost, S->leaf->stream);
  cudaStreamSynchronize(S->leaf->stream);
  cudaFree(d);
  KC_CUDA(e);
  KC_CUDA(cudaGetLastError());
  return 0;
}

int64_t kc_search_launch_count(const kc_search* S) {
  if(!S) return 0;
  int64_t n = S->launches + kc_games_launch_count(S->leaf);
  for(int h = 0; h < 2; h++) if(S->leafHalf[h]) n += kc_games_launch_count(S->leafHalf[h]);
  return n;
}

}  // extern "C"
