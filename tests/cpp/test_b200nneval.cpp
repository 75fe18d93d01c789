// Driver for katacoffee_b200/host/b200nneval.{h,cpp}: class NNEvaluator with the reference's interface (cpp/neuralnet/nneval.h:80-175)
// over the evaluator front end.  Reads like the reference's use of it (cpp/search/search.cpp:1054-1066, cpp/tests/testnnevalcanary.cpp):
// construct, spawnServerThreads, evaluate(board, history, nextPla, params, buf, skipCache, includeOwnerMap) from several threads.
//
//   test_b200nneval <model file> cpu    batch function = a pure function of the staged row (no GPU): interface, errors, cache, statistics
//   test_b200nneval <model file> gpu    the device: results equal kc_evaluator_evaluate on the same positions, probabilities are normalised
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "b200nneval.h"

static int g_fail = 0;
#define EXPECT(cond)                                                                  \
  do {                                                                                \
    if(!(cond)) { g_fail++; fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); } \
  } while(0)
template <class F>
static bool throwsWith(F f, const char* what) {
  try { f(); } catch(const StringError& e) { return strstr(e.what(), what) != nullptr; }
  return false;
}

static uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
static float valueOf(uint64_t b, uint64_t w, uint64_t m, int sym, int j) { return (float)(mix(b ^ mix(w) ^ mix(m ^ (uint64_t)sym * 77) ^ (uint64_t)j) >> 40) * (1.0f / 16777216.0f); }
static int fakeBackend(void*, int, const kc_eval_batch* b) {
  for(int i = 0; i < b->n; i++) {
    for(int j = 0; j < 100; j++) b->policyProbs[(size_t)i * 100 + j] = valueOf(b->black[i], b->white[i], b->misc[i], b->symmetry[i], j);
    b->whiteWinLoss[2 * i] = valueOf(b->black[i], b->white[i], b->misc[i], b->symmetry[i], 1000);
    b->whiteWinLoss[2 * i + 1] = valueOf(b->black[i], b->white[i], b->misc[i], b->symmetry[i], 1001);
    b->miscOut[2 * i] = 3.f; b->miscOut[2 * i + 1] = 0.5f;
    if(b->wantOwnership) for(int j = 0; j < 25; j++) b->ownership[(size_t)i * 25 + j] = 0.25f;
  }
  return 0;
}

// a position reached by alternating placements (rule legality is not what this driver checks: any empty cell, any direction)
struct Position { Board board; BoardHistory hist; Player nextPla; uint64_t black = 0, white = 0, misc = 0; };
static Position makePosition(uint64_t id, int W = 5, int H = 5) {
  Position p{Board(W, H, 4), BoardHistory(), P_BLACK};
  const int plies = (int)(mix(id) % 12);
  uint64_t r = id * 1234567ULL + 1;
  for(int t = 0; t < plies; t++) {
    r = mix(r);
    const int cell = (int)(r % (uint64_t)(W * H)), x = cell % W, y = cell / W;
    const Spot s = Location::getSpot(x, y, W);
    if(p.board.colors[s] != C_EMPTY) continue;
    // a direction whose line through the cell still has an empty cell (board.cpp:217-226: otherwise the move is illegal)
    static const int DX[4] = {0, -1, -1, 1}, DY[4] = {-1, 0, -1, -1};
    int dir = -1;
    for(int k = 0; k < 4 && dir < 0; k++) {
      const int d = (int)(((r >> 20) + (uint64_t)k) & 3);
      for(int sgn = -1; sgn <= 1 && dir < 0; sgn += 2)
        for(int xx = x + sgn * DX[d], yy = y + sgn * DY[d]; xx >= 0 && yy >= 0 && xx < W && yy < H; xx += sgn * DX[d], yy += sgn * DY[d])
          if(p.board.colors[Location::getSpot(xx, yy, W)] == C_EMPTY) { dir = d; break; }
    }
    if(dir < 0) continue;
    p.board.colors[s] = p.nextPla;
    p.hist.moveHistory.push_back(Move(Loc(s, (Direction)dir), p.nextPla));
    p.hist.numTurns++;
    p.nextPla = (Player)(p.nextPla ^ 3);
  }
  for(int y = 0; y < H; y++)
    for(int x = 0; x < W; x++) {
      const Color c = p.board.colors[Location::getSpot(x, y, W)];
      if(c == C_BLACK) p.black |= 1ULL << (y * (W + 1) + x);
      if(c == C_WHITE) p.white |= 1ULL << (y * (W + 1) + x);
    }
  const size_t n = p.hist.moveHistory.size();
  int lastDir = 4;
  for(int k = 0; k < 5 && (size_t)k < n; k++) {   // byte k = the move k plies ago: cell | player << 6
    const Move& m = p.hist.moveHistory[n - 1 - k];
    const int cell = Location::getY(m.loc.spot, W) * W + Location::getX(m.loc.spot, W);
    p.misc |= (uint64_t)(cell | (m.pla << 6)) << (8 * k);
    if(k == 0) lastDir = m.loc.dir;
  }
  p.misc |= ((uint64_t)lastDir << 40) | ((uint64_t)p.hist.numTurns << 48) | ((uint64_t)(p.nextPla << 3) << 56);
  return p;
}

static NNEvaluator* makeEvaluator(const std::string& modelFile, int maxBatch, int numThreads, bool doRandomize, int defaultSymmetry, int nnLen = 5) {
  return new NNEvaluator("b200", modelFile, "", nullptr, maxBatch, /*maxConcurrentEvals*/ 4 * maxBatch + 1, nnLen, nnLen, /*requireExactNNLen*/ true,
                         /*inputsUseNHWC*/ false, /*nnCacheSizePowerOfTwo*/ 14, /*nnMutexPoolSizePowerofTwo*/ 8, /*debugSkipNeuralNet*/ false, "", "", false,
                         enabled_t::Auto, enabled_t::Auto, numThreads, std::vector<int>((size_t)numThreads, 0), "seed", doRandomize, defaultSymmetry);
}

static void cpuMode(const std::string& modelFile) {
  // constructor errors keep the reference's wording (nneval.cpp:115-124)
  EXPECT(throwsWith([&] { delete makeEvaluator(modelFile, 0, 1, false, 0); }, "maxBatchSize is negative"));
  EXPECT(throwsWith([&] { delete makeEvaluator(modelFile, 8, 1, false, 0, 11); }, "Maximum supported nnEval board size"));
  EXPECT(throwsWith([&] { delete new NNEvaluator("m", modelFile, "", nullptr, 8, 8, 5, 5, true, false, 10, 4, false, "", "", false, enabled_t::Auto, enabled_t::Auto, 2,
                                                 std::vector<int>{0}, "s", false, 0); }, "gpuIdxByServerThread.size() != numThreads"));
  EXPECT(throwsWith([&] { delete makeEvaluator("/nonexistent/model.bin.gz", 8, 1, false, 0); }, ""));

  NNEvaluator* nnEval = makeEvaluator(modelFile, 8, 2, false, 3);
  nnEval->setBackendForTesting(fakeBackend, nullptr);
  EXPECT(nnEval->getNNXLen() == 5 && nnEval->getMaxBatchSize() == 8 && nnEval->getNumServerThreads() == 2 && nnEval->getNumGpus() == 1 && nnEval->getModelVersion() == 1);
  EXPECT(!nnEval->getInternalModelName().empty() && nnEval->getModelName() == "b200" && !nnEval->isNeuralNetLess() && nnEval->supportsShorttermError());
  MiscNNInputParams params;
  {
    Position p = makePosition(1);
    NNResultBuf buf;
    EXPECT(throwsWith([&] { nnEval->evaluate(p.board, p.hist, p.nextPla, params, buf, false, false); }, "before spawnServerThreads"));
  }
  nnEval->spawnServerThreads();
  EXPECT(throwsWith([&] { nnEval->setNumThreads({0}); }, "already running"));
  {
    Position big = makePosition(2, 6, 6), small = makePosition(3, 4, 4);
    NNResultBuf buf;
    EXPECT(throwsWith([&] { nnEval->evaluate(big.board, big.hist, big.nextPla, params, buf, false, false); }, "larger x or y size"));
    EXPECT(throwsWith([&] { nnEval->evaluate(small.board, small.hist, small.nextPla, params, buf, false, false); }, "different x or y size"));
    EXPECT(!buf.hasResult);
  }
  // four search threads, 300 evaluations each over 200 positions
  std::vector<Position> ps;
  for(int i = 0; i < 200; i++) ps.push_back(makePosition(100 + (uint64_t)i));
  std::vector<std::thread> threads;
  for(int t = 0; t < 4; t++)
    threads.emplace_back([&, t] {
      NNResultBuf buf;
      uint64_t r = (uint64_t)t;
      for(int i = 0; i < 300; i++) {
        r = mix(r);
        Position& p = ps[r % ps.size()];
        const bool own = (r >> 30) % 3 == 0;
        nnEval->evaluate(p.board, p.hist, p.nextPla, params, buf, false, own);
        EXPECT(buf.hasResult && buf.result != nullptr && buf.symmetry == 3);
        const NNOutput& o = *buf.result;
        bool ok = true;
        for(int j = 0; j < 100; j++) ok = ok && o.policyProbs[j] == valueOf(p.black, p.white, p.misc, 3, j);
        for(int j = 100; j < NNPos::MAX_NN_POLICY_SIZE; j++) ok = ok && o.policyProbs[j] == -1.0f;
        EXPECT(ok);
        EXPECT(o.whiteWinProb == valueOf(p.black, p.white, p.misc, 3, 1000) && o.whiteLossProb == valueOf(p.black, p.white, p.misc, 3, 1001));
        EXPECT(o.varTimeLeft == 3.f && o.shorttermWinlossError == 0.5f && o.nnXLen == 5 && o.nnYLen == 5);
        EXPECT((o.whiteOwnerMap != nullptr) == own);
        if(own) EXPECT(std::fabs(o.whiteOwnerMap[7] - (p.nextPla == P_WHITE ? 1.f : -1.f) * std::tanh(0.25f)) < 1e-7f);
        // NNOutput::nnHash = NNInputs::getHash
        int8_t stones[25]; int16_t moves[10];
        for(int c = 0; c < 25; c++) stones[c] = p.board.colors[Location::getSpot(c % 5, c / 5, 5)];
        const size_t n = p.hist.moveHistory.size();
        for(int k = 0; k < 5; k++) {
          const bool have = n >= (size_t)(5 - k);
          moves[2 * k] = have ? (int16_t)NNPos::locToPos(p.hist.moveHistory[n - 5 + k].loc, 5, 5, 5) : (int16_t)-1;
          moves[2 * k + 1] = have ? p.hist.moveHistory[n - 5 + k].pla : 0;
        }
        kc_eval_position ep{stones, moves, p.hist.numTurns, (int8_t)p.nextPla};
        uint64_t h[2], key[2];
        kc_eval_position_hash(5, 5, &ep, 1.0f, h, key);
        EXPECT(o.nnHash == Hash128(h[0], h[1]));
      }
    });
  for(auto& t : threads) t.join();
  const uint64_t rows0 = nnEval->numRowsProcessed();
  EXPECT(rows0 >= 100 && rows0 < 1200 && nnEval->numBatchesProcessed() > 0 && nnEval->averageProcessedBatchSize() >= 1.0 && nnEval->averageProcessedBatchSize() <= 8.0);
  // cache, clearCache, run-time symmetry settings, a second policy temperature
  NNResultBuf buf;
  nnEval->evaluate(ps[0].board, ps[0].hist, ps[0].nextPla, params, buf, false, false);
  const uint64_t rows = nnEval->numRowsProcessed();
  nnEval->evaluate(ps[0].board, ps[0].hist, ps[0].nextPla, params, buf, false, false);
  EXPECT(nnEval->numRowsProcessed() == rows);
  nnEval->clearCache();
  nnEval->evaluate(ps[0].board, ps[0].hist, ps[0].nextPla, params, buf, false, false);
  EXPECT(nnEval->numRowsProcessed() == rows + 1);
  nnEval->setDefaultSymmetry(5);
  EXPECT(nnEval->getDefaultSymmetry() == 5 && !nnEval->getDoRandomize());
  nnEval->evaluate(ps[1].board, ps[1].hist, ps[1].nextPla, params, buf, true, false);
  EXPECT(buf.symmetry == 5 && buf.result->policyProbs[0] == valueOf(ps[1].black, ps[1].white, ps[1].misc, 5, 0));
  params.symmetry = 6;
  nnEval->evaluate(ps[1].board, ps[1].hist, ps[1].nextPla, params, buf, true, false);
  EXPECT(buf.symmetry == 6);
  params.symmetry = NNInputs::SYMMETRY_NOTSPECIFIED;
  nnEval->setDoRandomize(true);
  int seen = 0;
  for(int i = 0; i < 40; i++) { nnEval->evaluate(ps[i].board, ps[i].hist, ps[i].nextPla, params, buf, true, false); seen |= 1 << buf.symmetry; }
  EXPECT(__builtin_popcount(seen) >= 4);
  MiscNNInputParams warm;
  warm.nnPolicyTemperature = 1.5f;
  nnEval->evaluate(ps[2].board, ps[2].hist, ps[2].nextPla, warm, buf, false, false);
  NNResultBuf buf2;
  nnEval->evaluate(ps[2].board, ps[2].hist, ps[2].nextPla, params, buf2, false, false);
  EXPECT(!(buf.result->nnHash == buf2.result->nnHash));   // the temperature is folded into the hash (nninputs.cpp:485-492)
  {   // the root under several symmetries at once, averaged as NNOutput's averaging constructor does (nninputs.cpp:95-170)
    const int syms[4] = {1, 4, 6, 3};
    const Position& p = ps[5];
    const uint64_t before = nnEval->numRowsProcessed();
    NNResultBuf avg;
    nnEval->evaluateAveragedOverSymmetries(ps[5].board, ps[5].hist, ps[5].nextPla, params, syms, 4, avg, true);
    EXPECT(avg.hasResult && nnEval->numRowsProcessed() == before + 4);
    bool ok = true;
    for(int j = 0; j < 100; j++) {
      float sum = 0.0f;
      for(int k = 0; k < 4; k++) sum += valueOf(p.black, p.white, p.misc, syms[k], j);
      ok = ok && avg.result->policyProbs[j] == sum / 4.0f;
    }
    for(int j = 100; j < NNPos::MAX_NN_POLICY_SIZE; j++) ok = ok && avg.result->policyProbs[j] == -1.0f;
    EXPECT(ok);
    float w = 0.0f;
    for(int k = 0; k < 4; k++) w += valueOf(p.black, p.white, p.misc, syms[k], 1000);
    EXPECT(avg.result->whiteWinProb == w / 4.0f && avg.result->varTimeLeft == 3.f);
    EXPECT(avg.result->whiteOwnerMap != nullptr && std::fabs(avg.result->whiteOwnerMap[3] - (p.nextPla == P_WHITE ? 1.f : -1.f) * std::tanh(0.25f)) < 1e-6f);
    EXPECT(throwsWith([&] { nnEval->evaluateAveragedOverSymmetries(ps[5].board, ps[5].hist, ps[5].nextPla, params, syms, 0, avg, false); }, "1..8 symmetries"));
    const int bad[2] = {0, 8};
    EXPECT(throwsWith([&] { nnEval->evaluateAveragedOverSymmetries(ps[5].board, ps[5].hist, ps[5].nextPla, params, bad, 2, avg, false); }, "out of range"));
  }
  nnEval->clearStats();
  EXPECT(nnEval->numRowsProcessed() == 0);
  nnEval->killServerThreads();
  nnEval->setNumThreads({0});
  EXPECT(nnEval->getNumServerThreads() == 1);
  nnEval->spawnServerThreads();
  nnEval->evaluate(ps[3].board, ps[3].hist, ps[3].nextPla, params, buf, false, false);
  EXPECT(buf.hasResult && nnEval->numRowsProcessed() == 1);
  delete nnEval;
}

static void gpuMode(const std::string& modelFile) {
  NNEvaluator* nnEval = makeEvaluator(modelFile, 64, 2, false, 0);
  nnEval->spawnServerThreads();
  std::vector<Position> ps;
  for(int i = 0; i < 160; i++) ps.push_back(makePosition(500 + (uint64_t)i));
  std::vector<std::shared_ptr<NNOutput>> results(ps.size());
  std::vector<std::thread> threads;
  for(int t = 0; t < 4; t++)
    threads.emplace_back([&, t] {
      NNResultBuf buf;
      MiscNNInputParams params;
      for(size_t i = (size_t)t; i < ps.size(); i += 4) {
        params.symmetry = (int)(i % 8);
        nnEval->evaluate(ps[i].board, ps[i].hist, ps[i].nextPla, params, buf, false, i % 2 == 0);
        results[i] = buf.result;
      }
    });
  for(auto& t : threads) t.join();
  for(size_t i = 0; i < ps.size(); i++) {
    const NNOutput& o = *results[i];
    double sum = 0; int legal = 0;
    for(int j = 0; j < 100; j++) if(o.policyProbs[j] >= 0) { sum += o.policyProbs[j]; legal++; }
    EXPECT(legal > 0 && std::fabs(sum - 1.0) < 1e-4);
    EXPECT(std::fabs(o.whiteWinProb + o.whiteLossProb - 1.0f) < 1e-5f && o.varTimeLeft >= 0 && o.shorttermWinlossError >= 0);
    if(i % 2 == 0) { bool ok = o.whiteOwnerMap != nullptr; for(int j = 0; ok && j < 25; j++) ok = std::fabs(o.whiteOwnerMap[j]) <= 1.0f; EXPECT(ok); }
    // an occupied cell is never legal
    for(int c = 0; c < 25; c++)
      if(ps[i].board.colors[Location::getSpot(c % 5, c / 5, 5)] != C_EMPTY)
        for(int d = 0; d < 4; d++) EXPECT(o.policyProbs[d * 25 + c] == -1.0f);
  }
  {   // all eight symmetries in one call = the mean of eight single evaluations
    const int syms[8] = {0, 1, 2, 3, 4, 5, 6, 7};
    const Position& p = ps[11];
    Position q = p;
    NNResultBuf avg, one;
    MiscNNInputParams params;
    nnEval->evaluateAveragedOverSymmetries(q.board, q.hist, q.nextPla, params, syms, 8, avg, false);
    std::vector<double> mean(100, 0.0);
    double win = 0.0;
    for(int s = 0; s < 8; s++) {
      params.symmetry = s;
      nnEval->evaluate(q.board, q.hist, q.nextPla, params, one, true, false);
      for(int j = 0; j < 100; j++) mean[j] += one.result->policyProbs[j] / 8.0;
      win += one.result->whiteWinProb / 8.0;
    }
    double worst = 0.0;
    for(int j = 0; j < 100; j++) worst = std::max(worst, std::fabs(mean[j] - (double)avg.result->policyProbs[j]));
    EXPECT(worst < 1e-4 && std::fabs(win - (double)avg.result->whiteWinProb) < 1e-4);
  }
  EXPECT(nnEval->numRowsProcessed() > 0 && nnEval->isAnyThreadUsingFP16());
  printf("NNEvaluator on the device: %llu rows in %llu batches\n", (unsigned long long)nnEval->numRowsProcessed(), (unsigned long long)nnEval->numBatchesProcessed());
  delete nnEval;
}

int main(int argc, char** argv) {
  if(argc < 3) { fprintf(stderr, "usage: test_b200nneval <model file> cpu|gpu\n"); return 2; }
  try {
    if(std::string(argv[2]) == "gpu") gpuMode(argv[1]); else cpuMode(argv[1]);
  } catch(const std::exception& e) {
    fprintf(stderr, "exception: %s\n", e.what());
    return 1;
  }
  if(g_fail) { fprintf(stderr, "%d check(s) failed\n", g_fail); return 1; }
  printf("test_b200nneval %s: ok\n", argv[2]);
  return 0;
}
