// b200nneval.cpp -- NNEvaluator (cpp/neuralnet/nneval.h) over kc_evaluator_* (see b200nneval.h).
//
// What each reference piece became:
//   ctor: loadModelFile + createComputeContext     nneval.cpp:142-170   the same calls into the B200 backend (b200backend.cpp)
//   spawnServerThreads / serve                      nneval.cpp:341-586   kc_evaluator_create_multi: one server thread per entry of
//                                                                        gpuIdxByServerThread, on that GPU's (kc_ctx, kc_model)
//   evaluate: getHash, cache, fillRowV1, queue,     nneval.cpp:588-845   hand the position to kc_evaluator_evaluate; copy the NNOutput fields
//             wait, post-process, cache store
//   NNCacheTable / clearCache / statistics          nneval.cpp:820-932   kc_evaluator_clear_cache / _get_stats / _clear_stats
// Exceptions: StringError with the reference's messages for the size checks (nneval.cpp:599-611); any failure of the C ABI is
// rethrown as StringError(kc_last_error()).
#include "b200nneval.h"

#include <algorithm>
#include <cstring>

using namespace std;

#ifdef KC_IN_REFERENCE_TREE
namespace NeuralNet {   // exported by b200backend.cpp next to the nninterface.h functions (declared in reftypes.h for the standalone build)
void getB200ContextAndModel(ComputeContext* context, int gpuIdx, void** kcCtx, void** kcModel);
bool getB200UseFP32Check(const ComputeContext* context);
}
#endif

static uint64_t hashSeed(const string& s) {   // FNV-1a: the seed only has to differ between evaluators
  uint64_t h = 1469598103934665603ULL;
  for(unsigned char c : s) { h ^= c; h *= 1099511628211ULL; }
  return h;
}

NNEvaluator::NNEvaluator(const string& mName, const string& mFileName, const string& expectedSha256, Logger* lg, int maxBatch, int maxConcurrent, int xLen,
                         int yLen, bool rExactNNLen, bool iUseNHWC, int cacheSizePowerOfTwo, int mutexPoolSizePowerofTwo, bool skipNeuralNet,
                         const string& openCLTunerFile, const string& homeDataDirOverride, bool openCLReTunePerBoardSize, enabled_t useFP16Mode,
                         enabled_t useNHWCMode, int numThr, const vector<int>& gpuIdxByServerThr, const string& rSeed, bool doRandomize, int defaultSymmetry)
  : modelName(mName), modelFileName(mFileName), nnXLen(xLen), nnYLen(yLen), requireExactNNLen(rExactNNLen), policySize(NNPos::getPolicySize(xLen, yLen)),
    usingFP16Mode(useFP16Mode), usingNHWCMode(useNHWCMode), maxBatchSize(maxBatch), maxConcurrentEvals(maxConcurrent),
    nnCacheSizePowerOfTwo(cacheSizePowerOfTwo), nnMutexPoolSizePowerofTwo(mutexPoolSizePowerofTwo), gpuIdxByServerThread(gpuIdxByServerThr),
    randSeedHash(hashSeed(rSeed)), logger(lg), loadedModel(nullptr), computeContext(nullptr), modelVersion(-1), currentDoRandomize(doRandomize),
    currentDefaultSymmetry(defaultSymmetry), spawned(false), lastInstance(nullptr), testFn(nullptr), testUser(nullptr) {
  (void)iUseNHWC;   // rows are not filled on the host any more
  if(nnXLen > NNPos::MAX_BOARD_LEN) throw StringError("Maximum supported nnEval board size is " + to_string(NNPos::MAX_BOARD_LEN));   // nneval.cpp:115-118
  if(nnYLen > NNPos::MAX_BOARD_LEN) throw StringError("Maximum supported nnEval board size is " + to_string(NNPos::MAX_BOARD_LEN));
  if(maxConcurrentEvals <= 0) throw StringError("maxConcurrentEvals is negative: " + to_string(maxConcurrentEvals));                   // nneval.cpp:119-122
  if(maxBatchSize <= 0) throw StringError("maxBatchSize is negative: " + to_string(maxBatchSize));
  if((int)gpuIdxByServerThread.size() != numThr) throw StringError("gpuIdxByServerThread.size() != numThreads");                        // nneval.cpp:123-124
  if(skipNeuralNet) throw StringError("B200 NNEvaluator: debugSkipNeuralNet is not supported (there is no CPU path)");
  if(!requireExactNNLen) throw StringError("B200 NNEvaluator: requireExactNNLen is required (boards exactly nnXLen x nnYLen)");
  loadedModel = NeuralNet::loadModelFile(modelFileName, expectedSha256);                                                                 // nneval.cpp:154-157
  modelVersion = NeuralNet::getModelVersion(loadedModel);
  vector<int> gpuIdxs = gpuIdxByServerThread;
  sort(gpuIdxs.begin(), gpuIdxs.end());
  gpuIdxs.erase(unique(gpuIdxs.begin(), gpuIdxs.end()), gpuIdxs.end());
  computeContext = NeuralNet::createComputeContext(gpuIdxs, logger, nnXLen, nnYLen, openCLTunerFile, homeDataDirOverride, openCLReTunePerBoardSize,
                                                   useFP16Mode, useNHWCMode, loadedModel);                                               // nneval.cpp:158-168
}

NNEvaluator::~NNEvaluator() {
  killServerThreads();
  NeuralNet::freeComputeContext(computeContext);   // nneval.cpp:194-199
  NeuralNet::freeLoadedModel(loadedModel);
}

string NNEvaluator::getInternalModelName() const { return NeuralNet::getModelName(loadedModel); }

void NNEvaluator::spawnServerThreads() { spawned = true; }   // the front ends are created on first use: win_len arrives with the first board

void NNEvaluator::killServerThreads() {
  lock_guard<mutex> lock(instancesMutex);
  lastInstance.store(nullptr);
  for(Instance* i : instances) { kc_evaluator_destroy(i->ev); delete i; }
  instances.clear();
  spawned = false;
}

void NNEvaluator::setNumThreads(const vector<int>& gpuIdxByServerThr) {
  if(spawned) throw StringError("NNEvaluator::setNumThreads called when threads were already running!");   // nneval.cpp:284-285
  gpuIdxByServerThread = gpuIdxByServerThr;
}

kc_evaluator* NNEvaluator::instanceFor(int winLen, float temperature) {
  const Instance* last = lastInstance.load(memory_order_acquire);
  if(last && last->winLen == winLen && last->temperature == temperature) return last->ev;
  lock_guard<mutex> lock(instancesMutex);
  if(!spawned) throw StringError("NNEvaluator::evaluate called before spawnServerThreads");
  for(Instance* i : instances)
    if(i->winLen == winLen && i->temperature == temperature) { lastInstance.store(i, memory_order_release); return i->ev; }
  kc_evaluator_config cfg{};
  cfg.nnXLen = nnXLen; cfg.nnYLen = nnYLen; cfg.winLen = winLen; cfg.maxBatch = maxBatchSize; cfg.maxConcurrentEvals = maxConcurrentEvals;
  cfg.numServerThreads = (int)gpuIdxByServerThread.size(); cfg.cacheSizePowerOfTwo = nnCacheSizePowerOfTwo; cfg.mutexPoolSizePowerOfTwo = nnMutexPoolSizePowerofTwo;
  cfg.doRandomize = 1;   // evaluate() passes the default symmetry itself when randomisation is off: setDoRandomize works while running
  cfg.defaultSymmetry = 0; cfg.randSeed = randSeedHash; cfg.policyTemperature = temperature;
  cfg.handleFlags = NeuralNet::getB200UseFP32Check(computeContext) ? KC_FLAG_FP32_CHECK : 0u;
  kc_evaluator* ev = nullptr;
  if(testFn) {
    if(kc_evaluator_create_custom(&cfg, testFn, testUser, &ev)) throw StringError(string("B200 NNEvaluator: ") + kc_last_error());
  } else {
    vector<kc_ctx*> ctxs; vector<const kc_model*> models;
    for(int gpu : gpuIdxByServerThread) {
      void *c = nullptr, *m = nullptr;
      NeuralNet::getB200ContextAndModel(computeContext, gpu, &c, &m);
      ctxs.push_back(static_cast<kc_ctx*>(c)); models.push_back(static_cast<const kc_model*>(m));
    }
    if(kc_evaluator_create_multi(cfg.numServerThreads, ctxs.data(), models.data(), &cfg, &ev)) throw StringError(string("B200 NNEvaluator: ") + kc_last_error());
  }
  Instance* inst = new Instance{winLen, temperature, ev};
  instances.push_back(inst);
  lastInstance.store(inst, memory_order_release);
  return ev;
}

void NNEvaluator::checkBoard(const Board& board, const MiscNNInputParams& nnInputParams) const {
  if(board.x_size > nnXLen || board.y_size > nnYLen)   // nneval.cpp:599-603
    throw StringError("NNEvaluator was configured with nnXLen = " + to_string(nnXLen) + " nnYLen = " + to_string(nnYLen) +
                      " but was asked to evaluate board with larger x or y size");
  if(board.x_size != nnXLen || board.y_size != nnYLen)   // nneval.cpp:604-610 (requireExactNNLen is always on here)
    throw StringError("NNEvaluator was configured with nnXLen = " + to_string(nnXLen) + " nnYLen = " + to_string(nnYLen) +
                      " and requireExactNNLen, but was asked to evaluate board with different x or y size");
  if(nnInputParams.playoutDoublingAdvantage != 0 || nnInputParams.policyOptimism > 0)
    throw StringError("B200 NNEvaluator: playoutDoublingAdvantage / policyOptimism are not inputs of the Coffee V1 features");
}

// (Board, BoardHistory) -> the position format of the C ABI: stones [H*W], the last five moves oldest first
static kc_eval_position toPosition(const Board& board, const BoardHistory& history, Player nextPlayer, int nnXLen, int nnYLen, int8_t* stones, int16_t* moves) {
  for(int y = 0; y < board.y_size; y++)
    for(int x = 0; x < board.x_size; x++) stones[y * board.x_size + x] = board.colors[Location::getSpot(x, y, board.x_size)];
  const size_t n = history.moveHistory.size();
  for(int k = 0; k < 5; k++) {
    const bool have = n >= (size_t)(5 - k);
    if(have) {
      const Move& m = history.moveHistory[n - 5 + k];
      moves[2 * k] = (int16_t)NNPos::locToPos(m.loc, board.x_size, nnXLen, nnYLen); moves[2 * k + 1] = m.pla;
    } else { moves[2 * k] = -1; moves[2 * k + 1] = 0; }
  }
  return kc_eval_position{stones, moves, history.numTurns, (int8_t)nextPlayer};
}

static std::shared_ptr<NNOutput> newOutput() {
#ifdef KC_IN_REFERENCE_TREE
  return std::make_shared<NNOutput>();   // NNOutput's destructor frees the owner map
#else
  return std::shared_ptr<NNOutput>(new NNOutput(), [](NNOutput* o) { delete[] o->whiteOwnerMap; delete o; });
#endif
}

void NNEvaluator::evaluateAveragedOverSymmetries(Board& board, const BoardHistory& history, Player nextPlayer, const MiscNNInputParams& nnInputParams,
                                                 const int* symmetries, int numSymmetries, NNResultBuf& buf, bool includeOwnerMap) {
  buf.hasResult = false;
  if(numSymmetries < 1 || numSymmetries > 8 || symmetries == nullptr) throw StringError("NNEvaluator::evaluateAveragedOverSymmetries: 1..8 symmetries are required");
  checkBoard(board, nnInputParams);
  kc_evaluator* ev = instanceFor(board.win_len, nnInputParams.nnPolicyTemperature);
  int8_t stones[NNPos::MAX_BOARD_LEN * NNPos::MAX_BOARD_LEN];
  int16_t moves[10];
  const kc_eval_position one = toPosition(board, history, nextPlayer, nnXLen, nnYLen, stones, moves);
  const int area = nnXLen * nnYLen;
  vector<kc_eval_position> pos((size_t)numSymmetries, one);
  vector<kc_eval_output> outs((size_t)numSymmetries);
  vector<int8_t> syms((size_t)numSymmetries);
  vector<float> policies((size_t)numSymmetries * policySize), owners(includeOwnerMap ? (size_t)numSymmetries * area : 0);
  for(int i = 0; i < numSymmetries; i++) {
    if(symmetries[i] < 0 || symmetries[i] > 7) throw StringError("NNEvaluator::evaluateAveragedOverSymmetries: symmetry out of range");
    syms[i] = (int8_t)symmetries[i];
    outs[i] = kc_eval_output{};
    outs[i].policyProbs = policies.data() + (size_t)i * policySize;
    outs[i].whiteOwnerMap = includeOwnerMap ? owners.data() + (size_t)i * area : nullptr;
  }
  if(kc_evaluator_evaluate_many(ev, numSymmetries, pos.data(), syms.data(), /*skipCache*/ 1, includeOwnerMap ? 1 : 0, outs.data()))
    throw StringError(string("B200 NNEvaluator: ") + kc_last_error());
  buf.result = newOutput();
  NNOutput& o = *buf.result;
  const float len = (float)numSymmetries;   // nninputs.cpp:95-170: plain means; the legal sets agree because it is one position
  o.nnXLen = nnXLen; o.nnYLen = nnYLen;
  o.nnHash = Hash128(outs[0].nnHash[0], outs[0].nnHash[1]);
  o.whiteWinProb = o.whiteLossProb = o.varTimeLeft = o.shorttermWinlossError = 0.0f;
  for(int i = 0; i < numSymmetries; i++) {
    o.whiteWinProb += outs[i].whiteWinProb; o.whiteLossProb += outs[i].whiteLossProb;
    o.varTimeLeft += outs[i].varTimeLeft; o.shorttermWinlossError += outs[i].shorttermWinlossError;
  }
  o.whiteWinProb /= len; o.whiteLossProb /= len; o.varTimeLeft /= len; o.shorttermWinlossError /= len;
  for(int p = 0; p < policySize; p++) {
    float sum = 0.0f;
    for(int i = 0; i < numSymmetries; i++) sum += policies[(size_t)i * policySize + p];
    o.policyProbs[p] = sum / len;
  }
  for(int p = policySize; p < NNPos::MAX_NN_POLICY_SIZE; p++) o.policyProbs[p] = -1.0f;
  o.whiteOwnerMap = nullptr;
  if(includeOwnerMap) {
    o.whiteOwnerMap = new float[(size_t)area];
    for(int p = 0; p < area; p++) {
      float sum = 0.0f;
      for(int i = 0; i < numSymmetries; i++) sum += owners[(size_t)i * area + p];
      o.whiteOwnerMap[p] = sum / len;
    }
  }
  buf.symmetry = symmetries[0];
  buf.includeOwnerMap = includeOwnerMap;
  buf.hasResult = true;
}

void NNEvaluator::evaluate(Board& board, const BoardHistory& history, Player nextPlayer, const MiscNNInputParams& nnInputParams, NNResultBuf& buf,
                           bool skipCache, bool includeOwnerMap) {
  buf.hasResult = false;
  checkBoard(board, nnInputParams);
  kc_evaluator* ev = instanceFor(board.win_len, nnInputParams.nnPolicyTemperature);

  int8_t stones[NNPos::MAX_BOARD_LEN * NNPos::MAX_BOARD_LEN];
  int16_t moves[10];
  kc_eval_position pos = toPosition(board, history, nextPlayer, nnXLen, nnYLen, stones, moves);
  buf.result = newOutput();
  NNOutput& o = *buf.result;
  o.nnXLen = nnXLen; o.nnYLen = nnYLen;
  o.whiteOwnerMap = includeOwnerMap ? new float[(size_t)nnXLen * nnYLen] : nullptr;
  kc_eval_output out{};
  out.policyProbs = o.policyProbs; out.whiteOwnerMap = o.whiteOwnerMap;
  int symmetry = nnInputParams.symmetry;   // nneval.cpp:518-528
  if(symmetry == NNInputs::SYMMETRY_NOTSPECIFIED && !currentDoRandomize.load()) symmetry = currentDefaultSymmetry.load();
  if(kc_evaluator_evaluate(ev, &pos, symmetry, skipCache ? 1 : 0, includeOwnerMap ? 1 : 0, &out)) {
    buf.result = nullptr;
    throw StringError(string("B200 NNEvaluator: ") + kc_last_error());
  }
  o.whiteWinProb = out.whiteWinProb; o.whiteLossProb = out.whiteLossProb; o.varTimeLeft = out.varTimeLeft; o.shorttermWinlossError = out.shorttermWinlossError;
  o.nnHash = Hash128(out.nnHash[0], out.nnHash[1]);
  for(int i = policySize; i < NNPos::MAX_NN_POLICY_SIZE; i++) o.policyProbs[i] = -1.0f;   // nneval.cpp:762-764
  buf.symmetry = out.symmetry;
  buf.includeOwnerMap = includeOwnerMap;
  buf.hasResult = true;
}

void NNEvaluator::clearCache() {
  lock_guard<mutex> lock(instancesMutex);
  for(Instance* i : instances) kc_evaluator_clear_cache(i->ev);
}
void NNEvaluator::clearStats() {
  lock_guard<mutex> lock(instancesMutex);
  for(Instance* i : instances) kc_evaluator_clear_stats(i->ev);
}
uint64_t NNEvaluator::numRowsProcessed() const {
  lock_guard<mutex> lock(instancesMutex);
  uint64_t n = 0;
  for(Instance* i : instances) { kc_evaluator_stats st{}; kc_evaluator_get_stats(i->ev, &st); n += st.rowsProcessed; }
  return n;
}
uint64_t NNEvaluator::numBatchesProcessed() const {
  lock_guard<mutex> lock(instancesMutex);
  uint64_t n = 0;
  for(Instance* i : instances) { kc_evaluator_stats st{}; kc_evaluator_get_stats(i->ev, &st); n += st.batchesProcessed; }
  return n;
}
double NNEvaluator::averageProcessedBatchSize() const {   // nneval.cpp:263-265
  return (double)numRowsProcessed() / (double)std::max<uint64_t>(numBatchesProcessed(), 1);
}
