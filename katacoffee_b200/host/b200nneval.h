// b200nneval.h -- class NNEvaluator with the public interface of cpp/neuralnet/nneval.h:80-175, implemented over the
// evaluator front end of libkatacoffee_b200.so (kc_evaluator_*, include/katacoffee_b200.h).  In a KataCoffee checkout this file
// replaces neuralnet/nneval.h (the Search, the commands and the tests keep compiling: same constructor, same methods, same
// exceptions); what disappears is nneval.cpp's row filling, queue, server loop, post-processing and NNCacheTable.
#pragma once
#ifdef KC_IN_REFERENCE_TREE
#include "../core/global.h"
#include "../core/logger.h"
#include "../game/board.h"
#include "../game/boardhistory.h"
#include "../neuralnet/nninputs.h"
#include "../neuralnet/nninterface.h"
#else
#include "reftypes.h"
#endif

#include <atomic>
#include <mutex>
#include <set>
#include <string>
#include <vector>

#include "katacoffee_b200.h"

class NNEvaluator {
 public:
  NNEvaluator(const std::string& modelName, const std::string& modelFileName, const std::string& expectedSha256, Logger* logger, int maxBatchSize,
              int maxConcurrentEvals, int nnXLen, int nnYLen, bool requireExactNNLen, bool inputsUseNHWC, int nnCacheSizePowerOfTwo,
              int nnMutexPoolSizePowerofTwo, bool debugSkipNeuralNet, const std::string& openCLTunerFile, const std::string& homeDataDirOverride,
              bool openCLReTunePerBoardSize, enabled_t useFP16Mode, enabled_t useNHWCMode, int numThreads, const std::vector<int>& gpuIdxByServerThread,
              const std::string& randSeed, bool doRandomize, int defaultSymmetry);
  ~NNEvaluator();
  NNEvaluator(const NNEvaluator&) = delete;
  NNEvaluator& operator=(const NNEvaluator&) = delete;

  std::string getModelName() const { return modelName; }
  std::string getModelFileName() const { return modelFileName; }
  std::string getInternalModelName() const;
  Logger* getLogger() { return logger; }
  bool isNeuralNetLess() const { return false; }
  int getMaxBatchSize() const { return maxBatchSize; }
  int getNumGpus() const { return (int)getGpuIdxs().size(); }
  int getNumServerThreads() const { return (int)gpuIdxByServerThread.size(); }
  std::set<int> getGpuIdxs() const { return std::set<int>(gpuIdxByServerThread.begin(), gpuIdxByServerThread.end()); }
  int getNNXLen() const { return nnXLen; }
  int getNNYLen() const { return nnYLen; }
  int getModelVersion() const { return modelVersion; }
  enabled_t getUsingFP16Mode() const { return usingFP16Mode; }
  enabled_t getUsingNHWCMode() const { return usingNHWCMode; }
  bool supportsShorttermError() const { return true; }

  void clearCache();
  // Blocks until the result is in buf.result (nneval.cpp:588-845); thread-safe.
  void evaluate(Board& board, const BoardHistory& history, Player nextPlayer, const MiscNNInputParams& nnInputParams, NNResultBuf& buf, bool skipCache,
                bool includeOwnerMap);
  // The root evaluation of Search::initNodeNNOutput under rootNumSymmetriesToSample > 1 (cpp/search/searchnnhelpers.cpp:67-83) in one call: the
  // position goes into the queue once per symmetry (cache skipped, as there), the rows travel in the same batch, and buf.result is their average
  // exactly as NNOutput::NNOutput(const vector<shared_ptr<NNOutput>>&) forms it (cpp/neuralnet/nninputs.cpp:95-170).  Not in the reference's
  // class: there the search thread blocks on numSymmetries consecutive evaluate() calls (which also works with this class).
  void evaluateAveragedOverSymmetries(Board& board, const BoardHistory& history, Player nextPlayer, const MiscNNInputParams& nnInputParams,
                                      const int* symmetries, int numSymmetries, NNResultBuf& buf, bool includeOwnerMap);
  void waitForNextNNEvalIfAny() {}   // only used to pace pondering threads in the reference; evaluations here never stall a caller that has none pending
  void spawnServerThreads();
  void killServerThreads();
  void setNumThreads(const std::vector<int>& gpuIdxByServerThr);
  bool isAnyThreadUsingFP16() const { return !(usingFP16Mode == enabled_t::False); }   // bf16 tensor-core path unless useFP16 = false
  bool getDoRandomize() const { return currentDoRandomize.load(); }
  int getDefaultSymmetry() const { return currentDefaultSymmetry.load(); }
  void setDoRandomize(bool b) { currentDoRandomize.store(b); }
  void setDefaultSymmetry(int s) { currentDefaultSymmetry.store(s); }
  uint64_t numRowsProcessed() const;
  uint64_t numBatchesProcessed() const;
  double averageProcessedBatchSize() const;
  void clearStats();

  // Test hook (no counterpart in the reference): evaluators created from now on call this batch function instead of the device.
  void setBackendForTesting(kc_eval_backend_fn fn, void* user) { testFn = fn; testUser = user; }

 private:
  struct Instance { int winLen; float temperature; kc_evaluator* ev; };   // one front end per (win_len, nnPolicyTemperature) seen
  kc_evaluator* instanceFor(int winLen, float temperature);
  void checkBoard(const Board& board, const MiscNNInputParams& nnInputParams) const;

  const std::string modelName, modelFileName;
  const int nnXLen, nnYLen;
  const bool requireExactNNLen;
  const int policySize;
  const enabled_t usingFP16Mode, usingNHWCMode;
  const int maxBatchSize, maxConcurrentEvals, nnCacheSizePowerOfTwo, nnMutexPoolSizePowerofTwo;
  std::vector<int> gpuIdxByServerThread;
  const uint64_t randSeedHash;
  Logger* logger;
  LoadedModel* loadedModel;
  ComputeContext* computeContext;
  int modelVersion;
  std::atomic<bool> currentDoRandomize;
  std::atomic<int> currentDefaultSymmetry;
  bool spawned;
  mutable std::mutex instancesMutex;
  std::vector<Instance*> instances;         // under instancesMutex
  std::atomic<const Instance*> lastInstance;   // fast path
  kc_eval_backend_fn testFn;
  void* testUser;
};
