// Compile-time stand-in for the reference declarations b200backend.cpp is written against.
//
// In a KataCoffee checkout the backend includes the real headers (cpp/neuralnet/nninterface.h,
// desc.h, nninputs.h, nneval.h) -- build it with -DKC_IN_REFERENCE_TREE.  Those headers do not
// compile at the surveyed revision, so this repository builds and tests the
// backend against the minimal mirror below: same type names, same member names, same meaning
// (field lists from cpp/neuralnet/desc.h:13-304, nninputs.h:75-118, nneval.h:45-65,
// core/commontypes.h:4-5).  Nothing here is used by the CUDA library itself.
#pragma once
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

struct StringError : public std::runtime_error { using std::runtime_error::runtime_error; };  // core/global.h
struct enabled_t { enum X : int { False, True, Auto }; X x; enabled_t(X a = Auto) : x(a) {} bool operator==(X o) const { return x == o; } };
class Logger;

static constexpr int ACTIVATION_IDENTITY = 0, ACTIVATION_RELU = 1, ACTIVATION_MISH = 2;  // activations.h

struct ConvLayerDesc { std::string name; int convYSize = 0, convXSize = 0, inChannels = 0, outChannels = 0, dilationY = 1, dilationX = 1; std::vector<float> weights; };
struct BatchNormLayerDesc { std::string name; int numChannels = 0; float epsilon = 0; bool hasScale = false, hasBias = false; std::vector<float> mean, variance, scale, bias; };
struct ActivationLayerDesc { std::string name; int activation = ACTIVATION_RELU; };
struct MatMulLayerDesc { std::string name; int inChannels = 0, outChannels = 0; std::vector<float> weights; };
struct MatBiasLayerDesc { std::string name; int numChannels = 0; std::vector<float> weights; };
struct ResidualBlockDesc {
  std::string name; BatchNormLayerDesc preBN; ActivationLayerDesc preActivation; ConvLayerDesc regularConv;
  BatchNormLayerDesc midBN; ActivationLayerDesc midActivation; ConvLayerDesc finalConv;
};
struct GlobalPoolingResidualBlockDesc {
  std::string name; int version = 1; BatchNormLayerDesc preBN; ActivationLayerDesc preActivation; ConvLayerDesc regularConv, gpoolConv;
  BatchNormLayerDesc gpoolBN; ActivationLayerDesc gpoolActivation; MatMulLayerDesc gpoolToBiasMul;
  BatchNormLayerDesc midBN; ActivationLayerDesc midActivation; ConvLayerDesc finalConv;
};
static constexpr int ORDINARY_BLOCK_KIND = 0, GLOBAL_POOLING_BLOCK_KIND = 2, NESTED_BOTTLENECK_BLOCK_KIND = 3;  // desc.h:173-175
using unique_ptr_void = std::unique_ptr<void, void (*)(const void*)>;
struct TrunkDesc {
  std::string name; int version = 1, numBlocks = 0, trunkNumChannels = 0, midNumChannels = 0, regularNumChannels = 0, gpoolNumChannels = 0;
  ConvLayerDesc initialConv; MatMulLayerDesc initialMatMul; std::vector<std::pair<int, unique_ptr_void>> blocks;
  BatchNormLayerDesc trunkTipBN; ActivationLayerDesc trunkTipActivation;
};
struct PolicyHeadDesc {
  std::string name; int version = 1; ConvLayerDesc p1Conv, g1Conv; BatchNormLayerDesc g1BN; ActivationLayerDesc g1Activation;
  MatMulLayerDesc gpoolToBiasMul; BatchNormLayerDesc p1BN; ActivationLayerDesc p1Activation; ConvLayerDesc p2Conv; MatMulLayerDesc gpoolToPassMul;
};
struct ValueHeadDesc {
  std::string name; int version = 1; ConvLayerDesc v1Conv; BatchNormLayerDesc v1BN; ActivationLayerDesc v1Activation; MatMulLayerDesc v2Mul;
  MatBiasLayerDesc v2Bias; ActivationLayerDesc v2Activation; MatMulLayerDesc v3Mul; MatBiasLayerDesc v3Bias, sv3Bias; MatMulLayerDesc sv3Mul; ConvLayerDesc vOwnershipConv;
};
struct ModelPostProcessParams { double varianceTimeMultiplier = 40.0, shorttermValueErrorMultiplier = 0.25; };  // desc.cpp:956-963
struct ModelDesc {
  std::string name; int version = 1, numInputChannels = 15, numInputGlobalChannels = 1;
  ModelPostProcessParams postProcessParams; TrunkDesc trunk; PolicyHeadDesc policyHead; ValueHeadDesc valueHead;
};

namespace NNPos { constexpr int MAX_BOARD_LEN = 10; constexpr int MAX_NN_POLICY_SIZE = MAX_BOARD_LEN * MAX_BOARD_LEN * 4; }  // nninputs.h:13-16
struct Hash128 { uint64_t hash0 = 0, hash1 = 0; Hash128() {} Hash128(uint64_t a, uint64_t b) : hash0(a), hash1(b) {}   // core/hash.h
                 bool operator==(const Hash128& o) const { return hash0 == o.hash0 && hash1 == o.hash1; } };
struct NNOutput {   // nninputs.h:75-118 (fields the backend and the evaluator write)
  Hash128 nnHash;
  float whiteWinProb, whiteLossProb, varTimeLeft, shorttermWinlossError;
  float policyProbs[NNPos::MAX_NN_POLICY_SIZE];
  int nnXLen, nnYLen;
  float* whiteOwnerMap = nullptr;
  float* noisedPolicyProbs = nullptr;
};
struct NNResultBuf {  // nneval.h:45-65 (fields the backend reads, and the result the evaluator hands back)
  bool hasResult = false;
  bool includeOwnerMap = false;
  int rowSpatialSize = 0, rowGlobalSize = 0;
  float* rowSpatial = nullptr; float* rowGlobal = nullptr;
  std::shared_ptr<NNOutput> result;
  int symmetry = 0; double policyOptimism = 0.0;
};

// ---- game types NNEvaluator::evaluate reads (cpp/game/board.h:24-48, 74-75, 216-222; boardhistory.h:12-31; nninputs.h:36-46) ----
typedef int8_t Player; typedef int8_t Color; typedef int8_t Direction; typedef short Spot;
static constexpr Player P_BLACK = 1, P_WHITE = 2;
static constexpr Color C_EMPTY = 0, C_BLACK = 1, C_WHITE = 2, C_WALL = 3;
static constexpr Direction D_NORTH = 0, D_WEST = 1, D_NORTHWEST = 2, D_NORTHEAST = 3, D_NONE = 4;
struct Loc { Spot spot; Direction dir; Loc() : spot(0), dir(D_NONE) {} Loc(Spot s, Direction d) : spot(s), dir(d) {} };
struct Move { Loc loc; Player pla; Move() : pla(0) {} Move(Loc l, Player p) : loc(l), pla(p) {} };
namespace Location {
inline Spot getSpot(int x, int y, int x_size) { return (Spot)((x + 1) + (y + 1) * (x_size + 1)); }
inline int getX(Spot spot, int x_size) { return spot % (x_size + 1) - 1; }
inline int getY(Spot spot, int x_size) { return spot / (x_size + 1) - 1; }
}
struct Board {
  static constexpr int MAX_LEN = 10, MAX_ARR_SIZE = (MAX_LEN + 1) * (MAX_LEN + 2) + 1;
  int x_size, y_size, win_len;
  Color colors[MAX_ARR_SIZE];
  Board(int x = 5, int y = 5, int k = 4) : x_size(x), y_size(y), win_len(k) {
    for(int i = 0; i < MAX_ARR_SIZE; i++) colors[i] = C_WALL;
    for(int yy = 0; yy < y; yy++) for(int xx = 0; xx < x; xx++) colors[Location::getSpot(xx, yy, x)] = C_EMPTY;
  }
};
struct BoardHistory { std::vector<Move> moveHistory; int numTurns = 0; bool isGameFinished = false; };
namespace NNInputs { constexpr int SYMMETRY_NOTSPECIFIED = -1; }
struct MiscNNInputParams { double playoutDoublingAdvantage = 0.0; float nnPolicyTemperature = 1.0f; int symmetry = NNInputs::SYMMETRY_NOTSPECIFIED; double policyOptimism = 0.0; };
namespace NNPos {
inline int getPolicySize(int nnXLen, int nnYLen) { return nnXLen * nnYLen * 4; }   // nninputs.cpp:47-49
inline int locToPos(Loc loc, int boardXSize, int nnXLen, int nnYLen) {            // nninputs.cpp:6-14
  return (int)loc.dir * nnXLen * nnYLen + Location::getY(loc.spot, boardXSize) * nnXLen + Location::getX(loc.spot, boardXSize);
}
}

struct ComputeContext; struct ComputeHandle; struct InputBuffers; struct LoadedModel;
namespace NeuralNet {
void globalInitialize(); void globalCleanup(); void printDevices();
LoadedModel* loadModelFile(const std::string& file, const std::string& expectedSha256);
LoadedModel* loadModelFromDesc(ModelDesc&& desc);   // standalone helper: the file parser is the reference's desc.cpp
const void* getB200ModelDescPOD(const LoadedModel*);   // the kc_model_desc view handed to the C ABI (test hook)
// the (kc_ctx*, kc_model*) pair of a GPU inside a ComputeContext (created on first use), for b200nneval.cpp
void getB200ContextAndModel(ComputeContext* context, int gpuIdx, void** kcCtx, void** kcModel);
bool getB200UseFP32Check(const ComputeContext* context);
void freeLoadedModel(LoadedModel*);
std::string getModelName(const LoadedModel*); int getModelVersion(const LoadedModel*);
ModelPostProcessParams getPostProcessParams(const LoadedModel*);
ComputeContext* createComputeContext(const std::vector<int>& gpuIdxs, Logger* logger, int nnXLen, int nnYLen, const std::string& openCLTunerFile,
                                     const std::string& homeDataDirOverride, bool openCLReTunePerBoardSize, enabled_t useFP16Mode, enabled_t useNHWCMode,
                                     const LoadedModel* loadedModel);
void freeComputeContext(ComputeContext*);
ComputeHandle* createComputeHandle(ComputeContext* context, const LoadedModel* loadedModel, Logger* logger, int maxBatchSize, bool requireExactNNLen,
                                   bool inputsUseNHWC, int gpuIdxForThisThread, int serverThreadIdx);
void freeComputeHandle(ComputeHandle*);
bool isUsingFP16(const ComputeHandle*);
InputBuffers* createInputBuffers(const LoadedModel* loadedModel, int maxBatchSize, int nnXLen, int nnYLen);
void freeInputBuffers(InputBuffers*);
void getOutput(ComputeHandle*, InputBuffers*, int numBatchEltsFilled, NNResultBuf** inputBufs, std::vector<NNOutput*>& outputs);
bool testEvaluateConv(const ConvLayerDesc*, int batchSize, int nnXLen, int nnYLen, bool useFP16, bool useNHWC, const std::vector<float>& in, std::vector<float>& out);
bool testEvaluateBatchNorm(const BatchNormLayerDesc*, int batchSize, int nnXLen, int nnYLen, bool useFP16, bool useNHWC, const std::vector<float>& in,
                           const std::vector<float>& mask, std::vector<float>& out);
bool testEvaluateResidualBlock(const ResidualBlockDesc*, int batchSize, int nnXLen, int nnYLen, bool useFP16, bool useNHWC, const std::vector<float>& in,
                               const std::vector<float>& mask, std::vector<float>& out);
bool testEvaluateGlobalPoolingResidualBlock(const GlobalPoolingResidualBlockDesc*, int batchSize, int nnXLen, int nnYLen, bool useFP16, bool useNHWC,
                                            const std::vector<float>& in, const std::vector<float>& mask, std::vector<float>& out);
}  // namespace NeuralNet
