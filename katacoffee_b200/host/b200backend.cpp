// b200backend.cpp -- the `USE_BACKEND=B200` translation unit: every free function of
// namespace NeuralNet declared in cpp/neuralnet/nninterface.h:31-171, implemented by forwarding to the
// C ABI of libkatacoffee_b200.so (include/katacoffee_b200.h).  Same role as the reference's
// eigenbackend.cpp / cudabackend.cpp; no arithmetic happens here.
//
//   in the reference tree : -DKC_IN_REFERENCE_TREE, placed at cpp/neuralnet/b200backend.cpp
//   in this repository    : compiled against host/reftypes.h (the reference headers do not compile)
//
// Error convention (nneval.cpp:327-330): C++ exceptions; a non-zero status from the C ABI is rethrown
// as StringError(kc_last_error()).
#ifdef KC_IN_REFERENCE_TREE
#include "../neuralnet/nninterface.h"
#include "../neuralnet/nneval.h"
#include "../neuralnet/modelversion.h"
#else
#include "reftypes.h"
#endif

#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>

#include "katacoffee_b200.h"

using namespace std;

static void kcCheck(int status, const char* what) {
  if(status != 0) throw StringError(string("B200 backend: ") + what + ": " + kc_last_error());
}

// ---------------------------------------------------------------------------------------------
struct LoadedModel {
  ModelDesc modelDesc;
  // POD view of modelDesc for the C ABI (pointers into modelDesc's vectors)
  vector<kc_block_desc> blocks;
  kc_model_desc pod;

  static kc_conv_desc conv(const ConvLayerDesc& d) {
    if(d.dilationX != 1 || d.dilationY != 1) throw StringError("B200 backend: dilated convolutions are not supported");
    return kc_conv_desc{d.convYSize, d.convXSize, d.inChannels, d.outChannels, d.weights.data()};
  }
  static kc_bn_desc bn(const BatchNormLayerDesc& d) {
    return kc_bn_desc{d.numChannels, d.epsilon, d.hasScale ? 1 : 0, d.hasBias ? 1 : 0, d.mean.data(), d.variance.data(),
                      d.scale.empty() ? nullptr : d.scale.data(), d.bias.empty() ? nullptr : d.bias.data()};
  }
  static kc_matmul_desc mm(const MatMulLayerDesc& d) { return kc_matmul_desc{d.inChannels, d.outChannels, d.weights.data()}; }
  static kc_matbias_desc mb(const MatBiasLayerDesc& d) { return kc_matbias_desc{d.numChannels, 0, d.weights.data()}; }
  static kc_block_desc block(const ResidualBlockDesc& b) {
    kc_block_desc k{};
    k.kind = 0; k.preActivation = b.preActivation.activation; k.midActivation = b.midActivation.activation; k.gpoolActivation = ACTIVATION_RELU;
    k.preBN = bn(b.preBN); k.regularConv = conv(b.regularConv); k.midBN = bn(b.midBN); k.finalConv = conv(b.finalConv);
    return k;
  }
  static kc_block_desc block(const GlobalPoolingResidualBlockDesc& b) {
    kc_block_desc k{};
    k.kind = 2; k.preActivation = b.preActivation.activation; k.midActivation = b.midActivation.activation; k.gpoolActivation = b.gpoolActivation.activation;
    k.preBN = bn(b.preBN); k.regularConv = conv(b.regularConv); k.gpoolConv = conv(b.gpoolConv); k.gpoolBN = bn(b.gpoolBN);
    k.gpoolToBiasMul = mm(b.gpoolToBiasMul); k.midBN = bn(b.midBN); k.finalConv = conv(b.finalConv);
    return k;
  }

  kc_modelfile* file = nullptr;   // standalone build: the parsed model file owning the weights `pod` points to
  // standalone loadModelFile: the file parser lives behind the C ABI (kc_modelfile_load); only the scalar fields of ModelDesc are filled
  explicit LoadedModel(kc_modelfile* f) : file(f) {
    pod = *kc_modelfile_desc(f);
    modelDesc.name = kc_modelfile_name(f);
    modelDesc.version = pod.version; modelDesc.numInputChannels = pod.numInputChannels; modelDesc.numInputGlobalChannels = pod.numInputGlobalChannels;
  }
  ~LoadedModel() { if(file) kc_modelfile_free(file); }
  LoadedModel(const LoadedModel&) = delete;
  LoadedModel& operator=(const LoadedModel&) = delete;

  explicit LoadedModel(ModelDesc&& d) : modelDesc(std::move(d)) {
    const TrunkDesc& t = modelDesc.trunk;
    for(const auto& b : t.blocks) {
      if(b.first == ORDINARY_BLOCK_KIND) blocks.push_back(block(*static_cast<const ResidualBlockDesc*>(b.second.get())));
      else if(b.first == GLOBAL_POOLING_BLOCK_KIND) blocks.push_back(block(*static_cast<const GlobalPoolingResidualBlockDesc*>(b.second.get())));
      else throw StringError("B200 backend: nested bottleneck blocks are not supported");
    }
    memset(&pod, 0, sizeof(pod));
    pod.version = modelDesc.version; pod.numInputChannels = modelDesc.numInputChannels; pod.numInputGlobalChannels = modelDesc.numInputGlobalChannels;
    pod.numBlocks = (int)blocks.size(); pod.trunkNumChannels = t.trunkNumChannels; pod.midNumChannels = t.midNumChannels;
    pod.regularNumChannels = t.regularNumChannels; pod.gpoolNumChannels = t.gpoolNumChannels;
    const PolicyHeadDesc& p = modelDesc.policyHead; const ValueHeadDesc& v = modelDesc.valueHead;
    pod.trunkTipActivation = t.trunkTipActivation.activation; pod.g1Activation = p.g1Activation.activation; pod.p1Activation = p.p1Activation.activation;
    pod.v1Activation = v.v1Activation.activation; pod.v2Activation = v.v2Activation.activation;
    pod.initialConv = conv(t.initialConv); pod.initialMatMul = mm(t.initialMatMul); pod.blocks = blocks.data(); pod.trunkTipBN = bn(t.trunkTipBN);
    pod.p1Conv = conv(p.p1Conv); pod.g1Conv = conv(p.g1Conv); pod.g1BN = bn(p.g1BN); pod.gpoolToBiasMul = mm(p.gpoolToBiasMul); pod.p1BN = bn(p.p1BN);
    pod.p2Conv = conv(p.p2Conv);
    pod.v1Conv = conv(v.v1Conv); pod.v1BN = bn(v.v1BN); pod.v2Mul = mm(v.v2Mul); pod.v2Bias = mb(v.v2Bias); pod.v3Mul = mm(v.v3Mul); pod.v3Bias = mb(v.v3Bias);
    pod.sv3Mul = mm(v.sv3Mul); pod.sv3Bias = mb(v.sv3Bias); pod.vOwnershipConv = conv(v.vOwnershipConv);
  }
};

struct ComputeContext {
  int nnXLen, nnYLen;
  bool useFP32Check;
  const LoadedModel* loadedModel;
  // one kc_ctx / kc_model per GPU, created on first use by a handle (weights are replicated per GPU,
  // like the reference: one server thread + handle per GPU over a shared LoadedModel, setup.cpp:190-229)
  mutex mu;
  vector<pair<int, pair<kc_ctx*, kc_model*>>> perGpu;
  pair<kc_ctx*, kc_model*> get(int gpu) {
    lock_guard<mutex> lock(mu);
    for(auto& e : perGpu) if(e.first == gpu) return e.second;
    kc_ctx* c = nullptr; kc_model* m = nullptr;
    kcCheck(kc_ctx_create(gpu, &c), "kc_ctx_create");
    kcCheck(kc_model_create(c, &loadedModel->pod, &m), "kc_model_create");
    perGpu.push_back({gpu, {c, m}});
    return {c, m};
  }
};

struct ComputeHandle {
  ComputeContext* context;
  kc_handle* handle;
  int maxBatchSize, nnXLen, nnYLen;
  bool inputsUseNHWC, usingBF16;
};

struct InputBuffers {
  int maxBatchSize, singleInputElts, singleInputGlobalElts, policySize, hw;
  // per-batch pointer tables handed to kc_forward_rows; the row staging itself is page-locked memory owned by the compute handle
  vector<const float*> spatialRows, globalRows;
  vector<float*> policyRows, scalarRows, ownerRows;
  vector<int8_t> symmetry;
};

namespace NeuralNet {

void globalInitialize() {}
void globalCleanup() {}
void printDevices() {
  int n = 0;
  if(kc_device_count(&n) != 0) n = 0;
  printf("B200 backend: %d CUDA device(s) visible\n", n);
}

LoadedModel* loadModelFromDesc(ModelDesc&& desc) { return new LoadedModel(std::move(desc)); }
const void* getB200ModelDescPOD(const LoadedModel* m) { return &m->pod; }
void getB200ContextAndModel(ComputeContext* context, int gpuIdx, void** kcCtx, void** kcModel) {
  auto cm = context->get(gpuIdx < 0 ? 0 : gpuIdx);
  *kcCtx = cm.first; *kcModel = cm.second;
}
bool getB200UseFP32Check(const ComputeContext* context) { return context->useFP32Check; }
LoadedModel* loadModelFile(const string& file, const string& expectedSha256) {
#ifdef KC_IN_REFERENCE_TREE
  ModelDesc desc;
  ModelDesc::loadFromFileMaybeGZipped(file, desc, expectedSha256);   // the reference's own parser (desc.cpp:1146-1204)
  return new LoadedModel(std::move(desc));
#else
  // standalone: the Coffee model-file parser behind the C ABI (csrc/modelfile.cpp restates desc.cpp's format)
  kc_modelfile* f = nullptr;
  kcCheck(kc_modelfile_load(file.c_str(), expectedSha256.c_str(), &f), "loadModelFile");
  return new LoadedModel(f);
#endif
}
void freeLoadedModel(LoadedModel* m) { delete m; }
string getModelName(const LoadedModel* m) { return m->modelDesc.name; }
int getModelVersion(const LoadedModel* m) { return m->modelDesc.version; }
ModelPostProcessParams getPostProcessParams(const LoadedModel* m) { return m->modelDesc.postProcessParams; }

ComputeContext* createComputeContext(const vector<int>& gpuIdxs, Logger* logger, int nnXLen, int nnYLen, const string& openCLTunerFile,
                                     const string& homeDataDirOverride, bool openCLReTunePerBoardSize, enabled_t useFP16Mode, enabled_t useNHWCMode,
                                     const LoadedModel* loadedModel) {
  (void)gpuIdxs; (void)logger; (void)openCLTunerFile; (void)homeDataDirOverride; (void)openCLReTunePerBoardSize; (void)useNHWCMode;
  ComputeContext* c = new ComputeContext();
  c->nnXLen = nnXLen; c->nnYLen = nnYLen; c->loadedModel = loadedModel;
  // useFP16 = false selects the fp32 check path; true / auto select the bf16 tensor-core path
  c->useFP32Check = (useFP16Mode == enabled_t::False);
  return c;
}
void freeComputeContext(ComputeContext* c) {
  if(!c) return;
  for(auto& e : c->perGpu) { kc_model_destroy(e.second.second); kc_ctx_destroy(e.second.first); }
  delete c;
}

ComputeHandle* createComputeHandle(ComputeContext* context, const LoadedModel* loadedModel, Logger* logger, int maxBatchSize, bool requireExactNNLen,
                                   bool inputsUseNHWC, int gpuIdxForThisThread, int serverThreadIdx) {
  (void)logger; (void)serverThreadIdx; (void)loadedModel;
  auto cm = context->get(gpuIdxForThisThread < 0 ? 0 : gpuIdxForThisThread);
  // requireExactNNLen = false: boards may be smaller than the net's slot; the tensor-core kernel then takes input channel 0 as the mask
  // (the fp32 check path always does), as the reference backends do (nninterface.h:73-76, eigenbackend.cpp:1438)
  unsigned flags = (context->useFP32Check ? KC_FLAG_FP32_CHECK : 0u) | (inputsUseNHWC ? KC_FLAG_INPUTS_NHWC : 0u) | (requireExactNNLen ? 0u : KC_FLAG_MASKED_BOARDS);
  ComputeHandle* h = new ComputeHandle();
  h->context = context; h->maxBatchSize = maxBatchSize; h->nnXLen = context->nnXLen; h->nnYLen = context->nnYLen; h->inputsUseNHWC = inputsUseNHWC;
  kcCheck(kc_handle_create(cm.first, cm.second, maxBatchSize, context->nnXLen, context->nnYLen, flags, &h->handle), "kc_handle_create");
  h->usingBF16 = kc_handle_uses_bf16(h->handle) != 0;
  return h;
}
void freeComputeHandle(ComputeHandle* h) { if(h) { kc_handle_destroy(h->handle); delete h; } }
bool isUsingFP16(const ComputeHandle* h) { return h->usingBF16; }

InputBuffers* createInputBuffers(const LoadedModel* loadedModel, int maxBatchSize, int nnXLen, int nnYLen) {
  InputBuffers* b = new InputBuffers();
  b->maxBatchSize = maxBatchSize; b->hw = nnXLen * nnYLen;
  b->singleInputElts = loadedModel->modelDesc.numInputChannels * b->hw;
  b->singleInputGlobalElts = loadedModel->modelDesc.numInputGlobalChannels;
  b->policySize = 4 * b->hw;
  b->symmetry.resize(maxBatchSize);   // the row staging itself is page-locked memory owned by the compute handle (kc_forward_rows)
  return b;
}
void freeInputBuffers(InputBuffers* b) { delete b; }

void getOutput(ComputeHandle* h, InputBuffers* b, int numBatchEltsFilled, NNResultBuf** inputBufs, vector<NNOutput*>& outputs) {
  const int n = numBatchEltsFilled;
  if(n <= 0 || n > b->maxBatchSize || (int)outputs.size() != n) throw StringError("B200 backend: getOutput called with a bad batch size");
  for(int i = 0; i < n; i++)
    if(inputBufs[i]->rowSpatialSize != b->singleInputElts || inputBufs[i]->rowGlobalSize != b->singleInputGlobalElts)
      throw StringError("B200 backend: row sizes do not match the model");
  // Rows and outputs stay where they are: the C ABI takes the NNResultBuf / NNOutput pointers and gathers, copies, evaluates and
  // scatters chunk by chunk (kc_forward_rows), so the host copies every backend makes here (e.g. eigenbackend.cpp:1700-1730,
  // 1755-1843) overlap the device work instead of running before and after it.
  static_assert(offsetof(NNOutput, whiteLossProb) == offsetof(NNOutput, whiteWinProb) + 4 && offsetof(NNOutput, varTimeLeft) == offsetof(NNOutput, whiteWinProb) + 8 &&
                offsetof(NNOutput, shorttermWinlossError) == offsetof(NNOutput, whiteWinProb) + 12, "NNOutput's four scalars are contiguous (nninputs.h:75-90)");
  b->spatialRows.resize(n); b->globalRows.resize(n); b->policyRows.resize(n); b->scalarRows.resize(n); b->ownerRows.resize(n);
  bool anyOwner = false;
  for(int i = 0; i < n; i++) {
    const NNResultBuf* r = inputBufs[i];
    NNOutput* o = outputs[i];
    b->spatialRows[i] = r->rowSpatial; b->globalRows[i] = r->rowGlobal; b->symmetry[i] = (int8_t)r->symmetry;
    // logits, already in NNPos order and inverse-symmetrised; nnHash / noisedPolicyProbs are not touched (eigenbackend.cpp:1765-1767)
    b->policyRows[i] = o->policyProbs; b->scalarRows[i] = &o->whiteWinProb; b->ownerRows[i] = o->whiteOwnerMap;
    anyOwner = anyOwner || o->whiteOwnerMap != nullptr;
    o->nnXLen = h->nnXLen; o->nnYLen = h->nnYLen;
  }
  kcCheck(kc_forward_rows(h->handle, n, b->spatialRows.data(), b->globalRows.data(), b->symmetry.data(), b->policyRows.data(), b->scalarRows.data(),
                          anyOwner ? b->ownerRows.data() : nullptr), "kc_forward_rows");
}

static kc_ctx* testCtx() {
  static kc_ctx* c = nullptr;
  if(!c) kcCheck(kc_ctx_create(0, &c), "kc_ctx_create");
  return c;
}
bool testEvaluateConv(const ConvLayerDesc* d, int batchSize, int nnXLen, int nnYLen, bool useFP16, bool useNHWC, const vector<float>& in, vector<float>& out) {
  if(useFP16) return false;   // layer hooks run on the fp32 check path only
  kc_conv_desc k = LoadedModel::conv(*d);
  out.resize((size_t)batchSize * nnXLen * nnYLen * d->outChannels);
  kcCheck(kc_test_conv(testCtx(), &k, batchSize, nnXLen, nnYLen, useNHWC, in.data(), out.data()), "kc_test_conv");
  return true;
}
bool testEvaluateBatchNorm(const BatchNormLayerDesc* d, int batchSize, int nnXLen, int nnYLen, bool useFP16, bool useNHWC, const vector<float>& in,
                           const vector<float>& mask, vector<float>& out) {
  if(useFP16) return false;
  kc_bn_desc k = LoadedModel::bn(*d);
  out.resize(in.size());
  kcCheck(kc_test_batchnorm(testCtx(), &k, ACTIVATION_IDENTITY, batchSize, nnXLen, nnYLen, useNHWC, in.data(), mask.data(), out.data()), "kc_test_batchnorm");
  return true;
}
bool testEvaluateResidualBlock(const ResidualBlockDesc* d, int batchSize, int nnXLen, int nnYLen, bool useFP16, bool useNHWC, const vector<float>& in,
                               const vector<float>& mask, vector<float>& out) {
  if(useFP16) return false;
  kc_block_desc k = LoadedModel::block(*d);
  out.resize(in.size());
  kcCheck(kc_test_resblock(testCtx(), &k, batchSize, nnXLen, nnYLen, useNHWC, in.data(), mask.data(), out.data()), "kc_test_resblock");
  return true;
}
bool testEvaluateGlobalPoolingResidualBlock(const GlobalPoolingResidualBlockDesc* d, int batchSize, int nnXLen, int nnYLen, bool useFP16, bool useNHWC,
                                            const vector<float>& in, const vector<float>& mask, vector<float>& out) {
  if(useFP16) return false;
  kc_block_desc k = LoadedModel::block(*d);
  out.resize(in.size());
  kcCheck(kc_test_resblock(testCtx(), &k, batchSize, nnXLen, nnYLen, useNHWC, in.data(), mask.data(), out.data()), "kc_test_resblock");
  return true;
}

}  // namespace NeuralNet
