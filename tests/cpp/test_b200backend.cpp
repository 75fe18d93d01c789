// Exercises katacoffee_b200/host/b200backend.cpp the way the reference's NNEvaluator::serve and
// tests/testnn.cpp drive a backend (nneval.cpp:386-567, testnn.cpp:105-135): loadModel ->
// createComputeContext -> createComputeHandle -> createInputBuffers -> getOutput, and testEvaluateConv.
// Checks: (1) the shim's outputs equal a direct kc_forward call bit for bit, for the bf16 and the
// fp32 path, with per-row symmetry and with/without the owner map; (2) the fp32 and bf16 paths agree
// to bf16 tolerance; (3) the literal 1x1-convolution vector of testnn.cpp:161-207.  Exit code 0 = pass.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "reftypes.h"
#include "katacoffee_b200.h"

static uint64_t rngState = 88172645463325252ULL;
static float rnd() {   // xorshift, uniform in (-1, 1)
  rngState ^= rngState << 13; rngState ^= rngState >> 7; rngState ^= rngState << 17;
  return (float)((rngState >> 11) * (1.0 / 9007199254740992.0)) * 2.0f - 1.0f;
}
static ConvLayerDesc conv(int k, int ic, int oc, float gain = 1.0f) {
  ConvLayerDesc d; d.convYSize = d.convXSize = k; d.inChannels = ic; d.outChannels = oc;
  d.weights.resize((size_t)k * k * ic * oc);
  float s = gain * std::sqrt(6.0f / (k * k * ic));
  for(float& w : d.weights) w = rnd() * s;
  return d;
}
static BatchNormLayerDesc bn(int c) {
  BatchNormLayerDesc d; d.numChannels = c; d.epsilon = 1e-4f; d.hasScale = true; d.hasBias = true;
  d.mean.assign(c, 0.f); d.variance.assign(c, 1.f); d.scale.assign(c, 1.f); d.bias.resize(c);
  for(int i = 0; i < c; i++) { d.mean[i] = rnd() * 0.1f; d.variance[i] = 0.8f + 0.4f * std::fabs(rnd()); d.bias[i] = rnd() * 0.1f; }
  return d;
}
static MatMulLayerDesc mm(int ic, int oc) {
  MatMulLayerDesc d; d.inChannels = ic; d.outChannels = oc; d.weights.resize((size_t)ic * oc);
  float s = std::sqrt(3.0f / ic);
  for(float& w : d.weights) w = rnd() * s;
  return d;
}
static MatBiasLayerDesc mb(int c) { MatBiasLayerDesc d; d.numChannels = c; d.weights.resize(c); for(float& w : d.weights) w = rnd() * 0.1f; return d; }

static ModelDesc makeModel(int C, int mid, int reg, int gp, int nblocks) {
  ModelDesc m; m.name = "cpp-test"; m.version = 1;
  TrunkDesc& t = m.trunk;
  t.numBlocks = nblocks; t.trunkNumChannels = C; t.midNumChannels = mid; t.regularNumChannels = reg; t.gpoolNumChannels = gp;
  t.initialConv = conv(3, 15, C); t.initialMatMul = mm(1, C);
  for(int i = 0; i < nblocks; i++) {
    if(i % 2 == 1) {
      auto* b = new GlobalPoolingResidualBlockDesc();
      b->preBN = bn(C); b->regularConv = conv(3, C, reg); b->gpoolConv = conv(3, C, gp); b->gpoolBN = bn(gp); b->gpoolToBiasMul = mm(3 * gp, reg);
      b->midBN = bn(reg); b->finalConv = conv(3, reg, C, 0.5f);
      t.blocks.emplace_back(GLOBAL_POOLING_BLOCK_KIND, unique_ptr_void(b, [](const void* p) { delete static_cast<const GlobalPoolingResidualBlockDesc*>(p); }));
    } else {
      auto* b = new ResidualBlockDesc();
      b->preBN = bn(C); b->regularConv = conv(3, C, mid); b->midBN = bn(mid); b->finalConv = conv(3, mid, C, 0.5f);
      t.blocks.emplace_back(ORDINARY_BLOCK_KIND, unique_ptr_void(b, [](const void* p) { delete static_cast<const ResidualBlockDesc*>(p); }));
    }
  }
  t.trunkTipBN = bn(C);
  PolicyHeadDesc& p = m.policyHead;
  p.p1Conv = conv(1, C, 32); p.g1Conv = conv(1, C, 32); p.g1BN = bn(32); p.gpoolToBiasMul = mm(96, 32); p.p1BN = bn(32); p.p2Conv = conv(1, 32, 4);
  ValueHeadDesc& v = m.valueHead;
  v.v1Conv = conv(1, C, 32); v.v1BN = bn(32); v.v2Mul = mm(96, 48); v.v2Bias = mb(48); v.v3Mul = mm(48, 2); v.v3Bias = mb(2);
  v.sv3Mul = mm(48, 2); v.sv3Bias = mb(2); v.vOwnershipConv = conv(1, 32, 1);
  return m;
}

#define REQUIRE(c) do { if(!(c)) { printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); return 1; } } while(0)

int main() {
  const int W = 5, H = 5, HW = 25, N = 37;
  NeuralNet::globalInitialize();
  LoadedModel* model = NeuralNet::loadModelFromDesc(makeModel(64, 64, 48, 16, 3));
  REQUIRE(NeuralNet::getModelVersion(model) == 1 && NeuralNet::getModelName(model) == "cpp-test");
  // inputs: plausible V1 rows (channel 0 = on-board mask, 0/1 planes), NCHW
  std::vector<std::vector<float>> spatial(N, std::vector<float>(15 * HW)), global(N, std::vector<float>(1, 4.0f));
  for(int i = 0; i < N; i++)
    for(int c = 0; c < 15; c++)
      for(int p = 0; p < HW; p++) spatial[i][c * HW + p] = c == 0 ? 1.0f : (rnd() > 0.4f ? 1.0f : 0.0f);
  std::vector<float> results[2];
  for(int fp32 = 0; fp32 < 2; fp32++) {
    ComputeContext* ctx = NeuralNet::createComputeContext({0}, nullptr, W, H, "", "", false, fp32 ? enabled_t::False : enabled_t::True, enabled_t::Auto, model);
    ComputeHandle* h = NeuralNet::createComputeHandle(ctx, model, nullptr, 64, true, false, 0, 0);
    REQUIRE(NeuralNet::isUsingFP16(h) == !fp32);
    InputBuffers* ib = NeuralNet::createInputBuffers(model, 64, W, H);
    std::vector<NNResultBuf> bufs(N);
    std::vector<NNResultBuf*> bufPtrs(N);
    std::vector<NNOutput> outs(N);
    std::vector<NNOutput*> outPtrs(N);
    std::vector<std::vector<float>> owner(N, std::vector<float>(HW));
    for(int i = 0; i < N; i++) {
      bufs[i].rowSpatial = spatial[i].data(); bufs[i].rowGlobal = global[i].data();
      bufs[i].rowSpatialSize = 15 * HW; bufs[i].rowGlobalSize = 1; bufs[i].symmetry = i % 8;
      outs[i].whiteOwnerMap = (i % 2) ? owner[i].data() : nullptr;
      bufPtrs[i] = &bufs[i]; outPtrs[i] = &outs[i];
    }
    NeuralNet::getOutput(h, ib, N, bufPtrs.data(), outPtrs);
    // the same through the C ABI directly
    kc_ctx* kc = nullptr; kc_model* km = nullptr; kc_handle* kh = nullptr;
    REQUIRE(kc_ctx_create(0, &kc) == 0);
    REQUIRE(kc_model_create(kc, static_cast<const kc_model_desc*>(NeuralNet::getB200ModelDescPOD(model)), &km) == 0);
    REQUIRE(kc_handle_create(kc, km, 64, W, H, fp32 ? KC_FLAG_FP32_CHECK : 0u, &kh) == 0);
    std::vector<float> sp((size_t)N * 15 * HW), gl(N, 4.0f), pol((size_t)N * 4 * HW), val(2 * N), misc(2 * N), own((size_t)N * HW);
    std::vector<int8_t> sym(N);
    for(int i = 0; i < N; i++) { memcpy(&sp[(size_t)i * 15 * HW], spatial[i].data(), 15 * HW * 4); sym[i] = (int8_t)(i % 8); }
    REQUIRE(kc_forward(kh, N, sp.data(), gl.data(), sym.data(), pol.data(), val.data(), misc.data(), own.data()) == 0);
    for(int i = 0; i < N; i++) {
      REQUIRE(memcmp(outs[i].policyProbs, &pol[(size_t)i * 4 * HW], 4 * HW * 4) == 0);
      REQUIRE(outs[i].whiteWinProb == val[2 * i] && outs[i].whiteLossProb == val[2 * i + 1]);
      REQUIRE(outs[i].varTimeLeft == misc[2 * i] && outs[i].shorttermWinlossError == misc[2 * i + 1]);
      REQUIRE(outs[i].nnXLen == W && outs[i].nnYLen == H);
      if(i % 2) REQUIRE(memcmp(owner[i].data(), &own[(size_t)i * HW], HW * 4) == 0);
      for(int k = 0; k < 4 * HW; k++) REQUIRE(std::isfinite(outs[i].policyProbs[k]));
    }
    results[fp32] = pol;
    results[fp32].insert(results[fp32].end(), val.begin(), val.end());
    kc_handle_destroy(kh); kc_model_destroy(km); kc_ctx_destroy(kc);
    NeuralNet::freeInputBuffers(ib); NeuralNet::freeComputeHandle(h); NeuralNet::freeComputeContext(ctx);
  }
  double maxDiff = 0, maxAbs = 0;
  for(size_t i = 0; i < results[0].size(); i++) {
    maxDiff = std::fmax(maxDiff, std::fabs(results[0][i] - results[1][i]));
    maxAbs = std::fmax(maxAbs, std::fabs(results[1][i]));
  }
  printf("bf16 vs fp32 path: max abs diff %.4g (max |logit| %.3g)\n", maxDiff, maxAbs);
  REQUIRE(maxAbs > 0.1 && maxDiff < 0.03 * std::fmax(maxAbs, 3.0));
  {
    // testnn.cpp:161-207 "1x1 convolution"
    ConvLayerDesc d; d.convYSize = d.convXSize = 1; d.inChannels = 2; d.outChannels = 3; d.weights = {0, 1, 1, -1, 10, 0.1f};
    std::vector<float> in = {5,5,4,4, 5,5,4,4, 1,1,8,8,  0,1,2,3, 3,4,5,6, 8,7,6,5,  0,1,0,2, 3,0,4,0, 0,5,0,6,  1,0,0,2, 0,2,2,0, 0,2,2,0};
    std::vector<float> expectFirst = {0, 1, 2, 3, 3, 4, 5, 6, 8, 7, 6, 5, 5, 4, 2, 1};
    std::vector<float> out;
    REQUIRE(NeuralNet::testEvaluateConv(&d, 2, 4, 3, false, false, in, out));
    REQUIRE(out.size() == 72);
    for(size_t i = 0; i < expectFirst.size(); i++) REQUIRE(std::fabs(out[i] - expectFirst[i]) < 1e-4);
    REQUIRE(!NeuralNet::testEvaluateConv(&d, 2, 4, 3, true, false, in, out));   // fp16 layer hooks: unsupported -> false
  }
  {
    // throughput of the drop-in call itself at a self-play sized batch: rows gathered from NNResultBufs, one getOutput, logits
    // scattered into NNOutputs (b10c128-shaped trunk, bf16 path)
    const int B = 148 * 128;
    LoadedModel* big = NeuralNet::loadModelFromDesc(makeModel(128, 128, 96, 32, 10));
    ComputeContext* ctx = NeuralNet::createComputeContext({0}, nullptr, W, H, "", "", false, enabled_t::True, enabled_t::Auto, big);
    ComputeHandle* h = NeuralNet::createComputeHandle(ctx, big, nullptr, B, true, false, 0, 0);
    InputBuffers* ib = NeuralNet::createInputBuffers(big, B, W, H);
    std::vector<NNResultBuf> bufs(B); std::vector<NNResultBuf*> bufPtrs(B);
    std::vector<NNOutput> outs(B); std::vector<NNOutput*> outPtrs(B);
    for(int i = 0; i < B; i++) {
      bufs[i].rowSpatial = spatial[i % N].data(); bufs[i].rowGlobal = global[i % N].data();
      bufs[i].rowSpatialSize = 15 * HW; bufs[i].rowGlobalSize = 1; bufs[i].symmetry = i % 8;
      bufPtrs[i] = &bufs[i]; outPtrs[i] = &outs[i];
    }
    for(int rep = 0; rep < 3; rep++) NeuralNet::getOutput(h, ib, B, bufPtrs.data(), outPtrs);
    const auto t0 = std::chrono::steady_clock::now();
    const int reps = 10;
    for(int rep = 0; rep < reps; rep++) NeuralNet::getOutput(h, ib, B, bufPtrs.data(), outPtrs);
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("NeuralNet::getOutput, %d rows per call (b10c128 shape, bf16): %.3f ms per call, %.3f M evals/s incl. the host gather / scatter\n", B,
           sec / reps * 1e3, B * reps / sec / 1e6);
    for(int i = 0; i < N; i++) REQUIRE(memcmp(outs[i].policyProbs, outs[i + N * 8].policyProbs, 4 * HW * 4) == 0);   // same row, same symmetry
    {
      // the pipelined, multi-threaded gather / scatter (kc_forward_rows) against one plain kc_forward over the gathered rows: every
      // row of the large batch bit for bit, with owner maps on every third row
      std::vector<std::vector<float>> owner(B);
      for(int i = 0; i < B; i += 3) { owner[i].assign(HW, -7.f); outs[i].whiteOwnerMap = owner[i].data(); }
      NeuralNet::getOutput(h, ib, B, bufPtrs.data(), outPtrs);
      kc_ctx* kc = nullptr; kc_model* km = nullptr; kc_handle* kh = nullptr;
      REQUIRE(kc_ctx_create(0, &kc) == 0);
      REQUIRE(kc_model_create(kc, static_cast<const kc_model_desc*>(NeuralNet::getB200ModelDescPOD(big)), &km) == 0);
      REQUIRE(kc_handle_create(kc, km, B, W, H, 0u, &kh) == 0);
      std::vector<float> sp((size_t)B * 15 * HW), gl(B, 4.0f), pol((size_t)B * 4 * HW), val(2 * B), misc(2 * B), own((size_t)B * HW);
      std::vector<int8_t> sym(B);
      for(int i = 0; i < B; i++) { memcpy(&sp[(size_t)i * 15 * HW], spatial[i % N].data(), 15 * HW * 4); sym[i] = (int8_t)(i % 8); }
      REQUIRE(kc_forward(kh, B, sp.data(), gl.data(), sym.data(), pol.data(), val.data(), misc.data(), own.data()) == 0);
      for(int i = 0; i < B; i++) {
        REQUIRE(memcmp(outs[i].policyProbs, &pol[(size_t)i * 4 * HW], 4 * HW * 4) == 0);
        REQUIRE(outs[i].whiteWinProb == val[2 * i] && outs[i].whiteLossProb == val[2 * i + 1]);
        REQUIRE(outs[i].varTimeLeft == misc[2 * i] && outs[i].shorttermWinlossError == misc[2 * i + 1]);
        if(i % 3 == 0) REQUIRE(memcmp(owner[i].data(), &own[(size_t)i * HW], HW * 4) == 0);
      }
      for(int i = 0; i < B; i += 3) outs[i].whiteOwnerMap = nullptr;
      kc_handle_destroy(kh); kc_model_destroy(km); kc_ctx_destroy(kc);
    }
    NeuralNet::freeInputBuffers(ib); NeuralNet::freeComputeHandle(h); NeuralNet::freeComputeContext(ctx); NeuralNet::freeLoadedModel(big);
  }
  bool threw = false;
  try { NeuralNet::loadModelFile("nonexistent.bin.gz", ""); } catch(const StringError&) { threw = true; }
  REQUIRE(threw);
  {
    // NeuralNet::loadModelFile: write the model in the reference's file format, load it back through the backend API and
    // check that the loaded model evaluates bit-identically to the in-memory one
    const char* path = "/tmp/kc_cpp_test_model.bin.gz";
    REQUIRE(kc_modelfile_write(static_cast<const kc_model_desc*>(NeuralNet::getB200ModelDescPOD(model)), "cpp-test-file", path) == 0);
    LoadedModel* fromFile = NeuralNet::loadModelFile(path, "");
    REQUIRE(NeuralNet::getModelName(fromFile) == "cpp-test-file" && NeuralNet::getModelVersion(fromFile) == 1);
    std::vector<float> pols[2];
    LoadedModel* ms[2] = {model, fromFile};
    for(int k = 0; k < 2; k++) {
      ComputeContext* ctx = NeuralNet::createComputeContext({0}, nullptr, W, H, "", "", false, enabled_t::True, enabled_t::Auto, ms[k]);
      ComputeHandle* h = NeuralNet::createComputeHandle(ctx, ms[k], nullptr, 64, true, false, 0, 0);
      InputBuffers* ib = NeuralNet::createInputBuffers(ms[k], 64, W, H);
      std::vector<NNResultBuf> bufs(N); std::vector<NNResultBuf*> bufPtrs(N);
      std::vector<NNOutput> outs(N); std::vector<NNOutput*> outPtrs(N);
      for(int i = 0; i < N; i++) {
        bufs[i].rowSpatial = spatial[i].data(); bufs[i].rowGlobal = global[i].data();
        bufs[i].rowSpatialSize = 15 * HW; bufs[i].rowGlobalSize = 1; bufs[i].symmetry = i % 8;
        bufPtrs[i] = &bufs[i]; outPtrs[i] = &outs[i];
      }
      NeuralNet::getOutput(h, ib, N, bufPtrs.data(), outPtrs);
      for(int i = 0; i < N; i++) { pols[k].insert(pols[k].end(), outs[i].policyProbs, outs[i].policyProbs + 4 * HW); pols[k].push_back(outs[i].whiteWinProb); }
      NeuralNet::freeInputBuffers(ib); NeuralNet::freeComputeHandle(h); NeuralNet::freeComputeContext(ctx);
    }
    REQUIRE(pols[0].size() == pols[1].size() && memcmp(pols[0].data(), pols[1].data(), pols[0].size() * 4) == 0);
    std::string sha;
    { kc_modelfile* f = nullptr; REQUIRE(kc_modelfile_load(path, nullptr, &f) == 0); sha = kc_modelfile_sha256(f); kc_modelfile_free(f); }
    NeuralNet::freeLoadedModel(NeuralNet::loadModelFile(path, sha));   // the right hash passes
    threw = false;
    try { NeuralNet::loadModelFile(path, std::string(64, 'a')); } catch(const StringError&) { threw = true; }
    REQUIRE(threw);
    NeuralNet::freeLoadedModel(fromFile);
    remove(path);
  }
  NeuralNet::freeLoadedModel(model);
  NeuralNet::globalCleanup();
  printf("b200backend ok\n");
  return 0;
}
