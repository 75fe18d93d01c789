// Native throughput / latency driver of the drop-in call itself: NeuralNet::getOutput (katacoffee_b200/host/b200backend.cpp) fed
// from NNResultBufs and writing NNOutputs, exactly as NNEvaluator::serve calls a backend (cpp/neuralnet/nneval.cpp:386-567).
//   bench_getoutput MODEL.bin.gz --size 5 --batches 1024,18944 [--reps 200]
// Prints one JSON object: per batch size the calls per second, evals/s and p50 / p99 latency of a call (host rows in, host outputs
// out: gather, H2D, kernels, D2H and scatter all inside).  BASELINE.json configs[2] is the 1024-row case.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "reftypes.h"
#include "katacoffee_b200.h"

int main(int argc, char** argv) {
  if(argc < 2) { fprintf(stderr, "usage: bench_getoutput MODEL [--size N] [--batches a,b,..] [--reps R]\n"); return 2; }
  int size = 5, reps = 200;
  std::vector<int> batches = {1024};
  for(int i = 2; i + 1 < argc; i += 2) {
    const std::string k = argv[i];
    if(k == "--size") size = atoi(argv[i + 1]);
    else if(k == "--reps") reps = atoi(argv[i + 1]);
    else if(k == "--batches") {
      batches.clear();
      for(const char* q = argv[i + 1]; *q;) { batches.push_back(atoi(q)); while(*q && *q != ',') q++; if(*q == ',') q++; }
    }
  }
  try {
    const int W = size, H = size, HW = W * H;
    NeuralNet::globalInitialize();
    LoadedModel* model = NeuralNet::loadModelFile(argv[1], "");
    const int maxB = *std::max_element(batches.begin(), batches.end());
    ComputeContext* ctx = NeuralNet::createComputeContext({0}, nullptr, W, H, "", "", false, enabled_t::True, enabled_t::Auto, model);
    ComputeHandle* h = NeuralNet::createComputeHandle(ctx, model, nullptr, maxB, true, false, 0, 0);
    InputBuffers* ib = NeuralNet::createInputBuffers(model, maxB, W, H);
    // 64 distinct plausible V1 rows (channel 0 = the on-board mask), shared by the requests like positions of concurrent games
    uint64_t st = 88172645463325252ULL;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
    std::vector<std::vector<float>> spatial(64, std::vector<float>(15 * HW)), global(64, std::vector<float>(1, 4.0f));
    for(auto& row : spatial)
      for(int c = 0; c < 15; c++)
        for(int p = 0; p < HW; p++) row[c * HW + p] = c == 0 ? 1.0f : (rnd() < 0.25 ? 1.0f : 0.0f);
    std::vector<NNResultBuf> bufs(maxB); std::vector<NNResultBuf*> bufPtrs(maxB);
    std::vector<NNOutput> outs(maxB); std::vector<NNOutput*> outPtrs(maxB);
    for(int i = 0; i < maxB; i++) {
      bufs[i].rowSpatial = spatial[i % 64].data(); bufs[i].rowGlobal = global[i % 64].data();
      bufs[i].rowSpatialSize = 15 * HW; bufs[i].rowGlobalSize = 1; bufs[i].symmetry = i % 8;
      bufPtrs[i] = &bufs[i]; outPtrs[i] = &outs[i];
    }
    printf("{\"api\": \"NeuralNet::getOutput (C++ shim over kc_forward_rows): NNResultBuf rows in host memory -> NNOutput logits in host memory\", \"board\": \"%dx%d\", \"batches\": [", W, H);
    for(size_t bi = 0; bi < batches.size(); bi++) {
      const int B = batches[bi];
      std::vector<NNOutput*> op(outPtrs.begin(), outPtrs.begin() + B);
      const int r = B >= 8192 ? std::max(10, reps / 10) : reps;
      for(int k = 0; k < 5; k++) NeuralNet::getOutput(h, ib, B, bufPtrs.data(), op);
      std::vector<double> lat(r);
      const auto t0 = std::chrono::steady_clock::now();
      for(int k = 0; k < r; k++) {
        const auto a = std::chrono::steady_clock::now();
        NeuralNet::getOutput(h, ib, B, bufPtrs.data(), op);
        lat[k] = std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
      }
      const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      std::sort(lat.begin(), lat.end());
      printf("%s{\"rows_per_call\": %d, \"calls\": %d, \"evals_per_s\": %.1f, \"ms_per_call_mean\": %.4f, \"ms_p50\": %.4f, \"ms_p99\": %.4f, "
             "\"h2d_bytes_per_call\": %lld, \"d2h_bytes_per_call\": %lld}",
             bi ? ", " : "", B, r, B * r / sec, sec / r * 1e3, lat[r / 2] * 1e3, lat[std::min(r - 1, (int)(r * 0.99))] * 1e3,
             (long long)B * (15 * HW * 4 + 4 + 1), (long long)B * (4 * HW * 4 + 16));
    }
    printf("]}\n");
    NeuralNet::freeInputBuffers(ib); NeuralNet::freeComputeHandle(h); NeuralNet::freeComputeContext(ctx); NeuralNet::freeLoadedModel(model);
  } catch(const StringError& e) {
    fprintf(stderr, "bench_getoutput: %s\n", e.what());
    return 1;
  }
  return 0;
}
