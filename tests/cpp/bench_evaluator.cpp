// Throughput of the evaluator front end on the device (kc_evaluator_create): T native client threads submit positions
// (kc_evaluator_evaluate_many in chunks, or single blocking kc_evaluator_evaluate calls with --single) to S server threads.
// Positions are random stone placements with no forced line (every one has legal moves), all distinct unless --repeat R
// makes each client cycle through R positions (cache hits).  Prints one JSON line; correctness is tests/test_z_gpu_evaluator.py's job.
//
//   bench_evaluator <model file> [--clients T] [--rows N per client] [--batch B] [--servers S] [--cache P] [--single] [--repeat R] [--chunk C]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "katacoffee_b200.h"

static uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}

#define CHECK(call)                                                              \
  do {                                                                           \
    if((call) != 0) { fprintf(stderr, "%s: %s\n", #call, kc_last_error()); return 2; } \
  } while(0)

int main(int argc, char** argv) {
  if(argc < 2) { fprintf(stderr, "usage: bench_evaluator <model file> [options]\n"); return 2; }
  int clients = 8, rows = 200000, batch = 4096, servers = 2, cachePow = -1, single = 0, repeat = 0, chunk = 512, W = 5, H = 5, K = 4;
  for(int i = 2; i < argc; i++) {
    const std::string a = argv[i];
    auto val = [&] { return i + 1 < argc ? atoi(argv[++i]) : 0; };
    if(a == "--clients") clients = val(); else if(a == "--rows") rows = val(); else if(a == "--batch") batch = val();
    else if(a == "--servers") servers = val(); else if(a == "--cache") cachePow = val(); else if(a == "--single") single = 1;
    else if(a == "--repeat") repeat = val(); else if(a == "--chunk") chunk = val(); else if(a == "--size") { W = H = val(); }
    else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 2; }
  }
  const int HW = W * H, P = 4 * HW;
  kc_ctx* ctx = nullptr; kc_modelfile* mf = nullptr; kc_model* model = nullptr; kc_evaluator* ev = nullptr;
  CHECK(kc_ctx_create(0, &ctx));
  CHECK(kc_modelfile_load(argv[1], "", &mf));
  CHECK(kc_model_create(ctx, kc_modelfile_desc(mf), &model));
  kc_evaluator_config cfg{};
  cfg.nnXLen = W; cfg.nnYLen = H; cfg.winLen = K; cfg.maxBatch = batch; cfg.maxConcurrentEvals = clients * (single ? 1 : chunk) + batch;
  cfg.numServerThreads = servers; cfg.cacheSizePowerOfTwo = cachePow; cfg.mutexPoolSizePowerOfTwo = 12;
  cfg.doRandomize = 1; cfg.randSeed = 7; cfg.policyTemperature = 1.0f;
  CHECK(kc_evaluator_create(ctx, model, &cfg, &ev));
  std::atomic<long> failures{0};
  std::atomic<double> checksum{0.0};
  auto run = [&](int nRows) {
    std::vector<std::thread> threads;
    for(int t = 0; t < clients; t++)
      threads.emplace_back([&, t] {
        const int C = single ? 1 : chunk;
        std::vector<int8_t> stones((size_t)C * HW);
        std::vector<kc_eval_position> ps(C);
        std::vector<kc_eval_output> os(C);
        std::vector<float> pol((size_t)C * P);
        double sum = 0;
        uint64_t r = mix((uint64_t)t * 1315423911ULL + 1);
        for(int done = 0; done < nRows; done += C) {
          for(int i = 0; i < C; i++) {
            const uint64_t id = repeat ? (uint64_t)((done + i) % repeat) + (uint64_t)t * 1000003ULL : ((uint64_t)t << 40) + (uint64_t)(done + i);
            r = mix(id);
            int8_t* s = stones.data() + (size_t)i * HW;
            const int nst = (int)(r % 16);
            memset(s, 0, HW);
            uint64_t q = r;
            for(int k = 0; k < nst; k++) { q = mix(q); s[q % HW] = (int8_t)(1 + ((q >> 20) & 1)); }
            ps[i] = kc_eval_position{s, nullptr, nst, (int8_t)(1 + ((r >> 33) & 1))};
            os[i] = kc_eval_output{};
            os[i].policyProbs = pol.data() + (size_t)i * P;
          }
          const int st = single ? kc_evaluator_evaluate(ev, &ps[0], KC_SYMMETRY_NOTSPECIFIED, 0, 0, &os[0])
                                : kc_evaluator_evaluate_many(ev, C, ps.data(), nullptr, 0, 0, os.data());
          if(st) { if(failures.fetch_add(1) == 0) fprintf(stderr, "evaluate: %s\n", kc_last_error()); return; }
          sum += os[0].whiteWinProb;
        }
        double c = checksum.load();
        while(!checksum.compare_exchange_weak(c, c + sum)) {}
      });
    for(auto& th : threads) th.join();
  };
  run(rows / 10 > 0 ? rows / 10 : 1);   // warm-up
  kc_evaluator_clear_stats(ev);
  const auto t0 = std::chrono::steady_clock::now();
  run(rows);
  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  kc_evaluator_stats st{};
  kc_evaluator_get_stats(ev, &st);
  const double requests = (double)clients * (double)(((rows + (single ? 1 : chunk) - 1) / (single ? 1 : chunk)) * (single ? 1 : chunk));
  printf("{\"metric\": \"evaluator_positions_per_s\", \"value\": %.1f, \"seconds\": %.4f, \"requests\": %.0f, \"clients\": %d, \"servers\": %d, \"maxBatch\": %d, "
         "\"mode\": \"%s\", \"chunk\": %d, \"rowsProcessed\": %llu, \"batches\": %llu, \"avgBatch\": %.1f, \"cacheHits\": %llu, \"backpressureWaits\": %llu, "
         "\"failures\": %ld, \"board\": \"%dx%d\", \"h2d_bytes_per_row\": 41, \"d2h_bytes_per_row\": %d}\n",
         requests / sec, sec, requests, clients, servers, batch, single ? "single" : "many", single ? 1 : chunk, (unsigned long long)st.rowsProcessed,
         (unsigned long long)st.batchesProcessed, (double)st.rowsProcessed / (double)(st.batchesProcessed ? st.batchesProcessed : 1),
         (unsigned long long)st.cacheHits, (unsigned long long)st.backpressureWaits, failures.load(), W, H, P * 4 + 16);
  kc_evaluator_destroy(ev);
  kc_model_destroy(model);
  kc_modelfile_free(mf);
  kc_ctx_destroy(ctx);
  return failures.load() ? 1 : 0;
}
