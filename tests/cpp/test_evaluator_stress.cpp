// Stress / race test of the evaluator front end (include/katacoffee_b200.h "Evaluator front end", csrc/evaluator.cpp):
// many client threads against a few server threads with a tiny staging ring and a tiny cache, a batch function whose
// outputs are a pure function of the staged row.  Every result is recomputed by the client and compared bit for bit, so a
// row delivered to the wrong client, a buffer recycled too early or a torn cache entry fails the run.
//
// Built twice by tests/test_evaluator.py::test_cpp_stress: plain (against libkatacoffee_b200.so) and with
// -fsanitize=thread together with csrc/evaluator.cpp + csrc/zobrist.cpp compiled -DKC_EVALUATOR_HOST_ONLY.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "katacoffee_b200.h"

static const int W = 5, H = 5, HW = 25, P = 100;

static uint64_t mix(uint64_t x) {
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
// what the batch function writes for a row: depends on everything that is staged for it
static float policyOf(uint64_t b, uint64_t w, uint64_t m, int sym, int j) { return (float)(mix(b ^ mix(w) ^ mix(m ^ (uint64_t)sym * 77) ^ (uint64_t)j) >> 40) * (1.0f / 16777216.0f); }
static float ownerOf(uint64_t b, uint64_t w, int j) { return (float)((int)(mix(b * 3 + w + (uint64_t)j) >> 56) - 128) / 64.0f; }

static std::atomic<long> g_rows{0}, g_batches{0}, g_maxBatch{0};
static int backendFn(void*, int, const kc_eval_batch* b) {
  g_rows += b->n; g_batches += 1;
  long mb = g_maxBatch.load();
  while(b->n > mb && !g_maxBatch.compare_exchange_weak(mb, b->n)) {}
  for(int i = 0; i < b->n; i++) {
    for(int j = 0; j < P; j++) b->policyProbs[(size_t)i * P + j] = policyOf(b->black[i], b->white[i], b->misc[i], b->symmetry[i], j);
    b->whiteWinLoss[2 * i] = policyOf(b->black[i], b->white[i], b->misc[i], b->symmetry[i], 1000);
    b->whiteWinLoss[2 * i + 1] = policyOf(b->black[i], b->white[i], b->misc[i], b->symmetry[i], 1001);
    b->miscOut[2 * i] = policyOf(b->black[i], b->white[i], b->misc[i], b->symmetry[i], 1002);
    b->miscOut[2 * i + 1] = policyOf(b->black[i], b->white[i], b->misc[i], b->symmetry[i], 1003);
    if(b->wantOwnership)
      for(int j = 0; j < HW; j++) b->ownership[(size_t)i * HW + j] = ownerOf(b->black[i], b->white[i], j);
  }
  return 0;
}

struct Position {
  int8_t stones[HW];
  int16_t moves[10];
  int numTurns;
  int8_t nextPla;
  uint64_t black, white, misc;   // what the front end will stage for it
};

static Position makePosition(uint64_t id) {
  Position p{};
  uint64_t r = mix(id);
  int n = 0;
  for(int c = 0; c < HW; c++) {
    const int v = (int)((r >> (2 * (c % 30))) & 3);
    if(c == 29) r = mix(r);
    p.stones[c] = v == 3 ? 0 : (int8_t)v;
    n += p.stones[c] != 0;
  }
  p.nextPla = (int8_t)(1 + (mix(id ^ 5) & 1));
  p.numTurns = n;
  for(int k = 0; k < 5; k++) { p.moves[2 * k] = -1; p.moves[2 * k + 1] = 0; }
  const uint64_t h = mix(id ^ 99);
  p.moves[8] = (int16_t)(h % 100); p.moves[9] = (int16_t)(1 + ((h >> 8) & 1));
  p.black = p.white = 0;
  for(int y = 0; y < H; y++)
    for(int x = 0; x < W; x++) {
      if(p.stones[y * W + x] == 1) p.black |= 1ULL << (y * (W + 1) + x);
      if(p.stones[y * W + x] == 2) p.white |= 1ULL << (y * (W + 1) + x);
    }
  p.misc = (uint64_t)((p.moves[8] % HW) | (p.moves[9] << 6)) | ((uint64_t)(p.moves[8] / HW) << 40) | ((uint64_t)p.numTurns << 48) |
           ((uint64_t)(p.nextPla << 3) << 56);
  return p;
}

static std::atomic<long> g_errors{0};
static void fail(const char* what, uint64_t id) {
  if(g_errors.fetch_add(1) < 10) fprintf(stderr, "MISMATCH %s (position %llu)\n", what, (unsigned long long)id);
}

static void verify(const Position& p, uint64_t id, const kc_eval_output& o, const float* pol, const float* own, bool wantOwn) {
  const int sym = o.symmetry;
  if(sym < 0 || sym > 7) { fail("symmetry", id); return; }
  for(int j = 0; j < P; j++) if(pol[j] != policyOf(p.black, p.white, p.misc, sym, j)) { fail("policy", id); return; }
  if(o.whiteWinProb != policyOf(p.black, p.white, p.misc, sym, 1000) || o.whiteLossProb != policyOf(p.black, p.white, p.misc, sym, 1001) ||
     o.varTimeLeft != policyOf(p.black, p.white, p.misc, sym, 1002) || o.shorttermWinlossError != policyOf(p.black, p.white, p.misc, sym, 1003)) fail("values", id);
  if(wantOwn) {
    const float sign = p.nextPla == 2 ? 1.f : -1.f;
    for(int j = 0; j < HW; j++) if(own[j] != sign * std::tanh(ownerOf(p.black, p.white, j))) { fail("owner map", id); return; }
  }
}

int main(int argc, char** argv) {
  const int clients = argc > 1 ? atoi(argv[1]) : 12;
  const int perClient = argc > 2 ? atoi(argv[2]) : 20000;
  const int distinct = argc > 3 ? atoi(argv[3]) : 3000;
  kc_evaluator_config cfg{};
  cfg.nnXLen = W; cfg.nnYLen = H; cfg.winLen = 4;
  cfg.maxBatch = 16; cfg.maxConcurrentEvals = 16; cfg.numServerThreads = 3;
  cfg.cacheSizePowerOfTwo = 9; cfg.mutexPoolSizePowerOfTwo = 3;   // 512 entries for `distinct` positions: constant eviction
  cfg.doRandomize = 1; cfg.defaultSymmetry = 0; cfg.randSeed = 12345; cfg.policyTemperature = 1.0f;
  kc_evaluator* ev = nullptr;
  if(kc_evaluator_create_custom(&cfg, backendFn, nullptr, &ev)) { fprintf(stderr, "create: %s\n", kc_last_error()); return 2; }
  std::vector<Position> positions;
  for(int i = 0; i < distinct; i++) positions.push_back(makePosition((uint64_t)i));
  std::vector<std::thread> threads;
  std::atomic<long> hits{0}, total{0};
  for(int t = 0; t < clients; t++)
    threads.emplace_back([&, t] {
      std::vector<float> pol(P), own(HW);
      uint64_t r = mix((uint64_t)t + 1000);
      if(t % 3 == 2) {
        // a batched client: evaluate_many over chunks of 40 rows (more than two staging buffers)
        const int chunk = 40;
        std::vector<kc_eval_position> ps(chunk);
        std::vector<kc_eval_output> os(chunk);
        std::vector<float> pols((size_t)chunk * P), owns((size_t)chunk * HW);
        std::vector<uint64_t> ids(chunk);
        for(int done = 0; done < perClient; done += chunk) {
          for(int i = 0; i < chunk; i++) {
            r = mix(r);
            ids[i] = r % (uint64_t)distinct;
            const Position& p = positions[ids[i]];
            ps[i] = kc_eval_position{p.stones, p.moves, p.numTurns, p.nextPla};
            os[i] = kc_eval_output{};
            os[i].policyProbs = pols.data() + (size_t)i * P; os[i].whiteOwnerMap = owns.data() + (size_t)i * HW;
          }
          if(kc_evaluator_evaluate_many(ev, chunk, ps.data(), nullptr, 0, 1, os.data())) { fail(kc_last_error(), 0); return; }
          for(int i = 0; i < chunk; i++) {
            verify(positions[ids[i]], ids[i], os[i], os[i].policyProbs, os[i].whiteOwnerMap, true);
            hits += os[i].cacheHit; total += 1;
          }
        }
        return;
      }
      for(int i = 0; i < perClient; i++) {
        r = mix(r);
        const uint64_t id = r % (uint64_t)distinct;
        const Position& p = positions[id];
        kc_eval_position ep{p.stones, p.moves, p.numTurns, p.nextPla};
        kc_eval_output o{};
        const bool wantOwn = (r >> 40) % 4 == 0;
        o.policyProbs = pol.data(); o.whiteOwnerMap = wantOwn ? own.data() : nullptr;
        if(kc_evaluator_evaluate(ev, &ep, KC_SYMMETRY_NOTSPECIFIED, (r >> 50) % 16 == 0, wantOwn, &o)) { fail(kc_last_error(), id); return; }
        verify(p, id, o, pol.data(), own.data(), wantOwn);
        hits += o.cacheHit; total += 1;
      }
    });
  for(auto& t : threads) t.join();
  kc_evaluator_stats st{};
  kc_evaluator_get_stats(ev, &st);
  kc_evaluator_destroy(ev);
  const bool accounting = (long)st.rowsProcessed == g_rows.load() && (long)st.batchesProcessed == g_batches.load() && (long)st.cacheHits == hits.load() &&
                          g_maxBatch.load() <= cfg.maxBatch;
  printf("{\"clients\": %d, \"requests\": %ld, \"rows\": %llu, \"batches\": %llu, \"avgBatch\": %.2f, \"cacheHits\": %llu, \"upgrades\": %llu, "
         "\"backpressureWaits\": %llu, \"mismatches\": %ld, \"accounting\": %s}\n",
         clients, total.load(), (unsigned long long)st.rowsProcessed, (unsigned long long)st.batchesProcessed,
         (double)st.rowsProcessed / (double)(st.batchesProcessed ? st.batchesProcessed : 1), (unsigned long long)st.cacheHits,
         (unsigned long long)st.ownerMapUpgrades, (unsigned long long)st.backpressureWaits, g_errors.load(), accounting ? "true" : "false");
  return (g_errors.load() == 0 && accounting) ? 0 : 1;
}
