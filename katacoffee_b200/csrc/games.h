// Host-side object behind kc_games (shared by games.cu and search.cu).
#pragma once
#include <vector>

#include "games_device.cuh"
#include "kc_internal.h"

struct kc_games {
  kc_ctx* ctx = nullptr;
  kc::Geom geom;
  kc::State st;
  bool big = false;                   // a board beyond 7x7: 128-bit bitboards, games_big.cuh
  void* d_bigGeom = nullptr;          // kc::BigGeom (line masks) for those boards
  uint64_t* d_zob = nullptr;          // [HW][2 colours][2]
  int16_t* d_moves = nullptr;
  uint32_t* d_legal = nullptr; uint32_t* d_status = nullptr; uint64_t* d_sitHash = nullptr; int16_t* d_played = nullptr;   // 4 slots of G entries each (slot 0 = current position; 1..3 = the multi-ply launches' per-ply ring)
  unsigned long long* d_stats = nullptr;  // 8 counters
  float* d_planes = nullptr; float* d_global = nullptr;  // fp32 feature outputs
  int8_t* d_sym = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;
  float lastKernelMs = 0.f;
  std::vector<cudaEvent_t> evPool;   // one (start, stop) pair per ply of kc_games_run_timed
  void* d_flush = nullptr; size_t flushBytes = 0;
  // rules+features-only timing: consecutive plies write their planes to different ring slots (4 x G x 15*HW fp32 > L2),
  // so a ply never overwrites lines of the previous one that are still dirty in L2
  float* d_planesRing[3] = {nullptr, nullptr, nullptr};
  int lastRunPlies = 0; bool lastRunRing = false;
  const float* lastRunPlanes = nullptr;   // where the last ply of the last rules+features kc_games_run wrote its planes
  // kc_games_postprocess outputs, allocated on first use and kept: [G][4*HW] probabilities, [G][2] win/loss, [G][2] misc, [G][2] nnHash
  float *d_ppPolicy = nullptr, *d_ppWinLoss = nullptr, *d_ppMisc = nullptr; uint64_t* d_ppHash = nullptr;
};


namespace kc {
int gamesRefreshOutputs(kc_games* G);   // games.cu: launches on G->stream, does not synchronise
// kc_games_eval with an optional device-side row count (bf16 path): only the first *nDev lanes are evaluated
int gamesEval(kc_games* G, kc_handle* h, const int8_t* symmetry, const int* nDev, int rowOffset = 0, bool smallCtas = false, bool symOnDevice = false);
}
