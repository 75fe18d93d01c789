// Internal declarations shared by the translation units of libkatacoffee_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/katacoffee_b200.h"

namespace kc {

void setError(const std::string& msg);
int fail(const std::string& msg);  // sets the thread-local error, returns 1

#define KC_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t kc_err__ = (expr);                                                            \
    if(kc_err__ != cudaSuccess)                                                               \
      return kc::fail(std::string(#expr) + ": " + cudaGetErrorString(kc_err__) + " (" +       \
                      __FILE__ + ":" + std::to_string(__LINE__) + ")");                       \
  } while(0)

#define KC_CHECK(cond, msg)                                                                   \
  do {                                                                                        \
    if(!(cond)) return kc::fail(std::string(msg));                                            \
  } while(0)

// Host-side Zobrist generation (zobrist.cpp)
struct ZobristTables {
  uint64_t board[KC_MAX_ARR_SIZE][4][2];
  uint64_t player[4][2];
  uint64_t sizeX[KC_MAX_LEN + 1][2];
  uint64_t sizeY[KC_MAX_LEN + 1][2];
};
const ZobristTables& zobrist();
constexpr uint64_t ZOBRIST_GAME_IS_OVER0 = 0xb6f9e465597a77eeULL;  // cpp/game/board.cpp:26-27
constexpr uint64_t ZOBRIST_GAME_IS_OVER1 = 0xf1d583d960a4ce7fULL;

// Spatial symmetry map of SymmetryHelpers::copyInputsWithSymmetry / copyOutputsWithSymmetry
// (cpp/neuralnet/nninputs.cpp:252-357): dstPos[srcPos] for an H x W plane.
void symmetryDstOfSrc(int H, int W, int symmetry, bool reverse, int* dstOfSrc);
int symDir(int dir, int symmetry);  // nninputs.cpp:409-433 (ledger J)

}  // namespace kc

struct kc_ctx {
  int device = 0;
  int smCount = 0;
  size_t smemOptin = 0;
};
