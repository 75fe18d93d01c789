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
// The same map for device code, and the channel permutation of play mode (ledger K): V1 input channels 3..6 are the last
// move's direction, so under symmetry s channel 3+d of the source is channel 3+symDir(d, s) of the symmetric position
// (symDir is an involution and equal for s and its inverse, so the same function serves inputs and policy outputs).
__host__ __device__ inline int symDirInline(int dir, int symmetry) {
  const bool tr = (symmetry & 4) != 0, fx = (symmetry & 2) != 0, fy = (symmetry & 1) != 0;
  if(fx != fy) dir = dir == 2 ? 3 : (dir == 3 ? 2 : dir);
  if(tr) dir = dir == 0 ? 1 : (dir == 1 ? 0 : dir);
  return dir;
}
__host__ __device__ inline int playModeChannel(int c, int symmetry) { return (c >= 3 && c <= 6) ? 3 + symDirInline(c - 3, symmetry) : c; }

}  // namespace kc

struct kc_ctx {
  int device = 0;
  int smCount = 0;
  size_t smemOptin = 0;
};
