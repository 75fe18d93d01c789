// Internal interface between the games kernels, the C ABI glue and the two net paths.
#pragma once
#include "kc_internal.h"

namespace kc {

// ---- trunk tile geometry (bf16 path) ----------------------------------------------------------
// One MMA tile = 128 activation rows = NB boards laid side by side with one zero pad column each:
//   row = y * (NB*(W+1)) + b*(W+1) + x        (x == W is the pad column, rows >= H*NB*(W+1) are dead)
// so that a 3x3 tap (dy,dx) is the constant row shift (dy-1)*NB*(W+1) + (dx-1) and vertical
// out-of-board taps fall into the zero halo above/below the tile.
inline int boardsPerTile(int W, int H) { int nb = 128 / (H * (W + 1)); return nb > 4 ? 4 : nb; }
constexpr int TILE_ROWS = 128;
constexpr int HALO_ROWS = 28;                    // >= NB*(W+1)+1 for every supported size (5x5: 25, 6x6: 22); 28 keeps chunks 128-byte aligned
constexpr int ACT_ROWS = TILE_ROWS + 2 * HALO_ROWS;

// ---- handle accessors used by games.cu ----------------------------------------------------------
int handleCheckGeometry(kc_handle* h, int W, int H, int n);
bool handleIsBf16(const kc_handle* h);
bool handlePermutesDirs(const kc_handle* h);   // KC_FLAG_SYM_PERMUTE_DIRS
void* handleInputTiles(kc_handle* h);      // tensor path: [tiles][2][128] 16-byte chunks, 8 channels each in the handle's 16-bit operand format
// bit patterns of 1.0 and of `k` (the win length, a small integer) in the handle's operand format: what a tile producer writes.
// (tcgen05 kind::f16 wants A and B in ONE format: an MMA with bf16 tiles against fp16 weights is an illegal instruction on B200.)
void handleTileConstants(const kc_handle* h, float k, uint32_t* one, uint32_t* kBits);
float* handleRawInput(kc_handle* h);       // both paths: kc_forward's device copy of the raw rows [n][15*H*W] fp32
float* handleRawGlobal(kc_handle* h);      // [n]
int handleConvertRaw(kc_handle* h, int n, const int8_t* sym_dev, cudaStream_t stream);   // tensor path: raw NCHW rows -> symmetrised input tiles
float* handleInputNHWC(kc_handle* h);      // fp32 path: [n][H*W][15]
float* handleInputGlobal(kc_handle* h);    // fp32 path: [n][1]
// Runs the net on the handle's (already symmetrised) input buffer on `stream`; symmetry_dev (device
// pointer, may be null) is used for the inverse symmetry of the spatial outputs.
int handleRunOnStream(kc_handle* h, int n, cudaStream_t stream, const int8_t* symmetry_dev, const int* nDev = nullptr, int rowOffset = 0,
                      bool symIsLocal = false);   // nDev, rowOffset: bf16 path only; symIsLocal: symmetry_dev[0] belongs to row rowOffset
int handleCheckAbort(kc_handle* h);   // after a synchronise
int handleTilesPerItem(const kc_handle* h);   // tensor path: activation tiles one CTA work item holds (TrunkCfg::NT)
bool handleCanLeaveRegisters(const kc_handle* h);
void handleLeaveRegisters(kc_handle* h, bool on);   // bf16 pair-mode trunk: use the setmaxnreg variant (16 k registers per SM stay free)
// NNEvaluator::evaluate post-processing (nneval.cpp:702-815) of the handle's last outputs, on `stream`
void launchPostprocess(kc_handle* h, int n, int LW, const uint32_t* legal_dev, const uint32_t* status_dev, const uint64_t* sitHash_dev,
                       float policyTemperature, float* policy_dev, float* winLoss_dev, float* misc_dev, uint64_t* nnHash_dev, cudaStream_t stream,
                       int rowOffset = 0, int* nonfinite_dev = nullptr);   // rowOffset: first row of the handle's outputs to read;
                       // nonfinite_dev: set to 1 if a row's policy sum or win / loss probabilities are not finite (nneval.cpp:745-750, 789-793)

}  // namespace kc
