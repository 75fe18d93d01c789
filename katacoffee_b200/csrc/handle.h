// Compute handle shared by the fp32 check path (net_fp32.cu) and the bf16 tcgen05 path (net_bf16.cu).
#pragma once
#include "model.h"
#include "net.h"

namespace kc {
struct Fp32Buffers {
  float *in = nullptr, *global = nullptr, *mask = nullptr, *maskSum = nullptr;
  float *trunk = nullptr, *tip = nullptr, *a = nullptr, *b = nullptr, *c = nullptr, *pool = nullptr, *bias = nullptr;
};
}  // namespace kc

namespace kc { struct RowStaging; }

struct kc_handle {
  kc_ctx* ctx = nullptr;
  const kc_model* model = nullptr;
  int maxBatch = 0, W = 0, H = 0;
  bool leaveRegisters = false;   // launch the trunk variant that leaves registers to co-resident kernels (set by a pipelined search)
  unsigned flags = 0;
  bool bf16 = true;
  cudaStream_t stream = nullptr;
  cudaStream_t h2dStream = nullptr, d2hStream = nullptr;   // kc_forward pipelines copy / compute / copy over row chunks
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::vector<cudaEvent_t> chunkEvents;
  // staging of kc_forward's host rows
  float* d_raw = nullptr; float* d_rawGlobal = nullptr; int8_t* d_sym = nullptr;
  uint8_t* d_dstOfSrc = nullptr;     // [8][HW] copyInputsWithSymmetry map
  uint8_t* d_dstOfSrcRev = nullptr;  // [8][HW] copyOutputsWithSymmetry map
  // outputs (device): policy [n][4*HW], value [n][2], misc [n][2], ownership [n][HW]
  float *d_policy = nullptr, *d_value = nullptr, *d_misc = nullptr, *d_own = nullptr;
  kc::Fp32Buffers f32;
  // bf16 path
  void* d_tiles = nullptr;           // [numTiles][2][128] x 16 B input tiles
  int numTilesAlloc = 0;
  long long* d_dbg = nullptr;        // KC_TRUNK_PROBE=1: timeline of one layer boundary (kc_handle_trunk_probe)
  int* d_abort = nullptr;            // set by the trunk kernel if an mbarrier wait timed out
  int64_t launches = 0;
  float lastTrunkMs = 0.f;
  // CUDA-event pairs around every trunk launch since the last kc_handle_trunk_time() call
  std::vector<cudaEvent_t> evPool;
  int evUsed = 0;
  int lastN = 0;
  kc::RowStaging* rows = nullptr;   // kc_forward_rows: page-locked staging + gather / scatter worker threads, created on first use
};

namespace kc {
// net_bf16.cu
int allocTrunkBuffers(kc_handle* h);
void freeTrunkBuffers(kc_handle* h);
// rowOffset (a multiple of 2*NB rows) selects a chunk of the batch: inputs, tiles and outputs are all offset by it
int convertInputToTiles(kc_handle* h, int n, int rawNHWC, const int8_t* sym_dev, cudaStream_t st, int rowOffset = 0);
// nDev (device pointer, may be null): the live row count of a batch that was compacted on the device; n is then its upper bound
int runTrunkBf16(kc_handle* h, int n, cudaStream_t st, const int8_t* sym_dev, int rowOffset = 0, const int* nDev = nullptr, bool symIsLocal = false);
int checkTrunkAbort(kc_handle* h);   // after a synchronise: non-zero (and error set) if the kernel bailed out
}  // namespace kc
