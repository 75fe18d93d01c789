// The tensor-core production path: the whole residual net as ONE persistent tcgen05/TMEM kernel.
//
// What it computes is the reference's Model::apply (cpp/neuralnet/eigenbackend.cpp:1420-1472:
// trunk :1202-1226, blocks :912-930 / :968-1005, policy head :1265-1298, value head :1341-1376), for
// boards that fill nnXLen x nnYLen or -- KC_FLAG_MASKED_BOARDS -- smaller ones in that slot, with fp16
// (default) or bf16 operands and fp32 accumulation.  The bullets describe the basic two-tile form
// (TrunkCfg<128, 2, 7>); the CTA-pair, 192- and 256-channel forms, the deferred value head and what
// the epilogues were measured to cost are in DESIGN.md 4.2.  How it computes it is B200-native:
//
//  * A CTA owns two 128-row activation tiles (NB boards each, laid side by side with one zero pad
//    column per board, see net.h) for the whole depth of the net.  Activations never leave the SM:
//    they live in shared memory as bf16 in the canonical K-major no-swizzle UMMA layout
//    [channel chunk of 8][row][8], where a 3x3 tap is a +/- row offset on the A descriptor
//    (implicit GEMM without im2col; the zero halo rows implement the zero padding).
//  * The fp32 residual stream lives in TMEM for the whole trunk: region T (128 columns per tile) is
//    the trunk, region S the block-internal tensor.  conv2 of a block is issued with the accumulate
//    flag set on its first MMA, so `trunk += conv2(...)` (eigenbackend.cpp:928-929) costs nothing.
//  * Weights are pre-tiled on the host into the exact shared-memory image of each (16-channel,
//    3-tap) stage and streamed L2 -> smem by the TMA engine (cp.async.bulk, UBLKCP) through a
//    7-deep mbarrier ring, shared by both tiles (each stage feeds two MMAs).
//  * Warp roles: warp 0 = TMA producer, warps 1 / 2 = MMA issuer of tile 0 / 1 (one thread each;
//    warp 1 also owns the TMEM allocation), warps 4-7 / 8-11 = epilogue of tile 0 / 1
//    (thread = TMEM lane = activation row).
//  * The epilogue (folded BN + ReLU + pad-row masking + bf16 pack) publishes the next layer's input
//    16 channels at a time; the MMA warp starts the next layer on chunk c as soon as both tiles have
//    published chunk c, so the tensor pipe only idles for the first chunk of every layer.
//  * Global pooling, the pooled matmuls, both heads, the inverse output symmetry and the final
//    stores are done by the epilogue warps in fp32 on CUDA cores (they are < 0.5 % of the FLOPs).
//  * The kernel is far larger than the 32 KB instruction cache, so the per-layer path is kept lean:
//    once-per-item code is out of line and shared (poolReduce, pooledMatmul4, mishf), each part of
//    the head epilogue has one call site, and the clock probes of kc_handle_trunk_probe are compiled
//    into separate instantiations (TrunkCfg::PROBE) that only KC_TRUNK_PROBE=1 launches.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "handle.h"
#include "umma.cuh"

namespace kc {

using namespace ptx;

constexpr int CHUNK_BYTES = ACT_ROWS * 16;             // one 8-channel chunk of an activation tile
constexpr int SCR_STRIDE = 17;
constexpr int MAX_NB = 4;
constexpr int HEADC = 32;   // p1 = g1 = v1 channels of the BASELINE nets; TrunkCfg::HC is what a kernel instantiation runs
constexpr int MAX_HEADC = 64;
constexpr int MAX_V2 = 128;

// Compile-time shape of one trunk-kernel instantiation.  <128, 2, 7>: trunks up to 128 channels, two activation tiles
// per CTA (TMEM: 2 x (T 128 + S 128) columns).  <192, 1, 6>: trunks up to 192 channels (b15c192), one tile per CTA
// (TMEM: T 192 + S 192 columns of the 512 allocated).
// PAIR: the kernel runs as clusters of two CTAs on one TPC and every MMA is a tcgen05 cta_group::2 instruction (M = 256:
// 128 activation rows from each CTA, the weight block split in halves between the two shared memories), so each SM reads
// and stages only half of the weights: shared-memory operand traffic drops from 128 to 96 B/clk and the L2 -> SM weight
// stream halves.
// KSTEPS: K-steps (16 input channels x 1 tap each) per weight stage.  3 = one kernel row of a 16-channel chunk; 9 = the whole chunk
// (pair mode): one barrier round and one commit per 9 MMAs, because the single issuing thread -- not the tensor pipe -- is what
// runs out first with small stages (an MMA of a CTA pair is 64 clk of pipe time; three waits + a commit per 192 clk did not fit).
template <int MAXC_, int NT_, int NSTAGES_, bool PAIR_ = false, bool REALLOC_ = false, int KSTEPS_ = 3, int HEADC_ = 32, bool PROBE_ = false>
struct TrunkCfg {
  static constexpr bool PROBE = PROBE_;   // the in-kernel clock probes (kc_handle_trunk_probe) are compiled into their own instantiations: they
                                          // are ~3 KB of a per-layer path that has to share a 32 KB instruction cache
  static constexpr int MAXC = MAXC_, NT = NT_, NSTAGES = NSTAGES_, KSTEPS = KSTEPS_;
  static constexpr int HC = HEADC_;   // p1 = g1 = v1 channels: 32 (b6c96 .. b15c192), 48 (b20c256), 64 (b40c256) -- python/modelconfigs.py
  static_assert(HC % 16 == 0 && HC >= 32 && HC <= MAX_HEADC && 3 * HC <= MAXC, "the head convolution's 3 x HC columns live in TMEM region S");
  static_assert(KSTEPS == 3 || KSTEPS == 9, "a stage is one kernel row or one whole 16-channel chunk of a 3x3 layer");
  static constexpr bool PAIR = PAIR_;
  static constexpr bool REALLOC = REALLOC_;   // launch with 128 registers per thread and re-allocate between the warpgroups (setmaxnreg)
  static constexpr int NCTA = PAIR ? 2 : 1;
  static constexpr int NCH = MAXC / 16;                       // 16-channel chunks the epilogue publishes
  static constexpr int MAXG = MAXC / 4 < 32 ? 32 : MAXC / 3;  // gpool channels: 32 (c128), 64 (c192)
  static constexpr int POOLW = 3 * MAXG;                      // pooled vector per board (>= 96 for the heads)
  static constexpr int THREADS = 128 + NT * 128;              // warp 0 TMA producer, 1..NT MMA issuers, 4.. epilogue (4 warps per tile)
  static constexpr int ACT_BYTES = (MAXC / 8) * CHUNK_BYTES;
  static constexpr int STAGE_BYTES = KSTEPS * MAXC * 32 / NCTA;   // per CTA
  static constexpr int OFF_ACT = 0;
  static constexpr int OFF_RING = OFF_ACT + NT * ACT_BYTES;
  static constexpr int OFF_SCR = OFF_RING + NSTAGES * STAGE_BYTES;
  static constexpr int OFF_POOLA = OFF_SCR + NT * 128 * SCR_STRIDE * 4;
  static constexpr int OFF_POOLB = OFF_POOLA + NT * MAX_NB * POOLW * 4;
  // the per-board bias buffer [MAX_NB][MAXC] and the v2 buffer [MAX_NB][MAX_V2] of the gpool / head epilogues live in the pooling
  // scratch: they are written after the last pooling pass has been read (named barrier in between) and read before the next one
  static_assert((MAX_NB * MAXC + MAX_NB * MAX_V2) * 4 <= 128 * SCR_STRIDE * 4, "bias + v2 buffers alias the pooling scratch");
  static constexpr int HEAD_ADD_OFF = MAX_NB * MAXC + MAX_NB * MAX_V2;   // floats into a tile's pooling scratch: the 32-channel head's folded policy bias [MAX_NB][32]
  static_assert(HEAD_ADD_OFF + MAX_NB * 32 <= 128 * SCR_STRIDE && 4 * MAX_NB * 32 <= MAX_NB * MAXC, "folded policy bias and its K-quarter parts");
  static constexpr int OFF_SYM = OFF_POOLB + NT * MAX_NB * POOLW * 4;
  static constexpr int OFF_MASK = OFF_SYM + 800;                  // (OFF_SYM: the inverse symmetry maps, 8 x H*W bytes, up to 10x10)                  // masked boards: [NT][128] row mask + [NT][MAX_NB][4] per-board pooling constants (fp32)
  static constexpr int OFF_PAR = OFF_MASK + NT * (128 + MAX_NB * 4) * 4;                   // [tile][2 buffers][scale MAXC | bias MAXC] fp32: the next layer's
  static constexpr int OFF_BAR = OFF_PAR + NT * 2 * 2 * MAXC * 4; // folded BN, staged while the tensor core is still busy with it
  // barriers (8 bytes each)
  static constexpr int BAR_FULL = 0, BAR_EMPTY = BAR_FULL + NSTAGES, BAR_ACC = BAR_EMPTY + NSTAGES, BAR_ACTFREE = BAR_ACC + NT,
                       BAR_IN = BAR_ACTFREE + NT, BAR_HEAD = BAR_IN + NT, BAR_CHUNK = BAR_HEAD + NT, BAR_FULLP = BAR_CHUNK + NT * NCH,
                       BAR_INP = BAR_FULLP + NSTAGES, BAR_SKEW = BAR_INP + NT, NUM_BARS = BAR_SKEW + 1;   // FULLP / INP: the peer CTA's weights / input tile landed
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int OFF_PROG = OFF_TMEM + 8;     // stages issued so far by tile 0's MMA issuer (read by tile 1's: the enforced lag)
  static constexpr int SMEM = OFF_TMEM + 16;
  static_assert(NT * 2 * MAXC <= 512, "TMEM: every tile needs a trunk region and a block-internal region");
  static_assert(SMEM <= 232448, "shared memory budget");
  static_assert(POOLW >= 3 * HC, "the heads pool HC channels three ways");
};
using Cfg128 = TrunkCfg<128, 2, 7>;
using Cfg128P = TrunkCfg<128, 2, 5, true, false, 9>;
using Cfg128PProbe = TrunkCfg<128, 2, 5, true, false, 9, 32, true>;
using Cfg128PR = TrunkCfg<128, 2, 5, true, true, 9>;   // the variant that leaves 16 k registers per SM to co-resident kernels (search half batches)
using Cfg192 = TrunkCfg<192, 1, 6>;
using Cfg256 = TrunkCfg<256, 1, 4>;
using Cfg256H48 = TrunkCfg<256, 1, 4, false, false, 3, 48>;   // b20c256's 48-channel heads (modelconfigs.py:284-286)
using Cfg256H64 = TrunkCfg<256, 1, 4, false, false, 3, 64>;   // b40c256's 64-channel heads                    // trunks up to 256 channels (b20c256 / b40c256 shapes): one tile, T 256 + S 256 = all 512 TMEM columns,
                                                        // single CTA (a pair's 9-tap stages of 256 output channels do not fit beside the 94 KB activation tile)
using Cfg192PProbe = TrunkCfg<192, 1, 4, true, false, 9, 32, true>;
using Cfg192P = TrunkCfg<192, 1, 4, true, false, 9>;   // b15c192 as CTA pairs: each CTA stages half of the 192 output channels (27 KB stages)

enum { EPI_BN = 0, EPI_GPOOL = 1, EPI_HEAD = 2 };

struct LayerDesc {
  int nk;          // K-steps (16 input channels x 1 tap each)
  int ntaps;       // 9 (3x3) or 1 (1x1)
  int N;           // MMA N = output channels padded to 16
  int outSel;      // 0 = TMEM region T (trunk), 1 = region S
  int accumulate;  // first MMA accumulates onto the region (residual add)
  int epi;         // epilogue kind
  int epiC;        // channels published to the activation tile (multiple of 16)
  int gpoolC;      // EPI_GPOOL: gpool channels (columns epiC .. epiC+gpoolC)
  int act;         // activation after the folded BN this layer's epilogue applies (0 identity, 1 ReLU, 2 Mish)
  int gpoolAct;    // EPI_GPOOL: activation of the gpool branch
  unsigned wOffset; // byte offset of this layer's first stage in the weight stream
  int pOff;        // float offset of this layer's parameters
};
constexpr int KC_TRUNK_PROBE_WORDS = 64 + 2 * 48 * 8;   // kc_handle_trunk_probe: 64 point probes + the whole-item timeline [tile][layer][8]
constexpr int MAX_LAYERS = 48;   // 1 + 2*blocks + 1; the table travels in the kernel parameter block (constant bank:
                                 // warp-uniform loads, so the MMA issue loop stays on the uniform datapath)

struct TrunkProgram {
  std::vector<LayerDesc> layers;
  // weight streams per operand format (index = the tcgen05 format code: 0 fp16, 1 bf16)
  uint8_t* d_w[2] = {nullptr, nullptr};
  uint8_t* d_wPair[2] = {nullptr, nullptr};   // the same stream with every stage split in halves of the output channels (pair mode)
  float* d_params = nullptr;
  LayerDesc* d_layers = nullptr;
  size_t wBytes = 0;
  int v2C = 0;
  int cfg = 0;     // 0: TrunkCfg<128, 2, 7>, 1: TrunkCfg<192, 1, 6>, 2: TrunkCfg<256, 1, 4>
  int headC = 32;  // p1 = g1 = v1 channels (cfg 2 also runs 48 and 64)
  double flopsPerEval = 0;
};

struct TrunkParams {
  const uint4* tiles; const uint8_t* wstream; const float* params;
  LayerDesc layers[MAX_LAYERS];
  int numLayers, numItems, n;
  const int* nDev;   // if non-null: the number of rows is read from device memory (a batch compacted on the device)
  int NB, W, H, HW, stride, tileRowW;
  int wMagic;        // 65536 / W + 1: cell / W = cell * wMagic >> 16 for cell < 128, W <= 10
  const int8_t* sym; const uint8_t* dstOfSrcRev;
  float *policy, *value, *misc, *own;
  int permuteDirs;   // KC_FLAG_SYM_PERMUTE_DIRS
  int opFmt;         // tensor-core operand format of the input tiles, the weights and the activations: 0 fp16 (default), 1 bf16
                     // (KC_FLAG_OPERANDS_BF16).  One format for A and B: a kind::f16 MMA with mixed formats is an illegal instruction.
  int skew;          // two tiles per CTA: tile 1's MMA issuer starts EVERY layer only once tile 0's has issued this many weight stages of
                     // it, so that the layer boundary of one tile (accumulator -> epilogue -> first chunk published, ~1,200 clk of no
                     // MMAs) is covered by the other tile's MMAs.  Enforced per layer: left alone the two issuers fall into lock step
                     // (they share the weight ring, and the one ahead is the one that waits for slots)
  int g1Act, p1Act, v1Act, v2Act;   // head activations
  int* abortFlag;
  long long* dbg;    // diagnostic: SM clock at the hand-over points of one layer boundary (CTA 0, first item, tile 0), or null
  float poolScale1, poolScale2, invHW;
  int v2C;
  int deferHead;     // 32-channel heads: run an item's value head after the next item's first-layer epilogue (KC_TRUNK_DEFER_HEAD=0: off)
  int masked;        // KC_FLAG_MASKED_BOARDS: boards may be smaller than nnXLen x nnYLen; input channel 0 (the on-board plane) is the mask
                     // (eigenbackend.cpp:1438) and the pooling divides by each board's own cell count (:141-166)
};

// ------------------------------------------------------------------------------------------------
// device code
// ------------------------------------------------------------------------------------------------
// Activations (cpp/neuralnet/activations.h:4-6): 0 identity, 1 ReLU, 2 Mish = x*tanh(softplus(x)) (eigenbackend.cpp:729, linear
// above 20).  With n = e^x, tanh(log(1+n)) = n(n+2) / (n(n+2) + 2): one exponential and one division.  `act` is warp-uniform.
// Loads through shared-window addresses: a pointer handed to an out-of-line function is generic, and a generic load is slower than LDS
__device__ __forceinline__ float lds32(uint32_t addr) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr)); return v; }
__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// The kernel's code is far larger than the 32 KB instruction cache and its once-per-item epilogues (global pooling, heads) run cold --
// a cold pass over them was measured at 2x a warm one -- so whatever is not on the per-layer path is kept out of line and shared.
__device__ __noinline__ float mishf(float x) {
  const float n = __expf(fminf(x, 20.f));
  const float t = n * (n + 2.f);
  return x > 20.f ? x : x * __fdividef(t, t + 2.f);
}
__device__ __forceinline__ float actf(float x, int act) {
  if(act == 1) return fmaxf(x, 0.f);
  if(act == 2) return mishf(x);
  return x;
}

__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// two activations as tensor-core operands: fp16 (saturating: an activation beyond 65504 stays finite) or bf16; `f16` is warp-uniform
__device__ __forceinline__ uint32_t packOp(float a, float b, bool f16) {
  if(f16) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // first source -> upper half
    return r;
  }
  return pack2(a, b);
}

struct EpiCtx {
  int t;            // tile within the CTA
  int r;            // row = TMEM lane
  int e;            // thread index within the tile's epilogue group (== r)
  bool valid;       // row is a board cell
  int b, cell;      // board within tile, dense cell index
  int slotB, slotCell;   // the same for every cell of the board's nnXLen x nnYLen slot (slotCell -1: a pad row), whatever the mask says
  uint32_t tmemLane;  // tmem base + lane offset + tile column offset
  uint8_t* act;     // this tile's activation buffer
  uint32_t barChunk;  // address of actReady[t][0] (pair mode: shared::cluster address in the leader CTA)
  bool remote;        // pair mode, peer CTA: barriers the MMA issuer waits on live in the leader
  bool f16;           // activations are published as fp16 (else bf16)
  float* scr; float* poolA; float* poolB; float* biasBuf; float* v2buf;
  const float* par;  // staged folded BN of the layer being finished: scale[MAXC] | bias[MAXC]
  const float* maskRow;   // masked boards: [128] 1.0 / 0.0 per row of this tile (shared), else null
  const uint8_t* cellRow; // [H*W] tile row (y * tileRowW + x) of a board's cell, without the board offset (shared)
  const float* boardK;    // masked boards: [MAX_NB][4] = 1 / cells, (sqrt(cells) - 14) * 0.1, (sqrt(cells) - 14)^2 * 0.01 - 0.1 per board of this tile
  long long* dbg;    // non-null for the one thread / layer whose timeline is recorded
};

// folded BN + ReLU + mask for 16 columns -> two 16-byte chunks of the activation tile
template <bool RELU>
__device__ __forceinline__ void publish16T(const EpiCtx& c, int cc, const float v[16], const float* scale, const float* bias, const float* add, int act) {
  uint32_t pk[8];
#pragma unroll
  for(int q = 0; q < 4; q++) {
    float4 s = *(reinterpret_cast<const float4*>(scale + cc * 16) + q);     // shared memory (c.par), staged before the accumulator wait
    float4 bb = *(reinterpret_cast<const float4*>(bias + cc * 16) + q);
    float x0 = v[4 * q], x1 = v[4 * q + 1], x2 = v[4 * q + 2], x3 = v[4 * q + 3];
    if(add) { const float4 ad = *(reinterpret_cast<const float4*>(add + cc * 16) + q); x0 += ad.x; x1 += ad.y; x2 += ad.z; x3 += ad.w; }
    float a0, a1, a2, a3;
    if(RELU) {
      a0 = fmaxf(fmaf(x0, s.x, bb.x), 0.f); a1 = fmaxf(fmaf(x1, s.y, bb.y), 0.f);
      a2 = fmaxf(fmaf(x2, s.z, bb.z), 0.f); a3 = fmaxf(fmaf(x3, s.w, bb.w), 0.f);
    } else {
      a0 = actf(fmaf(x0, s.x, bb.x), act); a1 = actf(fmaf(x1, s.y, bb.y), act);
      a2 = actf(fmaf(x2, s.z, bb.z), act); a3 = actf(fmaf(x3, s.w, bb.w), act);
    }
    if(!c.valid) { a0 = a1 = a2 = a3 = 0.f; }
    pk[2 * q] = packOp(a0, a1, c.f16);
    pk[2 * q + 1] = packOp(a2, a3, c.f16);
  }
  uint8_t* dst = c.act + (size_t)(2 * cc) * CHUNK_BYTES + (size_t)(HALO_ROWS + c.r) * 16;
  *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(dst + CHUNK_BYTES) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  fence_proxy_async();
  __syncwarp();
  if((c.r & 31) == 0) { if(c.remote) mbar_arrive_cluster(c.barChunk + cc * 8); else mbar_arrive(c.barChunk + cc * 8); }
}

// per-board pooling of 16 channels held one row per thread: writes sum and max per (board, channel)
// into outSum/outMax[b*16 + j] (shared), using the tile's scratch.  All 128 threads of the tile call it.
// the ReLU instantiation is the hot one: keeping the other activations out of it keeps its code as it was
__device__ __forceinline__ void publish16(const EpiCtx& c, int cc, const float v[16], const float* scale, const float* bias, const float* add, int act) {
  if(act == 1) publish16T<true>(c, cc, v, scale, bias, add, act);
  else publish16T<false>(c, cc, v, scale, bias, add, act);
}

// Global pooling of 16 channels held one row per thread.  Two threads per (board, channel): even lane = even cells, odd lane = odd
// cells, combined by one shuffle; the even lane (e even, e >> 1 = board*16 + channel < NB*16) returns the board's sum (.x) and max (.y).
// The reduction is one out-of-line copy for every caller (two gpool layers, four passes in the heads): it stays in the instruction cache.
// The tile row of a cell is computed (cell / W = cell * wMagic >> 16), not looked up: shared-memory loads are slow while the tensor
// core streams its operands from the same memory.
__device__ __noinline__ float2 poolReduce(uint32_t scr, uint32_t maskRow, int e, int NB, int stride, int HW, int wMagic, int dRow) {
  const int o = e >> 1, part = e & 1;
  const int b = o >> 4, j = o & 15;
  float s = 0.f, m = -1.0f;   // eigenbackend.cpp:145: max starts at -1
  if(b < NB) {
    const uint32_t base = scr + (uint32_t)(b * stride * SCR_STRIDE + j) * 4u;
    auto crf = [&](int i) { return i + (int)(((unsigned)i * (unsigned)wMagic) >> 16) * dRow; };
    auto val = [&](int r) { return lds32(base + (uint32_t)(r * SCR_STRIDE) * 4u); };
    float sB = 0.f, mB = -1.0f;
    int i = part;
    if(maskRow) {   // off-board cells count as -1 in the maximum (eigenbackend.cpp:150-155); they are 0 in the sum already
      const uint32_t mk = maskRow + (uint32_t)(b * stride) * 4u;   // same summation order as below: a board that fills its slot gives the same bits
      for(; i + 6 < HW; i += 8) {
        const int r0 = crf(i), r1 = crf(i + 2), r2 = crf(i + 4), r3 = crf(i + 6);
        const float v0 = val(r0), v1 = val(r1), v2 = val(r2), v3 = val(r3);
        const float k0 = lds32(mk + r0 * 4u) - 1.0f, k1 = lds32(mk + r1 * 4u) - 1.0f, k2 = lds32(mk + r2 * 4u) - 1.0f, k3 = lds32(mk + r3 * 4u) - 1.0f;
        s += v0; sB += v1; s += v2; sB += v3;
        m = fmaxf(m, fmaxf(v0 + k0, v2 + k2)); mB = fmaxf(mB, fmaxf(v1 + k1, v3 + k3));
      }
      for(; i < HW; i += 2) {
        const int r = crf(i);
        const float v = val(r);
        s += v;
        m = fmaxf(m, v + (lds32(mk + r * 4u) - 1.0f));
      }
    } else {
      for(; i + 6 < HW; i += 8) {
        const int r0 = crf(i), r1 = crf(i + 2), r2 = crf(i + 4), r3 = crf(i + 6);
        const float v0 = val(r0), v1 = val(r1), v2 = val(r2), v3 = val(r3);
        s += v0; sB += v1; s += v2; sB += v3;
        m = fmaxf(m, fmaxf(v0, v2)); mB = fmaxf(mB, fmaxf(v1, v3));
      }
      for(; i < HW; i += 2) {
        const float v = val(crf(i));
        s += v;
        m = fmaxf(m, v);
      }
    }
    s += sB; m = fmaxf(m, mB);
  }
  const float s2 = __shfl_xor_sync(0xffffffffu, s, 1), m2 = __shfl_xor_sync(0xffffffffu, m, 1);
  return make_float2(s + s2, fmaxf(m, m2));   // even lane: even cells + odd cells
}
__device__ __forceinline__ void poolBoards16(const TrunkParams& P, const EpiCtx& c, const float g[16], float& sum, float& mx, long long* pd = nullptr) {
  if(pd) pd[0] = clock64();
  named_bar_sync(1 + c.t, 128);   // the previous use of the scratch has been read
  if(pd) pd[1] = clock64();
#pragma unroll
  for(int j = 0; j < 16; j++) c.scr[c.r * SCR_STRIDE + j] = g[j];
  named_bar_sync(1 + c.t, 128);
  if(pd) pd[2] = clock64();
  const float2 r = poolReduce(smem_u32(c.scr), c.maskRow ? smem_u32(c.maskRow) : 0u, c.e, P.NB, P.stride, P.HW, P.wMagic, P.tileRowW - P.W);
  sum = r.x; mx = r.y;
}

// acc[b] = sum_k in[b*inStride + k] * W[k*OC + oc] for the NB boards of a tile, k ascending (same summation order as a plain
// loop), with the weight loads issued 32 at a time: these tiny matmuls are pure L2 latency, not bandwidth.  Kdim % 16 == 0; `in` rows
// are 16-byte aligned and read four values per shared-memory access (broadcast: one wavefront per instruction whatever its width).
template <int NW>
__device__ __forceinline__ void pooledFma(uint32_t in, int inStride, int k0, const float (&w)[NW], int NB, float acc[MAX_NB]) {
#pragma unroll
  for(int j = 0; j < NW; j += 4)
#pragma unroll
    for(int b = 0; b < MAX_NB; b++)
      if(b < NB) {
        const float4 x = lds128(in + (uint32_t)(b * inStride + k0 + j) * 4u);
        acc[b] = fmaf(x.x, w[j], acc[b]); acc[b] = fmaf(x.y, w[j + 1], acc[b]);
        acc[b] = fmaf(x.z, w[j + 2], acc[b]); acc[b] = fmaf(x.w, w[j + 3], acc[b]);
      }
}
static_assert(MAX_NB == 4, "pooledMatmul returns one accumulator per board in a float4");
__device__ __noinline__ float4 pooledMatmul4(uint32_t in, int inStride, const float* __restrict__ W, int Kdim, int OC, int oc, int NB) {
  float acc[MAX_NB];
#pragma unroll
  for(int b = 0; b < MAX_NB; b++) acc[b] = 0.f;
  int k0 = 0;
  for(; k0 + 32 <= Kdim; k0 += 32) {
    float w[32];
#pragma unroll
    for(int j = 0; j < 32; j++) w[j] = __ldg(W + (size_t)(k0 + j) * OC + oc);
    pooledFma<32>(in, inStride, k0, w, NB, acc);
  }
  for(; k0 < Kdim; k0 += 16) {
    float w[16];
#pragma unroll
    for(int j = 0; j < 16; j++) w[j] = __ldg(W + (size_t)(k0 + j) * OC + oc);
    pooledFma<16>(in, inStride, k0, w, NB, acc);
  }
  return make_float4(acc[0], acc[1], acc[2], acc[3]);
}
__device__ __forceinline__ void pooledMatmul(const float* in, int inStride, const float* __restrict__ W, int Kdim, int OC, int oc, int NB,
                                             float acc[MAX_NB]) {
  const float4 r = pooledMatmul4(smem_u32(in), inStride, W, Kdim, OC, oc, NB);
  acc[0] = r.x; acc[1] = r.y; acc[2] = r.z; acc[3] = r.w;
}

#define KC_DBG (K::PROBE ? P.dbg : static_cast<long long*>(nullptr))
template <class K>
__device__ void epilogueBN(const TrunkParams& P, const LayerDesc& L, const EpiCtx& c) {
  const float* scale = c.par;
  const float* bias = c.par + K::MAXC;
  uint32_t src = c.tmemLane + (L.outSel ? K::MAXC : 0);
  const int nch = L.epiC / 16;
  // software pipeline: the TMEM load of chunk cc+1 is in flight while chunk cc is normalised, packed and published
  uint32_t ra[16], rb[16];
  tmem_ld16_issue(src, ra);
  tmem_ld16_wait(ra);
  if(c.dbg) c.dbg[2] = clock64();
  for(int cc = 0; cc < nch; cc += 2) {
    float v[16];
    if(cc + 1 < nch) tmem_ld16_issue(src + (cc + 1) * 16, rb);
#pragma unroll
    for(int i = 0; i < 16; i++) v[i] = __uint_as_float(ra[i]);
    publish16(c, cc, v, scale, bias, nullptr, L.act);
    if(c.dbg && cc == 0) c.dbg[3] = clock64();
    if(cc + 1 >= nch) break;
    tmem_ld16_wait(rb);
    if(cc + 2 < nch) tmem_ld16_issue(src + (cc + 2) * 16, ra);
#pragma unroll
    for(int i = 0; i < 16; i++) v[i] = __uint_as_float(rb[i]);
    publish16(c, cc + 1, v, scale, bias, nullptr, L.act);
    if(cc + 2 < nch) tmem_ld16_wait(ra);
  }
  if(c.dbg) c.dbg[7] = clock64();
}

// params: gpoolBN scale[G] bias[G] | Wg [3G][R] | midBN scale[R] bias[R]
template <class K>
__device__ void epilogueGPool(const TrunkParams& P, const LayerDesc& L, const EpiCtx& c) {
  const int R = L.epiC, G = L.gpoolC;
  const float* Wg = P.params + L.pOff + 2 * G;
  const float* ms = c.par;            // staged in shared memory before the accumulator wait: midBN scale[R], then gpoolBN scale[G]
  const float* mb = c.par + K::MAXC;  //                                                       midBN bias[R],  then gpoolBN bias[G]
  const float* gs = ms + R;
  const float* gb = mb + R;
  uint32_t src = c.tmemLane + K::MAXC;   // region S
  float* pooled = c.poolA;           // [NB][3G]
  for(int half = 0; half < G / 16; half++) {
    float v[16], g[16];
    tmem_ld16(src + R + half * 16, v);
#pragma unroll
    for(int j = 0; j < 16; j++) {
      float a = actf(fmaf(v[j], gs[half * 16 + j], gb[half * 16 + j]), L.gpoolAct);
      g[j] = c.valid ? a : 0.f;
    }
    float sum, mx;
    poolBoards16(P, c, g, sum, mx);
    if(!(c.e & 1) && (c.e >> 1) < P.NB * 16) {
      int b = c.e >> 5, j = (c.e >> 1) & 15;
      float mean = sum * (c.boardK ? c.boardK[b * 4] : P.invHW);
      pooled[b * 3 * G + half * 16 + j] = mean;
      pooled[b * 3 * G + G + half * 16 + j] = mean * (c.boardK ? c.boardK[b * 4 + 1] : P.poolScale1);
      pooled[b * 3 * G + 2 * G + half * 16 + j] = mx;
    }
  }
  named_bar_sync(1 + c.t, 128);
  for(int oc = c.e; oc < R; oc += 128) {
    float acc[MAX_NB];
    pooledMatmul(pooled, 3 * G, Wg, 3 * G, R, oc, P.NB, acc);
#pragma unroll
    for(int b = 0; b < MAX_NB; b++) if(b < P.NB) c.biasBuf[b * K::MAXC + oc] = acc[b];
  }
  named_bar_sync(1 + c.t, 128);
  const float* add = c.biasBuf + c.b * K::MAXC;
  for(int cc = 0; cc < R / 16; cc++) {
    float v[16];
    tmem_ld16(src + cc * 16, v);
    publish16(c, cc, v, ms, mb, add, L.act);
  }
}

__shared__ __align__(16) float sW3[2][4][MAX_V2 + 1];   // value / misc output matrices of the head, [tile][output][k], [V2] = bias
__shared__ __align__(16) float sHeadPar[2][11 * MAX_HEADC + MAX_V2];   // read with 128-bit loads (W2)
__shared__ uint8_t sCellRow[112];            // tile row of cell i of a board (pooling), up to 10x10   // small head parameters per tile (layout in epilogueHead), staged once per kernel

template <int HC>
__device__ __forceinline__ void stageHeadParams(const TrunkParams& P, const LayerDesc& L, int e, int t, float* par) {
  constexpr int HEADC = HC;   // shadows the global: every offset below is in units of this instantiation's head width
  const int V2 = P.v2C;
  const float* base = P.params + L.pOff;
  const float* p1s = base + 2 * HEADC + 3 * HEADC * HEADC;              // p1BN s, b | W2 | v1BN s, b: 256 contiguous floats
  const float* Wv2 = p1s + 8 * HEADC;
  const float* b2 = Wv2 + 3 * HEADC * V2;
  const float* Wv3 = b2 + V2;
  const float* b3 = Wv3 + V2 * 2;
  const float* Wsv3 = b3 + 2;
  const float* bsv3 = Wsv3 + V2 * 2;
  const float* Wown = bsv3 + 2;
  for(int i = e; i < 2 * HEADC; i += 128) par[i] = __ldg(base + i);
  for(int i = e; i < 8 * HEADC; i += 128) par[2 * HEADC + i] = __ldg(p1s + i);
  if(e < HEADC) par[10 * HEADC + e] = __ldg(Wown + e);
  for(int i = e; i < V2; i += 128) par[11 * HEADC + i] = __ldg(b2 + i);
  for(int i = e; i < 4 * V2; i += 128) {
    const int o = i / V2, k = i - o * V2;
    sW3[t][o][k] = __ldg(((o < 2) ? Wv3 : Wsv3) + k * 2 + (o & 1));
  }
  if(e < 4) sW3[t][e][V2] = __ldg(((e < 2) ? b3 : bsv3) + (e & 1));
}

// params (HC = K::HC head channels, PW = 3 HC): g1BN s[HC] b[HC] | Wpb [PW][HC] | p1BN s[HC] b[HC] | W2 [HC][4] | v1BN s[HC] b[HC] |
//         Wv2 [PW][V2] | b2 [V2] | Wv3 [V2][2] | b3[2] | Wsv3 [V2][2] | bsv3[2] | Wown [HC]
template <class K, bool PART_A, bool PART_B>
__device__ __forceinline__ void epilogueHead(const TrunkParams& P, const LayerDesc& L, const EpiCtx& c, int tileIndex, uint32_t barHead,
                                             const uint8_t* sSym, const int nRows, const float* par) {
  // Two parts.  A: everything that reads TMEM or sits between this item and the next one's layers -- pooling, ownership, policy.  B: the
  // value head (v2 matmul = three round trips to L2, value / misc outputs), which reads shared memory only and, for 32-channel heads,
  // runs after the next item's first-layer epilogue, i.e. under that item's second layer (TrunkParams::deferHead).
  static_assert(PART_A || PART_B, "");
  constexpr int HC = K::HC, PW = 3 * HC;
  // 32-channel heads keep p1 in registers and release TMEM region S before the pooled matmuls (the next item's first layer can be issued
  // under the rest of this epilogue); wider heads read p1 from TMEM 16 columns at a time when the policy is formed, and release then
  constexpr bool EARLY = HC <= 32;
  const int V2 = P.v2C;
  // small parameters were staged in shared memory before the accumulator wait (stageHeadParams): layout of c.par
  //   [0,HC) g1BN s | [HC,2HC) g1BN b | [2HC,3HC) p1BN s | [3HC,4HC) p1BN b | [4HC,8HC) W2 [HC][4] | [8HC,9HC) v1BN s | [9HC,10HC) v1BN b |
  //   [10HC,11HC) Wown | [11HC,11HC+V2) b2 ;  sW3[tile][o][k] = value / misc output matrices, [o][V2] = their biases
  const float* g1s = par;
  const float* g1b = par + HC;
  const float* p1s = par + 2 * HC;
  const float* p1b = par + 3 * HC;
  const float* W2 = par + 4 * HC;
  const float* v1s = par + 8 * HC;
  const float* v1b = par + 9 * HC;
  const float* Wown = par + 10 * HC;
  const float* b2 = par + 11 * HC;
  const float* Wpb = P.params + L.pOff + 2 * HC;
  const float* Wv2 = Wpb + PW * HC + 2 * HC + 4 * HC + 2 * HC;
  uint32_t src = c.tmemLane + K::MAXC;
  float* pooledG = c.poolA;   // [NB][PW]
  float* pooledV = c.poolB;   // [NB][PW]
  const bool hp = KC_DBG && blockIdx.x == 0 && c.t == 0 && c.e == 0 && tileIndex == 0;
  const int gameBase = tileIndex * P.NB;
  if constexpr(PART_A) {
  float own = 0.f;            // ownership = v1 (after BN / activation / mask) . Wown, formed while v1 passes through the pooling
  // 32-channel heads: the pooled policy bias is formed by all 128 threads, each one K quarter (24 pooled features) of one output
  // channel; its weights are requested now and arrive under the pooling
  constexpr int KQ = PW / 4;
  const int pbOc = c.e % HC, pbK = c.e / HC;
  float wpb[EARLY ? KQ : 1];
  if constexpr(EARLY) {
#pragma unroll
    for(int j = 0; j < KQ; j++) wpb[j] = __ldg(Wpb + (size_t)(pbK * KQ + j) * HC + pbOc);
  }
  // this row's output symmetry, fetched now: a global load takes 1,500+ clk here (L2 is streaming the next item's weights), and the
  // policy stores at the end of this epilogue depend on it
  const int symEarly = (P.sym && tileIndex * P.NB + c.b < nRows) ? P.sym[tileIndex * P.NB + c.b] : 0;
  if(hp) KC_DBG[24] = clock64();
  const bool poolOut = !(c.e & 1) && (c.e >> 1) < P.NB * 16;
  const int pb = c.e >> 5, pj = (c.e >> 1) & 15;
  // g1 -> BN -> ReLU -> gpool (eigenbackend.cpp:1290-1291) ; v1 -> BN -> ReLU -> value pool (:1364-1366)
  for(int half = 0; half < HC / 16; half++) {
    float v[16], g[16], sum, mx;
    tmem_ld16(src + HC + half * 16, v);
    if(hp && half < 2) KC_DBG[30 + 4 * half] = clock64();
#pragma unroll
    for(int j = 0; j < 16; j++) {
      float a = actf(fmaf(v[j], g1s[half * 16 + j], g1b[half * 16 + j]), P.g1Act);
      g[j] = c.valid ? a : 0.f;
    }
    poolBoards16(P, c, g, sum, mx, hp && half == 0 ? KC_DBG + 40 : nullptr);
    if(hp && half < 2) KC_DBG[31 + 4 * half] = clock64();
    const float kInv = c.boardK ? c.boardK[pb * 4] : P.invHW, kS1 = c.boardK ? c.boardK[pb * 4 + 1] : P.poolScale1,
                kS2 = c.boardK ? c.boardK[pb * 4 + 2] : P.poolScale2;
    if(poolOut) {
      float mean = sum * kInv;
      pooledG[pb * PW + half * 16 + pj] = mean;
      pooledG[pb * PW + HC + half * 16 + pj] = mean * kS1;
      pooledG[pb * PW + 2 * HC + half * 16 + pj] = mx;
    }
    tmem_ld16(src + 2 * HC + half * 16, v);
    if(hp && half < 2) KC_DBG[32 + 4 * half] = clock64();
#pragma unroll
    for(int j = 0; j < 16; j++) {
      float a = actf(fmaf(v[j], v1s[half * 16 + j], v1b[half * 16 + j]), P.v1Act);
      g[j] = c.valid ? a : 0.f;
      own = fmaf(g[j], Wown[half * 16 + j], own);
    }
    poolBoards16(P, c, g, sum, mx);
    if(hp && half < 2) KC_DBG[33 + 4 * half] = clock64();
    if(poolOut) {
      float mean = sum * kInv;
      pooledV[pb * PW + half * 16 + pj] = mean;
      pooledV[pb * PW + HC + half * 16 + pj] = mean * kS1;
      pooledV[pb * PW + 2 * HC + half * 16 + pj] = mean * kS2;
    }
  }
  if(hp) KC_DBG[25] = clock64();
  float p1[EARLY ? HC : 1];
  auto releaseTmem = [&]() {
    // all TMEM reads of this tile are done: the MMA warp may overwrite region S for the next item
    tc_fence_before();
    __syncwarp();
    if((c.r & 31) == 0) { if(c.remote) mbar_arrive_cluster(barHead); else mbar_arrive(barHead); }
    if(KC_DBG && blockIdx.x == 0 && c.t == 0 && c.e == 0 && tileIndex == 0) KC_DBG[18] = clock64();
  };
  if constexpr(EARLY) {
    float v[16];
#pragma unroll
    for(int q = 0; q < HC / 16; q++) {
      tmem_ld16(src + q * 16, v);
#pragma unroll
      for(int j = 0; j < 16; j++) p1[q * 16 + j] = v[j];
    }
    releaseTmem();
  }
  named_bar_sync(1 + c.t, 128);   // pooledG / pooledV complete, the pooling scratch has been read
  if(hp) KC_DBG[26] = clock64();
  if constexpr(EARLY) {
    float part[MAX_NB] = {0.f, 0.f, 0.f, 0.f};
    pooledFma<KQ>(smem_u32(pooledG), PW, pbK * KQ, wpb, P.NB, part);
#pragma unroll
    for(int b = 0; b < MAX_NB; b++) if(b < P.NB) c.biasBuf[(pbK * MAX_NB + b) * HC + pbOc] = part[b];   // 4 * MAX_NB * HC <= MAX_NB * MAXC floats
    named_bar_sync(1 + c.t, 128);
    // thread (board, channel): the four K quarters summed and folded into p1's normalisation: (p1 + bias) * s + b = p1 * s + (bias * s + b)
    if(c.e < P.NB * HC) {
      const float* q4 = c.biasBuf + c.e;   // c.e = board * HC + channel
      const float bias = (q4[0] + q4[MAX_NB * HC]) + (q4[2 * MAX_NB * HC] + q4[3 * MAX_NB * HC]);
      c.scr[K::HEAD_ADD_OFF + c.e] = fmaf(bias, p1s[c.e % HC], p1b[c.e % HC]);
    }
  } else if(c.e < HC) {
    float acc[MAX_NB];
    pooledMatmul(pooledG, PW, Wpb, PW, HC, c.e, P.NB, acc);
#pragma unroll
    for(int b = 0; b < MAX_NB; b++) if(b < P.NB) c.biasBuf[b * PW + c.e] = acc[b];
  }
  if(hp) KC_DBG[27] = clock64();
  named_bar_sync(1 + c.t, 128);
  if(hp) KC_DBG[28] = clock64();
  int game = gameBase + c.b;
  if(c.maskRow && !c.valid && c.slotCell >= 0 && gameBase + c.slotB < nRows) {
    // a cell of the nnXLen x nnYLen slot that is off this board: masked p1 / v1 through the bias-free 1x1 output convolutions = 0
    const int g2 = gameBase + c.slotB;
    const int s = P.sym ? P.sym[g2] : 0;
    const int dst = sSym[s * P.HW + c.slotCell];
    float* pol = P.policy + (size_t)g2 * 4 * P.HW;
    pol[dst] = 0.f; pol[P.HW + dst] = 0.f; pol[2 * P.HW + dst] = 0.f; pol[3 * P.HW + dst] = 0.f;
    P.own[(size_t)g2 * P.HW + dst] = 0.f;
  }
  // the policy: p1 + pooled bias -> BN -> activation -> the 1x1 p2 convolution (eigenbackend.cpp:1292-1298).  Every thread runs the loop
  // (the late form reads TMEM warp-wide); only the threads of board cells store.
  const float* add = EARLY ? c.scr + K::HEAD_ADD_OFF + c.b * HC : c.biasBuf + c.b * PW;
  float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f;
  if constexpr(EARLY) {
#pragma unroll
    for(int k = 0; k < HC; k++) {
      float a = actf(fmaf(p1[k], p1s[k], add[k]), P.p1Act);   // add = pooled bias * s + b
      float4 w = *(reinterpret_cast<const float4*>(W2) + k);
      o0 = fmaf(a, w.x, o0); o1 = fmaf(a, w.y, o1); o2 = fmaf(a, w.z, o2); o3 = fmaf(a, w.w, o3);
    }
  } else {
    for(int q = 0; q < HC / 16; q++) {
      float v[16];
      tmem_ld16(src + q * 16, v);
#pragma unroll
      for(int j = 0; j < 16; j++) {
        const int k = q * 16 + j;
        float a = actf(fmaf(v[j] + add[k], p1s[k], p1b[k]), P.p1Act);
        float4 w = *(reinterpret_cast<const float4*>(W2) + k);
        o0 = fmaf(a, w.x, o0); o1 = fmaf(a, w.y, o1); o2 = fmaf(a, w.z, o2); o3 = fmaf(a, w.w, o3);
      }
    }
    releaseTmem();
  }
  if(c.valid && game < nRows) {
    int s = symEarly;
    int dst = sSym[s * P.HW + c.cell];
    float* pol = P.policy + (size_t)game * 4 * P.HW;
    if(P.permuteDirs) {   // play mode: the net's direction channel d is direction symDir(d, s) of the original position
      pol[symDirInline(0, s) * P.HW + dst] = o0; pol[symDirInline(1, s) * P.HW + dst] = o1;
      pol[symDirInline(2, s) * P.HW + dst] = o2; pol[symDirInline(3, s) * P.HW + dst] = o3;
    } else {
      pol[dst] = o0; pol[P.HW + dst] = o1; pol[2 * P.HW + dst] = o2; pol[3 * P.HW + dst] = o3;
    }
    P.own[(size_t)game * P.HW + dst] = own;
  }
  }   // PART_A
  if constexpr(PART_B) {
    for(int oc = c.e; oc < V2; oc += 128) {
      float acc[MAX_NB];
      pooledMatmul(pooledV, PW, Wv2, PW, V2, oc, P.NB, acc);
      const float bias2 = b2[oc];
#pragma unroll
      for(int b = 0; b < MAX_NB; b++) if(b < P.NB) c.v2buf[b * MAX_V2 + oc] = actf(acc[b] + bias2, P.v2Act);
    }
    named_bar_sync(1 + c.t, 128);
    // value / misc outputs: 8 lanes per (board, output), each a slice of k, combined by an xor butterfly
    const int q = c.e >> 3, seg = c.e & 7;
    const int b = q >> 2, o = q & 3;
    float acc = 0.f;
    if(b < P.NB) {
      const int kn = V2 >> 3;
      for(int k = seg * kn; k < (seg + 1) * kn; k++) acc = fmaf(c.v2buf[b * MAX_V2 + k], sW3[c.t][o][k], acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    const int game = gameBase + b;
    if(seg == 0 && b < P.NB && game < nRows) {
      acc += sW3[c.t][o][V2];
      if(o < 2) P.value[(size_t)game * 2 + (o & 1)] = acc; else P.misc[(size_t)game * 2 + (o & 1)] = acc;
    }
    if(hp) KC_DBG[29] = clock64();
  }   // PART_B
}

// The MMA issuer of one tile: a whole warp runs the (warp-uniform) control flow and waits on the
// barriers; one elected lane issues the tcgen05.mma / tcgen05.commit instructions.  Descriptors are kept
// as (hi, lo) words: hi holds SBO = 128 B and the sm_100 version bit, lo = start address >> 4 |
// LBO >> 4 << 16; a tap or a K-chunk is a plain add on lo.  For 3x3 layers the K loop is chunk-major:
// 16 input channels x (3 stages x 3 taps), each stage = one kernel row (dy), so the row shift is
// (dy-1)*tileRowW + (dx-1).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <class K>
__device__ __forceinline__ void mmaIssuer(const TrunkParams& P, const int t, const uint32_t sbase, const uint32_t bars,
                                          const uint32_t tmemBase, volatile int* abortFlag, const int numItems) {
  uint32_t slot = 0, phase = 0, itemCount = 0, chunkPhase = 0;
  const bool leader = elect_one();
  const uint32_t descHi = (128u >> 4) | (1u << 14);
  const uint32_t aLo0 = (((sbase + K::OFF_ACT + t * K::ACT_BYTES + HALO_ROWS * 16) & 0x3FFFFu) >> 4) | ((uint32_t)(CHUNK_BYTES >> 4) << 16);
  const uint32_t ringLo0 = ((sbase + K::OFF_RING) & 0x3FFFFu) >> 4;
  const uint32_t barFull = bars + K::BAR_FULL * 8, barEmpty = bars + K::BAR_EMPTY * 8, barChunk = bars + (K::BAR_CHUNK + t * K::NCH) * 8;
  const uint32_t barFullP = bars + K::BAR_FULLP * 8;
  auto desc = [descHi](uint32_t lo) { return ((uint64_t)descHi << 32) | lo; };
  // pair mode: this is the leader CTA; one MMA covers tile t of both CTAs, barriers carry both CTAs' arrivals
  auto mma = [&](uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if constexpr(K::PAIR) umma_bf16_2cta(d, a, b, idesc, acc); else umma_bf16(d, a, b, idesc, acc);
  };
  auto commit = [&](uint32_t bar) { if constexpr(K::PAIR) umma_commit_2cta(bar, 3); else umma_commit(bar); };
  auto waitc = [&](uint32_t bar, uint32_t parity, int code) {   // barriers that guard data written by threads (of either CTA)
    return mbar_wait(bar, parity, abortFlag, code);
  };
  const int itemStep = K::PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
  // enforced lag between the two tiles of a CTA (TrunkParams::skew): tile 0 publishes how many stages it has issued, tile 1 starts a
  // layer only when tile 0 is `skew` stages into it (or through it, if the layer is shorter)
  const bool skewed = K::NT == 2 && P.skew > 0;
  volatile int* progress = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(__cvta_shared_to_generic(sbase)) + K::OFF_PROG);
  int stagesIssued = 0;   // by this issuer, over all layers and items
  auto stageIssued = [&]() {
    stagesIssued++;
    if(skewed && t == 0 && leader) *progress = stagesIssued;
  };
  auto waitForLeadTile = [&](int nst) -> bool {   // tile 1, at the start of a layer of nst stages
    if(!(skewed && t == 1)) return true;
    const int need = stagesIssued + min(P.skew, nst);
    if(*progress >= need) return true;
    long long t0 = clock64();
    for(uint32_t it = 1;; it++) {
      if(*progress >= need) return true;
      if((it & 63u) == 0) {
        if(*abortFlag != 0) return false;
        if(clock64() - t0 > (1LL << 32)) { *abortFlag = 28; return false; }
      }
    }
  };
  for(int item = K::PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x; item < numItems; item += itemStep, itemCount++) {   // pair mode: item = item pair
    if(!mbar_wait(bars + (K::BAR_IN + t) * 8, itemCount & 1, abortFlag, 21)) return;
    if(K::PAIR && !waitc(bars + (K::BAR_INP + t) * 8, itemCount & 1, 26)) return;
    bool needHead = itemCount > 0;   // region S still belongs to the previous item's head epilogue; region T (layer 0) does not
    tc_fence_after();
    if(KC_DBG && blockIdx.x == 0 && t == 0 && itemCount == 1 && leader) KC_DBG[20] = clock64();
    for(int l = 0; l < P.numLayers; l++) {
      const int nk = P.layers[l].nk, ntaps = P.layers[l].ntaps, N = P.layers[l].N;
      // whole-item timeline (KC_TRUNK_PROBE): CTA 0, second item (steady state): [64 + ((t * MAX_LAYERS + l) * 8 + k)]
      //   k = 0 issuer reaches the layer, 1 first chunk available, 2 layer issued and committed, 3 epilogue sees the accumulator, 4 epilogue done
      long long* tl = (KC_DBG && blockIdx.x == 0 && itemCount == 1 && leader) ? KC_DBG + 64 + (t * MAX_LAYERS + l) * 8 : nullptr;
      if(tl) tl[0] = clock64();
      if(needHead && P.layers[l].outSel) {
        if(!waitc(bars + (K::BAR_HEAD + t) * 8, (itemCount - 1) & 1, 22)) return;
        tc_fence_after();
        needHead = false;
      }
      const uint32_t idesc = idesc_f16kind_f32(128 * K::NCTA, N, (uint32_t)P.opFmt, (uint32_t)P.opFmt);   // A and B must share the format
      const uint32_t d = tmemBase + t * (2 * K::MAXC) + (P.layers[l].outSel ? K::MAXC : 0);
      const uint32_t bStep = 2 * N / K::NCTA;                    // one K-step of weights = N*32 bytes (per CTA: its half of the rows)
      const uint32_t bLbo = (uint32_t)(N / K::NCTA) << 16;       // LBO = rows*16 bytes
      uint32_t accum = P.layers[l].accumulate ? 1u : 0u;
      if(ntaps == 9) {
        const int nchunks = nk / 9;
        constexpr int ROWS_PER_STAGE = K::KSTEPS / 3;          // kernel rows (dy) per weight stage: 1 or 3
        if(!waitForLeadTile(nchunks * (3 / ROWS_PER_STAGE))) return;
        for(int cc = 0; cc < nchunks; cc++) {
          const bool probe = KC_DBG && blockIdx.x == 0 && t == 0 && itemCount == 0 && l == 6 && cc == 0 && leader;
          if(probe) KC_DBG[6] = clock64();
          if(l > 0) {
            const uint32_t bit = 1u << cc;
            if(!waitc(barChunk + cc * 8, (chunkPhase & bit) ? 1 : 0, 24)) return;
            chunkPhase ^= bit;
          }
          if(probe) KC_DBG[4] = clock64();
          if(tl && cc == 0) tl[1] = clock64();
          if(KC_DBG && blockIdx.x == 0 && t == 0 && itemCount == 1 && l == 1 && cc == 0 && leader) KC_DBG[23] = clock64();
          const uint32_t aLoC = aLo0 + cc * (2 * CHUNK_BYTES >> 4);
#pragma unroll
          for(int dy0 = 0; dy0 < 3; dy0 += ROWS_PER_STAGE) {
            if(!mbar_wait(barFull + slot * 8, phase, abortFlag, 23)) return;
            if(K::PAIR && !mbar_wait(barFullP + slot * 8, phase, abortFlag, 27)) return;
            tc_fence_after();
            const uint32_t bLo = (ringLo0 + slot * (K::STAGE_BYTES >> 4)) | bLbo;
            if(leader) {
#pragma unroll
              for(int r = 0; r < ROWS_PER_STAGE; r++) {
                const uint32_t aLoR = aLoC + (dy0 + r - 1) * P.tileRowW - 1;
                const uint32_t bLoR = bLo + 3 * r * bStep;
                mma(d, desc(aLoR), desc(bLoR), idesc, accum);
                mma(d, desc(aLoR + 1), desc(bLoR + bStep), idesc, 1u);
                mma(d, desc(aLoR + 2), desc(bLoR + 2 * bStep), idesc, 1u);
                accum = 1u;
              }
              commit(barEmpty + slot * 8);
              if(probe && dy0 == 0) KC_DBG[5] = clock64();
              if(KC_DBG && blockIdx.x == 0 && t == 0 && itemCount == 0 && l == 6 && dy0 + ROWS_PER_STAGE == 3 && cc < 7) KC_DBG[8 + cc] = clock64();
            }
            __syncwarp();
            accum = 1u;
            stageIssued();
            if(++slot == K::NSTAGES) { slot = 0; phase ^= 1; }
          }
        }
      } else {
        // 1x1 layers: K-step = one 16-channel chunk, stages of up to 3 K-steps
        const int nst = (nk + K::KSTEPS - 1) / K::KSTEPS;
        if(!waitForLeadTile(nst)) return;
        for(int s = 0; s < nst; s++) {
          if(!mbar_wait(barFull + slot * 8, phase, abortFlag, 23)) return;
          if(K::PAIR && !mbar_wait(barFullP + slot * 8, phase, abortFlag, 27)) return;
          const uint32_t bLo = (ringLo0 + slot * (K::STAGE_BYTES >> 4)) | bLbo;
          const int ks = min(K::KSTEPS, nk - s * K::KSTEPS);
          for(int kk = 0; kk < ks; kk++) {
            const int cc = s * K::KSTEPS + kk;
            if(l > 0) {
              const uint32_t bit = 1u << cc;
              if(!waitc(barChunk + cc * 8, (chunkPhase & bit) ? 1 : 0, 24)) return;
              chunkPhase ^= bit;
            }
            tc_fence_after();
            if(leader) mma(d, desc(aLo0 + cc * (2 * CHUNK_BYTES >> 4)), desc(bLo + kk * bStep), idesc, accum);
            __syncwarp();
            accum = 1u;
          }
          if(leader) commit(barEmpty + slot * 8);
          __syncwarp();
          stageIssued();
          if(++slot == K::NSTAGES) { slot = 0; phase ^= 1; }
        }
      }
      if(leader) {
        commit(bars + (K::BAR_ACC + t) * 8);
        if(tl) tl[2] = clock64();
        if(l == P.numLayers - 1) commit(bars + (K::BAR_ACTFREE + t) * 8);
        if(KC_DBG && blockIdx.x == 0 && t == 0 && itemCount == 0 && l == 5) KC_DBG[0] = clock64();
        if(KC_DBG && blockIdx.x == 0 && t == 0 && itemCount == 0 && l == 6) KC_DBG[15] = clock64();
        if(KC_DBG && blockIdx.x == 0 && t == 0 && itemCount == 0 && l == P.numLayers - 1) KC_DBG[16] = clock64();
        if(KC_DBG && blockIdx.x == 0 && t == 0 && itemCount == 1 && l == 0) KC_DBG[21] = clock64();
      }
      __syncwarp();
    }
  }
}

template <class K>
__global__ void __launch_bounds__(K::REALLOC ? 512 : K::THREADS, 1) trunk_kernel(const TrunkParams P) {
  constexpr int NT = K::NT;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + K::OFF_BAR;
  volatile int* abortFlag = P.abortFlag;
  uint8_t* sSym = smem + K::OFF_SYM;
  const int nRows = P.nDev ? min(__ldg(P.nDev), P.n) : P.n;
  const int numItems = P.nDev ? ((nRows + P.NB - 1) / P.NB + NT - 1) / NT : P.numItems;
  // work distribution: a "unit" is what one CTA (or one CTA pair) takes per round; in pair mode the two CTAs of a cluster
  // take the two items of a unit and run the same number of rounds (the odd item out is a ghost: its tiles are read
  // from the last real item, its outputs land beyond nRows and are dropped)
  const uint32_t rank = K::PAIR ? cluster_ctarank() : 0u;
  const int unit0 = K::PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x, unitStep = K::PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int numUnits = K::PAIR ? (numItems + 1) / 2 : numItems;
  auto itemOf = [&](int unit) { return K::PAIR ? 2 * unit + (int)rank : unit; };

  // ---- one-time setup ----
  for(int i = threadIdx.x; i < NT * K::ACT_BYTES / 16; i += K::THREADS) reinterpret_cast<uint4*>(smem + K::OFF_ACT)[i] = make_uint4(0, 0, 0, 0);
  for(int i = threadIdx.x; i < 8 * P.HW; i += K::THREADS) sSym[i] = P.dstOfSrcRev[i];
  for(int i = threadIdx.x; i < P.HW; i += K::THREADS) sCellRow[i] = (uint8_t)((i / P.W) * P.tileRowW + i % P.W);
  if(threadIdx.x == 0) {
    *reinterpret_cast<volatile int*>(smem + K::OFF_PROG) = 0;
    for(int i = 0; i < K::NSTAGES; i++) {
      mbar_init(bars + (K::BAR_FULL + i) * 8, 1); mbar_init(bars + (K::BAR_EMPTY + i) * 8, NT);   // every MMA issuer releases a slot
      mbar_init(bars + (K::BAR_FULLP + i) * 8, 1);
      if(i == 0) mbar_init(bars + K::BAR_SKEW * 8, 1);
    }
    for(int t = 0; t < NT; t++) {
      mbar_init(bars + (K::BAR_ACC + t) * 8, 1); mbar_init(bars + (K::BAR_ACTFREE + t) * 8, 1); mbar_init(bars + (K::BAR_IN + t) * 8, 1);
      mbar_init(bars + (K::BAR_INP + t) * 8, 1);
      mbar_init(bars + (K::BAR_HEAD + t) * 8, 4 * K::NCTA);   // one arrival per epilogue warp (of both CTAs in pair mode)
      for(int c = 0; c < K::NCH; c++) mbar_init(bars + (K::BAR_CHUNK + t * K::NCH + c) * 8, 4 * K::NCTA);
    }
    fence_mbar_init();
  }
  fence_proxy_async();
  if(warp == 1) {
    if constexpr(K::PAIR) { tmem_alloc2(sbase + K::OFF_TMEM, 512); tmem_relinquish2(); }
    else { tmem_alloc(sbase + K::OFF_TMEM, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr(K::PAIR) cluster_sync(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = *reinterpret_cast<volatile uint32_t*>(smem + K::OFF_TMEM);
  if constexpr(K::REALLOC) {
    // register re-allocation between the warpgroups: the launch takes 128 registers per thread (launch bound 512), the producer /
    // issuer / relay warpgroup gives back all but 56, the two epilogue warpgroups grow to 160
    if(warp < 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 160;");
  }

  if(warp == 0) {
    // =========================== TMA producer (every CTA; pair mode: its half of each weight stage) ===========================
    if(lane == 0) {
      uint32_t slot = 0, phase = 0, itemCount = 0;
      bool alive = true;
      for(int unit = unit0; unit < numUnits && alive; unit += unitStep, itemCount++) {
        const int item = min(itemOf(unit), numItems - 1);
        for(int t = 0; t < NT && alive; t++) {
          if(itemCount > 0) alive = mbar_wait(bars + (K::BAR_ACTFREE + t) * 8, (itemCount - 1) & 1, abortFlag, 11);
          if(!alive) break;
          uint32_t bar = bars + (K::BAR_IN + t) * 8;
          mbar_arrive_expect_tx(bar, 2 * TILE_ROWS * 16);
          const uint4* src = P.tiles + (size_t)(item * NT + t) * 2 * TILE_ROWS;
          uint32_t dst = sbase + K::OFF_ACT + t * K::ACT_BYTES + HALO_ROWS * 16;
          bulk_g2s(dst, src, TILE_ROWS * 16, bar);
          bulk_g2s(dst + CHUNK_BYTES, src + TILE_ROWS, TILE_ROWS * 16, bar);
        }
        for(int l = 0; l < P.numLayers && alive; l++) {
          const LayerDesc L = P.layers[l];
          const int nst = (L.nk + K::KSTEPS - 1) / K::KSTEPS;
          const uint8_t* w = P.wstream + L.wOffset;
          for(int s = 0; s < nst; s++) {
            int ks = min(K::KSTEPS, L.nk - s * K::KSTEPS);
            uint32_t bytes = (uint32_t)ks * L.N * 32;          // whole stage; the stream holds [rank 0 half][rank 1 half] per stage
            alive = mbar_wait(bars + (K::BAR_EMPTY + slot) * 8, phase ^ 1, abortFlag, 12);
            if(!alive) break;
            uint32_t bar = bars + (K::BAR_FULL + slot) * 8;
            mbar_arrive_expect_tx(bar, bytes / K::NCTA);
            bulk_g2s(sbase + K::OFF_RING + slot * K::STAGE_BYTES, w + rank * (bytes / K::NCTA), bytes / K::NCTA, bar);
            w += bytes;
            if(++slot == K::NSTAGES) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if(warp >= 1 && warp <= NT) {
    // =========================== MMA issuers: one thread per tile (pair mode: in the leader CTA, for both CTAs) ===========================
    if(rank == 0) mmaIssuer<K>(P, warp - 1, sbase, bars, tmemBase, abortFlag, numUnits);
  } else if(warp < 4) {
    // warp 3.  Pair mode, peer CTA: relay -- tells the leader's issuers when this CTA's input tiles and weight halves have
    // landed (a bulk copy can only signal a barrier of its own CTA).  Otherwise idle.
    if(K::PAIR && rank == 1 && warp == 3 && lane == 0) {
      uint32_t slot = 0, phase = 0, itemCount = 0;
      bool alive = true;
      for(int unit = unit0; unit < numUnits && alive; unit += unitStep, itemCount++) {
        for(int t = 0; t < NT && alive; t++) {
          alive = mbar_wait(bars + (K::BAR_IN + t) * 8, itemCount & 1, abortFlag, 13);
          if(alive) mbar_arrive_cluster(mapa(bars + (K::BAR_INP + t) * 8, 0));
        }
        for(int l = 0; l < P.numLayers && alive; l++) {
          const int nst = (P.layers[l].nk + K::KSTEPS - 1) / K::KSTEPS;
          for(int s = 0; s < nst && alive; s++) {
            alive = mbar_wait(bars + (K::BAR_FULL + slot) * 8, phase, abortFlag, 14);
            if(alive) mbar_arrive_cluster(mapa(bars + (K::BAR_FULLP + slot) * 8, 0));
            if(++slot == K::NSTAGES) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else {
    // =========================== epilogue warps ===========================
    EpiCtx c;
    c.t = (warp - 4) >> 2;
    const int q = warp & 3;               // TMEM lane quadrant this warp may access
    c.r = q * 32 + lane;
    c.e = c.r;
    int y = c.r / P.tileRowW, rr = c.r - y * P.tileRowW;
    c.b = rr / P.stride;
    int x = rr - c.b * P.stride;
    const bool slotValid = (y < P.H) && (x < P.W);
    c.valid = slotValid;
    c.cell = y * P.W + x;
    if(!c.valid) { c.b = 0; c.cell = 0; }
    c.slotB = c.b; c.slotCell = slotValid ? c.cell : -1;
    float* maskRowW = reinterpret_cast<float*>(smem + K::OFF_MASK) + c.t * 128;
    float* boardKW = reinterpret_cast<float*>(smem + K::OFF_MASK) + NT * 128 + c.t * MAX_NB * 4;
    c.cellRow = sCellRow;
    c.maskRow = P.masked ? maskRowW : nullptr;
    c.boardK = P.masked ? boardKW : nullptr;
    uint32_t itemCount = 0;
    c.tmemLane = tmemBase + ((uint32_t)(q * 32) << 16) + c.t * (2 * K::MAXC);
    c.act = smem + K::OFF_ACT + c.t * K::ACT_BYTES;
    c.remote = K::PAIR && rank != 0;
    c.f16 = P.opFmt == 0;
    c.barChunk = bars + (K::BAR_CHUNK + c.t * K::NCH) * 8;
    uint32_t barHead = bars + (K::BAR_HEAD + c.t) * 8;
    if(c.remote) { c.barChunk = mapa(c.barChunk, 0); barHead = mapa(barHead, 0); }
    c.scr = reinterpret_cast<float*>(smem + K::OFF_SCR) + c.t * 128 * SCR_STRIDE;
    c.poolA = reinterpret_cast<float*>(smem + K::OFF_POOLA) + c.t * MAX_NB * K::POOLW;
    c.poolB = reinterpret_cast<float*>(smem + K::OFF_POOLB) + c.t * MAX_NB * K::POOLW;
    c.biasBuf = c.scr;                          // aliases (see TrunkCfg)
    c.v2buf = c.scr + MAX_NB * K::MAXC;
    uint32_t layerCount = 0;
    bool alive = true;
    int pendingTile = -1;   // tile index of the item whose value head (part B of the head epilogue) is still to run
    stageHeadParams<K::HC>(P, P.layers[P.numLayers - 1], c.e, c.t, sHeadPar[c.t]);   // read after the first layer's named barrier at the earliest
    for(int unit = unit0; unit < numUnits && alive; unit += unitStep) {
      const int item = itemOf(unit);
      for(int l = 0; l < P.numLayers && alive; l++, layerCount++) {
        const LayerDesc L = P.layers[l];
        {
          // stage this layer's folded BN (the one applied to its output) while its MMAs are still running
          float* par = reinterpret_cast<float*>(smem + K::OFF_PAR) + (c.t * 2 + (layerCount & 1)) * 2 * K::MAXC;
          if(L.epi != EPI_HEAD) {
            const float* sc = P.params + L.pOff + (L.epi == EPI_GPOOL ? 2 * L.gpoolC + 3 * L.gpoolC * L.epiC : 0);
            for(int i = c.e; i < L.epiC; i += 128) {
              par[i] = __ldg(sc + i);
              par[K::MAXC + i] = __ldg(sc + L.epiC + i);
            }
            if(L.epi == EPI_GPOOL && c.e < L.gpoolC) {   // gpoolBN behind midBN (epiC + gpoolC <= MAXC)
              par[L.epiC + c.e] = __ldg(P.params + L.pOff + c.e);
              par[K::MAXC + L.epiC + c.e] = __ldg(P.params + L.pOff + L.gpoolC + c.e);
            }
          } else {
            par = sHeadPar[c.t];   // staged once, before the first item
          }
          c.par = par;
          named_bar_sync(1 + c.t, 128);
        }
        alive = mbar_wait(bars + (K::BAR_ACC + c.t) * 8, layerCount & 1, abortFlag, 31);
        if(!alive) break;
        if(P.masked && l == 0) {
          // the mask of this item's boards = input channel 0, still in the activation tile (the first layer's MMAs are done, its
          // epilogue has not published yet).  The tile was written by the bulk copy: observe its barrier before reading.
          alive = mbar_wait(bars + (K::BAR_IN + c.t) * 8, itemCount & 1, abortFlag, 32);
          if(!alive) break;
          const uint16_t onBoard = *reinterpret_cast<const volatile uint16_t*>(c.act + (size_t)(HALO_ROWS + c.r) * 16);
          c.valid = c.slotCell >= 0 && onBoard != 0;
          maskRowW[c.r] = c.valid ? 1.0f : 0.0f;
          named_bar_sync(1 + c.t, 128);
          if(c.e < P.NB) {
            float cnt = 0.f;
            for(int yy = 0; yy < P.H; yy++)
              for(int xx = 0; xx < P.W; xx++) cnt += maskRowW[yy * P.tileRowW + c.e * P.stride + xx];
            const float sq = sqrtf(cnt);
            boardKW[c.e * 4] = cnt > 0.f ? 1.0f / cnt : 0.f;
            boardKW[c.e * 4 + 1] = (sq - 14.0f) * 0.1f;
            boardKW[c.e * 4 + 2] = (sq - 14.0f) * (sq - 14.0f) * 0.01f - 0.1f;
          }
          named_bar_sync(1 + c.t, 128);
        }
        long long* tl = (KC_DBG && blockIdx.x == 0 && c.e == 0 && (int)layerCount >= P.numLayers && (int)layerCount < 2 * P.numLayers)
                          ? KC_DBG + 64 + (c.t * MAX_LAYERS + l) * 8 : nullptr;
        if(tl) tl[3] = clock64();
        c.dbg = (KC_DBG && blockIdx.x == 0 && c.t == 0 && c.e == 0 && layerCount == 5) ? KC_DBG : nullptr;
        if(c.dbg) c.dbg[1] = clock64();
        const bool probeHead = KC_DBG && blockIdx.x == 0 && c.t == 0 && c.e == 0 && (int)layerCount == P.numLayers - 1;
        if(probeHead) KC_DBG[17] = clock64();
        if(KC_DBG && blockIdx.x == 0 && c.t == 0 && c.e == 0 && (int)layerCount == P.numLayers) KC_DBG[22] = clock64();
        tc_fence_after();
        // the head epilogue is two calls: part A here, part B (the value head: shared memory only) either right away or -- 32-channel
        // heads with another item to come -- after that item's first-layer epilogue, under its second layer.  One call site each: the
        // code is inlined, and the kernel has to share a 32 KB instruction cache.
        bool valueHeadNow = false;
        if(L.epi == EPI_BN) {
          epilogueBN<K>(P, L, c);
          valueHeadNow = pendingTile >= 0;
        } else if(L.epi == EPI_GPOOL) epilogueGPool<K>(P, L, c);
        else {
          epilogueHead<K, true, false>(P, L, c, item * NT + c.t, barHead, sSym, nRows, sHeadPar[c.t]);
          pendingTile = item * NT + c.t;
          valueHeadNow = !(K::HC <= 32 && P.deferHead && unit + unitStep < numUnits);
        }
        if(valueHeadNow) {
          epilogueHead<K, false, true>(P, P.layers[P.numLayers - 1], c, pendingTile, barHead, sSym, nRows, sHeadPar[c.t]);
          pendingTile = -1;
        }
        if(probeHead) KC_DBG[19] = clock64();
        if(tl) tl[4] = clock64();
      }
      itemCount++;
    }
  }
  // ---- teardown ----
  tc_fence_before();
  if constexpr(K::PAIR) cluster_sync(); else __syncthreads();
  if(warp == 1) {
    tc_fence_after();
    if constexpr(K::PAIR) tmem_dealloc2(tmemBase, 512); else tmem_dealloc(tmemBase, 512);
  }
}

#undef KC_DBG
// kc_forward input conversion: raw fp32 rows (NCHW/NHWC) + global -> symmetrised bf16 tiles
__global__ void k_convert_tiles(const float* __restrict__ raw, const float* __restrict__ rawGlobal, const int8_t* __restrict__ sym,
                                const uint8_t* __restrict__ dstOfSrc, uint4* __restrict__ tiles, int n, int numTiles,
                                int NB, int W, int H, int rawNHWC, int permuteDirs, int f16) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if(j >= numTiles * 256) return;
  const int HW = W * H, stride = W + 1, tileRowW = NB * stride;
  int tile = j >> 8, chunk = (j >> 7) & 1, row = j & 127;
  int y = row / tileRowW, rr = row - y * tileRowW;
  int b = rr / stride, x = rr - b * stride;
  int game = tile * NB + b;
  uint4 v = make_uint4(0, 0, 0, 0);
  if(y < H && x < W && game < n) {
    int s = sym ? sym[game] : 0;
    // destination cell (y,x) comes from the source cell whose image under the symmetry is (y,x)
    int dstCell = y * W + x, srcCell = 0;
    for(int p = 0; p < HW; p++) if(dstOfSrc[s * HW + p] == dstCell) srcCell = p;
    float f[8];
#pragma unroll
    for(int q = 0; q < 8; q++) {
      int c = chunk * 8 + q;
      if(permuteDirs) c = playModeChannel(c, s);   // destination channel c shows the source's channel playModeChannel(c)
      if(c < 15) f[q] = rawNHWC ? raw[((size_t)game * HW + srcCell) * 15 + c] : raw[((size_t)game * 15 + c) * HW + srcCell];
      else f[q] = rawGlobal[game];
    }
    v = make_uint4(packOp(f[0], f[1], f16), packOp(f[2], f[3], f16), packOp(f[4], f[5], f16), packOp(f[6], f[7], f16));
  }
  tiles[((size_t)tile * 2 + chunk) * 128 + row] = v;
}

// UMMA self-test kernel: D[128][N] = A[shift .. shift+128][K] * B[N][K]^T with the same descriptor
// construction as the trunk (row-shifted A operand), one CTA.
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(const __nv_bfloat16* __restrict__ A, const __nv_bfloat16* __restrict__ B,
                                                            float* __restrict__ D, int rowsA, int N, int K, int shift, int ws, int* abortFlag) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t barStorage;
  __shared__ uint32_t tmemSlot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sA = smem;                                   // [K/8][rowsA][8]
  uint8_t* sB = smem + (size_t)(K / 8) * rowsA * 16;    // [K/8][N][8]
  for(int i = threadIdx.x; i < rowsA * K; i += 128) {
    int r = i / K, k = i % K;
    reinterpret_cast<__nv_bfloat16*>(sA)[((size_t)(k / 8) * rowsA + r) * 8 + (k % 8)] = A[i];
  }
  for(int i = threadIdx.x; i < N * K; i += 128) {
    int r = i / K, k = i % K;
    reinterpret_cast<__nv_bfloat16*>(sB)[((size_t)(k / 8) * N + r) * 8 + (k % 8)] = B[i];
  }
  uint32_t bar = smem_u32(&barStorage);
  if(threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
  fence_proxy_async();
  if(warp == 0) { tmem_alloc(smem_u32(&tmemSlot), 256); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmemBase = *reinterpret_cast<volatile uint32_t*>(&tmemSlot);
  if(threadIdx.x == 0 && ws >= 10) {
    // throughput probe: `ws` back-to-back MMAs over the same operands, two accumulators alternating like the
    // trunk's two tiles; D[0] receives cycles per MMA
    uint32_t idesc = idesc_bf16_f32(128, N);
    uint64_t adesc = smem_desc(smem_u32(sA) + shift * 16, (uint32_t)rowsA * 16, 128);
    uint64_t bdesc = smem_desc(smem_u32(sB), (uint32_t)N * 16, 128);
    long long t0 = clock64();
    const uint32_t dA = tmemBase, dB = tmemBase + N;
    for(int i = 0; i < ws; i += 8) {
      umma_bf16(dA, adesc, bdesc, idesc, 1u); umma_bf16(dB, adesc + 1, bdesc, idesc, 1u);
      umma_bf16(dA, adesc + 2, bdesc, idesc, 1u); umma_bf16(dB, adesc, bdesc, idesc, 1u);
      umma_bf16(dA, adesc + 1, bdesc, idesc, 1u); umma_bf16(dB, adesc + 2, bdesc, idesc, 1u);
      umma_bf16(dA, adesc, bdesc, idesc, 1u); umma_bf16(dB, adesc + 1, bdesc, idesc, 1u);
    }
    umma_commit(bar);
    mbar_wait(bar, 0, abortFlag, 42);
    long long t1 = clock64();
    D[0] = (float)(t1 - t0) / (float)ws;
  } else if(threadIdx.x == 0) {
    uint32_t idesc = idesc_bf16_f32(128, N);
    for(int kc = 0; kc < K / 16; kc++) {
      uint64_t adesc = smem_desc(smem_u32(sA) + (2 * kc) * rowsA * 16 + shift * 16, (uint32_t)rowsA * 16, 128);
      uint64_t bdesc = smem_desc(smem_u32(sB) + (2 * kc) * N * 16, (uint32_t)N * 16, 128);
      if(!ws) umma_bf16(tmemBase, adesc, bdesc, idesc, kc > 0 ? 1u : 0u);
      else {
        // weight-stationary pair: D0 = A[shift..] B^T (fills collector b0), D1 = A[shift+1..] B^T (re-uses it)
        uint64_t adesc1 = smem_desc(smem_u32(sA) + (2 * kc) * rowsA * 16 + (shift + 1) * 16, (uint32_t)rowsA * 16, 128);
        if(kc & 1) {
          umma_bf16_ws<1, 0>(tmemBase, adesc, bdesc, idesc, kc > 0 ? 1u : 0u);
          umma_bf16_ws<1, 2>(tmemBase + N, adesc1, bdesc, idesc, kc > 0 ? 1u : 0u);
        } else {
          umma_bf16_ws<0, 0>(tmemBase, adesc, bdesc, idesc, kc > 0 ? 1u : 0u);
          umma_bf16_ws<0, 2>(tmemBase + N, adesc1, bdesc, idesc, kc > 0 ? 1u : 0u);
        }
      }
    }
    umma_commit(bar);
  }
  bool ok = mbar_wait(bar, 0, abortFlag, 41);
  tc_fence_after();
  if(ok && ws < 10) {
    int row = warp * 32 + lane;
    for(int half = 0; half < (ws ? 2 : 1); half++)
      for(int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmemBase + ((uint32_t)(warp * 32) << 16) + half * N + c0, v);
        for(int j = 0; j < 16; j++) D[((size_t)half * 128 + row) * N + c0 + j] = v[j];
      }
  }
  tc_fence_before();
  __syncthreads();
  if(warp == 0) { tc_fence_after(); tmem_dealloc(tmemBase, 256); }
}

// ------------------------------------------------------------------------------------------------
// host: program construction
// ------------------------------------------------------------------------------------------------
namespace {

struct Packer {
  std::vector<uint8_t> w;
  std::vector<float> p;
  // append the stages of one conv layer. rows(n, cin, tap) returns the weight.
  int fmt = 0;   // operand format the stream is packed in: 0 fp16 (saturating), 1 bf16
  template <class F>
  unsigned addConv(int N, int cinPadded, int ntaps, F weightAt) {
    unsigned off = (unsigned)w.size();
    int chunks = cinPadded / 16;
    for(int cc = 0; cc < chunks; cc++)
      for(int tap = 0; tap < ntaps; tap++) {
        size_t base = w.size();
        w.resize(base + (size_t)N * 32);
        uint16_t* dst = reinterpret_cast<uint16_t*>(w.data() + base);
        for(int h = 0; h < 2; h++)
          for(int n = 0; n < N; n++)
            for(int j = 0; j < 8; j++) {
              const float v = weightAt(n, cc * 16 + h * 8 + j, tap);
              uint16_t bits;
              if(fmt == 1) { __nv_bfloat16 q = __float2bfloat16(v); memcpy(&bits, &q, 2); }
              else { __half q = __float2half_rn(std::min(std::max(v, -65504.0f), 65504.0f)); memcpy(&bits, &q, 2); }
              dst[((size_t)h * N + n) * 8 + j] = bits;
            }
      }
    return off;
  }
  int addParams(const std::vector<float>& v) {
    int off = (int)p.size();
    p.insert(p.end(), v.begin(), v.end());
    while(p.size() % 4) p.push_back(0.f);   // keep every block float4-aligned
    return off;
  }
};

float convW(const ConvW& c, int oc, int ic, int tap) {   // oc,ic,y,x ; tap = y*kx + x
  if(oc >= c.oc || ic >= c.ic) return 0.f;
  return c.h[((size_t)oc * c.ic + ic) * c.ky * c.kx + tap];
}
std::vector<float> cat(std::initializer_list<const std::vector<float>*> parts) {
  std::vector<float> r;
  for(auto* v : parts) r.insert(r.end(), v->begin(), v->end());
  return r;
}

}  // namespace

int buildTrunkProgram(kc_model* m) {
  auto unsupported = [&](const std::string& why) { m->trunk = nullptr; m->trunkUnsupportedWhy = why; return 0; };
  const int C = m->trunkC;
  if(C % 16 != 0 || C > Cfg256::MAXC) return unsupported("trunk channels must be a multiple of 16 and <= 256 for the tcgen05 kernel");
  // <= 128 channels: two activation tiles per CTA; up to 192 (b15c192) and up to 256 (b20c256): one tile per CTA (TMEM budget)
  const int cfg = C <= Cfg128::MAXC ? 0 : C <= Cfg192::MAXC ? 1 : 2;
  const int maxC = cfg == 0 ? Cfg128::MAXC : cfg == 1 ? Cfg192::MAXC : Cfg256::MAXC, maxG = cfg == 0 ? Cfg128::MAXG : cfg == 1 ? Cfg192::MAXG : Cfg256::MAXG;
  if(m->initialConv.ky != 3 || m->initialConv.kx != 3) return unsupported("initial conv must be 3x3");
  const int HC = m->p1Conv.oc;   // shadows nothing: the packing below is in units of the model's head width
  if(m->g1Conv.oc != HC || m->v1Conv.oc != HC) return unsupported("the three head convolutions must have the same number of channels");
  if(HC != 32 && !(cfg == 2 && (HC == 48 || HC == 64)))
    return unsupported("head convolutions must have 32 channels (48 or 64 with a trunk of 193..256 channels: the b20c256 / b40c256 shapes)");
  if(m->gpoolToBiasMul.oc != HC || m->p2Conv.ic != HC || m->vOwnershipConv.ic != HC) return unsupported("head shapes do not fit together");
  if(m->p1Conv.ky != 1 || m->g1Conv.ky != 1 || m->v1Conv.ky != 1 || m->p2Conv.ky != 1 || m->vOwnershipConv.ky != 1)
    return unsupported("head convs must be 1x1");
  if(m->v2Mul.oc > MAX_V2 || m->v2Mul.oc % 16 != 0) return unsupported("v2 size must be a multiple of 16, <= 128");
  for(const BlockW& b : m->blocks) {
    if(b.regularConv.ky != 3 || b.regularConv.kx != 3 || b.finalConv.ky != 3 || b.finalConv.kx != 3) return unsupported("block convs must be 3x3");
    if(b.kind == 0 && (b.regularConv.oc % 16 != 0 || b.regularConv.oc > maxC)) return unsupported("mid channels must be a multiple of 16 and within the trunk width class");
    if(b.kind == 2) {
      if(b.gpoolConv.ky != 3 || b.gpoolConv.oc % 16 != 0 || b.gpoolConv.oc > maxG) return unsupported("gpool conv must be 3x3 with a multiple of 16 channels, <= 32 (trunk <= 128), <= 64 (trunk <= 192) or <= 80 (trunk <= 256)");
      if(b.regularConv.oc % 16 != 0 || b.regularConv.oc + b.gpoolConv.oc > maxC) return unsupported("gpool block: regular + gpool channels must fit the trunk width class");
    }
  }
  if(2 * m->blocks.size() + 2 > (size_t)MAX_LAYERS) return unsupported("too many blocks for the tcgen05 kernel's layer table");
  TrunkProgram* T = new TrunkProgram();
  T->cfg = cfg;
  T->headC = HC;
  double macs = 0;
  // the layer table and the fp32 parameters do not depend on the operand format; the weight stream is packed once per format
  auto packAll = [&](Packer& pk) {
  T->layers.clear();
  macs = 0;
  auto bnOf = [&](size_t blockIdx) -> const BNW& { return blockIdx < m->blocks.size() ? m->blocks[blockIdx].preBN : m->trunkTipBN; };
  {
    // layer 0: initial 3x3 conv over 15 planes + the global feature as a 16th input channel whose
    // weight sits on the centre tap only (== initialMatMul bias added to every cell, eigenbackend.cpp:1218-1220)
    LayerDesc L{};
    L.nk = 9; L.ntaps = 9; L.N = C; L.outSel = 0; L.accumulate = 0; L.epi = EPI_BN; L.epiC = C;
    L.wOffset = pk.addConv(C, 16, 9, [&](int n, int ic, int tap) {
      if(ic < 15) return convW(m->initialConv, n, ic, tap);
      return tap == 4 ? m->initialMatMul.h[n] : 0.f;
    });
    const BNW& bn = bnOf(0);
    L.act = bn.act;
    L.pOff = pk.addParams(cat({&bn.scale, &bn.bias}));
    T->layers.push_back(L);
    macs += 9.0 * 15 * C + C;
  }
  for(size_t bi = 0; bi < m->blocks.size(); bi++) {
    const BlockW& b = m->blocks[bi];
    const int R = b.regularConv.oc, G = b.kind == 2 ? b.gpoolConv.oc : 0;
    LayerDesc L1{};
    L1.nk = 9 * (C / 16); L1.ntaps = 9; L1.N = R + G; L1.outSel = 1; L1.accumulate = 0;
    L1.epi = b.kind == 2 ? EPI_GPOOL : EPI_BN; L1.epiC = R; L1.gpoolC = G;
    L1.wOffset = pk.addConv(R + G, C, 9, [&](int n, int ic, int tap) {
      return n < R ? convW(b.regularConv, n, ic, tap) : convW(b.gpoolConv, n - R, ic, tap);
    });
    L1.act = b.midBN.act; L1.gpoolAct = b.kind == 2 ? b.gpoolBN.act : 0;
    if(b.kind == 2) L1.pOff = pk.addParams(cat({&b.gpoolBN.scale, &b.gpoolBN.bias, &b.gpoolToBias.h, &b.midBN.scale, &b.midBN.bias}));
    else L1.pOff = pk.addParams(cat({&b.midBN.scale, &b.midBN.bias}));
    T->layers.push_back(L1);
    LayerDesc L2{};
    L2.nk = 9 * (R / 16); L2.ntaps = 9; L2.N = C; L2.outSel = 0; L2.accumulate = 1; L2.epi = EPI_BN; L2.epiC = C;
    L2.wOffset = pk.addConv(C, R, 9, [&](int n, int ic, int tap) { return convW(b.finalConv, n, ic, tap); });
    const BNW& bn = bnOf(bi + 1);
    L2.act = bn.act;
    L2.pOff = pk.addParams(cat({&bn.scale, &bn.bias}));
    T->layers.push_back(L2);
    macs += 9.0 * C * (R + G) + 9.0 * R * C + 3.0 * G * R;
  }
  {
    LayerDesc L{};
    L.nk = C / 16; L.ntaps = 1; L.N = 3 * HC; L.outSel = 1; L.accumulate = 0; L.epi = EPI_HEAD; L.epiC = 0;
    L.wOffset = pk.addConv(3 * HC, C, 1, [&](int n, int ic, int) {
      if(n < HC) return convW(m->p1Conv, n, ic, 0);
      if(n < 2 * HC) return convW(m->g1Conv, n - HC, ic, 0);
      return convW(m->v1Conv, n - 2 * HC, ic, 0);
    });
    // p2Conv oc,ic,1,1 -> [ic][4]
    std::vector<float> w2((size_t)HC * 4), wown(HC);
    for(int k = 0; k < HC; k++) {
      for(int d = 0; d < 4; d++) w2[(size_t)k * 4 + d] = m->p2Conv.h[(size_t)d * HC + k];
      wown[k] = m->vOwnershipConv.h[k];
    }
    L.pOff = pk.addParams(cat({&m->g1BN.scale, &m->g1BN.bias, &m->gpoolToBiasMul.h, &m->p1BN.scale, &m->p1BN.bias, &w2, &m->v1BN.scale,
                               &m->v1BN.bias, &m->v2Mul.h, &m->v2Bias.h, &m->v3Mul.h, &m->v3Bias.h, &m->sv3Mul.h, &m->sv3Bias.h, &wown}));
    T->layers.push_back(L);
    macs += 3.0 * HC * C;
  }
  };
  Packer pk, pkB;
  pk.fmt = 0; pkB.fmt = 1;
  packAll(pkB);
  packAll(pk);
  // Note: cat() above relies on each sub-block keeping the exact sizes the kernel indexes with
  // (HC, 3*HC*HC, ...); alignment padding is only appended after the whole block.
  T->v2C = m->v2Mul.oc;
  T->wBytes = pk.w.size();
  T->flopsPerEval = 2.0 * macs;   // per board cell; multiplied by H*W by the caller
  if(cudaMalloc(&T->d_params, pk.p.size() * 4) != cudaSuccess ||
     cudaMalloc(&T->d_layers, T->layers.size() * sizeof(LayerDesc)) != cudaSuccess) {
    delete T;
    return kc::fail("buildTrunkProgram: out of device memory");
  }
  for(int fmt = 0; fmt < 2; fmt++) {
    const std::vector<uint8_t>& ws = fmt == 0 ? pk.w : pkB.w;
    if(cudaMalloc(&T->d_w[fmt], ws.size()) != cudaSuccess) { delete T; return kc::fail("buildTrunkProgram: out of device memory"); }
    cudaMemcpy(T->d_w[fmt], ws.data(), ws.size(), cudaMemcpyHostToDevice);
    // pair-mode stream: per stage [rank 0: rows 0..N/2 of every K-step][rank 1: rows N/2..N], each K-step still [2][rows][8]
    std::vector<uint8_t> w2(ws.size());
    for(const LayerDesc& L : T->layers) {
      const int N = L.N, half = N / 2;
      const uint8_t* src = ws.data() + L.wOffset;
      uint8_t* dst = w2.data() + L.wOffset;
      static_assert(Cfg128P::KSTEPS == Cfg192P::KSTEPS, "one pair-mode stream layout");
      for(int k0 = 0; k0 < L.nk; k0 += Cfg128P::KSTEPS) {
        const int ks = std::min(Cfg128P::KSTEPS, L.nk - k0);
        for(int h = 0; h < 2; h++)
          for(int k = 0; k < ks; k++)
            for(int kc = 0; kc < 2; kc++) {
              memcpy(dst, src + (size_t)(k0 + k) * N * 32 + (size_t)kc * N * 16 + (size_t)h * half * 16, (size_t)half * 16);
              dst += (size_t)half * 16;
            }
      }
    }
    if(cudaMalloc(&T->d_wPair[fmt], w2.size()) != cudaSuccess) { delete T; return kc::fail("buildTrunkProgram: out of device memory"); }
    cudaMemcpy(T->d_wPair[fmt], w2.data(), w2.size(), cudaMemcpyHostToDevice);
  }
  cudaMemcpy(T->d_params, pk.p.data(), pk.p.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(T->d_layers, T->layers.data(), T->layers.size() * sizeof(LayerDesc), cudaMemcpyHostToDevice);
  m->trunk = T;
  return 0;
}

void freeTrunkProgram(kc_model* m) {
  if(!m->trunk) return;
  for(int fmt = 0; fmt < 2; fmt++) { cudaFree(m->trunk->d_w[fmt]); cudaFree(m->trunk->d_wPair[fmt]); }
  cudaFree(m->trunk->d_params); cudaFree(m->trunk->d_layers);
  delete m->trunk;
  m->trunk = nullptr;
}

int allocTrunkBuffers(kc_handle* h) {
  const int NB = boardsPerTile(h->W, h->H);
  int numTiles = (h->maxBatch + NB - 1) / NB;
  numTiles = (numTiles + 1) & ~1;   // items are tile pairs
  // round up to the feature kernel's CTA granularity (32 tiles) so partially filled CTAs stay in bounds
  numTiles = (numTiles + 31) / 32 * 32;
  h->numTilesAlloc = numTiles;
  size_t bytes = (size_t)numTiles * 2 * TILE_ROWS * 16;
  KC_CUDA(cudaMalloc(&h->d_tiles, bytes));
  KC_CUDA(cudaMemset(h->d_tiles, 0, bytes));
  KC_CUDA(cudaMalloc(&h->d_abort, 4));
  KC_CUDA(cudaMemset(h->d_abort, 0, 4));
  if(getenv("KC_TRUNK_PROBE")) { KC_CUDA(cudaMalloc(&h->d_dbg, KC_TRUNK_PROBE_WORDS * 8)); KC_CUDA(cudaMemset(h->d_dbg, 0, KC_TRUNK_PROBE_WORDS * 8)); }
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg128::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg192>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg192::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg192P>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg192P::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg256>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg256::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg256H48>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg256H48::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg256H64>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg256H64::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg128P>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg128P::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg128PR>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg128PR::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg128PProbe>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg128PProbe::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg192PProbe>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg192PProbe::SMEM));
  return 0;
}
void freeTrunkBuffers(kc_handle* h) {
  cudaFree(h->d_abort); cudaFree(h->d_dbg);
  for(cudaEvent_t e : h->evPool) cudaEventDestroy(e);
  h->evPool.clear();
}
int checkTrunkAbort(kc_handle* h) {
  if(!h->bf16 || !h->d_abort) return 0;
  int code = 0;
  KC_CUDA(cudaMemcpy(&code, h->d_abort, 4, cudaMemcpyDeviceToHost));
  KC_CHECK(code == 0, "trunk kernel aborted: an mbarrier wait timed out (code " + std::to_string(code) + ")");
  return 0;
}

int convertInputToTiles(kc_handle* h, int n, int rawNHWC, const int8_t* sym_dev, cudaStream_t st, int rowOffset) {
  const int NB = boardsPerTile(h->W, h->H), HW = h->W * h->H;
  int numTiles = ((n + NB - 1) / NB + 1) & ~1;
  uint4* tiles = (uint4*)h->d_tiles + (size_t)(rowOffset / NB) * 2 * TILE_ROWS;
  k_convert_tiles<<<(numTiles * 256 + 255) / 256, 256, 0, st>>>(h->d_raw + (size_t)rowOffset * 15 * HW, h->d_rawGlobal + rowOffset,
                                                               sym_dev ? sym_dev + rowOffset : nullptr, h->d_dstOfSrc, tiles, n, numTiles,
                                                               NB, h->W, h->H, rawNHWC, (h->flags & KC_FLAG_SYM_PERMUTE_DIRS) ? 1 : 0,
                                                               (h->flags & KC_FLAG_OPERANDS_BF16) ? 0 : 1);
  h->launches++;
  KC_CUDA(cudaGetLastError());
  return 0;
}

void handleTileConstants(const kc_handle* h, float k, uint32_t* one, uint32_t* kBits) {
  uint16_t a, b;
  if(h->flags & KC_FLAG_OPERANDS_BF16) { __nv_bfloat16 x = __float2bfloat16(1.0f), y = __float2bfloat16(k); memcpy(&a, &x, 2); memcpy(&b, &y, 2); }
  else { __half x = __float2half_rn(1.0f), y = __float2half_rn(k); memcpy(&a, &x, 2); memcpy(&b, &y, 2); }
  *one = a; *kBits = b;
}

int handleTilesPerItem(const kc_handle* h) { return h->model->trunk && h->model->trunk->cfg != 0 ? Cfg192::NT : Cfg128::NT; }   // (Cfg256::NT == Cfg192::NT)

static bool trunkUsesPairs() { static const bool usePair = [] { const char* e = getenv("KC_TRUNK_PAIR"); return !e || atoi(e) != 0; }(); return usePair; }
// true if this handle's trunk kernel has a variant that leaves registers to co-resident kernels (pair mode, trunks up to 128 channels)
bool handleCanLeaveRegisters(const kc_handle* h) { return h->bf16 && h->model->trunk && h->model->trunk->cfg == 0 && trunkUsesPairs(); }

int runTrunkBf16(kc_handle* h, int n, cudaStream_t st, const int8_t* sym_dev, int rowOffset, const int* nDev, bool symIsLocal) {
  const kc_model* m = h->model;
  const TrunkProgram* T = m->trunk;
  TrunkParams P{};
  P.opFmt = (h->flags & KC_FLAG_OPERANDS_BF16) ? 1 : 0;
  P.wstream = T->d_w[P.opFmt]; P.params = T->d_params;
  P.numLayers = (int)T->layers.size();
  for(int i = 0; i < P.numLayers; i++) P.layers[i] = T->layers[i];
  P.NB = boardsPerTile(h->W, h->H); P.W = h->W; P.H = h->H; P.HW = h->W * h->H; P.stride = h->W + 1; P.tileRowW = P.NB * P.stride; P.wMagic = 65536 / h->W + 1;
  int numTiles = (n + P.NB - 1) / P.NB;
  const int NT = T->cfg == 0 ? Cfg128::NT : Cfg192::NT;
  P.numItems = (numTiles + NT - 1) / NT;
  P.n = n;
  P.nDev = nDev;
  P.tiles = (const uint4*)h->d_tiles + (size_t)(rowOffset / P.NB) * 2 * TILE_ROWS;
  P.sym = sym_dev ? sym_dev + (symIsLocal ? 0 : rowOffset) : nullptr; P.dstOfSrcRev = h->d_dstOfSrcRev;
  P.policy = h->d_policy + (size_t)rowOffset * 4 * P.HW; P.value = h->d_value + (size_t)rowOffset * 2;
  P.misc = h->d_misc + (size_t)rowOffset * 2; P.own = h->d_own + (size_t)rowOffset * P.HW;
  P.abortFlag = h->d_abort;
  P.permuteDirs = (h->flags & KC_FLAG_SYM_PERMUTE_DIRS) ? 1 : 0;
  P.masked = (h->flags & KC_FLAG_MASKED_BOARDS) ? 1 : 0;
  P.g1Act = m->g1BN.act; P.p1Act = m->p1BN.act; P.v1Act = m->v1BN.act; P.v2Act = m->v2Act;
  P.dbg = h->d_dbg;
  float sq = sqrtf((float)P.HW);
  P.poolScale1 = (sq - 14.0f) * 0.1f;
  P.poolScale2 = (sq - 14.0f) * (sq - 14.0f) * 0.01f - 0.1f;
  P.invHW = 1.0f / (float)P.HW;
  P.v2C = T->v2C;
  int grid = std::min(P.numItems, h->ctx->smCount);
  if((int)h->evPool.size() < h->evUsed + 2 && h->evPool.size() < 4096) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    h->evPool.push_back(a); h->evPool.push_back(b);
  }
  bool timed = (int)h->evPool.size() >= h->evUsed + 2;
  if(timed) cudaEventRecord(h->evPool[h->evUsed], st);
  // CTA pairs (cta_group::2) are the default for trunks up to 128 channels; KC_TRUNK_PAIR=0 selects the single-CTA kernel
  const bool usePair = trunkUsesPairs() && T->cfg != 2;   // the 256-channel configuration has no pair form
  // tile skew (see TrunkParams::skew): 5 of the 14 ring stages in pair mode, measured +1.5 % burst / +1 % under the power cap
  // (0: 7.165, 2: 7.195, 4: 7.22, 5-8: 7.27-7.285, 10: 7.26 M evals/s); 2 of 7 in the single-CTA kernel; KC_TRUNK_SKEW overrides
  static const int skewEnv = [] { const char* e = getenv("KC_TRUNK_SKEW"); return e ? atoi(e) : -1; }();
  static const int deferEnv = [] { const char* e = getenv("KC_TRUNK_DEFER_HEAD"); return e ? atoi(e) : 1; }();
  P.deferHead = deferEnv;
  P.skew = T->cfg != 0 ? 0 : skewEnv >= 0 ? std::min(skewEnv, usePair ? Cfg128P::NSTAGES - 2 : Cfg128::NSTAGES - 2) : usePair ? 2 : 2;
  if(usePair) {
    // clusters of two CTAs (one TPC), cta_group::2 MMAs; a cluster takes two items per round
    P.wstream = T->d_wPair[P.opFmt];
    const int numUnits = nDev ? h->ctx->smCount / 2 : (P.numItems + 1) / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * std::min(numUnits, h->ctx->smCount / 2)); cfg.blockDim = dim3(T->cfg == 0 ? Cfg128P::THREADS : Cfg192P::THREADS);
    cfg.dynamicSmemBytes = T->cfg == 0 ? Cfg128P::SMEM : Cfg192P::SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    // the re-allocating variant only where something wants to run beside the trunk (the search's half batches); KC_TRUNK_REALLOC forces
    static const int reallocEnv = [] { const char* e = getenv("KC_TRUNK_REALLOC"); return e ? atoi(e) : -1; }();
    if(T->cfg != 0) KC_CUDA(cudaLaunchKernelEx(&cfg, P.dbg ? trunk_kernel<Cfg192PProbe> : trunk_kernel<Cfg192P>, P));
    else if(P.dbg) KC_CUDA(cudaLaunchKernelEx(&cfg, trunk_kernel<Cfg128PProbe>, P));
    else if(reallocEnv >= 0 ? reallocEnv != 0 : h->leaveRegisters) KC_CUDA(cudaLaunchKernelEx(&cfg, trunk_kernel<Cfg128PR>, P));
    else KC_CUDA(cudaLaunchKernelEx(&cfg, trunk_kernel<Cfg128P>, P));
  }
  else if(T->cfg == 0) trunk_kernel<Cfg128><<<grid, Cfg128::THREADS, Cfg128::SMEM, st>>>(P);
  else if(T->cfg == 1) trunk_kernel<Cfg192><<<grid, Cfg192::THREADS, Cfg192::SMEM, st>>>(P);
  else if(T->headC == 48) trunk_kernel<Cfg256H48><<<grid, Cfg256H48::THREADS, Cfg256H48::SMEM, st>>>(P);
  else if(T->headC == 64) trunk_kernel<Cfg256H64><<<grid, Cfg256H64::THREADS, Cfg256H64::SMEM, st>>>(P);
  else trunk_kernel<Cfg256><<<grid, Cfg256::THREADS, Cfg256::SMEM, st>>>(P);
  if(timed) { cudaEventRecord(h->evPool[h->evUsed + 1], st); h->evUsed += 2; }
  h->launches++;
  KC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace kc

extern "C" {
int kc_handle_trunk_probe(kc_handle* h, int64_t* out) {
  KC_CHECK(h && out, "kc_handle_trunk_probe: null argument");
  KC_CHECK(h->d_dbg, "kc_handle_trunk_probe: set KC_TRUNK_PROBE=1 before creating the handle");
  KC_CUDA(cudaSetDevice(h->ctx->device));
  KC_CUDA(cudaDeviceSynchronize());
  KC_CUDA(cudaMemcpy(out, h->d_dbg, kc::KC_TRUNK_PROBE_WORDS * 8, cudaMemcpyDeviceToHost));
  return 0;
}
// Self-test of the UMMA descriptor conventions the trunk kernel relies on (row-shifted K-major
// no-swizzle A operand): computes D = A[shift:shift+128] * B^T on the tensor core and returns it.
// A [rowsA][K], B [N][K] as bf16 bit patterns (uint16), D [128][N] fp32. K % 16 == 0, N % 16 == 0.
int kc_selftest_umma(kc_ctx* ctx, const uint16_t* A, const uint16_t* B, float* D, int rowsA, int N, int K, int shift, int ws) {
  using namespace kc;
  KC_CHECK(ctx && A && B && D, "kc_selftest_umma: null argument");
  KC_CHECK(K % 16 == 0 && N % 16 == 0 && N <= 256 && shift >= 0 && shift + 128 + (ws ? 1 : 0) <= rowsA, "kc_selftest_umma: bad shape");
  KC_CHECK(ws == 0 || ws >= 10 || N == 64 || N == 128, "kc_selftest_umma: the weight-stationary form needs N in {64, 128} here (two accumulators)");
  KC_CHECK(ws < 10 || N <= 128, "kc_selftest_umma: throughput probe uses two accumulators");
  const int nOut = ws ? 2 : 1;
  size_t smemBytes = (size_t)(K / 8) * (rowsA + N) * 16;
  KC_CHECK(smemBytes <= 200 * 1024, "kc_selftest_umma: operands do not fit shared memory");
  KC_CUDA(cudaSetDevice(ctx->device));
  __nv_bfloat16 *dA, *dB; float* dD; int* dAbort;
  KC_CUDA(cudaMalloc(&dA, (size_t)rowsA * K * 2)); KC_CUDA(cudaMalloc(&dB, (size_t)N * K * 2));
  KC_CUDA(cudaMalloc(&dD, (size_t)nOut * 128 * N * 4)); KC_CUDA(cudaMalloc(&dAbort, 4));
  KC_CUDA(cudaMemcpy(dA, A, (size_t)rowsA * K * 2, cudaMemcpyHostToDevice));
  KC_CUDA(cudaMemcpy(dB, B, (size_t)N * K * 2, cudaMemcpyHostToDevice));
  KC_CUDA(cudaMemset(dD, 0, (size_t)nOut * 128 * N * 4)); KC_CUDA(cudaMemset(dAbort, 0, 4));
  KC_CUDA(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemBytes));
  umma_probe_kernel<<<1, 128, smemBytes>>>(dA, dB, dD, rowsA, N, K, shift, ws, dAbort);
  KC_CUDA(cudaGetLastError());
  KC_CUDA(cudaDeviceSynchronize());
  int abortCode = 0;
  KC_CUDA(cudaMemcpy(&abortCode, dAbort, 4, cudaMemcpyDeviceToHost));
  KC_CUDA(cudaMemcpy(D, dD, (size_t)nOut * 128 * N * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dAbort);
  KC_CHECK(abortCode == 0, "kc_selftest_umma: kernel timed out waiting on an mbarrier (code " + std::to_string(abortCode) + ")");
  return 0;
}
}
