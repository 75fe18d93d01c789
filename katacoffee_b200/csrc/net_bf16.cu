// The bf16 production path: the whole residual net as ONE persistent tcgen05/TMEM kernel.
//
// What it computes is the reference's Model::apply (cpp/neuralnet/eigenbackend.cpp:1420-1472:
// trunk :1202-1226, blocks :912-930 / :968-1005, policy head :1265-1298, value head :1341-1376) for
// boards that exactly fill nnXLen x nnYLen (mask == 1, ledger 8.1-M), with bf16 operands and fp32
// accumulation.  How it computes it is B200-native:
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
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "handle.h"
#include "umma.cuh"

namespace kc {

using namespace ptx;

constexpr int CHUNK_BYTES = ACT_ROWS * 16;             // one 8-channel chunk of an activation tile
constexpr int KSTEPS_PER_STAGE = 3;
constexpr int SCR_STRIDE = 17;
constexpr int MAX_NB = 4;
constexpr int HEADC = 32;   // p1 = g1 = v1 = 32 channels
constexpr int MAX_V2 = 128;

// Compile-time shape of one trunk-kernel instantiation.  <128, 2, 7>: trunks up to 128 channels, two activation tiles
// per CTA (TMEM: 2 x (T 128 + S 128) columns).  <192, 1, 6>: trunks up to 192 channels (b15c192), one tile per CTA
// (TMEM: T 192 + S 192 columns of the 512 allocated).
template <int MAXC_, int NT_, int NSTAGES_>
struct TrunkCfg {
  static constexpr int MAXC = MAXC_, NT = NT_, NSTAGES = NSTAGES_;
  static constexpr int NCH = MAXC / 16;                       // 16-channel chunks the epilogue publishes
  static constexpr int MAXG = MAXC / 4 < 32 ? 32 : MAXC / 3;  // gpool channels: 32 (c128), 64 (c192)
  static constexpr int POOLW = 3 * MAXG;                      // pooled vector per board (>= 96 for the heads)
  static constexpr int THREADS = 128 + NT * 128;              // warp 0 TMA producer, 1..NT MMA issuers, 4.. epilogue (4 warps per tile)
  static constexpr int ACT_BYTES = (MAXC / 8) * CHUNK_BYTES;
  static constexpr int STAGE_BYTES = KSTEPS_PER_STAGE * MAXC * 32;
  static constexpr int OFF_ACT = 0;
  static constexpr int OFF_RING = OFF_ACT + NT * ACT_BYTES;
  static constexpr int OFF_SCR = OFF_RING + NSTAGES * STAGE_BYTES;
  static constexpr int OFF_POOLA = OFF_SCR + NT * 128 * SCR_STRIDE * 4;
  static constexpr int OFF_POOLB = OFF_POOLA + NT * MAX_NB * POOLW * 4;
  static constexpr int OFF_BIAS = OFF_POOLB + NT * MAX_NB * POOLW * 4;
  static constexpr int OFF_V2 = OFF_BIAS + NT * MAX_NB * MAXC * 4;
  static constexpr int OFF_SYM = OFF_V2 + NT * MAX_NB * MAX_V2 * 4;
  static constexpr int OFF_PAR = OFF_SYM + 448;                   // [tile][2 buffers][scale MAXC | bias MAXC] fp32: the next layer's
  static constexpr int OFF_BAR = OFF_PAR + NT * 2 * 2 * MAXC * 4; // folded BN, staged while the tensor core is still busy with it
  // barriers (8 bytes each)
  static constexpr int BAR_FULL = 0, BAR_EMPTY = BAR_FULL + NSTAGES, BAR_ACC = BAR_EMPTY + NSTAGES, BAR_ACTFREE = BAR_ACC + NT,
                       BAR_IN = BAR_ACTFREE + NT, BAR_HEAD = BAR_IN + NT, BAR_CHUNK = BAR_HEAD + NT, NUM_BARS = BAR_CHUNK + NT * NCH;
  static constexpr int OFF_TMEM = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM = OFF_TMEM + 16;
  static_assert(NT * 2 * MAXC <= 512, "TMEM: every tile needs a trunk region and a block-internal region");
  static_assert(SMEM <= 232448, "shared memory budget");
  static_assert(POOLW >= 96, "the heads pool 32 channels three ways");
};
using Cfg128 = TrunkCfg<128, 2, 7>;
using Cfg192 = TrunkCfg<192, 1, 6>;

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
  unsigned wOffset; // byte offset of this layer's first stage in the weight stream
  int pOff;        // float offset of this layer's parameters
};
constexpr int MAX_LAYERS = 48;   // 1 + 2*blocks + 1; the table travels in the kernel parameter block (constant bank:
                                 // warp-uniform loads, so the MMA issue loop stays on the uniform datapath)

struct TrunkProgram {
  std::vector<LayerDesc> layers;
  uint8_t* d_w = nullptr;
  float* d_params = nullptr;
  LayerDesc* d_layers = nullptr;
  size_t wBytes = 0;
  int v2C = 0;
  int cfg = 0;     // 0: TrunkCfg<128, 2, 7>, 1: TrunkCfg<192, 1, 6>
  double flopsPerEval = 0;
};

struct TrunkParams {
  const uint4* tiles; const uint8_t* wstream; const float* params;
  LayerDesc layers[MAX_LAYERS];
  int numLayers, numItems, n;
  const int* nDev;   // if non-null: the number of rows is read from device memory (a batch compacted on the device)
  int NB, W, H, HW, stride, tileRowW;
  const int8_t* sym; const uint8_t* dstOfSrcRev;
  float *policy, *value, *misc, *own;
  int* abortFlag;
  float poolScale1, poolScale2, invHW;
  int v2C;
};

// ------------------------------------------------------------------------------------------------
// device code
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

struct EpiCtx {
  int t;            // tile within the CTA
  int r;            // row = TMEM lane
  int e;            // thread index within the tile's epilogue group (== r)
  bool valid;       // row is a board cell
  int b, cell;      // board within tile, dense cell index
  uint32_t tmemLane;  // tmem base + lane offset + tile column offset
  uint8_t* act;     // this tile's activation buffer
  uint32_t barChunk;  // smem address of actReady[t][0]
  float* scr; float* poolA; float* poolB; float* biasBuf; float* v2buf;
  const float* par;  // staged folded BN of the layer being finished: scale[K::MAXC] | bias[K::MAXC]
};

// folded BN + ReLU + mask for 16 columns -> two 16-byte chunks of the activation tile
__device__ __forceinline__ void publish16(const EpiCtx& c, int cc, const float v[16], const float* scale, const float* bias, const float* add) {
  uint32_t pk[8];
#pragma unroll
  for(int q = 0; q < 4; q++) {
    float4 s = *(reinterpret_cast<const float4*>(scale + cc * 16) + q);     // shared memory (c.par), staged before the accumulator wait
    float4 bb = *(reinterpret_cast<const float4*>(bias + cc * 16) + q);
    float x0 = v[4 * q], x1 = v[4 * q + 1], x2 = v[4 * q + 2], x3 = v[4 * q + 3];
    if(add) { x0 += add[cc * 16 + 4 * q]; x1 += add[cc * 16 + 4 * q + 1]; x2 += add[cc * 16 + 4 * q + 2]; x3 += add[cc * 16 + 4 * q + 3]; }
    float a0 = fmaxf(fmaf(x0, s.x, bb.x), 0.f), a1 = fmaxf(fmaf(x1, s.y, bb.y), 0.f);
    float a2 = fmaxf(fmaf(x2, s.z, bb.z), 0.f), a3 = fmaxf(fmaf(x3, s.w, bb.w), 0.f);
    if(!c.valid) { a0 = a1 = a2 = a3 = 0.f; }
    pk[2 * q] = pack2(a0, a1);
    pk[2 * q + 1] = pack2(a2, a3);
  }
  uint8_t* dst = c.act + (size_t)(2 * cc) * CHUNK_BYTES + (size_t)(HALO_ROWS + c.r) * 16;
  *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  *reinterpret_cast<uint4*>(dst + CHUNK_BYTES) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
  fence_proxy_async();
  mbar_arrive(c.barChunk + cc * 8);
}

// per-board pooling of 16 channels held one row per thread: writes sum and max per (board, channel)
// into outSum/outMax[b*16 + j] (shared), using the tile's scratch.  All 128 threads of the tile call it.
__device__ __forceinline__ void poolBoards16(const TrunkParams& P, const EpiCtx& c, const float g[16], float* sums, float* maxs) {
#pragma unroll
  for(int j = 0; j < 16; j++) c.scr[c.r * SCR_STRIDE + j] = g[j];
  named_bar_sync(1 + c.t, 128);
  if(c.e < P.NB * 16) {
    int b = c.e >> 4, j = c.e & 15;
    float s = 0.f, m = -1.0f;   // eigenbackend.cpp:145: max starts at -1
    for(int y = 0; y < P.H; y++)
      for(int x = 0; x < P.W; x++) {
        float v = c.scr[(y * P.tileRowW + b * P.stride + x) * SCR_STRIDE + j];
        s += v;
        m = fmaxf(m, v);
      }
    sums[c.e] = s;
    maxs[c.e] = m;
  }
  named_bar_sync(1 + c.t, 128);
}

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
  for(int cc = 0; cc < nch; cc += 2) {
    float v[16];
    if(cc + 1 < nch) tmem_ld16_issue(src + (cc + 1) * 16, rb);
#pragma unroll
    for(int i = 0; i < 16; i++) v[i] = __uint_as_float(ra[i]);
    publish16(c, cc, v, scale, bias, nullptr);
    if(cc + 1 >= nch) break;
    tmem_ld16_wait(rb);
    if(cc + 2 < nch) tmem_ld16_issue(src + (cc + 2) * 16, ra);
#pragma unroll
    for(int i = 0; i < 16; i++) v[i] = __uint_as_float(rb[i]);
    publish16(c, cc + 1, v, scale, bias, nullptr);
    if(cc + 2 < nch) tmem_ld16_wait(ra);
  }
}

// params: gpoolBN scale[G] bias[G] | Wg [3G][R] | midBN scale[R] bias[R]
template <class K>
__device__ void epilogueGPool(const TrunkParams& P, const LayerDesc& L, const EpiCtx& c) {
  const int R = L.epiC, G = L.gpoolC;
  const float* gs = P.params + L.pOff;
  const float* gb = gs + G;
  const float* Wg = gb + G;
  const float* ms = c.par;            // midBN staged in shared memory
  const float* mb = c.par + K::MAXC;
  uint32_t src = c.tmemLane + K::MAXC;   // region S
  float* pooled = c.poolA;           // [NB][3G]
  __shared__ float sSum[2][MAX_NB * 16], sMax[2][MAX_NB * 16];
  for(int half = 0; half < G / 16; half++) {
    float v[16], g[16];
    tmem_ld16(src + R + half * 16, v);
#pragma unroll
    for(int j = 0; j < 16; j++) {
      float a = fmaxf(fmaf(v[j], __ldg(gs + half * 16 + j), __ldg(gb + half * 16 + j)), 0.f);
      g[j] = c.valid ? a : 0.f;
    }
    poolBoards16(P, c, g, sSum[c.t], sMax[c.t]);
    if(c.e < P.NB * 16) {
      int b = c.e >> 4, j = c.e & 15;
      float mean = sSum[c.t][c.e] * P.invHW;
      pooled[b * 3 * G + half * 16 + j] = mean;
      pooled[b * 3 * G + G + half * 16 + j] = mean * P.poolScale1;
      pooled[b * 3 * G + 2 * G + half * 16 + j] = sMax[c.t][c.e];
    }
  }
  named_bar_sync(1 + c.t, 128);
  for(int idx = c.e; idx < P.NB * R; idx += 128) {
    int b = idx / R, oc = idx - b * R;
    float acc = 0.f;
    for(int k = 0; k < 3 * G; k++) acc = fmaf(pooled[b * 3 * G + k], __ldg(Wg + k * R + oc), acc);
    c.biasBuf[b * K::MAXC + oc] = acc;
  }
  named_bar_sync(1 + c.t, 128);
  const float* add = c.biasBuf + c.b * K::MAXC;
  for(int cc = 0; cc < R / 16; cc++) {
    float v[16];
    tmem_ld16(src + cc * 16, v);
    publish16(c, cc, v, ms, mb, add);
  }
}

// params: g1BN s[32] b[32] | Wpb [96][32] | p1BN s[32] b[32] | W2 [32][4] | v1BN s[32] b[32] |
//         Wv2 [96][V2] | b2 [V2] | Wv3 [V2][2] | b3[2] | Wsv3 [V2][2] | bsv3[2] | Wown [32]
template <class K>
__device__ void epilogueHead(const TrunkParams& P, const LayerDesc& L, const EpiCtx& c, int tileIndex, uint32_t barHead,
                             const uint8_t* sSym, const int nRows) {
  const int V2 = P.v2C;
  const float* g1s = P.params + L.pOff;
  const float* g1b = g1s + HEADC;
  const float* Wpb = g1b + HEADC;
  const float* p1s = Wpb + 96 * HEADC;
  const float* p1b = p1s + HEADC;
  const float* W2 = p1b + HEADC;
  const float* v1s = W2 + HEADC * 4;
  const float* v1b = v1s + HEADC;
  const float* Wv2 = v1b + HEADC;
  const float* b2 = Wv2 + 96 * V2;
  const float* Wv3 = b2 + V2;
  const float* b3 = Wv3 + V2 * 2;
  const float* Wsv3 = b3 + 2;
  const float* bsv3 = Wsv3 + V2 * 2;
  const float* Wown = bsv3 + 2;
  uint32_t src = c.tmemLane + K::MAXC;
  __shared__ float sSum[2][MAX_NB * 16], sMax[2][MAX_NB * 16];
  float* pooledG = c.poolA;   // [NB][96]
  float* pooledV = c.poolB;   // [NB][96]
  float v1a[HEADC];
  // g1 -> BN -> ReLU -> gpool (eigenbackend.cpp:1290-1291) ; v1 -> BN -> ReLU -> value pool (:1364-1366)
  for(int half = 0; half < 2; half++) {
    float v[16], g[16];
    tmem_ld16(src + HEADC + half * 16, v);
#pragma unroll
    for(int j = 0; j < 16; j++) {
      float a = fmaxf(fmaf(v[j], __ldg(g1s + half * 16 + j), __ldg(g1b + half * 16 + j)), 0.f);
      g[j] = c.valid ? a : 0.f;
    }
    poolBoards16(P, c, g, sSum[c.t], sMax[c.t]);
    if(c.e < P.NB * 16) {
      int b = c.e >> 4, j = c.e & 15;
      float mean = sSum[c.t][c.e] * P.invHW;
      pooledG[b * 96 + half * 16 + j] = mean;
      pooledG[b * 96 + 32 + half * 16 + j] = mean * P.poolScale1;
      pooledG[b * 96 + 64 + half * 16 + j] = sMax[c.t][c.e];
    }
    tmem_ld16(src + 2 * HEADC + half * 16, v);
#pragma unroll
    for(int j = 0; j < 16; j++) {
      float a = fmaxf(fmaf(v[j], __ldg(v1s + half * 16 + j), __ldg(v1b + half * 16 + j)), 0.f);
      g[j] = c.valid ? a : 0.f;
      v1a[half * 16 + j] = g[j];
    }
    poolBoards16(P, c, g, sSum[c.t], sMax[c.t]);
    if(c.e < P.NB * 16) {
      int b = c.e >> 4, j = c.e & 15;
      float mean = sSum[c.t][c.e] * P.invHW;
      pooledV[b * 96 + half * 16 + j] = mean;
      pooledV[b * 96 + 32 + half * 16 + j] = mean * P.poolScale1;
      pooledV[b * 96 + 64 + half * 16 + j] = mean * P.poolScale2;
    }
  }
  float p1[HEADC];
  {
    float v[16];
    tmem_ld16(src, v);
#pragma unroll
    for(int j = 0; j < 16; j++) p1[j] = v[j];
    tmem_ld16(src + 16, v);
#pragma unroll
    for(int j = 0; j < 16; j++) p1[16 + j] = v[j];
  }
  // all TMEM reads of this tile are done: the MMA warp may overwrite region S for the next item
  tc_fence_before();
  mbar_arrive(barHead);
  named_bar_sync(1 + c.t, 128);
  // pooled matmuls: policy bias (NB*32 outputs) and v2 (NB*V2 outputs)
  if(c.e < P.NB * HEADC) {
    int b = c.e / HEADC, oc = c.e % HEADC;
    float acc = 0.f;
    for(int k = 0; k < 96; k++) acc = fmaf(pooledG[b * 96 + k], __ldg(Wpb + k * HEADC + oc), acc);
    c.biasBuf[b * 96 + oc] = acc;
  }
  for(int idx = c.e; idx < P.NB * V2; idx += 128) {
    int b = idx / V2, oc = idx - b * V2;
    float acc = 0.f;
    for(int k = 0; k < 96; k++) acc = fmaf(pooledV[b * 96 + k], __ldg(Wv2 + k * V2 + oc), acc);
    c.v2buf[b * MAX_V2 + oc] = fmaxf(acc + __ldg(b2 + oc), 0.f);
  }
  named_bar_sync(1 + c.t, 128);
  const int gameBase = tileIndex * P.NB;
  if(c.e < P.NB * 4) {
    int b = c.e >> 2, o = c.e & 3;
    int game = gameBase + b;
    if(game < nRows) {
      const float* Wm = (o < 2) ? Wv3 : Wsv3;
      int oo = o & 1;
      float acc = (o < 2) ? __ldg(b3 + oo) : __ldg(bsv3 + oo);
      for(int k = 0; k < V2; k++) acc = fmaf(c.v2buf[b * MAX_V2 + k], __ldg(Wm + k * 2 + oo), acc);
      if(o < 2) P.value[(size_t)game * 2 + oo] = acc; else P.misc[(size_t)game * 2 + oo] = acc;
    }
  }
  int game = gameBase + c.b;
  if(c.valid && game < nRows) {
    const float* add = c.biasBuf + c.b * 96;
    float o0 = 0.f, o1 = 0.f, o2 = 0.f, o3 = 0.f, own = 0.f;
#pragma unroll
    for(int k = 0; k < HEADC; k++) {
      float a = fmaxf(fmaf(p1[k] + add[k], __ldg(p1s + k), __ldg(p1b + k)), 0.f);
      float4 w = __ldg(reinterpret_cast<const float4*>(W2) + k);
      o0 = fmaf(a, w.x, o0); o1 = fmaf(a, w.y, o1); o2 = fmaf(a, w.z, o2); o3 = fmaf(a, w.w, o3);
      own = fmaf(v1a[k], __ldg(Wown + k), own);
    }
    int s = P.sym ? P.sym[game] : 0;
    int dst = sSym[s * P.HW + c.cell];
    float* pol = P.policy + (size_t)game * 4 * P.HW;
    pol[dst] = o0; pol[P.HW + dst] = o1; pol[2 * P.HW + dst] = o2; pol[3 * P.HW + dst] = o3;
    P.own[(size_t)game * P.HW + dst] = own;
  }
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
  auto desc = [descHi](uint32_t lo) { return ((uint64_t)descHi << 32) | lo; };
  for(int item = blockIdx.x; item < numItems; item += gridDim.x, itemCount++) {
    if(!mbar_wait(bars + (K::BAR_IN + t) * 8, itemCount & 1, abortFlag, 21)) return;
    if(itemCount > 0 && !mbar_wait(bars + (K::BAR_HEAD + t) * 8, (itemCount - 1) & 1, abortFlag, 22)) return;
    tc_fence_after();
    for(int l = 0; l < P.numLayers; l++) {
      const int nk = P.layers[l].nk, ntaps = P.layers[l].ntaps, N = P.layers[l].N;
      const uint32_t idesc = idesc_bf16_f32(128, N);
      const uint32_t d = tmemBase + t * (2 * K::MAXC) + (P.layers[l].outSel ? K::MAXC : 0);
      const uint32_t bStep = 2 * N;                    // one K-step of weights = N*32 bytes
      const uint32_t bLbo = (uint32_t)N << 16;         // LBO = N*16 bytes
      uint32_t accum = P.layers[l].accumulate ? 1u : 0u;
      if(ntaps == 9) {
        const int nchunks = nk / 9;
        for(int cc = 0; cc < nchunks; cc++) {
          if(l > 0) {
            const uint32_t bit = 1u << cc;
            if(!mbar_wait(barChunk + cc * 8, (chunkPhase & bit) ? 1 : 0, abortFlag, 24)) return;
            chunkPhase ^= bit;
          }
          const uint32_t aLoC = aLo0 + cc * (2 * CHUNK_BYTES >> 4);
#pragma unroll
          for(int dy = 0; dy < 3; dy++) {
            if(!mbar_wait(barFull + slot * 8, phase, abortFlag, 23)) return;
            tc_fence_after();
            const uint32_t bLo = (ringLo0 + slot * (K::STAGE_BYTES >> 4)) | bLbo;
            const uint32_t aLoR = aLoC + (dy - 1) * P.tileRowW - 1;
            if(leader) {
              umma_bf16(d, desc(aLoR), desc(bLo), idesc, accum);
              umma_bf16(d, desc(aLoR + 1), desc(bLo + bStep), idesc, 1u);
              umma_bf16(d, desc(aLoR + 2), desc(bLo + 2 * bStep), idesc, 1u);
              umma_commit(barEmpty + slot * 8);
            }
            __syncwarp();
            accum = 1u;
            if(++slot == K::NSTAGES) { slot = 0; phase ^= 1; }
          }
        }
      } else {
        // 1x1 layers: K-step = one 16-channel chunk, stages of up to 3 K-steps
        const int nst = (nk + KSTEPS_PER_STAGE - 1) / KSTEPS_PER_STAGE;
        for(int s = 0; s < nst; s++) {
          if(!mbar_wait(barFull + slot * 8, phase, abortFlag, 23)) return;
          const uint32_t bLo = (ringLo0 + slot * (K::STAGE_BYTES >> 4)) | bLbo;
          const int ks = min(KSTEPS_PER_STAGE, nk - s * KSTEPS_PER_STAGE);
          for(int kk = 0; kk < ks; kk++) {
            const int cc = s * KSTEPS_PER_STAGE + kk;
            if(l > 0) {
              const uint32_t bit = 1u << cc;
              if(!mbar_wait(barChunk + cc * 8, (chunkPhase & bit) ? 1 : 0, abortFlag, 24)) return;
              chunkPhase ^= bit;
            }
            tc_fence_after();
            if(leader) umma_bf16(d, desc(aLo0 + cc * (2 * CHUNK_BYTES >> 4)), desc(bLo + kk * bStep), idesc, accum);
            __syncwarp();
            accum = 1u;
          }
          if(leader) umma_commit(barEmpty + slot * 8);
          __syncwarp();
          if(++slot == K::NSTAGES) { slot = 0; phase ^= 1; }
        }
      }
      if(leader) {
        umma_commit(bars + (K::BAR_ACC + t) * 8);
        if(l == P.numLayers - 1) umma_commit(bars + (K::BAR_ACTFREE + t) * 8);
      }
      __syncwarp();
    }
  }
}

template <class K>
__global__ void __launch_bounds__(K::THREADS, 1) trunk_kernel(const TrunkParams P) {
  constexpr int NT = K::NT;
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + K::OFF_BAR;
  volatile int* abortFlag = P.abortFlag;
  uint8_t* sSym = smem + K::OFF_SYM;
  const int nRows = P.nDev ? min(__ldg(P.nDev), P.n) : P.n;
  const int numItems = P.nDev ? ((nRows + P.NB - 1) / P.NB + NT - 1) / NT : P.numItems;

  // ---- one-time setup ----
  for(int i = threadIdx.x; i < NT * K::ACT_BYTES / 16; i += K::THREADS) reinterpret_cast<uint4*>(smem + K::OFF_ACT)[i] = make_uint4(0, 0, 0, 0);
  for(int i = threadIdx.x; i < 8 * P.HW; i += K::THREADS) sSym[i] = P.dstOfSrcRev[i];
  if(threadIdx.x == 0) {
    for(int i = 0; i < K::NSTAGES; i++) { mbar_init(bars + (K::BAR_FULL + i) * 8, 1); mbar_init(bars + (K::BAR_EMPTY + i) * 8, NT); }   // every MMA issuer releases a slot
    for(int t = 0; t < NT; t++) {
      mbar_init(bars + (K::BAR_ACC + t) * 8, 1); mbar_init(bars + (K::BAR_ACTFREE + t) * 8, 1); mbar_init(bars + (K::BAR_IN + t) * 8, 1);
      mbar_init(bars + (K::BAR_HEAD + t) * 8, 128);
      for(int c = 0; c < K::NCH; c++) mbar_init(bars + (K::BAR_CHUNK + t * K::NCH + c) * 8, 128);
    }
    fence_mbar_init();
  }
  fence_proxy_async();
  if(warp == 1) { tmem_alloc(sbase + K::OFF_TMEM, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmemBase = *reinterpret_cast<volatile uint32_t*>(smem + K::OFF_TMEM);

  if(warp == 0) {
    // =========================== TMA producer ===========================
    if(lane == 0) {
      uint32_t slot = 0, phase = 0, itemCount = 0;
      bool alive = true;
      for(int item = blockIdx.x; item < numItems && alive; item += gridDim.x, itemCount++) {
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
          const int nst = (L.nk + KSTEPS_PER_STAGE - 1) / KSTEPS_PER_STAGE;
          const uint8_t* w = P.wstream + L.wOffset;
          for(int s = 0; s < nst; s++) {
            int ks = min(KSTEPS_PER_STAGE, L.nk - s * KSTEPS_PER_STAGE);
            uint32_t bytes = (uint32_t)ks * L.N * 32;
            alive = mbar_wait(bars + (K::BAR_EMPTY + slot) * 8, phase ^ 1, abortFlag, 12);
            if(!alive) break;
            uint32_t bar = bars + (K::BAR_FULL + slot) * 8;
            mbar_arrive_expect_tx(bar, bytes);
            bulk_g2s(sbase + K::OFF_RING + slot * K::STAGE_BYTES, w, bytes, bar);
            w += bytes;
            if(++slot == K::NSTAGES) { slot = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if(warp >= 1 && warp <= NT) {
    // =========================== MMA issuers: one thread per tile ===========================
    mmaIssuer<K>(P, warp - 1, sbase, bars, tmemBase, abortFlag, numItems);
  } else if(warp < 4) {
    // idle (keeps the epilogue warps aligned to TMEM lane quadrants: warp % 4 == quadrant)
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
    c.valid = (y < P.H) && (x < P.W);
    c.cell = y * P.W + x;
    if(!c.valid) { c.b = 0; c.cell = 0; }
    c.tmemLane = tmemBase + ((uint32_t)(q * 32) << 16) + c.t * (2 * K::MAXC);
    c.act = smem + K::OFF_ACT + c.t * K::ACT_BYTES;
    c.barChunk = bars + (K::BAR_CHUNK + c.t * K::NCH) * 8;
    c.scr = reinterpret_cast<float*>(smem + K::OFF_SCR) + c.t * 128 * SCR_STRIDE;
    c.poolA = reinterpret_cast<float*>(smem + K::OFF_POOLA) + c.t * MAX_NB * K::POOLW;
    c.poolB = reinterpret_cast<float*>(smem + K::OFF_POOLB) + c.t * MAX_NB * K::POOLW;
    c.biasBuf = reinterpret_cast<float*>(smem + K::OFF_BIAS) + c.t * MAX_NB * K::MAXC;
    c.v2buf = reinterpret_cast<float*>(smem + K::OFF_V2) + c.t * MAX_NB * MAX_V2;
    uint32_t layerCount = 0;
    bool alive = true;
    for(int item = blockIdx.x; item < numItems && alive; item += gridDim.x) {
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
          }
          c.par = par;
          named_bar_sync(1 + c.t, 128);
        }
        alive = mbar_wait(bars + (K::BAR_ACC + c.t) * 8, layerCount & 1, abortFlag, 31);
        if(!alive) break;
        tc_fence_after();
        if(L.epi == EPI_BN) epilogueBN<K>(P, L, c);
        else if(L.epi == EPI_GPOOL) epilogueGPool<K>(P, L, c);
        else epilogueHead<K>(P, L, c, item * NT + c.t, bars + (K::BAR_HEAD + c.t) * 8, sSym, nRows);
      }
    }
  }
  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if(warp == 1) { tc_fence_after(); tmem_dealloc(tmemBase, 512); }
}

// kc_forward input conversion: raw fp32 rows (NCHW/NHWC) + global -> symmetrised bf16 tiles
__global__ void k_convert_tiles(const float* __restrict__ raw, const float* __restrict__ rawGlobal, const int8_t* __restrict__ sym,
                                const uint8_t* __restrict__ dstOfSrc, uint4* __restrict__ tiles, int n, int numTiles,
                                int NB, int W, int H, int rawNHWC) {
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
      if(c < 15) f[q] = rawNHWC ? raw[((size_t)game * HW + srcCell) * 15 + c] : raw[((size_t)game * 15 + c) * HW + srcCell];
      else f[q] = rawGlobal[game];
    }
    v = make_uint4(pack2(f[0], f[1]), pack2(f[2], f[3]), pack2(f[4], f[5]), pack2(f[6], f[7]));
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
  template <class F>
  unsigned addConv(int N, int cinPadded, int ntaps, F weightAt) {
    unsigned off = (unsigned)w.size();
    int chunks = cinPadded / 16;
    for(int cc = 0; cc < chunks; cc++)
      for(int tap = 0; tap < ntaps; tap++) {
        size_t base = w.size();
        w.resize(base + (size_t)N * 32);
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(w.data() + base);
        for(int h = 0; h < 2; h++)
          for(int n = 0; n < N; n++)
            for(int j = 0; j < 8; j++)
              dst[((size_t)h * N + n) * 8 + j] = __float2bfloat16(weightAt(n, cc * 16 + h * 8 + j, tap));
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
  if(C % 16 != 0 || C > Cfg192::MAXC) return unsupported("trunk channels must be a multiple of 16 and <= 192 for the tcgen05 kernel");
  // <= 128 channels: two activation tiles per CTA; up to 192 (b15c192): one tile per CTA (TMEM budget)
  const int cfg = C <= Cfg128::MAXC ? 0 : 1;
  const int maxC = cfg == 0 ? Cfg128::MAXC : Cfg192::MAXC, maxG = cfg == 0 ? Cfg128::MAXG : Cfg192::MAXG;
  if(m->initialConv.ky != 3 || m->initialConv.kx != 3) return unsupported("initial conv must be 3x3");
  if(m->p1Conv.oc != HEADC || m->g1Conv.oc != HEADC || m->v1Conv.oc != HEADC) return unsupported("head convs must have 32 channels");
  if(m->p1Conv.ky != 1 || m->g1Conv.ky != 1 || m->v1Conv.ky != 1 || m->p2Conv.ky != 1 || m->vOwnershipConv.ky != 1)
    return unsupported("head convs must be 1x1");
  if(m->v2Mul.oc > MAX_V2) return unsupported("v2 size must be <= 128");
  if(m->trunkTipBN.act != 1 || m->g1BN.act != 1 || m->p1BN.act != 1 || m->v1BN.act != 1 || m->v2Act != 1)
    return unsupported("only ReLU activations are implemented in the tcgen05 kernel");
  for(const BlockW& b : m->blocks) {
    if(b.preBN.act != 1 || b.midBN.act != 1 || (b.kind == 2 && b.gpoolBN.act != 1)) return unsupported("only ReLU activations are implemented in the tcgen05 kernel");
    if(b.regularConv.ky != 3 || b.regularConv.kx != 3 || b.finalConv.ky != 3 || b.finalConv.kx != 3) return unsupported("block convs must be 3x3");
    if(b.kind == 0 && (b.regularConv.oc % 16 != 0 || b.regularConv.oc > maxC)) return unsupported("mid channels must be a multiple of 16 and within the trunk width class");
    if(b.kind == 2) {
      if(b.gpoolConv.ky != 3 || b.gpoolConv.oc % 16 != 0 || b.gpoolConv.oc > maxG) return unsupported("gpool conv must be 3x3 with a multiple of 16 channels, <= 32 (trunk <= 128) or <= 64 (trunk <= 192)");
      if(b.regularConv.oc % 16 != 0 || b.regularConv.oc + b.gpoolConv.oc > maxC) return unsupported("gpool block: regular + gpool channels must fit the trunk width class");
    }
  }
  if(2 * m->blocks.size() + 2 > (size_t)MAX_LAYERS) return unsupported("too many blocks for the tcgen05 kernel's layer table");
  TrunkProgram* T = new TrunkProgram();
  T->cfg = cfg;
  Packer pk;
  double macs = 0;
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
    if(b.kind == 2) L1.pOff = pk.addParams(cat({&b.gpoolBN.scale, &b.gpoolBN.bias, &b.gpoolToBias.h, &b.midBN.scale, &b.midBN.bias}));
    else L1.pOff = pk.addParams(cat({&b.midBN.scale, &b.midBN.bias}));
    T->layers.push_back(L1);
    LayerDesc L2{};
    L2.nk = 9 * (R / 16); L2.ntaps = 9; L2.N = C; L2.outSel = 0; L2.accumulate = 1; L2.epi = EPI_BN; L2.epiC = C;
    L2.wOffset = pk.addConv(C, R, 9, [&](int n, int ic, int tap) { return convW(b.finalConv, n, ic, tap); });
    const BNW& bn = bnOf(bi + 1);
    L2.pOff = pk.addParams(cat({&bn.scale, &bn.bias}));
    T->layers.push_back(L2);
    macs += 9.0 * C * (R + G) + 9.0 * R * C + 3.0 * G * R;
  }
  {
    LayerDesc L{};
    L.nk = C / 16; L.ntaps = 1; L.N = 3 * HEADC; L.outSel = 1; L.accumulate = 0; L.epi = EPI_HEAD; L.epiC = 0;
    L.wOffset = pk.addConv(3 * HEADC, C, 1, [&](int n, int ic, int) {
      if(n < HEADC) return convW(m->p1Conv, n, ic, 0);
      if(n < 2 * HEADC) return convW(m->g1Conv, n - HEADC, ic, 0);
      return convW(m->v1Conv, n - 2 * HEADC, ic, 0);
    });
    // p2Conv oc,ic,1,1 -> [ic][4]
    std::vector<float> w2((size_t)HEADC * 4), wown(HEADC);
    for(int k = 0; k < HEADC; k++) {
      for(int d = 0; d < 4; d++) w2[(size_t)k * 4 + d] = m->p2Conv.h[(size_t)d * HEADC + k];
      wown[k] = m->vOwnershipConv.h[k];
    }
    L.pOff = pk.addParams(cat({&m->g1BN.scale, &m->g1BN.bias, &m->gpoolToBiasMul.h, &m->p1BN.scale, &m->p1BN.bias, &w2, &m->v1BN.scale,
                               &m->v1BN.bias, &m->v2Mul.h, &m->v2Bias.h, &m->v3Mul.h, &m->v3Bias.h, &m->sv3Mul.h, &m->sv3Bias.h, &wown}));
    T->layers.push_back(L);
    macs += 3.0 * HEADC * C;
  }
  // Note: cat() above relies on each sub-block keeping the exact sizes the kernel indexes with
  // (HEADC, 96*HEADC, ...); alignment padding is only appended after the whole block.
  T->v2C = m->v2Mul.oc;
  T->wBytes = pk.w.size();
  T->flopsPerEval = 2.0 * macs;   // per board cell; multiplied by H*W by the caller
  if(cudaMalloc(&T->d_w, pk.w.size()) != cudaSuccess || cudaMalloc(&T->d_params, pk.p.size() * 4) != cudaSuccess ||
     cudaMalloc(&T->d_layers, T->layers.size() * sizeof(LayerDesc)) != cudaSuccess) {
    delete T;
    return kc::fail("buildTrunkProgram: out of device memory");
  }
  cudaMemcpy(T->d_w, pk.w.data(), pk.w.size(), cudaMemcpyHostToDevice);
  cudaMemcpy(T->d_params, pk.p.data(), pk.p.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(T->d_layers, T->layers.data(), T->layers.size() * sizeof(LayerDesc), cudaMemcpyHostToDevice);
  m->trunk = T;
  return 0;
}

void freeTrunkProgram(kc_model* m) {
  if(!m->trunk) return;
  cudaFree(m->trunk->d_w); cudaFree(m->trunk->d_params); cudaFree(m->trunk->d_layers);
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
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg128>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg128::SMEM));
  KC_CUDA(cudaFuncSetAttribute(trunk_kernel<Cfg192>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg192::SMEM));
  return 0;
}
void freeTrunkBuffers(kc_handle* h) {
  cudaFree(h->d_abort);
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
                                                               NB, h->W, h->H, rawNHWC);
  h->launches++;
  KC_CUDA(cudaGetLastError());
  return 0;
}

int runTrunkBf16(kc_handle* h, int n, cudaStream_t st, const int8_t* sym_dev, int rowOffset, const int* nDev) {
  const kc_model* m = h->model;
  const TrunkProgram* T = m->trunk;
  TrunkParams P{};
  P.wstream = T->d_w; P.params = T->d_params;
  P.numLayers = (int)T->layers.size();
  for(int i = 0; i < P.numLayers; i++) P.layers[i] = T->layers[i];
  P.NB = boardsPerTile(h->W, h->H); P.W = h->W; P.H = h->H; P.HW = h->W * h->H; P.stride = h->W + 1; P.tileRowW = P.NB * P.stride;
  int numTiles = (n + P.NB - 1) / P.NB;
  const int NT = T->cfg == 0 ? Cfg128::NT : Cfg192::NT;
  P.numItems = (numTiles + NT - 1) / NT;
  P.n = n;
  P.nDev = nDev;
  P.tiles = (const uint4*)h->d_tiles + (size_t)(rowOffset / P.NB) * 2 * TILE_ROWS;
  P.sym = sym_dev ? sym_dev + rowOffset : nullptr; P.dstOfSrcRev = h->d_dstOfSrcRev;
  P.policy = h->d_policy + (size_t)rowOffset * 4 * P.HW; P.value = h->d_value + (size_t)rowOffset * 2;
  P.misc = h->d_misc + (size_t)rowOffset * 2; P.own = h->d_own + (size_t)rowOffset * P.HW;
  P.abortFlag = h->d_abort;
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
  if(T->cfg == 0) trunk_kernel<Cfg128><<<grid, Cfg128::THREADS, Cfg128::SMEM, st>>>(P);
  else trunk_kernel<Cfg192><<<grid, Cfg192::THREADS, Cfg192::SMEM, st>>>(P);
  if(timed) { cudaEventRecord(h->evPool[h->evUsed + 1], st); h->evUsed += 2; }
  h->launches++;
  KC_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace kc

extern "C" {
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
