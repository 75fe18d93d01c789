// Model upload, compute handles and the fp32 "check mode" forward (CUDA-core FMA, NHWC).
//
// The check path mirrors the reference layer by layer (cpp/neuralnet/eigenbackend.cpp:1202-1226
// trunk, :912-930 / :968-1005 blocks, :1265-1298 policy head, :1341-1376 value head, :1675-1844
// getOutput) with direct convolutions, so that policy/value agree with the Eigen-algorithm oracle
// within 1e-4; it also backs the layer-level hooks (nninterface.h:127-169).  The production path
// is the bf16 tcgen05 trunk in net_bf16.cu; this file owns the handle and dispatches to it.
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <functional>
#include <mutex>
#include <thread>

#include "handle.h"

namespace kc {

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float activate(float x, int act) {
  if(act == 1) return fmaxf(x, 0.0f);
  if(act == 2) return x * tanhf(log1pf(expf(fminf(x, 20.0f))) + (fmaxf(x, 20.0f) - 20.0f));  // eigenbackend.cpp:729
  return x;
}

// raw rows (NCHW or NHWC, as the caller filled them) -> symmetrised NHWC fp32 (copyInputsWithSymmetry)
__global__ void k_convert_input(const float* __restrict__ raw, const int8_t* __restrict__ sym, const uint8_t* __restrict__ dstOfSrc,
                                float* __restrict__ out, int n, int HW, int C, int rawNHWC, int permuteDirs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n * HW * C) return;
  int c = i % C, p = (i / C) % HW, b = i / (C * HW);
  int s = sym ? sym[b] : 0;
  float v = rawNHWC ? raw[((size_t)b * HW + p) * C + c] : raw[((size_t)b * C + c) * HW + p];
  const int cd = permuteDirs ? playModeChannel(c, s) : c;   // play mode (ledger K): direction channels follow the symmetry
  out[((size_t)b * HW + dstOfSrc[s * HW + p]) * C + cd] = v;
}

__global__ void k_mask_from_input(const float* __restrict__ in, float* __restrict__ mask, float* __restrict__ maskSum, int n, int HW, int C) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= n) return;
  float s = 0.f;
  for(int p = 0; p < HW; p++) { float m = in[((size_t)b * HW + p) * C]; mask[(size_t)b * HW + p] = m; s += m; }
  maskSum[b] = s;
}
__global__ void k_mask_sum(const float* __restrict__ mask, float* __restrict__ maskSum, int n, int HW) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if(b >= n) return;
  float s = 0.f;
  for(int p = 0; p < HW; p++) s += mask[(size_t)b * HW + p];
  maskSum[b] = s;
}

// direct cross-correlation with zero padding; one thread per (row, oc); w is [tap][ic][oc]
__global__ void k_conv(const float* __restrict__ in, const float* __restrict__ w, float* __restrict__ out,
                       int n, int H, int W, int ic, int oc, int ky, int kx, int accumulate) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long total = (long long)n * H * W * oc;
  if(i >= total) return;
  int o = (int)(i % oc);
  long long row = i / oc;
  int x = (int)(row % W), y = (int)((row / W) % H);
  long long b = row / ((long long)W * H);
  int cy = ky / 2, cx = kx / 2;
  float acc = 0.f;
  for(int dy = 0; dy < ky; dy++) {
    int yy = y + dy - cy;
    if(yy < 0 || yy >= H) continue;
    for(int dx = 0; dx < kx; dx++) {
      int xx = x + dx - cx;
      if(xx < 0 || xx >= W) continue;
      const float* ip = in + ((b * H + yy) * W + xx) * ic;
      const float* wp = w + (size_t)(dy * kx + dx) * ic * oc + o;
      for(int k = 0; k < ic; k++) acc = fmaf(ip[k], wp[(size_t)k * oc], acc);
    }
  }
  if(accumulate) out[i] += acc; else out[i] = acc;
}

__global__ void k_bn(const float* __restrict__ in, const float* __restrict__ scale, const float* __restrict__ bias,
                     const float* __restrict__ mask, float* __restrict__ out, long long rows, int c, int act) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= rows * c) return;
  int k = (int)(i % c);
  long long row = i / c;
  bool on = mask == nullptr || mask[row] == 1.0f;
  out[i] = on ? activate(fmaf(in[i], scale[k], bias[k]), act) : 0.0f;
}

// eigenbackend.cpp:141-166 ; out [n][3c]
__global__ void k_gpool(const float* __restrict__ in, const float* __restrict__ mask, const float* __restrict__ maskSum,
                        float* __restrict__ out, int n, int HW, int c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n * c) return;
  int k = i % c, b = i / c;
  float s = 0.f, m = -1.0f;
  for(int p = 0; p < HW; p++) {
    float x = in[((size_t)b * HW + p) * c + k];
    s += x;
    float mv = mask ? mask[(size_t)b * HW + p] : 1.0f;
    m = fmaxf(m, x + (mv - 1.0f));
  }
  float div = maskSum[b], sq = sqrtf(div), mean = s / div;
  out[(size_t)b * 3 * c + k] = mean;
  out[(size_t)b * 3 * c + c + k] = mean * (sq - 14.0f) * 0.1f;
  out[(size_t)b * 3 * c + 2 * c + k] = m;
}
// eigenbackend.cpp:168-186
__global__ void k_valuepool(const float* __restrict__ in, const float* __restrict__ maskSum, float* __restrict__ out, int n, int HW, int c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n * c) return;
  int k = i % c, b = i / c;
  float s = 0.f;
  for(int p = 0; p < HW; p++) s += in[((size_t)b * HW + p) * c + k];
  float div = maskSum[b], sq = sqrtf(div), mean = s / div;
  out[(size_t)b * 3 * c + k] = mean;
  out[(size_t)b * 3 * c + c + k] = mean * (sq - 14.0f) * 0.1f;
  out[(size_t)b * 3 * c + 2 * c + k] = mean * ((sq - 14.0f) * (sq - 14.0f) * 0.01f - 0.1f);
}
// out[b][o] = act(sum_i in[b][i] w[i][o] + bias[o])
__global__ void k_matmul(const float* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                         float* __restrict__ out, int n, int ic, int oc, int act) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n * oc) return;
  int o = i % oc, b = i / oc;
  float acc = 0.f;
  for(int k = 0; k < ic; k++) acc = fmaf(in[(size_t)b * ic + k], w[(size_t)k * oc + o], acc);
  if(bias) acc += bias[o];
  out[i] = activate(acc, act);
}
__global__ void k_add_nc_bias(float* __restrict__ x, const float* __restrict__ bias, int n, int HW, int c) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= (long long)n * HW * c) return;
  int k = (int)(i % c);
  long long b = i / ((long long)HW * c);
  x[i] += bias[b * c + k];
}
// NHWC [n][HW][D] -> [n][d*HW + dst(p)] with the inverse spatial symmetry (copyOutputsWithSymmetry)
__global__ void k_spatial_out(const float* __restrict__ in, const int8_t* __restrict__ sym, const uint8_t* __restrict__ dstOfSrcRev,
                              float* __restrict__ out, int n, int HW, int D, int permuteDirs) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n * HW * D) return;
  int d = i % D, p = (i / D) % HW, b = i / (D * HW);
  int s = sym ? sym[b] : 0;
  const int dd = (permuteDirs && D == 4) ? symDirInline(d, s) : d;
  out[((size_t)b * D + dd) * HW + dstOfSrcRev[s * HW + p]] = in[i];
}
__global__ void k_nchw_to_nhwc(const float* __restrict__ in, float* __restrict__ out, int n, int c, int HW) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n * c * HW) return;
  int k = i % c, p = (i / c) % HW, b = i / (c * HW);
  out[i] = in[((size_t)b * c + k) * HW + p];
}
__global__ void k_nhwc_to_nchw(const float* __restrict__ in, float* __restrict__ out, int n, int c, int HW) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n * c * HW) return;
  int k = i % c, p = (i / c) % HW, b = i / (c * HW);
  out[((size_t)b * c + k) * HW + p] = in[i];
}

// NNEvaluator::evaluate post-processing (nneval.cpp:702-815), one warp per row: legal-masked softmax of the
// policy logits with temperature (illegal -> -1, all-underflow -> uniform), 2-way softmax of win/loss logits,
// softplus heads (desc.cpp:956-963 multipliers), flip to white's view, NNInputs::getHash with default
// parameters (nninputs.cpp:463-470: sit-hash ^ GAME_IS_OVER when finished).
__global__ void k_postprocess(const float* __restrict__ policyLogits, const float* __restrict__ valueLogits, const float* __restrict__ miscLogits,
                              const uint32_t* __restrict__ legal, const uint32_t* __restrict__ status, const uint64_t* __restrict__ sitHash,
                              int n, int policySize, int LW, float invTemp, float* __restrict__ policy, float* __restrict__ winLoss,
                              float* __restrict__ misc, uint64_t* __restrict__ nnHash, int* __restrict__ nonfinite) {
  int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if(row >= n) return;
  const float* p = policyLogits + (size_t)row * policySize;
  float mx = -1e25f;
  int cnt = 0;
  for(int i = lane; i < policySize; i += 32) {
    bool ok = (legal[(size_t)row * LW + (i >> 5)] >> (i & 31)) & 1u;
    float v = ok ? p[i] * invTemp : -1e30f;
    cnt += ok;
    mx = fmaxf(mx, v);
  }
  for(int o = 16; o > 0; o >>= 1) { mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o)); cnt += __shfl_xor_sync(0xffffffffu, cnt, o); }
  float sum = 0.f;
  for(int i = lane; i < policySize; i += 32) {
    bool ok = (legal[(size_t)row * LW + (i >> 5)] >> (i & 31)) & 1u;
    sum += expf((ok ? p[i] * invTemp : -1e30f) - mx);
  }
  for(int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  for(int i = lane; i < policySize; i += 32) {
    bool ok = (legal[(size_t)row * LW + (i >> 5)] >> (i & 31)) & 1u;
    float v;
    if(!ok) v = -1.0f;
    else if(sum <= 0.f) v = cnt > 0 ? 1.0f / cnt : 0.f;
    else v = expf(p[i] * invTemp - mx) / sum;
    policy[(size_t)row * policySize + i] = v;
  }
  if(lane == 0) {
    uint32_t st = status[row];
    int nextPla = (st >> 11) & 3;
    double w = valueLogits[2 * (size_t)row], l = valueLogits[2 * (size_t)row + 1];
    double m = fmax(w, l), ew = exp(w - m), el = exp(l - m), ps = ew + el;
    double winProb = ew / ps, lossProb = el / ps;
    // NNEvaluator::evaluate throws "Got nonfinite for policy sum" / "... nneval value" (nneval.cpp:745-750, 789-793): flag the batch
    if(nonfinite && (!isfinite(sum) || !isfinite(winProb) || !isfinite(lossProb))) atomicOr(nonfinite, 1);
    auto softPlus = [](double x) { return x > 40.0 ? x : log(1.0 + exp(x)); };
    double vtl = softPlus((double)miscLogits[2 * (size_t)row]) * 40.0;
    double s = softPlus((double)miscLogits[2 * (size_t)row + 1] * 0.5);
    winLoss[2 * (size_t)row] = (float)(nextPla == 2 ? winProb : lossProb);
    winLoss[2 * (size_t)row + 1] = (float)(nextPla == 2 ? lossProb : winProb);
    misc[2 * (size_t)row] = (float)vtl;
    misc[2 * (size_t)row + 1] = (float)sqrt(s * s * 0.25);
    bool fin = (st >> 8) & 1;
    nnHash[2 * (size_t)row] = sitHash[2 * (size_t)row] ^ (fin ? ZOBRIST_GAME_IS_OVER0 : 0ULL);
    nnHash[2 * (size_t)row + 1] = sitHash[2 * (size_t)row + 1] ^ (fin ? ZOBRIST_GAME_IS_OVER1 : 0ULL);
  }
}

static inline int blocksFor(long long total, int threads = 256) { return (int)((total + threads - 1) / threads); }

// ------------------------------------------------------------------------------------------------
// model upload
// ------------------------------------------------------------------------------------------------
static int uploadFloats(const std::vector<float>& h, float** d) {
  *d = nullptr;
  if(h.empty()) return 0;
  KC_CUDA(cudaMalloc(d, h.size() * sizeof(float)));
  KC_CUDA(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return 0;
}
static int initConv(ConvW& c, const kc_conv_desc& d, const char* name) {
  KC_CHECK(d.weights != nullptr, std::string("conv ") + name + ": null weights");
  KC_CHECK(d.convYSize > 0 && d.convXSize > 0 && (d.convYSize & 1) && (d.convXSize & 1) && d.inChannels > 0 && d.outChannels > 0,
           std::string("conv ") + name + ": bad shape");
  c.ky = d.convYSize; c.kx = d.convXSize; c.ic = d.inChannels; c.oc = d.outChannels;
  size_t n = (size_t)c.ky * c.kx * c.ic * c.oc;
  c.h.assign(d.weights, d.weights + n);
  std::vector<float> tap(n);
  for(int o = 0; o < c.oc; o++)
    for(int i = 0; i < c.ic; i++)
      for(int y = 0; y < c.ky; y++)
        for(int x = 0; x < c.kx; x++)
          tap[((size_t)(y * c.kx + x) * c.ic + i) * c.oc + o] = c.h[(((size_t)o * c.ic + i) * c.ky + y) * c.kx + x];
  return uploadFloats(tap, &c.d_tap);
}
static int initBN(BNW& b, const kc_bn_desc& d, int act, const char* name) {
  KC_CHECK(d.numChannels > 0 && d.mean && d.variance, std::string("batchnorm ") + name + ": bad description");
  b.c = d.numChannels; b.act = act;
  b.scale.resize(b.c); b.bias.resize(b.c);
  for(int i = 0; i < b.c; i++) {
    // vectors are used unconditionally (eigenbackend.cpp:707-710); null = the parser's default (desc.cpp:198-214)
    float s = d.scale ? d.scale[i] : 1.0f;
    float bb = d.bias ? d.bias[i] : 0.0f;
    b.scale[i] = s / sqrtf(d.variance[i] + d.epsilon);
    b.bias[i] = bb - b.scale[i] * d.mean[i];
  }
  if(uploadFloats(b.scale, &b.d_scale)) return 1;
  return uploadFloats(b.bias, &b.d_bias);
}
static int initMat(MatW& m, const kc_matmul_desc& d, const char* name) {
  KC_CHECK(d.inChannels > 0 && d.outChannels > 0 && d.weights, std::string("matmul ") + name + ": bad description");
  m.ic = d.inChannels; m.oc = d.outChannels;
  m.h.assign(d.weights, d.weights + (size_t)m.ic * m.oc);
  return uploadFloats(m.h, &m.d);
}
static int initBias(BiasW& b, const kc_matbias_desc& d, const char* name) {
  KC_CHECK(d.numChannels > 0 && d.weights, std::string("bias ") + name + ": bad description");
  b.c = d.numChannels;
  b.h.assign(d.weights, d.weights + b.c);
  return uploadFloats(b.h, &b.d);
}
static int initBlock(BlockW& b, const kc_block_desc& d) {
  KC_CHECK(d.kind == 0 || d.kind == 2, "block kind must be 0 (ordinary) or 2 (global pooling); nested bottleneck blocks are not supported");
  b.kind = d.kind;
  if(initBN(b.preBN, d.preBN, d.preActivation, "preBN")) return 1;
  if(initConv(b.regularConv, d.regularConv, "regularConv")) return 1;
  if(initBN(b.midBN, d.midBN, d.midActivation, "midBN")) return 1;
  if(initConv(b.finalConv, d.finalConv, "finalConv")) return 1;
  KC_CHECK(b.preBN.c == b.regularConv.ic && b.midBN.c == b.regularConv.oc && b.finalConv.ic == b.regularConv.oc && b.finalConv.oc == b.preBN.c,
           "block: channel counts do not chain");
  if(b.kind == 2) {
    if(initConv(b.gpoolConv, d.gpoolConv, "gpoolConv")) return 1;
    if(initBN(b.gpoolBN, d.gpoolBN, d.gpoolActivation, "gpoolBN")) return 1;
    if(initMat(b.gpoolToBias, d.gpoolToBiasMul, "gpoolToBiasMul")) return 1;
    KC_CHECK(b.gpoolConv.ic == b.preBN.c && b.gpoolBN.c == b.gpoolConv.oc && b.gpoolToBias.ic == 3 * b.gpoolConv.oc && b.gpoolToBias.oc == b.regularConv.oc,
             "gpool block: channel counts do not chain");
  }
  return 0;
}
static void freeConv(ConvW& c) { cudaFree(c.d_tap); }
static void freeBN(BNW& b) { cudaFree(b.d_scale); cudaFree(b.d_bias); }
static void freeBlock(BlockW& b) {
  freeBN(b.preBN); freeBN(b.gpoolBN); freeBN(b.midBN); freeConv(b.regularConv); freeConv(b.gpoolConv); freeConv(b.finalConv);
  cudaFree(b.gpoolToBias.d);
}

// one residual block on NHWC fp32 buffers (eigenbackend.cpp:912-930, 968-1005); trunk is updated in place
struct Fp32Scratch { float *a, *b, *c, *pool, *bias; };
static int runBlock(const BlockW& blk, int n, int H, int W, float* trunk, const float* mask, const float* maskSum,
                    Fp32Scratch s, cudaStream_t st, int64_t& launches) {
  int HW = H * W;
  long long rows = (long long)n * HW;
  k_bn<<<blocksFor(rows * blk.preBN.c), 256, 0, st>>>(trunk, blk.preBN.d_scale, blk.preBN.d_bias, mask, s.a, rows, blk.preBN.c, blk.preBN.act);
  k_conv<<<blocksFor(rows * blk.regularConv.oc), 256, 0, st>>>(s.a, blk.regularConv.d_tap, s.b, n, H, W, blk.regularConv.ic, blk.regularConv.oc,
                                                             blk.regularConv.ky, blk.regularConv.kx, 0);
  launches += 2;
  if(blk.kind == 2) {
    int gc = blk.gpoolConv.oc;
    k_conv<<<blocksFor(rows * gc), 256, 0, st>>>(s.a, blk.gpoolConv.d_tap, s.c, n, H, W, blk.gpoolConv.ic, gc, blk.gpoolConv.ky, blk.gpoolConv.kx, 0);
    k_bn<<<blocksFor(rows * gc), 256, 0, st>>>(s.c, blk.gpoolBN.d_scale, blk.gpoolBN.d_bias, mask, s.c, rows, gc, blk.gpoolBN.act);
    k_gpool<<<blocksFor((long long)n * gc), 256, 0, st>>>(s.c, mask, maskSum, s.pool, n, HW, gc);
    k_matmul<<<blocksFor((long long)n * blk.gpoolToBias.oc), 256, 0, st>>>(s.pool, blk.gpoolToBias.d, nullptr, s.bias, n, blk.gpoolToBias.ic, blk.gpoolToBias.oc, 0);
    k_add_nc_bias<<<blocksFor(rows * blk.regularConv.oc), 256, 0, st>>>(s.b, s.bias, n, HW, blk.regularConv.oc);
    launches += 5;
  }
  k_bn<<<blocksFor(rows * blk.midBN.c), 256, 0, st>>>(s.b, blk.midBN.d_scale, blk.midBN.d_bias, mask, s.a, rows, blk.midBN.c, blk.midBN.act);
  k_conv<<<blocksFor(rows * blk.finalConv.oc), 256, 0, st>>>(s.a, blk.finalConv.d_tap, trunk, n, H, W, blk.finalConv.ic, blk.finalConv.oc,
                                                           blk.finalConv.ky, blk.finalConv.kx, 1);
  launches += 2;
  KC_CUDA(cudaGetLastError());
  return 0;
}

// whole fp32 forward on the handle's symmetrised NHWC input
static int runFp32(kc_handle* h, int n, cudaStream_t st, const int8_t* sym_dev) {
  const kc_model* m = h->model;
  const int H = h->H, W = h->W, HW = H * W, C = m->trunkC;
  long long rows = (long long)n * HW;
  Fp32Buffers& f = h->f32;
  k_mask_from_input<<<blocksFor(n), 256, 0, st>>>(f.in, f.mask, f.maskSum, n, HW, m->numInputChannels);
  k_conv<<<blocksFor(rows * C), 256, 0, st>>>(f.in, m->initialConv.d_tap, f.trunk, n, H, W, m->initialConv.ic, C, m->initialConv.ky, m->initialConv.kx, 0);
  k_matmul<<<blocksFor((long long)n * C), 256, 0, st>>>(f.global, m->initialMatMul.d, nullptr, f.bias, n, m->initialMatMul.ic, C, 0);
  k_add_nc_bias<<<blocksFor(rows * C), 256, 0, st>>>(f.trunk, f.bias, n, HW, C);
  h->launches += 4;
  Fp32Scratch s{f.a, f.b, f.c, f.pool, f.bias};
  for(const BlockW& blk : m->blocks)
    if(runBlock(blk, n, H, W, f.trunk, f.mask, f.maskSum, s, st, h->launches)) return 1;
  k_bn<<<blocksFor(rows * C), 256, 0, st>>>(f.trunk, m->trunkTipBN.d_scale, m->trunkTipBN.d_bias, f.mask, f.tip, rows, C, m->trunkTipBN.act);
  // policy head
  int pc = m->p1Conv.oc, gc = m->g1Conv.oc, D = m->p2Conv.oc;
  k_conv<<<blocksFor(rows * pc), 256, 0, st>>>(f.tip, m->p1Conv.d_tap, f.a, n, H, W, C, pc, m->p1Conv.ky, m->p1Conv.kx, 0);
  k_conv<<<blocksFor(rows * gc), 256, 0, st>>>(f.tip, m->g1Conv.d_tap, f.b, n, H, W, C, gc, m->g1Conv.ky, m->g1Conv.kx, 0);
  k_bn<<<blocksFor(rows * gc), 256, 0, st>>>(f.b, m->g1BN.d_scale, m->g1BN.d_bias, f.mask, f.b, rows, gc, m->g1BN.act);
  k_gpool<<<blocksFor((long long)n * gc), 256, 0, st>>>(f.b, f.mask, f.maskSum, f.pool, n, HW, gc);
  k_matmul<<<blocksFor((long long)n * pc), 256, 0, st>>>(f.pool, m->gpoolToBiasMul.d, nullptr, f.bias, n, 3 * gc, pc, 0);
  k_add_nc_bias<<<blocksFor(rows * pc), 256, 0, st>>>(f.a, f.bias, n, HW, pc);
  k_bn<<<blocksFor(rows * pc), 256, 0, st>>>(f.a, m->p1BN.d_scale, m->p1BN.d_bias, f.mask, f.a, rows, pc, m->p1BN.act);
  k_conv<<<blocksFor(rows * D), 256, 0, st>>>(f.a, m->p2Conv.d_tap, f.c, n, H, W, pc, D, m->p2Conv.ky, m->p2Conv.kx, 0);
  k_spatial_out<<<blocksFor(rows * D), 256, 0, st>>>(f.c, sym_dev, h->d_dstOfSrcRev, h->d_policy, n, HW, D, (h->flags & KC_FLAG_SYM_PERMUTE_DIRS) ? 1 : 0);
  // value head
  int vc = m->v1Conv.oc, v2c = m->v2Mul.oc;
  k_conv<<<blocksFor(rows * vc), 256, 0, st>>>(f.tip, m->v1Conv.d_tap, f.a, n, H, W, C, vc, m->v1Conv.ky, m->v1Conv.kx, 0);
  k_bn<<<blocksFor(rows * vc), 256, 0, st>>>(f.a, m->v1BN.d_scale, m->v1BN.d_bias, f.mask, f.a, rows, vc, m->v1BN.act);
  k_valuepool<<<blocksFor((long long)n * vc), 256, 0, st>>>(f.a, f.maskSum, f.pool, n, HW, vc);
  k_matmul<<<blocksFor((long long)n * v2c), 256, 0, st>>>(f.pool, m->v2Mul.d, m->v2Bias.d, f.bias, n, 3 * vc, v2c, m->v2Act);
  k_matmul<<<blocksFor((long long)n * m->v3Mul.oc), 256, 0, st>>>(f.bias, m->v3Mul.d, m->v3Bias.d, h->d_value, n, v2c, m->v3Mul.oc, 0);
  k_matmul<<<blocksFor((long long)n * m->sv3Mul.oc), 256, 0, st>>>(f.bias, m->sv3Mul.d, m->sv3Bias.d, h->d_misc, n, v2c, m->sv3Mul.oc, 0);
  k_conv<<<blocksFor(rows), 256, 0, st>>>(f.a, m->vOwnershipConv.d_tap, f.c, n, H, W, vc, 1, m->vOwnershipConv.ky, m->vOwnershipConv.kx, 0);
  k_spatial_out<<<blocksFor(rows), 256, 0, st>>>(f.c, sym_dev, h->d_dstOfSrcRev, h->d_own, n, HW, 1, 0);
  h->launches += 18;
  KC_CUDA(cudaGetLastError());
  return 0;
}

// ---- accessors for games.cu ---------------------------------------------------------------------
int handleCheckGeometry(kc_handle* h, int W, int H, int n) {
  KC_CHECK(h->W == W && h->H == H, "games board size differs from the handle's nnXLen/nnYLen (exact size required)");
  KC_CHECK(n <= h->maxBatch, "numGames exceeds the handle's maxBatch");
  return 0;
}
bool handleIsBf16(const kc_handle* h) { return h->bf16; }
bool handlePermutesDirs(const kc_handle* h) { return (h->flags & KC_FLAG_SYM_PERMUTE_DIRS) != 0; }
void* handleInputTiles(kc_handle* h) { return h->d_tiles; }
float* handleRawInput(kc_handle* h) { return h->d_raw; }
float* handleRawGlobal(kc_handle* h) { return h->d_rawGlobal; }
int handleConvertRaw(kc_handle* h, int n, const int8_t* sym_dev, cudaStream_t stream) { return convertInputToTiles(h, n, /*rawNHWC=*/0, sym_dev, stream, 0); }
float* handleInputNHWC(kc_handle* h) { return h->f32.in; }
float* handleInputGlobal(kc_handle* h) { return h->f32.global; }
int handleCheckAbort(kc_handle* h) { return checkTrunkAbort(h); }
void launchPostprocess(kc_handle* h, int n, int LW, const uint32_t* legal_dev, const uint32_t* status_dev, const uint64_t* sitHash_dev,
                       float policyTemperature, float* policy_dev, float* winLoss_dev, float* misc_dev, uint64_t* nnHash_dev, cudaStream_t stream,
                       int rowOffset, int* nonfinite_dev) {
  const size_t ro = (size_t)rowOffset;
  k_postprocess<<<blocksFor((long long)n * 32), 256, 0, stream>>>(h->d_policy + ro * 4 * h->W * h->H, h->d_value + ro * 2, h->d_misc + ro * 2, legal_dev, status_dev, sitHash_dev, n,
                                                                 4 * h->W * h->H, LW, 1.0f / policyTemperature, policy_dev, winLoss_dev, misc_dev, nnHash_dev, nonfinite_dev);
}
void handleLeaveRegisters(kc_handle* h, bool on) { h->leaveRegisters = on; }
int handleRunOnStream(kc_handle* h, int n, cudaStream_t stream, const int8_t* sym_dev, const int* nDev, int rowOffset, bool symIsLocal) {
  h->lastN = n + rowOffset;
  if(h->bf16) return runTrunkBf16(h, n, stream, sym_dev, rowOffset, nDev, symIsLocal);
  return runFp32(h, n, stream, sym_dev);
}

}  // namespace kc

using namespace kc;

namespace kc {
// kc_forward / kc_forward_rows on the tensor path are pipelined over row chunks: H2D of chunk i+1 and D2H of chunk i-1 overlap
// the trunk kernel of chunk i (three streams, one event pair per chunk).  A chunk is a whole number of waves (rows that give every
// SM one work item).  Returns the chunk sizes in waves.
static std::vector<int> forwardChunkWaves(const kc_handle* h, int n, int* waveRows, bool hostGather = false) {
  const int NB = boardsPerTile(h->W, h->H);
  const int NT = handleTilesPerItem(h);
  const int wave = NT * NB * h->ctx->smCount;
  *waveRows = wave;
  // Tapered schedule (in waves): a small first chunk so that compute starts after a short copy, a small last chunk so
  // that little is left to copy back when the last kernel ends: 16 waves -> 1,3,11,1.
  const int totalWaves = (n + wave - 1) / wave;
  std::vector<int> sizes;
  if(hostGather && !getenv("KC_FORWARD_SCHEDULE")) {
    // kc_forward_rows: the host gathers chunk i+1 and scatters chunk i-1 while chunk i is on the device, so the chunks are even and
    // small enough that the first gather and the last scatter -- the two steps nothing overlaps -- stay short
    // Worker threads gather different chunks at the same time, so chunk c only has to be ready when the device gets to it: the
    // first chunks are one wave (the device starts after one short gather), then they grow (every chunk costs ~12 runtime calls on
    // the enqueueing thread), and the last one is a single wave again (its copy back and scatter overlap nothing).
    int left = totalWaves;
    const int ramp[] = {1, 1, 2};
    for(int i = 0; i < 3 && left > 1; i++) { const int v = std::min(ramp[i], left - 1); sizes.push_back(v); left -= v; }
    while(left > 1) { const int v = std::min(4, left - 1); sizes.push_back(v); left -= v; }
    if(left > 0) sizes.push_back(left);
    return sizes;
  }
  if(const char* env = getenv("KC_FORWARD_SCHEDULE")) {   // diagnostic override: chunk sizes in waves, e.g. "1,3,6,5,1" (the rest goes to a last chunk)
    int left = totalWaves;
    for(const char* q = env; *q && left > 0;) {
      const int v = std::max(1, std::min(left, atoi(q)));
      sizes.push_back(v); left -= v;
      while(*q && *q != ',') q++;
      if(*q == ',') q++;
    }
    if(left > 0) sizes.push_back(left);
  } else if(totalWaves >= 8) {
    // 1, 3, the bulk, 1: every chunk boundary costs a kernel prologue + the un-overlapped epilogue of its last work item
    // (about 25 us), so few chunks; measured at 16 waves: 1,2,4,4,4,1 6.63 / 1,2,4,8,1 6.64 / 1,4,10,1 6.64 / 1,3,11,1 6.75 M evals/s
    sizes.push_back(1); sizes.push_back(3); sizes.push_back(totalWaves - 5); sizes.push_back(1);
  } else {
    for(int w = 0; w < totalWaves; w += 2) sizes.push_back(std::min(2, totalWaves - w));
  }
  return sizes;
}

// ---- kc_forward_rows: scattered host rows in, scattered outputs back ------------------------------------------------------------
// Worker threads for the host-side gather / scatter.  One wake-up per call: a job is a function every worker (and the caller) runs
// once; inside, the threads take whole row chunks from shared atomic counters, so there is no fork / join per chunk.
struct RowWorkers {
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable wake, done;
  const std::function<void()>* job = nullptr;
  int generation = 0, pending = 0;
  bool stop = false;
  explicit RowWorkers(int n) {
    for(int i = 0; i < n; i++)
      threads.emplace_back([this] {
        int seen = 0;
        for(;;) {
          std::unique_lock<std::mutex> lk(m);
          wake.wait(lk, [&] { return stop || generation != seen; });
          if(stop) return;
          seen = generation;
          const std::function<void()>* f = job;
          lk.unlock();
          (*f)();
          lk.lock();
          if(--pending == 0) done.notify_one();
        }
      });
  }
  ~RowWorkers() {
    { std::lock_guard<std::mutex> lk(m); stop = true; }
    wake.notify_all();
    for(std::thread& t : threads) t.join();
  }
  void start(const std::function<void()>& f) {   // f must stay alive until finish() returns
    if(threads.empty()) return;
    { std::lock_guard<std::mutex> lk(m); job = &f; pending = (int)threads.size(); generation++; }
    wake.notify_all();
  }
  void finish() {
    if(threads.empty()) return;
    std::unique_lock<std::mutex> lk(m);
    done.wait(lk, [&] { return pending == 0; });
  }
};
struct RowStaging {
  float *spatial = nullptr, *global = nullptr, *policy = nullptr, *value = nullptr, *misc = nullptr, *own = nullptr;   // page-locked, [maxBatch] rows
  int8_t* sym = nullptr;
  RowWorkers workers;
  explicit RowStaging(int nWorkers) : workers(nWorkers) {}
  ~RowStaging() { cudaFreeHost(spatial); cudaFreeHost(global); cudaFreeHost(policy); cudaFreeHost(value); cudaFreeHost(misc); cudaFreeHost(own); cudaFreeHost(sym); }
};
void freeRowStaging(kc_handle* h) { delete h->rows; h->rows = nullptr; }

static int rowStaging(kc_handle* h) {
  if(h->rows) return 0;
  static const int nEnv = [] { const char* e = getenv("KC_FORWARD_ROWS_THREADS"); return e ? atoi(e) : -1; }();
  const int hc = (int)std::thread::hardware_concurrency();
  const int nThreads = nEnv >= 0 ? nEnv : std::max(1, std::min(4, hc > 0 ? hc / 2 : 2));   // gather / scatter workers beside the caller (which enqueues)
  RowStaging* r = new RowStaging(h->maxBatch >= 4096 ? nThreads : 0);
  h->rows = r;
  const size_t B = (size_t)h->maxBatch, HW = (size_t)h->W * h->H;
  KC_CUDA(cudaHostAlloc(&r->spatial, B * 15 * HW * 4, cudaHostAllocDefault)); KC_CUDA(cudaHostAlloc(&r->global, B * 4, cudaHostAllocDefault));
  KC_CUDA(cudaHostAlloc(&r->policy, B * 4 * HW * 4, cudaHostAllocDefault)); KC_CUDA(cudaHostAlloc(&r->value, B * 8, cudaHostAllocDefault));
  KC_CUDA(cudaHostAlloc(&r->misc, B * 8, cudaHostAllocDefault)); KC_CUDA(cudaHostAlloc(&r->own, B * HW * 4, cudaHostAllocDefault));
  KC_CUDA(cudaHostAlloc(&r->sym, B, cudaHostAllocDefault));
  return 0;
}

}  // namespace kc

extern "C" {

int kc_device_count(int* count) {
  KC_CHECK(count, "kc_device_count: null argument");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if(e != cudaSuccess) { *count = 0; return kc::fail(std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
  *count = n;
  return 0;
}

int kc_ctx_create(int device, kc_ctx** out) {
  KC_CHECK(out, "kc_ctx_create: null argument");
  int n = 0;
  KC_CUDA(cudaGetDeviceCount(&n));
  KC_CHECK(n > 0, "kc_ctx_create: no CUDA device visible (this library has no CPU fallback)");
  if(device < 0) device = 0;
  KC_CHECK(device < n, "kc_ctx_create: device index out of range");
  cudaDeviceProp prop;
  KC_CUDA(cudaGetDeviceProperties(&prop, device));
  KC_CHECK(prop.major == 10, std::string("kc_ctx_create: device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                 "; this library is built for sm_100a (B200) only");
  KC_CUDA(cudaSetDevice(device));
  kc_ctx* c = new kc_ctx();
  c->device = device;
  c->smCount = prop.multiProcessorCount;
  c->smemOptin = prop.sharedMemPerBlockOptin;
  (void)zobrist();
  *out = c;
  return 0;
}
int kc_ctx_destroy(kc_ctx* ctx) { delete ctx; return 0; }

int kc_model_create(kc_ctx* ctx, const kc_model_desc* d, kc_model** out) {
  KC_CHECK(ctx && d && out, "kc_model_create: null argument");
  KC_CUDA(cudaSetDevice(ctx->device));
  KC_CHECK(d->version == 1, "kc_model_create: only model version 1 exists for Coffee (cpp/neuralnet/modelversion.h:7-12)");
  KC_CHECK(d->numInputChannels == KC_NUM_SPATIAL_V1 && d->numInputGlobalChannels == KC_NUM_GLOBAL_V1,
           "kc_model_create: V1 inputs are 15 spatial + 1 global channels");
  KC_CHECK(d->numBlocks >= 0 && (d->numBlocks == 0 || d->blocks), "kc_model_create: bad block list");
  kc_model* m = new kc_model();
  m->ctx = ctx;
  m->numInputChannels = d->numInputChannels; m->numInputGlobalChannels = d->numInputGlobalChannels;
  m->trunkC = d->trunkNumChannels;
  int rc = 0;
  rc |= initConv(m->initialConv, d->initialConv, "initialConv");
  rc = rc || initMat(m->initialMatMul, d->initialMatMul, "initialMatMul");
  m->blocks.resize(d->numBlocks);
  for(int i = 0; i < d->numBlocks && !rc; i++) rc = initBlock(m->blocks[i], d->blocks[i]);
  rc = rc || initBN(m->trunkTipBN, d->trunkTipBN, d->trunkTipActivation, "trunkTipBN");
  rc = rc || initConv(m->p1Conv, d->p1Conv, "p1Conv") || initConv(m->g1Conv, d->g1Conv, "g1Conv") ||
       initBN(m->g1BN, d->g1BN, d->g1Activation, "g1BN") || initMat(m->gpoolToBiasMul, d->gpoolToBiasMul, "policy gpoolToBiasMul") ||
       initBN(m->p1BN, d->p1BN, d->p1Activation, "p1BN") || initConv(m->p2Conv, d->p2Conv, "p2Conv");
  rc = rc || initConv(m->v1Conv, d->v1Conv, "v1Conv") || initBN(m->v1BN, d->v1BN, d->v1Activation, "v1BN") ||
       initMat(m->v2Mul, d->v2Mul, "v2Mul") || initBias(m->v2Bias, d->v2Bias, "v2Bias") || initMat(m->v3Mul, d->v3Mul, "v3Mul") ||
       initBias(m->v3Bias, d->v3Bias, "v3Bias") || initMat(m->sv3Mul, d->sv3Mul, "sv3Mul") || initBias(m->sv3Bias, d->sv3Bias, "sv3Bias") ||
       initConv(m->vOwnershipConv, d->vOwnershipConv, "vOwnershipConv");
  m->v2Act = d->v2Activation;
  if(!rc) {
    // Coffee head shapes and chaining
    bool ok = m->initialConv.ic == 15 && m->initialConv.oc == m->trunkC && m->initialMatMul.ic == 1 && m->initialMatMul.oc == m->trunkC &&
              m->trunkTipBN.c == m->trunkC && m->p1Conv.ic == m->trunkC && m->g1Conv.ic == m->trunkC && m->v1Conv.ic == m->trunkC &&
              m->g1BN.c == m->g1Conv.oc && m->gpoolToBiasMul.ic == 3 * m->g1Conv.oc && m->gpoolToBiasMul.oc == m->p1Conv.oc &&
              m->p1BN.c == m->p1Conv.oc && m->p2Conv.ic == m->p1Conv.oc && m->p2Conv.oc == 4 && m->v1BN.c == m->v1Conv.oc &&
              m->v2Mul.ic == 3 * m->v1Conv.oc && m->v2Bias.c == m->v2Mul.oc && m->v3Mul.ic == m->v2Mul.oc && m->v3Mul.oc == 2 && m->v3Bias.c == 2 &&
              m->sv3Mul.ic == m->v2Mul.oc && m->sv3Mul.oc == 2 && m->sv3Bias.c == 2 && m->vOwnershipConv.ic == m->v1Conv.oc && m->vOwnershipConv.oc == 1;
    for(const BlockW& b : m->blocks) ok = ok && b.preBN.c == m->trunkC;
    if(!ok) rc = kc::fail("kc_model_create: layer shapes do not chain or are not the Coffee head shapes (policy 4 channels, value 2, misc 2, ownership 1)");
  }
  if(!rc) rc = buildTrunkProgram(m);
  if(rc) { kc_model_destroy(m); return 1; }
  *out = m;
  return 0;
}

int kc_model_destroy(kc_model* m) {
  if(!m) return 0;
  cudaSetDevice(m->ctx->device);
  freeTrunkProgram(m);
  freeConv(m->initialConv); cudaFree(m->initialMatMul.d);
  for(BlockW& b : m->blocks) freeBlock(b);
  freeBN(m->trunkTipBN);
  freeConv(m->p1Conv); freeConv(m->g1Conv); freeConv(m->p2Conv); freeBN(m->g1BN); freeBN(m->p1BN); cudaFree(m->gpoolToBiasMul.d);
  freeConv(m->v1Conv); freeConv(m->vOwnershipConv); freeBN(m->v1BN);
  cudaFree(m->v2Mul.d); cudaFree(m->v3Mul.d); cudaFree(m->sv3Mul.d); cudaFree(m->v2Bias.d); cudaFree(m->v3Bias.d); cudaFree(m->sv3Bias.d);
  delete m;
  return 0;
}

int kc_handle_create(kc_ctx* ctx, const kc_model* model, int maxBatch, int nnXLen, int nnYLen, unsigned flags, kc_handle** out) {
  KC_CHECK(ctx && model && out, "kc_handle_create: null argument");
  KC_CHECK(maxBatch > 0, "kc_handle_create: maxBatch must be positive");
  KC_CHECK(nnXLen >= 2 && nnYLen >= 2 && nnXLen <= KC_MAX_LEN && nnYLen <= KC_MAX_LEN, "kc_handle_create: nnXLen/nnYLen out of range");
  KC_CUDA(cudaSetDevice(ctx->device));
  kc_handle* h = new kc_handle();
  h->ctx = ctx; h->model = model; h->maxBatch = maxBatch; h->W = nnXLen; h->H = nnYLen; h->flags = flags;
  h->bf16 = !(flags & KC_FLAG_FP32_CHECK);
  const int HW = nnXLen * nnYLen, C = model->trunkC;
  size_t nb = (size_t)maxBatch;
  int rc = 0;
  auto alloc = [&](auto** p, size_t bytes) { if(!rc && cudaMalloc((void**)p, bytes) != cudaSuccess) rc = kc::fail("kc_handle_create: out of device memory"); };
  if(h->bf16) {
    if(!model->trunk) { delete h; return kc::fail("kc_handle_create: bf16 tcgen05 path unavailable for this model: " + model->trunkUnsupportedWhy); }
    // a 128-row activation tile holds whole boards with one pad column each: every size up to the reference's 10x10 (board.h:120) fits
    // (10 x 11 = 110 rows, one board per tile); the halo must cover one tile row + 1
    if(boardsPerTile(nnXLen, nnYLen) < 1 || boardsPerTile(nnXLen, nnYLen) * (nnXLen + 1) + 1 > HALO_ROWS) {
      delete h; return kc::fail("kc_handle_create: the tensor path needs H*(W+1) <= 128");
    }
  }
  if(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess ||
     cudaStreamCreateWithFlags(&h->h2dStream, cudaStreamNonBlocking) != cudaSuccess ||
     cudaStreamCreateWithFlags(&h->d2hStream, cudaStreamNonBlocking) != cudaSuccess) rc = kc::fail("kc_handle_create: cudaStreamCreate failed");
  cudaEventCreate(&h->ev0); cudaEventCreate(&h->ev1);
  alloc(&h->d_raw, nb * 15 * HW * 4); alloc(&h->d_rawGlobal, nb * 4); alloc(&h->d_sym, nb);
  alloc(&h->d_policy, nb * 4 * HW * 4); alloc(&h->d_value, nb * 2 * 4); alloc(&h->d_misc, nb * 2 * 4); alloc(&h->d_own, nb * HW * 4);
  alloc(&h->d_dstOfSrc, 8 * HW); alloc(&h->d_dstOfSrcRev, 8 * HW);
  if(!rc) {
    std::vector<uint8_t> fwd(8 * HW), rev(8 * HW);
    std::vector<int> tmp(HW);
    for(int s = 0; s < 8; s++) {
      symmetryDstOfSrc(nnYLen, nnXLen, s, false, tmp.data());
      for(int p = 0; p < HW; p++) fwd[s * HW + p] = (uint8_t)tmp[p];
      symmetryDstOfSrc(nnYLen, nnXLen, s, true, tmp.data());
      for(int p = 0; p < HW; p++) rev[s * HW + p] = (uint8_t)tmp[p];
    }
    cudaMemcpy(h->d_dstOfSrc, fwd.data(), fwd.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(h->d_dstOfSrcRev, rev.data(), rev.size(), cudaMemcpyHostToDevice);
  }
  if(h->bf16) {
    if(!rc) rc = allocTrunkBuffers(h);
  } else {
    Fp32Buffers& f = h->f32;
    int maxC = C;
    for(const BlockW& b : model->blocks) maxC = std::max(maxC, std::max(b.regularConv.oc, b.gpoolConv.oc));
    maxC = std::max(maxC, std::max(model->p1Conv.oc, std::max(model->g1Conv.oc, model->v1Conv.oc)));
    size_t big = nb * HW * (size_t)maxC * 4;
    alloc(&f.in, nb * HW * 15 * 4); alloc(&f.global, nb * 4); alloc(&f.mask, nb * HW * 4); alloc(&f.maskSum, nb * 4);
    alloc(&f.trunk, big); alloc(&f.tip, big); alloc(&f.a, big); alloc(&f.b, big); alloc(&f.c, big);
    alloc(&f.pool, nb * 3 * (size_t)maxC * 4); alloc(&f.bias, nb * (size_t)std::max(maxC, model->v2Mul.oc) * 4);
  }
  if(rc) { kc_handle_destroy(h); return 1; }
  *out = h;
  return 0;
}

int kc_handle_destroy(kc_handle* h) {
  if(!h) return 0;
  cudaSetDevice(h->ctx->device);
  if(h->stream) cudaStreamSynchronize(h->stream);
  cudaFree(h->d_raw); cudaFree(h->d_rawGlobal); cudaFree(h->d_sym); cudaFree(h->d_policy); cudaFree(h->d_value); cudaFree(h->d_misc);
  cudaFree(h->d_own); cudaFree(h->d_dstOfSrc); cudaFree(h->d_dstOfSrcRev); cudaFree(h->d_tiles);
  Fp32Buffers& f = h->f32;
  cudaFree(f.in); cudaFree(f.global); cudaFree(f.mask); cudaFree(f.maskSum); cudaFree(f.trunk); cudaFree(f.tip); cudaFree(f.a); cudaFree(f.b);
  cudaFree(f.c); cudaFree(f.pool); cudaFree(f.bias);
  freeTrunkBuffers(h);
  freeRowStaging(h);
  if(h->ev0) cudaEventDestroy(h->ev0);
  if(h->ev1) cudaEventDestroy(h->ev1);
  if(h->stream) cudaStreamDestroy(h->stream);
  if(h->h2dStream) cudaStreamDestroy(h->h2dStream);
  if(h->d2hStream) cudaStreamDestroy(h->d2hStream);
  for(cudaEvent_t e : h->chunkEvents) cudaEventDestroy(e);
  delete h;
  return 0;
}

int kc_handle_uses_bf16(const kc_handle* h) { return h && h->bf16 ? 1 : 0; }
int kc_handle_operand_format(const kc_handle* h) { return !h || !h->bf16 ? 2 : (h->flags & KC_FLAG_OPERANDS_BF16) ? 1 : 0; }
int64_t kc_handle_launch_count(const kc_handle* h) { return h ? h->launches : 0; }
int kc_handle_trunk_time(kc_handle* h, float* sumMs, int* count) {
  KC_CHECK(h && sumMs && count, "kc_handle_trunk_time: null argument");
  KC_CUDA(cudaSetDevice(h->ctx->device));
  KC_CUDA(cudaDeviceSynchronize());
  float sum = 0.f;
  for(int i = 0; i + 1 < h->evUsed; i += 2) {
    float ms = 0.f;
    KC_CUDA(cudaEventElapsedTime(&ms, h->evPool[i], h->evPool[i + 1]));
    sum += ms;
  }
  *sumMs = sum; *count = h->evUsed / 2;
  h->evUsed = 0;
  return 0;
}

int kc_handle_read_outputs(kc_handle* h, int n, float* policy, float* value, float* misc, float* ownership) {
  KC_CHECK(h && n > 0 && n <= h->maxBatch, "kc_handle_read_outputs: bad argument");
  KC_CUDA(cudaSetDevice(h->ctx->device));
  const int HW = h->H * h->W;
  KC_CUDA(cudaDeviceSynchronize());   // the evaluation may have run on a games stream
  if(checkTrunkAbort(h)) return 1;
  if(policy) KC_CUDA(cudaMemcpy(policy, h->d_policy, (size_t)n * 4 * HW * 4, cudaMemcpyDeviceToHost));
  if(value) KC_CUDA(cudaMemcpy(value, h->d_value, (size_t)n * 2 * 4, cudaMemcpyDeviceToHost));
  if(misc) KC_CUDA(cudaMemcpy(misc, h->d_misc, (size_t)n * 2 * 4, cudaMemcpyDeviceToHost));
  if(ownership) KC_CUDA(cudaMemcpy(ownership, h->d_own, (size_t)n * HW * 4, cudaMemcpyDeviceToHost));
  return 0;
}

int kc_host_alloc(size_t bytes, void** out) {
  KC_CHECK(out && bytes > 0, "kc_host_alloc: bad argument");
  KC_CUDA(cudaHostAlloc(out, bytes, cudaHostAllocDefault));
  return 0;
}
int kc_host_free(void* p) {
  if(p) KC_CUDA(cudaFreeHost(p));
  return 0;
}

int kc_forward(kc_handle* h, int n, const float* spatial, const float* global, const int8_t* symmetry,
               float* policy, float* value, float* misc, float* ownership) {
  KC_CHECK(h && spatial && global && policy && value && misc, "kc_forward: null argument");
  KC_CHECK(n > 0 && n <= h->maxBatch, "kc_forward: need 0 < n <= maxBatch (nninterface.h:112-117)");
  KC_CUDA(cudaSetDevice(h->ctx->device));
  const int HW = h->H * h->W;
  cudaStream_t st = h->stream;
  if(symmetry)
    for(int i = 0; i < n; i++) KC_CHECK(symmetry[i] >= 0 && symmetry[i] < 8, "kc_forward: symmetry must be within 0..7");
  int rawNHWC = (h->flags & KC_FLAG_INPUTS_NHWC) ? 1 : 0;
  if(h->bf16) {
    // Pipelined over row chunks: H2D of chunk i+1 and D2H of chunk i-1 overlap the trunk kernel of chunk i
    // (three streams, one event pair per chunk).  A chunk is a whole number of CTA work items per SM.
    int wave = 0;
    const std::vector<int> sizes = forwardChunkWaves(h, n, &wave);
    const int numChunks = (int)sizes.size();
    while((int)h->chunkEvents.size() < 2 * numChunks) {
      cudaEvent_t e;
      KC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      h->chunkEvents.push_back(e);
    }
    int r0 = 0;
    for(int c = 0; c < numChunks; r0 += sizes[c] * wave, c++) {
      const int rows = std::min(sizes[c] * wave, n - r0);
      KC_CUDA(cudaMemcpyAsync(h->d_raw + (size_t)r0 * 15 * HW, spatial + (size_t)r0 * 15 * HW, (size_t)rows * 15 * HW * 4, cudaMemcpyHostToDevice, h->h2dStream));
      KC_CUDA(cudaMemcpyAsync(h->d_rawGlobal + r0, global + r0, (size_t)rows * 4, cudaMemcpyHostToDevice, h->h2dStream));
      if(symmetry) KC_CUDA(cudaMemcpyAsync(h->d_sym + r0, symmetry + r0, (size_t)rows, cudaMemcpyHostToDevice, h->h2dStream));
      KC_CUDA(cudaEventRecord(h->chunkEvents[2 * c], h->h2dStream));
      KC_CUDA(cudaStreamWaitEvent(st, h->chunkEvents[2 * c], 0));
      if(convertInputToTiles(h, rows, rawNHWC, symmetry ? h->d_sym : nullptr, st, r0)) return 1;
      if(runTrunkBf16(h, rows, st, symmetry ? h->d_sym : nullptr, r0)) return 1;
      KC_CUDA(cudaEventRecord(h->chunkEvents[2 * c + 1], st));
      KC_CUDA(cudaStreamWaitEvent(h->d2hStream, h->chunkEvents[2 * c + 1], 0));
      KC_CUDA(cudaMemcpyAsync(policy + (size_t)r0 * 4 * HW, h->d_policy + (size_t)r0 * 4 * HW, (size_t)rows * 4 * HW * 4, cudaMemcpyDeviceToHost, h->d2hStream));
      KC_CUDA(cudaMemcpyAsync(value + (size_t)r0 * 2, h->d_value + (size_t)r0 * 2, (size_t)rows * 2 * 4, cudaMemcpyDeviceToHost, h->d2hStream));
      KC_CUDA(cudaMemcpyAsync(misc + (size_t)r0 * 2, h->d_misc + (size_t)r0 * 2, (size_t)rows * 2 * 4, cudaMemcpyDeviceToHost, h->d2hStream));
      if(ownership) KC_CUDA(cudaMemcpyAsync(ownership + (size_t)r0 * HW, h->d_own + (size_t)r0 * HW, (size_t)rows * HW * 4, cudaMemcpyDeviceToHost, h->d2hStream));
    }
    h->lastN = n;
    KC_CUDA(cudaStreamSynchronize(h->d2hStream));
    KC_CUDA(cudaStreamSynchronize(st));
    return checkTrunkAbort(h);
  }
  KC_CUDA(cudaMemcpyAsync(h->d_raw, spatial, (size_t)n * 15 * HW * 4, cudaMemcpyHostToDevice, st));
  KC_CUDA(cudaMemcpyAsync(h->d_rawGlobal, global, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  if(symmetry) KC_CUDA(cudaMemcpyAsync(h->d_sym, symmetry, (size_t)n, cudaMemcpyHostToDevice, st));
  const int8_t* sym_dev = symmetry ? h->d_sym : nullptr;
  k_convert_input<<<blocksFor((long long)n * HW * 15), 256, 0, st>>>(h->d_raw, sym_dev, h->d_dstOfSrc, h->f32.in, n, HW, 15, rawNHWC, (h->flags & KC_FLAG_SYM_PERMUTE_DIRS) ? 1 : 0);
  KC_CUDA(cudaMemcpyAsync(h->f32.global, h->d_rawGlobal, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
  h->launches += 1;
  if(handleRunOnStream(h, n, st, sym_dev)) return 1;
  KC_CUDA(cudaMemcpyAsync(policy, h->d_policy, (size_t)n * 4 * HW * 4, cudaMemcpyDeviceToHost, st));
  KC_CUDA(cudaMemcpyAsync(value, h->d_value, (size_t)n * 2 * 4, cudaMemcpyDeviceToHost, st));
  KC_CUDA(cudaMemcpyAsync(misc, h->d_misc, (size_t)n * 2 * 4, cudaMemcpyDeviceToHost, st));
  if(ownership) KC_CUDA(cudaMemcpyAsync(ownership, h->d_own, (size_t)n * HW * 4, cudaMemcpyDeviceToHost, st));
  KC_CUDA(cudaStreamSynchronize(st));
  return 0;
}

// NeuralNet::getOutput over rows that lie scattered in host memory (the NNResultBuf / NNOutput objects of nneval.h:45-65,
// nninputs.h:75-118): the gather into page-locked staging, the copies, the kernels and the scatter of chunk i run while the host
// gathers chunk i+1 and scatters chunk i-1.  Same results as kc_forward on the gathered rows, bit for bit.
int kc_forward_rows(kc_handle* h, int n, const float* const* spatialRows, const float* const* globalRows, const int8_t* symmetry,
                    float* const* policyRows, float* const* scalarRows, float* const* ownerRows) {
  KC_CHECK(h && spatialRows && globalRows && policyRows && scalarRows, "kc_forward_rows: null argument");
  KC_CHECK(n > 0 && n <= h->maxBatch, "kc_forward_rows: need 0 < n <= maxBatch (nninterface.h:112-117)");
  KC_CUDA(cudaSetDevice(h->ctx->device));
  if(symmetry)
    for(int i = 0; i < n; i++) KC_CHECK(symmetry[i] >= 0 && symmetry[i] < 8, "kc_forward_rows: symmetry must be within 0..7");
  if(rowStaging(h)) return 1;
  RowStaging& R = *h->rows;
  const int HW = h->H * h->W;
  const size_t rowBytes = (size_t)15 * HW * 4;
  auto gather = [&](int r0, int rows) {
    for(int i = r0; i < r0 + rows; i++) {
      memcpy(R.spatial + (size_t)i * 15 * HW, spatialRows[i], rowBytes);
      R.global[i] = globalRows[i][0];
      R.sym[i] = symmetry ? symmetry[i] : 0;
    }
  };
  auto scatter = [&](int r0, int rows) {
    for(int i = r0; i < r0 + rows; i++) {
      memcpy(policyRows[i], R.policy + (size_t)i * 4 * HW, (size_t)4 * HW * 4);
      float* sc = scalarRows[i];
      sc[0] = R.value[2 * i]; sc[1] = R.value[2 * i + 1]; sc[2] = R.misc[2 * i]; sc[3] = R.misc[2 * i + 1];
      if(ownerRows && ownerRows[i]) memcpy(ownerRows[i], R.own + (size_t)i * HW, (size_t)HW * 4);
    }
  };
  if(!h->bf16) {   // check path: no pipeline
    gather(0, n);
    if(kc_forward(h, n, R.spatial, R.global, symmetry ? R.sym : nullptr, R.policy, R.value, R.misc, R.own)) return 1;
    scatter(0, n);
    return 0;
  }
  cudaStream_t st = h->stream;
  const int rawNHWC = (h->flags & KC_FLAG_INPUTS_NHWC) ? 1 : 0;
  int wave = 0;
  const std::vector<int> sizes = forwardChunkWaves(h, n, &wave, true);
  const int numChunks = (int)sizes.size();
  while((int)h->chunkEvents.size() < 3 * numChunks) {
    cudaEvent_t e;
    KC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    h->chunkEvents.push_back(e);
  }
  std::vector<int> start(numChunks + 1, 0);
  for(int c = 0; c < numChunks; c++) start[c + 1] = std::min(n, start[c] + sizes[c] * wave);
  // Host side: every thread takes whole chunks -- first to gather (in chunk order), then, once a chunk's copy back has been enqueued
  // and has finished, to scatter.  The caller's thread enqueues chunk c as soon as it is gathered, then joins the scattering.
  std::vector<std::atomic<int>> gathered(numChunks), enqueued(numChunks);
  for(int c = 0; c < numChunks; c++) { gathered[c].store(0); enqueued[c].store(0); }
  std::atomic<int> nextGather{0}, nextScatter{0}, failed{0};
  const int device = h->ctx->device;
  auto scatterLoop = [&]() {
    for(int c; (c = nextScatter.fetch_add(1)) < numChunks;) {
      while(!enqueued[c].load(std::memory_order_acquire)) {
        if(failed.load()) return;
        std::this_thread::yield();
      }
      if(cudaEventSynchronize(h->chunkEvents[3 * c + 2]) != cudaSuccess) { failed.store(1); return; }
      scatter(start[c], start[c + 1] - start[c]);
    }
  };
  const std::function<void()> workerJob = [&]() {
    cudaSetDevice(device);
    for(int c; (c = nextGather.fetch_add(1)) < numChunks;) {
      gather(start[c], start[c + 1] - start[c]);
      gathered[c].store(1, std::memory_order_release);
    }
    scatterLoop();
  };
  const bool useWorkers = numChunks >= 3 && !R.workers.threads.empty();   // small batches: the caller does everything, no wake-up
  if(useWorkers) R.workers.start(workerJob);
  int status = 0;
  for(int c = 0; c < numChunks && !status; c++) {
    const int r0 = start[c], rows = start[c + 1] - r0;
    if(!useWorkers) {
      if(nextGather.fetch_add(1) == c) { gather(r0, rows); gathered[c].store(1); }
    }
    while(!gathered[c].load(std::memory_order_acquire)) std::this_thread::yield();
    auto enqueue = [&]() -> int {
      if(rows <= 0) return 0;
      KC_CUDA(cudaMemcpyAsync(h->d_raw + (size_t)r0 * 15 * HW, R.spatial + (size_t)r0 * 15 * HW, (size_t)rows * rowBytes, cudaMemcpyHostToDevice, h->h2dStream));
      KC_CUDA(cudaMemcpyAsync(h->d_rawGlobal + r0, R.global + r0, (size_t)rows * 4, cudaMemcpyHostToDevice, h->h2dStream));
      if(symmetry) KC_CUDA(cudaMemcpyAsync(h->d_sym + r0, R.sym + r0, (size_t)rows, cudaMemcpyHostToDevice, h->h2dStream));
      KC_CUDA(cudaEventRecord(h->chunkEvents[3 * c], h->h2dStream));
      KC_CUDA(cudaStreamWaitEvent(st, h->chunkEvents[3 * c], 0));
      if(convertInputToTiles(h, rows, rawNHWC, symmetry ? h->d_sym : nullptr, st, r0)) return 1;
      if(runTrunkBf16(h, rows, st, symmetry ? h->d_sym : nullptr, r0)) return 1;
      KC_CUDA(cudaEventRecord(h->chunkEvents[3 * c + 1], st));
      KC_CUDA(cudaStreamWaitEvent(h->d2hStream, h->chunkEvents[3 * c + 1], 0));
      KC_CUDA(cudaMemcpyAsync(R.policy + (size_t)r0 * 4 * HW, h->d_policy + (size_t)r0 * 4 * HW, (size_t)rows * 4 * HW * 4, cudaMemcpyDeviceToHost, h->d2hStream));
      KC_CUDA(cudaMemcpyAsync(R.value + (size_t)r0 * 2, h->d_value + (size_t)r0 * 2, (size_t)rows * 8, cudaMemcpyDeviceToHost, h->d2hStream));
      KC_CUDA(cudaMemcpyAsync(R.misc + (size_t)r0 * 2, h->d_misc + (size_t)r0 * 2, (size_t)rows * 8, cudaMemcpyDeviceToHost, h->d2hStream));
      if(ownerRows) KC_CUDA(cudaMemcpyAsync(R.own + (size_t)r0 * HW, h->d_own + (size_t)r0 * HW, (size_t)rows * HW * 4, cudaMemcpyDeviceToHost, h->d2hStream));
      KC_CUDA(cudaEventRecord(h->chunkEvents[3 * c + 2], h->d2hStream));
      return 0;
    };
    status = enqueue();
    if(status) failed.store(1);
    else enqueued[c].store(1, std::memory_order_release);
  }
  if(!status) scatterLoop();
  if(useWorkers) R.workers.finish();
  h->lastN = n;
  if(status) return 1;
  KC_CHECK(!failed.load(), "kc_forward_rows: a device copy failed");
  KC_CUDA(cudaStreamSynchronize(h->d2hStream));
  KC_CUDA(cudaStreamSynchronize(st));
  return checkTrunkAbort(h);
}

// ---- layer hooks (nninterface.h:127-169) ---------------------------------------------------------
static int layoutIn(const float* host, float** d_nhwc, int n, int c, int HW, int useNHWC) {
  float* raw = nullptr;
  size_t bytes = (size_t)n * c * HW * 4;
  KC_CUDA(cudaMalloc(&raw, bytes));
  KC_CUDA(cudaMemcpy(raw, host, bytes, cudaMemcpyHostToDevice));
  if(useNHWC) { *d_nhwc = raw; return 0; }
  KC_CUDA(cudaMalloc(d_nhwc, bytes));
  k_nchw_to_nhwc<<<blocksFor((long long)n * c * HW), 256>>>(raw, *d_nhwc, n, c, HW);
  KC_CUDA(cudaDeviceSynchronize());
  cudaFree(raw);
  return 0;
}
static int layoutOut(float* d_nhwc, float* host, int n, int c, int HW, int useNHWC) {
  size_t bytes = (size_t)n * c * HW * 4;
  if(useNHWC) { KC_CUDA(cudaMemcpy(host, d_nhwc, bytes, cudaMemcpyDeviceToHost)); return 0; }
  float* tmp = nullptr;
  KC_CUDA(cudaMalloc(&tmp, bytes));
  k_nhwc_to_nchw<<<blocksFor((long long)n * c * HW), 256>>>(d_nhwc, tmp, n, c, HW);
  KC_CUDA(cudaMemcpy(host, tmp, bytes, cudaMemcpyDeviceToHost));
  cudaFree(tmp);
  return 0;
}

int kc_test_conv(kc_ctx* ctx, const kc_conv_desc* d, int n, int xLen, int yLen, int useNHWC, const float* in, float* out) {
  KC_CHECK(ctx && d && in && out && n > 0, "kc_test_conv: bad argument");
  KC_CUDA(cudaSetDevice(ctx->device));
  ConvW c;
  if(initConv(c, *d, "test")) return 1;
  int HW = xLen * yLen;
  float *din = nullptr, *dout = nullptr;
  if(layoutIn(in, &din, n, c.ic, HW, useNHWC)) return 1;
  KC_CUDA(cudaMalloc(&dout, (size_t)n * HW * c.oc * 4));
  k_conv<<<blocksFor((long long)n * HW * c.oc), 256>>>(din, c.d_tap, dout, n, yLen, xLen, c.ic, c.oc, c.ky, c.kx, 0);
  KC_CUDA(cudaGetLastError());
  int rc = layoutOut(dout, out, n, c.oc, HW, useNHWC);
  cudaFree(din); cudaFree(dout); freeConv(c);
  return rc;
}

int kc_test_batchnorm(kc_ctx* ctx, const kc_bn_desc* d, int activation, int n, int xLen, int yLen, int useNHWC,
                      const float* in, const float* mask, float* out) {
  KC_CHECK(ctx && d && in && out && n > 0, "kc_test_batchnorm: bad argument");
  KC_CUDA(cudaSetDevice(ctx->device));
  BNW b;
  if(initBN(b, *d, activation, "test")) return 1;
  int HW = xLen * yLen;
  float *din = nullptr, *dmask = nullptr;
  if(layoutIn(in, &din, n, b.c, HW, useNHWC)) return 1;
  if(mask) { KC_CUDA(cudaMalloc(&dmask, (size_t)n * HW * 4)); KC_CUDA(cudaMemcpy(dmask, mask, (size_t)n * HW * 4, cudaMemcpyHostToDevice)); }
  k_bn<<<blocksFor((long long)n * HW * b.c), 256>>>(din, b.d_scale, b.d_bias, dmask, din, (long long)n * HW, b.c, b.act);
  KC_CUDA(cudaGetLastError());
  int rc = layoutOut(din, out, n, b.c, HW, useNHWC);
  cudaFree(din); cudaFree(dmask); freeBN(b);
  return rc;
}

int kc_test_resblock(kc_ctx* ctx, const kc_block_desc* d, int n, int xLen, int yLen, int useNHWC,
                     const float* in, const float* mask, float* out) {
  KC_CHECK(ctx && d && in && out && mask && n > 0, "kc_test_resblock: bad argument");
  KC_CUDA(cudaSetDevice(ctx->device));
  BlockW b;
  if(initBlock(b, *d)) return 1;
  int HW = xLen * yLen, C = b.preBN.c;
  int maxC = std::max(C, std::max(b.regularConv.oc, b.gpoolConv.oc));
  float *trunk = nullptr, *dmask = nullptr, *dsum = nullptr;
  if(layoutIn(in, &trunk, n, C, HW, useNHWC)) return 1;
  KC_CUDA(cudaMalloc(&dmask, (size_t)n * HW * 4)); KC_CUDA(cudaMemcpy(dmask, mask, (size_t)n * HW * 4, cudaMemcpyHostToDevice));
  KC_CUDA(cudaMalloc(&dsum, (size_t)n * 4));
  k_mask_sum<<<blocksFor(n), 256>>>(dmask, dsum, n, HW);
  Fp32Scratch s{};
  size_t big = (size_t)n * HW * maxC * 4;
  KC_CUDA(cudaMalloc(&s.a, big)); KC_CUDA(cudaMalloc(&s.b, big)); KC_CUDA(cudaMalloc(&s.c, big));
  KC_CUDA(cudaMalloc(&s.pool, (size_t)n * 3 * maxC * 4)); KC_CUDA(cudaMalloc(&s.bias, (size_t)n * maxC * 4));
  int64_t launches = 0;
  int rc = runBlock(b, n, yLen, xLen, trunk, dmask, dsum, s, 0, launches);
  if(!rc) rc = layoutOut(trunk, out, n, C, HW, useNHWC);
  cudaFree(trunk); cudaFree(dmask); cudaFree(dsum); cudaFree(s.a); cudaFree(s.b); cudaFree(s.c); cudaFree(s.pool); cudaFree(s.bias);
  freeBlock(b);
  return rc;
}

}  // extern "C"
