// ORACLE (test infrastructure only, see kc_oracle.h): the residual policy/value net forward.
//
// fp32 CPU restatement of what the reference's Eigen backend computes
//   cpp/neuralnet/eigenbackend.cpp:113-186   mask sum, per-(n,c) bias, gpool, value-head pool
//   cpp/neuralnet/eigenbackend.cpp:270-680   ConvLayer (Winograd F(4x4,3x3) / F(2x2,5x5) + GEMM, im2col 1x1)
//   cpp/neuralnet/eigenbackend.cpp:684-734   BatchNormLayer (merged scale/bias, activation, mask)
//   cpp/neuralnet/eigenbackend.cpp:888-1015  ResidualBlock, GlobalPoolingResidualBlock
//   cpp/neuralnet/eigenbackend.cpp:1169-1377 Trunk, PolicyHead, ValueHead
//   cpp/neuralnet/eigenbackend.cpp:1675-1844 getOutput (symmetry in, inverse symmetry out)
//   cpp/neuralnet/nneval.cpp:702-815         post-processing
// with the Coffee head shapes of SURVEY.md 8.1-H (4 policy channels, 2 value, 2 misc, 1 ownership).
// All tensors are NHWC internally like the Eigen backend.  mode 0 = direct convolution (the
// checker: summation in tap-major/channel-minor order), mode 1 = Winograd + 36 GEMMs (the
// algorithm the Eigen backend runs; used as the timed CPU baseline).  Eigen itself is an
// un-vendored dependency that is not in this image; nothing here is Eigen code.
// Pinned by the literal layer vectors of cpp/tests/testnn.cpp (tests/golden/nn_layers_golden.json).
#include "kc_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr int ACT_IDENTITY = 0, ACT_RELU = 1, ACT_MISH = 2;  // activations.h:4-6

struct Conv {
  int ky = 0, kx = 0, ic = 0, oc = 0;
  std::vector<float> w;       // oc,ic,y,x (desc.cpp:131-152)
  std::vector<float> wTap;    // [ky*kx][ic][oc] for the direct path
  std::vector<float> wWino;   // [36][ic][oc]    for the Winograd path (3x3 only)
  void init(const ko_conv_desc& d) {
    ky = d.convYSize; kx = d.convXSize; ic = d.inChannels; oc = d.outChannels;
    size_t n = (size_t)ky * kx * ic * oc;
    w.assign(d.weights, d.weights + n);
    wTap.resize(n);
    for(int o = 0; o < oc; o++)
      for(int i = 0; i < ic; i++)
        for(int y = 0; y < ky; y++)
          for(int x = 0; x < kx; x++)
            wTap[((size_t)(y * kx + x) * ic + i) * oc + o] = w[(((size_t)o * ic + i) * ky + y) * kx + x];
    if(ky == 3 && kx == 3) {
      // G g G^T with the F(4x4,3x3) kernel transform G (6x3); eigenbackend.cpp:322-329
      static const double G[6][3] = {{0.25, 0, 0},
                                     {-1.0 / 6, -1.0 / 6, -1.0 / 6},
                                     {-1.0 / 6, 1.0 / 6, -1.0 / 6},
                                     {1.0 / 24, 2.0 / 24, 4.0 / 24},
                                     {1.0 / 24, -2.0 / 24, 4.0 / 24},
                                     {0, 0, 1}};
      wWino.assign((size_t)36 * ic * oc, 0.0f);
      for(int o = 0; o < oc; o++)
        for(int i = 0; i < ic; i++) {
          const float* g = &w[((size_t)o * ic + i) * 9];
          float tmp[3][6];  // rows transformed along x first, in float like the reference
          for(int y = 0; y < 3; y++)
            for(int a = 0; a < 6; a++)
              tmp[y][a] = (float)(G[a][0] * g[y * 3 + 0] + G[a][1] * g[y * 3 + 1] + G[a][2] * g[y * 3 + 2]);
          for(int b = 0; b < 6; b++)
            for(int a = 0; a < 6; a++) {
              float v = (float)(G[b][0] * tmp[0][a] + G[b][1] * tmp[1][a] + G[b][2] * tmp[2][a]);
              wWino[((size_t)(b * 6 + a) * ic + i) * oc + o] = v;
            }
        }
    }
  }
};

struct BN {
  int c = 0;
  std::vector<float> scale, bias;  // merged: eigenbackend.cpp:707-710
  void init(const ko_bn_desc& d) {
    c = d.numChannels;
    scale.resize(c); bias.resize(c);
    for(int i = 0; i < c; i++) {
      // The backend uses the scale/bias vectors unconditionally (eigenbackend.cpp:707-710); hasScale /
      // hasBias only tell the file parser to fill them with 1 / 0 (desc.cpp:198-214), which a null
      // pointer stands for here.  (testnn.cpp:818-823 relies on this: hasBias=false, bias={0,-2}.)
      float s = d.scale ? d.scale[i] : 1.0f;
      float b = d.bias ? d.bias[i] : 0.0f;
      scale[i] = s / std::sqrt(d.variance[i] + d.epsilon);
      bias[i] = b - scale[i] * d.mean[i];
    }
  }
};

struct MatMul {
  int ic = 0, oc = 0;
  std::vector<float> w;  // ic,oc (desc.cpp:284-299)
  void init(const ko_matmul_desc& d) {
    ic = d.inChannels; oc = d.outChannels;
    w.assign(d.weights, d.weights + (size_t)ic * oc);
  }
  void apply(const float* in, float* out) const {  // one row
    for(int o = 0; o < oc; o++) out[o] = 0.0f;
    for(int i = 0; i < ic; i++) {
      float v = in[i];
      const float* wr = &w[(size_t)i * oc];
      for(int o = 0; o < oc; o++) out[o] += v * wr[o];
    }
  }
};

inline float activate(float x, int act) {
  if(act == ACT_RELU) return x > 0.0f ? x : 0.0f;
  if(act == ACT_MISH) return x * std::tanh(std::log1p(std::exp(std::min(x, 20.0f))) + (std::max(x, 20.0f) - 20.0f));
  return x;
}

// out = mask ? act(x*s+b) : 0   (eigenbackend.cpp:719-732); in/out [n][hw][c], mask [n][hw] or null (=1)
void bnApply(const BN& bn, int act, int n, int hw, const float* in, const float* mask, float* out) {
  int c = bn.c;
  for(int i = 0; i < n * hw; i++) {
    bool on = mask == nullptr || mask[i] == 1.0f;
    const float* x = in + (size_t)i * c;
    float* y = out + (size_t)i * c;
    if(on) for(int k = 0; k < c; k++) y[k] = activate(x[k] * bn.scale[k] + bn.bias[k], act);
    else for(int k = 0; k < c; k++) y[k] = 0.0f;
  }
}

// Direct cross-correlation, zero padding: out[y,x,oc] (+)= sum w[oc,ic,dy,dx] in[y+dy-cy, x+dx-cx, ic]
void convDirect(const Conv& cv, int n, int H, int W, const float* in, float* out, bool accumulate) {
  int ic = cv.ic, oc = cv.oc, cy = cv.ky / 2, cx = cv.kx / 2;
  std::vector<float> acc(oc);
  for(int b = 0; b < n; b++)
    for(int y = 0; y < H; y++)
      for(int x = 0; x < W; x++) {
        std::fill(acc.begin(), acc.end(), 0.0f);
        for(int dy = 0; dy < cv.ky; dy++) {
          int yy = y + dy - cy;
          if(yy < 0 || yy >= H) continue;
          for(int dx = 0; dx < cv.kx; dx++) {
            int xx = x + dx - cx;
            if(xx < 0 || xx >= W) continue;
            const float* ip = in + ((size_t)(b * H + yy) * W + xx) * ic;
            const float* wp = &cv.wTap[(size_t)(dy * cv.kx + dx) * ic * oc];
            for(int i = 0; i < ic; i++) {
              float v = ip[i];
              if(v == 0.0f) continue;
              const float* wr = wp + (size_t)i * oc;
              for(int o = 0; o < oc; o++) acc[o] += v * wr[o];
            }
          }
        }
        float* op = out + ((size_t)(b * H + y) * W + x) * oc;
        if(accumulate) for(int o = 0; o < oc; o++) op[o] += acc[o];
        else for(int o = 0; o < oc; o++) op[o] = acc[o];
      }
}

// C[M][N] = A[M][K] * B[K][N], row-major, plain blocked loops the compiler vectorises over N.
void sgemm(int M, int N, int K, const float* A, const float* B, float* C) {
  constexpr int MB = 4;
  int m = 0;
  for(; m + MB <= M; m += MB) {
    float* c0 = C + (size_t)m * N; float* c1 = c0 + N; float* c2 = c1 + N; float* c3 = c2 + N;
    for(int j = 0; j < N; j++) { c0[j] = 0; c1[j] = 0; c2[j] = 0; c3[j] = 0; }
    for(int k = 0; k < K; k++) {
      float a0 = A[(size_t)m * K + k], a1 = A[(size_t)(m + 1) * K + k];
      float a2 = A[(size_t)(m + 2) * K + k], a3 = A[(size_t)(m + 3) * K + k];
      const float* b = B + (size_t)k * N;
      for(int j = 0; j < N; j++) {
        float bv = b[j];
        c0[j] += a0 * bv; c1[j] += a1 * bv; c2[j] += a2 * bv; c3[j] += a3 * bv;
      }
    }
  }
  for(; m < M; m++) {
    float* c0 = C + (size_t)m * N;
    for(int j = 0; j < N; j++) c0[j] = 0;
    for(int k = 0; k < K; k++) {
      float a0 = A[(size_t)m * K + k];
      const float* b = B + (size_t)k * N;
      for(int j = 0; j < N; j++) c0[j] += a0 * b[j];
    }
  }
}

// Winograd F(4x4,3x3): V = B^T d B per 6x6 input tile, 36 GEMMs [tiles x ic]*[ic x oc], Y = A^T M A.
// (eigenbackend.cpp:425-667 structure; B^T and A^T are the standard Lavin matrices it hard-codes.)
void convWinograd3x3(const Conv& cv, int n, int H, int W, const float* in, float* out, bool accumulate) {
  static const float BT[6][6] = {{4, 0, -5, 0, 1, 0},  {0, -4, -4, 1, 1, 0}, {0, 4, -4, -1, 1, 0},
                                 {0, -2, -1, 2, 1, 0}, {0, 2, -1, -2, 1, 0}, {0, 4, 0, -5, 0, 1}};
  static const float AT[4][6] = {{1, 1, 1, 1, 1, 0}, {0, 1, -1, 2, -2, 0}, {0, 1, 1, 4, 4, 0}, {0, 1, -1, 8, -8, 1}};
  const int ic = cv.ic, oc = cv.oc;
  const int tx = (W + 3) / 4, ty = (H + 3) / 4;
  const int T = n * ty * tx;
  std::vector<float> V((size_t)36 * T * ic), M((size_t)36 * T * oc);
  std::vector<float> d((size_t)36 * ic), t((size_t)36 * ic);
  for(int b = 0; b < n; b++)
    for(int yt = 0; yt < ty; yt++)
      for(int xt = 0; xt < tx; xt++) {
        for(int dy = 0; dy < 6; dy++)
          for(int dx = 0; dx < 6; dx++) {
            int y = yt * 4 + dy - 1, x = xt * 4 + dx - 1;
            float* dst = &d[(size_t)(dy * 6 + dx) * ic];
            if(x < 0 || y < 0 || x >= W || y >= H) std::fill(dst, dst + ic, 0.0f);
            else memcpy(dst, in + ((size_t)(b * H + y) * W + x) * ic, sizeof(float) * ic);
          }
        // rows (along x), then columns (along y)
        for(int dy = 0; dy < 6; dy++)
          for(int a = 0; a < 6; a++) {
            float* dst = &t[(size_t)(dy * 6 + a) * ic];
            std::fill(dst, dst + ic, 0.0f);
            for(int k = 0; k < 6; k++) {
              float c = BT[a][k];
              if(c == 0.0f) continue;
              const float* s = &d[(size_t)(dy * 6 + k) * ic];
              for(int i = 0; i < ic; i++) dst[i] += c * s[i];
            }
          }
        int tile = (b * ty + yt) * tx + xt;
        for(int bb = 0; bb < 6; bb++)
          for(int a = 0; a < 6; a++) {
            float* dst = &V[((size_t)(bb * 6 + a) * T + tile) * ic];
            std::fill(dst, dst + ic, 0.0f);
            for(int k = 0; k < 6; k++) {
              float c = BT[bb][k];
              if(c == 0.0f) continue;
              const float* s = &t[(size_t)(k * 6 + a) * ic];
              for(int i = 0; i < ic; i++) dst[i] += c * s[i];
            }
          }
      }
  for(int s = 0; s < 36; s++)
    sgemm(T, oc, ic, &V[(size_t)s * T * ic], &cv.wWino[(size_t)s * ic * oc], &M[(size_t)s * T * oc]);
  std::vector<float> m((size_t)36 * oc), u((size_t)24 * oc), yv(oc);
  for(int b = 0; b < n; b++)
    for(int yt = 0; yt < ty; yt++)
      for(int xt = 0; xt < tx; xt++) {
        int tile = (b * ty + yt) * tx + xt;
        for(int s = 0; s < 36; s++) memcpy(&m[(size_t)s * oc], &M[((size_t)s * T + tile) * oc], sizeof(float) * oc);
        // u[dy][a] = sum_k AT[a][k] m[dy][k]   (6 rows x 4)
        for(int dy = 0; dy < 6; dy++)
          for(int a = 0; a < 4; a++) {
            float* dst = &u[(size_t)(dy * 4 + a) * oc];
            std::fill(dst, dst + oc, 0.0f);
            for(int k = 0; k < 6; k++) {
              float c = AT[a][k];
              if(c == 0.0f) continue;
              const float* s = &m[(size_t)(dy * 6 + k) * oc];
              for(int o = 0; o < oc; o++) dst[o] += c * s[o];
            }
          }
        for(int bb = 0; bb < 4; bb++)
          for(int a = 0; a < 4; a++) {
            int y = yt * 4 + bb, x = xt * 4 + a;
            if(x >= W || y >= H) continue;
            std::fill(yv.begin(), yv.end(), 0.0f);
            for(int k = 0; k < 6; k++) {
              float c = AT[bb][k];
              if(c == 0.0f) continue;
              const float* s = &u[(size_t)(k * 4 + a) * oc];
              for(int o = 0; o < oc; o++) yv[o] += c * s[o];
            }
            float* op = out + ((size_t)(b * H + y) * W + x) * oc;
            if(accumulate) for(int o = 0; o < oc; o++) op[o] += yv[o];
            else for(int o = 0; o < oc; o++) op[o] = yv[o];
          }
      }
}

// 1x1 as one GEMM (the reference's im2col branch, eigenbackend.cpp:668-678)
void conv1x1Gemm(const Conv& cv, int n, int H, int W, const float* in, float* out, bool accumulate) {
  int rows = n * H * W;
  if(!accumulate) { sgemm(rows, cv.oc, cv.ic, in, cv.wTap.data(), out); return; }
  std::vector<float> tmp((size_t)rows * cv.oc);
  sgemm(rows, cv.oc, cv.ic, in, cv.wTap.data(), tmp.data());
  for(size_t i = 0; i < tmp.size(); i++) out[i] += tmp[i];
}

// round-to-nearest-even to bfloat16 precision (what the tensor-core path stores its operands in)
inline float bf16r(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if((u & 0x7f800000u) == 0x7f800000u) return x;
  u += 0x7fffu + ((u >> 16) & 1u);
  u &= 0xffff0000u;
  memcpy(&x, &u, 4);
  return x;
}

// Reduced-precision emulation ("mode & 15 == 2"): the same arithmetic as mode 0 with the convolution's weights and
// input activations rounded to the tensor-core operand formats first (products are then exact in fp32, accumulation
// stays fp32) -- the precision model of the tcgen05 path, used to tell rounding apart from kernel bugs.
// Operand formats ride in the mode word: bits 4-5 = activations, bits 6-7 = weights; 0 bf16, 1 fp16, 2 fp32 (not rounded).
// So mode 2 = bf16 x bf16, KO_MODE_EMUL(1, 1) = 2 | 16 | 64 = fp16 x fp16 (tcgen05 kind::f16 takes either format at one rate).
inline float f16r(float x) {
  if(!(std::fabs(x) <= 65504.0f)) return x > 0 ? 65504.0f : x < 0 ? -65504.0f : x;   // cvt.rn.satfinite
  return (float)(_Float16)x;
}
inline float roundFmt(float x, int fmt) { return fmt == 0 ? bf16r(x) : fmt == 1 ? f16r(x) : x; }
void convApplyEmul(const Conv& cv, int n, int H, int W, const float* in, float* out, bool accumulate, int mode) {
  const int aFmt = (mode >> 4) & 3, wFmt = (mode >> 6) & 3;
  Conv q = cv;
  for(float& v : q.wTap) v = roundFmt(v, wFmt);
  std::vector<float> a((size_t)n * H * W * cv.ic);
  for(size_t i = 0; i < a.size(); i++) a[i] = roundFmt(in[i], aFmt);
  convDirect(q, n, H, W, a.data(), out, accumulate);
}

void convApply(const Conv& cv, int n, int H, int W, const float* in, float* out, bool accumulate, int mode) {
  if((mode & 15) == 2) convApplyEmul(cv, n, H, W, in, out, accumulate, mode);
  else if(mode == 1 && cv.ky == 3 && cv.kx == 3) convWinograd3x3(cv, n, H, W, in, out, accumulate);
  else if(mode == 1 && cv.ky == 1 && cv.kx == 1) conv1x1Gemm(cv, n, H, W, in, out, accumulate);
  else convDirect(cv, n, H, W, in, out, accumulate);
}

// eigenbackend.cpp:141-166 (in [n][hw][c], out [n][3c])
void gpool(int n, int hw, int c, const float* in, const float* mask, const float* maskSum, float* out) {
  for(int b = 0; b < n; b++)
    for(int k = 0; k < c; k++) {
      float s = 0.0f, m = -1.0f;
      for(int p = 0; p < hw; p++) {
        float x = in[((size_t)b * hw + p) * c + k];
        s += x;
        float mv = mask ? mask[b * hw + p] : 1.0f;
        m = std::max(m, x + (mv - 1.0f));
      }
      float div = maskSum[b], sq = std::sqrt(div), mean = s / div;
      out[(size_t)b * 3 * c + k] = mean;
      out[(size_t)b * 3 * c + c + k] = mean * (sq - 14.0f) * 0.1f;
      out[(size_t)b * 3 * c + 2 * c + k] = m;
    }
}
// eigenbackend.cpp:168-186
void valuePool(int n, int hw, int c, const float* in, const float* maskSum, float* out) {
  for(int b = 0; b < n; b++)
    for(int k = 0; k < c; k++) {
      float s = 0.0f;
      for(int p = 0; p < hw; p++) s += in[((size_t)b * hw + p) * c + k];
      float div = maskSum[b], sq = std::sqrt(div), mean = s / div;
      out[(size_t)b * 3 * c + k] = mean;
      out[(size_t)b * 3 * c + c + k] = mean * (sq - 14.0f) * 0.1f;
      out[(size_t)b * 3 * c + 2 * c + k] = mean * ((sq - 14.0f) * (sq - 14.0f) * 0.01f - 0.1f);
    }
}
void addNCBias(int n, int hw, int c, float* x, const float* bias) {  // eigenbackend.cpp:126-137
  for(int b = 0; b < n; b++)
    for(int p = 0; p < hw; p++)
      for(int k = 0; k < c; k++) x[((size_t)b * hw + p) * c + k] += bias[(size_t)b * c + k];
}

struct Block {
  int kind = 0, preAct = 1, gpoolAct = 1, midAct = 1;
  BN preBN, gpoolBN, midBN;
  Conv regularConv, gpoolConv, finalConv;
  MatMul gpoolToBias;
  void init(const ko_block_desc& d) {
    kind = d.kind; preAct = d.preActivation; gpoolAct = d.gpoolActivation; midAct = d.midActivation;
    preBN.init(d.preBN); regularConv.init(d.regularConv); midBN.init(d.midBN); finalConv.init(d.finalConv);
    if(kind == 2) { gpoolConv.init(d.gpoolConv); gpoolBN.init(d.gpoolBN); gpoolToBias.init(d.gpoolToBiasMul); }
  }
  // trunk += ... (eigenbackend.cpp:912-930, 968-1005)
  void apply(int n, int H, int W, float* trunk, const float* mask, const float* maskSum, int mode) const {
    int hw = H * W;
    std::vector<float> pre((size_t)n * hw * preBN.c);
    bnApply(preBN, preAct, n, hw, trunk, mask, pre.data());
    std::vector<float> reg((size_t)n * hw * regularConv.oc);
    convApply(regularConv, n, H, W, pre.data(), reg.data(), false, mode);
    if(kind == 2) {
      int g = gpoolConv.oc;
      std::vector<float> go((size_t)n * hw * g), go2((size_t)n * hw * g), cat((size_t)n * 3 * g),
        bias((size_t)n * regularConv.oc);
      convApply(gpoolConv, n, H, W, pre.data(), go.data(), false, mode);
      bnApply(gpoolBN, gpoolAct, n, hw, go.data(), mask, go2.data());
      gpool(n, hw, g, go2.data(), mask, maskSum, cat.data());
      for(int b = 0; b < n; b++) gpoolToBias.apply(&cat[(size_t)b * 3 * g], &bias[(size_t)b * regularConv.oc]);
      addNCBias(n, hw, regularConv.oc, reg.data(), bias.data());
    }
    std::vector<float> mid((size_t)n * hw * midBN.c);
    bnApply(midBN, midAct, n, hw, reg.data(), mask, mid.data());
    convApply(finalConv, n, H, W, mid.data(), trunk, true, mode);
  }
};

}  // namespace

struct ko_model {
  int version, numInputChannels, numInputGlobalChannels, numBlocks, trunkC;
  int trunkTipAct, g1Act, p1Act, v1Act, v2Act;
  Conv initialConv; MatMul initialMatMul;
  std::vector<Block> blocks;
  BN trunkTipBN;
  Conv p1Conv, g1Conv, p2Conv; BN g1BN, p1BN; MatMul gpoolToBiasMul;
  Conv v1Conv, vOwnershipConv; BN v1BN; MatMul v2Mul, v3Mul, sv3Mul;
  std::vector<float> v2Bias, v3Bias, sv3Bias;

  void forwardChunk(int n, int H, int W, int inputsNHWC, const float* rowSpatial, const float* rowGlobal,
                    const int8_t* symmetry, float* policy, float* value, float* misc, float* ownership,
                    int mode, float* trace = nullptr) const {
    const int hw = H * W, C = numInputChannels;
    // eigenbackend.cpp:1696-1704: per row, copy global and copyInputsWithSymmetry into the NHWC batch
    std::vector<float> in((size_t)n * hw * C);
    std::vector<float> tmp((size_t)hw * C), tmp2((size_t)hw * C);
    for(int b = 0; b < n; b++) {
      const float* src = rowSpatial + (size_t)b * hw * C;
      int sym = symmetry ? symmetry[b] : 0;
      if(inputsNHWC) ko_copy_inputs_with_symmetry(src, &in[(size_t)b * hw * C], 1, H, W, C, 1, sym);
      else {
        ko_copy_inputs_with_symmetry(src, tmp.data(), 1, H, W, C, 0, sym);
        for(int c = 0; c < C; c++)
          for(int p = 0; p < hw; p++) in[((size_t)b * hw + p) * C + c] = tmp[(size_t)c * hw + p];
      }
    }
    // eigenbackend.cpp:1438-1439: mask = channel 0, maskSum
    std::vector<float> mask((size_t)n * hw), maskSum(n);
    for(int b = 0; b < n; b++) {
      float s = 0;
      for(int p = 0; p < hw; p++) { mask[b * hw + p] = in[((size_t)b * hw + p) * C]; s += mask[b * hw + p]; }
      maskSum[b] = s;
    }
    // trunk (eigenbackend.cpp:1202-1226)
    std::vector<float> trunk((size_t)n * hw * trunkC), gb((size_t)n * trunkC);
    convApply(initialConv, n, H, W, in.data(), trunk.data(), false, mode);
    if((mode & 15) == 2) {   // the tensor-core path folds this matmul into the initial conv as a 16th reduced-precision input channel
      MatMul q = initialMatMul;
      for(float& v : q.w) v = roundFmt(v, (mode >> 6) & 3);
      for(int b = 0; b < n; b++) q.apply(rowGlobal + (size_t)b * numInputGlobalChannels, &gb[(size_t)b * trunkC]);
    } else
    for(int b = 0; b < n; b++) initialMatMul.apply(rowGlobal + (size_t)b * numInputGlobalChannels, &gb[(size_t)b * trunkC]);
    addNCBias(n, hw, trunkC, trunk.data(), gb.data());
    // trace (diagnostic): the trunk after the initial conv, after every block, and the tip, [numBlocks + 2][n][hw][trunkC]
    const size_t tsz = trunk.size();
    if(trace) std::copy(trunk.begin(), trunk.end(), trace);
    for(size_t bi = 0; bi < blocks.size(); bi++) {
      blocks[bi].apply(n, H, W, trunk.data(), mask.data(), maskSum.data(), mode);
      if(trace) std::copy(trunk.begin(), trunk.end(), trace + (bi + 1) * tsz);
    }
    std::vector<float> tip((size_t)n * hw * trunkC);
    bnApply(trunkTipBN, trunkTipAct, n, hw, trunk.data(), mask.data(), tip.data());
    if(trace) std::copy(tip.begin(), tip.end(), trace + (blocks.size() + 1) * tsz);
    // policy head (eigenbackend.cpp:1265-1298), no pass output (ledger H)
    {
      int pc = p1Conv.oc, gc = g1Conv.oc;
      std::vector<float> p1((size_t)n * hw * pc), p12((size_t)n * hw * pc), g1((size_t)n * hw * gc),
        g12((size_t)n * hw * gc), cat((size_t)n * 3 * gc), bias((size_t)n * pc), pol((size_t)n * hw * p2Conv.oc);
      convApply(p1Conv, n, H, W, tip.data(), p1.data(), false, mode);
      convApply(g1Conv, n, H, W, tip.data(), g1.data(), false, mode);
      bnApply(g1BN, g1Act, n, hw, g1.data(), mask.data(), g12.data());
      gpool(n, hw, gc, g12.data(), mask.data(), maskSum.data(), cat.data());
      for(int b = 0; b < n; b++) gpoolToBiasMul.apply(&cat[(size_t)b * 3 * gc], &bias[(size_t)b * pc]);
      addNCBias(n, hw, pc, p1.data(), bias.data());
      bnApply(p1BN, p1Act, n, hw, p1.data(), mask.data(), p12.data());
      convApply(p2Conv, n, H, W, p12.data(), pol.data(), false, (mode & 15) == 2 ? 0 : mode);   // fp32 on CUDA cores in the tensor-core path
      // NHWC [hw][4] -> NNPos order dir*HW + y*W + x, inverse spatial symmetry per direction channel
      // (eigenbackend.cpp:1776: copyOutputsWithSymmetry; ledger H, K parity mode)
      int D = p2Conv.oc;
      std::vector<float> planes((size_t)D * hw);
      for(int b = 0; b < n; b++) {
        for(int d = 0; d < D; d++)
          for(int p = 0; p < hw; p++) planes[(size_t)d * hw + p] = pol[((size_t)b * hw + p) * D + d];
        ko_copy_outputs_with_symmetry(planes.data(), policy + (size_t)b * D * hw, D, H, W, symmetry ? symmetry[b] : 0);
      }
    }
    // value head (eigenbackend.cpp:1341-1376)
    {
      int vc = v1Conv.oc;
      std::vector<float> v1((size_t)n * hw * vc), v12((size_t)n * hw * vc), pool((size_t)n * 3 * vc),
        v2((size_t)n * v2Mul.oc);
      convApply(v1Conv, n, H, W, tip.data(), v1.data(), false, mode);
      bnApply(v1BN, v1Act, n, hw, v1.data(), mask.data(), v12.data());
      valuePool(n, hw, vc, v12.data(), maskSum.data(), pool.data());
      for(int b = 0; b < n; b++) {
        float* o = &v2[(size_t)b * v2Mul.oc];
        v2Mul.apply(&pool[(size_t)b * 3 * vc], o);
        for(int k = 0; k < v2Mul.oc; k++) o[k] = activate(o[k] + v2Bias[k], v2Act);
        float* vo = value + (size_t)b * v3Mul.oc;
        v3Mul.apply(o, vo);
        for(int k = 0; k < v3Mul.oc; k++) vo[k] += v3Bias[k];
        float* mo = misc + (size_t)b * sv3Mul.oc;
        sv3Mul.apply(o, mo);
        for(int k = 0; k < sv3Mul.oc; k++) mo[k] += sv3Bias[k];
      }
      if(ownership) {
        std::vector<float> own((size_t)n * hw);
        convApply(vOwnershipConv, n, H, W, v12.data(), own.data(), false, (mode & 15) == 2 ? 0 : mode);
        for(int b = 0; b < n; b++)
          ko_copy_outputs_with_symmetry(&own[(size_t)b * hw], ownership + (size_t)b * hw, 1, H, W,
                                        symmetry ? symmetry[b] : 0);
      }
    }
  }
};

extern "C" {

ko_model* ko_model_create(const ko_model_desc* d) {
  ko_model* m = new ko_model();
  m->version = d->version;
  m->numInputChannels = d->numInputChannels;
  m->numInputGlobalChannels = d->numInputGlobalChannels;
  m->numBlocks = d->numBlocks;
  m->trunkC = d->trunkNumChannels;
  m->trunkTipAct = d->trunkTipActivation; m->g1Act = d->g1Activation; m->p1Act = d->p1Activation;
  m->v1Act = d->v1Activation; m->v2Act = d->v2Activation;
  m->initialConv.init(d->initialConv);
  m->initialMatMul.init(d->initialMatMul);
  m->blocks.resize(d->numBlocks);
  for(int i = 0; i < d->numBlocks; i++) m->blocks[i].init(d->blocks[i]);
  m->trunkTipBN.init(d->trunkTipBN);
  m->p1Conv.init(d->p1Conv); m->g1Conv.init(d->g1Conv); m->g1BN.init(d->g1BN);
  m->gpoolToBiasMul.init(d->gpoolToBiasMul); m->p1BN.init(d->p1BN); m->p2Conv.init(d->p2Conv);
  m->v1Conv.init(d->v1Conv); m->v1BN.init(d->v1BN);
  m->v2Mul.init(d->v2Mul); m->v3Mul.init(d->v3Mul); m->sv3Mul.init(d->sv3Mul);
  m->v2Bias.assign(d->v2Bias.weights, d->v2Bias.weights + d->v2Bias.numChannels);
  m->v3Bias.assign(d->v3Bias.weights, d->v3Bias.weights + d->v3Bias.numChannels);
  m->sv3Bias.assign(d->sv3Bias.weights, d->sv3Bias.weights + d->sv3Bias.numChannels);
  m->vOwnershipConv.init(d->vOwnershipConv);
  return m;
}
void ko_model_destroy(ko_model* m) { delete m; }

void ko_model_forward(const ko_model* m, int n, int nnXLen, int nnYLen, int inputsNHWC,
                      const float* rowSpatial, const float* rowGlobal, const int8_t* symmetry,
                      float* policy, float* value, float* misc, float* ownership, int mode, int threads) {
  const int H = nnYLen, W = nnXLen, hw = H * W, C = m->numInputChannels;
  if(threads < 1) threads = 1;
  // The reference runs Eigen with batches of nnMaxBatchSize = 4 per server thread (setup.cpp:299);
  // chunks of 4 rows are distributed over the threads.
  const int chunk = 4;
  int numChunks = (n + chunk - 1) / chunk;
  auto work = [&](int t) {
    for(int ci = t; ci < numChunks; ci += threads) {
      int b0 = ci * chunk, nb = std::min(chunk, n - b0);
      m->forwardChunk(nb, H, W, inputsNHWC, rowSpatial + (size_t)b0 * hw * C,
                      rowGlobal + (size_t)b0 * m->numInputGlobalChannels, symmetry ? symmetry + b0 : nullptr,
                      policy + (size_t)b0 * m->p2Conv.oc * hw, value + (size_t)b0 * m->v3Mul.oc,
                      misc + (size_t)b0 * m->sv3Mul.oc, ownership ? ownership + (size_t)b0 * hw : nullptr, mode);
    }
  };
  if(threads == 1) { work(0); return; }
  std::vector<std::thread> th;
  for(int t = 0; t < threads; t++) th.emplace_back(work, t);
  for(auto& t : th) t.join();
}

void ko_model_forward_trace(const ko_model* m, int n, int nnXLen, int nnYLen, const float* rowSpatial, const float* rowGlobal,
                            float* policy, float* value, float* misc, float* ownership, int mode, float* trace) {
  m->forwardChunk(n, nnYLen, nnXLen, 0, rowSpatial, rowGlobal, nullptr, policy, value, misc, ownership, mode, trace);
}

static void toNHWC(const float* in, float* out, int n, int c, int hw) {
  for(int b = 0; b < n; b++)
    for(int k = 0; k < c; k++)
      for(int p = 0; p < hw; p++) out[((size_t)b * hw + p) * c + k] = in[((size_t)b * c + k) * hw + p];
}
static void toNCHW(const float* in, float* out, int n, int c, int hw) {
  for(int b = 0; b < n; b++)
    for(int k = 0; k < c; k++)
      for(int p = 0; p < hw; p++) out[((size_t)b * c + k) * hw + p] = in[((size_t)b * hw + p) * c + k];
}

void ko_test_conv(const ko_conv_desc* d, int n, int xLen, int yLen, int useNHWC, const float* in, float* out, int mode) {
  Conv cv; cv.init(*d);
  int hw = xLen * yLen;
  std::vector<float> a((size_t)n * hw * cv.ic), o((size_t)n * hw * cv.oc);
  if(useNHWC) memcpy(a.data(), in, a.size() * sizeof(float)); else toNHWC(in, a.data(), n, cv.ic, hw);
  convApply(cv, n, yLen, xLen, a.data(), o.data(), false, mode);
  if(useNHWC) memcpy(out, o.data(), o.size() * sizeof(float)); else toNCHW(o.data(), out, n, cv.oc, hw);
}
void ko_test_batchnorm(const ko_bn_desc* d, int activation, int n, int xLen, int yLen, int useNHWC,
                       const float* in, const float* mask, float* out) {
  BN bn; bn.init(*d);
  int hw = xLen * yLen;
  std::vector<float> a((size_t)n * hw * bn.c), o((size_t)n * hw * bn.c);
  if(useNHWC) memcpy(a.data(), in, a.size() * sizeof(float)); else toNHWC(in, a.data(), n, bn.c, hw);
  bnApply(bn, activation, n, hw, a.data(), mask, o.data());
  if(useNHWC) memcpy(out, o.data(), o.size() * sizeof(float)); else toNCHW(o.data(), out, n, bn.c, hw);
}
void ko_test_resblock(const ko_block_desc* d, int n, int xLen, int yLen, int useNHWC, const float* in,
                      const float* mask, float* out, int mode) {
  Block b; b.init(*d);
  int hw = xLen * yLen, c = b.preBN.c;
  std::vector<float> t((size_t)n * hw * c), ms(n);
  if(useNHWC) memcpy(t.data(), in, t.size() * sizeof(float)); else toNHWC(in, t.data(), n, c, hw);
  for(int i = 0; i < n; i++) { float s = 0; for(int p = 0; p < hw; p++) s += mask[i * hw + p]; ms[i] = s; }
  b.apply(n, yLen, xLen, t.data(), mask, ms.data(), mode);
  if(useNHWC) memcpy(out, t.data(), t.size() * sizeof(float)); else toNCHW(t.data(), out, n, c, hw);
}

// nneval.cpp:702-815 for one row.  policy: logits in, probabilities out (illegal = -1);
// value2: (win, loss) logits in -> (whiteWinProb, whiteLossProb); misc2: (varTimeLeft,
// shorttermWinlossError) pre-activation in -> post-processed (desc.cpp:956-963 multipliers).
void ko_postprocess(float* policy, int policySize, const uint32_t* legalMask, float policyTemp,
                    float* value2, float* misc2, int nextPla) {
  float maxP = -1e25f;
  int legalCount = 0;
  float invT = 1.0f / policyTemp;
  auto isLegal = [&](int i) { return ((legalMask[i >> 5] >> (i & 31)) & 1u) != 0; };
  for(int i = 0; i < policySize; i++) {
    float v;
    if(isLegal(i)) { legalCount++; v = policy[i] * invT; }
    else v = -1e30f;
    policy[i] = v;
    if(v > maxP) maxP = v;
  }
  float sum = 0.0f;
  for(int i = 0; i < policySize; i++) { policy[i] = std::exp(policy[i] - maxP); sum += policy[i]; }
  if(sum <= 0.0f) {
    float uniform = legalCount > 0 ? 1.0f / legalCount : 0.0f;
    for(int i = 0; i < policySize; i++) policy[i] = isLegal(i) ? uniform : -1.0f;
  } else {
    for(int i = 0; i < policySize; i++) policy[i] = isLegal(i) ? policy[i] / sum : -1.0f;
  }
  // 2-way softmax of win/loss logits, softplus heads, flip to white's view (nneval.cpp:769-815)
  double w = value2[0], l = value2[1];
  double mx = std::max(w, l);
  double winProb = std::exp(w - mx), lossProb = std::exp(l - mx);
  double probSum = winProb + lossProb;
  winProb /= probSum; lossProb /= probSum;
  auto softPlus = [](double x) { return x > 40.0 ? x : std::log(1.0 + std::exp(x)); };  // nneval.cpp:580-586
  double varTimeLeft = softPlus(misc2[0]) * 40.0;   // varianceTimeMultiplier, desc.cpp:956-963
  double s = softPlus(misc2[1] * 0.5);
  double stErr = std::sqrt(s * s * 0.25);          // shorttermValueErrorMultiplier
  if(nextPla == 2) { value2[0] = (float)winProb; value2[1] = (float)lossProb; }
  else { value2[0] = (float)lossProb; value2[1] = (float)winProb; }
  misc2[0] = (float)varTimeLeft;
  misc2[1] = (float)stErr;
}

}  // extern "C"
