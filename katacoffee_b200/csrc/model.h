// Device-side model: folded, re-laid-out copies of a kc_model_desc for both net paths.
#pragma once
#include <vector>

#include "kc_internal.h"

namespace kc {

struct ConvW {
  int ky = 0, kx = 0, ic = 0, oc = 0;
  std::vector<float> h;      // host copy, oc,ic,y,x (desc.cpp:131-152)
  float* d_tap = nullptr;    // device fp32 [ky*kx][ic][oc] for the check path
};
struct BNW {
  int c = 0, act = 1;
  std::vector<float> scale, bias;   // folded: scale/sqrt(var+eps), bias - mean*scale' (eigenbackend.cpp:707-710)
  float* d_scale = nullptr; float* d_bias = nullptr;
};
struct MatW {
  int ic = 0, oc = 0;
  std::vector<float> h;      // ic,oc (desc.cpp:284-299)
  float* d = nullptr;
};
struct BiasW {
  int c = 0;
  std::vector<float> h;
  float* d = nullptr;
};
struct BlockW {
  int kind = 0;
  BNW preBN, gpoolBN, midBN;
  ConvW regularConv, gpoolConv, finalConv;
  MatW gpoolToBias;
};

// Packed program for the bf16 tcgen05 trunk kernel (built by net_bf16.cu)
struct TrunkProgram;

}  // namespace kc

struct kc_model {
  kc_ctx* ctx = nullptr;
  int numInputChannels = 0, numInputGlobalChannels = 0, trunkC = 0;
  kc::ConvW initialConv;
  kc::MatW initialMatMul;
  std::vector<kc::BlockW> blocks;
  kc::BNW trunkTipBN;
  kc::ConvW p1Conv, g1Conv, p2Conv;
  kc::BNW g1BN, p1BN;
  kc::MatW gpoolToBiasMul;
  kc::ConvW v1Conv, vOwnershipConv;
  kc::BNW v1BN;
  kc::MatW v2Mul, v3Mul, sv3Mul;
  kc::BiasW v2Bias, v3Bias, sv3Bias;
  int v2Act = 1;
  kc::TrunkProgram* trunk = nullptr;   // null if the shapes are outside what the tcgen05 kernel supports
  std::string trunkUnsupportedWhy;
};

namespace kc {
// net_bf16.cu
int buildTrunkProgram(kc_model* m);
void freeTrunkProgram(kc_model* m);
}  // namespace kc
