// Model files: the KataGo model format as the reference's parser reads it, for Coffee nets.
//
// Replaces, for the standalone backend, ModelDesc::loadFromFileMaybeGZipped (cpp/neuralnet/desc.cpp:1146-1204), the
// ModelDesc / TrunkDesc / PolicyHeadDesc / ValueHeadDesc constructors (:977-1094, :648-696, :751-810, :844-925), the block
// stack (:562-640), the layer parsers (conv :107-155 with the y,x,ic,oc -> oc,ic,y,x re-layout, batch norm :175-219,
// activation :239-258, matmul :274-302, matbias :318-340), readFloats incl. the "@BIN@" little-endian blocks (:27-92) and
// FileUtils::loadFileIntoString / uncompressAndLoadFileIntoString (cpp/core/fileutils.cpp:113-215: the SHA-256 is taken over
// the file as stored, i.e. over the compressed bytes of a .gz).
//
// Canonical readings:
//  * version must be 1 (cpp/neuralnet/modelversion.h:7-12).  As written desc.cpp:989-998 rejects every version (< 3 and > 1).
//  * an activation layer is its name, optionally followed by ACTIVATION_IDENTITY / ACTIVATION_RELU / ACTIVATION_MISH (the
//    reference reads the kind only for version >= 11, desc.cpp:243-256; without it the activation is ReLU).  The writer emits
//    ReLU layers without the token (= the reference's version-1 format) and the token only for identity / Mish (an extension).
//  * sizes from the file are bounded (channels <= 8192, filters <= 9, floats <= what the rest of the file can hold) before any
//    allocation, and every exception is turned into an error return at the C ABI.
//  * Coffee head shapes: p2Conv 4 output channels (one per direction); gpoolToPassMul is part of the format and is parsed,
//    checked (inChannels = 3 x g1 channels) and ignored -- Coffee has no pass; v3 2 outputs (win, loss); sv3 2 outputs
//    (varTimeLeft, shorttermWinlossError); ownership 1 channel.
//  * nested bottleneck blocks are rejected (not supported by either device path).
// No GPU is needed for anything in this file.
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <memory>

#include "kc_internal.h"

namespace {

struct ParseError { std::string msg; };

// ---- SHA-256 (FIPS 180-4) of the file bytes, hex ----
std::string sha256Hex(const std::string& data) {
  static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
    0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
    0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
    0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
    0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
    0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  std::string msg = data;
  const uint64_t bitLen = (uint64_t)data.size() * 8;
  msg.push_back((char)0x80);
  while(msg.size() % 64 != 56) msg.push_back('\0');
  for(int i = 7; i >= 0; i--) msg.push_back((char)((bitLen >> (8 * i)) & 0xff));
  auto rotr = [](uint32_t x, int r) { return (x >> r) | (x << (32 - r)); };
  for(size_t off = 0; off < msg.size(); off += 64) {
    uint32_t w[64];
    for(int i = 0; i < 16; i++) {
      const unsigned char* p = (const unsigned char*)msg.data() + off + 4 * i;
      w[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
    }
    for(int i = 16; i < 64; i++) {
      const uint32_t s0 = rotr(w[i - 15], 7) ^ rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
      const uint32_t s1 = rotr(w[i - 2], 17) ^ rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for(int i = 0; i < 64; i++) {
      const uint32_t S1 = rotr(e, 6) ^ rotr(e, 11) ^ rotr(e, 25), ch = (e & f) ^ (~e & g);
      const uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      const uint32_t S0 = rotr(a, 2) ^ rotr(a, 13) ^ rotr(a, 22), maj = (a & b) ^ (a & c) ^ (b & c);
      const uint32_t t2 = S0 + maj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  char buf[65];
  for(int i = 0; i < 8; i++) snprintf(buf + 8 * i, 9, "%08x", h[i]);
  return std::string(buf, 64);
}

std::string toLower(std::string s) { for(char& c : s) c = (char)tolower((unsigned char)c); return s; }
bool isSuffix(const std::string& s, const std::string& suf) { return s.size() >= suf.size() && s.compare(s.size() - suf.size(), suf.size(), suf) == 0; }

std::string gunzip(const std::string& in, const std::string& file) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if(inflateInit2(&zs, 15 + 32) != Z_OK) throw ParseError{"Error while ungzipping file. Invalid file? File: " + file};
  std::string out;
  zs.next_in = (Bytef*)in.data();
  zs.avail_in = (uInt)in.size();
  std::vector<char> chunk(1 << 18);
  int ret = Z_OK;
  while(ret != Z_STREAM_END) {
    zs.next_out = (Bytef*)chunk.data();
    zs.avail_out = (uInt)chunk.size();
    ret = inflate(&zs, Z_NO_FLUSH);
    if(ret != Z_OK && ret != Z_STREAM_END) { inflateEnd(&zs); throw ParseError{"Error while ungzipping file, zlib error " + std::to_string(ret) + ". Invalid file? File: " + file}; }
    out.append(chunk.data(), chunk.size() - zs.avail_out);
    if(ret == Z_OK && zs.avail_in == 0 && zs.avail_out != 0) { inflateEnd(&zs); throw ParseError{"Error while ungzipping file: truncated. File: " + file}; }
  }
  inflateEnd(&zs);
  return out;
}

// ---- the parsed model: owns every weight array; `desc` points into them ----
struct Layers {
  std::vector<std::unique_ptr<std::vector<float>>> arrays;
  const float* keep(std::vector<float>&& v) { arrays.emplace_back(new std::vector<float>(std::move(v))); return arrays.back()->data(); }
};

struct Cursor {
  const std::string& s;
  size_t pos = 0;
  bool binary;
  Cursor(const std::string& str, bool bin) : s(str), binary(bin) {}
  void skipWs() { while(pos < s.size() && isspace((unsigned char)s[pos])) pos++; }
  std::string token(const std::string& what) {
    skipWs();
    const size_t b = pos;
    while(pos < s.size() && !isspace((unsigned char)s[pos])) pos++;
    if(b == pos) throw ParseError{what + ": unexpected end of model file"};
    return s.substr(b, pos - b);
  }
  bool peekStartsWith(const char* prefix) {
    skipWs();
    return s.compare(pos, strlen(prefix), prefix) == 0;
  }
  int integer(const std::string& what) {
    const std::string t = token(what);
    char* end = nullptr;
    const long v = strtol(t.c_str(), &end, 10);
    if(end == t.c_str() || *end != '\0') throw ParseError{what + ": expected an integer, found '" + t + "'"};
    return (int)v;
  }
  float real(const std::string& what) {
    const std::string t = token(what);
    char* end = nullptr;
    const float v = strtof(t.c_str(), &end);
    if(end == t.c_str()) throw ParseError{what + ": expected a number, found '" + t + "'"};
    return v;
  }
  // readFloats (desc.cpp:37-92)
  std::vector<float> floats(size_t n, const std::string& name) {
    // a header can claim any size: nothing is allocated that the rest of the file could not fill (a binary float takes 4 bytes, a text
    // float at least 2 characters incl. its separator)
    const size_t remaining = s.size() - std::min(pos, s.size());
    if(n > remaining / (binary ? 4 : 2) + 1) throw ParseError{name + ": did not find the expected number of floats (the file is too short for " + std::to_string(n) + " weights)"};
    std::vector<float> buf(n);
    if(!binary) {
      for(size_t i = 0; i < n; i++) {
        skipWs();
        const char* b = s.c_str() + pos;
        char* end = nullptr;
        buf[i] = strtof(b, &end);
        if(end == b) throw ParseError{name + ": could not read float weights. Invalid model - perhaps you are trying to load a .bin.gz model as a .txt.gz model?"};
        pos += (size_t)(end - b);
      }
    } else {
      int before = 0;
      while(true) {
        if(pos >= s.size() || before > 100)
          throw ParseError{name + ": could not read float weights. Invalid model - perhaps you are trying to load a .txt.gz model as a .bin.gz model?"};
        if(s[pos++] == '@') break;
        before++;
      }
      if(s.compare(pos, 4, "BIN@") != 0) throw ParseError{name + ": did not find expected header for binary float block"};
      pos += 4;
      if(n > (s.size() - pos) / 4) throw ParseError{name + ": did not find the expected number of floats in binary float block"};
      const unsigned char* p = (const unsigned char*)s.data() + pos;
      for(size_t i = 0; i < n; i++) {   // little-endian on every host
        const uint32_t u = (uint32_t)p[4 * i] | ((uint32_t)p[4 * i + 1] << 8) | ((uint32_t)p[4 * i + 2] << 16) | ((uint32_t)p[4 * i + 3] << 24);
        memcpy(&buf[i], &u, 4);
      }
      pos += n * 4;
    }
    for(size_t i = 0; i < n; i++)
      if(!std::isfinite(buf[i])) throw ParseError{name + ": Nan or infinite neural net weight or parameter"};
    return buf;
  }
};

constexpr int MAX_CHANNELS = 8192, MAX_FILTER = 9;   // far above any KataGo-family net; keeps every size product inside size_t

kc_conv_desc parseConv(Cursor& c, Layers& L) {
  const std::string name = c.token("convlayer");
  kc_conv_desc d{};
  d.convYSize = c.integer(name); d.convXSize = c.integer(name); d.inChannels = c.integer(name); d.outChannels = c.integer(name);
  const int dilY = c.integer(name), dilX = c.integer(name);
  if(d.convXSize <= 0 || d.convYSize <= 0) throw ParseError{name + ": convolution filter sizes must be positive"};
  if(d.inChannels <= 0 || d.outChannels <= 0) throw ParseError{name + ": number of in and out channels must be positive"};
  if(dilX <= 0 || dilY <= 0) throw ParseError{name + ": dilation factors must be positive"};
  if(d.convXSize % 2 != 1 || d.convYSize % 2 != 1) throw ParseError{name + ": convolution filter sizes must be odd, found even sizes"};
  if(dilX != 1 || dilY != 1) throw ParseError{name + ": dilated convolutions are not supported by the B200 backend"};
  // bounds before any size arithmetic: 9 * 9 * 8192 * 8192 fits size_t with room to spare
  if(d.convXSize > MAX_FILTER || d.convYSize > MAX_FILTER) throw ParseError{name + ": convolution filter size above " + std::to_string(MAX_FILTER)};
  if(d.inChannels > MAX_CHANNELS || d.outChannels > MAX_CHANNELS) throw ParseError{name + ": more than " + std::to_string(MAX_CHANNELS) + " channels"};
  const std::vector<float> f = c.floats((size_t)d.convYSize * d.convXSize * d.inChannels * d.outChannels, name);
  // file order y,x,ic,oc -> oc,ic,y,x (desc.cpp:131-152)
  std::vector<float> w(f.size());
  size_t idx = 0;
  for(int y = 0; y < d.convYSize; y++)
    for(int x = 0; x < d.convXSize; x++)
      for(int ic = 0; ic < d.inChannels; ic++)
        for(int oc = 0; oc < d.outChannels; oc++)
          w[(((size_t)oc * d.inChannels + ic) * d.convYSize + y) * d.convXSize + x] = f[idx++];
  d.weights = L.keep(std::move(w));
  return d;
}

kc_bn_desc parseBN(Cursor& c, Layers& L) {
  const std::string name = c.token("bnlayer");
  kc_bn_desc d{};
  d.numChannels = c.integer(name); d.epsilon = c.real(name); d.hasScale = c.integer(name); d.hasBias = c.integer(name);
  if(d.numChannels < 1) throw ParseError{name + ": numChannels (" + std::to_string(d.numChannels) + ") < 1"};
  if(d.numChannels > MAX_CHANNELS) throw ParseError{name + ": more than " + std::to_string(MAX_CHANNELS) + " channels"};
  if(!(d.epsilon > 0)) throw ParseError{name + ": epsilon (" + std::to_string(d.epsilon) + ") <= 0"};
  const size_t n = (size_t)d.numChannels;
  d.mean = L.keep(c.floats(n, name));
  d.variance = L.keep(c.floats(n, name));
  d.scale = L.keep(d.hasScale ? c.floats(n, name) : std::vector<float>(n, 1.0f));   // desc.cpp:198-212: absent = 1 / 0
  d.bias = L.keep(d.hasBias ? c.floats(n, name) : std::vector<float>(n, 0.0f));
  d.hasScale = 1; d.hasBias = 1;
  return d;
}

int parseActivation(Cursor& c) {
  const std::string name = c.token("activation");
  if(!c.peekStartsWith("ACTIVATION_")) return 1;   // ReLU
  const std::string kind = c.token(name);
  if(kind == "ACTIVATION_IDENTITY") return 0;
  if(kind == "ACTIVATION_RELU") return 1;
  if(kind == "ACTIVATION_MISH") return 2;
  throw ParseError{name + ": unknown activation " + kind};
}

kc_matmul_desc parseMatMul(Cursor& c, Layers& L) {
  const std::string name = c.token("matmullayer");
  kc_matmul_desc d{};
  d.inChannels = c.integer(name); d.outChannels = c.integer(name);
  if(d.inChannels <= 0 || d.outChannels <= 0) throw ParseError{name + ": number of in and out channels must be positive"};
  if(d.inChannels > 3 * MAX_CHANNELS || d.outChannels > MAX_CHANNELS) throw ParseError{name + ": matrix dimensions out of range"};
  d.weights = L.keep(c.floats((size_t)d.inChannels * d.outChannels, name));   // file order ic,oc is the layout used (desc.cpp:284-299)
  return d;
}

kc_matbias_desc parseMatBias(Cursor& c, Layers& L) {
  const std::string name = c.token("matbiaslayer");
  kc_matbias_desc d{};
  d.numChannels = c.integer(name);
  if(d.numChannels <= 0) throw ParseError{name + ": number of channels must be positive"};
  if(d.numChannels > MAX_CHANNELS) throw ParseError{name + ": more than " + std::to_string(MAX_CHANNELS) + " channels"};
  d.weights = L.keep(c.floats((size_t)d.numChannels, name));
  return d;
}

#define MF_REQUIRE(cond, msg) do { if(!(cond)) throw ParseError{msg}; } while(0)
std::string mism(const std::string& where, const char* a, int av, const char* b, int bv) {
  return where + ": " + a + " (" + std::to_string(av) + ") != " + b + " (" + std::to_string(bv) + ")";
}

}  // namespace

struct kc_modelfile {
  std::string name, sha256;
  kc_model_desc desc;
  std::vector<kc_block_desc> blocks;
  Layers layers;
};

namespace {

void parseModel(const std::string& text, bool binary, kc_modelfile& m) {
  Cursor c(text, binary);
  Layers& L = m.layers;
  kc_model_desc& d = m.desc;
  memset(&d, 0, sizeof(d));
  m.name = c.token("model name");
  d.version = c.integer(m.name + ": model failed to parse version");
  MF_REQUIRE(d.version >= 0, m.name + ": model version must be non-negative");
  MF_REQUIRE(d.version == 1, m.name + ": model version " + std::to_string(d.version) + " is not a Coffee model version (modelversion.h: only version 1 exists)");
  d.numInputChannels = c.integer(m.name + ": numInputChannels");
  MF_REQUIRE(d.numInputChannels > 0, m.name + ": model numInputChannels must be positive");
  d.numInputGlobalChannels = c.integer(m.name + ": numInputGlobalChannels");
  MF_REQUIRE(d.numInputGlobalChannels > 0, m.name + ": model numInputGlobalChannels must be positive");
  MF_REQUIRE(d.numInputChannels == KC_NUM_SPATIAL_V1 && d.numInputGlobalChannels == KC_NUM_GLOBAL_V1,
             m.name + ": a version 1 model takes " + std::to_string(KC_NUM_SPATIAL_V1) + " spatial and " + std::to_string(KC_NUM_GLOBAL_V1) + " global input channels");
  // ---- trunk (desc.cpp:648-696)
  const std::string tname = c.token("trunk");
  d.numBlocks = c.integer(tname); d.trunkNumChannels = c.integer(tname); d.midNumChannels = c.integer(tname); d.regularNumChannels = c.integer(tname);
  (void)c.integer(tname);   // dilatedNumChannels, unused
  d.gpoolNumChannels = c.integer(tname);
  MF_REQUIRE(d.numBlocks >= 1, tname + ": trunk num blocks must be positive");
  MF_REQUIRE(d.trunkNumChannels > 0 && d.midNumChannels > 0 && d.regularNumChannels > 0 && d.gpoolNumChannels > 0, tname + ": all numbers of channels must be positive");
  d.initialConv = parseConv(c, L);
  MF_REQUIRE(d.initialConv.outChannels == d.trunkNumChannels, mism(tname, "initialConv.outChannels", d.initialConv.outChannels, "trunkNumChannels", d.trunkNumChannels));
  d.initialMatMul = parseMatMul(c, L);
  MF_REQUIRE(d.initialMatMul.outChannels == d.trunkNumChannels, mism(tname, "initialMatMul.outChannels", d.initialMatMul.outChannels, "trunkNumChannels", d.trunkNumChannels));
  m.blocks.resize(d.numBlocks);
  for(int i = 0; i < d.numBlocks; i++) {
    kc_block_desc& b = m.blocks[i];
    memset(&b, 0, sizeof(b));
    const std::string kind = c.token(tname + ": block kind");
    if(kind == "ordinary_block") {
      const std::string bn = c.token("res block");
      b.kind = 0;
      b.preBN = parseBN(c, L); b.preActivation = parseActivation(c); b.regularConv = parseConv(c, L);
      b.midBN = parseBN(c, L); b.midActivation = parseActivation(c); b.finalConv = parseConv(c, L);
      b.gpoolActivation = 1;
      MF_REQUIRE(b.preBN.numChannels == b.regularConv.inChannels, mism(bn, "preBN.numChannels", b.preBN.numChannels, "regularConv.inChannels", b.regularConv.inChannels));
      MF_REQUIRE(b.midBN.numChannels == b.regularConv.outChannels, mism(bn, "midBN.numChannels", b.midBN.numChannels, "regularConv.outChannels", b.regularConv.outChannels));
      MF_REQUIRE(b.midBN.numChannels == b.finalConv.inChannels, mism(bn, "midBN.numChannels", b.midBN.numChannels, "finalConv.inChannels", b.finalConv.inChannels));
    } else if(kind == "gpool_block") {
      const std::string bn = c.token("gpool res block");
      b.kind = 2;
      b.preBN = parseBN(c, L); b.preActivation = parseActivation(c); b.regularConv = parseConv(c, L); b.gpoolConv = parseConv(c, L);
      b.gpoolBN = parseBN(c, L); b.gpoolActivation = parseActivation(c); b.gpoolToBiasMul = parseMatMul(c, L);
      b.midBN = parseBN(c, L); b.midActivation = parseActivation(c); b.finalConv = parseConv(c, L);
      MF_REQUIRE(b.preBN.numChannels == b.regularConv.inChannels, mism(bn, "preBN.numChannels", b.preBN.numChannels, "regularConv.inChannels", b.regularConv.inChannels));
      MF_REQUIRE(b.preBN.numChannels == b.gpoolConv.inChannels, mism(bn, "preBN.numChannels", b.preBN.numChannels, "gpoolConv.inChannels", b.gpoolConv.inChannels));
      MF_REQUIRE(b.gpoolBN.numChannels == b.gpoolConv.outChannels, mism(bn, "gpoolBN.numChannels", b.gpoolBN.numChannels, "gpoolConv.outChannels", b.gpoolConv.outChannels));
      MF_REQUIRE(b.gpoolBN.numChannels * 3 == b.gpoolToBiasMul.inChannels, mism(bn, "gpoolBN.numChannels * 3", b.gpoolBN.numChannels * 3, "gpoolToBiasMul.inChannels", b.gpoolToBiasMul.inChannels));
      MF_REQUIRE(b.midBN.numChannels == b.regularConv.outChannels, mism(bn, "midBN.numChannels", b.midBN.numChannels, "regularConv.outChannels", b.regularConv.outChannels));
      MF_REQUIRE(b.midBN.numChannels == b.gpoolToBiasMul.outChannels, mism(bn, "midBN.numChannels", b.midBN.numChannels, "gpoolToBiasMul.outChannels", b.gpoolToBiasMul.outChannels));
      MF_REQUIRE(b.midBN.numChannels == b.finalConv.inChannels, mism(bn, "midBN.numChannels", b.midBN.numChannels, "finalConv.inChannels", b.finalConv.inChannels));
    } else if(kind == "nested_bottleneck_block") {
      throw ParseError{tname + ": nested bottleneck blocks are not supported by the B200 backend"};
    } else {
      throw ParseError{tname + ": found unknown block kind: " + kind};
    }
    MF_REQUIRE(b.preBN.numChannels == d.trunkNumChannels, mism(tname, "block preBN.numChannels", b.preBN.numChannels, "trunkNumChannels", d.trunkNumChannels));
    MF_REQUIRE(b.finalConv.outChannels == d.trunkNumChannels, mism(tname, "block finalConv.outChannels", b.finalConv.outChannels, "trunkNumChannels", d.trunkNumChannels));
  }
  d.blocks = m.blocks.data();
  d.trunkTipBN = parseBN(c, L);
  d.trunkTipActivation = parseActivation(c);
  MF_REQUIRE(d.trunkTipBN.numChannels == d.trunkNumChannels, mism(tname, "trunkTipBN.numChannels", d.trunkTipBN.numChannels, "trunkNumChannels", d.trunkNumChannels));
  // ---- policy head (desc.cpp:751-810)
  const std::string pname = c.token("policy head");
  d.p1Conv = parseConv(c, L); d.g1Conv = parseConv(c, L); d.g1BN = parseBN(c, L); d.g1Activation = parseActivation(c);
  d.gpoolToBiasMul = parseMatMul(c, L); d.p1BN = parseBN(c, L); d.p1Activation = parseActivation(c); d.p2Conv = parseConv(c, L);
  const kc_matmul_desc passMul = parseMatMul(c, L);   // part of the format; Coffee has no pass move
  MF_REQUIRE(d.p1Conv.outChannels == d.p1BN.numChannels, mism(pname, "p1Conv.outChannels", d.p1Conv.outChannels, "p1BN.numChannels", d.p1BN.numChannels));
  MF_REQUIRE(d.g1Conv.outChannels == d.g1BN.numChannels, mism(pname, "g1Conv.outChannels", d.g1Conv.outChannels, "g1BN.numChannels", d.g1BN.numChannels));
  MF_REQUIRE(d.gpoolToBiasMul.inChannels == d.g1BN.numChannels * 3, mism(pname, "gpoolToBiasMul.inChannels", d.gpoolToBiasMul.inChannels, "g1BN.numChannels * 3", d.g1BN.numChannels * 3));
  MF_REQUIRE(d.gpoolToBiasMul.outChannels == d.p1BN.numChannels, mism(pname, "gpoolToBiasMul.outChannels", d.gpoolToBiasMul.outChannels, "p1BN.numChannels", d.p1BN.numChannels));
  MF_REQUIRE(d.p2Conv.inChannels == d.p1BN.numChannels, mism(pname, "p2Conv.inChannels", d.p2Conv.inChannels, "p1BN.numChannels", d.p1BN.numChannels));
  MF_REQUIRE(passMul.inChannels == d.g1BN.numChannels * 3, mism(pname, "gpoolToPassMul.inChannels", passMul.inChannels, "g1BN.numChannels * 3", d.g1BN.numChannels * 3));
  MF_REQUIRE(d.p2Conv.outChannels == 4, pname + ": p2Conv.outChannels (" + std::to_string(d.p2Conv.outChannels) + ") != 4 (one policy channel per direction)");
  // ---- value head (desc.cpp:844-925)
  const std::string vname = c.token("value head");
  d.v1Conv = parseConv(c, L); d.v1BN = parseBN(c, L); d.v1Activation = parseActivation(c); d.v2Mul = parseMatMul(c, L); d.v2Bias = parseMatBias(c, L);
  d.v2Activation = parseActivation(c); d.v3Mul = parseMatMul(c, L); d.v3Bias = parseMatBias(c, L); d.sv3Mul = parseMatMul(c, L); d.sv3Bias = parseMatBias(c, L);
  d.vOwnershipConv = parseConv(c, L);
  MF_REQUIRE(d.v1Conv.outChannels == d.v1BN.numChannels, mism(vname, "v1Conv.outChannels", d.v1Conv.outChannels, "v1BN.numChannels", d.v1BN.numChannels));
  MF_REQUIRE(d.v2Mul.inChannels == d.v1BN.numChannels * 3, mism(vname, "v2Mul.inChannels", d.v2Mul.inChannels, "v1BN.numChannels * 3", d.v1BN.numChannels * 3));
  MF_REQUIRE(d.v2Mul.outChannels == d.v2Bias.numChannels, mism(vname, "v2Mul.outChannels", d.v2Mul.outChannels, "v2Bias.numChannels", d.v2Bias.numChannels));
  MF_REQUIRE(d.v2Mul.outChannels == d.v3Mul.inChannels, mism(vname, "v2Mul.outChannels", d.v2Mul.outChannels, "v3Mul.inChannels", d.v3Mul.inChannels));
  MF_REQUIRE(d.v3Mul.outChannels == 2, vname + ": v3Mul.outChannels (" + std::to_string(d.v3Mul.outChannels) + ") != 2 (win, loss)");
  MF_REQUIRE(d.v3Bias.numChannels == 2, vname + ": v3Bias.numChannels (" + std::to_string(d.v3Bias.numChannels) + ") != 2");
  MF_REQUIRE(d.sv3Mul.inChannels == d.v2Mul.outChannels, mism(vname, "sv3Mul.inChannels", d.sv3Mul.inChannels, "v2Mul.outChannels", d.v2Mul.outChannels));
  MF_REQUIRE(d.sv3Mul.outChannels == 2, vname + ": sv3Mul.outChannels (" + std::to_string(d.sv3Mul.outChannels) + ") != 2 (varTimeLeft, shorttermWinlossError)");
  MF_REQUIRE(d.sv3Bias.numChannels == 2, vname + ": sv3Bias.numChannels (" + std::to_string(d.sv3Bias.numChannels) + ") != 2");
  MF_REQUIRE(d.vOwnershipConv.inChannels == d.v1Conv.outChannels, mism(vname, "vOwnershipConv.inChannels", d.vOwnershipConv.inChannels, "v1Conv.outChannels", d.v1Conv.outChannels));
  MF_REQUIRE(d.vOwnershipConv.outChannels == 1, vname + ": vOwnershipConv.outChannels (" + std::to_string(d.vOwnershipConv.outChannels) + ") != 1");
  // ---- whole model (desc.cpp:1064-1094)
  MF_REQUIRE(d.numInputChannels == d.initialConv.inChannels, mism(m.name, "numInputChannels", d.numInputChannels, "trunk.initialConv.inChannels", d.initialConv.inChannels));
  MF_REQUIRE(d.numInputGlobalChannels == d.initialMatMul.inChannels, mism(m.name, "numInputGlobalChannels", d.numInputGlobalChannels, "trunk.initialMatMul.inChannels", d.initialMatMul.inChannels));
  MF_REQUIRE(d.trunkNumChannels == d.p1Conv.inChannels, mism(m.name, "trunk.trunkNumChannels", d.trunkNumChannels, "policyHead.p1Conv.inChannels", d.p1Conv.inChannels));
  MF_REQUIRE(d.trunkNumChannels == d.g1Conv.inChannels, mism(m.name, "trunk.trunkNumChannels", d.trunkNumChannels, "policyHead.g1Conv.inChannels", d.g1Conv.inChannels));
  MF_REQUIRE(d.trunkNumChannels == d.v1Conv.inChannels, mism(m.name, "trunk.trunkNumChannels", d.trunkNumChannels, "valueHead.v1Conv.inChannels", d.v1Conv.inChannels));
}

// ---- writer: the same format, so that a random-init or converted net can be handed to anything that reads model files ----
struct Writer {
  std::string out;
  bool binary;
  void tok(const std::string& s) { out += s; out += '\n'; }
  void ints(std::initializer_list<int> v) { std::string l; for(int x : v) { if(!l.empty()) l += ' '; l += std::to_string(x); } tok(l); }
  void floats(const float* p, size_t n) {
    if(binary) {
      out += "@BIN@";
      for(size_t i = 0; i < n; i++) { uint32_t u; memcpy(&u, &p[i], 4); for(int k = 0; k < 4; k++) out.push_back((char)((u >> (8 * k)) & 0xff)); }
      out += '\n';
    } else {
      char buf[32];
      for(size_t i = 0; i < n; i++) { snprintf(buf, sizeof(buf), "%.9g", (double)p[i]); out += buf; out += (i + 1 == n || i % 8 == 7) ? '\n' : ' '; }
    }
  }
  void conv(const std::string& name, const kc_conv_desc& d) {
    tok(name); ints({d.convYSize, d.convXSize, d.inChannels, d.outChannels, 1, 1});
    std::vector<float> f((size_t)d.convYSize * d.convXSize * d.inChannels * d.outChannels);
    size_t idx = 0;
    for(int y = 0; y < d.convYSize; y++)
      for(int x = 0; x < d.convXSize; x++)
        for(int ic = 0; ic < d.inChannels; ic++)
          for(int oc = 0; oc < d.outChannels; oc++)
            f[idx++] = d.weights[(((size_t)oc * d.inChannels + ic) * d.convYSize + y) * d.convXSize + x];
    floats(f.data(), f.size());
  }
  void bn(const std::string& name, const kc_bn_desc& d) {
    tok(name);
    char eps[32]; snprintf(eps, sizeof(eps), "%.9g", (double)d.epsilon);
    const bool hs = d.hasScale && d.scale, hb = d.hasBias && d.bias;
    tok(std::to_string(d.numChannels) + " " + eps + " " + (hs ? "1" : "0") + " " + (hb ? "1" : "0"));
    floats(d.mean, d.numChannels); floats(d.variance, d.numChannels);
    if(hs) floats(d.scale, d.numChannels);
    if(hb) floats(d.bias, d.numChannels);
  }
  // Version 1 < 11: the reference's ActivationLayerDesc reads no kind token and means ReLU (desc.cpp:243-256), so a ReLU layer is written
  // as its name alone -- a file of a ReLU net is in the reference's format.  Identity / Mish have no version-1 spelling: their kind
  // token is this parser's extension (see the header), and such a file is readable by kc_modelfile_load only.
  void act(const std::string& name, int a) { tok(name); if(a != 1) tok(a == 0 ? "ACTIVATION_IDENTITY" : "ACTIVATION_MISH"); }
  void matmul(const std::string& name, const kc_matmul_desc& d) { tok(name); ints({d.inChannels, d.outChannels}); floats(d.weights, (size_t)d.inChannels * d.outChannels); }
  void matbias(const std::string& name, const kc_matbias_desc& d) { tok(name); ints({d.numChannels}); floats(d.weights, d.numChannels); }
};

std::string gzipString(const std::string& in) {
  z_stream zs;
  memset(&zs, 0, sizeof(zs));
  if(deflateInit2(&zs, 6, Z_DEFLATED, 15 + 16, 8, Z_DEFAULT_STRATEGY) != Z_OK) throw ParseError{"deflateInit2 failed"};
  std::string out(deflateBound(&zs, (uLong)in.size()) + 64, '\0');
  zs.next_in = (Bytef*)in.data(); zs.avail_in = (uInt)in.size();
  zs.next_out = (Bytef*)&out[0]; zs.avail_out = (uInt)out.size();
  const int ret = deflate(&zs, Z_FINISH);
  if(ret != Z_STREAM_END) { deflateEnd(&zs); throw ParseError{"deflate failed"}; }
  out.resize(out.size() - zs.avail_out);
  deflateEnd(&zs);
  return out;
}

}  // namespace

extern "C" {

int kc_modelfile_load(const char* path, const char* expectedSha256, kc_modelfile** out) {
  KC_CHECK(path && out, "kc_modelfile_load: null argument");
  const std::string file(path);
  std::unique_ptr<kc_modelfile> m(new kc_modelfile());
  try {
    std::ifstream in(file, std::ios::in | std::ios::binary);
    if(!in.good()) throw ParseError{"could not open file"};
    std::string raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    m->sha256 = sha256Hex(raw);
    if(expectedSha256 && expectedSha256[0] != '\0' && toLower(expectedSha256) != m->sha256)
      throw ParseError{"File " + file + " sha256 was " + m->sha256 + " which does not match the expected sha256 " + expectedSha256};
    const std::string lower = toLower(file);
    if(isSuffix(lower, ".txt")) parseModel(raw, false, *m);
    else if(isSuffix(lower, ".bin")) parseModel(raw, true, *m);
    else if(isSuffix(lower, ".gz")) {
      const std::string text = gunzip(raw, file);
      const bool binary = !isSuffix(lower, ".txt.gz");
      try {
        parseModel(text, binary, *m);
      } catch(const ParseError& e) {
        if(!binary || isSuffix(lower, ".bin.gz")) throw;
        try {   // ambiguous extension: try again as text (desc.cpp:1180-1195)
          m.reset(new kc_modelfile());
          m->sha256 = sha256Hex(raw);
          parseModel(text, false, *m);
        } catch(const ParseError& e2) {
          throw ParseError{"Could neither parse .gz model as .txt.gz model nor as .bin.gz model, errors were:\n" + e2.msg + "\n" + e.msg};
        }
      }
    } else
      throw ParseError{"Model file should end with .txt, .bin, .txt.gz, .bin.gz, or possibly just .gz. (If it doesn't have one of these extensions already, "
                       "it's probably the wrong file, renaming will probably NOT help)."};
  } catch(const ParseError& e) {
    return kc::fail("Error loading or parsing model file " + file + ": " + e.msg);
  } catch(const std::exception& e) {   // bad_alloc / length_error from a corrupt header: no C++ exception crosses the C ABI
    return kc::fail("Error loading or parsing model file " + file + ": " + e.what());
  }
  *out = m.release();
  return 0;
}

int kc_modelfile_free(kc_modelfile* f) { delete f; return 0; }
const kc_model_desc* kc_modelfile_desc(const kc_modelfile* f) { return f ? &f->desc : nullptr; }
const char* kc_modelfile_name(const kc_modelfile* f) { return f ? f->name.c_str() : ""; }
const char* kc_modelfile_sha256(const kc_modelfile* f) { return f ? f->sha256.c_str() : ""; }

int kc_modelfile_write(const kc_model_desc* d, const char* name, const char* path) {
  KC_CHECK(d && name && path, "kc_modelfile_write: null argument");
  const std::string file(path), lower = toLower(file);
  KC_CHECK(isSuffix(lower, ".txt") || isSuffix(lower, ".bin") || isSuffix(lower, ".txt.gz") || isSuffix(lower, ".bin.gz"),
           "kc_modelfile_write: the file name must end with .txt, .bin, .txt.gz or .bin.gz");
  KC_CHECK(d->numBlocks >= 1, "kc_modelfile_write: the format needs at least one block (desc.cpp:661-662)");
  try {   // no C++ exception (bad_alloc on a huge description) crosses the C ABI
  Writer w;
  w.binary = isSuffix(lower, ".bin") || isSuffix(lower, ".bin.gz");
  w.tok(name);
  w.ints({d->version}); w.ints({d->numInputChannels}); w.ints({d->numInputGlobalChannels});
  w.tok("trunk");
  w.ints({d->numBlocks, d->trunkNumChannels, d->midNumChannels, d->regularNumChannels, d->regularNumChannels, d->gpoolNumChannels});
  w.conv("conv1", d->initialConv);
  w.matmul("ginputw", d->initialMatMul);
  for(int i = 0; i < d->numBlocks; i++) {
    const kc_block_desc& b = d->blocks[i];
    const std::string p = "block" + std::to_string(i);
    KC_CHECK(b.kind == 0 || b.kind == 2, "kc_modelfile_write: unknown block kind");
    w.tok(b.kind == 2 ? "gpool_block" : "ordinary_block");
    w.tok(p);
    w.bn(p + "/norm1", b.preBN); w.act(p + "/actv1", b.preActivation); w.conv(p + "/w1a", b.regularConv);
    if(b.kind == 2) {
      w.conv(p + "/w1b", b.gpoolConv); w.bn(p + "/norm1b", b.gpoolBN); w.act(p + "/actv1b", b.gpoolActivation); w.matmul(p + "/w1r", b.gpoolToBiasMul);
    }
    w.bn(p + "/norm2", b.midBN); w.act(p + "/actv2", b.midActivation); w.conv(p + "/w2", b.finalConv);
  }
  w.bn("trunk/norm", d->trunkTipBN); w.act("trunk/actv", d->trunkTipActivation);
  w.tok("policyhead");
  w.conv("p1/w", d->p1Conv); w.conv("g1/w", d->g1Conv); w.bn("g1/norm", d->g1BN); w.act("g1/actv", d->g1Activation);
  w.matmul("matmulg2w", d->gpoolToBiasMul); w.bn("p1/norm", d->p1BN); w.act("p1/actv", d->p1Activation); w.conv("p2/w", d->p2Conv);
  {   // gpoolToPassMul: required by the format, unused by Coffee -- zeros
    std::vector<float> z((size_t)d->gpoolToBiasMul.inChannels, 0.0f);
    const kc_matmul_desc pass{d->gpoolToBiasMul.inChannels, 1, z.data()};
    w.matmul("matmulpass", pass);
  }
  w.tok("valuehead");
  w.conv("v1/w", d->v1Conv); w.bn("v1/norm", d->v1BN); w.act("v1/actv", d->v1Activation); w.matmul("v2/w", d->v2Mul); w.matbias("v2/b", d->v2Bias);
  w.act("v2/actv", d->v2Activation); w.matmul("v3/w", d->v3Mul); w.matbias("v3/b", d->v3Bias); w.matmul("sv3/w", d->sv3Mul); w.matbias("sv3/b", d->sv3Bias);
  w.conv("vownership/w", d->vOwnershipConv);
  std::string bytes;
  try {
    bytes = isSuffix(lower, ".gz") ? gzipString(w.out) : w.out;
  } catch(const ParseError& e) { return kc::fail("kc_modelfile_write: " + e.msg); }
  std::ofstream out(file, std::ios::out | std::ios::binary | std::ios::trunc);
  KC_CHECK(out.good(), "kc_modelfile_write: could not open " + file);
  out.write(bytes.data(), (std::streamsize)bytes.size());
  out.close();
  KC_CHECK(out.good(), "kc_modelfile_write: write to " + file + " failed");
  } catch(const std::exception& e) { return kc::fail(std::string("kc_modelfile_write: ") + e.what()); }
  return 0;
}

}  // extern "C"
