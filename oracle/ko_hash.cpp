// ORACLE (test infrastructure only, see kc_oracle.h): hashing, PRNG and Zobrist tables.
//
// Restates, from the published algorithms, what the reference gets from
//   cpp/core/md5.cpp (RFC 1321), cpp/core/sha2.cpp (FIPS 180-4),
//   cpp/core/rand_helpers.h:29-66 (xorshift1024*, PCG32), cpp/core/rand.cpp:276-318 (seeding),
//   cpp/core/rand.h (nextUInt/nextUInt64/nextDouble/nextGaussian), cpp/core/hash.cpp:27-73 (mixers),
//   cpp/game/board.cpp:134-178 (Board::initHash).
// Pinned by the reference golden vectors in tests/test_oracle_hash.py and by oracle/_ref.
#include "kc_oracle.h"

#include <cmath>
#include <cstring>
#include <string>
#include <vector>

namespace {

// ---------------------------------------------------------------- MD5 (RFC 1321)
inline uint32_t rotl32(uint32_t x, int c) { return (x << c) | (x >> (32 - c)); }

void md5_impl(const uint8_t* msg, size_t len, uint32_t out[4]) {
  static const int S[64] = {7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22, 7, 12, 17, 22,
                            5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20, 5, 9,  14, 20,
                            4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23, 4, 11, 16, 23,
                            6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21, 6, 10, 15, 21};
  uint32_t K[64];
  for(int i = 0; i < 64; i++)
    K[i] = (uint32_t)(int64_t)std::floor(std::fabs(std::sin((double)(i + 1))) * 4294967296.0);
  uint32_t a0 = 0x67452301u, b0 = 0xefcdab89u, c0 = 0x98badcfeu, d0 = 0x10325476u;
  size_t padded = ((len + 8) / 64 + 1) * 64;
  std::vector<uint8_t> buf(padded, 0);
  memcpy(buf.data(), msg, len);
  buf[len] = 0x80;
  uint64_t bits = (uint64_t)len * 8;
  for(int i = 0; i < 8; i++)
    buf[padded - 8 + i] = (uint8_t)(bits >> (8 * i));
  for(size_t off = 0; off < padded; off += 64) {
    uint32_t M[16];
    for(int i = 0; i < 16; i++)
      M[i] = (uint32_t)buf[off + 4 * i] | ((uint32_t)buf[off + 4 * i + 1] << 8) |
             ((uint32_t)buf[off + 4 * i + 2] << 16) | ((uint32_t)buf[off + 4 * i + 3] << 24);
    uint32_t A = a0, B = b0, C = c0, D = d0;
    for(int i = 0; i < 64; i++) {
      uint32_t F;
      int g;
      if(i < 16) { F = (B & C) | (~B & D); g = i; }
      else if(i < 32) { F = (D & B) | (~D & C); g = (5 * i + 1) & 15; }
      else if(i < 48) { F = B ^ C ^ D; g = (3 * i + 5) & 15; }
      else { F = C ^ (B | ~D); g = (7 * i) & 15; }
      F = F + A + K[i] + M[g];
      A = D; D = C; C = B;
      B = B + rotl32(F, S[i]);
    }
    a0 += A; b0 += B; c0 += C; d0 += D;
  }
  // md5.cpp:136-139 returns the four state words as native uint32 (h0..h3)
  out[0] = a0; out[1] = b0; out[2] = c0; out[3] = d0;
}

// ---------------------------------------------------------------- SHA-256 (FIPS 180-4)
inline uint32_t rotr32(uint32_t x, int c) { return (x >> c) | (x << (32 - c)); }

void sha256_impl(const uint8_t* msg, size_t len, uint8_t out[32]) {
  static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
  uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a,
                   0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
  size_t padded = ((len + 8) / 64 + 1) * 64;
  std::vector<uint8_t> buf(padded, 0);
  memcpy(buf.data(), msg, len);
  buf[len] = 0x80;
  uint64_t bits = (uint64_t)len * 8;
  for(int i = 0; i < 8; i++)
    buf[padded - 1 - i] = (uint8_t)(bits >> (8 * i));
  for(size_t off = 0; off < padded; off += 64) {
    uint32_t w[64];
    for(int i = 0; i < 16; i++)
      w[i] = ((uint32_t)buf[off + 4 * i] << 24) | ((uint32_t)buf[off + 4 * i + 1] << 16) |
             ((uint32_t)buf[off + 4 * i + 2] << 8) | (uint32_t)buf[off + 4 * i + 3];
    for(int i = 16; i < 64; i++) {
      uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
      uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
      w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for(int i = 0; i < 64; i++) {
      uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
      uint32_t ch = (e & f) ^ (~e & g);
      uint32_t t1 = hh + S1 + ch + K[i] + w[i];
      uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
      uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
      uint32_t t2 = S0 + mj;
      hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
  }
  for(int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(h[i] >> 24);
    out[4 * i + 1] = (uint8_t)(h[i] >> 16);
    out[4 * i + 2] = (uint8_t)(h[i] >> 8);
    out[4 * i + 3] = (uint8_t)h[i];
  }
}

}  // namespace

// ---------------------------------------------------------------- Rand (rand.cpp / rand.h / rand_helpers.h)
struct ko_rand {
  uint64_t a[16];
  uint64_t a_idx;
  uint64_t s;
  bool hasGaussian;
  double storedGaussian;

  // rand_helpers.h:29-41
  uint32_t xormNext() {
    uint64_t a0 = a[a_idx];
    uint64_t a1 = a[a_idx = (a_idx + 1) & 15];
    a1 ^= a1 << 31;
    a1 ^= a1 >> 11;
    a0 ^= a0 >> 30;
    a[a_idx] = a0 ^ a1;
    uint64_t r = a[a_idx] * 1181783497276652981ULL;
    return (uint32_t)(r >> 32);
  }
  // rand_helpers.h:60-66
  uint32_t pcgNext() {
    s = s * 6364136223846793005ULL + 1442695040888963407ULL;
    uint32_t x = (uint32_t)(((s >> 18) ^ s) >> 27);
    int rot = (int)(s >> 59);
    return rot == 0 ? x : ((x >> rot) | (x << (32 - rot)));
  }
  // rand.h: nextUInt = pcg32 + xorm (in that evaluation order; both are independent streams)
  uint32_t nextUInt() {
    uint32_t p = pcgNext();
    uint32_t x = xormNext();
    return p + x;
  }
  uint64_t nextUInt64() {
    uint64_t lower = (uint64_t)nextUInt();
    uint64_t upper = (uint64_t)nextUInt() << 32;
    return lower | upper;
  }
  double nextDouble() {
    double x;
    do {
      uint64_t bits = nextUInt64() & ((1ULL << 53) - 1ULL);
      x = (double)bits / (double)(1ULL << 53);
    } while(!(x >= 0.0 && x < 1.0));
    return x;
  }
  double nextGaussian() {
    if(hasGaussian) {
      hasGaussian = false;
      return storedGaussian;
    }
    double v1, v2, ss;
    do {
      v1 = nextDouble() * 2.0 - 1.0;
      v2 = nextDouble() * 2.0 - 1.0;
      ss = v1 * v1 + v2 * v2;
    } while(ss >= 1 || ss == 0);
    double mult = std::sqrt(-2 * std::log(ss) / ss);
    storedGaussian = v2 * mult;
    hasGaussian = true;
    return v1 * mult;
  }
  // rand.cpp:276-318
  void init(const std::string& seed) {
    std::string str;
    {
      uint32_t h[4];
      md5_impl((const uint8_t*)seed.data(), seed.size(), h);
      str += "|";
      str += std::to_string(h[0]);
      str += "|";
      str += seed;
    }
    int counter = 0;
    int nextHashIdx = 4;
    uint64_t hash[4];
    auto getNonzero = [&]() -> uint64_t {
      uint64_t v;
      do {
        if(nextHashIdx >= 4) {
          std::string tmp = std::to_string(counter) + str;
          counter += 37;
          ko_sha256_u64((const uint8_t*)tmp.data(), tmp.size(), hash);
          nextHashIdx = 0;
        }
        v = hash[nextHashIdx];
        nextHashIdx += 1;
      } while(v == 0);
      return v;
    };
    for(int i = 0; i < 16; i++)
      a[i] = getNonzero();
    a_idx = 0;
    s = getNonzero();
    hasGaussian = false;
    storedGaussian = 0.0;
  }
};

extern "C" {

void ko_md5(const uint8_t* msg, size_t len, uint32_t out[4]) { md5_impl(msg, len, out); }
void ko_sha256(const uint8_t* msg, size_t len, uint8_t out[32]) { sha256_impl(msg, len, out); }
// sha2.cpp:1218-1228 (CONVERT_DIGEST_UINT64): big-endian packing of digest bytes
void ko_sha256_u64(const uint8_t* msg, size_t len, uint64_t out[4]) {
  uint8_t d[32];
  sha256_impl(msg, len, d);
  for(int i = 0; i < 4; i++) {
    uint64_t v = 0;
    for(int j = 0; j < 8; j++)
      v = (v << 8) | d[i * 8 + j];
    out[i] = v;
  }
}

// hash.cpp:27-73
uint64_t ko_basic_lcong(uint64_t x) { return 2862933555777941757ULL * x + 3037000493ULL; }
uint64_t ko_basic_lcong2(uint64_t x) { return 6364136223846793005ULL * x + 1442695040888963407ULL; }
uint64_t ko_murmurmix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
uint64_t ko_splitmix64(uint64_t x) {
  x = x + 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}
static inline uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }
uint64_t ko_rrmxmx(uint64_t x) {
  x ^= rotr64(x, 49) ^ rotr64(x, 24);
  x *= 0x9fb21c651e98df25ULL;
  x ^= x >> 28;
  x *= 0x9fb21c651e98df25ULL;
  return x ^ (x >> 28);
}

ko_rand* ko_rand_create(const char* seed) {
  ko_rand* r = new ko_rand();
  r->init(std::string(seed));
  return r;
}
void ko_rand_destroy(ko_rand* r) { delete r; }
uint32_t ko_rand_next_uint(ko_rand* r) { return r->nextUInt(); }
uint64_t ko_rand_next_uint64(ko_rand* r) { return r->nextUInt64(); }
double ko_rand_next_double(ko_rand* r) { return r->nextDouble(); }
double ko_rand_next_gaussian(ko_rand* r) { return r->nextGaussian(); }

void ko_xorshift1024_test(const uint64_t init_a[16], int n, uint32_t* out) {
  ko_rand r;
  for(int i = 0; i < 16; i++) r.a[i] = init_a[i];
  r.a_idx = 0;
  for(int i = 0; i < n; i++) out[i] = r.xormNext();
}
void ko_pcg32_test(uint64_t state, int n, uint32_t* out) {
  ko_rand r;
  r.s = state;
  for(int i = 0; i < n; i++) out[i] = r.pcgNext();
}

// Board::initHash, board.cpp:134-178 (ZOBRIST_BOARD_HASH2 is not on the hot path and is omitted)
void ko_zobrist_tables(uint64_t* board, uint64_t* player, uint64_t* size_x, uint64_t* size_y) {
  ko_rand rand;
  rand.init("Board::initHash()");
  auto nextHash = [&rand](uint64_t* dst) {
    dst[0] = rand.nextUInt64();
    dst[1] = rand.nextUInt64();
  };
  for(int i = 0; i < 4; i++)
    nextHash(player + 2 * i);
  for(int i = 0; i < KO_MAX_ARR_SIZE; i++) {
    for(int j = 0; j < 4; j++) {
      uint64_t* dst = board + ((size_t)i * 4 + j) * 2;
      if(j == 0 || j == 3) { dst[0] = 0; dst[1] = 0; }
      else nextHash(dst);
    }
  }
  rand.init("Board::initHash() for ZOBRIST_SIZE hashes");
  for(int i = 0; i < KO_MAX_LEN + 1; i++) {
    nextHash(size_x + 2 * i);
    nextHash(size_y + 2 * i);
  }
}

}  // extern "C"
