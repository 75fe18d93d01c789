// Host-side generation of the Zobrist tables the device kernels hash with.
//
// Bit-exact with the reference's Board::initHash (cpp/game/board.cpp:134-178), which draws
// 128-bit values from Rand(seed-string) (cpp/core/rand.cpp:276-318): the seed string is expanded
// by MD5 (one word, as a decimal prefix) and SHA-256 (counter-prefixed blocks, big-endian 64-bit
// words) into the state of an xorshift1024* generator and a PCG32 generator whose 32-bit outputs
// are summed (cpp/core/rand.h, rand_helpers.h:29-66).  Also error plumbing and the symmetry maps
// of cpp/neuralnet/nninputs.cpp:252-433 shared by the CUDA translation units.
#include <cmath>
#include <cstring>
#include <mutex>
#include <string>

#include "kc_internal.h"

namespace kc {

static thread_local std::string g_lastError;
void setError(const std::string& msg) { g_lastError = msg; }
int fail(const std::string& msg) { g_lastError = msg; return 1; }
const char* lastErrorCStr() { return g_lastError.c_str(); }

namespace {

struct Bytes {
  std::string data;
  void padTo64(bool bigEndianLength) {
    uint64_t bits = (uint64_t)data.size() * 8;
    data.push_back((char)0x80);
    while(data.size() % 64 != 56) data.push_back((char)0);
    for(int i = 0; i < 8; i++) {
      int shift = bigEndianLength ? 8 * (7 - i) : 8 * i;
      data.push_back((char)((bits >> shift) & 0xff));
    }
  }
  uint32_t le32(size_t off) const {
    const unsigned char* p = (const unsigned char*)data.data() + off;
    return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
  }
  uint32_t be32(size_t off) const {
    const unsigned char* p = (const unsigned char*)data.data() + off;
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
  }
};

inline uint32_t rol(uint32_t v, unsigned s) { return (v << s) | (v >> (32u - s)); }
inline uint32_t ror(uint32_t v, unsigned s) { return (v >> s) | (v << (32u - s)); }

// First state word of MD5(msg) (RFC 1321); Rand::init only uses hash[0].
uint32_t md5FirstWord(const std::string& msg) {
  static const unsigned rot[4][4] = {{7, 12, 17, 22}, {5, 9, 14, 20}, {4, 11, 16, 23}, {6, 10, 15, 21}};
  Bytes b{msg};
  b.padTo64(false);
  uint32_t st[4] = {0x67452301u, 0xefcdab89u, 0x98badcfeu, 0x10325476u};
  for(size_t blk = 0; blk < b.data.size(); blk += 64) {
    uint32_t v[4] = {st[0], st[1], st[2], st[3]};
    for(unsigned i = 0; i < 64; i++) {
      unsigned round = i / 16;
      uint32_t f;
      unsigned g;
      switch(round) {
        case 0: f = (v[1] & v[2]) | (~v[1] & v[3]); g = i; break;
        case 1: f = (v[3] & v[1]) | (~v[3] & v[2]); g = (5 * i + 1) % 16; break;
        case 2: f = v[1] ^ v[2] ^ v[3]; g = (3 * i + 5) % 16; break;
        default: f = v[2] ^ (v[1] | ~v[3]); g = (7 * i) % 16; break;
      }
      uint32_t k = (uint32_t)(uint64_t)std::floor(std::fabs(std::sin((double)i + 1.0)) * 4294967296.0);
      uint32_t sum = v[0] + f + k + b.le32(blk + 4 * g);
      uint32_t nb = v[1] + rol(sum, rot[round][i % 4]);
      v[0] = v[3]; v[3] = v[2]; v[2] = v[1]; v[1] = nb;
    }
    for(int i = 0; i < 4; i++) st[i] += v[i];
  }
  return st[0];
}

// SHA-256 (FIPS 180-4) digest as four big-endian 64-bit words (what SHA2::get256(.., uint64_t[4]) yields).
void sha256Words(const std::string& msg, uint64_t out[4]) {
  // round constants = fractional parts of cube roots of the first 64 primes
  static uint32_t K[64];
  static bool init = false;
  if(!init) {
    int n = 0;
    for(int p = 2; n < 64; p++) {
      bool prime = true;
      for(int q = 2; q * q <= p; q++) if(p % q == 0) { prime = false; break; }
      if(!prime) continue;
      long double c = cbrtl((long double)p);
      K[n++] = (uint32_t)((c - floorl(c)) * 4294967296.0L);
    }
    init = true;
  }
  Bytes b{msg};
  b.padTo64(true);
  uint32_t h[8] = {0x6a09e667u, 0xbb67ae85u, 0x3c6ef372u, 0xa54ff53au, 0x510e527fu, 0x9b05688cu, 0x1f83d9abu, 0x5be0cd19u};
  for(size_t blk = 0; blk < b.data.size(); blk += 64) {
    uint32_t w[64];
    for(int i = 0; i < 16; i++) w[i] = b.be32(blk + 4 * i);
    for(int i = 16; i < 64; i++)
      w[i] = w[i - 16] + (ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] +
             (ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10));
    uint32_t s[8];
    memcpy(s, h, sizeof(s));
    for(int i = 0; i < 64; i++) {
      uint32_t t1 = s[7] + (ror(s[4], 6) ^ ror(s[4], 11) ^ ror(s[4], 25)) + ((s[4] & s[5]) ^ (~s[4] & s[6])) + K[i] + w[i];
      uint32_t t2 = (ror(s[0], 2) ^ ror(s[0], 13) ^ ror(s[0], 22)) + ((s[0] & s[1]) ^ (s[0] & s[2]) ^ (s[1] & s[2]));
      for(int j = 7; j > 0; j--) s[j] = s[j - 1];
      s[4] += t1;
      s[0] = t1 + t2;
    }
    for(int i = 0; i < 8; i++) h[i] += s[i];
  }
  for(int i = 0; i < 4; i++) out[i] = ((uint64_t)h[2 * i] << 32) | h[2 * i + 1];
}

// Rand(seed) reduced to what initHash needs: a stream of uint64 built from two uint32 draws.
class SeededStream {
 public:
  explicit SeededStream(const std::string& seed) {
    std::string suffix = "|" + std::to_string(md5FirstWord(seed)) + "|" + seed;
    int counter = 0, have = 0;
    uint64_t block[4];
    auto nextNonzero = [&]() {
      for(;;) {
        if(have == 0) {
          sha256Words(std::to_string(counter) + suffix, block);
          counter += 37;
          have = 4;
        }
        uint64_t v = block[4 - have];
        have--;
        if(v != 0) return v;
      }
    };
    for(int i = 0; i < 16; i++) xs_[i] = nextNonzero();
    idx_ = 0;
    pcg_ = nextNonzero();
  }
  uint64_t next64() {
    uint64_t lo = next32();
    uint64_t hi = next32();
    return lo | (hi << 32);
  }

 private:
  uint32_t next32() {
    // PCG32 step (rand_helpers.h:60-66)
    pcg_ = pcg_ * 6364136223846793005ULL + 1442695040888963407ULL;
    uint32_t x = (uint32_t)(((pcg_ >> 18) ^ pcg_) >> 27);
    unsigned rot = (unsigned)(pcg_ >> 59);
    uint32_t p = rot ? ror(x, rot) : x;
    // xorshift1024* step (rand_helpers.h:29-41)
    uint64_t s0 = xs_[idx_];
    idx_ = (idx_ + 1) & 15;
    uint64_t s1 = xs_[idx_];
    s1 ^= s1 << 31;
    s1 ^= s1 >> 11;
    s0 ^= s0 >> 30;
    xs_[idx_] = s0 ^ s1;
    uint32_t q = (uint32_t)((xs_[idx_] * 1181783497276652981ULL) >> 32);
    return p + q;
  }
  uint64_t xs_[16];
  unsigned idx_;
  uint64_t pcg_;
};

ZobristTables* buildTables() {
  ZobristTables* t = new ZobristTables();
  memset(t, 0, sizeof(*t));
  {
    SeededStream s("Board::initHash()");
    for(int i = 0; i < 4; i++) { t->player[i][0] = s.next64(); t->player[i][1] = s.next64(); }
    for(int spot = 0; spot < KC_MAX_ARR_SIZE; spot++)
      for(int color = 1; color <= 2; color++) {  // empty (0) and wall (3) stay zero
        t->board[spot][color][0] = s.next64();
        t->board[spot][color][1] = s.next64();
      }
  }
  {
    SeededStream s("Board::initHash() for ZOBRIST_SIZE hashes");
    for(int i = 0; i <= KC_MAX_LEN; i++) {
      t->sizeX[i][0] = s.next64(); t->sizeX[i][1] = s.next64();
      t->sizeY[i][0] = s.next64(); t->sizeY[i][1] = s.next64();
    }
  }
  return t;
}

}  // namespace

const ZobristTables& zobrist() {
  static ZobristTables* tables = buildTables();
  return *tables;
}

void symmetryDstOfSrc(int H, int W, int symmetry, bool reverse, int* dstOfSrc) {
  bool transpose = (symmetry & 4) != 0 && H == W;
  bool flipX = (symmetry & 2) != 0, flipY = (symmetry & 1) != 0;
  if(transpose && !reverse) { bool t = flipX; flipX = flipY; flipY = t; }
  for(int h = 0; h < H; h++)
    for(int w = 0; w < W; w++) {
      // destination strides in units of cells: start from (row stride W, col stride 1), negate the
      // flipped ones around the far edge, then exchange them when transposing
      int rowStep = W, colStep = 1, base = 0;
      if(flipY) { base += (H - 1) * rowStep; rowStep = -rowStep; }
      if(flipX) { base += (W - 1) * colStep; colStep = -colStep; }
      if(transpose) { int t = rowStep; rowStep = colStep; colStep = t; }
      dstOfSrc[h * W + w] = base + h * rowStep + w * colStep;
    }
}

int symDir(int dir, int symmetry) {
  if(dir < 0 || dir > 3) return dir;
  bool tr = (symmetry & 4) != 0, fx = (symmetry & 2) != 0, fy = (symmetry & 1) != 0;
  if(fx != fy) dir = dir == 2 ? 3 : (dir == 3 ? 2 : dir);
  if(tr) dir = dir == 0 ? 1 : (dir == 1 ? 0 : dir);
  return dir;
}

}  // namespace kc

extern "C" {
const char* kc_last_error(void) { return kc::lastErrorCStr(); }
int kc_abi_version(void) { return 1; }
int kc_zobrist_tables(uint64_t* board, uint64_t* player, uint64_t* size_x, uint64_t* size_y) {
  const kc::ZobristTables& z = kc::zobrist();
  if(board) memcpy(board, z.board, sizeof(z.board));
  if(player) memcpy(player, z.player, sizeof(z.player));
  if(size_x) memcpy(size_x, z.sizeX, sizeof(z.sizeX));
  if(size_y) memcpy(size_y, z.sizeY, sizeof(z.sizeY));
  return 0;
}
}
