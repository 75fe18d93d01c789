// Device-side Coffee rules shared by the batched games kernel (games.cu) and the batched tree search (search.cu):
// bitboard geometry, per-game state, legal Locs, one ply, the V1 plane bitboards.  Reference semantics and the
// canonical readings are listed at the top of games.cu.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace kc {

struct Geom {
  int W, H, K, HW, stride;   // stride = W + 1
  int LW;                    // legal mask words = ceil(4*HW/32)
  int numGames;
  uint64_t all;              // on-board cells (padded layout)
  uint64_t rowMask;          // (1<<W)-1
  uint64_t lines[4][13];     // [dir][line index]: N: x, W: y, NW: x-y+H-1, NE: x+y
  uint64_t playerHash[4][2]; // ZOBRIST_PLAYER_HASH
  uint64_t sizeHash[2];      // SIZE_X[W] ^ SIZE_Y[H]
  uint64_t seed;
  int autoRefill;
  // bf16 tile layout for the trunk (see net.h): boards side by side, NB per 128-row tile
  int NB, tileRowW;          // tileRowW = NB*(W+1)
};

struct State {   // SoA device arrays, one entry per game lane
  uint64_t* black; uint64_t* white; uint64_t* hash0; uint64_t* hash1; uint64_t* gameId; uint64_t* misc;
  uint64_t* blackHi; uint64_t* whiteHi;   // boards beyond 7x7 only (games_big.cuh): bits 64.. of the padded bitboards, else null
};
// misc: bytes 0..4 = last five moves, most recent first: bits 0-5 dense cell, bits 6-7 player (0 = none)
//       byte 5 = direction of the most recent move (4 = none), byte 6 = numTurns,
//       byte 7 = bit0 finished, bits 1-2 winner, bits 3-4 next player

struct StepOut {   // device pointers, any may be null
  uint32_t* legal; uint32_t* status; uint64_t* sitHash; int16_t* played; unsigned long long* stats;
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {  // cpp/core/hash.cpp:50-56
  x += 0x9e3779b97f4a7c15ULL;
  x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ULL;
  x = (x ^ (x >> 27)) * 0x94d049bb133111ebULL;
  return x ^ (x >> 31);
}

// Board dimensions: compile-time for the hot 5x5 k=4 configuration (30-bit bitboards in 32-bit registers,
// every loop unrolls into straight-line code), run-time otherwise (64-bit bitboards, rolled loops: the
// generic kernel is instruction-cache sensitive).
template <bool Narrow> struct StaticBB { using type = uint32_t; };
template <> struct StaticBB<false> { using type = uint64_t; };
template <int CW, int CH, int CK>
struct StaticDims {
  static_assert(CH * (CW + 1) <= 64, "static boards must fit a 64-bit bitboard");
  static constexpr bool kStatic = true;
  static constexpr int kUnroll = 16;
  using BB = typename StaticBB<(CH * (CW + 1) <= 32)>::type;   // 5x5: 30 bits in a 32-bit register; 6x6: 42 bits
  __device__ __forceinline__ explicit StaticDims(const Geom&) {}
  __device__ __forceinline__ constexpr int W() const { return CW; }
  __device__ __forceinline__ constexpr int H() const { return CH; }
  __device__ __forceinline__ constexpr int K() const { return CK; }
  __device__ __forceinline__ constexpr int HW() const { return CW * CH; }
  __device__ __forceinline__ constexpr int stride() const { return CW + 1; }
};
struct DynDims {
  static constexpr bool kStatic = false;
  static constexpr int kUnroll = 1;
  using BB = uint64_t;
  int w, h, k;
  __device__ __forceinline__ explicit DynDims(const Geom& g) : w(g.W), h(g.H), k(g.K) {}
  __device__ __forceinline__ int W() const { return w; }
  __device__ __forceinline__ int H() const { return h; }
  __device__ __forceinline__ int K() const { return k; }
  __device__ __forceinline__ int HW() const { return w * h; }
  __device__ __forceinline__ int stride() const { return w + 1; }
};

__device__ __forceinline__ int popcBB(uint32_t v) { return __popc(v); }
__device__ __forceinline__ int popcBB(uint64_t v) { return __popcll(v); }
__device__ __forceinline__ int nthSetBit(uint32_t m, int n) { return __fns(m, 0, n + 1); }   // n is 0-based
__device__ __forceinline__ int nthSetBit(uint64_t m, int n) {
  uint32_t lo = (uint32_t)m, hi = (uint32_t)(m >> 32);
  int c = __popc(lo);
  if(n < c) return __fns(lo, 0, n + 1);
  return 32 + __fns(hi, 0, n - c + 1);
}

template <class D>
__device__ __forceinline__ int padOf(const D& dm, int cell) { return cell + cell / dm.W(); }
template <class D>
__device__ __forceinline__ int shiftOf(const D& dm, int d) {
  return d == 0 ? dm.stride() : d == 1 ? 1 : d == 2 ? dm.stride() + 1 : dm.stride() - 1;
}

// Legal Locs of the player to move, one padded bitboard per direction (board.cpp:185-227).
template <class D>
__device__ __forceinline__ void legalMasks(const D& dm, const Geom& g, typename D::BB empty, int lastCell, int lastDir, typename D::BB L[4]) {
  using BB = typename D::BB;
  BB cand = empty;
  if(lastDir < 4 && lastCell >= 0) {
    int x = lastCell % dm.W(), y = lastCell / dm.W();
    int li = lastDir == 0 ? x : lastDir == 1 ? y : lastDir == 2 ? (x - y + dm.H() - 1) : (x + y);
    cand &= (BB)g.lines[lastDir][li];
  }
#pragma unroll D::kUnroll
  for(int d = 0; d < 4; d++) {
    int nl = (d == 0) ? dm.W() : (d == 1) ? dm.H() : (dm.W() + dm.H() - 1);
    BB ok = 0;
#pragma unroll D::kUnroll
    for(int i = 0; i < nl; i++) {
      BB e = empty & (BB)g.lines[d][i];
      ok |= (popcBB(e) >= 2) ? e : (BB)0;   // another empty cell anywhere on the same line
    }
    L[d] = cand & ok;
  }
}

template <class D>
__device__ __forceinline__ uint64_t toDense(const D& dm, typename D::BB m) {
  using BB = typename D::BB;
  const BB rowMask = ((BB)1 << dm.W()) - 1;
  uint64_t r = 0;
#pragma unroll D::kUnroll
  for(int y = 0; y < dm.H(); y++) r |= (uint64_t)((m >> (y * dm.stride())) & rowMask) << (y * dm.W());
  return r;
}

// cells covered by a same-colour run of length >= n along shift s
template <class D>
__device__ __forceinline__ typename D::BB coverAtLeast(typename D::BB m, int s, int n) {
  typename D::BB starts = m;
#pragma unroll D::kUnroll
  for(int i = 1; i < n; i++) starts &= (m >> (i * s));
  typename D::BB c = starts;
#pragma unroll D::kUnroll
  for(int i = 1; i < n; i++) c |= (starts << (i * s));
  return c;
}

template <class BB>
struct GameRegs {
  BB black, white;
  uint64_t h0, h1, id, misc;
};

__device__ __forceinline__ int histCell(uint64_t misc, int i) { return (int)((misc >> (8 * i)) & 0x3f); }
__device__ __forceinline__ int histPla(uint64_t misc, int i) { return (int)((misc >> (8 * i + 6)) & 0x3); }
__device__ __forceinline__ int lastDirOf(uint64_t misc) { return (int)((misc >> 40) & 0xff); }
__device__ __forceinline__ int numTurnsOf(uint64_t misc) { return (int)((misc >> 48) & 0xff); }
__device__ __forceinline__ int flagsOf(uint64_t misc) { return (int)((misc >> 56) & 0xff); }

template <class BB>
__device__ __forceinline__ void resetGame(const Geom& g, GameRegs<BB>& s, uint64_t id) {
  s.black = 0; s.white = 0;
  s.h0 = g.sizeHash[0]; s.h1 = g.sizeHash[1];
  s.id = id;
  // no history, lastDir none, numTurns 0, not finished, winner 0, next player black (boardhistory.cpp:7-18)
  s.misc = (4ULL << 40) | ((uint64_t)(1 << 3) << 56);
}

// The 15 V1 planes as padded bitboards (nninputs.cpp:508-657, ledger F/G).
template <class D>
__device__ __forceinline__ void v1Planes(const D& dm, const Geom& g, const GameRegs<typename D::BB>& s, const typename D::BB L[4],
                                         uint64_t* P /*[15], stride pstride*/, int pstride) {
  using BB = typename D::BB;
  int fl = flagsOf(s.misc);
  int pla = (fl >> 3) & 3, opp = pla ^ 3;
  BB own = pla == 1 ? s.black : s.white, other = pla == 1 ? s.white : s.black;
  int nt = numTurnsOf(s.misc);
  P[0 * pstride] = g.all;
  P[1 * pstride] = own;
  P[2 * pstride] = other;
  uint64_t lastBit = histPla(s.misc, 0) ? (1ULL << padOf(dm, histCell(s.misc, 0))) : 0ULL;
  int ld = lastDirOf(s.misc);
#pragma unroll
  for(int d = 0; d < 4; d++) P[(3 + d) * pstride] = (ld == d) ? lastBit : 0ULL;
  // moves 2..5 plies ago: alternation chain own, opp, own, opp; breaks at the first failure
  bool ok = true;
#pragma unroll
  for(int i = 1; i < 5; i++) {
    int want = (i & 1) ? pla : opp;
    ok = ok && nt >= i + 1 && histPla(s.misc, i) == want;
    P[(6 + i) * pstride] = ok ? (1ULL << padOf(dm, histCell(s.misc, i))) : 0ULL;
  }
  P[11 * pstride] = L[0] | L[1] | L[2] | L[3];
  // stones in a maximal same-colour run of length exactly k-1, k-2, k-3 in some direction
  BB ex[3] = {0, 0, 0};
#pragma unroll D::kUnroll
  for(int cd = 0; cd < 8; cd++) {
    const int d = cd & 3;
    const BB m = (cd & 4) ? s.white : s.black;
    const int shift = shiftOf(dm, d);
    BB hi = coverAtLeast<D>(m, shift, dm.K());      // >= k
#pragma unroll
    for(int j = 0; j < 3; j++) {
      int len = dm.K() - 1 - j;
      BB lo = len >= 1 ? coverAtLeast<D>(m, shift, len) : (BB)0;
      ex[j] |= lo & ~hi;
      hi = lo;
    }
  }
  P[12 * pstride] = ex[0];
  P[13 * pstride] = ex[1];
  P[14 * pstride] = ex[2];
}

// One ply for one game. Returns the policy index played (-1 none).  L = legal masks afterwards.
template <class D>
__device__ __forceinline__ int stepGame(const D& dm, const Geom& g, GameRegs<typename D::BB>& s, int forcedMove, bool useForced,
                                        const uint64_t* __restrict__ zob, typename D::BB L[4], bool& illegal) {
  using BB = typename D::BB;
  illegal = false;
  int fl = flagsOf(s.misc);
  if((fl & 1) && g.autoRefill) {
    resetGame(g, s, s.id + (uint64_t)g.numGames);
    fl = flagsOf(s.misc);
  }
  int pla = (fl >> 3) & 3;
  BB empty = (BB)g.all & ~(s.black | s.white);
  int lastCell = histPla(s.misc, 0) ? histCell(s.misc, 0) : -1;
  legalMasks(dm, g, empty, lastCell, lastDirOf(s.misc), L);
  if(fl & 1) return -1;                         // finished, not refilled
  int dir = -1, cellPad = 0;
  if(useForced && forcedMove != -2) {   // -2: "play the counter-RNG move" inside a forced-move batch
    if(forcedMove < 0) return -1;
    if(forcedMove >= 4 * dm.HW()) { illegal = true; return -1; }
    dir = forcedMove / dm.HW();
    int cell = forcedMove % dm.HW();               // ledger I
    cellPad = padOf(dm, cell);
    if(!((L[dir] >> cellPad) & (BB)1)) { illegal = true; return -1; }
  } else {
    int c0 = popcBB(L[0]), c1 = popcBB(L[1]), c2 = popcBB(L[2]), c3 = popcBB(L[3]);
    int n = c0 + c1 + c2 + c3;
    if(n == 0) return -1;
    uint64_t r = splitmix64(g.seed ^ (s.id * 0x9E3779B97F4A7C15ULL) ^ (uint64_t)numTurnsOf(s.misc));
    int k = (int)(r % (uint64_t)n);
    if(k < c0) dir = 0;
    else if(k < c0 + c1) { dir = 1; k -= c0; }
    else if(k < c0 + c1 + c2) { dir = 2; k -= c0 + c1; }
    else { dir = 3; k -= c0 + c1 + c2; }
    cellPad = nthSetBit(L[dir], k);
  }
  int y = cellPad / dm.stride(), x = cellPad - y * dm.stride();
  int cell = y * dm.W() + x;
  // play (board.cpp:427-435) and history (boardhistory.cpp:157-176)
  BB bit = (BB)1 << cellPad;
  if(pla == 1) s.black |= bit; else s.white |= bit;
  const uint64_t* z = zob + ((size_t)cell * 2 + (pla - 1)) * 2;
  s.h0 ^= z[0]; s.h1 ^= z[1];
  uint64_t hist = ((s.misc & 0xffffffffULL) << 8) | (uint64_t)(cell | (pla << 6));
  int nt = numTurnsOf(s.misc) + 1;
  // win through the last move (board.cpp:376-383), overlines count
  BB mine = pla == 1 ? s.black : s.white;
  bool win = false;
#pragma unroll D::kUnroll
  for(int d = 0; d < 4; d++) win = win || ((coverAtLeast<D>(mine, shiftOf(dm, d), dm.K()) & bit) != 0);
  int opp = pla ^ 3;
  // legal masks of the player now to move (also decides the draw, ledger C)
  BB empty2 = empty & ~bit;
  legalMasks(dm, g, empty2, cell, dir, L);
  bool none = (L[0] | L[1] | L[2] | L[3]) == 0;
  int finished = (win || none) ? 1 : 0;
  int winner = win ? pla : 0;
  int nfl = finished | (winner << 1) | (opp << 3);
  s.misc = (hist & 0xffffffffffULL) | ((uint64_t)dir << 40) | ((uint64_t)(nt & 0xff) << 48) | ((uint64_t)nfl << 56);
  return dir * dm.HW() + cell;
}

}  // namespace kc
