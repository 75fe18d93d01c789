// Coffee rules, history, sit-hash and V1 planes for boards BEYOND 7x7 -- up to the reference's maximum of 10x10
// (cpp/game/board.h:120: MAX_LEN = 10) -- where a padded bitboard (bit = y*(W+1) + x, at most 10*11 = 110 bits) no longer fits
// one 64-bit register.  Same semantics, same outputs and the same reference lines as games_device.cuh / games.cu; this is the general
// path, written for breadth rather than speed: 128-bit bitboards as two 64-bit words, one thread per game for the rules, the CTA
// for the plane stores.  The hot configurations (5x5, 6x6, everything up to 7x7) never come here.
//
// State: black / white low words in State::black / white, high words in State::blackHi / whiteHi.  A cell index needs 7 bits on these
// boards, so `misc` packs the history differently from the small-board layout: five 9-bit entries (bits 0-6 dense cell, 7-8 player,
// 0 = none), most recent first, in bits 0-44; bits 45-47 the last direction (4 = none); byte 6 numTurns and byte 7 the flags as
// everywhere else.
#pragma once
#include "games_device.cuh"

namespace kc {

struct B128 {
  uint64_t lo, hi;
  B128() = default;   // trivial (it lives in shared memory); value-initialise -- B128() / {} -- for an empty board
  __host__ __device__ __forceinline__ B128(uint64_t l, uint64_t h) : lo(l), hi(h) {}
  __host__ __device__ __forceinline__ B128 operator&(const B128& o) const { return B128(lo & o.lo, hi & o.hi); }
  __host__ __device__ __forceinline__ B128 operator|(const B128& o) const { return B128(lo | o.lo, hi | o.hi); }
  __host__ __device__ __forceinline__ B128 operator~() const { return B128(~lo, ~hi); }
  __host__ __device__ __forceinline__ B128& operator&=(const B128& o) { lo &= o.lo; hi &= o.hi; return *this; }
  __host__ __device__ __forceinline__ B128& operator|=(const B128& o) { lo |= o.lo; hi |= o.hi; return *this; }
  __host__ __device__ __forceinline__ B128 operator<<(int s) const {   // 0 <= s < 128
    if(s == 0) return *this;
    if(s >= 64) return B128(0, lo << (s - 64));
    return B128(lo << s, (hi << s) | (lo >> (64 - s)));
  }
  __host__ __device__ __forceinline__ B128 operator>>(int s) const {
    if(s == 0) return *this;
    if(s >= 64) return B128(hi >> (s - 64), 0);
    return B128((lo >> s) | (hi << (64 - s)), hi >> s);
  }
  __host__ __device__ __forceinline__ bool any() const { return (lo | hi) != 0; }
  __host__ __device__ __forceinline__ bool test(int b) const { return ((b < 64 ? lo >> b : hi >> (b - 64)) & 1ULL) != 0; }
  __host__ __device__ static __forceinline__ B128 bit(int b) { return b < 64 ? B128(1ULL << b, 0) : B128(0, 1ULL << (b - 64)); }
};
__device__ __forceinline__ int popcB(const B128& v) { return __popcll(v.lo) + __popcll(v.hi); }
__device__ __forceinline__ int nthSetBitB(const B128& m, int n) {   // n is 0-based
  const int c = __popcll(m.lo);
  return n < c ? nthSetBit(m.lo, n) : 64 + nthSetBit(m.hi, n - c);
}

struct BigGeom {   // device memory: the masks the small boards keep in the kernel parameter block
  B128 all;
  B128 lines[4][2 * KC_MAX_LEN - 1];   // [dir][line index]: N: x, W: y, NW: x-y+H-1, NE: x+y
};

__device__ __forceinline__ int bigHistCell(uint64_t misc, int i) { return (int)((misc >> (9 * i)) & 0x7f); }
__device__ __forceinline__ int bigHistPla(uint64_t misc, int i) { return (int)((misc >> (9 * i + 7)) & 0x3); }
__device__ __forceinline__ int bigLastDir(uint64_t misc) { return (int)((misc >> 45) & 0x7); }
constexpr uint64_t BIG_MISC_START = (4ULL << 45) | ((uint64_t)(1 << 3) << 56);   // no history, no last direction, black to move

struct BigRegs { B128 black, white; uint64_t h0, h1, id, misc; };

__device__ __forceinline__ int bigShift(const Geom& g, int d) { return d == 0 ? g.stride : d == 1 ? 1 : d == 2 ? g.stride + 1 : g.stride - 1; }

// Board::isLegal for every Loc (board.cpp:185-227): empty, on the forced line if there is one, and another empty cell on its own line
__device__ inline void bigLegalMasks(const Geom& g, const BigGeom& bg, const B128& empty, int lastCell, int lastDir, B128 L[4]) {
  B128 cand = empty;
  if(lastDir < 4 && lastCell >= 0) {
    const int x = lastCell % g.W, y = lastCell / g.W;
    const int li = lastDir == 0 ? x : lastDir == 1 ? y : lastDir == 2 ? (x - y + g.H - 1) : (x + y);
    cand &= bg.lines[lastDir][li];
  }
  for(int d = 0; d < 4; d++) {
    const int nl = d == 0 ? g.W : d == 1 ? g.H : g.W + g.H - 1;
    B128 ok{};
    for(int i = 0; i < nl; i++) {
      const B128 e = empty & bg.lines[d][i];
      if(popcB(e) >= 2) ok |= e;
    }
    L[d] = cand & ok;
  }
}

// cells covered by a same-colour run of length >= n along shift s
__device__ inline B128 bigCoverAtLeast(const B128& m, int s, int n) {
  B128 starts = m;
  for(int i = 1; i < n; i++) starts &= (m >> (i * s));
  B128 c = starts;
  for(int i = 1; i < n; i++) c |= (starts << (i * s));
  return c;
}

// One ply (stepGame of games_device.cuh on 128-bit boards).  Returns the policy index played (-1 none); L = legal masks afterwards.
__device__ inline int bigStepGame(const Geom& g, const BigGeom& bg, BigRegs& s, int forcedMove, bool useForced, const uint64_t* __restrict__ zob,
                                  B128 L[4], bool& illegal) {
  illegal = false;
  int fl = flagsOf(s.misc);
  if((fl & 1) && g.autoRefill) {
    s.black = B128(); s.white = B128(); s.h0 = g.sizeHash[0]; s.h1 = g.sizeHash[1];
    s.id = s.id + (uint64_t)g.numGames; s.misc = BIG_MISC_START;
    fl = flagsOf(s.misc);
  }
  const int pla = (fl >> 3) & 3;
  const B128 empty = bg.all & ~(s.black | s.white);
  const int lastCell = bigHistPla(s.misc, 0) ? bigHistCell(s.misc, 0) : -1;
  bigLegalMasks(g, bg, empty, lastCell, bigLastDir(s.misc), L);
  if(fl & 1) return -1;
  int dir = -1, cellPad = 0;
  if(useForced && forcedMove != -2) {
    if(forcedMove < 0) return -1;
    if(forcedMove >= 4 * g.HW) { illegal = true; return -1; }
    dir = forcedMove / g.HW;
    const int cell = forcedMove % g.HW;   // ledger I
    cellPad = cell + cell / g.W;
    if(!L[dir].test(cellPad)) { illegal = true; return -1; }
  } else {
    const int c0 = popcB(L[0]), c1 = popcB(L[1]), c2 = popcB(L[2]), c3 = popcB(L[3]);
    const int n = c0 + c1 + c2 + c3;
    if(n == 0) return -1;
    const uint64_t r = splitmix64(g.seed ^ (s.id * 0x9E3779B97F4A7C15ULL) ^ (uint64_t)numTurnsOf(s.misc));
    int k = (int)(r % (uint64_t)n);
    if(k < c0) dir = 0;
    else if(k < c0 + c1) { dir = 1; k -= c0; }
    else if(k < c0 + c1 + c2) { dir = 2; k -= c0 + c1; }
    else { dir = 3; k -= c0 + c1 + c2; }
    cellPad = nthSetBitB(L[dir], k);
  }
  const int y = cellPad / g.stride, x = cellPad - y * g.stride;
  const int cell = y * g.W + x;
  const B128 bit = B128::bit(cellPad);
  if(pla == 1) s.black |= bit; else s.white |= bit;                       // board.cpp:427-435
  const uint64_t* z = zob + ((size_t)cell * 2 + (pla - 1)) * 2;
  s.h0 ^= z[0]; s.h1 ^= z[1];
  const uint64_t hist = ((s.misc & ((1ULL << 36) - 1)) << 9) | (uint64_t)(cell | (pla << 7));   // boardhistory.cpp:157-176: four older entries move up
  const int nt = numTurnsOf(s.misc) + 1;
  const B128 mine = pla == 1 ? s.black : s.white;
  bool win = false;                                                       // board.cpp:376-383, overlines count (ledger N)
  for(int d = 0; d < 4; d++) win = win || (bigCoverAtLeast(mine, bigShift(g, d), g.K) & bit).any();
  const int opp = pla ^ 3;
  const B128 empty2 = empty & ~bit;
  bigLegalMasks(g, bg, empty2, cell, dir, L);
  const bool none = !(L[0] | L[1] | L[2] | L[3]).any();                   // ledger C: the player to move has no legal Loc
  const int finished = (win || none) ? 1 : 0;
  const int winner = win ? pla : 0;
  const int nfl = finished | (winner << 1) | (opp << 3);
  s.misc = (hist & ((1ULL << 45) - 1)) | ((uint64_t)dir << 45) | ((uint64_t)(nt & 0xff) << 48) | ((uint64_t)nfl << 56);
  return dir * g.HW + cell;
}

// The 15 V1 planes as padded 128-bit bitboards (nninputs.cpp:508-657, ledger F / G)
__device__ inline void bigV1Planes(const Geom& g, const BigRegs& s, const B128 L[4], B128* P, int pstride, const B128& all) {
  const int fl = flagsOf(s.misc);
  const int pla = (fl >> 3) & 3, opp = pla ^ 3;
  const B128 own = pla == 1 ? s.black : s.white, other = pla == 1 ? s.white : s.black;
  const int nt = numTurnsOf(s.misc);
  auto cellBit = [&](int cell) { return B128::bit(cell + cell / g.W); };
  P[0 * pstride] = all;
  P[1 * pstride] = own;
  P[2 * pstride] = other;
  const B128 lastBit = bigHistPla(s.misc, 0) ? cellBit(bigHistCell(s.misc, 0)) : B128();
  const int ld = bigLastDir(s.misc);
  for(int d = 0; d < 4; d++) P[(3 + d) * pstride] = ld == d ? lastBit : B128();
  bool ok = true;
  for(int i = 1; i < 5; i++) {
    const int want = (i & 1) ? pla : opp;
    ok = ok && nt >= i + 1 && bigHistPla(s.misc, i) == want;
    P[(6 + i) * pstride] = ok ? cellBit(bigHistCell(s.misc, i)) : B128();
  }
  P[11 * pstride] = L[0] | L[1] | L[2] | L[3];
  B128 ex[3] = {};
  for(int cd = 0; cd < 8; cd++) {
    const int d = cd & 3;
    const B128 m = (cd & 4) ? s.white : s.black;
    const int shift = bigShift(g, d);
    B128 hi = bigCoverAtLeast(m, shift, g.K);
    for(int j = 0; j < 3; j++) {
      const int len = g.K - 1 - j;
      const B128 lo = len >= 1 ? bigCoverAtLeast(m, shift, len) : B128();
      ex[j] |= lo & ~hi;
      hi = lo;
    }
  }
  P[12 * pstride] = ex[0];
  P[13 * pstride] = ex[1];
  P[14 * pstride] = ex[2];
}

constexpr int TB_BIG = 64;
constexpr int BIG_MAX_HW = KC_MAX_LEN * KC_MAX_LEN;

// feat: 0 no planes, 1 fp32 NCHW, 2 fp32 NHWC (with the optional per-game symmetry of copyInputsWithSymmetry, nninputs.cpp:252-357)
template <bool DO_STEP>
__global__ void __launch_bounds__(TB_BIG) games_big_kernel(const Geom g, State st, const BigGeom* __restrict__ bgp, const int16_t* __restrict__ moves, int useMoves,
                                                           const uint64_t* __restrict__ zob, StepOut so, float* __restrict__ planes, float* __restrict__ global,
                                                           const int8_t* __restrict__ symmetry, int permuteDirs, int feat) {
  __shared__ BigGeom bg;
  __shared__ B128 sPl[15][TB_BIG];
  __shared__ uint8_t sSrcPad[8][BIG_MAX_HW];   // [symmetry][dst cell] -> padded bit index of the source cell
  __shared__ int8_t sSym[TB_BIG];
  const int t = threadIdx.x;
  for(int i = t; i < (int)(sizeof(BigGeom) / 8); i += TB_BIG) reinterpret_cast<uint64_t*>(&bg)[i] = reinterpret_cast<const uint64_t*>(bgp)[i];
  const int gBase = blockIdx.x * TB_BIG;
  const int gi = gBase + t;
  const bool active = gi < g.numGames;
  if(feat != 0) {
    for(int i = t; i < 8 * g.HW; i += TB_BIG) {
      const int sym = i / g.HW, cell = i % g.HW;
      const int h = cell / g.W, w = cell % g.W;
      bool tr = (sym & 4) && g.H == g.W, fx = (sym & 2) != 0, fy = (sym & 1) != 0;
      if(tr) { const bool tmp = fx; fx = fy; fy = tmp; }
      int rowStep = g.W, colStep = 1, base = 0;
      if(fy) { base += (g.H - 1) * rowStep; rowStep = -rowStep; }
      if(fx) { base += (g.W - 1) * colStep; colStep = -colStep; }
      if(tr) { const int tmp = rowStep; rowStep = colStep; colStep = tmp; }
      sSrcPad[sym][base + h * rowStep + w * colStep] = (uint8_t)(cell + cell / g.W);
    }
    sSym[t] = (active && symmetry) ? symmetry[gi] : 0;
  }
  __syncthreads();
  unsigned long long cSteps = 0, cFin = 0, cB = 0, cW = 0, cD = 0, cXor = 0;
  if(active) {
    BigRegs s;
    s.black = B128(st.black[gi], st.blackHi[gi]); s.white = B128(st.white[gi], st.whiteHi[gi]);
    s.h0 = st.hash0[gi]; s.h1 = st.hash1[gi]; s.id = st.gameId[gi]; s.misc = st.misc[gi];
    B128 L[4];
    int played = -1;
    bool illegal = false;
    if(DO_STEP) {
      const int mv = useMoves ? (int)moves[gi] : -1;
      played = bigStepGame(g, bg, s, mv, useMoves != 0, zob, L, illegal);
      st.black[gi] = s.black.lo; st.blackHi[gi] = s.black.hi; st.white[gi] = s.white.lo; st.whiteHi[gi] = s.white.hi;
      st.hash0[gi] = s.h0; st.hash1[gi] = s.h1; st.gameId[gi] = s.id; st.misc[gi] = s.misc;
    } else {
      const B128 empty = bg.all & ~(s.black | s.white);
      const int lastCell = bigHistPla(s.misc, 0) ? bigHistCell(s.misc, 0) : -1;
      bigLegalMasks(g, bg, empty, lastCell, bigLastDir(s.misc), L);
    }
    const int fl = flagsOf(s.misc);
    const int nextPla = (fl >> 3) & 3;
    const uint64_t sh0 = s.h0 ^ g.playerHash[nextPla][0], sh1 = s.h1 ^ g.playerHash[nextPla][1];
    if(so.status) so.status[gi] = (uint32_t)numTurnsOf(s.misc) | ((uint32_t)(fl & 1) << 8) | ((uint32_t)((fl >> 1) & 3) << 9) |
                                  ((uint32_t)nextPla << 11) | (illegal ? (1u << 15) : 0u);
    if(so.sitHash) { so.sitHash[2 * (size_t)gi] = sh0; so.sitHash[2 * (size_t)gi + 1] = sh1; }
    if(so.played) so.played[gi] = (int16_t)played;
    if(so.legal) {   // policy order: bit = dir*HW + y*W + x (nninputs.cpp:6-14)
      uint32_t* out = so.legal + (size_t)gi * g.LW;
      uint32_t word = 0;
      int wi = 0;
      for(int pos = 0; pos < 4 * g.HW; pos++) {
        const int d = pos / g.HW, cell = pos - d * g.HW;
        if(L[d].test(cell + cell / g.W)) word |= 1u << (pos & 31);
        if((pos & 31) == 31) { out[wi++] = word; word = 0; }
      }
      if((4 * g.HW) & 31) out[wi] = word;
    }
    if(DO_STEP && played >= 0) {
      cSteps = 1;
      cXor = sh0;
      if(fl & 1) { cFin = 1; const int w = (fl >> 1) & 3; cB = w == 1; cW = w == 2; cD = w == 0; }
    }
    if(feat != 0) bigV1Planes(g, s, L, &sPl[0][t], TB_BIG, bg.all);
  }
  if(so.stats) {
    for(int o = 16; o > 0; o >>= 1) {
      cSteps += __shfl_xor_sync(0xffffffffu, cSteps, o); cFin += __shfl_xor_sync(0xffffffffu, cFin, o);
      cB += __shfl_xor_sync(0xffffffffu, cB, o); cW += __shfl_xor_sync(0xffffffffu, cW, o);
      cD += __shfl_xor_sync(0xffffffffu, cD, o); cXor ^= __shfl_xor_sync(0xffffffffu, cXor, o);
    }
    if((t & 31) == 0) {
      if(cSteps) atomicAdd(&so.stats[0], cSteps);
      if(cFin) atomicAdd(&so.stats[2], cFin);
      if(cB) atomicAdd(&so.stats[3], cB);
      if(cW) atomicAdd(&so.stats[4], cW);
      if(cD) atomicAdd(&so.stats[5], cD);
      if(cXor) atomicXor(&so.stats[6], cXor);
    }
  }
  if(feat == 0) return;
  __syncthreads();
  // the CTA writes the planes of its games: consecutive threads, consecutive floats
  const int ng = min(TB_BIG, g.numGames - gBase);
  const int E = 15 * g.HW;
  float* out = planes + (size_t)gBase * E;
  for(int e = t; e < ng * E; e += TB_BIG) {
    const int gl = e / E, r = e - gl * E;
    int c, cell;
    if(feat == 1) { c = r / g.HW; cell = r - c * g.HW; } else { cell = r / 15; c = r - cell * 15; }
    const int sym = sSym[gl];
    if(permuteDirs) c = playModeChannel(c, sym);   // destination channel c shows the source's channel playModeChannel(c)
    out[e] = sPl[c][gl].test(sSrcPad[sym][cell]) ? 1.0f : 0.0f;
  }
  if(global && t < ng) global[gBase + t] = (float)g.K;
}

}  // namespace kc
