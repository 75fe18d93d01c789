// ORACLE (test infrastructure only, see kc_oracle.h): Coffee rules, history, sit-hash, V1 features,
// symmetry helpers and the synthetic random-legal playout.
//
// Mailbox restatement that follows the reference line by line (deliberately NOT the bitboard
// formulation the CUDA kernels use, so the two are independent implementations):
//   cpp/game/board.h:24-52,74-75,120-124   types, Spot = (x+1)+(y+1)*(x_size+1), MAX_ARR_SIZE
//   cpp/game/board.cpp:82-85               ADJ1..4 = N, W, NW, NE mailbox offsets
//   cpp/game/board.cpp:111-132             Board::init
//   cpp/game/board.cpp:185-227             Board::isLegal
//   cpp/game/board.cpp:280-292,427-435     playMove / getSitHash / playMoveAssumeLegal
//   cpp/game/board.cpp:315-335,376-383     maxConsecutives / checkGameEnd
//   cpp/game/boardhistory.cpp:142-176      makeBoardMove / makeBoardMoveAssumeLegal
//   cpp/neuralnet/nninputs.cpp:6-49        NNPos
//   cpp/neuralnet/nninputs.cpp:252-433     symmetry helpers
//   cpp/neuralnet/nninputs.cpp:463-502     NNInputs::getHash
//   cpp/neuralnet/nninputs.cpp:508-657     NNInputs::fillRowV1 (+ board.cpp:392-420 fillRowWithLine)
// with the canonical resolutions of SURVEY.md section 8.1 (ledger ids quoted inline).
#include "kc_oracle.h"

#include <algorithm>
#include <cassert>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr int C_EMPTY = 0, C_BLACK = 1, C_WHITE = 2, C_WALL = 3;
constexpr int P_BLACK = 1, P_WHITE = 2;
constexpr int D_NONE = 4;
constexpr int NULL_LOC = 0;
constexpr int MAX_ARR = KO_MAX_ARR_SIZE;

struct Zobrist {
  uint64_t board[MAX_ARR][4][2];
  uint64_t player[4][2];
  uint64_t sizeX[KO_MAX_LEN + 1][2];
  uint64_t sizeY[KO_MAX_LEN + 1][2];
  Zobrist() { ko_zobrist_tables(&board[0][0][0], &player[0][0], &sizeX[0][0], &sizeY[0][0]); }
};
const Zobrist& zob() {
  static const Zobrist z;
  return z;
}
// board.cpp:26-27
constexpr uint64_t GAME_IS_OVER0 = 0xb6f9e465597a77eeULL, GAME_IS_OVER1 = 0xf1d583d960a4ce7fULL;
// nninputs.cpp:54-56
constexpr uint64_t ZPD0 = 0xa5e6114d380bfc1dULL, ZPD1 = 0x4160557f1222f4adULL;
constexpr uint64_t ZPT0 = 0xebcbdfeec6f4334bULL, ZPT1 = 0xb85e43ee243b5ad2ULL;
constexpr uint64_t ZPO0 = 0x88415c85c2801955ULL, ZPO1 = 0x39bdf76b2aaa5eb1ULL;

struct Move {
  int16_t spot;
  int8_t dir;
  int8_t pla;
};

}  // namespace

struct ko_game {
  // Board (board.h:216-222)
  int x_size, y_size, win_len;
  int8_t colors[MAX_ARR];
  int16_t lastSpot;
  int8_t lastDir;
  uint64_t pos_hash[2];
  // BoardHistory (boardhistory.h:12-31); ledger D: every played move is appended exactly once
  std::vector<Move> moveHistory;
  int numTurns;
  int initialPla;
  int presumedNextMovePla;
  bool isGameFinished;
  int winner;

  int spotOf(int x, int y) const { return (x + 1) + (y + 1) * (x_size + 1); }
  int getX(int spot) const { return (spot % (x_size + 1)) - 1; }
  int getY(int spot) const { return (spot / (x_size + 1)) - 1; }
  bool isOnBoard(int spot) const { return spot >= 0 && spot < MAX_ARR && colors[spot] != C_WALL; }
  int adj(int dir) const {
    // board.cpp:82-85
    switch(dir) {
      case 0: return -(x_size + 1);      // N
      case 1: return -1;                 // W
      case 2: return -(x_size + 1) - 1;  // NW
      default: return -(x_size + 1) + 1; // NE
    }
  }

  void init() {
    // board.cpp:111-132
    for(int i = 0; i < MAX_ARR; i++) colors[i] = C_WALL;
    for(int y = 0; y < y_size; y++)
      for(int x = 0; x < x_size; x++) colors[spotOf(x, y)] = C_EMPTY;
    pos_hash[0] = zob().sizeX[x_size][0] ^ zob().sizeY[y_size][0];
    pos_hash[1] = zob().sizeX[x_size][1] ^ zob().sizeY[y_size][1];
    lastSpot = NULL_LOC;
    lastDir = D_NONE;
    // boardhistory.cpp:99-120 (clear), default first player black (boardhistory.cpp:7-18)
    moveHistory.clear();
    numTurns = 0;
    initialPla = P_BLACK;
    presumedNextMovePla = P_BLACK;
    isGameFinished = false;
    winner = C_EMPTY;
  }

  // board.cpp:185-227 with ledger B (dir must be 0..3 and spot must be a board cell)
  bool isLegal(int spot, int dir, int pla) const {
    if(pla != P_BLACK && pla != P_WHITE) return false;
    if(dir < 0 || dir > 3) return false;
    if(spot < 0 || spot >= MAX_ARR) return false;
    if(colors[spot] != C_EMPTY) return false;
    int lastX = getX(lastSpot), lastY = getY(lastSpot);
    int x = getX(spot), y = getY(spot);
    int dx = x - lastX, dy = y - lastY;
    switch(lastDir) {
      case 0: if(dx != 0 || dy == 0) return false; break;
      case 1: if(dx == 0 || dy != 0) return false; break;
      case 2: if(dx != dy) return false; break;
      case 3: if(dx != -dy) return false; break;
      default: break;
    }
    int off = adj(dir);
    int t = spot;
    while(isOnBoard(t)) {
      t += off;
      if(t >= 0 && t < MAX_ARR && colors[t] == C_EMPTY) return true;
    }
    t = spot;
    while(isOnBoard(t)) {
      t -= off;
      if(t >= 0 && t < MAX_ARR && colors[t] == C_EMPTY) return true;
    }
    return false;
  }

  // board.cpp:315-335
  int maxConsecutives(int spot) const {
    int ans = 1;
    int color = colors[spot];
    for(int dir = 0; dir < 4; dir++) {
      int off = adj(dir);
      int c = 1;
      int a = spot - off;
      while(isOnBoard(a) && colors[a] == color) { c++; a -= off; }
      a = spot + off;
      while(isOnBoard(a) && colors[a] == color) { c++; a += off; }
      ans = std::max(ans, c);
    }
    return ans;
  }

  int countLegal(int pla) const {
    int n = 0;
    for(int y = 0; y < y_size; y++)
      for(int x = 0; x < x_size; x++)
        for(int d = 0; d < 4; d++)
          if(isLegal(spotOf(x, y), d, pla)) n++;
    return n;
  }

  // boardhistory.cpp:157-176 + board.cpp:427-435, ledger C (draw) and D (history)
  void makeMoveAssumeLegal(int spot, int dir, int pla) {
    isGameFinished = false;
    winner = C_EMPTY;
    colors[spot] = (int8_t)pla;
    pos_hash[0] ^= zob().board[spot][pla][0];
    pos_hash[1] ^= zob().board[spot][pla][1];
    lastSpot = (int16_t)spot;
    lastDir = (int8_t)dir;
    numTurns += 1;
    moveHistory.push_back(Move{(int16_t)spot, (int8_t)dir, (int8_t)pla});
    presumedNextMovePla = pla ^ 3;
    // board.cpp:376-383: win through the last move, overlines count (ledger N)
    if(maxConsecutives(spot) >= win_len) {
      isGameFinished = true;
      winner = pla;
    } else if(countLegal(pla ^ 3) == 0) {
      // ledger C: the player to move has no legal Loc -> finished without a winner (draw)
      isGameFinished = true;
      winner = C_EMPTY;
    }
  }
};

namespace {

// board.cpp:392-420 fillRowWithLine, canonical reading (ledger G): every stone, each of the 4 line
// directions, maximal same-colour run through it; set iff run length == len.
void fillLinePlane(const ko_game* g, int len, float* plane, int nnXLen, int posStride) {
  if(len <= 0) return;
  for(int y = 0; y < g->y_size; y++) {
    for(int x = 0; x < g->x_size; x++) {
      int spot = g->spotOf(x, y);
      int color = g->colors[spot];
      if(color != C_BLACK && color != C_WHITE) continue;
      for(int dir = 0; dir < 4; dir++) {
        int off = g->adj(dir);
        int c = 1;
        int a = spot - off;
        while(g->isOnBoard(a) && g->colors[a] == color) { c++; a -= off; }
        a = spot + off;
        while(g->isOnBoard(a) && g->colors[a] == color) { c++; a += off; }
        if(c == len) {
          plane[(y * nnXLen + x) * posStride] = 1.0f;
          break;
        }
      }
    }
  }
}

// nninputs.cpp:252-335 copyWithSymmetry
void copyWithSymmetry(const float* src, float* dst, int nSize, int hSize, int wSize, int cSize,
                      bool useNHWC, int symmetry, bool reverse) {
  bool transpose = (symmetry & 0x4) != 0 && hSize == wSize;
  bool flipX = (symmetry & 0x2) != 0;
  bool flipY = (symmetry & 0x1) != 0;
  if(transpose && !reverse) std::swap(flipX, flipY);
  if(useNHWC) {
    int nStride = hSize * wSize * cSize, hStride = wSize * cSize, wStride = cSize;
    int hBaseNew = 0, hStrideNew = hStride, wBaseNew = 0, wStrideNew = wStride;
    if(flipY) { hBaseNew = (hSize - 1) * hStrideNew; hStrideNew = -hStrideNew; }
    if(flipX) { wBaseNew = (wSize - 1) * wStrideNew; wStrideNew = -wStrideNew; }
    if(transpose) std::swap(hStrideNew, wStrideNew);
    for(int n = 0; n < nSize; n++)
      for(int h = 0; h < hSize; h++) {
        int nhOld = n * nStride + h * hStride;
        int nhNew = n * nStride + hBaseNew + h * hStrideNew;
        for(int w = 0; w < wSize; w++) {
          int nhwOld = nhOld + w * wStride;
          int nhwNew = nhNew + wBaseNew + w * wStrideNew;
          for(int c = 0; c < cSize; c++) dst[nhwNew + c] = src[nhwOld + c];
        }
      }
  } else {
    int ncSize = nSize * cSize, ncStride = hSize * wSize, hStride = wSize, wStride = 1;
    int hBaseNew = 0, hStrideNew = hStride, wBaseNew = 0, wStrideNew = wStride;
    if(flipY) { hBaseNew = (hSize - 1) * hStrideNew; hStrideNew = -hStrideNew; }
    if(flipX) { wBaseNew = (wSize - 1) * wStrideNew; wStrideNew = -wStrideNew; }
    if(transpose) std::swap(hStrideNew, wStrideNew);
    for(int nc = 0; nc < ncSize; nc++)
      for(int h = 0; h < hSize; h++) {
        int nchOld = nc * ncStride + h * hStride;
        int nchNew = nc * ncStride + hBaseNew + h * hStrideNew;
        for(int w = 0; w < wSize; w++) dst[nchNew + wBaseNew + w * wStrideNew] = src[nchOld + w * wStride];
      }
  }
}

}  // namespace

extern "C" {

ko_game* ko_game_create(int x_size, int y_size, int win_len) {
  if(x_size < 1 || y_size < 1 || x_size > KO_MAX_LEN || y_size > KO_MAX_LEN) return nullptr;
  ko_game* g = new ko_game();
  g->x_size = x_size;
  g->y_size = y_size;
  g->win_len = win_len;
  g->init();
  return g;
}
void ko_game_destroy(ko_game* g) { delete g; }
void ko_game_reset(ko_game* g) { g->init(); }
void ko_game_copy(ko_game* dst, const ko_game* src) { *dst = *src; }  // ledger A: all fields

int ko_game_set_stone(ko_game* g, int x, int y, int color) {
  if(x < 0 || y < 0 || x >= g->x_size || y >= g->y_size) return 0;
  if(color != C_BLACK && color != C_WHITE && color != C_EMPTY) return 0;
  int spot = g->spotOf(x, y);
  int old = g->colors[spot];
  // keep pos_hash consistent with the stones (the reference setStone forgets the hash; harmless here)
  g->pos_hash[0] ^= zob().board[spot][old][0] ^ zob().board[spot][color][0];
  g->pos_hash[1] ^= zob().board[spot][old][1] ^ zob().board[spot][color][1];
  g->colors[spot] = (int8_t)color;
  return 1;
}
void ko_game_set_last_loc(ko_game* g, int x, int y, int dir) {
  if(x < 0) { g->lastSpot = NULL_LOC; g->lastDir = D_NONE; }
  else { g->lastSpot = (int16_t)g->spotOf(x, y); g->lastDir = (int8_t)dir; }
}
// Test hook: replace the move list (oldest first, pos = policy index) and the turn counter; lastLoc
// follows the most recent move (Loc(NULL_LOC, D_NONE) if the list is empty).
void ko_game_set_history(ko_game* g, int n, const int32_t* pos, const int32_t* pla, int numTurns, int nextPla) {
  int hw = g->x_size * g->y_size;
  g->moveHistory.clear();
  for(int i = 0; i < n; i++) {
    int dir = pos[i] / hw, rem = pos[i] % hw;
    g->moveHistory.push_back(Move{(int16_t)g->spotOf(rem % g->x_size, rem / g->x_size), (int8_t)dir, (int8_t)pla[i]});
  }
  g->numTurns = numTurns;
  g->presumedNextMovePla = nextPla;
  if(n > 0) { g->lastSpot = g->moveHistory.back().spot; g->lastDir = g->moveHistory.back().dir; }
  else { g->lastSpot = NULL_LOC; g->lastDir = D_NONE; }
}
int ko_game_is_legal(const ko_game* g, int x, int y, int dir, int pla) {
  if(x < 0 || y < 0 || x >= g->x_size || y >= g->y_size) return 0;
  return g->isLegal(g->spotOf(x, y), dir, pla) ? 1 : 0;
}
int ko_game_legal_mask(const ko_game* g, int pla, uint32_t* out) {
  int hw = g->x_size * g->y_size;
  int words = (4 * hw + 31) / 32;
  for(int i = 0; i < words; i++) out[i] = 0;
  int n = 0;
  for(int d = 0; d < 4; d++)
    for(int y = 0; y < g->y_size; y++)
      for(int x = 0; x < g->x_size; x++)
        if(g->isLegal(g->spotOf(x, y), d, pla)) {
          int pos = d * hw + y * g->x_size + x;  // nninputs.cpp:6-8
          out[pos >> 5] |= 1u << (pos & 31);
          n++;
        }
  return n;
}
int ko_game_play(ko_game* g, int pos) {
  int hw = g->x_size * g->y_size;
  if(pos < 0 || pos >= 4 * hw) return 0;
  // ledger I: dir = pos / HW, rem = pos % HW
  int dir = pos / hw, rem = pos % hw;
  int x = rem % g->x_size, y = rem / g->x_size;
  int pla = g->presumedNextMovePla;
  int spot = g->spotOf(x, y);
  if(!g->isLegal(spot, dir, pla)) return 0;
  g->makeMoveAssumeLegal(spot, dir, pla);
  return 1;
}
int ko_game_next_pla(const ko_game* g) { return g->presumedNextMovePla; }
int ko_game_num_turns(const ko_game* g) { return g->numTurns; }
int ko_game_finished(const ko_game* g) { return g->isGameFinished ? 1 : 0; }
int ko_game_winner(const ko_game* g) { return g->winner; }
int ko_game_max_consecutives(const ko_game* g, int x, int y) { return g->maxConsecutives(g->spotOf(x, y)); }
int ko_game_color_at(const ko_game* g, int x, int y) { return g->colors[g->spotOf(x, y)]; }
// policy index (dir*H*W + y*W + x) of the k-th most recent move (0 = the last one), -1 if the history is shorter
int ko_game_recent_move_pos(const ko_game* g, int k) {
  const int n = (int)g->moveHistory.size();
  if(k < 0 || k >= n) return -1;
  const Move& m = g->moveHistory[n - 1 - k];
  return (int)m.dir * g->x_size * g->y_size + g->getY(m.spot) * g->x_size + g->getX(m.spot);
}
// GraphHash::getGraphHash (cpp/game/graphhash.cpp:3-28), literal: state hash = getSitHash(nextPla) ^ GAME_IS_OVER when finished;
// after any move the previous graph hash is chained in (hash0 = splitMix64(h0 ^ h1), hash1 = nasam(h1) + hash0, then += state hash).
static uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }
static uint64_t nasam(uint64_t x) {   // cpp/core/hash.cpp:72-80
  x ^= rotr64(x, 25) ^ rotr64(x, 47);
  x *= 0x9e6c63d0676a9a99ULL;
  x ^= (x >> 23) ^ (x >> 51);
  x *= 0x9e6d62d06f6a9a9bULL;
  x ^= (x >> 23) ^ (x >> 51);
  return x;
}
void ko_graph_hash(const uint64_t prev[2], const ko_game* g, int nextPla, uint64_t out[2]) {
  uint64_t st[2];
  ko_game_sit_hash(g, nextPla, st);
  if(g->isGameFinished) { st[0] ^= GAME_IS_OVER0; st[1] ^= GAME_IS_OVER1; }
  if(g->moveHistory.empty()) { out[0] = st[0]; out[1] = st[1]; return; }
  uint64_t h0 = ko_splitmix64(prev[0] ^ prev[1]);
  uint64_t h1 = nasam(prev[1]) + h0;
  out[0] = h0 + st[0];
  out[1] = h1 + st[1];
}
uint32_t ko_game_status(const ko_game* g) {
  return (uint32_t)(g->numTurns & 0xff) | ((g->isGameFinished ? 1u : 0u) << 8) |
         ((uint32_t)g->winner << 9) | ((uint32_t)g->presumedNextMovePla << 11);
}
void ko_game_sit_hash(const ko_game* g, int pla, uint64_t out[2]) {
  out[0] = g->pos_hash[0] ^ zob().player[pla][0];
  out[1] = g->pos_hash[1] ^ zob().player[pla][1];
}
void ko_game_nn_hash(const ko_game* g, int pla, double pda, float temp, double optimism, uint64_t out[2]) {
  uint64_t h0, h1;
  {
    uint64_t t[2];
    ko_game_sit_hash(g, pla, t);
    h0 = t[0]; h1 = t[1];
  }
  if(g->isGameFinished) { h0 ^= GAME_IS_OVER0; h1 ^= GAME_IS_OVER1; }
  if(pda != 0) {
    int64_t d = (int64_t)(pda * 256.0f);
    h0 += ko_splitmix64((uint64_t)d);
    h1 += ko_basic_lcong((uint64_t)d);
    h0 ^= ZPD0; h1 ^= ZPD1;
  }
  if(temp != 1.0f) {
    int64_t d = (int64_t)(temp * 2048.0f);
    h0 ^= ko_basic_lcong2((uint64_t)d);
    h1 = ko_splitmix64(h1 + (uint64_t)d);
    h0 += h1;
    h0 ^= ZPT0; h1 ^= ZPT1;
  }
  if(optimism > 0) {
    h0 ^= ZPO0; h1 ^= ZPO1;
    int64_t d = (int64_t)(optimism * 1024.0);
    h0 = ko_rrmxmx(ko_splitmix64(h0) + (uint64_t)d);
    h1 = ko_rrmxmx(h1 + h0 + (uint64_t)d);
  }
  out[0] = h0; out[1] = h1;
}

void ko_game_fill_row_v1(const ko_game* g, int pla, int nnXLen, int nnYLen, int useNHWC,
                         float* rowBin, float* rowGlobal) {
  const int C = 15;  // ledger F
  std::fill(rowBin, rowBin + C * nnXLen * nnYLen, 0.0f);
  rowGlobal[0] = 0.0f;
  int opp = pla ^ 3;
  int featureStride = useNHWC ? 1 : nnXLen * nnYLen;
  int posStride = useNHWC ? C : 1;
  auto set = [&](int pos, int feature) { rowBin[pos * posStride + feature * featureStride] = 1.0f; };
  auto spotToPos = [&](int spot) { return g->getY(spot) * nnXLen + g->getX(spot); };
  // features 0,1,2 (nninputs.cpp:541-559)
  for(int y = 0; y < g->y_size; y++)
    for(int x = 0; x < g->x_size; x++) {
      int pos = y * nnXLen + x;
      set(pos, 0);
      int stone = g->colors[g->spotOf(x, y)];
      if(stone == pla) set(pos, 1);
      else if(stone == opp) set(pos, 2);
    }
  // features 3..6: last move, channel = its direction (nninputs.cpp:562-568)
  size_t len = g->moveHistory.size();
  if(len > 0) {
    const Move& m = g->moveHistory[len - 1];
    set(spotToPos(m.spot), 3 + m.dir);
  }
  // features 7..10: moves 2,3,4,5 plies ago, alternation chain (nninputs.cpp:577-632), fixed channels
  int numTurns = g->numTurns;
  if(numTurns >= 2 && len >= 2 && g->moveHistory[len - 2].pla == pla) {
    set(spotToPos(g->moveHistory[len - 2].spot), 7);
    if(numTurns >= 3 && len >= 3 && g->moveHistory[len - 3].pla == opp) {
      set(spotToPos(g->moveHistory[len - 3].spot), 8);
      if(numTurns >= 4 && len >= 4 && g->moveHistory[len - 4].pla == pla) {
        set(spotToPos(g->moveHistory[len - 4].spot), 9);
        if(numTurns >= 5 && len >= 5 && g->moveHistory[len - 5].pla == opp)
          set(spotToPos(g->moveHistory[len - 5].spot), 10);
      }
    }
  }
  // feature 11: legal-spot mask, OR over the 4 directions (nninputs.cpp:635-647, ledger F)
  for(int x = 0; x < g->x_size; x++)
    for(int y = 0; y < g->y_size; y++) {
      int spot = g->spotOf(x, y);
      for(int d = 0; d < 4; d++)
        if(g->isLegal(spot, d, pla)) { set(y * nnXLen + x, 11); break; }
    }
  // features 12..14: runs of exactly k-1, k-2, k-3 (nninputs.cpp:650-653)
  for(int len2 = g->win_len - 1; len2 >= g->win_len - 3; --len2) {
    int feature = 12 + g->win_len - 1 - len2;
    fillLinePlane(g, len2, rowBin + feature * featureStride, nnXLen, posStride);
  }
  rowGlobal[0] = (float)g->win_len;  // nninputs.cpp:656
}

void ko_copy_inputs_with_symmetry(const float* src, float* dst, int n, int h, int w, int c,
                                  int useNHWC, int symmetry) {
  copyWithSymmetry(src, dst, n, h, w, c, useNHWC != 0, symmetry, false);
}
void ko_copy_outputs_with_symmetry(const float* src, float* dst, int n, int h, int w, int symmetry) {
  copyWithSymmetry(src, dst, n, h, w, 1, false, symmetry, true);
}
// nninputs.cpp:359-375
int ko_sym_invert(int s) { return s == 5 ? 6 : (s == 6 ? 5 : s); }
int ko_sym_compose(int first, int next) {
  if(first & 4) next = (next & 4) | ((next & 2) >> 1) | ((next & 1) << 1);
  return first ^ next;
}
// nninputs.cpp:409-433 with ledger J
int ko_sym_dir(int dir, int symmetry) {
  if(dir == D_NONE) return D_NONE;
  bool tr = (symmetry & 4) != 0, fx = (symmetry & 2) != 0, fy = (symmetry & 1) != 0;
  if(fx ^ fy) {
    if(dir == 3) dir = 2;
    else if(dir == 2) dir = 3;
  }
  if(tr) {
    if(dir == 0) dir = 1;
    else if(dir == 1) dir = 0;
  }
  return dir;
}
// nninputs.cpp:377-391
void ko_sym_xy(int x, int y, int xSize, int ySize, int symmetry, int* outX, int* outY) {
  bool tr = (symmetry & 4) != 0, fx = (symmetry & 2) != 0, fy = (symmetry & 1) != 0;
  if(fx) x = xSize - x - 1;
  if(fy) y = ySize - y - 1;
  if(tr) std::swap(x, y);
  *outX = x; *outY = y;
}

int ko_playout_choose(const ko_game* g, uint64_t seed, uint64_t gameIdx, uint64_t* rOut) {
  uint32_t mask[13];
  int n = ko_game_legal_mask(g, g->presumedNextMovePla, mask);
  uint64_t r = ko_splitmix64(seed ^ (gameIdx * 0x9E3779B97F4A7C15ULL) ^ (uint64_t)g->numTurns);
  if(rOut) *rOut = r;
  if(n == 0) return -1;
  int k = (int)(r % (uint64_t)n);
  int hw4 = 4 * g->x_size * g->y_size;
  for(int pos = 0; pos < hw4; pos++)
    if(mask[pos >> 5] >> (pos & 31) & 1) {
      if(k == 0) return pos;
      k--;
    }
  return -1;
}

long ko_playout_run(int x_size, int y_size, int win_len, uint64_t seed, uint64_t g0, int n,
                    int maxPlies, ko_step_record* records, long maxRecords, float* planes,
                    int useNHWC, float* globals, int threads) {
  if(threads < 1) threads = 1;
  const int hw = x_size * y_size;
  const int planeElts = 15 * hw;
  // pass 1 (serial, cheap): count records per game so that threads can write disjoint ranges
  // deterministically in game order.
  std::vector<long> offset(n + 1, 0);
  {
    std::vector<int> cnt(n, 0);
    auto countRange = [&](int lo, int hi) {
      ko_game* g = ko_game_create(x_size, y_size, win_len);
      for(int i = lo; i < hi; i++) {
        g->init();
        int c = 1;
        while(!g->isGameFinished && g->numTurns < maxPlies) {
          int pos = ko_playout_choose(g, seed, g0 + (uint64_t)i, nullptr);
          if(pos < 0) break;
          ko_game_play(g, pos);
          c++;
        }
        cnt[i] = c;
      }
      ko_game_destroy(g);
    };
    std::vector<std::thread> th;
    for(int t = 0; t < threads; t++) {
      int lo = (int)((long)n * t / threads), hi = (int)((long)n * (t + 1) / threads);
      th.emplace_back(countRange, lo, hi);
    }
    for(auto& t : th) t.join();
    for(int i = 0; i < n; i++) offset[i + 1] = offset[i] + cnt[i];
  }
  long total = offset[n];
  auto runRange = [&](int lo, int hi) {
    ko_game* g = ko_game_create(x_size, y_size, win_len);
    for(int i = lo; i < hi; i++) {
      g->init();
      long idx = offset[i];
      int movePos = -1;
      while(true) {
        if(idx < maxRecords) {
          int pla = g->presumedNextMovePla;
          if(records) {
            ko_step_record& r = records[idx];
            r.game = (uint32_t)(g0 + i);
            r.status = ko_game_status(g);
            for(int w = 0; w < 13; w++) r.legal[w] = 0;
            ko_game_legal_mask(g, pla, r.legal);  // raw isLegal mask, also on finished positions
            r.movePos = movePos;
            ko_game_sit_hash(g, pla, r.sitHash);
            ko_game_nn_hash(g, pla, 0.0, 1.0f, 0.0, r.nnHash);
          }
          if(planes) {
            float gl;
            ko_game_fill_row_v1(g, pla, x_size, y_size, useNHWC, planes + idx * planeElts, &gl);
            if(globals) globals[idx] = gl;
          }
        }
        idx++;
        if(g->isGameFinished || g->numTurns >= maxPlies) break;
        movePos = ko_playout_choose(g, seed, g0 + (uint64_t)i, nullptr);
        if(movePos < 0) break;
        ko_game_play(g, movePos);
      }
    }
    ko_game_destroy(g);
  };
  std::vector<std::thread> th;
  for(int t = 0; t < threads; t++) {
    int lo = (int)((long)n * t / threads), hi = (int)((long)n * (t + 1) / threads);
    th.emplace_back(runRange, lo, hi);
  }
  for(auto& t : th) t.join();
  return total < maxRecords ? total : maxRecords;
}

}  // extern "C"
