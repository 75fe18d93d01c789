// ORACLE (test infrastructure only).  C entry points over the REFERENCE'S OWN rules / hashing / NN-input code, compiled by
// oracle/ref_patch.sh from a patched scratch copy of /root/reference/cpp (game/board, game/boardhistory, neuralnet/nninputs,
// core/hash, core/rand ...) into oracle/_ref/libkc_ref_rules.so.  This file is ours: it contains no reference code, it only CALLS
//   Board::initHash / isLegal / playMoveAssumeLegal / maxConsecutives / checkGameEnd / getSitHash   (cpp/game/board.cpp)
//   BoardHistory::makeBoardMove / makeBoardMoveAssumeLegal                                          (cpp/game/boardhistory.cpp)
//   NNInputs::getHash / fillRowV1, NNPos::*, SymmetryHelpers::*                                     (cpp/neuralnet/nninputs.cpp)
//   GraphHash::getGraphHash / getGraphHashFromScratch                                                (cpp/game/graphhash.cpp)
//   Rand                                                                                             (cpp/core/rand.cpp)
// so that tests/test_oracle_ref_rules.py can hold the restatement in oracle/ko_game.cpp / ko_hash.cpp against the literal code.
//
// What the DRIVER below adds on top of the literal code, because the literal code has no such thing:
//   ledger C  a draw: after a non-winning move, a player to move without a legal Loc ends the game with winner C_EMPTY
//             (the literal BoardHistory never ends a game without a winner);
//   ledger D  moves are made with makeBoardMove (the one variant that appends moveHistory), so that fillRowV1's history planes see them;
//   ledger B  isLegal is only asked for dir 0..3 on on-board spots (D_NONE indexes ADJS out of bounds).
// The synthetic playout (move = (r mod popcount)-th legal Loc in policy order, r = Hash::splitMix64(seed ^ g * phi ^ ply)) is
// SURVEY.md 8(d)'s; it uses the reference's own Hash::splitMix64.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "core/hash.h"
#include "core/rand.h"
#include "game/board.h"
#include "game/boardhistory.h"
#include "game/graphhash.h"
#include "neuralnet/nninputs.h"

namespace {
struct StepRecord {   // == ko_step_record (oracle/kc_oracle.h)
  uint32_t game;
  uint32_t status;
  uint32_t legal[KC_REF_LEGAL_WORDS];
  int32_t movePos;
  uint64_t sitHash[2];
  uint64_t nnHash[2];
};

// raw Board::isLegal mask of `pla` in policy order dir*HW + y*W + x; returns the number of legal Locs
int legalMask(const Board& b, Player pla, uint32_t* mask, int words) {
  for(int w = 0; w < words; w++) mask[w] = 0;
  const int hw = b.x_size * b.y_size;
  int n = 0;
  for(int d = 0; d < NUM_ACTUAL_DIRECTIONS; d++)
    for(int y = 0; y < b.y_size; y++)
      for(int x = 0; x < b.x_size; x++)
        if(b.isLegal(Loc(Location::getSpot(x, y, b.x_size), (Direction)d), pla)) {
          const int pos = d * hw + y * b.x_size + x;
          mask[pos >> 5] |= 1u << (pos & 31);
          n++;
        }
  return n;
}
uint32_t statusOf(const BoardHistory& h) {
  return (uint32_t)(h.numTurns & 255) | ((h.isGameFinished ? 1u : 0u) << 8) | ((uint32_t)(h.winner & 3) << 9) | ((uint32_t)(h.presumedNextMovePla & 3) << 11);
}
}  // namespace

extern "C" {

void kc_ref_init() { Board::initHash(); }

// the Zobrist tables Board::initHash filled (board.cpp:134-178): board [MAX_ARR_SIZE][4][2], player [4][2], sizeX / sizeY [MAX_LEN+1][2]
int kc_ref_tables(uint64_t* board, uint64_t* player, uint64_t* sizeX, uint64_t* sizeY) {
  Board::initHash();
  for(int i = 0; i < Board::MAX_ARR_SIZE; i++)
    for(int c = 0; c < 4; c++) { board[(i * 4 + c) * 2] = Board::ZOBRIST_BOARD_HASH[i][c].hash0; board[(i * 4 + c) * 2 + 1] = Board::ZOBRIST_BOARD_HASH[i][c].hash1; }
  for(int c = 0; c < 4; c++) { player[c * 2] = Board::ZOBRIST_PLAYER_HASH[c].hash0; player[c * 2 + 1] = Board::ZOBRIST_PLAYER_HASH[c].hash1; }
  for(int i = 0; i <= Board::MAX_LEN; i++) {
    sizeX[i * 2] = Board::ZOBRIST_SIZE_X_HASH[i].hash0; sizeX[i * 2 + 1] = Board::ZOBRIST_SIZE_X_HASH[i].hash1;
    sizeY[i * 2] = Board::ZOBRIST_SIZE_Y_HASH[i].hash0; sizeY[i * 2 + 1] = Board::ZOBRIST_SIZE_Y_HASH[i].hash1;
  }
  return Board::MAX_ARR_SIZE;
}

// Rand(seed).nextUInt() x n and .nextUInt64() x n (rand.cpp:276-318, rand.h inline)
void kc_ref_rand(const char* seed, int n, uint32_t* out32, uint64_t* out64) {
  Rand r{std::string(seed)};
  for(int i = 0; i < n; i++) out32[i] = r.nextUInt();
  Rand q{std::string(seed)};
  for(int i = 0; i < n; i++) out64[i] = q.nextUInt64();
}

// The playouts of SURVEY.md 8(d) on the reference's Board / BoardHistory.  One record per position (initial position first), in game
// order, as oracle/ko_playout_run writes them.  planes (may be null): [records][16*H*W] = the literal fillRowV1 row (NCHW:
// channel stride H*W; NHWC: pos stride NUM_FEATURES_SPATIAL_V1 = 16), of which channels 0..10 are meaningful (ledger F, G).
long kc_ref_playout_run(int xSize, int ySize, int winLen, uint64_t seed, uint64_t g0, int n, int maxPlies, StepRecord* records,
                        long maxRecords, float* planes, int useNHWC, float* globals) {
  Board::initHash();
  const int hw = xSize * ySize;
  const int words = (4 * hw + 31) / 32;
  if(words > KC_REF_LEGAL_WORDS) return -1;
  long idx = 0;
  MiscNNInputParams params;
  std::vector<float> row((size_t)(NNInputs::NUM_FEATURES_SPATIAL_V1 + 8) * hw + 64);
  for(int i = 0; i < n; i++) {
    Board board(xSize, ySize, winLen);
    BoardHistory hist(board, P_BLACK);
    int movePos = -1;
    while(true) {
      const Player pla = hist.presumedNextMovePla;
      uint32_t mask[KC_REF_LEGAL_WORDS];
      const int numLegal = legalMask(board, pla, mask, KC_REF_LEGAL_WORDS);
      if(!hist.isGameFinished && numLegal == 0) { hist.isGameFinished = true; hist.winner = C_EMPTY; }   // ledger C (driver)
      if(idx < maxRecords) {
        if(records) {
          StepRecord& r = records[idx];
          r.game = (uint32_t)(g0 + i);
          r.status = statusOf(hist);
          memcpy(r.legal, mask, sizeof(mask));
          r.movePos = movePos;
          const Hash128 sh = board.getSitHash(pla);
          r.sitHash[0] = sh.hash0; r.sitHash[1] = sh.hash1;
          const Hash128 nh = NNInputs::getHash(board, hist, pla, params);
          r.nnHash[0] = nh.hash0; r.nnHash[1] = nh.hash1;
        }
        if(planes) {
          float gl = 0;
          std::fill(row.begin(), row.end(), 0.0f);
          NNInputs::fillRowV1(board, hist, pla, params, xSize, ySize, useNHWC != 0, row.data(), &gl);
          memcpy(planes + (size_t)idx * NNInputs::NUM_FEATURES_SPATIAL_V1 * hw, row.data(), sizeof(float) * NNInputs::NUM_FEATURES_SPATIAL_V1 * hw);
          if(globals) globals[idx] = gl;
        }
      }
      idx++;
      if(hist.isGameFinished || hist.numTurns >= maxPlies) break;
      const uint64_t r = Hash::splitMix64(seed ^ ((g0 + (uint64_t)i) * 0x9E3779B97F4A7C15ULL) ^ (uint64_t)hist.numTurns);
      int k = (int)(r % (uint64_t)numLegal);
      movePos = -1;
      for(int pos = 0; pos < 4 * hw; pos++)
        if(mask[pos >> 5] >> (pos & 31) & 1) {
          if(k == 0) { movePos = pos; break; }
          k--;
        }
      const int d = movePos / hw, rem = movePos % hw;
      hist.makeBoardMove(board, Loc(Location::getSpot(rem % xSize, rem / xSize, xSize), (Direction)d), pla);   // ledger D
    }
  }
  return idx < maxRecords ? idx : maxRecords;
}

// Arbitrary (not necessarily reachable) position: stones [H*W] 0/1/2, last move (lastX < 0: none).  Outputs: the raw isLegal mask of
// `pla`, Board::maxConsecutives per stone cell (0 on empty cells), getSitHash(pla), pos_hash.  Returns the number of legal Locs.
int kc_ref_position(int xSize, int ySize, int winLen, const int8_t* stones, int lastX, int lastY, int lastDir, int pla, uint32_t* legal,
                    int32_t* maxConsec, uint64_t* sitHash, uint64_t* posHash) {
  Board::initHash();
  Board b(xSize, ySize, winLen);
  for(int y = 0; y < ySize; y++)
    for(int x = 0; x < xSize; x++) {
      const int c = stones[y * xSize + x];
      if(c == C_BLACK || c == C_WHITE) {   // a stone placed by a move: playMoveAssumeLegal keeps pos_hash current (board.cpp:427-435)
        b.playMoveAssumeLegal(Loc(Location::getSpot(x, y, xSize), D_NORTH), (Player)c);
      }
    }
  b.lastLoc = lastX < 0 ? Loc(Board::NULL_LOC, D_NONE) : Loc(Location::getSpot(lastX, lastY, xSize), (Direction)lastDir);
  const int words = (4 * xSize * ySize + 31) / 32;
  const int n = legalMask(b, (Player)pla, legal, words);
  for(int y = 0; y < ySize; y++)
    for(int x = 0; x < xSize; x++) {
      const Spot s = Location::getSpot(x, y, xSize);
      maxConsec[y * xSize + x] = b.colors[s] == C_EMPTY ? 0 : b.maxConsecutives(s);
    }
  const Hash128 sh = b.getSitHash((Player)pla);
  sitHash[0] = sh.hash0; sitHash[1] = sh.hash1;
  posHash[0] = b.pos_hash.hash0; posHash[1] = b.pos_hash.hash1;
  return n;
}

// GraphHash::getGraphHash (graphhash.cpp:14-29) chained along a move sequence from the empty board: out[i] = the hash after i moves
// (out[0] = the hash of the initial position), and GraphHash::getGraphHashFromScratch of the final history in out[numMoves + 1]
void kc_ref_graph_hash_chain(int xSize, int ySize, int winLen, int numMoves, const int32_t* movePos, uint64_t* out) {
  Board::initHash();
  Board board(xSize, ySize, winLen);
  BoardHistory hist(board, P_BLACK);
  const int hw = xSize * ySize;
  Hash128 h = GraphHash::getGraphHash(Hash128(), hist, hist.presumedNextMovePla);
  out[0] = h.hash0; out[1] = h.hash1;
  for(int i = 0; i < numMoves; i++) {
    const int d = movePos[i] / hw, rem = movePos[i] % hw;
    hist.makeBoardMove(board, Loc(Location::getSpot(rem % xSize, rem / xSize, xSize), (Direction)d), hist.presumedNextMovePla);
    h = GraphHash::getGraphHash(h, hist, hist.presumedNextMovePla);
    out[2 * (i + 1)] = h.hash0; out[2 * (i + 1) + 1] = h.hash1;
  }
  const Hash128 s = GraphHash::getGraphHashFromScratch(hist, hist.presumedNextMovePla);
  out[2 * (numMoves + 1)] = s.hash0; out[2 * (numMoves + 1) + 1] = s.hash1;
}

// NNInputs::getHash on an empty board of the given size with every MiscNNInputParams fold (nninputs.cpp:463-502)
void kc_ref_nn_hash_params(int xSize, int ySize, int winLen, int pla, int finished, double pda, float temperature, double optimism, uint64_t* out) {
  Board::initHash();
  Board b(xSize, ySize, winLen);
  BoardHistory h(b, (Player)pla);
  h.isGameFinished = finished != 0;
  MiscNNInputParams p;
  p.playoutDoublingAdvantage = pda; p.nnPolicyTemperature = temperature; p.policyOptimism = optimism;
  const Hash128 r = NNInputs::getHash(b, h, (Player)pla, p);
  out[0] = r.hash0; out[1] = r.hash1;
}

void kc_ref_copy_inputs_with_symmetry(const float* src, float* dst, int n, int h, int w, int c, int useNHWC, int symmetry) {
  SymmetryHelpers::copyInputsWithSymmetry(src, dst, n, h, w, c, useNHWC != 0, symmetry);
}
void kc_ref_copy_outputs_with_symmetry(const float* src, float* dst, int n, int h, int w, int symmetry) {
  SymmetryHelpers::copyOutputsWithSymmetry(src, dst, n, h, w, symmetry);
}
int kc_ref_sym_invert(int s) { return SymmetryHelpers::invert(s); }
int kc_ref_sym_compose(int a, int b) { return SymmetryHelpers::compose(a, b); }
int kc_ref_sym_dir(int dir, int symmetry) { return SymmetryHelpers::getSymDir((Direction)dir, symmetry); }   // with ledger J's return
void kc_ref_sym_xy(int x, int y, int xSize, int ySize, int symmetry, int* ox, int* oy) {
  const Spot s = SymmetryHelpers::getSymSpot(x, y, xSize, ySize, symmetry);
  const int xs = (symmetry & 4) ? ySize : xSize;
  *ox = Location::getX(s, xs); *oy = Location::getY(s, xs);
}
// NNPos (nninputs.cpp:6-49).  posToLoc is the literal function incl. its `pos /= HW` (ledger I): outputs x, y, dir of the Loc it returns
int kc_ref_loc_to_pos(int x, int y, int dir, int xSize, int nnXLen, int nnYLen) {
  return NNPos::locToPos(Loc(Location::getSpot(x, y, xSize), (Direction)dir), xSize, nnXLen, nnYLen);
}
void kc_ref_pos_to_loc(int pos, int xSize, int ySize, int nnXLen, int nnYLen, int* x, int* y, int* dir) {
  const Loc l = NNPos::posToLoc(pos, xSize, ySize, nnXLen, nnYLen);
  *dir = l.dir;
  if(l.spot == Board::NULL_LOC) { *x = -1; *y = -1; return; }
  *x = Location::getX(l.spot, xSize); *y = Location::getY(l.spot, xSize);
}
int kc_ref_policy_size(int nnXLen, int nnYLen) { return NNPos::getPolicySize(nnXLen, nnYLen); }

}  // extern "C"
