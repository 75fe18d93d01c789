// Custom Coffee SGF: README.md:33-35 -- a move is B[xyd] / W[xyd]: column letter, row letter and a third
// letter a..d for the line direction | - \ / (= D_NORTH, D_WEST, D_NORTHWEST, D_NORTHEAST = 0..3, cpp/game/board.h:41-48);
// AB / AW placements give the starting position.  Restates the Coffee-adapted parts of cpp/dataio/sgf.cpp: coordinate and
// direction letters (:42-100, :132-154: a..z = 0..25, A..Z = 26..51), Sgf main-line traversal (properties SZ, AB, AW, B, W,
// RE) and WriteSgf::writeSgf's header (:1526-1548: "(;FF[4]GM[Coffee]SZ[n]" or SZ[x:y], the win length as WLL[k], PB, PW, RE).
// Canonical choice (ledger C): a draw is written RE[0]; the reference's gameResultNoSgfTag asserts on a game without winner
// (:1491-1501).  Host-only, no GPU.
#include <cstring>

#include "kc_internal.h"

namespace {

struct SgfError { std::string msg; };

int coordOf(char c) {
  if(c >= 'a' && c <= 'z') return c - 'a';
  if(c >= 'A' && c <= 'Z') return c - 'A' + 26;
  return -1;
}
char letterOf(int v) { return "abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"[v]; }
int dirOf(char c) {
  if(c >= 'a' && c <= 'd') return c - 'a';
  if(c >= 'A' && c <= 'D') return c - 'A';
  return -1;
}

}  // namespace

extern "C" {

int kc_sgf_write(int xSize, int ySize, int winLen, const char* blackName, const char* whiteName, const int8_t* initialStones, int numMoves,
                 const int16_t* movePos, const int8_t* movePla, int winner, char* out, size_t outCap, size_t* outLen) {
  KC_CHECK(xSize >= 1 && ySize >= 1 && xSize <= KC_MAX_LEN && ySize <= KC_MAX_LEN, "kc_sgf_write: board size out of range");
  KC_CHECK(numMoves >= 0 && (numMoves == 0 || (movePos && movePla)), "kc_sgf_write: bad move list");
  KC_CHECK(winner >= -1 && winner <= 2, "kc_sgf_write: winner must be -1 (unfinished), 0 (draw), 1 (black) or 2 (white)");
  const int HW = xSize * ySize;
  std::string s = "(;FF[4]GM[Coffee]";
  s += xSize == ySize ? "SZ[" + std::to_string(xSize) + "]" : "SZ[" + std::to_string(xSize) + ":" + std::to_string(ySize) + "]";
  s += "WLL[" + std::to_string(winLen) + "]";
  s += std::string("PB[") + (blackName ? blackName : "") + "]PW[" + (whiteName ? whiteName : "") + "]";
  if(winner >= 0) s += winner == 1 ? "RE[B+]" : winner == 2 ? "RE[W+]" : "RE[0]";
  for(int col = 1; col <= 2 && initialStones; col++) {
    bool any = false;
    for(int y = 0; y < ySize; y++)
      for(int x = 0; x < xSize; x++)
        if(initialStones[y * xSize + x] == col) {
          if(!any) { s += col == 1 ? "AB" : "AW"; any = true; }
          s += '['; s += letterOf(x); s += letterOf(y); s += ']';
        }
  }
  for(int i = 0; i < numMoves; i++) {
    KC_CHECK(movePla[i] == 1 || movePla[i] == 2, "kc_sgf_write: move player must be 1 or 2");
    KC_CHECK(movePos[i] >= -1 && movePos[i] < 4 * HW, "kc_sgf_write: move out of range");
    s += movePla[i] == 1 ? ";B[" : ";W[";
    if(movePos[i] >= 0) {
      const int dir = movePos[i] / HW, cell = movePos[i] % HW;
      s += letterOf(cell % xSize); s += letterOf(cell / xSize); s += "abcd"[dir];
    }
    s += ']';
  }
  s += ")\n";
  if(outLen) *outLen = s.size();
  KC_CHECK(out && outCap > s.size(), "kc_sgf_write: output buffer too small (need " + std::to_string(s.size() + 1) + " bytes)");
  memcpy(out, s.c_str(), s.size() + 1);
  return 0;
}

int kc_sgf_parse(const char* sgf, int* xSizeOut, int* ySizeOut, int* winLenOut, int8_t* initialStones, int maxMoves, int16_t* movePos,
                 int8_t* movePla, int* numMovesOut, int* winnerOut) {
  KC_CHECK(sgf && xSizeOut && ySizeOut && numMovesOut, "kc_sgf_parse: null argument");
  const std::string s(sgf);
  try {
    size_t pos = 0;
    auto skipWs = [&]() { while(pos < s.size() && isspace((unsigned char)s[pos])) pos++; };
    skipWs();
    if(pos >= s.size() || s[pos] != '(') throw SgfError{"expected '(' at the start of the game tree"};
    pos++;
    int xSize = 0, ySize = 0, winLen = 4, winner = -1, numMoves = 0, depth = 1;
    std::vector<std::pair<int, std::pair<int, int>>> placements;   // colour, (x, y): applied once the size is known
    std::vector<std::pair<std::string, int>> moves;                // value, player
    bool sawNode = false;
    while(true) {
      skipWs();
      if(pos >= s.size()) throw SgfError{"unexpected end of SGF"};
      const char c = s[pos];
      if(c == ';') { pos++; sawNode = true; continue; }
      if(c == '(') {   // a variation: follow the first one (main line), skip the others
        if(depth >= 1 && sawNode) { depth++; pos++; continue; }
        throw SgfError{"unexpected '('"};
      }
      if(c == ')') {
        // end of the main line: everything after the first closing parenthesis belongs to other variations
        break;
      }
      if(!isupper((unsigned char)c)) throw SgfError{std::string("unexpected character '") + c + "'"};
      std::string key;
      while(pos < s.size() && (isupper((unsigned char)s[pos]) || islower((unsigned char)s[pos]))) { if(isupper((unsigned char)s[pos])) key += s[pos]; pos++; }
      std::vector<std::string> values;
      while(true) {
        skipWs();
        if(pos >= s.size() || s[pos] != '[') break;
        pos++;
        std::string v;
        while(true) {
          if(pos >= s.size()) throw SgfError{"unterminated property value of " + key};
          if(s[pos] == '\\' && pos + 1 < s.size()) { v += s[pos + 1]; pos += 2; continue; }
          if(s[pos] == ']') { pos++; break; }
          v += s[pos++];
        }
        values.push_back(v);
      }
      if(values.empty()) throw SgfError{"property " + key + " has no value"};
      if(key == "SZ") {
        const std::string& v = values[0];
        const size_t colon = v.find(':');
        xSize = atoi(v.substr(0, colon).c_str());
        ySize = colon == std::string::npos ? xSize : atoi(v.substr(colon + 1).c_str());
        if(xSize < 1 || ySize < 1 || xSize > KC_MAX_LEN || ySize > KC_MAX_LEN) throw SgfError{"Invalid board size: " + v};
      } else if(key == "WLL") {
        winLen = atoi(values[0].c_str());
        if(winLen < 1) throw SgfError{"Invalid win length: " + values[0]};
      } else if(key == "RE") {
        const std::string& v = values[0];
        winner = (v.size() >= 2 && (v[0] == 'B' || v[0] == 'b') && v[1] == '+') ? 1 : (v.size() >= 2 && (v[0] == 'W' || v[0] == 'w') && v[1] == '+') ? 2
                 : (v == "0" || v == "Draw" || v == "draw") ? 0 : -1;
      } else if(key == "AB" || key == "AW" || key == "AE") {
        const int col = key == "AB" ? 1 : key == "AW" ? 2 : 0;
        for(const std::string& v : values) {
          // a point "xy" or a rectangle "xy:xy" (sgf.cpp:102-126)
          int x1, y1, x2, y2;
          if(v.size() == 5 && v[2] == ':') { x1 = coordOf(v[0]); y1 = coordOf(v[1]); x2 = coordOf(v[3]); y2 = coordOf(v[4]); }
          else if(v.size() == 2) { x1 = x2 = coordOf(v[0]); y1 = y2 = coordOf(v[1]); }
          else throw SgfError{"Invalid location: " + v};
          if(x1 < 0 || y1 < 0 || x2 < x1 || y2 < y1) throw SgfError{"Invalid location or location rect: " + v};
          for(int x = x1; x <= x2; x++) for(int y = y1; y <= y2; y++) placements.push_back({col, {x, y}});
        }
      } else if(key == "B" || key == "W") {
        if(values.size() != 1) throw SgfError{"SGF property is not a singleton: " + key};
        moves.push_back({values[0], key == "B" ? 1 : 2});
      }
      // every other property (FF, GM, PB, PW, C, ...) is ignored
    }
    if(xSize == 0) throw SgfError{"SGF does not contain property: SZ"};
    const int HW = xSize * ySize;
    if(initialStones) memset(initialStones, 0, (size_t)HW);
    for(const auto& p : placements) {
      const int x = p.second.first, y = p.second.second;
      if(x >= xSize || y >= ySize) throw SgfError{"Invalid location: placement outside the board"};
      if(initialStones) initialStones[y * xSize + x] = (int8_t)p.first;
    }
    for(const auto& mv : moves) {
      const std::string& v = mv.first;
      int p = -1;
      if(!v.empty()) {   // empty value = no move (parseSgfLoc, sgf.cpp:125-130)
        if(v.size() != 3) throw SgfError{"Invalid location: " + v};
        const int x = coordOf(v[0]), y = coordOf(v[1]), d = dirOf(v[2]);
        if(x < 0 || y < 0 || x >= xSize || y >= ySize || d < 0) throw SgfError{"Invalid location: " + v};
        p = d * HW + y * xSize + x;
      }
      if(numMoves < maxMoves) {
        if(movePos) movePos[numMoves] = (int16_t)p;
        if(movePla) movePla[numMoves] = (int8_t)mv.second;
      }
      numMoves++;
    }
    *xSizeOut = xSize; *ySizeOut = ySize;
    if(winLenOut) *winLenOut = winLen;
    if(winnerOut) *winnerOut = winner;
    *numMovesOut = numMoves;   // may exceed maxMoves: the caller sees how much room the game needs
  } catch(const SgfError& e) {
    return kc::fail(std::string("kc_sgf_parse: ") + e.msg);
  }
  return 0;
}

}  // extern "C"
