#!/bin/bash
# ORACLE (test infrastructure only).  Builds oracle/_ref/libkc_ref_rules.so: the reference's OWN rules / hashing / NN-input code --
# game/board.{h,cpp}, game/boardhistory.{h,cpp}, neuralnet/nninputs.{h,cpp}, core/{hash,rand,rand_helpers,md5,sha2,global}.* --
# compiled from a SCRATCH COPY under oracle/_ref/build/ after the minimal edits below.  The reference tree does not compile as it
# stands; every edit is one line (or one cut) and is keyed to the SURVEY.md 0.2 defect or 8.1 ledger row that
# makes it necessary.  Nothing is copied into the repository: the scratch copy and the .so live under the git-ignored oracle/_ref/.
# The result is what tests/test_oracle_ref_rules.py compares the oracle restatement (oracle/ko_game.cpp, ko_hash.cpp) against.
#
# Usage: oracle/ref_patch.sh [REF=/root/reference]     (called by `make -C oracle refrules`)
set -euo pipefail
REF="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
B="$HERE/_ref/build"
[ -f "$REF/cpp/game/board.cpp" ] || { echo "reference tree not present at $REF: keeping prebuilt oracle/_ref (if any)"; exit 0; }
rm -rf "$B"
mkdir -p "$B/core" "$B/game" "$B/neuralnet"
for f in core/global.h core/global.cpp core/hash.h core/hash.cpp core/rand.h core/rand.cpp core/rand_helpers.h core/rand_helpers.cpp \
         core/md5.h core/md5.cpp core/sha2.h core/sha2.cpp core/os.h core/test.h core/test.cpp core/timer.h core/timer.cpp core/bsearch.h \
         core/bsearch.cpp core/using.h core/commontypes.h game/board.h game/board.cpp game/boardhistory.h game/boardhistory.cpp game/graphhash.h game/graphhash.cpp \
         neuralnet/nninputs.h neuralnet/nninputs.cpp; do
  cp "$REF/cpp/$f" "$B/$f"
done
ln -s "$REF/cpp/external" "$B/external"     # board.h includes ../external/nlohmann_json/json.hpp; read in place, not copied

# ---- core/global.h:342-349 declares `string` globals without std:: and defines non-inline variables in a header ----
sed -i -E '/^namespace ColoredOutput \{/,/^\}/ s/^  string (BACKGROUND|WORD|END|RESET) =/  static const std::string \1 =/' "$B/core/global.h"
sed -i -E '/^namespace ColoredOutput \{/,/^\}/ s/^  string colorize\(string text/  std::string colorize(std::string text/' "$B/core/global.h"

# ---- game/board.h:24-26 uses Spot / Direction / Player before their typedefs (board.h:28,41,75) ----
sed -i 's|^STRUCT_NAMED_PAIR(Spot, spot, Direction, dir, Loc);|typedef short Spot; typedef int8_t Direction; typedef int8_t Player;  /* moved up: same typedefs as board.h:28,41,75 */\nSTRUCT_NAMED_PAIR(Spot, spot, Direction, dir, Loc);|' "$B/game/board.h"
# ---- game/board.h:178 illegal extra qualification `Hash128 Board::getSitHash` inside the class ----
sed -i 's|^  Hash128 Board::getSitHash(Player pla) const;|  Hash128 getSitHash(Player pla) const;|' "$B/game/board.h"
# ---- game/board.h:163 `vector<Placement>` without std:: in a header that has no using-directive ----
sed -i 's|^  bool setStones(vector<Placement> placements);|  bool setStones(std::vector<Placement> placements);|' "$B/game/board.h"

# ---- game/board.cpp:214,220 redeclare `tempSpot` (second time as a Loc) in isLegal: the second walk is a Spot too ----
sed -i 's|^  Loc tempSpot = loc;$|  tempSpot = loc.spot;|' "$B/game/board.cpp"
# ---- ledger 8.1-A: Board copy constructor does not copy win_len (board.cpp:101-109) ----
sed -i 's|^  y_size = other.y_size;$|  y_size = other.y_size;\n  win_len = other.win_len;  /* ledger A */|' "$B/game/board.cpp"
# ---- game/board.cpp:497-963 (text / JSON IO: redeclared `suc`, multi-char literal, Go-era parsers) is not on the path:
#      the file is cut at its "IO FUNCS" banner ----
n=$(grep -n '^// IO FUNCS' "$B/game/board.cpp" | head -1 | cut -d: -f1)
head -n $((n - 1)) "$B/game/board.cpp" > "$B/game/board.cpp.cut" && mv "$B/game/board.cpp.cut" "$B/game/board.cpp"

# ---- game/boardhistory.cpp:181 and :192 define BoardHistory::checkGameEnd twice: the second (which also sets `winner`
#      from moveHistory) is renamed away; printBasicInfo / printDebugInfo need the IO that was cut ----
awk '/^bool BoardHistory::checkGameEnd\(const Board& board\) \{/ { c++; if(c == 2) sub(/checkGameEnd/, "checkGameEndSecondDefinition") } { print }' \
  "$B/game/boardhistory.cpp" > "$B/game/boardhistory.cpp.t" && mv "$B/game/boardhistory.cpp.t" "$B/game/boardhistory.cpp"
sed -i 's|^  bool checkGameEnd(const Board& board);|  bool checkGameEnd(const Board\& board);\n  bool checkGameEndSecondDefinition(const Board\& board);|' "$B/game/boardhistory.h"
n=$(grep -n '^void BoardHistory::printBasicInfo' "$B/game/boardhistory.cpp" | head -1 | cut -d: -f1)
head -n $((n - 1)) "$B/game/boardhistory.cpp" > "$B/game/boardhistory.cpp.cut" && mv "$B/game/boardhistory.cpp.cut" "$B/game/boardhistory.cpp"

# ---- neuralnet/nninputs.h:49 uninitialised namespace-scope `const bool`, redefined non-const at nninputs.cpp:4 ----
sed -i 's|^  const bool historyChannelWithDirection;.*|  extern bool historyChannelWithDirection;|' "$B/neuralnet/nninputs.h"
# ---- ledger 8.1-J: getSymDir falls off the end (assert(false), nninputs.cpp:409-433) when the direction does not change ----
awk '/^Direction SymmetryHelpers::getSymDir/ { f = 1 } f && /^  assert\(false\);$/ { print "  return dir;  /* ledger J */"; f = 0; next } { print }' \
  "$B/neuralnet/nninputs.cpp" > "$B/neuralnet/nninputs.cpp.t" && mv "$B/neuralnet/nninputs.cpp.t" "$B/neuralnet/nninputs.cpp"
# ---- ledger 8.1-L: getSymBoard calls a (size, winLen) constructor with (x, y) (nninputs.cpp:439): give it the win length ----
sed -i 's|^  Board symBoard(transpose ? board.y_size : board.x_size, transpose ? board.x_size : board.y_size);|  Board symBoard(transpose ? board.y_size : board.x_size, transpose ? board.x_size : board.y_size, board.win_len);  /* ledger L */|' "$B/neuralnet/nninputs.cpp"
# ---- ledger 8.1-F/G: from "Feature 11" on fillRowV1 indexes the legal planes by spot instead of pos, runs past channel 15 and calls
#      fillRowWithLine, which walks wall spots out of bounds (board.cpp:392-420).  Planes 0..10 are literal and comparable; the rest
#      is cut off behind a switch so that the literal planes 0..10 can be read safely ----
sed -i 's|^  // Feature 11 or 20 - legal moves$|  if(!NNInputs::kcRefRunBrokenTail) { rowGlobal[0] = (float)board.win_len; return; }  /* ledger F, G */\n  // Feature 11 or 20 - legal moves|' "$B/neuralnet/nninputs.cpp"
sed -i 's|^bool NNInputs::historyChannelWithDirection = false;|bool NNInputs::historyChannelWithDirection = false;\nbool NNInputs::kcRefRunBrokenTail = false;|' "$B/neuralnet/nninputs.cpp"
sed -i 's|^  extern bool historyChannelWithDirection;|  extern bool historyChannelWithDirection;\n  extern bool kcRefRunBrokenTail;|' "$B/neuralnet/nninputs.h"

CXX="${CXX:-g++}"
# -DNDEBUG: fillRowV1 asserts currentFeatureIdx == 11 even when the history is short (ledger F) and whiteWinsOfWinner asserts on a draw
$CXX -std=c++17 -O2 -fPIC -shared -DNDEBUG -DKC_REF_LEGAL_WORDS=13 -w -I"$B" -o "$HERE/_ref/libkc_ref_rules.so" "$HERE/ref_rules_shim.cpp" \
  "$B/game/board.cpp" "$B/game/boardhistory.cpp" "$B/game/graphhash.cpp" "$B/neuralnet/nninputs.cpp" "$B/core/hash.cpp" "$B/core/rand.cpp" "$B/core/rand_helpers.cpp" \
  "$B/core/md5.cpp" "$B/core/sha2.cpp" "$B/core/global.cpp" "$B/core/test.cpp" "$B/core/timer.cpp" "$B/core/bsearch.cpp" -lpthread
echo "built oracle/_ref/libkc_ref_rules.so from $REF (patched scratch copy in oracle/_ref/build)"
