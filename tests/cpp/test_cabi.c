/* The boundary is a C ABI: this file is C99 (no C++), includes include/katacoffee_b200.h and drives the host-only entry points --
 * error reporting, Zobrist tables, the position hash, the evaluator front end over a C batch function, SGF text, the training-data file.
 * Built and run by tests/test_capi_exports.py::test_header_is_c_and_c_program_links (no GPU needed). */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "katacoffee_b200.h"

static int failures = 0;
#define EXPECT(c) do { if(!(c)) { failures++; fprintf(stderr, "FAILED line %d: %s\n", __LINE__, #c); } } while(0)

static int batch_fn(void* user, int server, const kc_eval_batch* b) {
  int i, j;
  (void)server;
  *(int*)user += b->n;
  for(i = 0; i < b->n; i++) {
    for(j = 0; j < 100; j++) b->policyProbs[i * 100 + j] = (float)((b->black[i] >> (j % 30)) & 1);
    b->whiteWinLoss[2 * i] = 0.25f; b->whiteWinLoss[2 * i + 1] = 0.75f;
    b->miscOut[2 * i] = 1.0f; b->miscOut[2 * i + 1] = 2.0f;
  }
  return 0;
}

int main(void) {
  uint64_t board[KC_MAX_ARR_SIZE * 4 * 2], player[4 * 2], sx[(KC_MAX_LEN + 1) * 2], sy[(KC_MAX_LEN + 1) * 2];
  int8_t stones[25];
  int16_t moves[10] = {-1, 0, -1, 0, -1, 0, -1, 0, 7, 1};   /* one move so far: black at cell 7, direction 0 */
  kc_eval_position pos;
  uint64_t h[2], key[2];
  kc_evaluator_config cfg;
  kc_evaluator* ev = NULL;
  kc_eval_output out;
  kc_evaluator_stats st;
  float policy[100];
  int rows = 0, count = -1;
  kc_ctx* ctx = NULL;

  EXPECT(kc_abi_version() == 1);
  EXPECT(kc_zobrist_tables(board, player, sx, sy) == 0);
  memset(stones, 0, sizeof stones);
  stones[7] = 1;
  pos.stones = stones; pos.moves = moves; pos.numTurns = 1; pos.nextPla = 2;
  EXPECT(kc_eval_position_hash(5, 5, &pos, 1.0f, h, key) == 0);
  /* NNInputs::getHash = SIZE_X[5] ^ SIZE_Y[5] ^ BOARD[spot(2,1)][black] ^ PLAYER[white]; spot = (x+1) + (y+1)*(W+1) = 15 */
  EXPECT(h[0] == (sx[5 * 2] ^ sy[5 * 2] ^ board[(15 * 4 + 1) * 2] ^ player[2 * 2]));
  EXPECT(h[1] == (sx[5 * 2 + 1] ^ sy[5 * 2 + 1] ^ board[(15 * 4 + 1) * 2 + 1] ^ player[2 * 2 + 1]));
  EXPECT(key[0] != h[0]);

  memset(&cfg, 0, sizeof cfg);
  cfg.nnXLen = 5; cfg.nnYLen = 5; cfg.winLen = 4; cfg.maxBatch = 4; cfg.maxConcurrentEvals = 8; cfg.numServerThreads = 1;
  cfg.cacheSizePowerOfTwo = 8; cfg.mutexPoolSizePowerOfTwo = 2; cfg.defaultSymmetry = 2; cfg.policyTemperature = 1.0f;
  EXPECT(kc_evaluator_create_custom(&cfg, batch_fn, &rows, &ev) == 0);
  memset(&out, 0, sizeof out);
  out.policyProbs = policy;
  EXPECT(kc_evaluator_evaluate(ev, &pos, KC_SYMMETRY_NOTSPECIFIED, 0, 0, &out) == 0);
  EXPECT(out.symmetry == 2 && out.cacheHit == 0 && out.whiteWinProb == 0.25f && out.shorttermWinlossError == 2.0f && out.nnHash[0] == h[0]);
  EXPECT(policy[8] == 1.0f && policy[7] == 0.0f);   /* cell (2,1) is bit 1*6+2 = 8 of the black bitboard */
  EXPECT(kc_evaluator_evaluate(ev, &pos, KC_SYMMETRY_NOTSPECIFIED, 0, 0, &out) == 0 && out.cacheHit == 1 && rows == 1);
  EXPECT(kc_evaluator_get_stats(ev, &st) == 0 && st.rowsProcessed == 1 && st.cacheHits == 1 && st.cacheMisses == 1);
  EXPECT(kc_evaluator_evaluate(ev, &pos, 9, 0, 0, &out) != 0 && strstr(kc_last_error(), "symmetry") != NULL);
  EXPECT(kc_evaluator_destroy(ev) == 0);

  {
    char sgf[512]; size_t len = 0;
    int16_t mv[2] = {7, 33}; int8_t pl[2] = {1, 2};
    EXPECT(kc_sgf_write(5, 5, 4, "b", "w", NULL, 2, mv, pl, -1, sgf, sizeof sgf, &len) == 0 && len > 0 && strstr(sgf, "SZ[5]") != NULL);
  }
  EXPECT(kc_training_write_npz("/tmp/kc_cabi_test.npz", 0, 5, 5, NULL, NULL, NULL, NULL, NULL) == 0);
  remove("/tmp/kc_cabi_test.npz");
  /* compute entry points fail loudly without a device (or succeed with one); never a silent fallback */
  if(kc_device_count(&count) != 0 || count == 0) EXPECT(kc_ctx_create(0, &ctx) != 0 && strlen(kc_last_error()) > 0);
  if(failures) { fprintf(stderr, "%d check(s) failed\n", failures); return 1; }
  printf("test_cabi: ok\n");
  return 0;
}
