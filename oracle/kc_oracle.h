/*
 * kc_oracle.h -- C API of the CPU ORACLE for the KataCoffee leaf-evaluation hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load libkc_oracle.so.  The product
 * (katacoffee_b200/csrc, include/katacoffee_b200.h) never links, imports or calls it.
 *
 * It is a restatement of the reference algorithm (the reference tree does not compile, SURVEY.md
 * section 0.2), with the canonical resolutions of SURVEY.md section 8.1.  Pinning status:
 *   - MD5 / SHA-256 / xorshift1024* / PCG32 / Rand(seed) / Zobrist chain: PINNED against the
 *     reference's own golden vectors (cpp/core/rand.cpp:41-149,386-439); MD5 and SHA-256 also
 *     against the real reference code compiled into oracle/_ref (cpp/core/md5.cpp, sha2.cpp -- the
 *     only reference translation units on this path that compile standalone).
 *   - conv / batchnorm / residual block / gpool residual block, symmetry copies: PINNED against
 *     the literal vectors of cpp/tests/testnn.cpp and cpp/tests/results/runOutputTests.txt
 *     (tests/golden/, extracted by tests/golden/make_golden.py).
 *   - Coffee rules and hashing over positions -- Board::initHash tables, the Rand stream, isLegal (full masks), playMoveAssumeLegal,
 *     maxConsecutives / checkGameEnd, BoardHistory::makeBoardMove (numTurns / finished / winner), getSitHash, NNInputs::getHash with
 *     every fold, fillRowV1 planes 0..10 + global (NCHW / NHWC), the SymmetryHelpers and NNPos::locToPos: PINNED against the
 *     reference's OWN code -- game/board.cpp, game/boardhistory.cpp, neuralnet/nninputs.cpp, core/hash.cpp, core/rand.cpp compiled
 *     from a patched scratch copy into oracle/_ref/libkc_ref_rules.so (oracle/ref_patch.sh lists each one-line edit with the
 *     SURVEY.md 0.2 defect / 8.1 ledger row behind it) -- over > 10^6 playout positions and arbitrary positions on boards from
 *     2x2 to 10x10 (tests/test_oracle_ref_rules.py).
 *   - PARITY UNPINNED, because the literal code is unusable there (ledger rows, asserted where it can be shown): the draw rule (C),
 *     V1 planes 11..14 (F: legal plane indexed by spot over 4 channels past channel 15; G: fillRowWithLine walks orthogonal
 *     neighbours and wall spots), NNPos::posToLoc (I); and whole-net outputs with Coffee head shapes and the search, for which
 *     no reference test or runnable reference code exists (eigenbackend.cpp needs Eigen and still carries Go heads).  The
 *     restatement follows the cited lines under the ledger and is the definition there.
 */
#ifndef KC_ORACLE_H_
#define KC_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ----------------------------------------------------------------------------------------------
 * Hashing / PRNG (cpp/core/md5.cpp, sha2.cpp, rand.cpp:276-318, rand_helpers.h:29-66, hash.cpp)
 * -------------------------------------------------------------------------------------------- */
void ko_md5(const uint8_t* msg, size_t len, uint32_t out[4]);
void ko_sha256(const uint8_t* msg, size_t len, uint8_t out[32]);
void ko_sha256_u64(const uint8_t* msg, size_t len, uint64_t out[4]);
uint64_t ko_splitmix64(uint64_t x);
uint64_t ko_murmurmix(uint64_t x);
uint64_t ko_rrmxmx(uint64_t x);
uint64_t ko_basic_lcong(uint64_t x);
uint64_t ko_basic_lcong2(uint64_t x);

typedef struct ko_rand ko_rand;
ko_rand* ko_rand_create(const char* seed);
void ko_rand_destroy(ko_rand* r);
uint32_t ko_rand_next_uint(ko_rand* r);
uint64_t ko_rand_next_uint64(ko_rand* r);
double ko_rand_next_double(ko_rand* r);
double ko_rand_next_gaussian(ko_rand* r);
void ko_xorshift1024_test(const uint64_t init_a[16], int n, uint32_t* out);
void ko_pcg32_test(uint64_t state, int n, uint32_t* out);

/* Zobrist tables, Board::initHash (cpp/game/board.cpp:134-178).  Layout (uint64 pairs hash0,hash1):
 *   board  [133][4][2]   (colour index 0..3; empty and wall are zero)
 *   player [4][2]
 *   size_x [11][2], size_y [11][2] */
#define KO_MAX_LEN 10
#define KO_MAX_ARR_SIZE ((KO_MAX_LEN + 1) * (KO_MAX_LEN + 2) + 1)
void ko_zobrist_tables(uint64_t* board, uint64_t* player, uint64_t* size_x, uint64_t* size_y);

/* ----------------------------------------------------------------------------------------------
 * Game: Board + BoardHistory restatement
 * -------------------------------------------------------------------------------------------- */
typedef struct ko_game ko_game;
ko_game* ko_game_create(int x_size, int y_size, int win_len);
void ko_game_destroy(ko_game* g);
void ko_game_reset(ko_game* g);
void ko_game_copy(ko_game* dst, const ko_game* src);
/* Places a stone without history (Board::setStone, board.cpp:258-265). colour 0/1/2. */
int ko_game_set_stone(ko_game* g, int x, int y, int color);
void ko_game_set_last_loc(ko_game* g, int x, int y, int dir); /* x<0 => NULL_LOC,D_NONE */
void ko_game_set_history(ko_game* g, int n, const int32_t* pos, const int32_t* pla, int numTurns, int nextPla);
/* Board::isLegal (board.cpp:185-227) for Loc(spot(x,y),dir) and player pla (1 black, 2 white) */
int ko_game_is_legal(const ko_game* g, int x, int y, int dir, int pla);
/* Full legal mask for the player to move, bit pos = dir*H*W + y*W + x (nninputs.cpp:6-14),
 * out has ceil(4*H*W/32) words.  Returns the number of legal Locs. */
int ko_game_legal_mask(const ko_game* g, int pla, uint32_t* out);
/* BoardHistory::makeBoardMove semantics (boardhistory.cpp:142-176) with ledger rows C, D:
 * returns 0 if illegal (nothing changes), 1 if played. */
int ko_game_play(ko_game* g, int pos);
int ko_game_next_pla(const ko_game* g);
int ko_game_num_turns(const ko_game* g);
int ko_game_finished(const ko_game* g);
int ko_game_winner(const ko_game* g);
int ko_game_max_consecutives(const ko_game* g, int x, int y);
/* status word shared with the CUDA path: bits 0-7 numTurns, bit 8 finished, bits 9-10 winner,
 * bits 11-12 next player */
uint32_t ko_game_status(const ko_game* g);
/* Board::getSitHash(pla) (board.cpp:288-292) */
void ko_game_sit_hash(const ko_game* g, int pla, uint64_t out[2]);
/* NNInputs::getHash (nninputs.cpp:463-502) */
void ko_game_nn_hash(const ko_game* g, int pla, double playoutDoublingAdvantage,
                     float nnPolicyTemperature, double policyOptimism, uint64_t out[2]);
/* NNInputs::fillRowV1 (nninputs.cpp:508-657), canonical 15-channel layout (ledger F,G).
 * rowBin: 15*nnXLen*nnYLen floats (NCHW or NHWC), rowGlobal: 1 float. */
void ko_game_fill_row_v1(const ko_game* g, int pla, int nnXLen, int nnYLen, int useNHWC,
                         float* rowBin, float* rowGlobal);
int ko_game_color_at(const ko_game* g, int x, int y);
/* policy index of the k-th most recent move (0 = last), -1 if there is none */
int ko_game_recent_move_pos(const ko_game* g, int k);
/* GraphHash::getGraphHash (cpp/game/graphhash.cpp:3-28), literal.  Because Coffee has no pass move the reference chains the
 * previous hash in after EVERY move, so the literal hash is path-dependent and never transposes (SURVEY ledger row L in
 * DESIGN.md); the search's transposition key is ko_search's canonical state key instead. */
void ko_graph_hash(const uint64_t prev[2], const ko_game* g, int nextPla, uint64_t out[2]);

/* SymmetryHelpers (nninputs.cpp:252-433) */
void ko_copy_inputs_with_symmetry(const float* src, float* dst, int n, int h, int w, int c,
                                  int useNHWC, int symmetry);
void ko_copy_outputs_with_symmetry(const float* src, float* dst, int n, int h, int w, int symmetry);
int ko_sym_invert(int symmetry);
int ko_sym_compose(int first, int next);
int ko_sym_dir(int dir, int symmetry);
void ko_sym_xy(int x, int y, int xSize, int ySize, int symmetry, int* outX, int* outY);

/* Synthetic random-legal playout of SURVEY.md 8(d): the move of game `gameIdx` at ply t is the
 * (r mod popcount)-th set bit of the legal mask, r = splitmix64(seed ^ gameIdx*0x9E3779B97F4A7C15 ^ t).
 * Returns the chosen pos or -1 if there is no legal move. Also returns r via *rOut if non-null. */
int ko_playout_choose(const ko_game* g, uint64_t seed, uint64_t gameIdx, uint64_t* rOut);

/* Batched trajectory generator used by tests and the CPU baseline: plays games [g0, g0+n) from the
 * empty board to terminal (or maxPlies) and records, per step, the record below AFTER each move
 * (plus the initial position as step 0 of every game).  Any output pointer may be NULL.
 * Returns the total number of records written (<= maxRecords). */
typedef struct {
  uint32_t game;       /* game index */
  uint32_t status;     /* ko_game_status */
  uint32_t legal[13];  /* isLegal mask of the player to move (raw, also when finished); 10x10 needs 400 bits */
  int32_t  movePos;    /* move that led here, -1 for the initial position */
  uint64_t sitHash[2]; /* getSitHash(next player) */
  uint64_t nnHash[2];  /* NNInputs::getHash with default params */
} ko_step_record;
long ko_playout_run(int x_size, int y_size, int win_len, uint64_t seed, uint64_t g0, int n,
                    int maxPlies, ko_step_record* records, long maxRecords,
                    float* planes /* [records][15*H*W] NCHW or NHWC */, int useNHWC,
                    float* globals /* [records] */, int threads);

/* ----------------------------------------------------------------------------------------------
 * Net: model description (POD mirror of cpp/neuralnet/desc.h:13-304; identical memory layout to
 * kc_model_desc in include/katacoffee_b200.h so Python builds one ctypes structure for both).
 * Weight layouts are the ones desc.cpp produces: conv oc,ic,y,x (desc.cpp:131-152), matmul ic,oc
 * (desc.cpp:284-299).
 * -------------------------------------------------------------------------------------------- */
typedef struct { int32_t convYSize, convXSize, inChannels, outChannels; const float* weights; } ko_conv_desc;
typedef struct { int32_t numChannels; float epsilon; int32_t hasScale, hasBias;
                 const float *mean, *variance, *scale, *bias; } ko_bn_desc;
typedef struct { int32_t inChannels, outChannels; const float* weights; } ko_matmul_desc;
typedef struct { int32_t numChannels; int32_t pad_; const float* weights; } ko_matbias_desc;
typedef struct {
  int32_t kind; /* 0 ordinary, 2 global pooling (desc.h:173-175) */
  int32_t preActivation, gpoolActivation, midActivation;
  ko_bn_desc preBN;
  ko_conv_desc regularConv;
  ko_conv_desc gpoolConv;      /* kind 2 only */
  ko_bn_desc gpoolBN;          /* kind 2 only */
  ko_matmul_desc gpoolToBiasMul; /* kind 2 only */
  ko_bn_desc midBN;
  ko_conv_desc finalConv;
} ko_block_desc;
typedef struct {
  int32_t version, numInputChannels, numInputGlobalChannels, numBlocks;
  int32_t trunkNumChannels, midNumChannels, regularNumChannels, gpoolNumChannels;
  int32_t trunkTipActivation, g1Activation, p1Activation, v1Activation, v2Activation, pad_;
  ko_conv_desc initialConv;
  ko_matmul_desc initialMatMul;
  const ko_block_desc* blocks;
  ko_bn_desc trunkTipBN;
  /* policy head (desc.h:216-240) */
  ko_conv_desc p1Conv, g1Conv;
  ko_bn_desc g1BN;
  ko_matmul_desc gpoolToBiasMul;
  ko_bn_desc p1BN;
  ko_conv_desc p2Conv;
  /* value head (desc.h:242-269) */
  ko_conv_desc v1Conv;
  ko_bn_desc v1BN;
  ko_matmul_desc v2Mul;
  ko_matbias_desc v2Bias;
  ko_matmul_desc v3Mul;
  ko_matbias_desc v3Bias;
  ko_matmul_desc sv3Mul;
  ko_matbias_desc sv3Bias;
  ko_conv_desc vOwnershipConv;
} ko_model_desc;

typedef struct ko_model ko_model;
ko_model* ko_model_create(const ko_model_desc* desc); /* deep-copies all weights */
void ko_model_destroy(ko_model* m);
/* NeuralNet::getOutput restatement (eigenbackend.cpp:1675-1844): inputs rowSpatial [n][15*H*W]
 * (NHWC iff inputsNHWC) + rowGlobal [n][1] + symmetry [n]; outputs logits, inverse-symmetrised,
 * policy in NNPos order [n][4*H*W], value [n][2], misc [n][2], ownership [n][H*W] (may be NULL).
 * mode 0 = direct fp32 convolution (the checker), 1 = Winograd F(4x4,3x3) + GEMM as the Eigen
 * backend does (eigenbackend.cpp:417-667; used for the CPU baseline), 2 = mode 0 with every
 * tensor-core convolution's weights and input activations rounded to bf16 (precision model of the
 * tcgen05 path: separates rounding from kernel bugs in the parity tests).  KO_MODE_EMUL(a, w) selects
 * the operand formats of that emulation: activations a, weights w; 0 bf16, 1 fp16, 2 fp32 (exact). */
#define KO_MODE_EMUL(a, w) (2 | ((a) << 4) | ((w) << 6))
void ko_model_forward(const ko_model* m, int n, int nnXLen, int nnYLen, int inputsNHWC,
                      const float* rowSpatial, const float* rowGlobal, const int8_t* symmetry,
                      float* policy, float* value, float* misc, float* ownership, int mode,
                      int threads);
/* Diagnostic: one un-chunked forward (NCHW rows, no symmetry) that also copies the trunk after the initial conv, after
 * every block and after trunkTipBN into trace [numBlocks + 2][n][H*W][trunkC] (NHWC). */
void ko_model_forward_trace(const ko_model* m, int n, int nnXLen, int nnYLen, const float* rowSpatial, const float* rowGlobal,
                            float* policy, float* value, float* misc, float* ownership, int mode, float* trace);
/* Layer-level hooks = NeuralNet::testEvaluate* (nninterface.h:127-169); buffers NHWC or NCHW. */
void ko_test_conv(const ko_conv_desc* d, int n, int xLen, int yLen, int useNHWC, const float* in,
                  float* out, int mode);
void ko_test_batchnorm(const ko_bn_desc* d, int activation, int n, int xLen, int yLen, int useNHWC,
                       const float* in, const float* mask, float* out);
void ko_test_resblock(const ko_block_desc* d, int n, int xLen, int yLen, int useNHWC,
                      const float* in, const float* mask, float* out, int mode);
/* NNEvaluator::evaluate post-processing (nneval.cpp:702-815): in-place on one row. */
void ko_postprocess(float* policy, int policySize, const uint32_t* legalMask, float policyTemp,
                    float* value2 /* win,loss logits -> whiteWin, whiteLoss */, float* misc2,
                    int nextPla);

/* ----------------------------------------------------------------------------------------------
 * Tree search (ko_search.cpp): single-game restatement of the reference's playout loop under SearchParams()
 * defaults with valueWeightExponent 0 (see the file header).  modelOrNull == NULL selects the integer-hash
 * evaluator that the CUDA search also implements, so trees can be compared exactly.
 * -------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t maxVisits, temperaturePlies, autoRefill, noCompaction, reuseTree, useGraphSearch;
  double cpuctExploration, fpuReductionMax, rootFpuReductionMax;
  double subtreeValueBiasFactor, subtreeValueBiasWeightExponent, subtreeValueBiasFreeProp;
  int32_t rootNoiseEnabled, fpuParentWeightByVisitedPolicy;
  double rootDirichletNoiseTotalConcentration, rootDirichletNoiseWeight;
  double rootPolicyTemperature, rootPolicyTemperatureEarly, chosenMoveTemperatureHalflife;
  double fpuParentWeightByVisitedPolicyPow, rootDesiredPerChildVisitsCoeff, valueWeightExponent;
  uint64_t noiseSeed, noiseGameId;   /* oracle only (one game per call): what kc_search_reset's seed and the game id are on the device */
  int32_t nnRandomize, pad3_;        /* leaves evaluated under a symmetry drawn from (noiseSeed, sit-hash) */
  /* the rest of cpp/configs/training/selfplay1.cfg:144-185 and of the GTP / analysis defaults (setup.cpp:520-560) */
  int32_t useLcbForSelection, useNonBuggyLcb;
  double lcbStdevs, minVisitPropForLCB;
  int32_t rootNumSymmetriesToSample, useNoisePruning;
  int32_t useUncertainty, pad4_;
  double uncertaintyCoeff, uncertaintyExponent, uncertaintyMaxWeight;
  double chosenMoveSubtract, chosenMovePrune;           /* also applied to the root's children in recomputeNodeStats when the root is noised */
  double noisePruneUtilityScale, noisePruningCap;       /* 0.15, 1e50 (searchparams.cpp:27-28) */
} ko_search_params;
void ko_search_last_play_selection(double* out, int P);
int ko_search_choose_values(const double* values, const uint8_t* order, int P, int boardArea, int ply, double tempEarly, double tempLate,
                            double halflife, double subtract, double prune, uint64_t seed, uint64_t gameId);
void ko_search_run(const ko_game* rootGame, int x_size, int y_size, const ko_search_params* p, const ko_model* modelOrNull,
                   int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits, double* edgeUtilitySum, float* policyOut,
                   uint8_t* orderOut, uint64_t counters[3] /* += visits, evaluations, terminal visits */);
/* Graph search (transpositions) and subtree value bias: see the block comment in ko_search.cpp.  counters[5] += visits,
 * evaluations, terminal visits, transposition hits, catch-up visits. */
void ko_search_run_graph(const ko_game* rootGame, int x_size, int y_size, const ko_search_params* p, const ko_model* modelOrNull,
                         int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits, double* edgeUtilitySum, float* policyOut,
                         uint8_t* orderOut, uint64_t counters[5], uint64_t* digest);
typedef struct ko_graph_search ko_graph_search;   /* persistent graph: continue() searches on, advance() re-roots (tree re-use) */
ko_graph_search* ko_graph_search_create(int x_size, int y_size, const ko_search_params* p);
void ko_graph_search_destroy(ko_graph_search* s);
void ko_graph_search_continue(ko_graph_search* s, const ko_game* rootGame, const ko_model* modelOrNull, int32_t* rootVisits, double* rootUtilitySum,
                              int32_t* edgeVisits, double* edgeUtilitySum, float* policyOut, uint8_t* orderOut, uint64_t counters[5], uint64_t* digest);
void ko_graph_search_advance(ko_graph_search* s, int movePos);
uint64_t ko_graph_search_digest(const ko_graph_search* s);
int ko_graph_search_num_nodes(const ko_graph_search* s);
typedef struct ko_search ko_search;   /* persistent tree: continue() searches on, advance() re-roots at the move played */
ko_search* ko_search_create(void);
void ko_search_destroy(ko_search* s);
void ko_search_clear(ko_search* s);
void ko_search_continue(ko_search* s, const ko_game* rootGame, int x_size, int y_size, const ko_search_params* p, const ko_model* modelOrNull,
                        int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits, double* edgeUtilitySum, float* policyOut,
                        uint8_t* orderOut, uint64_t counters[3]);
void ko_search_advance(ko_search* s, int movePos);
int ko_search_choose(const int32_t* edgeVisits, const uint8_t* order, int P, int ply, int temperaturePlies, uint64_t seed, uint64_t gameId);
/* the reference's temperature schedule for the move choice (see ko_search.cpp) */
int ko_search_choose_temperature(const int32_t* edgeVisits, const uint8_t* order, int P, int boardArea, int ply, double tempEarly, double tempLate,
                                 double halflife, double subtract, double prune, uint64_t seed, uint64_t gameId);
/* Training rows of one finished game (restatement of TrainingWriteBuffers::addRow, cpp/dataio/trainingwrite.cpp:316-566, with the
 * canonical choices documented at kc_search_read_training_rows); arrays as in the reference's npz, R rows. */
void ko_training_rows(int x_size, int y_size, int win_len, int R, const int32_t* movePos, const int32_t* rootN, const double* rootW,
                      const int16_t* visits, uint64_t gameId, uint8_t* bin, float* globalIn, int16_t* policy, float* globalT, int8_t* value);

#ifdef __cplusplus
}
#endif
#endif
