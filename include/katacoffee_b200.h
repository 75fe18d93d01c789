/*
 * katacoffee_b200.h -- C ABI of the B200-native KataCoffee leaf-evaluation hot path.
 *
 * This is the drop-in boundary: what a `USE_BACKEND=B200` build of the reference would link
 * (cpp/neuralnet/b200backend.cpp implementing cpp/neuralnet/nninterface.h:31-171 forwards to
 * these entry points; see INTEGRATION.md for the shim) plus the batched rules/features/hash
 * entry points that have no counterpart in the reference (it runs them on the CPU one position
 * at a time: cpp/game/board.cpp, cpp/game/boardhistory.cpp, cpp/neuralnet/nninputs.cpp).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; kc_last_error() returns the
 *    thread-local message (the C++ shim rethrows it as StringError, the reference's convention,
 *    cpp/neuralnet/nneval.cpp:327-330).
 *  - no C++ or torch types cross the ABI: plain pointers and sizes only.
 *  - pointers are HOST pointers unless the parameter name ends in `_dev`; the library owns all
 *    device memory and frees it in the matching _destroy.
 *  - kc_handle / kc_games own a CUDA stream and may be used by one thread at a time (mirrors
 *    nninterface.h:18-20: a ComputeHandle is used by exactly one server thread); kc_model is
 *    immutable and shareable (mirrors LoadedModel, nninterface.h:42-43).
 *  - there is no CPU fallback: every compute entry point fails if no sm_100 device is present.
 */
#ifndef KATACOFFEE_B200_H_
#define KATACOFFEE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KC_MAX_LEN 10                                        /* cpp/game/board.h:15  */
#define KC_MAX_ARR_SIZE ((KC_MAX_LEN + 1) * (KC_MAX_LEN + 2) + 1) /* board.h:120-124 */
#define KC_NUM_SPATIAL_V1 15                                 /* SURVEY.md 8.1-F      */
#define KC_NUM_GLOBAL_V1 1
#define KC_MAX_DEVICE_LEN 7  /* the 64-bit bitboard kernels and the device search: H*(W+1) <= 64.  kc_games_*, the net paths and the evaluator
                                front end take every size up to KC_MAX_LEN (beyond 7x7: general kernels on 128-bit bitboards) */

typedef struct kc_ctx kc_ctx;
typedef struct kc_model kc_model;
typedef struct kc_handle kc_handle;
typedef struct kc_games kc_games;

const char* kc_last_error(void);
int kc_abi_version(void);

/* ---------------------------------------------------------------------------------------------
 * Context.  Replaces NeuralNet::globalInitialize + createComputeContext (nninterface.h:33,52-66)
 * and Board::initHash (cpp/game/board.cpp:134-178: the Zobrist tables are generated on the host by
 * the same MD5 -> SHA-256 -> xorshift1024* + PCG32 chain and uploaded to __constant__ memory).
 * ------------------------------------------------------------------------------------------- */
int kc_device_count(int* count);
int kc_ctx_create(int device, kc_ctx** out);
int kc_ctx_destroy(kc_ctx* ctx);
/* Host-side Zobrist tables (uint64 pairs hash0,hash1): board [133][4][2] (colour index 0..3, empty
 * and wall zero), player [4][2], size_x [11][2], size_y [11][2].  No GPU needed. */
int kc_zobrist_tables(uint64_t* board, uint64_t* player, uint64_t* size_x, uint64_t* size_y);

/* ---------------------------------------------------------------------------------------------
 * Model description: POD mirror of ModelDesc (cpp/neuralnet/desc.h:13-304).  Weight layouts are
 * what desc.cpp produces: conv oc,ic,y,x (desc.cpp:131-152), matmul ic,oc (desc.cpp:284-299).
 * Activations: 0 identity, 1 ReLU, 2 Mish (cpp/neuralnet/activations.h:4-6).
 * ------------------------------------------------------------------------------------------- */
typedef struct { int32_t convYSize, convXSize, inChannels, outChannels; const float* weights; } kc_conv_desc;
typedef struct { int32_t numChannels; float epsilon; int32_t hasScale, hasBias;
                 const float *mean, *variance, *scale, *bias; } kc_bn_desc;
typedef struct { int32_t inChannels, outChannels; const float* weights; } kc_matmul_desc;
typedef struct { int32_t numChannels; int32_t pad_; const float* weights; } kc_matbias_desc;
typedef struct {
  int32_t kind; /* 0 ordinary, 2 global pooling (desc.h:173-175) */
  int32_t preActivation, gpoolActivation, midActivation;
  kc_bn_desc preBN;
  kc_conv_desc regularConv;
  kc_conv_desc gpoolConv;        /* kind 2 only */
  kc_bn_desc gpoolBN;            /* kind 2 only */
  kc_matmul_desc gpoolToBiasMul; /* kind 2 only */
  kc_bn_desc midBN;
  kc_conv_desc finalConv;
} kc_block_desc;
typedef struct {
  int32_t version, numInputChannels, numInputGlobalChannels, numBlocks;
  int32_t trunkNumChannels, midNumChannels, regularNumChannels, gpoolNumChannels;
  int32_t trunkTipActivation, g1Activation, p1Activation, v1Activation, v2Activation, pad_;
  kc_conv_desc initialConv;
  kc_matmul_desc initialMatMul;
  const kc_block_desc* blocks;
  kc_bn_desc trunkTipBN;
  kc_conv_desc p1Conv, g1Conv;   /* policy head (desc.h:216-240), Coffee shapes: p2Conv out = 4 */
  kc_bn_desc g1BN;
  kc_matmul_desc gpoolToBiasMul;
  kc_bn_desc p1BN;
  kc_conv_desc p2Conv;
  kc_conv_desc v1Conv;           /* value head (desc.h:242-269): v3 out = 2, sv3 out = 2 */
  kc_bn_desc v1BN;
  kc_matmul_desc v2Mul;
  kc_matbias_desc v2Bias;
  kc_matmul_desc v3Mul;
  kc_matbias_desc v3Bias;
  kc_matmul_desc sv3Mul;
  kc_matbias_desc sv3Bias;
  kc_conv_desc vOwnershipConv;
} kc_model_desc;

/* Replaces NeuralNet::loadModelFile + the backend Model constructor (nninterface.h:42,
 * eigenbackend.cpp:1380-1420): validates shapes, folds BN (scale/sqrt(var+eps), bias-mean*scale),
 * re-tiles and converts the weights for both device paths. */
int kc_model_create(kc_ctx* ctx, const kc_model_desc* desc, kc_model** out);
int kc_model_destroy(kc_model* model);

/* ---------------------------------------------------------------------------------------------
 * Model files.  Replaces ModelDesc::loadFromFileMaybeGZipped (cpp/neuralnet/desc.cpp:1146-1204) and
 * the ModelDesc / layer parsers behind it (desc.cpp:27-92 readFloats with "@BIN@" blocks, :107-155 conv, :175-219 batch norm,
 * :239-258 activation, :274-340 matmul / matbias, :354-456 blocks, :562-696 trunk, :751-925 heads, :977-1094 model) for Coffee
 * models: version 1, 15 + 1 input channels, policy 4 channels (+ the format's gpoolToPassMul, parsed and ignored), value 2,
 * misc 2, ownership 1.  File kinds by suffix: .txt, .bin, .txt.gz, .bin.gz, .gz (binary first, then text).
 * expectedSha256 (NULL or "" = do not check) is compared with the SHA-256 of the file as stored (cpp/core/fileutils.cpp:113-140).
 * No GPU is needed.  The returned description stays valid until kc_modelfile_free; kc_model_create copies what it needs.
 * ------------------------------------------------------------------------------------------- */
typedef struct kc_modelfile kc_modelfile;
int kc_modelfile_load(const char* path, const char* expectedSha256, kc_modelfile** out);
int kc_modelfile_free(kc_modelfile* f);
const kc_model_desc* kc_modelfile_desc(const kc_modelfile* f);
const char* kc_modelfile_name(const kc_modelfile* f);
const char* kc_modelfile_sha256(const kc_modelfile* f);
/* Writes `desc` in the same format (suffix .txt, .bin, .txt.gz or .bin.gz); activations are written with their kind. */
int kc_modelfile_write(const kc_model_desc* desc, const char* name, const char* path);

/* ---------------------------------------------------------------------------------------------
 * Custom Coffee SGF (README.md:33-35, cpp/dataio/sgf.cpp:42-154, 1526-1548): a move is B[xyd] / W[xyd],
 * x and y the column / row letters (a..z, A..Z) and d in a..d the direction | - \ / (0..3); header
 * "(;FF[4]GM[Coffee]SZ[n]WLL[k]PB[..]PW[..]RE[B+|W+|0]", AB / AW placements = the starting position.  Moves are policy indices
 * dir*H*W + y*W + x (-1 = no move), players 1 black / 2 white, winner -1 unfinished / 0 draw / 1 / 2.  Host-only.
 * kc_sgf_write: *outLen receives the length; fails if outCap is too small.  kc_sgf_parse follows the main line; initialStones
 * [H*W] (may be NULL); *numMoves is the number of moves in the file even if it exceeds maxMoves.
 * ------------------------------------------------------------------------------------------- */
int kc_sgf_write(int xSize, int ySize, int winLen, const char* blackName, const char* whiteName, const int8_t* initialStones, int numMoves,
                 const int16_t* movePos, const int8_t* movePla, int winner, char* out, size_t outCap, size_t* outLen);
int kc_sgf_parse(const char* sgf, int* xSize, int* ySize, int* winLen, int8_t* initialStones, int maxMoves, int16_t* movePos,
                 int8_t* movePla, int* numMoves, int* winner);

/* ---------------------------------------------------------------------------------------------
 * Compute handle.  Replaces createComputeHandle / createInputBuffers / getOutput
 * (nninterface.h:77-117).
 * ------------------------------------------------------------------------------------------- */
#define KC_FLAG_FP32_CHECK 1u       /* CUDA-core fp32 path (1e-4 gate); default is the tcgen05 tensor-core path */
#define KC_FLAG_INPUTS_NHWC 2u      /* kc_forward spatial rows are NHWC (inputsUseNHWC)          */
#define KC_FLAG_SYM_PERMUTE_DIRS 4u /* play mode, SURVEY.md 8.1-K: a symmetry also permutes the direction channels (inputs 3..6 by
                                       getSymDir, cpp/neuralnet/nninputs.cpp:409-433, the 4 policy channels by its inverse), which
                                       makes it a true symmetry of the game; default is the reference backends' spatial-only copy.
                                       Applies to kc_forward and kc_games_eval; kc_games_features is always spatial-only. */

#define KC_FLAG_OPERANDS_BF16 8u    /* tensor-core path: weights and activations as bf16 operands.  Default is fp16 operands (tcgen05
                                       kind::f16 takes either format at the same rate, fp32 accumulation and an fp32 residual stream
                                       in both): 3 more mantissa bits bring raw policy / ownership logits inside 1e-2 of the fp32
                                       reference (bf16 operands: 2-4e-2, profiles/r02_bf16_error_by_layer.json).  Activations and
                                       weights saturate at +-65504 instead of overflowing. */
#define KC_FLAG_MASKED_BOARDS 16u   /* tensor-core path: boards may be smaller than nnXLen x nnYLen (requireExactNNLen = false, nninterface.h:73-76).
                                     * Input channel 0 (the on-board plane) is the mask, as in the reference backends (eigenbackend.cpp:1438):
                                     * off-board cells are zeroed after every normalisation, the pooling layers divide by each board's own
                                     * cell count and take the maximum over on-board cells only (:141-166).  The fp32 check path always does this. */

/* Page-locked host memory for the buffers handed to kc_forward (InputBuffers of nninterface.h:92-93): with pinned rows the
 * chunked H2D / D2H copies of kc_forward are asynchronous DMA that overlap the kernels; pageable memory works but serialises. */
int kc_host_alloc(size_t bytes, void** out);
int kc_host_free(void* p);

int kc_handle_create(kc_ctx* ctx, const kc_model* model, int maxBatch, int nnXLen, int nnYLen,
                     unsigned flags, kc_handle** out);
int kc_handle_destroy(kc_handle* h);
int kc_handle_uses_bf16(const kc_handle* h); /* isUsingFP16 analogue, nninterface.h:88: 1 on the tensor-core path (either 16-bit operand format) */
int kc_handle_operand_format(const kc_handle* h); /* arithmetic type of the convolutions' operands: 0 fp16, 1 bf16 (both: fp32 accumulation), 2 fp32 (check path) */
/* NeuralNet::getOutput (nninterface.h:112-117).  Inputs: spatial [n][15*H*W] fp32 (NCHW, or NHWC
 * with KC_FLAG_INPUTS_NHWC), global [n][1], symmetry [n] in 0..7 (NULL = 0).  Outputs are raw
 * logits, inverse-symmetrised, NNPos order (cpp/neuralnet/nninputs.cpp:6-14): policy [n][4*H*W],
 * value [n][2] (win, loss), misc [n][2] (varTimeLeft, shorttermWinlossError pre-softplus),
 * ownership [n][H*W] or NULL.  Synchronous: results are valid on return.  0 < n <= maxBatch. */
int kc_forward(kc_handle* h, int n, const float* spatial, const float* global, const int8_t* symmetry,
               float* policy, float* value, float* misc, float* ownership);
/* The same call for rows that lie scattered in host memory -- what NeuralNet::getOutput is handed: spatialRows[i] / globalRows[i] are
 * NNResultBuf::rowSpatial / rowGlobal (nneval.h:45-65), policyRows[i] is NNOutput::policyProbs, scalarRows[i] points at NNOutput's
 * four contiguous scalars {whiteWinProb, whiteLossProb, varTimeLeft, shorttermWinlossError} (nninputs.h:75-90; all logits),
 * ownerRows[i] is NNOutput::whiteOwnerMap (the array or any entry may be NULL).  The library gathers into its own page-locked
 * staging with worker threads (KC_FORWARD_ROWS_THREADS, default min(4, cores / 2); batches under three chunks stay on the calling thread) and pipelines gather / H2D / kernels / D2H /
 * scatter over row chunks; results are bit-identical to kc_forward on the gathered rows. */
int kc_forward_rows(kc_handle* h, int n, const float* const* spatialRows, const float* const* globalRows, const int8_t* symmetry,
                    float* const* policyRows, float* const* scalarRows, float* const* ownerRows);
/* Copies the outputs of the last device-resident evaluation (kc_games_eval) to the host. */
int kc_handle_read_outputs(kc_handle* h, int n, float* policy, float* value, float* misc, float* ownership);
/* Number of kernel launches this handle has issued (for bench.py's gpu_launches). */
int64_t kc_handle_launch_count(const kc_handle* h);
/* Sum (ms) and count of the device durations of the trunk-kernel launches issued through this
 * handle since the previous call (CUDA events recorded on the launching stream); synchronises. */
int kc_handle_trunk_time(kc_handle* h, float* sumMs, int* count);
/* Diagnostic (handles created with KC_TRUNK_PROBE=1 in the environment): SM-clock timestamps of one layer boundary of
 * the last trunk launch, CTA 0 / first work item / tile 0, out[32]: 0 last MMA of layer 5 issued, 1 epilogue saw the
 * accumulator barrier, 2 first TMEM load done, 3 first 16 channels published, 7 last published, 6 issuer of layer 6
 * starts waiting, 4 its wait is over, 5 its first MMAs are issued, 8..14 the MMAs of its input chunks 0..6 are issued,
 * 15 its last MMA is issued; item boundary: 16 head conv issued, 17 head epilogue starts, 18 it has released TMEM,
 * 19 it ends, 20 issuer starts the next item, 21 has issued its layer 0, 22 epilogue sees layer 0, 23 issuer has
 * the first chunk of layer 1; 24..29 phases inside the head epilogue.  out[64 + ((tile * 48 + layer) * 8 + k)]: the whole second
 * item of CTA 0 -- k = 0 the MMA issuer reaches the layer, 1 its first input chunk is there, 2 the layer is issued, 3 the epilogue
 * sees the accumulator, 4 the epilogue is done (tests/diag_timeline.py prints it).  out must hold 64 + 2 * 48 * 8 words. */
int kc_handle_trunk_probe(kc_handle* h, int64_t* out);
/* Self-test of the tcgen05 descriptor conventions the trunk kernel relies on: D[128][N] =
 * A[shift:shift+128][K] * B[N][K]^T on the tensor core (A, B bf16 bit patterns, D fp32).  ws != 0 uses the
 * weight-stationary form: D gets two results [2][128][N], for row shifts `shift` and `shift+1`, the second MMA
 * re-using the B operand latched by the first. */
int kc_selftest_umma(kc_ctx* ctx, const uint16_t* A, const uint16_t* B, float* D, int rowsA, int N, int K, int shift, int ws);

/* Layer-level hooks = NeuralNet::testEvaluateConv / BatchNorm / ResidualBlock /
 * GlobalPoolingResidualBlock (nninterface.h:127-169), fp32 path; buffers NCHW or NHWC. */
int kc_test_conv(kc_ctx* ctx, const kc_conv_desc* d, int n, int xLen, int yLen, int useNHWC,
                 const float* in, float* out);
int kc_test_batchnorm(kc_ctx* ctx, const kc_bn_desc* d, int activation, int n, int xLen, int yLen,
                      int useNHWC, const float* in, const float* mask, float* out);
int kc_test_resblock(kc_ctx* ctx, const kc_block_desc* d, int n, int xLen, int yLen, int useNHWC,
                     const float* in, const float* mask, float* out);

/* ---------------------------------------------------------------------------------------------
 * Batched games: device-resident Board + BoardHistory for G concurrent games.
 * Semantics per game are exactly the reference's (ledger 8.1):
 * Board::isLegal (board.cpp:185-227), playMoveAssumeLegal (:427-435), checkGameEnd (:376-383),
 * BoardHistory::makeBoardMoveAssumeLegal (boardhistory.cpp:157-176), getSitHash (board.cpp:288-292),
 * NNInputs::fillRowV1 (nninputs.cpp:508-657), copyInputsWithSymmetry (nninputs.cpp:252-357).
 * ------------------------------------------------------------------------------------------- */
int kc_games_create(kc_ctx* ctx, int numGames, int xSize, int ySize, int winLen, kc_games** out);
int kc_games_destroy(kc_games* g);
/* All games back to the empty board, black to move; game i gets id firstGameId + i.  `autoRefill`
 * != 0 makes kc_games_step restart a finished game with a fresh id (ids continue after the last
 * one handed out) instead of leaving it finished. */
int kc_games_reset(kc_games* g, uint64_t seed, uint64_t firstGameId, int autoRefill);
/* Overwrites games [g0, g0+n) with explicit positions (test hook): stones [n][H*W] (0 empty,
 * 1 black, 2 white), nextPla [n], moves [n][5][2] = last five (pos, pla) pairs oldest first with
 * pos = -1 for "none" (pos is a policy index dir*H*W + y*W + x), numTurns [n]. */
int kc_games_load(kc_games* g, int g0, int n, const int8_t* stones, const int8_t* nextPla,
                  const int16_t* moves, const int32_t* numTurns);
/* One ply for every unfinished game.  movePos [G] (policy index, -1 = do nothing) or NULL for
 * the counter-RNG random-legal move of SURVEY.md 8(d) (also selected per game by movePos == -2):
 *   r = splitmix64(seed ^ gameId*0x9E3779B97F4A7C15 ^ ply), move = (r mod popcount)-th legal bit.
 * Illegal movePos leaves the game unchanged and sets bit 15 of its status word.
 * Outputs (any may be NULL), describing the position AFTER the move:
 *   legal   [G][ceil(4*H*W/32)]  isLegal mask of the player to move, bit = policy index
 *   status  [G]                  bits 0-7 numTurns, 8 finished, 9-10 winner, 11-12 next player
 *   sitHash [G][2]               Board::getSitHash(next player)
 *   played  [G]                  policy index played this ply (-1 if none)
 *   gameIds [G]                  id of the game occupying each lane after the step */
int kc_games_step(kc_games* g, const int16_t* movePos, uint32_t* legal, uint32_t* status,
                  uint64_t* sitHash, int16_t* played, uint64_t* gameIds);
/* NNInputs::fillRowV1 for every game's current position (+ optional per-game symmetry, exactly
 * copyInputsWithSymmetry).  layout 0 = NCHW, 1 = NHWC; planes [G][15*H*W] fp32; global [G][1]. */
int kc_games_features(kc_games* g, int layout, const int8_t* symmetry, float* planes, float* global);
/* Device-resident evaluation of the current positions: V1 planes are generated in bf16 (or fp32
 * for a KC_FLAG_FP32_CHECK handle) directly into the handle's input buffer and the net is run;
 * nothing crosses PCIe.  Read the results with kc_handle_read_outputs. numGames <= maxBatch. */
int kc_games_eval(kc_games* g, kc_handle* h, const int8_t* symmetry);
/* NNEvaluator::evaluate post-processing (cpp/neuralnet/nneval.cpp:702-815) of the last kc_games_eval, on the
 * device: policyProbs [G][4*H*W] = legal-masked softmax with temperature (illegal = -1), whiteWinLoss [G][2] =
 * (whiteWinProb, whiteLossProb), misc [G][2] = (varTimeLeft, shorttermWinlossError) after softplus and the
 * model's multipliers, nnHash [G][2] = NNInputs::getHash with default parameters (nninputs.cpp:463-470).
 * Any output may be NULL. */
int kc_games_postprocess(kc_games* g, kc_handle* h, float policyTemperature, float* policyProbs, float* whiteWinLoss, float* misc,
                         uint64_t* nnHash);
/* Fused hot-path step used by bench.py: [random-legal step (+refill) -> planes -> forward] x plies,
 * all on the device, no host synchronisation between plies.  h may be NULL (rules+features only:
 * fp32 NCHW planes, legal masks, status words and sit-hashes are written to device buffers owned
 * by the games object every ply). */
typedef struct {
  uint64_t steps;      /* game-steps executed (one per live game per ply) */
  uint64_t evals;      /* positions sent through the net */
  uint64_t gamesFinished, blackWins, whiteWins, draws;
  uint64_t checksum;   /* XOR of sitHash0 of every position reached (cheap cross-check vs oracle) */
} kc_stats;
int kc_games_run(kc_games* g, kc_handle* h, int plies, kc_stats* statsAccum);
/* Same, timed on the device: every ply is bracketed by a CUDA-event pair on the launching stream and
 * *msTotal receives the sum of the per-ply durations; flushL2Bytes > 0 overwrites a scratch buffer of
 * that size between plies (outside the timed windows) so no ply finds its inputs in L2 by accident. */
int kc_games_run_timed(kc_games* g, kc_handle* h, int plies, size_t flushL2Bytes, kc_stats* statsAccum, float* msTotal);
/* What the last ply of the last rules+features run (kc_games_run / kc_games_run_timed with h == NULL) left in the device buffers:
 * planes [G][15*H*W] fp32 NCHW, global [G], legal [G][LW], status [G], sitHash [G][2], played [G]; any may be NULL. */
int kc_games_read_run_outputs(kc_games* g, float* planes, float* global, uint32_t* legal, uint32_t* status, uint64_t* sitHash, int16_t* played);
/* The same for one of the last four plies of the last rules+features launch (pliesBack 0 = the last ply, 1 = the one before, ...;
 * must be below min(4, plies the last launch stepped)): every ply of a launch writes its own slot of a 4-slot ring, so per-step masks,
 * win/draw status words, sit-hashes, moves and (when the run used the plane ring, flushL2Bytes > 0) V1 planes are all delivered. */
int kc_games_read_run_ply(kc_games* g, int pliesBack, float* planes, uint32_t* legal, uint32_t* status, uint64_t* sitHash, int16_t* played);
int64_t kc_games_launch_count(const kc_games* g);
/* Average device time in ms per ply of the rules+features kernel in the last kc_games_run. */
float kc_games_last_kernel_ms(const kc_games* g);

/* ---------------------------------------------------------------------------------------------
 * Batched tree search: the caller of the leaf-evaluation path.  G games, one PUCT tree each,
 * advanced in lock step: per iteration every game descends to one leaf, the G leaves are evaluated as one batch
 * (rules -> V1 planes -> net -> NNEvaluator post-processing, all on the device) and backed up.
 * Reference: Search::playoutDescend (cpp/search/search.cpp:935-1160), selectBestChildToDescend and
 * getExploreSelectionValue (cpp/search/searchexplorehelpers.cpp:9-45, 323-451), addLeafValue / recomputeNodeStats
 * (cpp/search/searchupdatehelpers.cpp:12-76, 151-326), with SearchParams() defaults (cpp/search/searchparams.cpp:8-90)
 * and valueWeightExponent 0 (see the header of csrc/search.cu for the exact canonical semantics).
 * ------------------------------------------------------------------------------------------- */
typedef struct kc_search kc_search;
typedef struct {
  int32_t maxVisits;            /* SearchParams::maxVisits: visits of the root per move (incl. its own evaluation) */
  int32_t temperaturePlies;     /* plies played in proportion to visits (chosenMoveTemperature 1), afterwards the most visited move */
  int32_t autoRefill;           /* restart finished games with fresh ids before the next search */
  int32_t noCompaction;         /* 0 (default): on the bf16 path only the leaves that need the net are batched (terminal
                                   visits take no row); 1: one row per game, idle rows evaluated and ignored */
  int32_t reuseTree;            /* keep the subtree of the move played for the next search (Search::makeMove); its visits count
                                   towards maxVisits, so later searches need fewer evaluations */
  int32_t useGraphSearch;       /* SearchParams::useGraphSearch: positions with equal stones, player to move and last move share one
                                   node (cpp/search/search.cpp:704-757) */
  double cpuctExploration;      /* SearchParams::cpuctExploration (1.0) */
  double fpuReductionMax;       /* SearchParams::fpuReductionMax (0.2) */
  double rootFpuReductionMax;   /* SearchParams::rootFpuReductionMax (0.2) */
  double subtreeValueBiasFactor;          /* SearchParams::subtreeValueBiasFactor (0 = off; selfplay1.cfg:180 uses 0.30) */
  double subtreeValueBiasWeightExponent;  /* SearchParams::subtreeValueBiasWeightExponent (0.5 default, selfplay1.cfg:181 0.8) */
  double subtreeValueBiasFreeProp;        /* SearchParams::subtreeValueBiasFreeProp (0.8): share of a dropped node's contribution
                                             given back to its entry when the tree is re-used */
  /* Further SearchParams of the self-play configuration (cpp/configs/training/selfplay1.cfg:144-185); zero = off.  Any of them
   * (like useGraphSearch and the bias) selects the graph mode of the search. */
  int32_t rootNoiseEnabled;               /* shaped Dirichlet noise on the root policy (searchhelpers.cpp:51-120), drawn from a
                                             counter-based stream keyed by (kc_search_reset's seed, game id, ply) */
  int32_t fpuParentWeightByVisitedPolicy; /* FPU base = avgWeight * utilityAvg + (1 - avgWeight) * own evaluation, avgWeight =
                                             min(1, visitedPolicyMass ^ Pow) (searchexplorehelpers.cpp:279-282) */
  double rootDirichletNoiseTotalConcentration; /* 10.83 */
  double rootDirichletNoiseWeight;             /* 0.25 */
  double rootPolicyTemperature;           /* root policy ^ (1/T), T interpolated from ...Early by 0.5 ^ (turn / halflife * 19 / sqrt(area)) */
  double rootPolicyTemperatureEarly;      /* (searchhelpers.cpp:143-175, 463-467); 0 or 1 = off */
  double chosenMoveTemperatureHalflife;   /* 19 */
  double fpuParentWeightByVisitedPolicyPow;    /* 2.0 in selfplay1.cfg:185 */
  double rootDesiredPerChildVisitsCoeff;  /* a root child with weight < sqrt(prior * totalChildWeight * coeff) is searched first (2 in selfplay1.cfg:147) */
  double valueWeightExponent;             /* SearchParams::valueWeightExponent (0.5 in SearchParams() and selfplay1.cfg:179): children
                                             whose utility is implausibly low next to their siblings count less in the parent */
  /* Move choice under the reference's temperature schedule (searchresults.cpp:287-298, searchhelpers.cpp:12-49): if either
   * temperature is > 0 it replaces temperaturePlies -- edge visits, minus min(subtract, max/64), zero below min(prune, max/64),
   * raised to 1/T, T going from ...Early to chosenMoveTemperature with chosenMoveTemperatureHalflife (selfplay1.cfg:137-141). */
  double chosenMoveTemperature, chosenMoveTemperatureEarly, chosenMoveSubtract, chosenMovePrune;
  int32_t noPipeline;           /* With KC_SEARCH_PIPELINE=1 in the environment, the bf16 net and at least two trunk work items per SM pair the
                                   games are searched as two half batches on two streams, so that one half's select / expand kernels run under
                                   the other half's trunk kernel; 2 here selects that pipeline too, 1 forces one batch.  One batch is the default (it measured 2.6 % faster
                                   once the search kernels had been shortened); results are identical either way. */
  int32_t nnRandomize;          /* NNEvaluator's nnRandomize (cpp/neuralnet/nneval.cpp:515-524): every leaf is evaluated under one of the 8
                                   symmetries (inputs symmetrised, outputs mapped back).  The symmetry is drawn from the position's sit-hash
                                   and the seed, not from a shared stream, so a position gets the same one whichever game reaches it.  With a
                                   KC_FLAG_SYM_PERMUTE_DIRS handle it is a true symmetry of the game; without, the reference backends' spatial copy. */
  /* The rest of cpp/configs/training/selfplay1.cfg:144-185 (graph mode; zero = off). */
  int32_t useLcbForSelection;   /* selfplay1.cfg:151: the move choice raises the child with the best lower confidence bound above every child it
                                   beats (getPlaySelectionValues, cpp/search/searchresults.cpp:188-231; getSelfUtilityLCBAndRadius,
                                   searchhelpers.cpp:469-522, from utilitySqAvg / weightSqSum kept per node) */
  int32_t useNonBuggyLcb;       /* selfplay1.cfg:182; 0 = the historical form that never promotes the first-created child */
  double lcbStdevs;             /* 5.0 in selfplay1.cfg:152 */
  double minVisitPropForLCB;    /* 0.15 in selfplay1.cfg:153 */
  int32_t rootNumSymmetriesToSample;   /* selfplay1.cfg:149 (4): every search starts by evaluating its root under that many distinct symmetries and
                                          averaging the outputs (searchnnhelpers.cpp:67-83, 133-174; NNOutput's averaging constructor) */
  int32_t useNoisePruning;      /* SearchParams::useNoisePruning (pruneNoiseWeight, searchupdatehelpers.cpp:422-470): in creation order, a child below the
                                   weighted average of its elder siblings keeps at most twice its raw-policy share of their weight; off in self-play
                                   (setup.cpp:525), on in the GTP / analysis defaults */
  int32_t useUncertainty;       /* SearchParams::useUncertainty: a node's own evaluation weighs uncertaintyCoeff / (shorttermWinlossError ^ exponent +
                                   coeff / maxWeight) instead of 1 (computeWeightFromNNOutput, searchupdatehelpers.cpp:91-113); off in self-play */
  int32_t pad4_;
  double uncertaintyCoeff, uncertaintyExponent, uncertaintyMaxWeight;   /* 0.25 / 1.0 / 8.0 in the GTP defaults (setup.cpp:545-560) */
  double noisePruneUtilityScale, noisePruningCap;                        /* 0.15 / 1e50 (searchparams.cpp:27-28) */
  int32_t nnCacheSizePowerOfTwo;   /* NNCacheTable for the device search (nneval.cpp:874-932; selfplay1.cfg:121 uses 21): 2^n direct-mapped entries of
                                      post-processed outputs shared by all games of the search, keyed by the whole identity of the net's inputs, so
                                      a hit returns bit for bit what the evaluation would have returned -- fewer rows, same search.  0 = off. */
  int32_t pad5_;
} kc_search_params;
/* With any of chosenMoveTemperature[Early] or useLcbForSelection set in graph mode the move is chosen from the full
 * Search::getPlaySelectionValues (child weights; children other than the most stably explored one cut down to the weight its final
 * explore-selection value asks for, getReducedPlaySelectionWeight, rounded up; then LCB), and at a noised root the children that choice
 * would prune count nothing in the root's statistics either (recomputeNodeStats, searchupdatehelpers.cpp:196-206). */
typedef struct {
  uint64_t visits, netEvals, terminalVisits, movesPlayed, gamesFinished, blackWins, whiteWins, draws;
  uint64_t batchRows;           /* rows sent through the evaluator */
  uint64_t transpositionHits;   /* graph search: new edges that found their position already in the graph (no evaluation) */
  uint64_t catchUpVisits;       /* graph search: visits absorbed by an edge lagging behind its node (maybeCatchUpEdgeVisits) */
  uint64_t nnCacheHits;         /* leaves whose evaluation came from the search's NN cache (nnCacheSizePowerOfTwo > 0); not counted in netEvals */
} kc_search_stats;
/* handle == NULL selects the deterministic integer-hash evaluator (exact fp32 policy/value derived from the sit-hash),
 * which exists so that the search logic can be compared bit for bit with the CPU oracle; with a handle the leaves go
 * through kc_games_eval + kc_games_postprocess' kernels.  numGames <= the handle's maxBatch. */
int kc_search_create(kc_ctx* ctx, kc_handle* handleOrNull, int numGames, int xSize, int ySize, int winLen,
                     const kc_search_params* params, kc_search** out);
int kc_search_destroy(kc_search* s);
/* The games being played (owned by the search): kc_games_load / kc_games_step work on it between searches. */
kc_games* kc_search_games(kc_search* s);
int kc_search_reset(kc_search* s, uint64_t seed, uint64_t firstGameId);
/* One search (maxVisits iterations) from the current positions without playing a move. */
int kc_search_run_visits(kc_search* s);
/* Root statistics after kc_search_run_visits: rootVisits [G], rootUtilitySum [G] (white-positive), and per policy
 * index [G][4*H*W]: edgeVisits, edgeUtilitySum, the root's policy (-1 illegal), creation order (255 = no child). */
int kc_search_read_root(kc_search* s, int32_t* rootVisits, double* rootUtilitySum, int32_t* edgeVisits,
                        double* edgeUtilitySum, float* policy, uint8_t* order);
/* `moves` times: [refill] -> search -> choose and play the move.  chosenLast [G] receives the last moves played
 * (policy index, -1 none); *msTotal the device time of the whole call. */
int kc_search_play(kc_search* s, int moves, int16_t* chosenLast, kc_search_stats* statsAccum, float* msTotal);
/* Training rows (TrainingWriteBuffers::addRow, cpp/dataio/trainingwrite.cpp:316-566).  After
 * kc_search_enable_training_rows every move played by kc_search_play is recorded and, when its game ends, written out
 * as one row of the reference's npz arrays (rows of one game are consecutive, games in completion order):
 *   binaryInputNCHWPacked [N][15][ceil(H*W/8)] u8   fillRowV1 planes, bits big-endian in each byte (packBits :218-233)
 *   globalInputNC         [N][1] f32                win_len
 *   policyTargetsNCMove   [N][2][4*H*W] i16         visits of this turn's search; of the next turn's (all ones on the last turn)
 *   globalTargetsNC       [N][64] f32               the reference's literal indices: 0..9 win/loss td-targets from the player to move's
 *        view for nowFactor 0, 1/(1+A*0.176), 1/(1+A*0.056), 1/(1+A*0.016), 1 (fillValueTDTargets :286-314; the turn targets are the
 *        search's root estimate (1 +- utility)/2, the last one the result, draw = 0.5/0.5), 25 row weight 1, 26 policy weight 1,
 *        27 ownership weight 1, 28 next-policy weight, 33 future-position weight 1, 36..40 history masks 1 (the reference randomises
 *        them), 41..46 game hash in 22/22/20-bit chunks (splitmix64 of the game id), 51 turn index, 60 root visits, 63 version 1;
 *        every other entry 0 (Go-only or not produced by this search: score, lead, surprise / entropy statistics, net ages)
 *   valueTargetsNCHW      [N][5][H][W] i8           0 final stones (+1 own, -1 opponent's), 1 zero, 2 / 3 the stones 2 / 6 plies later
 *        (clipped to the end), 4 the longest same-colour run through each final stone (recordMaxConsecutives, stones only)
 * maxRows bounds the device buffer; rows of games that do not fit are counted in *numDropped. */
int kc_search_enable_training_rows(kc_search* s, int maxRows);
int kc_search_read_training_rows(kc_search* s, int* numRows, int* numDropped, uint8_t* binaryInputNCHWPacked, float* globalInputNC,
                                 int16_t* policyTargetsNCMove, float* globalTargetsNC, int8_t* valueTargetsNCHW, int clear);
/* Graph-mode searches (useGraphSearch or subtreeValueBiasFactor != 0): digest [G] = order-independent hash over every node of
 * each game's graph (visits, weightSum and utilityAvg bit patterns, edges with their visit counts and creation order), the
 * quantity the oracle's ko_search_run_graph reports, for whole-graph comparisons. */
/* The rows as the reference's training-data file (TrainingWriteBuffers::writeToZipFile, cpp/dataio/trainingwrite.cpp:566-587): a zip
 * archive of five deflated numpy arrays with NumpyBuffer's 256-byte headers (cpp/dataio/numpywrite.cpp:110-222) under the member names
 * binaryInputNCHWPacked [N,15,ceil(HW/8)] u1, globalInputNC [N,1] f4, policyTargetsNCMove [N,2,4HW] i2, globalTargetsNC [N,64] f4,
 * valueTargetsNCHW [N,5,H,W] i1 -- what python/shuffle.py and numpy.load read.  Host-only; written to path + ".tmp" and renamed. */
int kc_training_write_npz(const char* path, int numRows, int xSize, int ySize, const uint8_t* binaryInputNCHWPacked, const float* globalInputNC,
                          const int16_t* policyTargetsNCMove, const float* globalTargetsNC, const int8_t* valueTargetsNCHW);
/* The play-selection values [G][4*H*W] the last kc_search_play chose its last move from (0 for moves without a child). */
int kc_search_read_play_selection(kc_search* s, double* playSelection);
int kc_search_tree_digest(kc_search* s, uint64_t* digest);
int64_t kc_search_launch_count(const kc_search* s);

/* ---------------------------------------------------------------------------------------------
 * Self-play over several GPUs of one box from one process: one search pool (host thread, context, weights, compute handle,
 * kc_search of gamesPerDevice games) per device and one data-writer thread per pool -- the reference's one NN server thread +
 * ComputeHandle per GPU (cpp/program/setup.cpp:167-229, gpuIdxByServerThread) under the game loops and the data-writer thread of
 * cpp/command/selfplay.cpp:226-260,390-392.  Games never cross devices; pool i plays the game ids firstGameId + i * 2^40 + ...
 * The timed part plays `moves` moves per game lane in chunks of movesPerChunk; after every chunk the rows of the games that ended
 * are read back and written as outputDir/poolNN_CCCCCC.npz (kc_training_write_npz) while the next chunk is searched.  At the end
 * the pools' counters are summed on devices[0] with ONE ncclReduce (NCCL loaded with dlopen("libnccl.so.2"); with one device or
 * noNccl != 0 they are summed on the host) -- the only exchange of the whole job.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int32_t numDevices;
  const int32_t* devices;     /* CUDA device indices */
  int32_t gamesPerDevice, xSize, ySize, winLen;
  int32_t moves;              /* moves per game lane in the timed part */
  int32_t movesPerChunk;      /* 0 = one chunk */
  int32_t warmupMoves;        /* untimed moves first (with reuseTree they build the trees the timed moves re-use); their rows are discarded */
  int32_t staggerPlies;       /* lane g starts after (g mod staggerPlies) random-legal plies: the mix of a running self-play; 0 = empty boards */
  int32_t maxRowsPerChunk;    /* device row buffer per pool; 0 = no training rows */
  int32_t noNccl;             /* 1: sum the counters on the host */
  uint32_t handleFlags;       /* kc_handle_create flags */
  uint64_t seed, firstGameId;
  const char* outputDir;      /* NULL: rows are read back, no file is written */
} kc_selfplay_config;
typedef struct {
  double wallSeconds;         /* the timed part on the host clock: from all pools ready to the last pool's last file closed */
  double deviceMsMax;         /* max over pools of the device time of its kc_search_play calls */
  uint64_t rowsWritten, rowsDropped, filesWritten, bytesWritten, kernelLaunches;
  int32_t reducedWithNccl;
} kc_selfplay_report;
int kc_selfplay_run(const kc_selfplay_config* cfg, const kc_model_desc* desc, const kc_search_params* params, kc_search_stats* total,
                    kc_selfplay_report* report);


/* ---------------------------------------------------------------------------------------------
 * Evaluator front end.  Replaces NNEvaluator::evaluate / serve and NNCacheTable
 * (cpp/neuralnet/nneval.h:17-42,80-249; nneval.cpp:341-586 server loop, :588-815 evaluate, :820-932 cache) for callers
 * that search on the CPU with many threads (the reference's own Search):
 *  - clients hand over a POSITION (stones, player to move, last five moves), not filled rows: the 100 isLegal calls of
 *    fillRowV1 (nninputs.cpp:635-647) and the 100 of the post-processing (nneval.cpp:712-715) run on the device
 *    (rules kernel -> bf16 trunk tiles -> net -> masked softmax), 40 bytes per row cross PCIe instead of 1.5 KB;
 *  - the submit path takes no lock: a client claims a row with one atomic add on a ticket counter and packs its position
 *    straight into the page-locked staging of the batch that ticket belongs to (the reference serialises every client on
 *    one bufferMutex, nneval.cpp:663-676, and allocates an NNOutput per row, :505-515); a server thread closes a batch by
 *    moving the ticket counter to the next batch boundary, exactly when the reference's would (as soon as it is free and
 *    at least one row waits);
 *  - the cache keeps results inline in one flat table (no shared_ptr per entry) behind a striped mutex pool, keyed by
 *    NNInputs::getHash (nninputs.cpp:463-502) ^ mix(last five moves + last direction): SURVEY.md ledger 8.1-E -- the
 *    literal hash alone collides between positions whose stones agree and whose forced lines differ.  The literal
 *    hash is what kc_eval_output.nnHash reports.
 * kc_evaluator_evaluate is thread-safe and blocks until the result is there (NNEvaluator::evaluate's contract).
 * ------------------------------------------------------------------------------------------- */
typedef struct kc_evaluator kc_evaluator;
#define KC_SYMMETRY_NOTSPECIFIED (-1) /* NNInputs::SYMMETRY_NOTSPECIFIED: the evaluator chooses (doRandomize / defaultSymmetry) */
typedef struct {
  int nnXLen, nnYLen, winLen;  /* exact board size (requireExactNNLen) */
  int maxBatch;                /* rows per batch = NNEvaluator's maxBatchSize */
  int maxConcurrentEvals;      /* staging ring = maxConcurrentEvals / maxBatch + 3 (nneval.cpp:128-136) + numServerThreads batches, rounded up to a power of two */
  int numServerThreads;        /* each owns a compute handle on the context's device (nneval.cpp:341-362) */
  int cacheSizePowerOfTwo;     /* < 0: no cache (nneval.cpp:139-140) */
  int mutexPoolSizePowerOfTwo;
  int doRandomize;             /* rows without a symmetry get one drawn from (randSeed, cache key): a position is always evaluated the same way */
  int defaultSymmetry;         /* used when doRandomize == 0 */
  uint64_t randSeed;
  float policyTemperature;     /* MiscNNInputParams::nnPolicyTemperature, folded into the hash as nninputs.cpp:485-492 does */
  unsigned handleFlags;        /* KC_FLAG_FP32_CHECK | KC_FLAG_SYM_PERMUTE_DIRS for the server threads' handles */
} kc_evaluator_config;
/* A position as kc_games_load takes it: stones [H*W] (0 empty, 1 black, 2 white), nextPla 1 / 2, moves [5][2] = last five
 * (policy index, player) pairs oldest first, -1 = none (NULL = no history), numTurns.  It must have a legal move
 * (nneval.cpp:730 asserts the same). */
typedef struct {
  const int8_t* stones;
  const int16_t* moves;
  int32_t numTurns;
  int8_t nextPla;
} kc_eval_position;
/* NNOutput after NNEvaluator::evaluate's post-processing (nninputs.h:75-118, nneval.cpp:702-815): policyProbs [4*H*W]
 * (caller's buffer; illegal = -1), white's win / loss probabilities, varTimeLeft, shorttermWinlossError, whiteOwnerMap
 * [H*W] (caller's buffer or NULL; tanh, white's perspective). */
typedef struct {
  float* policyProbs;
  float* whiteOwnerMap;
  float whiteWinProb, whiteLossProb, varTimeLeft, shorttermWinlossError;
  uint64_t nnHash[2];
  int32_t symmetry;   /* the symmetry the row was evaluated under (of the evaluation that filled the cache, on a hit) */
  int32_t cacheHit;
} kc_eval_output;
typedef struct {
  uint64_t rowsProcessed, batchesProcessed;  /* NNEvaluator::numRowsProcessed / numBatchesProcessed */
  uint64_t cacheHits, cacheMisses, ownerMapUpgrades;
  uint64_t backpressureWaits;                /* submits that found the staging ring full (more than maxConcurrentEvals in flight) */
} kc_evaluator_stats;
/* One closed batch, as a server thread hands it to its backend.  Inputs are the packed device format of the games
 * kernels (csrc/games_device.cuh: bitboards with row stride W+1, pos_hash pair, misc = last five moves / direction /
 * numTurns / next player); kc_eval_unpack_position turns a row back into a kc_eval_position.  The backend fills
 * policyProbs [n][4*H*W], whiteWinLoss [n][2], miscOut [n][2] (all post-processed as nneval.cpp:702-801 does) and,
 * when wantOwnership != 0, ownership [n][H*W] = the net's raw ownership output (player to move's perspective). */
typedef struct {
  int n, wantOwnership;
  float policyTemperature;
  const uint64_t *black, *white, *hash0, *hash1, *misc;
  const int8_t* symmetry;
  float *policyProbs, *whiteWinLoss, *miscOut, *ownership;
  /* boards beyond 7x7 (up to the reference's 10x10, board.h:120): bits 64.. of the 128-bit bitboards, and misc in the wide format of
   * csrc/games_big.cuh (9-bit history entries, last direction at bit 45) -- kc_eval_unpack_position_wide; null otherwise */
  const uint64_t *blackHi, *whiteHi;
} kc_eval_batch;
typedef int (*kc_eval_backend_fn)(void* user, int serverThread, const kc_eval_batch* batch);

int kc_evaluator_create(kc_ctx* ctx, const kc_model* model, const kc_evaluator_config* cfg, kc_evaluator** out);
/* One evaluator over several GPUs, the reference's gpuIdxByServerThread (nneval.h:100, nneval.cpp:341-362; cpp/program/setup.cpp:190-229):
 * server thread i works on ctxs[i] with models[i] (a copy of the weights created on that context); numServers == cfg->numServerThreads.
 * The staging ring and the cache are shared, so the clients do not know which GPU serves them. */
int kc_evaluator_create_multi(int numServers, kc_ctx* const* ctxs, const kc_model* const* models, const kc_evaluator_config* cfg,
                              kc_evaluator** out);
/* The same front end over a caller-supplied batch function instead of the device (needs no GPU): the host-logic tests
 * drive it with the CPU oracle; a non-zero return fails every row of the batch with `kc_last_error` = "backend failed". */
int kc_evaluator_create_custom(const kc_evaluator_config* cfg, kc_eval_backend_fn fn, void* user, kc_evaluator** out);
/* Joins the server threads; no kc_evaluator_evaluate may be in flight (NNEvaluator::~NNEvaluator / killServerThreads). */
int kc_evaluator_destroy(kc_evaluator* ev);
int kc_evaluator_evaluate(kc_evaluator* ev, const kc_eval_position* pos, int symmetry, int skipCache, int includeOwnerMap,
                          kc_eval_output* out);
/* n positions from one client thread (a batched CPU search): rows are submitted without waiting for one another and
 * collected afterwards; results are exactly those of n kc_evaluator_evaluate calls. */
int kc_evaluator_evaluate_many(kc_evaluator* ev, int n, const kc_eval_position* pos, const int8_t* symmetryOrNull, int skipCache,
                               int includeOwnerMap, kc_eval_output* out);
int kc_evaluator_clear_cache(kc_evaluator* ev);  /* NNEvaluator::clearCache */
int kc_evaluator_get_stats(const kc_evaluator* ev, kc_evaluator_stats* out);
int kc_evaluator_clear_stats(kc_evaluator* ev);  /* NNEvaluator::clearStats */
/* Host helpers: the literal NNInputs::getHash of a position (policy temperature folded in when != 1) and the cache key
 * the evaluator uses for it; the inverse of the packing for custom backends. */
int kc_eval_position_hash(int xSize, int ySize, const kc_eval_position* pos, float policyTemperature, uint64_t nnHash[2], uint64_t cacheKey[2]);
int kc_eval_unpack_position(int xSize, int ySize, uint64_t black, uint64_t white, uint64_t misc, int8_t* stones, int8_t* nextPla,
                            int16_t* movesCellPla, int32_t* numTurns, int32_t* lastDir);
/* the same for a row of a board beyond 7x7 (kc_eval_batch::blackHi / whiteHi non-null) */
int kc_eval_unpack_position_wide(int xSize, int ySize, uint64_t blackLo, uint64_t blackHi, uint64_t whiteLo, uint64_t whiteHi, uint64_t misc,
                                 int8_t* stones, int8_t* nextPla, int16_t* movesCellPla, int32_t* numTurns, int32_t* lastDir);

#ifdef __cplusplus
}
#endif
#endif
