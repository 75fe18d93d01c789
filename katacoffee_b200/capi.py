"""ctypes bindings for include/katacoffee_b200.h (one prototype per exported symbol)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkatacoffee_b200.so")

c_float_p = C.POINTER(C.c_float)


class ConvDesc(C.Structure):
    _fields_ = [("convYSize", C.c_int32), ("convXSize", C.c_int32), ("inChannels", C.c_int32),
                ("outChannels", C.c_int32), ("weights", c_float_p)]


class BNDesc(C.Structure):
    _fields_ = [("numChannels", C.c_int32), ("epsilon", C.c_float), ("hasScale", C.c_int32),
                ("hasBias", C.c_int32), ("mean", c_float_p), ("variance", c_float_p),
                ("scale", c_float_p), ("bias", c_float_p)]


class MatMulDesc(C.Structure):
    _fields_ = [("inChannels", C.c_int32), ("outChannels", C.c_int32), ("weights", c_float_p)]


class MatBiasDesc(C.Structure):
    _fields_ = [("numChannels", C.c_int32), ("pad_", C.c_int32), ("weights", c_float_p)]


class BlockDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("preActivation", C.c_int32), ("gpoolActivation", C.c_int32),
                ("midActivation", C.c_int32), ("preBN", BNDesc), ("regularConv", ConvDesc),
                ("gpoolConv", ConvDesc), ("gpoolBN", BNDesc), ("gpoolToBiasMul", MatMulDesc),
                ("midBN", BNDesc), ("finalConv", ConvDesc)]


class ModelDesc(C.Structure):
    _fields_ = [("version", C.c_int32), ("numInputChannels", C.c_int32),
                ("numInputGlobalChannels", C.c_int32), ("numBlocks", C.c_int32),
                ("trunkNumChannels", C.c_int32), ("midNumChannels", C.c_int32),
                ("regularNumChannels", C.c_int32), ("gpoolNumChannels", C.c_int32),
                ("trunkTipActivation", C.c_int32), ("g1Activation", C.c_int32),
                ("p1Activation", C.c_int32), ("v1Activation", C.c_int32), ("v2Activation", C.c_int32),
                ("pad_", C.c_int32),
                ("initialConv", ConvDesc), ("initialMatMul", MatMulDesc),
                ("blocks", C.POINTER(BlockDesc)), ("trunkTipBN", BNDesc),
                ("p1Conv", ConvDesc), ("g1Conv", ConvDesc), ("g1BN", BNDesc),
                ("gpoolToBiasMul", MatMulDesc), ("p1BN", BNDesc), ("p2Conv", ConvDesc),
                ("v1Conv", ConvDesc), ("v1BN", BNDesc), ("v2Mul", MatMulDesc), ("v2Bias", MatBiasDesc),
                ("v3Mul", MatMulDesc), ("v3Bias", MatBiasDesc), ("sv3Mul", MatMulDesc),
                ("sv3Bias", MatBiasDesc), ("vOwnershipConv", ConvDesc)]


class Stats(C.Structure):
    _fields_ = [("steps", C.c_uint64), ("evals", C.c_uint64), ("gamesFinished", C.c_uint64),
                ("blackWins", C.c_uint64), ("whiteWins", C.c_uint64), ("draws", C.c_uint64),
                ("checksum", C.c_uint64)]


class SearchParams(C.Structure):
    _fields_ = [("maxVisits", C.c_int32), ("temperaturePlies", C.c_int32), ("autoRefill", C.c_int32), ("noCompaction", C.c_int32), ("reuseTree", C.c_int32), ("useGraphSearch", C.c_int32),
                ("cpuctExploration", C.c_double), ("fpuReductionMax", C.c_double), ("rootFpuReductionMax", C.c_double),
                ("subtreeValueBiasFactor", C.c_double), ("subtreeValueBiasWeightExponent", C.c_double), ("subtreeValueBiasFreeProp", C.c_double),
                ("rootNoiseEnabled", C.c_int32), ("fpuParentWeightByVisitedPolicy", C.c_int32),
                ("rootDirichletNoiseTotalConcentration", C.c_double), ("rootDirichletNoiseWeight", C.c_double),
                ("rootPolicyTemperature", C.c_double), ("rootPolicyTemperatureEarly", C.c_double), ("chosenMoveTemperatureHalflife", C.c_double),
                ("fpuParentWeightByVisitedPolicyPow", C.c_double), ("rootDesiredPerChildVisitsCoeff", C.c_double), ("valueWeightExponent", C.c_double),
                ("chosenMoveTemperature", C.c_double), ("chosenMoveTemperatureEarly", C.c_double), ("chosenMoveSubtract", C.c_double), ("chosenMovePrune", C.c_double),
                ("noPipeline", C.c_int32), ("nnRandomize", C.c_int32),
                ("useLcbForSelection", C.c_int32), ("useNonBuggyLcb", C.c_int32), ("lcbStdevs", C.c_double), ("minVisitPropForLCB", C.c_double),
                ("rootNumSymmetriesToSample", C.c_int32), ("useNoisePruning", C.c_int32), ("useUncertainty", C.c_int32), ("pad4_", C.c_int32),
                ("uncertaintyCoeff", C.c_double), ("uncertaintyExponent", C.c_double), ("uncertaintyMaxWeight", C.c_double),
                ("noisePruneUtilityScale", C.c_double), ("noisePruningCap", C.c_double),
                ("nnCacheSizePowerOfTwo", C.c_int32), ("pad5_", C.c_int32)]


class SearchStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("visits", "netEvals", "terminalVisits", "movesPlayed", "gamesFinished", "blackWins",
                                          "whiteWins", "draws", "batchRows", "transpositionHits", "catchUpVisits", "nnCacheHits")]


class SelfplayConfig(C.Structure):
    _fields_ = [("numDevices", C.c_int32), ("devices", C.POINTER(C.c_int32)), ("gamesPerDevice", C.c_int32), ("xSize", C.c_int32), ("ySize", C.c_int32),
                ("winLen", C.c_int32), ("moves", C.c_int32), ("movesPerChunk", C.c_int32), ("warmupMoves", C.c_int32), ("staggerPlies", C.c_int32),
                ("maxRowsPerChunk", C.c_int32), ("noNccl", C.c_int32), ("handleFlags", C.c_uint32), ("seed", C.c_uint64), ("firstGameId", C.c_uint64),
                ("outputDir", C.c_char_p)]


class SelfplayReport(C.Structure):
    _fields_ = [("wallSeconds", C.c_double), ("deviceMsMax", C.c_double), ("rowsWritten", C.c_uint64), ("rowsDropped", C.c_uint64),
                ("filesWritten", C.c_uint64), ("bytesWritten", C.c_uint64), ("kernelLaunches", C.c_uint64), ("reducedWithNccl", C.c_int32)]


class EvaluatorConfig(C.Structure):
    _fields_ = [("nnXLen", C.c_int32), ("nnYLen", C.c_int32), ("winLen", C.c_int32), ("maxBatch", C.c_int32),
                ("maxConcurrentEvals", C.c_int32), ("numServerThreads", C.c_int32), ("cacheSizePowerOfTwo", C.c_int32),
                ("mutexPoolSizePowerOfTwo", C.c_int32), ("doRandomize", C.c_int32), ("defaultSymmetry", C.c_int32),
                ("randSeed", C.c_uint64), ("policyTemperature", C.c_float), ("handleFlags", C.c_uint32)]


class EvalPosition(C.Structure):
    _fields_ = [("stones", C.POINTER(C.c_int8)), ("moves", C.POINTER(C.c_int16)), ("numTurns", C.c_int32), ("nextPla", C.c_int8)]


class EvalOutput(C.Structure):
    _fields_ = [("policyProbs", c_float_p), ("whiteOwnerMap", c_float_p), ("whiteWinProb", C.c_float), ("whiteLossProb", C.c_float),
                ("varTimeLeft", C.c_float), ("shorttermWinlossError", C.c_float), ("nnHash", C.c_uint64 * 2), ("symmetry", C.c_int32),
                ("cacheHit", C.c_int32)]


class EvaluatorStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rowsProcessed", "batchesProcessed", "cacheHits", "cacheMisses", "ownerMapUpgrades",
                                          "backpressureWaits")]


class EvalBatch(C.Structure):
    _fields_ = [("n", C.c_int32), ("wantOwnership", C.c_int32), ("policyTemperature", C.c_float),
                ("black", C.POINTER(C.c_uint64)), ("white", C.POINTER(C.c_uint64)), ("hash0", C.POINTER(C.c_uint64)),
                ("hash1", C.POINTER(C.c_uint64)), ("misc", C.POINTER(C.c_uint64)), ("symmetry", C.POINTER(C.c_int8)),
                ("policyProbs", c_float_p), ("whiteWinLoss", c_float_p), ("miscOut", c_float_p), ("ownership", c_float_p),
                ("blackHi", C.POINTER(C.c_uint64)), ("whiteHi", C.POINTER(C.c_uint64))]


EVAL_BACKEND_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(EvalBatch))

FLAG_FP32_CHECK = 1
FLAG_INPUTS_NHWC = 2
FLAG_SYM_PERMUTE_DIRS = 4
FLAG_OPERANDS_BF16 = 8
FLAG_MASKED_BOARDS = 16

vp = C.c_void_p
# name -> (restype, argtypes); must list every symbol include/katacoffee_b200.h declares
PROTOTYPES = {
    "kc_last_error": (C.c_char_p, []),
    "kc_abi_version": (C.c_int, []),
    "kc_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "kc_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
    "kc_ctx_destroy": (C.c_int, [vp]),
    "kc_zobrist_tables": (C.c_int, [vp, vp, vp, vp]),
    "kc_model_create": (C.c_int, [vp, C.POINTER(ModelDesc), C.POINTER(vp)]),
    "kc_model_destroy": (C.c_int, [vp]),
    "kc_handle_create": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_uint, C.POINTER(vp)]),
    "kc_handle_destroy": (C.c_int, [vp]),
    "kc_handle_uses_bf16": (C.c_int, [vp]),
    "kc_handle_operand_format": (C.c_int, [vp]),
    "kc_forward": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]),
    "kc_forward_rows": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp, vp]),
    "kc_handle_read_outputs": (C.c_int, [vp, C.c_int, vp, vp, vp, vp]),
    "kc_handle_launch_count": (C.c_int64, [vp]),
    "kc_handle_trunk_time": (C.c_int, [vp, C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "kc_handle_trunk_probe": (C.c_int, [vp, vp]),
    "kc_selftest_umma": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "kc_test_conv": (C.c_int, [vp, C.POINTER(ConvDesc), C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]),
    "kc_test_batchnorm": (C.c_int, [vp, C.POINTER(BNDesc), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "kc_test_resblock": (C.c_int, [vp, C.POINTER(BlockDesc), C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "kc_games_create": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)]),
    "kc_games_destroy": (C.c_int, [vp]),
    "kc_games_reset": (C.c_int, [vp, C.c_uint64, C.c_uint64, C.c_int]),
    "kc_games_load": (C.c_int, [vp, C.c_int, C.c_int, vp, vp, vp, vp]),
    "kc_games_step": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "kc_games_features": (C.c_int, [vp, C.c_int, vp, vp, vp]),
    "kc_games_eval": (C.c_int, [vp, vp, vp]),
    "kc_games_postprocess": (C.c_int, [vp, vp, C.c_float, vp, vp, vp, vp]),
    "kc_games_run": (C.c_int, [vp, vp, C.c_int, C.POINTER(Stats)]),
    "kc_games_run_timed": (C.c_int, [vp, vp, C.c_int, C.c_size_t, C.POINTER(Stats), C.POINTER(C.c_float)]),
    "kc_games_launch_count": (C.c_int64, [vp]),
    "kc_games_last_kernel_ms": (C.c_float, [vp]),
    "kc_modelfile_load": (C.c_int, [C.c_char_p, C.c_char_p, C.POINTER(vp)]),
    "kc_modelfile_free": (C.c_int, [vp]),
    "kc_modelfile_desc": (vp, [vp]),
    "kc_modelfile_name": (C.c_char_p, [vp]),
    "kc_modelfile_sha256": (C.c_char_p, [vp]),
    "kc_modelfile_write": (C.c_int, [vp, C.c_char_p, C.c_char_p]),
    "kc_sgf_write": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, vp, C.c_int, vp, vp, C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "kc_sgf_parse": (C.c_int, [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), vp, C.c_int, vp, vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "kc_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(vp)]),
    "kc_host_free": (C.c_int, [vp]),
    "kc_games_read_run_outputs": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "kc_games_read_run_ply": (C.c_int, [vp, C.c_int, vp, vp, vp, vp, vp]),
    "kc_search_create": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(SearchParams), C.POINTER(vp)]),
    "kc_search_destroy": (C.c_int, [vp]),
    "kc_search_games": (vp, [vp]),
    "kc_search_reset": (C.c_int, [vp, C.c_uint64, C.c_uint64]),
    "kc_search_run_visits": (C.c_int, [vp]),
    "kc_search_read_root": (C.c_int, [vp, vp, vp, vp, vp, vp, vp]),
    "kc_search_play": (C.c_int, [vp, C.c_int, vp, C.POINTER(SearchStats), C.POINTER(C.c_float)]),
    "kc_search_enable_training_rows": (C.c_int, [vp, C.c_int]),
    "kc_search_read_training_rows": (C.c_int, [vp, C.POINTER(C.c_int), C.POINTER(C.c_int), vp, vp, vp, vp, vp, C.c_int]),
    "kc_training_write_npz": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp]),
    "kc_selfplay_run": (C.c_int, [vp, vp, vp, vp, vp]),
    "kc_search_read_play_selection": (C.c_int, [vp, vp]),
    "kc_search_tree_digest": (C.c_int, [vp, vp]),
    "kc_search_launch_count": (C.c_int64, [vp]),
    "kc_evaluator_create": (C.c_int, [vp, vp, C.POINTER(EvaluatorConfig), C.POINTER(vp)]),
    "kc_evaluator_create_multi": (C.c_int, [C.c_int, C.POINTER(vp), C.POINTER(vp), C.POINTER(EvaluatorConfig), C.POINTER(vp)]),
    "kc_evaluator_create_custom": (C.c_int, [C.POINTER(EvaluatorConfig), EVAL_BACKEND_FN, vp, C.POINTER(vp)]),
    "kc_evaluator_destroy": (C.c_int, [vp]),
    "kc_evaluator_evaluate": (C.c_int, [vp, C.POINTER(EvalPosition), C.c_int, C.c_int, C.c_int, C.POINTER(EvalOutput)]),
    "kc_evaluator_evaluate_many": (C.c_int, [vp, C.c_int, C.POINTER(EvalPosition), vp, C.c_int, C.c_int, C.POINTER(EvalOutput)]),
    "kc_evaluator_clear_cache": (C.c_int, [vp]),
    "kc_evaluator_get_stats": (C.c_int, [vp, C.POINTER(EvaluatorStats)]),
    "kc_evaluator_clear_stats": (C.c_int, [vp]),
    "kc_eval_position_hash": (C.c_int, [C.c_int, C.c_int, C.POINTER(EvalPosition), C.c_float, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "kc_eval_unpack_position": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, vp, C.POINTER(C.c_int8), vp,
                                          C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "kc_eval_unpack_position_wide": (C.c_int, [C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, vp, C.POINTER(C.c_int8), vp,
                                               C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
}

_lib = None


class KCError(RuntimeError):
    """Raised for any non-zero status from the C ABI (the analogue of the reference's StringError)."""


def lib():
    """Loads the CUDA library. Fails loudly if it was not built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KCError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(katacoffee_b200 has no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(status):
    if status != 0:
        raise KCError(lib().kc_last_error().decode("utf-8", "replace"))


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)
