"""Host-side mirror of the reference's backend interface for this path.

Function names and argument meaning follow cpp/neuralnet/nninterface.h (createComputeContext :52,
createComputeHandle :77, getOutput :112, testEvaluate* :127-169) so the parity tests read like the
reference's own tests/testnn.cpp; every call goes through the C ABI (capi) and raises KCError
(the StringError analogue) on failure.  `Games` wraps the batched rules/features entry points.
"""
import ctypes as C
import os

import numpy as np

from . import capi
from .capi import check, lib, ptr


class ComputeContext:
    def __init__(self, device=0):
        self._p = C.c_void_p()
        check(lib().kc_ctx_create(device, C.byref(self._p)))

    def close(self):
        if self._p:
            lib().kc_ctx_destroy(self._p)
            self._p = C.c_void_p()


def createComputeContext(gpuIdx=0):
    return ComputeContext(gpuIdx)


class LoadedModel:
    def __init__(self, ctx, model):
        """model: katacoffee_b200.modeldesc.Model (or anything with a .desc ModelDesc)."""
        self.ctx, self.model = ctx, model
        self._p = C.c_void_p()
        check(lib().kc_model_create(ctx._p, C.byref(model.desc), C.byref(self._p)))

    def close(self):
        if self._p:
            lib().kc_model_destroy(self._p)
            self._p = C.c_void_p()


class ModelFile:
    """A parsed model file (NeuralNet::loadModelFile's parsing half; cpp/neuralnet/desc.cpp).  No GPU needed.
    `.desc` is a capi.ModelDesc whose pointers stay valid while this object lives."""

    def __init__(self, path, expectedSha256=""):
        self._p = C.c_void_p()
        check(lib().kc_modelfile_load(os.fsencode(path), (expectedSha256 or "").encode(), C.byref(self._p)))
        self.desc = C.cast(lib().kc_modelfile_desc(self._p), C.POINTER(capi.ModelDesc)).contents
        self.name = lib().kc_modelfile_name(self._p).decode()
        self.sha256 = lib().kc_modelfile_sha256(self._p).decode()

    def close(self):
        if self._p:
            lib().kc_modelfile_free(self._p)
            self._p = C.c_void_p()

    def __del__(self):
        self.close()


def writeModelFile(model, path, name=None):
    """Writes a modeldesc.Model (or anything with a .desc) in the reference's model format (.txt/.bin[.gz])."""
    check(lib().kc_modelfile_write(C.byref(model.desc), (name or getattr(model, "name", "model")).encode(), os.fsencode(path)))


def loadModelFile(ctx, path, expectedSha256=""):
    """NeuralNet::loadModelFile (cpp/neuralnet/nninterface.h:42): parse the file and build the device model."""
    mf = ModelFile(path, expectedSha256)
    return LoadedModel(ctx, mf)


def writeSgf(xSize, ySize, winLen, moves, players=None, winner=-1, blackName="B200", whiteName="B200", initialStones=None):
    """Custom Coffee SGF text of a game (README.md:33-35): moves are policy indices, players default to black first."""
    mv = np.ascontiguousarray(moves, np.int16)
    pl = np.ascontiguousarray(players if players is not None else [1 + (i % 2) for i in range(len(mv))], np.int8)
    st = None if initialStones is None else np.ascontiguousarray(initialStones, np.int8).reshape(-1)
    cap = 256 + 8 * len(mv) + 4 * xSize * ySize + len(blackName) + len(whiteName)
    buf = C.create_string_buffer(cap)
    n = C.c_size_t()
    check(lib().kc_sgf_write(xSize, ySize, winLen, blackName.encode(), whiteName.encode(), None if st is None else ptr(st), len(mv),
                             ptr(mv) if len(mv) else None, ptr(pl) if len(mv) else None, winner, buf, cap, C.byref(n)))
    return buf.value.decode()


def parseSgf(text, maxMoves=400):
    """dict(xSize, ySize, winLen, initialStones [H, W], moves, players, winner) of the main line of a Coffee SGF."""
    x, y, k, n, w = C.c_int(), C.c_int(), C.c_int(), C.c_int(), C.c_int()
    stones = np.zeros(100, np.int8); mv = np.zeros(maxMoves, np.int16); pl = np.zeros(maxMoves, np.int8)
    check(lib().kc_sgf_parse(text.encode(), C.byref(x), C.byref(y), C.byref(k), ptr(stones), maxMoves, ptr(mv), ptr(pl), C.byref(n), C.byref(w)))
    m = min(n.value, maxMoves)
    return dict(xSize=x.value, ySize=y.value, winLen=k.value, initialStones=stones[:x.value * y.value].reshape(y.value, x.value).copy(),
                moves=mv[:m].copy(), players=pl[:m].copy(), winner=w.value, numMoves=n.value)


class ComputeHandle:
    def __init__(self, ctx, loadedModel, maxBatchSize, nnXLen, nnYLen, useFP32Check=False, inputsUseNHWC=False, playModeSymmetry=False,
                 operandsBF16=False, requireExactNNLen=True):
        self.ctx, self.loadedModel = ctx, loadedModel
        self.maxBatch, self.nnXLen, self.nnYLen = maxBatchSize, nnXLen, nnYLen
        self.inputsUseNHWC = inputsUseNHWC
        flags = ((capi.FLAG_FP32_CHECK if useFP32Check else 0) | (capi.FLAG_INPUTS_NHWC if inputsUseNHWC else 0) |
                 (capi.FLAG_SYM_PERMUTE_DIRS if playModeSymmetry else 0) | (capi.FLAG_OPERANDS_BF16 if operandsBF16 else 0) |
                 (0 if requireExactNNLen else capi.FLAG_MASKED_BOARDS))
        self._p = C.c_void_p()
        check(lib().kc_handle_create(ctx._p, loadedModel._p, maxBatchSize, nnXLen, nnYLen, flags, C.byref(self._p)))

    def isUsingBF16(self):
        """isUsingFP16 analogue (nninterface.h:88): true on the tensor-core path, whichever 16-bit operand format it runs."""
        return bool(lib().kc_handle_uses_bf16(self._p))

    def operandFormat(self):
        return ("fp16", "bf16", "fp32")[lib().kc_handle_operand_format(self._p)]

    def launchCount(self):
        return int(lib().kc_handle_launch_count(self._p))

    def trunkTime(self):
        s, n = C.c_float(), C.c_int()
        check(lib().kc_handle_trunk_time(self._p, C.byref(s), C.byref(n)))
        return s.value, n.value

    def readOutputs(self, n, ownership=True):
        hw = self.nnXLen * self.nnYLen
        policy = np.empty((n, 4 * hw), np.float32)
        value = np.empty((n, 2), np.float32)
        misc = np.empty((n, 2), np.float32)
        own = np.empty((n, hw), np.float32) if ownership else None
        check(lib().kc_handle_read_outputs(self._p, n, ptr(policy), ptr(value), ptr(misc), ptr(own)))
        return policy, value, misc, own

    def close(self):
        if self._p:
            lib().kc_handle_destroy(self._p)
            self._p = C.c_void_p()


def createComputeHandle(context, loadedModel, maxBatchSize, nnXLen, nnYLen, useFP32Check=False, inputsUseNHWC=False, playModeSymmetry=False,
                        operandsBF16=False, requireExactNNLen=True):
    """NeuralNet::createComputeHandle (nninterface.h:66-86).  requireExactNNLen=False: boards may be smaller than nnXLen x nnYLen
    (input channel 0 is the mask), on either path."""
    return ComputeHandle(context, loadedModel, maxBatchSize, nnXLen, nnYLen, useFP32Check, inputsUseNHWC, playModeSymmetry, operandsBF16, requireExactNNLen)


def getOutput(handle, rowSpatial, rowGlobal, symmetry=None, ownership=True, out=None):
    """NeuralNet::getOutput: rowSpatial [n, 15*H*W] fp32, rowGlobal [n, 1], symmetry [n] int8.
    Returns raw logits (policy [n,4HW], value [n,2], misc [n,2], ownership [n,HW])."""
    n = rowSpatial.shape[0]
    hw = handle.nnXLen * handle.nnYLen
    rowSpatial = np.ascontiguousarray(rowSpatial, np.float32).reshape(n, 15 * hw)
    rowGlobal = np.ascontiguousarray(rowGlobal, np.float32).reshape(n, 1)
    sym = None if symmetry is None else np.ascontiguousarray(symmetry, np.int8)
    if out is None:
        policy = np.empty((n, 4 * hw), np.float32)
        value = np.empty((n, 2), np.float32)
        misc = np.empty((n, 2), np.float32)
        own = np.empty((n, hw), np.float32) if ownership else None
    else:
        policy, value, misc, own = out
    check(lib().kc_forward(handle._p, n, ptr(rowSpatial), ptr(rowGlobal), ptr(sym), ptr(policy), ptr(value), ptr(misc), ptr(own)))
    return policy, value, misc, own


def _desc_conv(d):
    w = np.ascontiguousarray(d["weights"], np.float32)
    return capi.ConvDesc(d["convYSize"], d["convXSize"], d["inChannels"], d["outChannels"], w.ctypes.data_as(capi.c_float_p)), w


def _desc_bn(d):
    arrs = [np.ascontiguousarray(d[k], np.float32) for k in ("mean", "variance", "scale", "bias")]
    return capi.BNDesc(d["numChannels"], d["epsilon"], int(d["hasScale"]), int(d["hasBias"]),
                       *[a.ctypes.data_as(capi.c_float_p) for a in arrs]), arrs


def _desc_matmul(d):
    w = np.ascontiguousarray(d["weights"], np.float32)
    return capi.MatMulDesc(d["inChannels"], d["outChannels"], w.ctypes.data_as(capi.c_float_p)), w


def block_desc_from_dict(d):
    """Builds a BlockDesc from a dict shaped like the reference's ResidualBlockDesc /
    GlobalPoolingResidualBlockDesc (as extracted into tests/golden/nn_layers_golden.json)."""
    keep = []
    b = capi.BlockDesc()
    b.kind = 2 if "gpoolConv" in d else 0
    b.preActivation = b.gpoolActivation = b.midActivation = 1   # ActivationLayerDesc default = ReLU
    for name, fn in (("preBN", _desc_bn), ("regularConv", _desc_conv), ("midBN", _desc_bn), ("finalConv", _desc_conv)):
        s, k = fn(d[name]); setattr(b, name, s); keep.append(k)
    if b.kind == 2:
        for name, fn in (("gpoolConv", _desc_conv), ("gpoolBN", _desc_bn), ("gpoolToBiasMul", _desc_matmul)):
            s, k = fn(d[name]); setattr(b, name, s); keep.append(k)
    return b, keep


def testEvaluateConv(ctx, desc, batchSize, nnXLen, nnYLen, useNHWC, inputBuffer):
    d, keep = _desc_conv(desc)
    inp = np.ascontiguousarray(inputBuffer, np.float32)
    out = np.empty(batchSize * nnXLen * nnYLen * desc["outChannels"], np.float32)
    check(lib().kc_test_conv(ctx._p, C.byref(d), batchSize, nnXLen, nnYLen, int(useNHWC), ptr(inp), ptr(out)))
    return out


def testEvaluateBatchNorm(ctx, desc, batchSize, nnXLen, nnYLen, useNHWC, inputBuffer, maskBuffer, activation=0):
    d, keep = _desc_bn(desc)
    inp = np.ascontiguousarray(inputBuffer, np.float32)
    mask = np.ascontiguousarray(maskBuffer, np.float32)
    out = np.empty_like(inp)
    check(lib().kc_test_batchnorm(ctx._p, C.byref(d), activation, batchSize, nnXLen, nnYLen, int(useNHWC), ptr(inp), ptr(mask), ptr(out)))
    return out


def testEvaluateResidualBlock(ctx, desc, batchSize, nnXLen, nnYLen, useNHWC, inputBuffer, maskBuffer):
    """Also serves testEvaluateGlobalPoolingResidualBlock (the desc decides)."""
    b, keep = block_desc_from_dict(desc)
    inp = np.ascontiguousarray(inputBuffer, np.float32)
    mask = np.ascontiguousarray(maskBuffer, np.float32)
    out = np.empty_like(inp)
    check(lib().kc_test_resblock(ctx._p, C.byref(b), batchSize, nnXLen, nnYLen, int(useNHWC), ptr(inp), ptr(mask), ptr(out)))
    return out


class Games:
    """G concurrent device-resident games (Board + BoardHistory of the reference, batched)."""

    def __init__(self, ctx, numGames, xSize=5, ySize=5, winLen=4):
        self.ctx, self.G, self.W, self.H, self.K = ctx, numGames, xSize, ySize, winLen
        self.HW = xSize * ySize
        self.LW = (4 * self.HW + 31) // 32
        self._p = C.c_void_p()
        check(lib().kc_games_create(ctx._p, numGames, xSize, ySize, winLen, C.byref(self._p)))

    def reset(self, seed=0, firstGameId=0, autoRefill=False):
        check(lib().kc_games_reset(self._p, seed, firstGameId, int(autoRefill)))

    def load(self, g0, stones, nextPla, moves=None, numTurns=None):
        stones = np.ascontiguousarray(stones, np.int8)
        n = stones.shape[0]
        nextPla = np.ascontiguousarray(nextPla, np.int8)
        moves = None if moves is None else np.ascontiguousarray(moves, np.int16)
        numTurns = None if numTurns is None else np.ascontiguousarray(numTurns, np.int32)
        check(lib().kc_games_load(self._p, g0, n, ptr(stones), ptr(nextPla), ptr(moves), ptr(numTurns)))

    def step(self, movePos=None):
        G = self.G
        mv = None if movePos is None else np.ascontiguousarray(movePos, np.int16)
        legal = np.empty((G, self.LW), np.uint32)
        status = np.empty(G, np.uint32)
        sitHash = np.empty((G, 2), np.uint64)
        played = np.empty(G, np.int16)
        ids = np.empty(G, np.uint64)
        check(lib().kc_games_step(self._p, ptr(mv), ptr(legal), ptr(status), ptr(sitHash), ptr(played), ptr(ids)))
        return dict(legal=legal, status=status, sitHash=sitHash, played=played, gameIds=ids)

    def features(self, nhwc=False, symmetry=None):
        planes = np.empty((self.G, 15 * self.HW), np.float32)
        glob = np.empty((self.G, 1), np.float32)
        sym = None if symmetry is None else np.ascontiguousarray(symmetry, np.int8)
        check(lib().kc_games_features(self._p, int(nhwc), ptr(sym), ptr(planes), ptr(glob)))
        return planes, glob

    def eval(self, handle, symmetry=None):
        sym = None if symmetry is None else np.ascontiguousarray(symmetry, np.int8)
        check(lib().kc_games_eval(self._p, handle._p, ptr(sym)))

    def postprocess(self, handle, policyTemperature=1.0):
        """NNEvaluator::evaluate post-processing of the last eval(): (policyProbs, whiteWinLoss, misc, nnHash)."""
        policy = np.empty((self.G, 4 * self.HW), np.float32)
        wl = np.empty((self.G, 2), np.float32)
        misc = np.empty((self.G, 2), np.float32)
        nnh = np.empty((self.G, 2), np.uint64)
        check(lib().kc_games_postprocess(self._p, handle._p, policyTemperature, ptr(policy), ptr(wl), ptr(misc), ptr(nnh)))
        return policy, wl, misc, nnh

    def run(self, handle, plies, stats=None):
        st = stats if stats is not None else capi.Stats()
        check(lib().kc_games_run(self._p, handle._p if handle is not None else None, plies, C.byref(st)))
        return st

    def runTimed(self, handle, plies, flushL2Bytes=0, stats=None):
        """Returns (stats, device milliseconds summed over the plies)."""
        st = stats if stats is not None else capi.Stats()
        ms = C.c_float()
        check(lib().kc_games_run_timed(self._p, handle._p if handle is not None else None, plies, flushL2Bytes, C.byref(st), C.byref(ms)))
        return st, ms.value

    def readRunOutputs(self):
        """What the last ply of the last rules+features run(None, plies) wrote: dict(planes [G, 15*HW], global, legal, status, sitHash, played)."""
        G, HW = self.G, self.HW
        out = dict(planes=np.empty((G, 15 * HW), np.float32), glob=np.empty(G, np.float32), legal=np.empty((G, self.LW), np.uint32),
                   status=np.empty(G, np.uint32), sitHash=np.empty((G, 2), np.uint64), played=np.empty(G, np.int16))
        check(lib().kc_games_read_run_outputs(self._p, ptr(out["planes"]), ptr(out["glob"]), ptr(out["legal"]), ptr(out["status"]),
                                              ptr(out["sitHash"]), ptr(out["played"])))
        return out

    def readRunPly(self, pliesBack, planes=True):
        """Outputs of one of the last four plies of the last rules+features launch (0 = last ply): dict(planes, legal, status, sitHash, played)."""
        G, HW = self.G, self.HW
        out = dict(planes=np.empty((G, 15 * HW), np.float32) if planes else None, legal=np.empty((G, self.LW), np.uint32),
                   status=np.empty(G, np.uint32), sitHash=np.empty((G, 2), np.uint64), played=np.empty(G, np.int16))
        check(lib().kc_games_read_run_ply(self._p, pliesBack, ptr(out["planes"]) if planes else None, ptr(out["legal"]), ptr(out["status"]),
                                          ptr(out["sitHash"]), ptr(out["played"])))
        return out

    def launchCount(self):
        return int(lib().kc_games_launch_count(self._p))

    def lastKernelMs(self):
        return float(lib().kc_games_last_kernel_ms(self._p))

    def close(self):
        if self._p:
            lib().kc_games_destroy(self._p)
            self._p = C.c_void_p()


class Search:
    """G concurrent PUCT tree searches in lock step on the device (the reference's Search, batched over games).
    handle=None selects the integer-hash evaluator used to test the search logic exactly against the oracle."""

    def __init__(self, ctx, handle, numGames, xSize=5, ySize=5, winLen=4, maxVisits=800, temperaturePlies=0, autoRefill=False,
                 cpuctExploration=1.0, fpuReductionMax=0.2, rootFpuReductionMax=0.2, noCompaction=False, reuseTree=False,
                 useGraphSearch=False, subtreeValueBiasFactor=0.0, subtreeValueBiasWeightExponent=0.5, subtreeValueBiasFreeProp=0.8, **options):
        """options: further kc_search_params fields by name (rootNoiseEnabled, rootDirichletNoiseTotalConcentration,
        rootDirichletNoiseWeight, rootPolicyTemperature[Early], chosenMoveTemperatureHalflife, fpuParentWeightByVisitedPolicy[Pow],
        rootDesiredPerChildVisitsCoeff)."""
        self.ctx, self.G, self.W, self.H = ctx, numGames, xSize, ySize
        self.P = 4 * xSize * ySize
        self.params = capi.SearchParams(maxVisits, temperaturePlies, int(autoRefill), int(noCompaction), int(reuseTree), int(useGraphSearch), cpuctExploration, fpuReductionMax, rootFpuReductionMax,
                                        subtreeValueBiasFactor, subtreeValueBiasWeightExponent, subtreeValueBiasFreeProp)
        for k, v in options.items():
            if k not in dict(capi.SearchParams._fields_):
                raise TypeError(f"unknown search option {k}")
            setattr(self.params, k, v)
        self._p = C.c_void_p()
        check(lib().kc_search_create(ctx._p, handle._p if handle is not None else None, numGames, xSize, ySize, winLen,
                                     C.byref(self.params), C.byref(self._p)))
        # a non-owning view of the games object inside the search
        self.games = Games.__new__(Games)
        self.games.ctx, self.games.G, self.games.W, self.games.H, self.games.K = ctx, numGames, xSize, ySize, winLen
        self.games.HW = xSize * ySize
        self.games.LW = (4 * self.games.HW + 31) // 32
        self.games._p = C.c_void_p(lib().kc_search_games(self._p))
        self.games.close = lambda: None

    def reset(self, seed=0, firstGameId=0):
        check(lib().kc_search_reset(self._p, seed, firstGameId))

    def runVisits(self):
        check(lib().kc_search_run_visits(self._p))

    def readRoot(self):
        G, P = self.G, self.P
        out = dict(rootVisits=np.zeros(G, np.int32), rootUtilitySum=np.zeros(G, np.float64), edgeVisits=np.zeros((G, P), np.int32),
                   edgeUtilitySum=np.zeros((G, P), np.float64), policy=np.zeros((G, P), np.float32), order=np.zeros((G, P), np.uint8))
        check(lib().kc_search_read_root(self._p, ptr(out["rootVisits"]), ptr(out["rootUtilitySum"]), ptr(out["edgeVisits"]),
                                        ptr(out["edgeUtilitySum"]), ptr(out["policy"]), ptr(out["order"])))
        return out

    def readPlaySelection(self):
        """The play-selection values [G, P] the last play() chose its last move from."""
        out = np.zeros((self.G, self.P), np.float64)
        check(lib().kc_search_read_play_selection(self._p, ptr(out)))
        return out

    def treeDigest(self):
        """Graph-mode searches: hash over every node of each game's graph (see kc_search_tree_digest)."""
        out = np.zeros(self.G, np.uint64)
        check(lib().kc_search_tree_digest(self._p, ptr(out)))
        return out

    def play(self, moves=1, stats=None):
        """Returns (stats, last moves played [G], device ms)."""
        st = stats if stats is not None else capi.SearchStats()
        chosen = np.empty(self.G, np.int16)
        ms = C.c_float()
        check(lib().kc_search_play(self._p, moves, ptr(chosen), C.byref(st), C.byref(ms)))
        return st, chosen, ms.value

    def enableTrainingRows(self, maxRows):
        self._maxRows = maxRows
        check(lib().kc_search_enable_training_rows(self._p, maxRows))

    def readTrainingRows(self, clear=True):
        """The finished games' rows as the reference's npz arrays (dict name -> array) + the number of dropped rows."""
        HW = self.W * self.H
        n = self._maxRows
        out = dict(binaryInputNCHWPacked=np.zeros((n, 15, (HW + 7) // 8), np.uint8), globalInputNC=np.zeros((n, 1), np.float32),
                   policyTargetsNCMove=np.zeros((n, 2, self.P), np.int16), globalTargetsNC=np.zeros((n, 64), np.float32),
                   valueTargetsNCHW=np.zeros((n, 5, self.H, self.W), np.int8))
        rows, dropped = C.c_int(), C.c_int()
        check(lib().kc_search_read_training_rows(self._p, C.byref(rows), C.byref(dropped), ptr(out["binaryInputNCHWPacked"]), ptr(out["globalInputNC"]),
                                                 ptr(out["policyTargetsNCMove"]), ptr(out["globalTargetsNC"]), ptr(out["valueTargetsNCHW"]), int(clear)))
        return {k: v[:rows.value] for k, v in out.items()}, dropped.value

    def writeTrainingNpz(self, path, clear=True):
        """numpy .npz with the reference's array names (TrainingWriteBuffers::writeToZipFile, trainingwrite.cpp:566-587)."""
        rows, dropped = self.readTrainingRows(clear)
        writeTrainingNpz(path, self.W, self.H, rows)
        return len(rows["globalInputNC"]), dropped

    def launchCount(self):
        return int(lib().kc_search_launch_count(self._p))

    def close(self):
        if self._p:
            lib().kc_search_destroy(self._p)
            self._p = C.c_void_p()


def selfplayRun(model, devices, gamesPerDevice, xSize=5, ySize=5, winLen=4, moves=8, movesPerChunk=0, warmupMoves=1, staggerPlies=0,
                maxRowsPerChunk=0, outputDir=None, seed=0, firstGameId=0, handleFlags=0, noNccl=False, maxVisits=800, **searchOptions):
    """kc_selfplay_run: self-play on several GPUs from this one process (one search pool + writer thread per device, one ncclReduce
    of the counters at the end).  searchOptions: kc_search_params fields by name.  Returns (capi.SearchStats totals, capi.SelfplayReport)."""
    params = capi.SearchParams()
    params.maxVisits, params.cpuctExploration, params.fpuReductionMax, params.rootFpuReductionMax = maxVisits, 1.0, 0.2, 0.2
    params.subtreeValueBiasWeightExponent, params.subtreeValueBiasFreeProp = 0.5, 0.8
    for k, v in searchOptions.items():
        if k not in dict(capi.SearchParams._fields_):
            raise TypeError(f"unknown search option {k}")
        setattr(params, k, v)
    devs = (C.c_int32 * len(devices))(*devices)
    cfg = capi.SelfplayConfig(len(devices), devs, gamesPerDevice, xSize, ySize, winLen, moves, movesPerChunk, warmupMoves, staggerPlies,
                              maxRowsPerChunk, int(noNccl), handleFlags, seed, firstGameId, os.fsencode(outputDir) if outputDir else None)
    total, report = capi.SearchStats(), capi.SelfplayReport()
    check(lib().kc_selfplay_run(C.byref(cfg), C.byref(model.desc), C.byref(params), C.byref(total), C.byref(report)))
    return total, report


class NNEvaluator:
    """The evaluator front end (kc_evaluator_*): NNEvaluator::evaluate / serve / NNCacheTable of cpp/neuralnet/nneval.h:17-249
    for CPU search threads.  `evaluate` is thread-safe and blocks (ctypes releases the GIL for its duration).
    With `customBackend` (a callable (serverThread, EvalBatch) -> status) no device is used: host-logic tests only."""

    def __init__(self, ctx=None, loadedModel=None, nnXLen=5, nnYLen=5, winLen=4, maxBatchSize=1024, maxConcurrentEvals=None, numThreads=1,
                 nnCacheSizePowerOfTwo=16, nnMutexPoolSizePowerofTwo=10, doRandomize=False, defaultSymmetry=0, randSeed=0,
                 nnPolicyTemperature=1.0, useFP32Check=False, playModeSymmetry=False, customBackend=None):
        self.nnXLen, self.nnYLen, self.winLen = nnXLen, nnYLen, winLen
        self.cfg = capi.EvaluatorConfig(nnXLen, nnYLen, winLen, maxBatchSize, maxConcurrentEvals or 2 * maxBatchSize, numThreads,
                                        nnCacheSizePowerOfTwo, nnMutexPoolSizePowerofTwo, int(doRandomize), defaultSymmetry, randSeed,
                                        nnPolicyTemperature,
                                        (capi.FLAG_FP32_CHECK if useFP32Check else 0) | (capi.FLAG_SYM_PERMUTE_DIRS if playModeSymmetry else 0))
        self._p = C.c_void_p()
        self._keep = (ctx, loadedModel)
        if customBackend is not None:
            def tramp(user, serverThread, batch):
                try:
                    return int(customBackend(serverThread, batch.contents) or 0)
                except Exception:   # an exception must not unwind through the C server thread
                    import traceback
                    traceback.print_exc()
                    return 99
            self._fn = capi.EVAL_BACKEND_FN(tramp)
            check(lib().kc_evaluator_create_custom(C.byref(self.cfg), self._fn, None, C.byref(self._p)))
        elif isinstance(ctx, (list, tuple)):   # gpuIdxByServerThread: one (context, model) pair per server thread
            assert len(ctx) == len(loadedModel) == numThreads
            cs = (C.c_void_p * numThreads)(*[c._p for c in ctx])
            ms = (C.c_void_p * numThreads)(*[m._p for m in loadedModel])
            check(lib().kc_evaluator_create_multi(numThreads, cs, ms, C.byref(self.cfg), C.byref(self._p)))
        else:
            check(lib().kc_evaluator_create(ctx._p, loadedModel._p, C.byref(self.cfg), C.byref(self._p)))

    def _position(self, stones, nextPla, moves, numTurns, keep):
        st = np.ascontiguousarray(stones, dtype=np.int8).reshape(-1)
        assert st.size == self.nnXLen * self.nnYLen
        keep.append(st)
        mv = None
        if moves is not None:
            mv = np.ascontiguousarray(moves, dtype=np.int16).reshape(-1)
            assert mv.size == 10
            keep.append(mv)
        return capi.EvalPosition(st.ctypes.data_as(C.POINTER(C.c_int8)), mv.ctypes.data_as(C.POINTER(C.c_int16)) if mv is not None else None,
                                 int(numTurns), int(nextPla))

    def _result(self, out, policy, own):
        return {"policyProbs": policy, "whiteOwnerMap": own, "whiteWinProb": out.whiteWinProb, "whiteLossProb": out.whiteLossProb,
                "varTimeLeft": out.varTimeLeft, "shorttermWinlossError": out.shorttermWinlossError,
                "nnHash": (int(out.nnHash[0]), int(out.nnHash[1])), "symmetry": int(out.symmetry), "cacheHit": bool(out.cacheHit)}

    def evaluate(self, stones, nextPla, moves=None, numTurns=0, symmetry=-1, skipCache=False, includeOwnerMap=False):
        """stones [H*W] (0/1/2), moves [5][2] last five (policy index, player) oldest first with -1 = none."""
        keep = []
        pos = self._position(stones, nextPla, moves, numTurns, keep)
        HW = self.nnXLen * self.nnYLen
        policy = np.empty(4 * HW, np.float32)
        own = np.empty(HW, np.float32) if includeOwnerMap else None
        out = capi.EvalOutput()
        out.policyProbs = policy.ctypes.data_as(capi.c_float_p)
        out.whiteOwnerMap = own.ctypes.data_as(capi.c_float_p) if own is not None else None
        check(lib().kc_evaluator_evaluate(self._p, C.byref(pos), symmetry, int(skipCache), int(includeOwnerMap), C.byref(out)))
        return self._result(out, policy, own)

    def evaluateMany(self, stones, nextPla, moves=None, numTurns=None, symmetry=None, skipCache=False, includeOwnerMap=False):
        """stones [n][H*W], nextPla [n], moves [n][5][2] or None, numTurns [n] or None, symmetry [n] (int8) or None."""
        stones = np.ascontiguousarray(stones, dtype=np.int8)
        n, HW = stones.shape[0], self.nnXLen * self.nnYLen
        keep = []
        pos = (capi.EvalPosition * n)()
        for i in range(n):
            pos[i] = self._position(stones[i], nextPla[i], None if moves is None else moves[i], 0 if numTurns is None else numTurns[i], keep)
        policy = np.empty((n, 4 * HW), np.float32)
        own = np.empty((n, HW), np.float32) if includeOwnerMap else None
        outs = (capi.EvalOutput * n)()
        for i in range(n):
            outs[i].policyProbs = policy[i].ctypes.data_as(capi.c_float_p)
            outs[i].whiteOwnerMap = own[i].ctypes.data_as(capi.c_float_p) if own is not None else None
        sym = None if symmetry is None else np.ascontiguousarray(symmetry, dtype=np.int8)
        check(lib().kc_evaluator_evaluate_many(self._p, n, pos, ptr(sym), int(skipCache), int(includeOwnerMap), outs))
        return [self._result(outs[i], policy[i], None if own is None else own[i]) for i in range(n)]

    def clearCache(self):
        check(lib().kc_evaluator_clear_cache(self._p))

    def stats(self):
        st = capi.EvaluatorStats()
        check(lib().kc_evaluator_get_stats(self._p, C.byref(st)))
        return {n: int(getattr(st, n)) for n, _ in st._fields_}

    def clearStats(self):
        check(lib().kc_evaluator_clear_stats(self._p))

    def numRowsProcessed(self):
        return self.stats()["rowsProcessed"]

    def numBatchesProcessed(self):
        return self.stats()["batchesProcessed"]

    def averageProcessedBatchSize(self):
        st = self.stats()
        return st["rowsProcessed"] / max(st["batchesProcessed"], 1)

    def close(self):
        if self._p:
            lib().kc_evaluator_destroy(self._p)
            self._p = C.c_void_p()


def evalPositionHash(xSize, ySize, stones, nextPla, moves=None, numTurns=0, policyTemperature=1.0):
    """(NNInputs::getHash, the evaluator's cache key) of a position, computed on the host."""
    st = np.ascontiguousarray(stones, dtype=np.int8).reshape(-1)
    mv = None if moves is None else np.ascontiguousarray(moves, dtype=np.int16).reshape(-1)
    pos = capi.EvalPosition(st.ctypes.data_as(C.POINTER(C.c_int8)), mv.ctypes.data_as(C.POINTER(C.c_int16)) if mv is not None else None,
                            int(numTurns), int(nextPla))
    h, k = (C.c_uint64 * 2)(), (C.c_uint64 * 2)()
    check(lib().kc_eval_position_hash(xSize, ySize, C.byref(pos), policyTemperature, h, k))
    return (int(h[0]), int(h[1])), (int(k[0]), int(k[1]))


def evalUnpackPosition(xSize, ySize, black, white, misc):
    """Inverse of the evaluator's row packing: (stones [H*W], nextPla, moves [5][2] = (cell, player) oldest first, numTurns, lastDir)."""
    stones = np.zeros(xSize * ySize, np.int8)
    moves = np.zeros((5, 2), np.int16)
    pla, nt, ld = C.c_int8(), C.c_int32(), C.c_int32()
    check(lib().kc_eval_unpack_position(xSize, ySize, int(black), int(white), int(misc), ptr(stones), C.byref(pla), ptr(moves), C.byref(nt), C.byref(ld)))
    return stones, int(pla.value), moves, int(nt.value), int(ld.value)


def evalUnpackPositionWide(xSize, ySize, blackLo, blackHi, whiteLo, whiteHi, misc):
    """The same for a row of a board beyond 7x7 (EvalBatch.blackHi / whiteHi non-null)."""
    stones = np.zeros(xSize * ySize, np.int8)
    moves = np.zeros((5, 2), np.int16)
    pla, nt, ld = C.c_int8(), C.c_int32(), C.c_int32()
    check(lib().kc_eval_unpack_position_wide(xSize, ySize, int(blackLo), int(blackHi), int(whiteLo), int(whiteHi), int(misc), ptr(stones),
                                             C.byref(pla), ptr(moves), C.byref(nt), C.byref(ld)))
    return stones, int(pla.value), moves, int(nt.value), int(ld.value)


def writeTrainingNpz(path, xSize, ySize, rows):
    """kc_training_write_npz: the reference's training-data file (trainingwrite.cpp:566-587) from a dict of the five row arrays."""
    n = len(rows["globalInputNC"])
    a = [np.ascontiguousarray(rows[k], dt) for k, dt in (("binaryInputNCHWPacked", np.uint8), ("globalInputNC", np.float32), ("policyTargetsNCMove", np.int16),
                                                          ("globalTargetsNC", np.float32), ("valueTargetsNCHW", np.int8))]
    HW = xSize * ySize
    assert a[0].shape == (n, 15, (HW + 7) // 8) and a[1].shape == (n, 1) and a[2].shape == (n, 2, 4 * HW) and a[3].shape == (n, 64) and a[4].shape == (n, 5, ySize, xSize)
    check(lib().kc_training_write_npz(os.fsencode(path), n, xSize, ySize, *[ptr(x) for x in a]))


def zobristTables():
    board = np.zeros((133, 4, 2), np.uint64)
    player = np.zeros((4, 2), np.uint64)
    sx = np.zeros((11, 2), np.uint64)
    sy = np.zeros((11, 2), np.uint64)
    check(lib().kc_zobrist_tables(ptr(board), ptr(player), ptr(sx), ptr(sy)))
    return board, player, sx, sy


def selftestUmma(ctx, A_bf16_bits, B_bf16_bits, shift, ws=False):
    rowsA, K = A_bf16_bits.shape
    N = B_bf16_bits.shape[0]
    D = np.zeros((2, 128, N) if ws else (128, N), np.float32)
    A = np.ascontiguousarray(A_bf16_bits, np.uint16)
    B = np.ascontiguousarray(B_bf16_bits, np.uint16)
    check(lib().kc_selftest_umma(ctx._p, ptr(A), ptr(B), ptr(D), rowsA, N, K, shift, int(ws)))
    return D
