"""Diagnostic (not a test): kc_selfplay_run over every visible GPU from this one process -- one pool + writer thread per device, one
ncclReduce of the counters at the end (csrc/selfplay.cpp).  Prints one JSON line per device count (1, then all).
Usage (2 GPUs): gpurun --gpus 2 -- python tests/diag_selfplay_multi.py [games_per_device] [moves]"""
import json, os, sys, tempfile, shutil
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
from katacoffee_b200 import backend, modeldesc
from katacoffee_b200.capi import lib
G = int(sys.argv[1]) if len(sys.argv) > 1 else 148 * 128
moves = int(sys.argv[2]) if len(sys.argv) > 2 else 4
n = C.c_int()
lib().kc_device_count(C.byref(n))
model = modeldesc.Model("b10c128", seed=1)
kw = dict(useGraphSearch=1, subtreeValueBiasFactor=0.30, subtreeValueBiasWeightExponent=0.8, reuseTree=1, cpuctExploration=1.1, rootFpuReductionMax=0.0,
          rootNoiseEnabled=1, rootDirichletNoiseTotalConcentration=10.83, rootDirichletNoiseWeight=0.25, rootPolicyTemperature=1.1, rootPolicyTemperatureEarly=1.25,
          chosenMoveTemperatureHalflife=19.0, fpuParentWeightByVisitedPolicy=1, fpuParentWeightByVisitedPolicyPow=2.0, rootDesiredPerChildVisitsCoeff=2.0,
          valueWeightExponent=0.5, chosenMoveTemperatureEarly=0.75, chosenMoveTemperature=0.15, chosenMovePrune=1.0, nnRandomize=1,
          rootNumSymmetriesToSample=4, useLcbForSelection=1, lcbStdevs=5.0, minVisitPropForLCB=0.15, useNonBuggyLcb=1,
          nnCacheSizePowerOfTwo=int(os.environ.get("KC_DIAG_NNCACHE", 21)))
for devs in ([0], list(range(n.value))):
    out = tempfile.mkdtemp(prefix="kc_sp_")
    try:
        tot, rep = backend.selfplayRun(model, devs, G, 5, 5, 4, moves=moves, movesPerChunk=2, warmupMoves=1, staggerPlies=20, maxRowsPerChunk=G * 6, outputDir=out,
                                       seed=7, maxVisits=800, autoRefill=1, temperaturePlies=30, **kw)
        files = len(os.listdir(out))
    finally:
        shutil.rmtree(out, ignore_errors=True)
    print(json.dumps({"devices": devs, "moves": tot.movesPlayed, "moves_per_s": tot.movesPlayed / rep.wallSeconds, "wall_s": rep.wallSeconds,
                      "device_s_max": rep.deviceMsMax * 1e-3, "visits": tot.visits, "games_finished": tot.gamesFinished, "rows_written": rep.rowsWritten,
                      "rows_dropped": rep.rowsDropped, "npz_files": files, "npz_bytes": rep.bytesWritten, "reduced_with_nccl": rep.reducedWithNccl}), flush=True)
    if len(devs) == n.value:
        break
