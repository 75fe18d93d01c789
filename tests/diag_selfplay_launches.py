"""Diagnostic (not a test): one self-play move of the bench configuration through kc_selfplay_run, for an ncu launch list
(ncu --metrics gpu__time_duration.sum --launch-skip 3000 --launch-count 6000 python tests/diag_selfplay_launches.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from katacoffee_b200 import backend, modeldesc
G = int(os.environ.get("KC_DIAG_GAMES", 148 * 128))
visits = int(os.environ.get("KC_DIAG_VISITS", 800))
moves = int(os.environ.get("KC_DIAG_MOVES", 1))
model = modeldesc.Model("b10c128", seed=1)
kw = dict(useGraphSearch=1, subtreeValueBiasFactor=0.30, subtreeValueBiasWeightExponent=0.8, reuseTree=1, cpuctExploration=1.1, rootFpuReductionMax=0.0,
          rootNoiseEnabled=1, rootDirichletNoiseTotalConcentration=10.83, rootDirichletNoiseWeight=0.25, rootPolicyTemperature=1.1, rootPolicyTemperatureEarly=1.25,
          chosenMoveTemperatureHalflife=19.0, fpuParentWeightByVisitedPolicy=1, fpuParentWeightByVisitedPolicyPow=2.0, rootDesiredPerChildVisitsCoeff=2.0,
          valueWeightExponent=0.5, chosenMoveTemperatureEarly=0.75, chosenMoveTemperature=0.15, chosenMovePrune=1.0, nnRandomize=1,
          rootNumSymmetriesToSample=4, useLcbForSelection=1, lcbStdevs=5.0, minVisitPropForLCB=0.15, useNonBuggyLcb=1)
if os.environ.get("KC_DIAG_NNCACHE"):
    kw["nnCacheSizePowerOfTwo"] = int(os.environ["KC_DIAG_NNCACHE"])
if os.environ.get("KC_DIAG_PLAIN"):
    kw = dict(reuseTree=1)
t0 = time.time()
tot, rep = backend.selfplayRun(model, [0], G, 5, 5, 4, moves=moves, movesPerChunk=moves, warmupMoves=int(os.environ.get("KC_DIAG_WARMUP", 0)), staggerPlies=20,
                               maxRowsPerChunk=0, seed=1, maxVisits=visits, autoRefill=1, temperaturePlies=30, **kw)
print(json.dumps({"moves": tot.movesPlayed, "visits": tot.visits, "netEvals": tot.netEvals, "batchRows": tot.batchRows, "wall_s": rep.wallSeconds,
                  "device_s": rep.deviceMsMax * 1e-3, "launches": rep.kernelLaunches, "nnCacheHits": tot.nnCacheHits, "rows_per_s": tot.batchRows / rep.wallSeconds,
                  "moves_per_s": tot.movesPlayed / rep.wallSeconds, "total_s": time.time() - t0}))
