"""Diagnostic (not a test): long self-play with every search option on, the bf16 net, auto-refill and training rows; checks the counters' invariants."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from katacoffee_b200 import backend, modeldesc, capi
ctx = backend.createComputeContext(0)
G, V, MOVES = 4096, 200, 60
lm = backend.LoadedModel(ctx, modeldesc.Model("b6c96", seed=3))
h = backend.createComputeHandle(ctx, lm, G, 5, 5)
s = backend.Search(ctx, h, G, 5, 5, 4, maxVisits=V, temperaturePlies=30, autoRefill=True, reuseTree=True, useGraphSearch=True,
                   subtreeValueBiasFactor=0.3, subtreeValueBiasWeightExponent=0.8, cpuctExploration=1.1, rootFpuReductionMax=0.0,
                   rootNoiseEnabled=1, rootDirichletNoiseTotalConcentration=10.83, rootDirichletNoiseWeight=0.25, rootPolicyTemperature=1.1,
                   rootPolicyTemperatureEarly=1.25, chosenMoveTemperatureHalflife=19.0, fpuParentWeightByVisitedPolicy=1,
                   fpuParentWeightByVisitedPolicyPow=2.0, rootDesiredPerChildVisitsCoeff=2.0, valueWeightExponent=0.5,
                   chosenMoveTemperatureEarly=0.75, chosenMoveTemperature=0.15, chosenMovePrune=1.0)
s.reset(seed=5)
s.enableTrainingRows(G * 25 * 4)
st = capi.SearchStats()
tot_ms = 0
for m in range(MOVES // 10):
    _, chosen, ms = s.play(10, st)
    tot_ms += ms
    print(f"after {10*(m+1)} moves: visits {st.visits} evals {st.netEvals} terminal {st.terminalVisits} transpositions {st.transpositionHits} catchup {st.catchUpVisits} "
          f"games finished {st.gamesFinished} (B {st.blackWins} W {st.whiteWins} D {st.draws}) {ms/10:.1f} ms/move-batch", flush=True)
assert st.visits == st.netEvals + st.terminalVisits + st.transpositionHits + st.catchUpVisits
assert st.gamesFinished == st.blackWins + st.whiteWins + st.draws and st.gamesFinished > 2 * G
rows, dropped = s.readTrainingRows()
n = len(rows["globalInputNC"])
print("training rows", n, "dropped", dropped, "moves played", st.movesPlayed)
assert dropped == 0 and n > 0
assert (rows["policyTargetsNCMove"][:, 0].sum(1) > 0).all()
gt = rows["globalTargetsNC"]
assert np.allclose(gt[:, 0] + gt[:, 1], 1.0, atol=1e-5) and np.isfinite(gt).all()
print("stress ok")
