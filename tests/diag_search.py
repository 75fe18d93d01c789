"""Diagnostic (not a test): a short batched tree search at the bench shape, for per-kernel timing under ncu."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend, modeldesc
ctx = backend.createComputeContext(0)
G, V = 18944, int(sys.argv[1]) if len(sys.argv) > 1 else 24
lm = backend.LoadedModel(ctx, modeldesc.Model("b10c128", seed=1))
h = backend.createComputeHandle(ctx, lm, G, 5, 5)
graph = len(sys.argv) > 2 and sys.argv[2] == "graph"
kw = dict(useGraphSearch=True, subtreeValueBiasFactor=0.3, subtreeValueBiasWeightExponent=0.8) if graph else {}
s = backend.Search(ctx, h, G, 5, 5, 4, maxVisits=V, temperaturePlies=30, autoRefill=True, **kw)
s.reset(seed=1)
lane = np.arange(G)
for t in range(20):
    s.games.step(np.where(lane % 20 > t, -2, -1).astype(np.int16))
h.trunkTime()
st, _, ms = s.play(1)
tms, tcnt = h.trunkTime()
print(f"{ms/V:.3f} ms/iteration, net evals {st.netEvals/st.visits:.3f} of visits; trunk launches {tcnt}, {tms/V:.3f} ms of trunk per iteration")
