import sys, os, json
sys.path.insert(0, os.getcwd())
import numpy as np
from katacoffee_b200 import backend, modeldesc
ctx = backend.createComputeContext(0)
for net, W, G in (("b20c256", 5, 148 * 4 * 16), ("b20c256", 6, 148 * 3 * 16)):
    model = modeldesc.Model(net, seed=1)
    lm = backend.LoadedModel(ctx, model)
    h = backend.createComputeHandle(ctx, lm, G, W, W)
    games = backend.Games(ctx, G, W, W, 4); games.reset(seed=1, autoRefill=True)
    games.runTimed(h, 3, 256 << 20); h.trunkTime()
    st, ms = games.runTimed(h, 10, 256 << 20)
    tms, tn = h.trunkTime()
    fl = modeldesc.flops_per_eval(net, W * W)
    print(json.dumps({"net": net, "board": W, "rows": G, "evals_per_s": G * 10 / (ms * 1e-3), "trunk_ms": tms / tn, "tflops": fl * G / (tms / tn * 1e-3) / 1e12, "flops_per_eval": fl}))
    games.close(); h.close(); lm.close()
