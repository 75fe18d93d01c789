"""Diagnostic (GPU box): evaluator device backend vs the direct games path, row by row, for several server / client settings."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import kc_oracle as oracle  # noqa: E402
from katacoffee_b200 import backend, modeldesc  # noqa: E402
from test_evaluator import H, K, W, expected_output, make_positions  # noqa: E402

oracle.build()
N = 300
ps = make_positions(oracle, N, seed=41)
stones = np.stack([p["stones"] for p in ps]); nextPla = np.array([p["nextPla"] for p in ps], np.int8)
moves = np.stack([p["moves"] for p in ps]); numTurns = np.array([p["numTurns"] for p in ps], np.int32)
model = modeldesc.Model("b2c32", seed=4)
omodel = oracle.Model(model)
ctx = backend.createComputeContext(0)
lm = backend.LoadedModel(ctx, model)
sym = (np.arange(N) % 8).astype(np.int8)
for mode in ("bf16", "fp32"):
    fp32 = mode == "fp32"
    h = backend.createComputeHandle(ctx, lm, N, W, H, useFP32Check=fp32)
    games = backend.Games(ctx, N, W, H, K)
    games.load(0, stones, nextPla, moves, numTurns)
    games.eval(h, sym)
    dpol, dwl, dmisc, dhash = games.postprocess(h, 1.0)
    # direct path again in a 64-lane object, 64 rows at a time and one row at a time
    h64 = backend.createComputeHandle(ctx, lm, 64, W, H, useFP32Check=fp32)
    g64 = backend.Games(ctx, 64, W, H, K)
    bad = []
    for i0 in range(0, 256, 64):
        g64.load(0, stones[i0:i0 + 64], nextPla[i0:i0 + 64], moves[i0:i0 + 64], numTurns[i0:i0 + 64])
        g64.eval(h64, sym[i0:i0 + 64])
        p64 = g64.postprocess(h64, 1.0)[0]
        bad += [i0 + j for j in range(64) if np.abs(p64[j] - dpol[i0 + j]).max() > 2e-6]
    print(mode, "direct 64-lane object, full batches: mismatching rows", bad, flush=True)
    bad = []
    for i in range(0, 64):
        g64.load(0, stones[i:i + 1], nextPla[i:i + 1], moves[i:i + 1], numTurns[i:i + 1])
        s64 = np.zeros(64, np.int8); s64[0] = sym[i]
        g64.eval(h64, s64)
        p64 = g64.postprocess(h64, 1.0)[0]
        if np.abs(p64[0] - dpol[i]).max() > 2e-6:
            bad.append((i, float(np.abs(p64[0] - dpol[i]).max())))
    print(mode, "direct 64-lane object, one row loaded at lane 0 (others stale): mismatching rows", bad, flush=True)
    for servers, clients in ((1, 1), (1, 6), (2, 6)):
        ev = backend.NNEvaluator(ctx, lm, nnXLen=W, nnYLen=H, winLen=K, maxBatchSize=64, maxConcurrentEvals=256, numThreads=servers, nnCacheSizePowerOfTwo=-1,
                                 useFP32Check=fp32)
        results = [None] * N

        def client(t):
            for i in range(t, N, clients):
                results[i] = ev.evaluate(stones[i], nextPla[i], moves[i], numTurns[i], symmetry=int(sym[i]))

        ths = [threading.Thread(target=client, args=(t,)) for t in range(clients)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        bad = [(i, int(sym[i]), round(float(np.abs(results[i]["policyProbs"] - dpol[i]).max()), 5)) for i in range(N)
               if np.abs(results[i]["policyProbs"] - dpol[i]).max() > 2e-6]
        print(mode, f"evaluator servers={servers} clients={clients}: {len(bad)} mismatching rows, avg batch {ev.averageProcessedBatchSize():.2f}", bad[:12], flush=True)
        for i, _, _ in bad[:3]:
            e = expected_output(oracle, omodel, ps[i], int(sym[i]))
            print("   row", i, "evaluator vs oracle", float(np.abs(results[i]["policyProbs"] - e["policy"]).max()), "direct vs oracle", float(np.abs(dpol[i] - e["policy"]).max()),
                  "numTurns", int(numTurns[i]), "moves", moves[i].tolist(), flush=True)
        # many path
        res = ev.evaluateMany(stones, nextPla, moves, numTurns, symmetry=sym)
        bad = [(i, int(sym[i]), round(float(np.abs(res[i]["policyProbs"] - dpol[i]).max()), 5)) for i in range(N) if np.abs(res[i]["policyProbs"] - dpol[i]).max() > 2e-6]
        print(mode, f"evaluateMany servers={servers}: {len(bad)} mismatching rows", bad[:12], flush=True)
        ev.close()
    for x in (g64, h64, games, h):
        x.close()
