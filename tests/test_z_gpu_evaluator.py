"""GPU checks of the evaluator front end's device backend (kc_evaluator_create): packed positions -> rules kernel -> bf16 trunk
tiles -> net -> masked softmax, batched from concurrent client threads, must give what kc_games_load + kc_games_eval +
kc_games_postprocess give for the same positions and symmetries, and stay within the reduced-precision bars of the fp32 oracle.
(Named test_z_* so that it runs after the parity suite.)"""
import json
import os
import subprocess
import threading

import numpy as np
import pytest

from test_evaluator import HW, H, K, W, expected_output, make_positions

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _arrays(ps):
    return (np.stack([p["stones"] for p in ps]), np.array([p["nextPla"] for p in ps], np.int8), np.stack([p["moves"] for p in ps]),
            np.array([p["numTurns"] for p in ps], np.int32))


@pytest.mark.timeout(600)
@pytest.mark.parametrize("mode", ["bf16", "fp32"])
def test_device_evaluator_matches_direct_path_and_oracle(ctx, oracle, mode):
    from katacoffee_b200 import backend, modeldesc
    # distinct positions only: the cache (like the reference's) ignores the symmetry, so a repeated position would be served
    # the result of its first evaluation, under that one's symmetry
    seen, ps = set(), []
    for p in make_positions(oracle, 330, seed=41):
        key = backend.evalPositionHash(W, H, p["stones"], p["nextPla"], p["moves"], p["numTurns"])[1]
        if key not in seen:
            seen.add(key)
            ps.append(p)
    N = len(ps)
    assert N >= 250
    stones, nextPla, moves, numTurns = _arrays(ps)
    model = modeldesc.Model("b2c32", seed=4)
    omodel = oracle.Model(model)
    lm = backend.LoadedModel(ctx, model)
    # the direct path: all positions in one games object, explicit symmetries
    sym = (np.arange(N) % 8).astype(np.int8)
    h = backend.createComputeHandle(ctx, lm, N, W, H, useFP32Check=(mode == "fp32"))
    games = backend.Games(ctx, N, W, H, K)
    games.load(0, stones, nextPla, moves, numTurns)
    games.eval(h, sym)
    dpol, dwl, dmisc, dhash = games.postprocess(h, 1.0)
    down = h.readOutputs(N)[3]
    # the evaluator: batches of at most 64 rows closed whenever a server is free, 2 servers, 6 client threads
    # (fp32: through kc_evaluator_create_multi, one (context, model) pair per server thread -- here the same device twice)
    ev = backend.NNEvaluator([ctx, ctx] if mode == "fp32" else ctx, [lm, lm] if mode == "fp32" else lm, nnXLen=W, nnYLen=H, winLen=K, maxBatchSize=64,
                             maxConcurrentEvals=256, numThreads=2, nnCacheSizePowerOfTwo=16, useFP32Check=(mode == "fp32"))
    results = [None] * N
    errors = []

    def client(t):
        try:
            idx = list(range(t, N, 6))
            if t % 2 == 0:
                for i in idx:
                    results[i] = ev.evaluate(stones[i], nextPla[i], moves[i], numTurns[i], symmetry=int(sym[i]), includeOwnerMap=True)
            else:
                res = ev.evaluateMany(stones[idx], nextPla[idx], moves[idx], numTurns[idx], symmetry=sym[idx], includeOwnerMap=True)
                for i, r in zip(idx, res):
                    results[i] = r
        except BaseException as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=client, args=(t,)) for t in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]
    tol = 2e-6
    worst = 0.0
    for i, r in enumerate(results):
        assert r["nnHash"] == (int(dhash[i][0]), int(dhash[i][1])), "host NNInputs::getHash != the device's"
        assert ((r["policyProbs"] == -1) == (dpol[i] == -1)).all(), "legal mask"
        assert np.abs(r["policyProbs"] - dpol[i]).max() <= tol, i
        assert abs(r["whiteWinProb"] - dwl[i][0]) <= tol and abs(r["whiteLossProb"] - dwl[i][1]) <= tol
        assert np.allclose([r["varTimeLeft"], r["shorttermWinlossError"]], dmisc[i], rtol=1e-6, atol=1e-6)
        sign = 1.0 if nextPla[i] == 2 else -1.0
        assert np.abs(r["whiteOwnerMap"] - sign * np.tanh(down[i])).max() <= 1e-5
        assert r["symmetry"] == sym[i] and not r["cacheHit"]
        if i % 5 == 0:   # against the fp32 oracle: nneval post-processed quantities
            e = expected_output(oracle, omodel, ps[i], int(sym[i]))
            bar = 1e-4 if mode == "fp32" else 1e-2
            worst = max(worst, np.abs(r["policyProbs"] - e["policy"]).max(), abs(r["whiteWinProb"] - e["winLoss"][0]))
            assert np.abs(r["policyProbs"] - e["policy"]).max() < bar and abs(r["whiteWinProb"] - e["winLoss"][0]) < bar
            assert np.abs(r["whiteOwnerMap"] - e["owner"]).max() < (1e-4 if mode == "fp32" else 1e-2)
    st = ev.stats()
    assert st["rowsProcessed"] == N and st["cacheMisses"] == N and st["batchesProcessed"] >= (N + 63) // 64
    # second pass: everything from the cache, identical
    hits = 0
    for i in range(0, N, 7):
        r = ev.evaluate(stones[i], nextPla[i], moves[i], numTurns[i], includeOwnerMap=True)
        if r["cacheHit"]:   # (a direct-mapped table: a few of the N entries have been evicted by another position)
            hits += 1
            assert (r["policyProbs"] == results[i]["policyProbs"]).all() and (r["whiteOwnerMap"] == results[i]["whiteOwnerMap"]).all() and r["symmetry"] == sym[i]
    checked = len(range(0, N, 7))
    assert hits >= 0.8 * checked and ev.stats()["rowsProcessed"] == N + checked - hits
    print(f"evaluator[{mode}]: {st['batchesProcessed']} batches, avg {ev.averageProcessedBatchSize():.1f} rows; worst |diff| vs fp32 oracle {worst:.2e}")
    for x in (ev, games, h, lm):
        x.close()


@pytest.mark.timeout(600)
def test_native_clients_throughput_tool(ctx, tmp_path):
    """tests/cpp/bench_evaluator.cpp (built by build_host) runs on the device: all rows processed, no failures, cache hits with --repeat."""
    from katacoffee_b200 import backend, modeldesc
    exe = os.path.join(ROOT, "katacoffee_b200", "host", "bench_evaluator")
    assert os.path.exists(exe), "run __graft_entry__.build()"
    path = str(tmp_path / "b6c96.bin.gz")
    backend.writeModelFile(modeldesc.Model("b6c96", seed=11), path)
    r = subprocess.run([exe, path, "--clients", "4", "--rows", "20000", "--batch", "2048", "--servers", "2", "--chunk", "256"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["failures"] == 0 and rep["rowsProcessed"] == rep["requests"] and rep["value"] > 0
    print("evaluator throughput (b6c96, 4 clients):", rep)
    r = subprocess.run([exe, path, "--clients", "4", "--rows", "20000", "--batch", "2048", "--cache", "16", "--repeat", "500", "--single"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["failures"] == 0 and rep["cacheHits"] > 0.9 * rep["requests"]


@pytest.mark.timeout(300)
def test_cpp_nnevaluator_class_on_the_device(ctx, tmp_path):
    """class NNEvaluator (host/b200nneval.cpp) with two server threads on GPU 0: four search threads, explicit symmetries, owner maps;
    normalised policies over legal moves only, win + loss = 1."""
    from katacoffee_b200 import backend, modeldesc
    from katacoffee_b200 import build as kb
    kb.build_host()
    exe = os.path.join(ROOT, "katacoffee_b200", "host", "test_b200nneval")
    path = str(tmp_path / "b6c96.bin.gz")
    backend.writeModelFile(modeldesc.Model("b6c96", seed=11), path)
    r = subprocess.run([exe, path, "gpu"], capture_output=True, text=True, timeout=200)
    assert r.returncode == 0 and "test_b200nneval gpu: ok" in r.stdout, r.stdout + r.stderr
    print(r.stdout.strip())


@pytest.mark.gpu
@pytest.mark.timeout(300)
def test_nonfinite_outputs_fail_the_batch(ctx, oracle):
    """NNEvaluator::evaluate throws 'Got nonfinite for policy sum / nneval value' (nneval.cpp:745-750, 789-793): a net whose value
    head produces NaN makes the evaluation fail with an error instead of handing out (and caching) NaN probabilities."""
    from katacoffee_b200 import backend, modeldesc
    model = modeldesc.Model("b2c32", seed=4)
    model._w(model.desc.v3Mul)[:] = np.nan          # the win / loss logits of every row become NaN
    lm = backend.LoadedModel(ctx, model)
    ev = backend.NNEvaluator(ctx, lm, nnXLen=W, nnYLen=H, winLen=K, maxBatchSize=16, maxConcurrentEvals=64, numThreads=1, nnCacheSizePowerOfTwo=10)
    p = make_positions(oracle, 3, seed=5)[0]
    with pytest.raises(RuntimeError, match="nonfinite"):
        ev.evaluate(p["stones"], p["nextPla"], p["moves"], p["numTurns"])
    with pytest.raises(RuntimeError, match="nonfinite"):   # and nothing was cached
        ev.evaluate(p["stones"], p["nextPla"], p["moves"], p["numTurns"])
    assert ev.stats()["cacheHits"] == 0
    ev.close(); lm.close()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("dims", [(9, 9, 5), (10, 10, 5), (8, 7, 4)])
def test_device_evaluator_on_boards_beyond_7x7(ctx, oracle, dims):
    """Boards up to the reference's 10x10 (board.h:120) through the evaluator front end: the rows are staged as 128-bit bitboards with the
    wide misc format and must come out as kc_games_load + kc_games_eval + kc_games_postprocess give them (that path is held to the oracle
    by test_gpu_parity), with the oracle's NNInputs::getHash."""
    from katacoffee_b200 import backend, modeldesc
    Wd, Hd, Kd = dims
    seen, ps = set(), []
    for p in make_positions(oracle, 90, seed=7, max_plies=30, dims=dims):
        key = backend.evalPositionHash(Wd, Hd, p["stones"], p["nextPla"], p["moves"], p["numTurns"])[1]
        if key not in seen:
            seen.add(key)
            ps.append(p)
    N = len(ps)
    assert N >= 60
    stones, nextPla, moves, numTurns = _arrays(ps)
    model = modeldesc.Model("b2c32", seed=9)
    lm = backend.LoadedModel(ctx, model)
    sym = (np.arange(N) % 8).astype(np.int8) if Wd == Hd else (np.arange(N) % 4).astype(np.int8)   # transposing symmetries need a square board
    h = backend.createComputeHandle(ctx, lm, N, Wd, Hd)
    games = backend.Games(ctx, N, Wd, Hd, Kd)
    games.load(0, stones, nextPla, moves, numTurns)
    games.eval(h, sym)
    dpol, dwl, dmisc, dhash = games.postprocess(h, 1.0)
    down = h.readOutputs(N)[3]
    ev = backend.NNEvaluator(ctx, lm, nnXLen=Wd, nnYLen=Hd, winLen=Kd, maxBatchSize=32, maxConcurrentEvals=128, numThreads=2, nnCacheSizePowerOfTwo=12)
    res = ev.evaluateMany(stones, nextPla, moves, numTurns, symmetry=sym, includeOwnerMap=True)
    for i, r in enumerate(res):
        og = oracle.Game(Wd, Hd, Kd)
        for c in range(Wd * Hd):
            if stones[i, c]:
                og.set_stone(c % Wd, c // Wd, int(stones[i, c]))
        og.set_history([(int(m[0]), int(m[1])) for m in moves[i] if m[0] >= 0], int(numTurns[i]), int(nextPla[i]))
        assert r["nnHash"] == tuple(int(x) for x in og.nn_hash()) == (int(dhash[i][0]), int(dhash[i][1]))
        assert ((r["policyProbs"] == -1) == (dpol[i] == -1)).all(), "legal mask"
        assert np.abs(r["policyProbs"] - dpol[i]).max() <= 2e-6, i
        assert abs(r["whiteWinProb"] - dwl[i][0]) <= 2e-6 and abs(r["whiteLossProb"] - dwl[i][1]) <= 2e-6
        assert np.allclose([r["varTimeLeft"], r["shorttermWinlossError"]], dmisc[i], rtol=1e-6, atol=1e-6)
        sign = 1.0 if nextPla[i] == 2 else -1.0
        assert np.abs(r["whiteOwnerMap"] - sign * np.tanh(down[i])).max() <= 1e-5
        assert r["symmetry"] == sym[i] and not r["cacheHit"]
    again = ev.evaluate(stones[0], nextPla[0], moves[0], numTurns[0], includeOwnerMap=True)
    assert again["cacheHit"] and (again["policyProbs"] == res[0]["policyProbs"]).all()
    for x in (ev, games, h, lm):
        x.close()
