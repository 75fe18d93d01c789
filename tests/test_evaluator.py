"""CPU checks of the evaluator front end (kc_evaluator_*): the lock-free submit path, batch closing,
staging-ring recycling, the inline cache, the owner-map upgrade and the error path, driven through
`kc_evaluator_create_custom` with the CPU oracle as the batch function (rules -> V1 planes -> net -> post-processing,
nneval.cpp:588-815).  The device backend of the same front end is covered by tests/test_z_gpu_evaluator.py."""
import threading

import numpy as np
import pytest

W, H, K = 5, 5, 4
HW = W * H


def make_positions(oracle, n, seed, min_plies=0, max_plies=14, dims=None):
    """Positions reached by the counter-RNG playouts of SURVEY.md 8(d): stones, player to move, last five moves, numTurns."""
    W, H, K = dims or (5, 5, 4)
    HW = W * H
    out = []
    g = 0
    og = oracle.Game(W, H, K)
    while len(out) < n:
        og.reset()
        stones = np.zeros(HW, np.int8)
        hist = []
        plies = min_plies + (g * 7 + seed) % (max_plies - min_plies + 1)
        for _ in range(plies):
            pos = og.choose(seed, g)
            pla = og.next_pla()
            if pos < 0 or not og.play(pos):
                break
            stones[pos % HW] = pla
            hist.append((pos, pla))
            if og.finished():
                break
        g += 1
        if og.finished() or og.legal_mask()[1] == 0:
            continue
        moves = np.full((5, 2), -1, np.int16)
        last = hist[-5:]
        for j, m in enumerate(last):
            moves[5 - len(last) + j] = m
        out.append({"stones": stones.copy(), "nextPla": og.next_pla(), "moves": moves, "numTurns": len(hist), "hist": list(hist)})
    return out


def oracle_game_of(oracle, stones, next_pla, hist, num_turns, dims=None):
    W, H, K = dims or (5, 5, 4)
    HW = W * H
    og = oracle.Game(W, H, K)
    for c in range(HW):
        if stones[c]:
            og.set_stone(c % W, c // W, int(stones[c]))
    og.set_history([(int(p), int(pl)) for p, pl in hist], int(num_turns), int(next_pla))
    return og


def expected_output(oracle, omodel, p, sym, temp=1.0, dims=None):
    """NNEvaluator::evaluate of one position with the oracle: planes -> net under `sym` -> post-processing."""
    W, H, K = dims or (5, 5, 4)
    og = oracle_game_of(oracle, p["stones"], p["nextPla"], p["hist"][-5:], p["numTurns"], dims)
    row, glob = og.fill_row_v1()
    pol, val, misc, own = omodel.forward(row[None], glob[None], W, H, symmetry=np.array([sym], np.int8))
    legal, n = og.legal_mask()
    assert n > 0
    ep, ev, em = oracle.postprocess(pol[0], legal, val[0], misc[0], p["nextPla"], temp)
    sign = 1.0 if p["nextPla"] == 2 else -1.0
    return {"policy": ep, "winLoss": ev, "misc": em, "owner": sign * np.tanh(own[0]), "nnHash": og.nn_hash(temp=temp)}


class OracleBackend:
    """The batch function: unpacks the rows the front end staged and evaluates them with the oracle."""

    def __init__(self, oracle, omodel, fail_on_batch=None, dims=None):
        self.oracle, self.omodel = oracle, omodel
        self.dims = dims or (5, 5, 4)
        self.batch_sizes = []
        self.servers = set()
        self.fail_on_batch = fail_on_batch
        self.lock = threading.Lock()

    def __call__(self, server, b):
        from katacoffee_b200 import backend
        with self.lock:
            self.batch_sizes.append(b.n)
            self.servers.add(server)
            idx = len(self.batch_sizes) - 1
        if self.fail_on_batch is not None and idx == self.fail_on_batch:
            return 7
        W, H, K = self.dims
        HW = W * H
        n = b.n
        pol = np.ctypeslib.as_array(b.policyProbs, shape=(n, 4 * HW))
        wl = np.ctypeslib.as_array(b.whiteWinLoss, shape=(n, 2))
        mo = np.ctypeslib.as_array(b.miscOut, shape=(n, 2))
        ow = np.ctypeslib.as_array(b.ownership, shape=(n, HW))
        for i in range(n):
            stones, pla, moves, nt, last_dir = backend.evalUnpackPosition(W, H, b.black[i], b.white[i], b.misc[i])
            hist = [(int(c), int(pl)) for c, pl in moves if pl]
            if hist:
                hist[-1] = (last_dir * HW + hist[-1][0], hist[-1][1])   # only the last move's direction is kept (and needed)
            og = oracle_game_of(self.oracle, stones, pla, hist, nt, self.dims)
            sh = og.sit_hash(pla)
            zp = backend.zobristTables()[1]
            assert int(b.hash0[i]) ^ int(zp[pla][0]) == int(sh[0]) and int(b.hash1[i]) ^ int(zp[pla][1]) == int(sh[1]), "packed pos_hash"
            row, glob = og.fill_row_v1()
            p, v, m, o = self.omodel.forward(row[None], glob[None], W, H, symmetry=np.array([b.symmetry[i]], np.int8))
            legal, _ = og.legal_mask()
            ep, ev, em = self.oracle.postprocess(p[0], legal, v[0], m[0], pla, b.policyTemperature)
            pol[i], wl[i], mo[i] = ep, ev, em
            if b.wantOwnership:
                ow[i] = o[0]
        return 0


@pytest.fixture(scope="module")
def omodel(oracle, built_lib):
    from katacoffee_b200 import modeldesc
    return oracle.Model(modeldesc.Model("b2c32", seed=4))


def check_result(r, e, owner=False):
    assert np.abs(r["policyProbs"] - e["policy"]).max() < 1e-6
    assert abs(r["whiteWinProb"] - e["winLoss"][0]) < 1e-6 and abs(r["whiteLossProb"] - e["winLoss"][1]) < 1e-6
    assert np.allclose([r["varTimeLeft"], r["shorttermWinlossError"]], e["misc"], rtol=1e-6, atol=1e-7)
    assert r["nnHash"] == (int(e["nnHash"][0]), int(e["nnHash"][1]))
    if owner:
        assert np.abs(r["whiteOwnerMap"] - e["owner"]).max() < 1e-6


def test_position_hash_and_cache_key(oracle, built_lib):
    """kc_eval_position_hash: literal NNInputs::getHash (nninputs.cpp:463-502, incl. the policy-temperature fold) and a cache
    key that separates what the literal hash confuses (ledger 8.1-E)."""
    from katacoffee_b200 import backend
    ps = make_positions(oracle, 60, seed=5)
    keys = set()
    for p in ps:
        og = oracle_game_of(oracle, p["stones"], p["nextPla"], p["hist"][-5:], p["numTurns"])
        for temp in (1.0, 0.7, 1.25):
            h, k = backend.evalPositionHash(W, H, p["stones"], p["nextPla"], p["moves"], p["numTurns"], temp)
            e = og.nn_hash(temp=temp)
            assert h == (int(e[0]), int(e[1])), temp
        keys.add(backend.evalPositionHash(W, H, p["stones"], p["nextPla"], p["moves"], p["numTurns"])[1])
    assert len(keys) == len({(p["stones"].tobytes(), p["nextPla"], p["moves"][:, 0].tobytes()) for p in ps})
    # same stones and player, different last direction: same literal hash, different legality, different key
    p = next(q for q in ps if q["numTurns"] >= 2)
    m2 = p["moves"].copy()
    m2[4, 0] = (m2[4, 0] + HW) % (4 * HW)
    h1, k1 = backend.evalPositionHash(W, H, p["stones"], p["nextPla"], p["moves"], p["numTurns"])
    h2, k2 = backend.evalPositionHash(W, H, p["stones"], p["nextPla"], m2, p["numTurns"])
    assert h1 == h2 and k1 != k2
    # numTurns: never in the literal hash; in the cache key as min(numTurns, 5), because the history planes 7..10 are gated on
    # numTurns >= 2..5 (nninputs.cpp:575-620) -- equal moves with numTurns 1 and 5 are different inputs
    q = next(q for q in ps if q["numTurns"] >= 5)
    hq, kq = backend.evalPositionHash(W, H, q["stones"], q["nextPla"], q["moves"], q["numTurns"])
    assert backend.evalPositionHash(W, H, q["stones"], q["nextPla"], q["moves"], q["numTurns"] + 2) == (hq, kq)
    for nt in (1, 2, 3, 4):
        hn, kn = backend.evalPositionHash(W, H, q["stones"], q["nextPla"], q["moves"], nt)
        assert hn == hq and kn != kq, nt
    # pinned to the values SURVEY.md 8(c) derived with the reference's own md5.cpp / sha2.cpp: the empty boards' hashes are
    # SIZE_X[n] ^ SIZE_Y[n] ^ ZOBRIST_PLAYER_HASH[pla]
    PLAYER = {1: (0xc535f97fd0cc7e76, 0x8a2a2a2ff24dbb6d), 2: (0x392045e5c8d9bd73, 0xd3c1c132e034dcb0)}
    SIZE = {5: ((0x478fc9704f6fb627, 0xb1ea08b4f90dfbf6), (0x8ae2f36ea4707544, 0x4a0b61a306b2d937)),
            6: ((0xcbfa3568128cb44f, 0x7e0be6bb36f9ec98), (0x30cb3b6e8ff75c22, 0xc24be82a513dfb9e))}
    for n, (sx, sy) in SIZE.items():
        for pla, ph in PLAYER.items():
            hh, _ = backend.evalPositionHash(n, n, np.zeros(n * n, np.int8), pla)
            assert hh == (sx[0] ^ sy[0] ^ ph[0], sx[1] ^ sy[1] ^ ph[1]), (n, pla)
    # round trip of the packing
    st, pla, mv, nt, ld = backend.evalUnpackPosition(W, H, 0b100001, 0b10, (3 | (1 << 6)) | (2 << 40) | (9 << 48) | ((2 << 3) << 56))
    assert st[0] == 1 and st[5 - 0] == 0 and st[1] == 2 and pla == 2 and nt == 9 and ld == 2 and tuple(mv[4]) == (3, 1) and mv[3][1] == 0


def test_cache_separates_equal_moves_with_different_num_turns(oracle, built_lib, omodel):
    """Two requests with the same stones, player and last five moves but numTurns 1 and 5 have different V1 planes (history planes
    7..10 need numTurns >= 2..5): each must get its own evaluation, also when the other one is already in the cache."""
    from katacoffee_b200 import backend
    p = next(q for q in make_positions(oracle, 40, seed=11, min_plies=6) if q["numTurns"] >= 5)
    be = OracleBackend(oracle, omodel)
    ev = backend.NNEvaluator(nnXLen=W, nnYLen=H, winLen=K, maxBatchSize=8, maxConcurrentEvals=16, numThreads=1,
                             nnCacheSizePowerOfTwo=10, customBackend=be)
    outs = {}
    for nt in (p["numTurns"], 1, p["numTurns"], 1):
        r = ev.evaluate(p["stones"], p["nextPla"], p["moves"], nt, symmetry=0)
        e = expected_output(oracle, omodel, dict(p, numTurns=nt), 0)
        check_result(r, e)
        outs.setdefault(nt, []).append(r)
    assert not outs[1][0]["cacheHit"] and outs[1][1]["cacheHit"] and outs[p["numTurns"]][1]["cacheHit"]
    assert np.abs(outs[1][0]["policyProbs"] - outs[p["numTurns"]][0]["policyProbs"]).max() > 1e-6   # the inputs really differ
    ev.close()


@pytest.mark.timeout(300)
def test_threads_batches_cache_against_oracle(oracle, built_lib, omodel):
    """8 client threads x 30 evaluations over 90 distinct positions, 2 server threads, batches of at most 8 rows: every result
    equals the oracle's NNEvaluator::evaluate; rows + hits account for every request; batches never exceed maxBatch."""
    from katacoffee_b200 import backend
    ps = make_positions(oracle, 90, seed=11)
    be = OracleBackend(oracle, omodel)
    ev = backend.NNEvaluator(nnXLen=W, nnYLen=H, winLen=K, maxBatchSize=8, maxConcurrentEvals=16, numThreads=2, nnCacheSizePowerOfTwo=12,
                             nnMutexPoolSizePowerofTwo=4, defaultSymmetry=3, customBackend=be)
    exp = [expected_output(oracle, omodel, p, 3) for p in ps]
    errors = []

    def client(t):
        try:
            rng = np.random.default_rng(t)
            for _ in range(30):
                i = int(rng.integers(0, len(ps)))
                p = ps[i]
                r = ev.evaluate(p["stones"], p["nextPla"], p["moves"], p["numTurns"])
                check_result(r, exp[i])
                assert r["symmetry"] == 3
        except BaseException as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=client, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]
    st = ev.stats()
    assert st["cacheHits"] + st["cacheMisses"] == 240 and st["cacheHits"] > 0
    assert st["rowsProcessed"] == st["cacheMisses"] == sum(be.batch_sizes)
    assert st["batchesProcessed"] == len(be.batch_sizes) and max(be.batch_sizes) <= 8
    assert abs(ev.averageProcessedBatchSize() - np.mean(be.batch_sizes)) < 1e-9
    # an explicit symmetry is honoured (skipCache: the row is evaluated again, nneval.cpp:611) and stored
    r = ev.evaluate(ps[0]["stones"], ps[0]["nextPla"], ps[0]["moves"], ps[0]["numTurns"], symmetry=6, skipCache=True)
    check_result(r, expected_output(oracle, omodel, ps[0], 6))
    r2 = ev.evaluate(ps[0]["stones"], ps[0]["nextPla"], ps[0]["moves"], ps[0]["numTurns"])
    assert r2["cacheHit"] and r2["symmetry"] == 6 and (r2["policyProbs"] == r["policyProbs"]).all()
    # clearCache / clearStats
    ev.clearCache()
    ev.clearStats()
    r3 = ev.evaluate(ps[0]["stones"], ps[0]["nextPla"], ps[0]["moves"], ps[0]["numTurns"])
    assert not r3["cacheHit"] and ev.stats()["rowsProcessed"] == 1 and ev.stats()["cacheMisses"] == 1
    ev.close()


@pytest.mark.timeout(300)
def test_owner_map_upgrade_and_randomized_symmetry(oracle, built_lib, omodel):
    """nneval.cpp:612-623, 689-701: a cached result without owner map is re-evaluated for the map only, its policy and
    values are kept; nneval.cpp:817-838: tanh and flip to white.  doRandomize draws the symmetry from (seed, position)."""
    from katacoffee_b200 import backend
    ps = make_positions(oracle, 24, seed=3, min_plies=2)
    be = OracleBackend(oracle, omodel)
    ev = backend.NNEvaluator(nnXLen=W, nnYLen=H, winLen=K, maxBatchSize=4, numThreads=1, nnCacheSizePowerOfTwo=10, doRandomize=True, randSeed=77,
                             customBackend=be)
    syms = set()
    for p in ps:
        a = ev.evaluate(p["stones"], p["nextPla"], p["moves"], p["numTurns"])
        assert not a["cacheHit"] and a["whiteOwnerMap"] is None
        syms.add(a["symmetry"])
        e = expected_output(oracle, omodel, p, a["symmetry"])
        check_result(a, e)
        b = ev.evaluate(p["stones"], p["nextPla"], p["moves"], p["numTurns"], includeOwnerMap=True)
        assert not b["cacheHit"] and (b["policyProbs"] == a["policyProbs"]).all() and b["whiteWinProb"] == a["whiteWinProb"]
        check_result(b, e, owner=True)
        c = ev.evaluate(p["stones"], p["nextPla"], p["moves"], p["numTurns"], includeOwnerMap=True)
        assert c["cacheHit"] and (c["whiteOwnerMap"] == b["whiteOwnerMap"]).all() and c["symmetry"] == a["symmetry"]
        d = ev.evaluate(p["stones"], p["nextPla"], p["moves"], p["numTurns"], skipCache=True)
        assert d["symmetry"] == a["symmetry"], "a position is always evaluated under the same symmetry"
    st = ev.stats()
    assert st["ownerMapUpgrades"] == len(ps) and st["cacheHits"] == len(ps) and st["cacheMisses"] == len(ps)
    assert st["rowsProcessed"] == 3 * len(ps)
    assert len(syms) >= 4, syms
    ev.close()


@pytest.mark.timeout(300)
def test_evaluate_many_with_a_full_ring(oracle, built_lib, omodel):
    """Three threads push 150 rows each through an 8-buffer ring of 4-row batches (far more than maxConcurrentEvals in
    flight): submits wait for recycled buffers, nothing deadlocks, every result is right; policy temperature 0.7."""
    from katacoffee_b200 import backend
    ps = make_positions(oracle, 150, seed=21)
    be = OracleBackend(oracle, omodel)
    ev = backend.NNEvaluator(nnXLen=W, nnYLen=H, winLen=K, maxBatchSize=4, maxConcurrentEvals=4, numThreads=2, nnCacheSizePowerOfTwo=-1,
                             defaultSymmetry=5, nnPolicyTemperature=0.7, customBackend=be)
    exp = [expected_output(oracle, omodel, p, 5, temp=0.7) for p in ps]
    stones = np.stack([p["stones"] for p in ps])
    nextPla = [p["nextPla"] for p in ps]
    moves = np.stack([p["moves"] for p in ps])
    numTurns = [p["numTurns"] for p in ps]
    errors = []

    def client(t):
        try:
            order = np.random.default_rng(t).permutation(len(ps))
            res = ev.evaluateMany(stones[order], [nextPla[i] for i in order], moves[order], [numTurns[i] for i in order], includeOwnerMap=(t == 1))
            for r, i in zip(res, order):
                check_result(r, exp[i], owner=(t == 1))
        except BaseException as e:   # noqa: BLE001
            errors.append(e)

    threads = [threading.Thread(target=client, args=(t,)) for t in range(3)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[0]
    st = ev.stats()
    assert st["rowsProcessed"] == 450 and st["cacheHits"] == 0 and st["cacheMisses"] == 0   # no cache: nothing counted
    assert st["backpressureWaits"] > 0
    assert max(be.batch_sizes) <= 4 and be.servers == {0, 1}
    ev.close()


@pytest.mark.timeout(120)
def test_backend_failure_and_argument_errors(oracle, built_lib, omodel):
    from katacoffee_b200 import backend, capi
    ps = make_positions(oracle, 6, seed=2)
    be = OracleBackend(oracle, omodel, fail_on_batch=1)
    ev = backend.NNEvaluator(nnXLen=W, nnYLen=H, winLen=K, maxBatchSize=4, numThreads=1, customBackend=be)
    p = ps[0]
    ev.evaluate(p["stones"], p["nextPla"], p["moves"], p["numTurns"])
    with pytest.raises(capi.KCError, match="backend failed"):
        ev.evaluate(ps[1]["stones"], ps[1]["nextPla"], ps[1]["moves"], ps[1]["numTurns"])
    # the failed row was not cached and the evaluator keeps serving
    r = ev.evaluate(ps[1]["stones"], ps[1]["nextPla"], ps[1]["moves"], ps[1]["numTurns"])
    assert not r["cacheHit"]
    check_result(r, expected_output(oracle, omodel, ps[1], 0))
    with pytest.raises(capi.KCError, match="symmetry"):
        ev.evaluate(p["stones"], p["nextPla"], p["moves"], p["numTurns"], symmetry=8)
    with pytest.raises(capi.KCError, match="nextPla"):
        ev.evaluate(p["stones"], 3, p["moves"], p["numTurns"])
    bad = p["stones"].copy()
    bad[0] = 5
    with pytest.raises(capi.KCError, match="stone colour"):
        ev.evaluate(bad, p["nextPla"], p["moves"], p["numTurns"])
    ev.close()
    for kw, msg in (({"maxBatchSize": 0}, "maxBatchSize"), ({"numThreads": 0}, "numServerThreads"), ({"nnXLen": 11}, "board size"),
                    ({"defaultSymmetry": 8}, "defaultSymmetry"), ({"nnPolicyTemperature": 0.0}, "policyTemperature"),
                    ({"nnCacheSizePowerOfTwo": 40}, "cacheSizePowerOfTwo")):
        with pytest.raises(capi.KCError, match=msg):
            backend.NNEvaluator(**{"nnXLen": W, "nnYLen": H, "winLen": K, "maxBatchSize": 4, "customBackend": be, **kw})
    # without a device the real backend must refuse (no CPU fallback)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(capi.KCError):
            backend.createComputeContext(0)


@pytest.mark.timeout(600)
@pytest.mark.parametrize("tsan", [False, True])
def test_cpp_stress(built_lib, tmp_path, tsan):
    """tests/cpp/test_evaluator_stress.cpp: 12 native client threads (single rows and evaluate_many chunks, owner maps,
    skipCache) against 3 servers, an 8-buffer ring and a 512-entry cache; every result is recomputed and compared bit for
    bit.  Second build: the queue and cache compiled alone under -fsanitize=thread (no data race reported)."""
    import json
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "cpp", "test_evaluator_stress.cpp")
    exe = str(tmp_path / ("stress_tsan" if tsan else "stress"))
    inc = ["-I" + os.path.join(root, "include")]
    if tsan:
        csrc = os.path.join(root, "katacoffee_b200", "csrc")
        cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-DKC_EVALUATOR_HOST_ONLY", "-w"] + inc + ["-I" + csrc, "-I/usr/local/cuda/include",
               "-x", "c++", src, os.path.join(csrc, "evaluator.cpp"), os.path.join(csrc, "zobrist.cpp"), "-o", exe, "-lpthread"]
    else:
        libdir = os.path.join(root, "katacoffee_b200")
        cmd = ["g++", "-std=c++17", "-O2", "-Wall"] + inc + [src, "-o", exe, "-L" + libdir, "-lkatacoffee_b200", "-Wl,-rpath," + libdir, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if tsan and r.returncode != 0 and "tsan" in r.stderr:
        pytest.skip("libtsan is not installed")
    assert r.returncode == 0, r.stderr
    args = ["12", "6000", "1500"] if tsan else ["12", "20000", "3000"]
    r = subprocess.run([exe] + args, capture_output=True, text=True, timeout=500, env={**os.environ, "TSAN_OPTIONS": "halt_on_error=1"})
    assert r.returncode == 0, r.stdout + r.stderr
    assert "ThreadSanitizer" not in r.stderr, r.stderr
    rep = json.loads(r.stdout.strip().splitlines()[-1])
    assert rep["mismatches"] == 0 and rep["accounting"] and rep["cacheHits"] > 0 and rep["upgrades"] > 0 and rep["backpressureWaits"] > 0


@pytest.mark.timeout(300)
def test_cpp_nnevaluator_class(built_lib, tmp_path):
    """host/b200nneval.cpp: class NNEvaluator with the reference's interface (nneval.h:80-175) over the front end, driven by
    tests/cpp/test_b200nneval.cpp like the reference's search threads drive it (constructor and size-check errors with the
    reference's wording, four threads, owner maps, cache / clearCache, run-time symmetry settings, a second policy temperature,
    kill / setNumThreads / respawn); the batch function is a pure function of the staged row."""
    import os
    import subprocess
    from katacoffee_b200 import backend, modeldesc
    from katacoffee_b200 import build as kb
    kb.build_host()
    exe = os.path.join(os.path.dirname(kb.HERE), "katacoffee_b200", "host", "test_b200nneval")
    path = str(tmp_path / "b2c32.bin.gz")
    backend.writeModelFile(modeldesc.Model("b2c32", seed=4), path)
    r = subprocess.run([exe, path, "cpu"], capture_output=True, text=True, timeout=200)
    assert r.returncode == 0 and "test_b200nneval cpu: ok" in r.stdout, r.stdout + r.stderr
    # the same driver with the class, the shim and the front end compiled under -fsanitize=thread
    root = os.path.dirname(kb.HERE)
    host, csrc, libdir = os.path.join(kb.HERE, "host"), os.path.join(kb.HERE, "csrc"), kb.HERE
    tsan_exe = str(tmp_path / "nneval_tsan")
    cmd = ["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-fopenmp", "-DKC_EVALUATOR_HOST_ONLY", "-w", "-I" + os.path.join(root, "include"), "-I" + host,
           "-I" + csrc, "-I/usr/local/cuda/include", os.path.join(root, "tests", "cpp", "test_b200nneval.cpp"), os.path.join(host, "b200nneval.cpp"),
           os.path.join(host, "b200backend.cpp"), "-x", "c++", os.path.join(csrc, "evaluator.cpp"), os.path.join(csrc, "zobrist.cpp"), "-o", tsan_exe,
           "-L" + libdir, "-lkatacoffee_b200", "-Wl,-rpath," + libdir, "-lpthread"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0 and "tsan" in r.stderr:
        pytest.skip("libtsan is not installed")
    assert r.returncode == 0, r.stderr
    r = subprocess.run([tsan_exe, path, "cpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "ThreadSanitizer" not in r.stderr and "test_b200nneval cpu: ok" in r.stdout, r.stdout + r.stderr[-2000:]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("dims", [(6, 6, 4), (4, 3, 3), (7, 7, 5)])
def test_other_board_sizes(oracle, built_lib, omodel, dims):
    """6x6 k=4 (BASELINE configs[4]), a non-square board (the transposing symmetries are then flips only, nninputs.cpp:252-357) and the
    largest device board: packing, hashes (SIZE_X / SIZE_Y Zobrist terms), symmetries and results against the oracle pipeline."""
    from katacoffee_b200 import backend
    Wd, Hd, Kd = dims
    ps = make_positions(oracle, 40, seed=9, max_plies=Wd * Hd // 2, dims=dims)
    be = OracleBackend(oracle, omodel, dims=dims)
    ev = backend.NNEvaluator(nnXLen=Wd, nnYLen=Hd, winLen=Kd, maxBatchSize=8, numThreads=2, nnCacheSizePowerOfTwo=10, customBackend=be)
    res = ev.evaluateMany(np.stack([p["stones"] for p in ps]), [p["nextPla"] for p in ps], np.stack([p["moves"] for p in ps]),
                          [p["numTurns"] for p in ps], symmetry=np.arange(len(ps)) % 8, includeOwnerMap=True)
    seen = set()
    for i, (p, r) in enumerate(zip(ps, res)):
        og = oracle_game_of(oracle, p["stones"], p["nextPla"], p["hist"][-5:], p["numTurns"], dims)
        h, k = backend.evalPositionHash(Wd, Hd, p["stones"], p["nextPla"], p["moves"], p["numTurns"])
        assert h == tuple(int(x) for x in og.nn_hash()) == r["nnHash"]
        if k in seen:
            assert r["cacheHit"] or True   # a repeated position may have been served from the cache under another symmetry
            continue
        seen.add(k)
        sym = r["symmetry"]
        e = expected_output(oracle, omodel, p, sym, dims=dims)
        assert np.abs(r["policyProbs"] - e["policy"]).max() < 1e-6 and np.abs(r["whiteOwnerMap"] - e["owner"]).max() < 1e-6
        assert abs(r["whiteWinProb"] - e["winLoss"][0]) < 1e-6
        assert (r["policyProbs"] >= 0).sum() == og.legal_mask()[1]
    ev.close()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("dims", [(2, 2, 2), (3, 7, 3), (5, 5, 4), (7, 7, 4), (7, 4, 4), (8, 8, 4), (10, 10, 5), (9, 3, 3), (2, 10, 2)])
def test_packing_round_trip_on_arbitrary_positions(oracle, built_lib, dims):
    """Random stone placements and histories (not only reachable ones), every supported shape: what the front end stages unpacks to what
    was submitted, its pos_hash / NNInputs::getHash are the oracle's, and the cache key separates any two different (stones, player,
    last five (cell, player), last direction) tuples."""
    from katacoffee_b200 import backend
    Wd, Hd, Kd = dims
    HWd = Wd * Hd
    rng = np.random.default_rng(Wd * 100 + Hd)
    n = 120
    stones = rng.integers(0, 3, (n, HWd)).astype(np.int8)
    nextPla = rng.integers(1, 3, n).astype(np.int8)
    moves = np.full((n, 5, 2), -1, np.int16)
    numTurns = rng.integers(0, 200, n).astype(np.int32)
    for i in range(n):
        k = int(rng.integers(0, 6))
        for j in range(5 - k, 5):
            moves[i, j] = (int(rng.integers(0, 4 * HWd)), int(rng.integers(1, 3)))
    staged = {}

    def record(server, b):
        for i in range(b.n):
            if Wd > 7 or Hd > 7:   # 128-bit boards, the wide misc format (games_big.cuh)
                assert b.blackHi and b.whiteHi
                st, pla, mv, nt, ld = backend.evalUnpackPositionWide(Wd, Hd, b.black[i], b.blackHi[i], b.white[i], b.whiteHi[i], b.misc[i])
            else:
                assert not b.blackHi and not b.whiteHi
                st, pla, mv, nt, ld = backend.evalUnpackPosition(Wd, Hd, b.black[i], b.white[i], b.misc[i])
            staged[(st.tobytes(), pla, mv.tobytes(), nt, ld)] = (int(b.hash0[i]), int(b.hash1[i]))
            np.ctypeslib.as_array(b.policyProbs, shape=(b.n, 4 * HWd))[i] = 0.0
        return 0

    ev = backend.NNEvaluator(nnXLen=Wd, nnYLen=Hd, winLen=Kd, maxBatchSize=16, numThreads=1, nnCacheSizePowerOfTwo=-1, customBackend=record)
    res = ev.evaluateMany(stones, nextPla, moves, numTurns)
    ev.close()
    zp = backend.zobristTables()[1]
    keys = {}
    for i in range(n):
        cells = np.array([[m[0] % HWd if m[0] >= 0 else -1, m[1] if m[0] >= 0 else 0] for m in moves[i]], np.int16)
        last_dir = int(moves[i, 4, 0]) // HWd if moves[i, 4, 0] >= 0 else 4
        key = (stones[i].tobytes(), int(nextPla[i]), cells.tobytes(), int(numTurns[i]), last_dir)
        assert key in staged, i
        og = oracle.Game(Wd, Hd, Kd)
        for c in range(HWd):
            if stones[i, c]:
                og.set_stone(c % Wd, c // Wd, int(stones[i, c]))
        hist = [(int(m[0]), int(m[1])) for m in moves[i] if m[0] >= 0]
        og.set_history(hist, int(numTurns[i]), int(nextPla[i]))
        sh = og.sit_hash(int(nextPla[i]))
        h0, h1 = staged[key]
        assert (h0 ^ int(zp[nextPla[i]][0]), h1 ^ int(zp[nextPla[i]][1])) == (int(sh[0]), int(sh[1]))
        assert res[i]["nnHash"] == tuple(int(x) for x in og.nn_hash())
        ck = backend.evalPositionHash(Wd, Hd, stones[i], nextPla[i], moves[i], numTurns[i])[1]
        ident = (stones[i].tobytes(), int(nextPla[i]), cells.tobytes(), last_dir)
        assert keys.setdefault(ck, ident) == ident
    assert len(keys) == len({v for v in keys.values()})
