"""CPU checks of the oracle's tree search and training rows (test infrastructure for the device search; no GPU)."""
import numpy as np
import pytest


def _selfplay(oracle, W, H, K, V, seed, gid, T=4):
    og = oracle.Game(W, H, K)
    moves, rn, rw, vis = [], [], [], []
    while not og.finished():
        r = oracle.search_run(og, V)
        assert r["rootVisits"] == V and r["edgeVisits"].sum() == V - 1          # one visit is the root's own evaluation
        assert abs(r["rootUtilitySum"]) <= V + 1e-9
        legal = og.legal_mask()[0]
        for pos in np.nonzero(r["edgeVisits"])[0]:
            assert (legal[pos >> 5] >> (pos & 31)) & 1                         # only legal moves are ever visited
        assert (r["policy"][r["policy"] >= 0].sum() - 1.0) < 1e-5
        mv = oracle.search_choose(r["edgeVisits"], r["order"], og.num_turns(), T, seed, gid)
        assert r["edgeVisits"][mv] > 0
        moves.append(mv); rn.append(r["rootVisits"]); rw.append(r["rootUtilitySum"])
        vis.append(np.where(r["order"] != 255, r["edgeVisits"], 0).astype(np.int16))
        og.play(mv)
    return og, moves, rn, rw, np.stack(vis)


@pytest.mark.parametrize("W,H,K", [(5, 5, 4), (4, 5, 3)])
def test_oracle_search_and_training_rows(oracle, W, H, K):
    V = 48
    og, moves, rn, rw, vis = _selfplay(oracle, W, H, K, V, seed=3, gid=7)
    rows = oracle.training_rows(W, H, K, moves, rn, rw, vis, 7)
    R, HW = len(moves), W * H
    bits = np.unpackbits(rows["binaryInputNCHWPacked"], axis=2)[:, :, :HW]
    replay = oracle.Game(W, H, K)
    for i in range(R):
        planes = replay.fill_row_v1()[0].reshape(15, HW)
        assert (bits[i] == planes).all()                                        # packBits is numpy's big-endian packbits
        assert rows["globalTargetsNC"][i, 51] == i and rows["globalTargetsNC"][i, 60] == V
        assert (rows["policyTargetsNCMove"][i, 0] == vis[i]).all()
        replay.play(moves[i])
    gt = rows["globalTargetsNC"]
    assert np.allclose(gt[:, 0:10:2] + gt[:, 1:10:2], 1.0, atol=1e-6)           # every td target is a win/loss pair
    winner = og.winner()
    for i in range(R):
        pla = 1 + (i % 2)                                                       # black moves first
        win = 0.5 if winner == 0 else float(winner == pla)
        assert gt[i, 0] == win and gt[i, 1] == 1.0 - win                       # nowFactor 0: the game result
    assert gt[R - 1, 28] == 0 and (gt[:R - 1, 28] == 1).all() and (rows["policyTargetsNCMove"][R - 1, 1] == 1).all()
    # final ownership is the final stones from the mover's view; the run-length plane marks the winning line
    fin = np.array([[og._color_at(x, y) if hasattr(og, "_color_at") else oracle.lib().ko_game_color_at(og._g, x, y) for x in range(W)] for y in range(H)])
    assert (np.abs(rows["valueTargetsNCHW"][:, 0]) == (fin != 0)).all()
    if winner != 0:
        assert rows["valueTargetsNCHW"][0, 4].max() >= K
    else:
        assert rows["valueTargetsNCHW"][0, 4].max() < K


def _midgame(oracle, W, H, K, seed, gid, plies):
    og = oracle.Game(W, H, K)
    for _ in range(plies):
        if og.finished():
            break
        og.play(og.choose(seed, gid))
    return og


def test_oracle_graph_search_invariants(oracle):
    """Graph mode of the oracle search: without transposition table and bias it visits exactly what the tree search visits
    (same selection rule on the same statistics, utilities of the hash evaluator are dyadic so even the sums agree);
    with the table, every visit is an evaluation, a terminal visit, a transposition hit or a catch-up visit."""
    V = 300
    hits = 0
    for gid in range(12):
        og = _midgame(oracle, 5, 5, 4, 17, gid, 5 + gid % 8)
        if og.finished():
            continue
        a = oracle.search_run(og, V)
        b = oracle.search_run_graph(og, V, graph=False)
        assert (a["edgeVisits"] == b["edgeVisits"]).all() and (a["order"] == b["order"]).all()
        assert abs(a["rootUtilitySum"] - b["rootUtilitySum"]) < 1e-9
        c = oracle.search_run_graph(og, V, graph=True)
        assert c["rootVisits"] == V and int(c["counters"][0]) == V == int(c["counters"][1:].sum())
        assert c["edgeVisits"].sum() == V - 1
        hits += int(c["counters"][3])
        d = oracle.search_run_graph(og, V, graph=True, bias_factor=0.3, bias_exponent=0.8)
        assert d["rootVisits"] == V and abs(d["rootUtilitySum"]) <= V
        assert d["digest"] != c["digest"]
    assert hits > 0


def test_oracle_det_pow_close_to_libm(oracle):
    """The bias weight W^exponent uses a pow built from IEEE basic operations (so the device can reproduce its bits);
    indirectly: exponent 0.5 through sqrt and through exp(0.5 log) give searches that agree to rounding noise."""
    og = _midgame(oracle, 5, 5, 4, 3, 1, 6)
    a = oracle.search_run_graph(og, 200, graph=True, bias_factor=0.3, bias_exponent=0.5)
    b = oracle.search_run_graph(og, 200, graph=True, bias_factor=0.3, bias_exponent=0.5000000001)
    assert abs(a["rootUtilitySum"] - b["rootUtilitySum"]) < 1e-6 * 200


def test_literal_graph_hash_is_path_dependent(oracle):
    """GraphHash::getGraphHash as written (graphhash.cpp:14-29) chains the previous hash after every move, so two move orders
    reaching the same stones / player / last move get different hashes -- why the search keys on the state instead."""
    import ctypes as C
    def chain(moves):
        og = oracle.Game(5, 5, 4)
        h = np.zeros(2, np.uint64)
        for mv in moves:
            nh = np.zeros(2, np.uint64)
            oracle.lib().ko_graph_hash(h.ctypes.data_as(C.c_void_p), og._g, og.next_pla(), nh.ctypes.data_as(C.c_void_p))
            h = nh
            assert og.play(mv)
        nh = np.zeros(2, np.uint64)
        oracle.lib().ko_graph_hash(h.ctypes.data_as(C.c_void_p), og._g, og.next_pla(), nh.ctypes.data_as(C.c_void_p))
        return og, nh
    HW = 25
    # black a, white b, black c, white d  vs  black c, white b, black a, white d: all on row 0 with direction W-E (dir 1)
    p = lambda x, y, d: d * HW + y * 5 + x
    g1, h1 = chain([p(0, 0, 1), p(1, 0, 1), p(2, 0, 1), p(4, 0, 1)])
    g2, h2 = chain([p(2, 0, 1), p(1, 0, 1), p(0, 0, 1), p(4, 0, 1)])
    assert tuple(g1.sit_hash()) == tuple(g2.sit_hash())
    assert tuple(h1) != tuple(h2)


SELFPLAY_OPTS = dict(rootNoiseEnabled=1, rootDirichletNoiseTotalConcentration=10.83, rootDirichletNoiseWeight=0.25, rootPolicyTemperature=1.1,
                     rootPolicyTemperatureEarly=1.25, chosenMoveTemperatureHalflife=19.0)


def test_oracle_root_noise_is_a_deterministic_dirichlet_mixture(oracle):
    """Root noise / temperature of the oracle search (the definition the device reproduces bit for bit): a probability vector over
    exactly the legal moves, a pure function of (seed, game id, ply), different for another game id, and on average
    0.75 * policy + 0.25 * alpha with alpha the shaped Dirichlet mean (searchhelpers.cpp:51-120)."""
    og = _midgame(oracle, 5, 5, 4, 3, 1, 4)
    base = oracle.search_run_graph(og, 1)["policy"]
    a = oracle.search_run_graph(og, 1, noiseSeed=5, noiseGameId=9, **SELFPLAY_OPTS)["policy"]
    b = oracle.search_run_graph(og, 1, noiseSeed=5, noiseGameId=9, **SELFPLAY_OPTS)["policy"]
    c = oracle.search_run_graph(og, 1, noiseSeed=5, noiseGameId=10, **SELFPLAY_OPTS)["policy"]
    assert (a == b).all() and (a != c).any() and (a != base).any()
    assert ((a < 0) == (base < 0)).all() and abs(float(a[a >= 0].sum()) - 1.0) < 1e-5
    legal = base >= 0
    p = base[legal].astype(np.float64)
    la = np.log(np.minimum(0.01, p) + 1e-20)
    la = np.maximum(0.0, la - la.mean())
    alpha = 0.5 * (la / la.sum() + 1.0 / legal.sum()) if la.sum() > 0 else np.full(legal.sum(), 1.0 / legal.sum())
    acc, n = np.zeros(legal.sum()), 400
    for k in range(n):
        acc += oracle.search_run_graph(og, 1, rootNoiseEnabled=1, rootDirichletNoiseTotalConcentration=10.83, rootDirichletNoiseWeight=0.25,
                                       noiseSeed=11, noiseGameId=k)["policy"][legal]
    assert np.abs(acc / n - (0.75 * p + 0.25 * alpha)).max() < 0.006


def test_oracle_move_choice_temperature_schedule(oracle):
    """chooseIndexWithTemperature restated (searchhelpers.cpp:12-49): T -> 0 picks the most visited move, prune removes rarely
    visited moves, and at T = 1 the choice frequencies follow the visit counts."""
    P = 100
    ev = np.zeros(P, np.int32); order = np.full(P, 255, np.uint8)
    ev[[3, 40, 77, 90]] = [60, 25, 14, 1]; order[[3, 40, 77, 90]] = [0, 1, 2, 3]
    assert oracle.search_choose_temperature(ev, order, 25, 0, 0.0, 0.0, 19.0, 0.0, 0.0, 1, 1) == 3
    counts = {3: 0, 40: 0, 77: 0, 90: 0}
    for gid in range(3000):
        counts[oracle.search_choose_temperature(ev, order, 25, 0, 1.0, 1.0, 19.0, 0.0, 0.0, 7, gid)] += 1
    assert abs(counts[3] / 3000 - 0.60) < 0.03 and abs(counts[40] / 3000 - 0.25) < 0.03 and 0 < counts[90] < 90
    pruned = [oracle.search_choose_temperature(ev, order, 25, 0, 1.0, 1.0, 19.0, 0.0, 1.0, 7, gid) for gid in range(3000)]
    assert 90 in pruned            # prune is min(1, max/64 = 0.94): a move with one visit survives (the reference's "< amountToPrune")
    pruned2 = [oracle.search_choose_temperature(ev * 2, order, 25, 0, 1.0, 1.0, 19.0, 0.0, 2.0, 7, gid) for gid in range(3000)]
    assert 90 in pruned2 and pruned2.count(3) > pruned2.count(40) > pruned2.count(77)
    # the schedule cools down with the turn number: late in the game the choice is nearly greedy
    late = [oracle.search_choose_temperature(ev, order, 25, 200, 0.75, 0.05, 19.0, 0.0, 0.0, 7, gid) for gid in range(500)]
    assert late.count(3) >= 495


def test_oracle_value_weighting_and_fpu_options_run(oracle):
    """valueWeightExponent / fpuParentWeightByVisitedPolicy / rootDesiredPerChildVisitsCoeff: searches complete with the visit
    budget spent and differ from the plain search."""
    og = _midgame(oracle, 5, 5, 4, 3, 2, 6)
    plain = oracle.search_run_graph(og, 300)
    for opts in (dict(valueWeightExponent=0.5), dict(fpuParentWeightByVisitedPolicy=1, fpuParentWeightByVisitedPolicyPow=2.0),
                 dict(rootDesiredPerChildVisitsCoeff=2.0)):
        r = oracle.search_run_graph(og, 300, **opts)
        assert r["rootVisits"] == 300 and r["edgeVisits"].sum() == 299
        assert r["digest"] != plain["digest"]
    wide = oracle.search_run_graph(og, 300, rootDesiredPerChildVisitsCoeff=2.0)
    assert (wide["edgeVisits"] > 0).sum() >= (plain["edgeVisits"] > 0).sum()    # the coefficient funnels visits into more root children


def test_oracle_play_selection_values_reduced_weights_and_lcb(oracle):
    """Search::getPlaySelectionValues as the oracle restates it (searchresults.cpp:66-231): without LCB the most stably explored child
    keeps its weight (= its visits here) and every other child is an integer no larger than its visits (getReducedPlaySelectionWeight,
    rounded up); with LCB at most one value differs -- the child with the best lower bound -- and it only grows; a very sharp
    bound (lcbStdevs small) makes the promotion happen, and useNonBuggyLcb = 0 never promotes the first-created child."""
    promoted = 0
    for gid in range(24):
        og = _midgame(oracle, 5, 5, 4, 11, gid, 3 + gid % 6)
        if og.finished():
            continue
        base = oracle.search_run_graph(og, 200, graph=True, cpuct=1.1, valueWeightExponent=0.5)
        psv, ev, order = base["playSelection"], base["edgeVisits"], base["order"]
        kids = order != 255
        assert (psv[~kids] == 0).all() and (psv[kids] <= ev[kids] + 1e-9).all()
        best = int(np.argmax(np.where(kids, ev, -1)))
        others = kids.copy(); others[best] = False
        assert (psv[others] == np.ceil(psv[others])).all()
        assert abs(psv[np.argmax(psv)] - ev[np.argmax(psv)]) < 1e-6          # the top value is an unreduced child weight
        for stdevs, nonbuggy in ((5.0, 1), (0.5, 1), (0.5, 0)):
            r = oracle.search_run_graph(og, 200, graph=True, cpuct=1.1, valueWeightExponent=0.5, useLcbForSelection=1, lcbStdevs=stdevs,
                                        minVisitPropForLCB=0.15, useNonBuggyLcb=nonbuggy)
            assert (r["edgeVisits"] == ev).all()                         # LCB touches the move choice, not the search
            diff = np.flatnonzero(r["playSelection"] != psv)
            assert len(diff) <= 1
            if len(diff):
                assert r["playSelection"][diff[0]] > psv[diff[0]]
                promoted += stdevs == 0.5 and nonbuggy == 1
                if not nonbuggy:
                    assert order[diff[0]] != 0
    assert promoted > 0


def test_oracle_root_symmetry_sampling_and_uncertainty(oracle):
    """rootNumSymmetriesToSample (searchnnhelpers.cpp:67-83): N extra evaluations per search, root priors that differ from a single
    evaluation, depend on (seed, game id) for N < 8 and -- all eight symmetries averaged -- only on summation order for N = 8.
    useUncertainty (searchupdatehelpers.cpp:91-113): the root's weight is no longer its visit count."""
    og = _midgame(oracle, 5, 5, 4, 3, 1, 4)
    one = oracle.search_run_graph(og, 50)
    four = oracle.search_run_graph(og, 50, rootNumSymmetriesToSample=4, noiseSeed=5, noiseGameId=9)
    again = oracle.search_run_graph(og, 50, rootNumSymmetriesToSample=4, noiseSeed=5, noiseGameId=9)
    other = oracle.search_run_graph(og, 50, rootNumSymmetriesToSample=4, noiseSeed=5, noiseGameId=10)
    assert int(four["counters"][1]) == int(one["counters"][1]) + 3 and four["rootVisits"] == one["rootVisits"] == 50
    assert (four["policy"] == again["policy"]).all() and (four["policy"] != one["policy"]).any() and (four["policy"] != other["policy"]).any()
    assert ((four["policy"] < 0) == (one["policy"] < 0)).all() and abs(float(four["policy"][four["policy"] >= 0].sum()) - 1.0) < 1e-5
    e1 = oracle.search_run_graph(og, 1, rootNumSymmetriesToSample=8, noiseSeed=5, noiseGameId=9)["policy"]
    e2 = oracle.search_run_graph(og, 1, rootNumSymmetriesToSample=8, noiseSeed=6, noiseGameId=11)["policy"]
    assert np.abs(e1 - e2).max() < 1e-6
    unc = oracle.search_run_graph(og, 50, useUncertainty=1, uncertaintyCoeff=0.25, uncertaintyExponent=1.0, uncertaintyMaxWeight=8.0)
    assert unc["rootVisits"] == 50 and unc["digest"] != one["digest"] and (unc["edgeVisits"] != one["edgeVisits"]).any()


def test_oracle_noise_pruning_changes_weights_not_visits_bookkeeping(oracle):
    """useNoisePruning (pruneNoiseWeight, searchupdatehelpers.cpp:422-470): the search still makes maxVisits visits, the graph differs
    from the unpruned one, a tighter utility scale prunes more (the root's weight drops further below its visit count), and a cap of 0
    switches the pruning off bit for bit."""
    changed = tighter = 0
    for gid in range(12):
        og = _midgame(oracle, 5, 5, 4, 3, gid, 2 + gid % 5)
        if og.finished():
            continue
        kw = dict(valueWeightExponent=0.5, noiseSeed=5, noiseGameId=gid, **SELFPLAY_OPTS)   # root noise sends visits to low-prior moves: what the pruning is for
        base = oracle.search_run_graph(og, 300, **kw)
        on = oracle.search_run_graph(og, 300, useNoisePruning=1, noisePruneUtilityScale=0.15, noisePruningCap=1e50, **kw)
        tight = oracle.search_run_graph(og, 300, useNoisePruning=1, noisePruneUtilityScale=0.01, noisePruningCap=1e50, **kw)
        off = oracle.search_run_graph(og, 300, useNoisePruning=1, noisePruneUtilityScale=0.15, noisePruningCap=0.0, chosenMovePrune=0.0, **kw)
        base0 = oracle.search_run_graph(og, 300, chosenMovePrune=0.0, **kw)
        assert on["rootVisits"] == tight["rootVisits"] == base["rootVisits"] == 300
        changed += on["digest"] != base["digest"]
        tighter += tight["digest"] != on["digest"]
        assert off["digest"] == base0["digest"] and (off["edgeVisits"] == base0["edgeVisits"]).all()
    assert changed >= 6 and tighter >= 6
