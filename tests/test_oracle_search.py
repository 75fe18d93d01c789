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
