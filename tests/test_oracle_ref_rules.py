"""Pins the oracle restatement (oracle/ko_game.cpp, ko_hash.cpp) to the REFERENCE'S OWN rules / hashing / NN-input code.

oracle/_ref/libkc_ref_rules.so is the reference's game/board.cpp, game/boardhistory.cpp, neuralnet/nninputs.cpp, core/hash.cpp,
core/rand.cpp ... compiled from a patched scratch copy (oracle/ref_patch.sh: one-line edits, each keyed to the SURVEY.md 0.2 defect
or the 8.1 ledger row that makes it necessary).  It exists where /root/reference does (this container), not on the GPU box: every
test here skips without it.  CPU only.

PINNED to the literal code by this file (>= 10^6 positions over five board shapes):
  a3  Board::initHash tables (board.cpp:134-178) and the Rand stream behind them (rand.cpp:276-318)
  a4  Board::isLegal (board.cpp:185-227): the full 4*H*W mask at every position, reachable and arbitrary
  a5  playMoveAssumeLegal (board.cpp:427-435): pos_hash
  a6  maxConsecutives / checkGameEnd (board.cpp:315-335, 376-383): per stone, and through the finished / winner bits
  a7  BoardHistory::makeBoardMove[AssumeLegal] (boardhistory.cpp:142-176): numTurns, isGameFinished, winner, next player
  a8  getSitHash (board.cpp:288-292)
  a9  NNInputs::getHash (nninputs.cpp:463-502) incl. the game-over, playout-doubling, temperature and optimism folds
  a10 fillRowV1 planes 0..10 and the global feature (nninputs.cpp:508-632, 656), NCHW and NHWC
  a11 copyInputs/OutputsWithSymmetry, getSymSpot, getSymDir, invert, compose (nninputs.cpp:252-433); NNPos::locToPos / getPolicySize
DIVERGES ON PURPOSE (SURVEY.md 8.1 ledger; the literal code is unusable there, asserted below where it can be shown):
  C  draw: the literal BoardHistory never ends a game without a winner -- the playout driver of the shim applies the canonical rule
  D  history: moves are made with makeBoardMove, the only variant that appends moveHistory
  F  fillRowV1 from "Feature 11" on (legal planes indexed by spot, 4 channels, running past channel 15): cut off, not compared;
     the literal code also needs NUM_FEATURES_SPATIAL_V1 = 16 as its NHWC pos stride where the canonical layout has 15 channels
  G  fillRowWithLine (orthogonal FOREACHADJ, walks wall spots): not compared
  I  NNPos::posToLoc: `pos /= HW` -- shown below to return the wrong cell for every dir >= 1
  J  getSymDir: returns `dir` where the literal code falls off its end (the one-line patch)
"""
import ctypes as C

import numpy as np
import pytest

SHAPES = [(5, 5, 4), (6, 6, 4), (4, 5, 3), (7, 7, 5), (3, 3, 3), (10, 10, 5), (9, 8, 4)]   # incl. the reference maximum (board.h:120)


@pytest.fixture(scope="module")
def ref(oracle):
    l = oracle.ref_rules_lib()
    if l is None:
        pytest.skip("oracle/_ref/libkc_ref_rules.so not built (needs /root/reference)")
    return l


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_zobrist_tables_and_rand_stream_equal_the_literal_code(oracle, ref):
    board = np.zeros((133, 4, 2), np.uint64); player = np.zeros((4, 2), np.uint64)
    sx = np.zeros((11, 2), np.uint64); sy = np.zeros((11, 2), np.uint64)
    assert ref.kc_ref_tables(_p(board), _p(player), _p(sx), _p(sy)) == 133
    t = oracle.zobrist_tables()
    assert (t["board"] == board).all() and (t["player"] == player).all() and (t["sizeX"] == sx).all() and (t["sizeY"] == sy).all()
    assert board[:, 1:3].any() and not board[:, 0].any() and not board[:, 3].any()     # empty / wall entries are zero (board.cpp:152-153)
    for seed in ("abc", "Board::initHash()", "katacoffee-b200:b10c128", ""):
        a32 = np.zeros(4096, np.uint32); a64 = np.zeros(4096, np.uint64)
        ref.kc_ref_rand(seed.encode(), 4096, _p(a32), _p(a64))
        o32, o64 = oracle.rand_stream(seed, 4096)
        assert (a32 == o32).all() and (a64 == o64).all(), seed


@pytest.mark.timeout(900)
def test_million_playout_positions_equal_the_literal_code(oracle, ref):
    """Every position of >= 10^6 random-legal plies: legal mask, status (numTurns / finished / winner / next player), move, sit-hash,
    NNInputs::getHash, and fillRowV1 planes 0..10 + global, restatement vs literal code."""
    total = 0
    for (W, H, K), games in zip(SHAPES, (30000, 12000, 6000, 3000, 2000)):
        HW = W * H
        for g0 in range(0, games, 3000):
            n = min(3000, games - g0)
            ro, po, go = oracle.playout_run(W, H, K, 20261018, g0, n, threads=8)
            rr, pr, gr = oracle.ref_playout_run(W, H, K, 20261018, g0, n)
            assert len(ro) == len(rr)
            for f in ("game", "status", "legal", "movePos", "sitHash", "nnHash"):
                assert (ro[f] == rr[f]).all(), (W, H, K, f)
            assert (go == gr).all() and (go == K).all()
            lit = pr.reshape(len(rr), 16, HW)[:, :11]
            assert (po.reshape(len(ro), 15, HW)[:, :11] == lit).all(), (W, H, K, "planes 0..10")
            assert not pr.reshape(len(rr), 16, HW)[:, 11:].any()      # the cut-off tail writes nothing (ledger F, G)
            total += len(ro)
        # the draw rule is the driver's (ledger C): such positions exist and are the only finished ones without a winner
        fin, win = (ro["status"] >> 8) & 1, (ro["status"] >> 9) & 3
        assert ((fin == 1) & (win == 0)).sum() > 0 or (W, H) == (7, 7)
    assert total >= 1_000_000, total


def test_nhwc_rows_of_the_literal_fill(oracle, ref):
    """NHWC: the literal code strides positions by NUM_FEATURES_SPATIAL_V1 = 16 (nninputs.h:50), the canonical layout by 15 (ledger F):
    channel c of cell p is the same value in both."""
    W, H, K = 5, 5, 4
    ro, po, _ = oracle.playout_run(W, H, K, 7, 0, 300, nhwc=True, threads=4)
    rr, pr, _ = oracle.ref_playout_run(W, H, K, 7, 0, 300, nhwc=True)
    assert len(ro) == len(rr) and len(ro) > 4000
    assert (po.reshape(len(ro), W * H, 15)[:, :, :11] == pr.reshape(len(rr), W * H, 16)[:, :, :11]).all()


@pytest.mark.parametrize("W,H,K", SHAPES + [(10, 10, 5), (9, 8, 5), (2, 2, 2)])
def test_arbitrary_positions_equal_the_literal_code(oracle, ref, W, H, K):
    """Positions no playout reaches (random stones, random last Loc incl. one on a stone or none): isLegal mask, maxConsecutives per
    stone, getSitHash, pos_hash."""
    rng = np.random.default_rng(W * 100 + H * 10 + K)
    HW = W * H
    LW = (4 * HW + 31) // 32
    og = oracle.Game(W, H, K)
    for it in range(1500 if HW <= 49 else 300):
        fill = rng.random()
        u = rng.random(HW)
        stones = np.where(u < fill / 2, 1, np.where(u < fill, 2, 0)).astype(np.int8)
        has_last = rng.random() < 0.9
        lx, ly, ld = (int(rng.integers(W)), int(rng.integers(H)), int(rng.integers(4))) if has_last else (-1, -1, 4)
        pla = int(rng.integers(1, 3))
        legal = np.zeros(13, np.uint32); mc = np.zeros(HW, np.int32); sh = np.zeros(2, np.uint64); ph = np.zeros(2, np.uint64)
        n = ref.kc_ref_position(W, H, K, _p(stones), lx, ly, ld, pla, _p(legal), _p(mc), _p(sh), _p(ph))
        og.reset()
        for c in range(HW):
            if stones[c]:
                og.set_stone(c % W, c // W, int(stones[c]))
        if has_last:
            og.set_last_loc(lx, ly, ld)
        em, en = og.legal_mask(pla)
        assert en == n and (em == legal[:LW]).all(), it
        assert (og.sit_hash(pla) == sh).all(), it
        for c in range(HW):
            if stones[c]:
                assert og.max_consecutives(c % W, c // W) == mc[c], (it, c)


def test_nn_hash_folds_equal_the_literal_code(oracle, ref):
    out = np.zeros(2, np.uint64)
    for (W, H, K) in ((5, 5, 4), (6, 6, 4), (10, 10, 5)):
        og = oracle.Game(W, H, K)
        for pla in (1, 2):
            for pda, temp, opt in ((0.0, 1.0, 0.0), (1.5, 1.0, 0.0), (-0.75, 1.0, 0.0), (0.0, 0.7, 0.0), (0.0, 1.25, 0.0), (0.0, 1.0, 0.6), (2.0, 0.8, 1.0)):
                ref.kc_ref_nn_hash_params(W, H, K, pla, 0, pda, temp, opt, _p(out))
                assert (og.nn_hash(pla, pda=pda, temp=temp, optimism=opt) == out).all(), (W, pla, pda, temp, opt)


def test_literal_graph_hash_chain(oracle, ref):
    """GraphHash::getGraphHash (graphhash.cpp:14-29) chained over whole games: the oracle's ko_graph_hash equals the literal code, and the
    literal getGraphHashFromScratch reproduces the chained value (with the getRecentBoard(0) the literal makeBoardMove leaves:
    recentBoards is pushed twice per move, ledger D, but slot 0 is always the current board)."""
    for (W, H, K) in ((5, 5, 4), (6, 6, 4)):
        for g in range(40):
            og = oracle.Game(W, H, K)
            moves, chain = [], []
            h = np.zeros(2, np.uint64)
            nh = np.zeros(2, np.uint64)
            oracle.lib().ko_graph_hash(_p(h), og._g, og.next_pla(), _p(nh))
            chain.append(nh.copy()); h = nh.copy()
            while not og.finished():
                pos = og.choose(99, g)
                if pos < 0:
                    break
                og.play(pos); moves.append(pos)
                nh = np.zeros(2, np.uint64)
                oracle.lib().ko_graph_hash(_p(h), og._g, og.next_pla(), _p(nh))
                chain.append(nh.copy()); h = nh.copy()
                if og.finished():
                    break
            mv = np.array(moves, np.int32)
            out = np.zeros((len(moves) + 2, 2), np.uint64)
            ref.kc_ref_graph_hash_chain(W, H, K, len(moves), _p(mv), _p(out))
            # the oracle's game may have ended in a draw the literal history does not know (ledger C): compare up to the last move
            # whose position the literal code also sees as unfinished or won
            drawn = og.finished() and og.winner() == 0
            upto = len(moves) + (0 if drawn else 1)
            assert (np.array(chain[:upto]) == out[:upto]).all(), (W, g)
            assert (out[len(moves)] == out[len(moves) + 1]).all()


def test_symmetry_helpers_equal_the_literal_code(oracle, ref):
    rng = np.random.default_rng(3)
    for (n, h, w, c) in ((2, 5, 5, 15), (1, 6, 6, 3), (3, 4, 5, 2), (1, 7, 3, 1)):
        src = rng.standard_normal(n * h * w * c).astype(np.float32)
        for sym in range(8):
            for nhwc in (False, True):
                a = np.zeros_like(src)
                ref.kc_ref_copy_inputs_with_symmetry(_p(src), _p(a), n, h, w, c, int(nhwc), sym)
                assert (a == oracle.copy_inputs_with_symmetry(src, n, h, w, c, nhwc, sym)).all(), (n, h, w, c, sym, nhwc)
            s1 = src[: n * h * w].copy()
            a = np.zeros_like(s1)
            ref.kc_ref_copy_outputs_with_symmetry(_p(s1), _p(a), n, h, w, sym)
            assert (a == oracle.copy_outputs_with_symmetry(s1, n, h, w, sym)).all()
    lib = oracle.lib()
    for sym in range(8):
        assert ref.kc_ref_sym_invert(sym) == lib.ko_sym_invert(sym)
        for s2 in range(8):
            assert ref.kc_ref_sym_compose(sym, s2) == lib.ko_sym_compose(sym, s2)
        for d in range(5):
            assert ref.kc_ref_sym_dir(d, sym) == lib.ko_sym_dir(d, sym)      # incl. the cases where the literal code fell off its end (ledger J)
        for (W, H) in ((5, 5), (6, 4)):
            for y in range(H):
                for x in range(W):
                    ax, ay, bx, by = C.c_int(), C.c_int(), C.c_int(), C.c_int()
                    ref.kc_ref_sym_xy(x, y, W, H, sym, C.byref(ax), C.byref(ay))
                    lib.ko_sym_xy(x, y, W, H, sym, C.byref(bx), C.byref(by))
                    assert (ax.value, ay.value) == (bx.value, by.value)


def test_nnpos_and_ledger_row_I(oracle, ref):
    """locToPos / getPolicySize are literal; the literal posToLoc divides where it should take the remainder (ledger I)."""
    for (W, H) in ((5, 5), (6, 6), (4, 7)):
        assert ref.kc_ref_policy_size(W, H) == 4 * W * H
        wrong = 0
        for d in range(4):
            for y in range(H):
                for x in range(W):
                    pos = ref.kc_ref_loc_to_pos(x, y, d, W, W, H)
                    assert pos == d * W * H + y * W + x                      # the policy index every component here uses
                    ox, oy, od = C.c_int(), C.c_int(), C.c_int()
                    ref.kc_ref_pos_to_loc(pos, W, H, W, H, C.byref(ox), C.byref(oy), C.byref(od))
                    assert od.value == d
                    if (ox.value, oy.value) != (x, y):
                        wrong += 1
                    else:
                        assert (x, y) == (d, 0)   # `pos /= HW` leaves pos = dir: the literal answer is cell (dir, 0)
        assert wrong == 4 * W * H - 4
