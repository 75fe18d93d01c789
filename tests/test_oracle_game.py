"""Hand-checked Coffee rule cases for the oracle restatement (cpp/game/board.cpp:185-227 isLegal,
:315-335/:376-383 win, cpp/game/boardhistory.cpp:157-176, ledger SURVEY.md 8.1) plus playout
properties.  The reference has no Coffee tests of its own: parity unpinned,
these cases pin the canonical reading instead."""
import numpy as np


def pos(d, x, y, W=5, H=5):
    return d * W * H + y * W + x


def test_empty_board_legal_count(oracle):
    g = oracle.Game(5, 5, 4)
    mask, n = g.legal_mask()
    assert n == 96   # 100 Locs minus the four corner diagonals whose line is a single cell
    assert not g.is_legal(0, 0, 3, 1) and not g.is_legal(4, 4, 3, 1)   # NE line x+y=0 / 8
    assert not g.is_legal(4, 0, 2, 1) and not g.is_legal(0, 4, 2, 1)   # NW line x-y=4 / -4
    assert g.is_legal(0, 0, 0, 1) and g.is_legal(0, 0, 1, 1) and g.is_legal(0, 0, 2, 1)
    assert not g.is_legal(0, 0, 0, 0) and not g.is_legal(0, 0, 4, 1) and not g.is_legal(5, 0, 0, 1)
    g6 = oracle.Game(6, 6, 4)
    assert g6.legal_mask()[1] == 140


def test_forced_line_follows_last_direction(oracle):
    g = oracle.Game(5, 5, 4)
    assert g.play(pos(0, 2, 2))            # black (2,2), direction N: white must answer in column 2
    assert g.next_pla() == 2 and g.num_turns() == 1
    for y in range(5):
        for x in range(5):
            for d in range(4):
                expect = (x == 2 and y != 2)
                assert g.is_legal(x, y, d, 2) == expect, (x, y, d)
    assert g.play(pos(1, 2, 0))            # white (2,0), direction W: black must answer in row 0
    for x in range(5):
        assert g.is_legal(x, 0, 0, 1) == (x != 2)
    assert not g.is_legal(2, 1, 0, 1)
    g.play(pos(2, 4, 0))                   # black (4,0) dir NW: line x-y=4 is that single cell -> illegal
    assert g.num_turns() == 2
    assert g.play(pos(3, 4, 0))            # black (4,0) dir NE: x+y=4 anti-diagonal
    for y in range(5):
        for x in range(5):
            assert g.is_legal(x, y, 0, 2) == (x + y == 4 and (x, y) != (4, 0) and (x, y) != (2, 2))
    assert not g.play(pos(0, 2, 2))        # occupied


def test_direction_needs_another_empty_cell_on_its_line(oracle):
    g = oracle.Game(5, 5, 4)
    # fill row 0 except (0,0): direction W from (0,0) has no other empty cell, N still has
    for x in range(1, 5):
        g.set_stone(x, 0, 1 + x % 2)
    assert not g.is_legal(0, 0, 1, 1)
    assert g.is_legal(0, 0, 0, 1) and g.is_legal(0, 0, 2, 1)
    # stones do not block the scan: an empty cell beyond a stone still counts
    g2 = oracle.Game(5, 5, 4)
    g2.set_stone(1, 0, 1); g2.set_stone(2, 0, 2); g2.set_stone(3, 0, 1)
    assert g2.is_legal(0, 0, 1, 1)


def test_win_overline_and_draw(oracle):
    g = oracle.Game(5, 5, 4)
    for x in (0, 1, 2):
        g.set_stone(x, 0, 1)
    assert g.play(pos(0, 3, 0)) and g.finished() and g.winner() == 1      # 4 in a row
    g = oracle.Game(5, 5, 4)
    for x in (0, 1, 3, 4):
        g.set_stone(x, 0, 1)
    assert g.play(pos(0, 2, 0)) and g.finished() and g.winner() == 1      # 5 in a row also wins (ledger N)
    g = oracle.Game(5, 5, 4)
    for x in (0, 1, 2):
        g.set_stone(x, x, 2)
    g.set_history([], 0, 2)
    assert g.play(pos(1, 3, 3)) and g.winner() == 2                       # diagonal, white
    # draw (ledger C): black plays (0,0) pointing N while column 0 has no other empty cell
    g = oracle.Game(5, 5, 4)
    for y in range(1, 5):
        g.set_stone(0, y, 1 + y % 2)
    assert g.play(pos(1, 0, 0)) is True     # direction W is legal (row 0 has empties) -> opponent forced to row 0
    assert not g.finished()
    g = oracle.Game(5, 5, 4)
    for y in range(1, 5):
        g.set_stone(0, y, 1 + y % 2)
    g.set_stone(1, 1, 2)
    # direction NW from (0,0): line x-y=0 has empties -> legal; after it, white must play on that diagonal
    assert g.play(pos(2, 0, 0)) and not g.finished()


def test_stalemate_is_a_draw(oracle):
    g = oracle.Game(5, 5, 4)
    # column 2 full except (2,0); black plays (2,0)?? needs another empty on column for dir N -> use dir W,
    # then fill row 0 so the forced row has no empty cell for white.
    g.set_stone(0, 0, 2); g.set_stone(1, 0, 1); g.set_stone(3, 0, 2)
    # row 0 now has empties (2,0) and (4,0): black plays (2,0) dir W -> legal; white forced to row 0: (4,0) only
    assert g.play(pos(1, 2, 0)) and not g.finished()
    mask, n = g.legal_mask()
    assert n >= 1 and all(((mask[p >> 5] >> (p & 31)) & 1) == 0 or (p % 25) == 4 for p in range(100))
    # white answers (4,0) dir W: row 0 has no other empty cell -> that Loc is illegal; dir N is fine
    assert not g.play(pos(1, 4, 0))
    assert g.play(pos(0, 4, 0))             # white (4,0) dir N -> black must play in column 4
    assert not g.finished()
    # now make the forced line empty-free: fill column 4 and let black be forced there
    g2 = oracle.Game(5, 5, 4)
    for y in range(1, 5):
        g2.set_stone(4, y, 1 + (y // 2) % 2)     # B B W W pattern broken: 1,1,2,2 -> no 4-run
    g2.set_stone(3, 0, 2)
    assert g2.play(pos(1, 4, 0))            # black (4,0) dir W (row 0 has empties) -> white forced to row 0, fine
    g3 = oracle.Game(5, 5, 4)
    for y in range(1, 5):
        g3.set_stone(4, y, 1 + (y // 2) % 2)
    for x in range(0, 3):
        g3.set_stone(x, 0, 1 + x % 2)
    # row 0: (3,0),(4,0) empty. black (3,0) dir W legal ((4,0) empty); white forced to row 0 -> only (4,0):
    assert g3.play(pos(1, 3, 0)) and not g3.finished()
    # white (4,0): dir N column 4 full -> illegal, dir W row full after placing? (row 0 has no other empty) illegal,
    # dir NW line x-y=4 single cell illegal, dir NE x+y=4: (3,1),(2,2),(1,3),(0,4) empty -> legal
    assert not g3.play(pos(0, 4, 0)) and not g3.play(pos(1, 4, 0)) and not g3.play(pos(2, 4, 0))
    assert g3.play(pos(3, 4, 0))
    # a true stalemate: forced line with no empty cell at all
    g4 = oracle.Game(5, 5, 4)
    for x in range(0, 4):
        g4.set_stone(x, 4, 1 + (x // 2) % 2)
    for y in range(0, 4):
        g4.set_stone(4, y, 1 + (y // 2) % 2)
    # black plays (4,4)? line choice: dir NW (x-y=0) has empties -> legal, white forced onto x-y=0 diagonal: fine.
    # Use direction N: column 4 has no other empty -> illegal. So stalemate must come from the forced line
    # being exhausted by the move itself: row 4 has only (4,4) empty; black (4,4) dir W illegal.  Construct instead:
    g5 = oracle.Game(5, 5, 4)
    for y in (1, 2, 3):
        g5.set_stone(0, y, 1 + y % 2)
    # column 0: (0,0),(0,4) empty. black (0,0) dir N legal; white forced to column 0 -> (0,4) only.
    assert g5.play(pos(0, 0, 0)) and not g5.finished()
    # white (0,4) with dir N: column 0 now full -> illegal; dir W legal -> black forced to row 4
    assert not g5.play(pos(0, 0, 4)) and g5.play(pos(1, 0, 4))
    assert not g5.finished()


def test_forced_line_exhausted_draw(oracle):
    """After a move whose direction line has no empty cell left for the opponent the game is a draw."""
    g = oracle.Game(5, 5, 4)
    # column 0: only (0,0) and (0,1) empty, alternating colours elsewhere (no 4-run)
    g.set_stone(0, 2, 1); g.set_stone(0, 3, 2); g.set_stone(0, 4, 1)
    assert g.play(pos(0, 0, 0))             # black (0,0) dir N: (0,1) is the other empty cell
    assert not g.finished()
    # white must play (0,1); choosing dir N is illegal (no other empty in the column), dir W is legal
    assert not g.play(pos(0, 0, 1))
    # make row 1 full so dir W is illegal too, and both diagonals through (0,1) full
    g = oracle.Game(5, 5, 4)
    g.set_stone(0, 2, 1); g.set_stone(0, 3, 2); g.set_stone(0, 4, 1)
    for x in range(1, 5):
        g.set_stone(x, 1, 1 + (x // 2) % 2)
    g.set_stone(1, 0, 2); g.set_stone(1, 2, 2); g.set_stone(2, 3, 1); g.set_stone(3, 4, 2)
    # black (0,0) dir N: legal; afterwards white's only cell (0,1) has no direction with another empty cell
    assert g.play(pos(0, 0, 0))
    assert g.finished() and g.winner() == 0
    assert g.legal_mask()[1] == 0


def test_sit_hash_is_zobrist_xor(oracle):
    import ctypes as C
    board = np.zeros((133, 4, 2), np.uint64)
    player = np.zeros((4, 2), np.uint64)
    sx = np.zeros((11, 2), np.uint64)
    sy = np.zeros((11, 2), np.uint64)
    vp = C.c_void_p
    oracle.lib().ko_zobrist_tables(board.ctypes.data_as(vp), player.ctypes.data_as(vp), sx.ctypes.data_as(vp), sy.ctypes.data_as(vp))
    g = oracle.Game(5, 5, 4)
    assert (g.sit_hash() == (sx[5] ^ sy[5] ^ player[1])).all()
    g.play(pos(0, 2, 2))
    spot = (2 + 1) + (2 + 1) * 6
    assert (g.sit_hash() == (sx[5] ^ sy[5] ^ board[spot][1] ^ player[2])).all()
    # ledger E: the literal hash ignores lastLoc -> same stones, other direction, same hash
    g2 = oracle.Game(5, 5, 4)
    g2.play(pos(1, 2, 2))
    assert (g2.sit_hash() == g.sit_hash()).all()
    # NNInputs::getHash adds GAME_IS_OVER only when finished
    assert (g.nn_hash() == g.sit_hash()).all()


def test_v1_planes_layout(oracle):
    g = oracle.Game(5, 5, 4)
    moves = [pos(0, 2, 2), pos(1, 2, 0), pos(3, 4, 0), pos(0, 3, 1), pos(1, 3, 3), pos(2, 0, 3)]
    for m in moves:
        assert g.play(m), m
    row, glob = g.fill_row_v1()
    P = row.reshape(15, 5, 5)
    assert glob[0] == 4.0 and (P[0] == 1).all()
    pla = g.next_pla()
    assert pla == 1
    own = {(2, 2), (4, 0), (3, 3)}
    opp = {(2, 0), (3, 1), (0, 3)}
    for y in range(5):
        for x in range(5):
            assert P[1, y, x] == ((x, y) in own) and P[2, y, x] == ((x, y) in opp)
    # last move (0,3) with direction 2 (NW) -> channel 3+2
    assert P[5, 3, 0] == 1 and P[3:7].sum() == 1
    # moves 2..5 plies ago in channels 7..10
    assert P[7, 3, 3] == 1 and P[8, 1, 3] == 1 and P[9, 0, 4] == 1 and P[10, 0, 2] == 1
    assert P[7:11].sum() == 4
    # legal-spot plane = OR over directions
    mask, n = g.legal_mask()
    spots = np.zeros(25, bool)
    for p in range(100):
        if (mask[p >> 5] >> (p & 31)) & 1:
            spots[p % 25] = True
    assert (P[11].reshape(-1) == spots).all()
    # NHWC is the transpose
    row2, _ = g.fill_row_v1(nhwc=True)
    assert (row2.reshape(25, 15).T.reshape(-1) == row).all()


def test_line_planes(oracle):
    g = oracle.Game(5, 5, 4)
    g.set_stone(0, 0, 1); g.set_stone(1, 0, 1); g.set_stone(2, 0, 1)     # run of 3 (k-1) along row 0
    g.set_stone(0, 2, 2); g.set_stone(1, 3, 2)                           # run of 2 (k-2) on a diagonal
    g.set_stone(4, 4, 1)                                                 # isolated: run of 1 (k-3) in every direction
    P = g.fill_row_v1()[0].reshape(15, 5, 5)
    assert {(x, y) for y in range(5) for x in range(5) if P[12, y, x]} == {(0, 0), (1, 0), (2, 0)}
    assert {(x, y) for y in range(5) for x in range(5) if P[13, y, x]} == {(0, 2), (1, 3)}
    # length-1 plane: every stone has some direction in which it stands alone
    assert {(x, y) for y in range(5) for x in range(5) if P[14, y, x]} == {(0, 0), (1, 0), (2, 0), (0, 2), (1, 3), (4, 4)}


def test_playout_statistics(oracle):
    recs, planes, glob = oracle.playout_run(5, 5, 4, seed=1, g0=0, n=2000, planes=False, threads=4)
    final = {}
    for r in recs:
        final[int(r["game"])] = int(r["status"])
    assert len(final) == 2000
    plies = np.array([s & 0xff for s in final.values()])
    assert all((s >> 8) & 1 for s in final.values())             # every game reaches a terminal position
    assert 7 <= plies.min() and plies.max() <= 24                # the last empty cell can never be filled
    assert 18 < plies.mean() < 23
    winners = np.array([(s >> 9) & 3 for s in final.values()])
    assert (winners == 0).mean() > 0.1 and (winners == 1).mean() > 0.2 and (winners == 2).mean() > 0.15
    # deterministic and independent of the thread count
    recs2, _, _ = oracle.playout_run(5, 5, 4, seed=1, g0=0, n=2000, planes=False, threads=1)
    assert recs.tobytes() == recs2.tobytes()
