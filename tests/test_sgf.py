"""Custom Coffee SGF (README.md:33-35, cpp/dataio/sgf.cpp): writer / parser of the host library.  No GPU."""
import numpy as np
import pytest

from katacoffee_b200 import backend, capi


def test_sgf_literal_format():
    HW = 25
    p = lambda x, y, d: d * HW + y * 5 + x
    text = backend.writeSgf(5, 5, 4, [p(0, 0, 0), p(0, 1, 1), p(4, 1, 3)], winner=1, blackName="bot-a", whiteName="bot-b")
    # B[aaa] = column a, row a, direction a (|); third letter a..d = | - \ /  (README.md:35)
    assert text == "(;FF[4]GM[Coffee]SZ[5]WLL[4]PB[bot-a]PW[bot-b]RE[B+];B[aaa];W[abb];B[ebd])\n"
    st = np.zeros((6, 5), np.int8); st[2, 3] = 1; st[5, 0] = 2
    text = backend.writeSgf(5, 6, 3, [], winner=0, initialStones=st)
    assert text == "(;FF[4]GM[Coffee]SZ[5:6]WLL[3]PB[B200]PW[B200]RE[0]AB[dc]AW[af])\n"


def test_sgf_round_trip_of_oracle_games(oracle):
    for W, H, K, gid in ((5, 5, 4, 1), (6, 6, 4, 2), (4, 7, 3, 3)):
        og = oracle.Game(W, H, K)
        moves = []
        while not og.finished():
            mv = og.choose(9, gid)
            moves.append(mv)
            og.play(mv)
        text = backend.writeSgf(W, H, K, moves, winner=og.winner())
        g = backend.parseSgf(text)
        assert (g["xSize"], g["ySize"], g["winLen"], g["winner"]) == (W, H, K, og.winner())
        assert list(g["moves"]) == moves and list(g["players"]) == [1 + (i % 2) for i in range(len(moves))]
        assert not g["initialStones"].any()
        # replaying the parsed game reproduces the position
        rg = oracle.Game(W, H, K)
        for mv in g["moves"]:
            assert rg.play(int(mv))
        assert rg.status() == og.status() and tuple(rg.sit_hash()) == tuple(og.sit_hash())


def test_sgf_parser_details_and_errors():
    g = backend.parseSgf("(;GM[Coffee]FF[4]SZ[5]AB[aa:bb][ee]AW[cd]C[a \\] comment];B[cca];W[];B[dab](;W[eeb])(;W[aec]))")
    assert g["winLen"] == 4 and g["winner"] == -1
    assert g["initialStones"].tolist() == [[1, 1, 0, 0, 0], [1, 1, 0, 0, 0], [0, 0, 0, 0, 0], [0, 0, 2, 0, 0], [0, 0, 0, 0, 1]]
    # upper-case letters are 26.. for coordinates (off a 5x5 board -> rejected), but the direction letter is case-insensitive
    assert backend.parseSgf("(;SZ[5];B[ccA])")["moves"].tolist() == [0 * 25 + 2 * 5 + 2]
    assert g is not None
    g = backend.parseSgf("(;SZ[5];B[aab];W[];B[dab](;W[eeb])(;W[aec]))")
    assert g["moves"].tolist() == [25, -1, 25 + 3, 25 + 24] and g["players"].tolist() == [1, 2, 1, 2]     # main line = first variation
    assert backend.parseSgf("(;SZ[5]RE[W+R])")["winner"] == 2 and backend.parseSgf("(;SZ[5]RE[Draw])")["winner"] == 0
    assert backend.parseSgf("(;SZ[5];B[aaa];W[bbb];B[ccc])", maxMoves=2)["numMoves"] == 3
    for bad, why in (("(;FF[4];B[aaa])", "SZ"), ("(;SZ[5];B[aa])", "Invalid location"), ("(;SZ[5];B[aae])", "Invalid location"),
                     ("(;SZ[5];B[faa])", "Invalid location"), ("(;SZ[11])", "board size"), (";SZ[5]", "expected"), ("(;SZ[5];B[aaa", "unterminated"),
                     ("(;SZ[5]AB[zz])", "outside")):
        with pytest.raises(capi.KCError, match=why):
            backend.parseSgf(bad)
