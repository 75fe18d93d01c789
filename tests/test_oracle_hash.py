"""Pins the oracle's hashing / PRNG / Zobrist chain to the reference's own golden vectors
(cpp/core/rand.cpp:41-149, 386-507 -> tests/golden/hash_golden.json) and to the real reference
md5.cpp / sha2.cpp compiled into oracle/_ref; also checks the product's independent host-side
Zobrist generator against the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import golden


def test_md5_known_answer(oracle):
    g = golden("hash_golden.json")["md5"]
    out = (C.c_uint32 * 4)()
    msg = g["msg"].encode()
    oracle.lib().ko_md5(msg, len(msg), out)
    assert list(out) == g["expected"]


def test_sha256_known_answers(oracle):
    for case in golden("hash_golden.json")["sha256"]:
        out = (C.c_uint8 * 32)()
        msg = case["msg"].encode()
        oracle.lib().ko_sha256(msg, len(msg), out)
        assert bytes(out).hex() == case["hex"]


def test_xorshift1024_golden(oracle):
    g = golden("hash_golden.json")["xorshift1024"]
    init = np.array([int(x) for x in g["init_a"]], np.uint64)
    out = np.zeros(len(g["expected"]), np.uint32)
    oracle.lib().ko_xorshift1024_test(init.ctypes.data_as(C.c_void_p), len(out), out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == g["expected"]


def test_pcg32_golden(oracle):
    g = golden("hash_golden.json")["pcg32"]
    out = np.zeros(len(g["expected"]), np.uint32)
    oracle.lib().ko_pcg32_test(g["state"], len(out), out.ctypes.data_as(C.c_void_p))
    assert out.tolist() == g["expected"]


def test_rand_seeded_golden(oracle):
    g = golden("hash_golden.json")["rand"]
    r = oracle.lib().ko_rand_create(g["seed"].encode())
    vals = [oracle.lib().ko_rand_next_uint(r) for _ in g["expected"]]
    oracle.lib().ko_rand_destroy(r)
    assert vals == g["expected"]


def test_against_real_reference_hashes(oracle):
    ref = oracle.ref_hash_lib()
    if ref is None:
        pytest.skip("oracle/_ref not built (reference tree absent)")
    rng = np.random.default_rng(5)
    for n in [0, 1, 3, 55, 56, 57, 63, 64, 65, 119, 120, 121, 127, 128, 1000]:
        msg = bytes(rng.integers(1, 255, n, dtype=np.uint8))   # no NULs: the reference API is strlen-free here
        a, b = (C.c_uint32 * 4)(), (C.c_uint32 * 4)()
        oracle.lib().ko_md5(msg, n, a)
        ref.kref_md5(msg, n, b)
        assert list(a) == list(b), n
        a, b = (C.c_uint64 * 4)(), (C.c_uint64 * 4)()
        oracle.lib().ko_sha256_u64(msg, n, a)
        ref.kref_sha256_u64(msg, n, b)
        assert list(a) == list(b), n


def test_zobrist_survey_probe_values(oracle):
    """Values measured by the survey's probe of the real seeding chain."""
    board = np.zeros((133, 4, 2), np.uint64)
    player = np.zeros((4, 2), np.uint64)
    sx = np.zeros((11, 2), np.uint64)
    sy = np.zeros((11, 2), np.uint64)
    vp = C.c_void_p
    oracle.lib().ko_zobrist_tables(board.ctypes.data_as(vp), player.ctypes.data_as(vp), sx.ctypes.data_as(vp), sy.ctypes.data_as(vp))
    assert [hex(int(v)) for v in player[1]] == ["0xc535f97fd0cc7e76", "0x8a2a2a2ff24dbb6d"]
    assert [hex(int(v)) for v in player[2]] == ["0x392045e5c8d9bd73", "0xd3c1c132e034dcb0"]
    assert [hex(int(v)) for v in player[0]] == ["0x15e4ab319abb2cc6", "0x53219f89c01c557f"]
    assert [hex(int(v)) for v in player[3]] == ["0x9f4cdf6bbbf10bff", "0xc420fed9f31f05c8"]
    assert [hex(int(v)) for v in board[0][1]] == ["0xad00229ec323b548", "0xf4b593539cd89728"]
    assert [hex(int(v)) for v in sx[5]] == ["0x478fc9704f6fb627", "0xb1ea08b4f90dfbf6"]
    assert [hex(int(v)) for v in sy[5]] == ["0x8ae2f36ea4707544", "0x4a0b61a306b2d937"]
    assert [hex(int(v)) for v in sx[6]] == ["0xcbfa3568128cb44f", "0x7e0be6bb36f9ec98"]
    assert [hex(int(v)) for v in sy[6]] == ["0x30cb3b6e8ff75c22", "0xc24be82a513dfb9e"]
    assert not board[:, 0].any() and not board[:, 3].any()


def test_product_zobrist_equals_oracle(oracle, built_lib):
    from katacoffee_b200 import backend
    pb, pp, px, py = backend.zobristTables()
    board = np.zeros((133, 4, 2), np.uint64)
    player = np.zeros((4, 2), np.uint64)
    sx = np.zeros((11, 2), np.uint64)
    sy = np.zeros((11, 2), np.uint64)
    vp = C.c_void_p
    oracle.lib().ko_zobrist_tables(board.ctypes.data_as(vp), player.ctypes.data_as(vp), sx.ctypes.data_as(vp), sy.ctypes.data_as(vp))
    assert (pb == board).all() and (pp == player).all() and (px == sx).all() and (py == sy).all()
