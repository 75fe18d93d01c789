"""ctypes binding of the CPU ORACLE (oracle/kc_oracle.h).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product package katacoffee_b200 never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libkc_oracle.so")
REF_HASH_PATH = os.path.join(_HERE, "_ref", "libkc_ref_hash.so")
REF_RULES_PATH = os.path.join(_HERE, "_ref", "libkc_ref_rules.so")


def build(force=False):
    """make -C oracle: the restatement always; oracle/_ref only where /root/reference exists."""
    srcs = [os.path.join(_HERE, f) for f in ("ko_hash.cpp", "ko_game.cpp", "ko_net.cpp", "ko_search.cpp", "kc_oracle.h", "Makefile")]
    newest = max(os.path.getmtime(s) for s in srcs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        subprocess.run(["make", "-C", _HERE, "libkc_oracle.so"], check=True, stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/cpp/core/sha2.cpp") and (force or not os.path.exists(REF_HASH_PATH)):
        subprocess.run(["make", "-C", _HERE, "ref"], check=True, stdout=subprocess.DEVNULL)
    rules_srcs = [os.path.join(_HERE, f) for f in ("ref_patch.sh", "ref_rules_shim.cpp")]
    if os.path.exists("/root/reference/cpp/game/board.cpp") and (
            force or not os.path.exists(REF_RULES_PATH) or os.path.getmtime(REF_RULES_PATH) < max(os.path.getmtime(s) for s in rules_srcs)):
        subprocess.run(["make", "-C", _HERE, "refrules"], check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


class StepRecord(C.Structure):
    _fields_ = [("game", C.c_uint32), ("status", C.c_uint32), ("legal", C.c_uint32 * 13), ("movePos", C.c_int32),
                ("sitHash", C.c_uint64 * 2), ("nnHash", C.c_uint64 * 2)]


STEP_DTYPE = np.dtype([("game", "<u4"), ("status", "<u4"), ("legal", "<u4", (13,)), ("movePos", "<i4"),
                       ("sitHash", "<u8", (2,)), ("nnHash", "<u8", (2,))], align=True)

_lib = None
vp = C.c_void_p


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        l = C.CDLL(LIB_PATH)
        l.ko_splitmix64.restype = C.c_uint64
        l.ko_splitmix64.argtypes = [C.c_uint64]
        l.ko_rand_create.restype = vp
        l.ko_rand_create.argtypes = [C.c_char_p]
        l.ko_rand_destroy.argtypes = [vp]
        l.ko_rand_next_uint.restype = C.c_uint32
        l.ko_rand_next_uint.argtypes = [vp]
        l.ko_rand_next_uint64.restype = C.c_uint64
        l.ko_rand_next_uint64.argtypes = [vp]
        l.ko_md5.argtypes = [C.c_char_p, C.c_size_t, vp]
        l.ko_sha256.argtypes = [C.c_char_p, C.c_size_t, vp]
        l.ko_sha256_u64.argtypes = [C.c_char_p, C.c_size_t, vp]
        l.ko_xorshift1024_test.argtypes = [vp, C.c_int, vp]
        l.ko_pcg32_test.argtypes = [C.c_uint64, C.c_int, vp]
        l.ko_zobrist_tables.argtypes = [vp, vp, vp, vp]
        l.ko_game_create.restype = vp
        l.ko_game_create.argtypes = [C.c_int, C.c_int, C.c_int]
        for name in ("ko_game_destroy", "ko_game_reset"):
            getattr(l, name).argtypes = [vp]
        l.ko_game_set_stone.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        l.ko_game_set_last_loc.argtypes = [vp, C.c_int, C.c_int, C.c_int]
        l.ko_game_set_history.argtypes = [vp, C.c_int, vp, vp, C.c_int, C.c_int]
        l.ko_game_is_legal.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        l.ko_game_legal_mask.argtypes = [vp, C.c_int, vp]
        l.ko_game_play.argtypes = [vp, C.c_int]
        for name in ("ko_game_next_pla", "ko_game_num_turns", "ko_game_finished", "ko_game_winner"):
            getattr(l, name).argtypes = [vp]
        l.ko_game_status.restype = C.c_uint32
        l.ko_game_status.argtypes = [vp]
        l.ko_game_max_consecutives.argtypes = [vp, C.c_int, C.c_int]
        l.ko_game_sit_hash.argtypes = [vp, C.c_int, vp]
        l.ko_game_nn_hash.argtypes = [vp, C.c_int, C.c_double, C.c_float, C.c_double, vp]
        l.ko_game_fill_row_v1.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        l.ko_copy_inputs_with_symmetry.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        l.ko_copy_outputs_with_symmetry.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        l.ko_sym_dir.argtypes = [C.c_int, C.c_int]
        l.ko_playout_choose.argtypes = [vp, C.c_uint64, C.c_uint64, vp]
        l.ko_playout_run.restype = C.c_long
        l.ko_playout_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, vp, C.c_long,
                                     vp, C.c_int, vp, C.c_int]
        l.ko_model_create.restype = vp
        l.ko_model_create.argtypes = [vp]
        l.ko_model_destroy.argtypes = [vp]
        l.ko_model_forward.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int]
        l.ko_model_forward_trace.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, vp]
        l.ko_test_conv.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int]
        l.ko_test_batchnorm.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp]
        l.ko_test_resblock.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int]
        l.ko_postprocess.argtypes = [vp, C.c_int, vp, C.c_float, vp, vp, C.c_int]
        l.ko_search_run.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        l.ko_search_choose.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64]
        l.ko_search_choose_temperature.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64]
        l.ko_search_choose_values.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_uint64, C.c_uint64]
        l.ko_search_last_play_selection.argtypes = [vp, C.c_int]
        l.ko_search_run_graph.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        l.ko_graph_hash.argtypes = [vp, vp, C.c_int, vp]
        l.ko_game_recent_move_pos.argtypes = [vp, C.c_int]
        l.ko_graph_search_create.restype = vp
        l.ko_graph_search_create.argtypes = [C.c_int, C.c_int, vp]
        l.ko_graph_search_destroy.argtypes = [vp]
        l.ko_graph_search_continue.argtypes = [vp] * 11
        l.ko_graph_search_advance.argtypes = [vp, C.c_int]
        l.ko_graph_search_digest.argtypes = [vp]
        l.ko_graph_search_digest.restype = C.c_uint64
        l.ko_graph_search_num_nodes.argtypes = [vp]
        l.ko_search_create.restype = vp
        l.ko_search_destroy.argtypes = [vp]
        l.ko_search_clear.argtypes = [vp]
        l.ko_search_continue.argtypes = [vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
        l.ko_search_advance.argtypes = [vp, C.c_int]
        l.ko_training_rows.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, C.c_uint64, vp, vp, vp, vp, vp]
        l.ko_game_color_at.argtypes = [vp, C.c_int, C.c_int]
        _lib = l
    return _lib


def ref_hash_lib():
    """The REAL reference md5/sha2 (oracle/_ref), or None if it was not built/shipped."""
    if not os.path.exists(REF_HASH_PATH):
        return None
    l = C.CDLL(REF_HASH_PATH)
    l.kref_md5.argtypes = [C.c_char_p, C.c_size_t, vp]
    l.kref_sha256_u64.argtypes = [C.c_char_p, C.c_size_t, vp]
    l.kref_sha256_bytes.argtypes = [C.c_char_p, C.c_size_t, vp]
    return l


_ref_rules = None


def ref_rules_lib():
    """oracle/_ref/libkc_ref_rules.so: the reference's OWN board / boardhistory / nninputs / hash / rand code (patched scratch copy,
    oracle/ref_patch.sh) behind the C entry points of oracle/ref_rules_shim.cpp, or None where it has not been built."""
    global _ref_rules
    if _ref_rules is None:
        if not os.path.exists(REF_RULES_PATH):
            return None
        l = C.CDLL(REF_RULES_PATH)
        l.kc_ref_tables.restype = C.c_int
        l.kc_ref_tables.argtypes = [vp, vp, vp, vp]
        l.kc_ref_rand.argtypes = [C.c_char_p, C.c_int, vp, vp]
        l.kc_ref_playout_run.restype = C.c_long
        l.kc_ref_playout_run.argtypes = [C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_uint64, C.c_int, C.c_int, vp, C.c_long, vp, C.c_int, vp]
        l.kc_ref_position.restype = C.c_int
        l.kc_ref_position.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp]
        l.kc_ref_nn_hash_params.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_float, C.c_double, vp]
        l.kc_ref_copy_inputs_with_symmetry.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
        l.kc_ref_copy_outputs_with_symmetry.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int]
        l.kc_ref_sym_xy.argtypes = [C.c_int] * 5 + [vp, vp]
        l.kc_ref_pos_to_loc.argtypes = [C.c_int] * 5 + [vp, vp, vp]
        l.kc_ref_graph_hash_chain.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        l.kc_ref_init()
        _ref_rules = l
    return _ref_rules


def ref_playout_run(x, y, k, seed, g0, n, max_plies=255, planes=True, nhwc=False, max_records=None):
    """The same synthetic playouts as playout_run, played on the reference's own Board / BoardHistory; planes are the literal
    fillRowV1 rows [records][16*H*W] (channels 0..10 meaningful, see oracle/ref_rules_shim.cpp)."""
    l = ref_rules_lib()
    cap = max_records if max_records is not None else n * (x * y + 2)
    recs = np.zeros(cap, STEP_DTYPE)
    pl = np.zeros((cap, 16 * x * y), np.float32) if planes else None
    gl = np.zeros(cap, np.float32) if planes else None
    got = l.kc_ref_playout_run(x, y, k, seed, g0, n, max_plies, _p(recs), cap, _p(pl), int(nhwc), _p(gl))
    assert got >= 0
    return recs[:got], (pl[:got] if planes else None), (gl[:got] if planes else None)


def zobrist_tables():
    """The restatement's Board::initHash tables: board [133][4][2], player [4][2], sizeX / sizeY [11][2] (hash0, hash1)."""
    board = np.zeros((133, 4, 2), np.uint64); player = np.zeros((4, 2), np.uint64)
    sx = np.zeros((11, 2), np.uint64); sy = np.zeros((11, 2), np.uint64)
    lib().ko_zobrist_tables(_p(board), _p(player), _p(sx), _p(sy))
    return {"board": board, "player": player, "sizeX": sx, "sizeY": sy}


def rand_stream(seed, n):
    """n values of Rand(seed).nextUInt() and, from a fresh generator, n of nextUInt64()."""
    l = lib()
    a = l.ko_rand_create(seed.encode())
    v32 = np.array([l.ko_rand_next_uint(a) for _ in range(n)], np.uint32)
    l.ko_rand_destroy(a)
    b = l.ko_rand_create(seed.encode())
    v64 = np.array([l.ko_rand_next_uint64(b) for _ in range(n)], np.uint64)
    l.ko_rand_destroy(b)
    return v32, v64


def _p(a):
    return None if a is None else a.ctypes.data_as(vp)


class Game:
    def __init__(self, x=5, y=5, k=4):
        self.W, self.H, self.K = x, y, k
        self.HW = x * y
        self.LW = (4 * self.HW + 31) // 32
        self._g = lib().ko_game_create(x, y, k)

    def __del__(self):
        if getattr(self, "_g", None):
            lib().ko_game_destroy(self._g)
            self._g = None

    def reset(self):
        lib().ko_game_reset(self._g)

    def set_stone(self, x, y, color):
        return lib().ko_game_set_stone(self._g, x, y, color)

    def set_last_loc(self, x, y, d):
        lib().ko_game_set_last_loc(self._g, x, y, d)

    def set_history(self, moves, num_turns, next_pla):
        """moves: list of (pos, pla) oldest first (pos = policy index)."""
        n = len(moves)
        pos = np.array([m[0] for m in moves], np.int32)
        pla = np.array([m[1] for m in moves], np.int32)
        lib().ko_game_set_history(self._g, n, _p(pos), _p(pla), num_turns, next_pla)

    def is_legal(self, x, y, d, pla):
        return bool(lib().ko_game_is_legal(self._g, x, y, d, pla))

    def legal_mask(self, pla=None):
        out = np.zeros(13, np.uint32)
        n = lib().ko_game_legal_mask(self._g, self.next_pla() if pla is None else pla, _p(out))
        return out[:self.LW].copy(), n

    def play(self, pos):
        return bool(lib().ko_game_play(self._g, pos))

    def next_pla(self):
        return lib().ko_game_next_pla(self._g)

    def num_turns(self):
        return lib().ko_game_num_turns(self._g)

    def finished(self):
        return bool(lib().ko_game_finished(self._g))

    def winner(self):
        return lib().ko_game_winner(self._g)

    def status(self):
        return int(lib().ko_game_status(self._g))

    def sit_hash(self, pla=None):
        out = np.zeros(2, np.uint64)
        lib().ko_game_sit_hash(self._g, self.next_pla() if pla is None else pla, _p(out))
        return out

    def max_consecutives(self, x, y):
        return lib().ko_game_max_consecutives(self._g, x, y)

    def nn_hash(self, pla=None, pda=0.0, temp=1.0, optimism=0.0):
        out = np.zeros(2, np.uint64)
        lib().ko_game_nn_hash(self._g, self.next_pla() if pla is None else pla, pda, temp, optimism, _p(out))
        return out

    def fill_row_v1(self, nhwc=False, pla=None):
        row = np.zeros(15 * self.HW, np.float32)
        g = np.zeros(1, np.float32)
        lib().ko_game_fill_row_v1(self._g, self.next_pla() if pla is None else pla, self.W, self.H, int(nhwc), _p(row), _p(g))
        return row, g

    def choose(self, seed, game_idx):
        return lib().ko_playout_choose(self._g, seed, game_idx, None)


def playout_run(x, y, k, seed, g0, n, max_plies=255, planes=True, nhwc=False, threads=1, max_records=None):
    hw = x * y
    if max_records is None:
        max_records = n * (hw + 1)
    recs = np.zeros(max_records, STEP_DTYPE)
    assert STEP_DTYPE.itemsize == C.sizeof(StepRecord)
    pl = np.zeros((max_records, 15 * hw), np.float32) if planes else None
    gl = np.zeros(max_records, np.float32) if planes else None
    total = lib().ko_playout_run(x, y, k, seed, g0, n, max_plies, _p(recs), max_records, _p(pl), int(nhwc), _p(gl), threads)
    return recs[:total], (pl[:total] if planes else None), (gl[:total] if planes else None)


def copy_inputs_with_symmetry(src, n, h, w, c, nhwc, sym):
    src = np.ascontiguousarray(src, np.float32)
    dst = np.zeros_like(src)
    lib().ko_copy_inputs_with_symmetry(_p(src), _p(dst), n, h, w, c, int(nhwc), sym)
    return dst


def copy_outputs_with_symmetry(src, n, h, w, sym):
    src = np.ascontiguousarray(src, np.float32)
    dst = np.zeros_like(src)
    lib().ko_copy_outputs_with_symmetry(_p(src), _p(dst), n, h, w, sym)
    return dst


class Model:
    """Oracle copy of a katacoffee_b200.modeldesc.Model (same POD description)."""

    def __init__(self, model):
        self.model = model
        self._m = lib().ko_model_create(C.byref(model.desc))

    def __del__(self):
        if getattr(self, "_m", None):
            lib().ko_model_destroy(self._m)
            self._m = None

    def forward(self, rowSpatial, rowGlobal, x, y, symmetry=None, nhwc=False, mode=0, threads=1):
        n = rowSpatial.shape[0]
        hw = x * y
        rs = np.ascontiguousarray(rowSpatial, np.float32)
        rg = np.ascontiguousarray(rowGlobal, np.float32)
        sym = None if symmetry is None else np.ascontiguousarray(symmetry, np.int8)
        policy = np.zeros((n, 4 * hw), np.float32)
        value = np.zeros((n, 2), np.float32)
        misc = np.zeros((n, 2), np.float32)
        own = np.zeros((n, hw), np.float32)
        lib().ko_model_forward(self._m, n, x, y, int(nhwc), _p(rs), _p(rg), _p(sym), _p(policy), _p(value), _p(misc), _p(own), mode, threads)
        return policy, value, misc, own


def mode_emul(act_fmt, w_fmt):
    """Operand formats of the reduced-precision emulation (KO_MODE_EMUL): 'bf16', 'fp16' or 'fp32' for activations / weights."""
    f = {"bf16": 0, "fp16": 1, "fp32": 2}
    return 2 | (f[act_fmt] << 4) | (f[w_fmt] << 6)


def forward_trace(om, rowSpatial, rowGlobal, x, y, mode=0):
    """(policy, value, misc, own, trace) with trace [numBlocks + 2][n][hw][trunkC]: the trunk after the initial conv, each block, the tip."""
    n, hw = rowSpatial.shape[0], x * y
    rs = np.ascontiguousarray(rowSpatial, np.float32)
    rg = np.ascontiguousarray(rowGlobal, np.float32)
    policy = np.zeros((n, 4 * hw), np.float32)
    value = np.zeros((n, 2), np.float32)
    misc = np.zeros((n, 2), np.float32)
    own = np.zeros((n, hw), np.float32)
    trace = np.zeros((om.model.num_blocks + 2, n, hw, om.model.trunk), np.float32)
    lib().ko_model_forward_trace(om._m, n, x, y, _p(rs), _p(rg), _p(policy), _p(value), _p(misc), _p(own), mode, _p(trace))
    return policy, value, misc, own, trace


def test_conv(desc_struct, n, xlen, ylen, nhwc, inp, out_channels, mode=0):
    inp = np.ascontiguousarray(inp, np.float32)
    out = np.zeros(n * xlen * ylen * out_channels, np.float32)
    lib().ko_test_conv(C.byref(desc_struct), n, xlen, ylen, int(nhwc), _p(inp), _p(out), mode)
    return out


def test_batchnorm(desc_struct, activation, n, xlen, ylen, nhwc, inp, mask):
    inp = np.ascontiguousarray(inp, np.float32)
    mask = np.ascontiguousarray(mask, np.float32)
    out = np.zeros_like(inp)
    lib().ko_test_batchnorm(C.byref(desc_struct), activation, n, xlen, ylen, int(nhwc), _p(inp), _p(mask), _p(out))
    return out


def test_resblock(block_struct, n, xlen, ylen, nhwc, inp, mask, mode=0):
    inp = np.ascontiguousarray(inp, np.float32)
    mask = np.ascontiguousarray(mask, np.float32)
    out = np.zeros_like(inp)
    lib().ko_test_resblock(C.byref(block_struct), n, xlen, ylen, int(nhwc), _p(inp), _p(mask), _p(out), mode)
    return out


def postprocess(policy, legal_mask, value2, misc2, next_pla, temp=1.0):
    p = np.ascontiguousarray(policy, np.float32).copy()
    v = np.ascontiguousarray(value2, np.float32).copy()
    m = np.ascontiguousarray(misc2, np.float32).copy()
    lm = np.ascontiguousarray(legal_mask, np.uint32)
    lib().ko_postprocess(_p(p), p.shape[0], _p(lm), temp, _p(v), _p(m), next_pla)
    return p, v, m


class SearchParams(C.Structure):
    """Same layout as kc_search_params (include/katacoffee_b200.h)."""
    _fields_ = [("maxVisits", C.c_int32), ("temperaturePlies", C.c_int32), ("autoRefill", C.c_int32), ("noCompaction", C.c_int32), ("reuseTree", C.c_int32), ("useGraphSearch", C.c_int32),
                ("cpuctExploration", C.c_double), ("fpuReductionMax", C.c_double), ("rootFpuReductionMax", C.c_double),
                ("subtreeValueBiasFactor", C.c_double), ("subtreeValueBiasWeightExponent", C.c_double), ("subtreeValueBiasFreeProp", C.c_double),
                ("rootNoiseEnabled", C.c_int32), ("fpuParentWeightByVisitedPolicy", C.c_int32),
                ("rootDirichletNoiseTotalConcentration", C.c_double), ("rootDirichletNoiseWeight", C.c_double),
                ("rootPolicyTemperature", C.c_double), ("rootPolicyTemperatureEarly", C.c_double), ("chosenMoveTemperatureHalflife", C.c_double),
                ("fpuParentWeightByVisitedPolicyPow", C.c_double), ("rootDesiredPerChildVisitsCoeff", C.c_double), ("valueWeightExponent", C.c_double),
                ("noiseSeed", C.c_uint64), ("noiseGameId", C.c_uint64), ("nnRandomize", C.c_int32), ("pad3_", C.c_int32),   # oracle-only tail
                ("useLcbForSelection", C.c_int32), ("useNonBuggyLcb", C.c_int32), ("lcbStdevs", C.c_double), ("minVisitPropForLCB", C.c_double),
                ("rootNumSymmetriesToSample", C.c_int32), ("useNoisePruning", C.c_int32), ("useUncertainty", C.c_int32), ("pad4_", C.c_int32),
                ("uncertaintyCoeff", C.c_double), ("uncertaintyExponent", C.c_double), ("uncertaintyMaxWeight", C.c_double),
                ("chosenMoveSubtract", C.c_double), ("chosenMovePrune", C.c_double),
                ("noisePruneUtilityScale", C.c_double), ("noisePruningCap", C.c_double)]


def _with_extras(sp, extra):
    """Sets the optional search options (root noise / temperature, FPU parent weight, ...) given as keyword arguments."""
    for k, v in extra.items():
        if not hasattr(sp, k):
            raise TypeError(f"unknown search option {k}")
        setattr(sp, k, v)
    return sp


def search_run(game, max_visits, model=None, cpuct=1.0, fpu=0.2, root_fpu=0.2, **extra):
    """One oracle search from `game` (a Game); model=None uses the integer-hash evaluator.  Returns a dict with the
    root statistics (arrays over the policy index) and the visit counters."""
    P = 4 * game.HW
    sp = _with_extras(SearchParams(max_visits, 0, 0, 0, 0, 0, cpuct, fpu, root_fpu), extra)
    rv = np.zeros(1, np.int32); rw = np.zeros(1, np.float64)
    ev = np.zeros(P, np.int32); ew = np.zeros(P, np.float64); pol = np.zeros(P, np.float32); order = np.zeros(P, np.uint8)
    cnt = np.zeros(3, np.uint64)
    lib().ko_search_run(game._g, game.W, game.H, C.byref(sp), None if model is None else model._m, _p(rv), _p(rw), _p(ev), _p(ew), _p(pol),
                        _p(order), _p(cnt))
    return {"rootVisits": int(rv[0]), "rootUtilitySum": float(rw[0]), "edgeVisits": ev, "edgeUtilitySum": ew, "policy": pol, "order": order,
            "counters": cnt}


def search_run_graph(game, max_visits, model=None, cpuct=1.0, fpu=0.2, root_fpu=0.2, graph=True, bias_factor=0.0, bias_exponent=0.5, **extra):
    """One oracle graph search (transpositions + subtree value bias) from `game`; see ko_search.cpp.  counters = visits,
    evaluations, terminal visits, transposition hits, catch-up visits; digest = hash over the whole graph."""
    P = 4 * game.HW
    sp = _with_extras(SearchParams(max_visits, 0, 0, 0, 0, int(graph), cpuct, fpu, root_fpu, bias_factor, bias_exponent), extra)
    rv = np.zeros(1, np.int32); rw = np.zeros(1, np.float64)
    ev = np.zeros(P, np.int32); ew = np.zeros(P, np.float64); pol = np.zeros(P, np.float32); order = np.zeros(P, np.uint8)
    cnt = np.zeros(5, np.uint64); dg = np.zeros(1, np.uint64)
    lib().ko_search_run_graph(game._g, game.W, game.H, C.byref(sp), None if model is None else model._m, _p(rv), _p(rw), _p(ev), _p(ew),
                              _p(pol), _p(order), _p(cnt), _p(dg))
    return {"rootVisits": int(rv[0]), "rootUtilitySum": float(rw[0]), "edgeVisits": ev, "edgeUtilitySum": ew, "policy": pol, "order": order,
            "counters": cnt, "digest": int(dg[0]), "playSelection": _last_play_selection(P)}


def _last_play_selection(P):
    """Search::getPlaySelectionValues at the root of the graph search that just ran on this thread (see ko_search.cpp)."""
    out = np.zeros(P, np.float64)
    lib().ko_search_last_play_selection(_p(out), P)
    return out


class PersistentGraphSearch:
    """Oracle graph search (transpositions + subtree value bias) that re-roots its graph at the move played (tree re-use)."""

    def __init__(self, W, H, max_visits, cpuct=1.0, fpu=0.2, root_fpu=0.2, graph=True, bias_factor=0.0, bias_exponent=0.5, free_prop=0.8, **extra):
        self.P = 4 * W * H
        sp = _with_extras(SearchParams(max_visits, 0, 0, 0, 1, int(graph), cpuct, fpu, root_fpu, bias_factor, bias_exponent, free_prop), extra)
        self._s = lib().ko_graph_search_create(W, H, C.byref(sp))

    def __del__(self):
        if getattr(self, "_s", None):
            lib().ko_graph_search_destroy(self._s)
            self._s = None

    def run(self, game, model=None):
        P = self.P
        rv = np.zeros(1, np.int32); rw = np.zeros(1, np.float64)
        ev = np.zeros(P, np.int32); ew = np.zeros(P, np.float64); pol = np.zeros(P, np.float32); order = np.zeros(P, np.uint8)
        cnt = np.zeros(5, np.uint64); dg = np.zeros(1, np.uint64)
        lib().ko_graph_search_continue(self._s, game._g, None if model is None else model._m, _p(rv), _p(rw), _p(ev), _p(ew), _p(pol), _p(order),
                                       _p(cnt), _p(dg))
        return {"rootVisits": int(rv[0]), "rootUtilitySum": float(rw[0]), "edgeVisits": ev, "edgeUtilitySum": ew, "policy": pol, "order": order,
                "counters": cnt, "digest": int(dg[0]), "playSelection": _last_play_selection(P)}

    def advance(self, move_pos):
        lib().ko_graph_search_advance(self._s, int(move_pos))

    def digest(self):
        return int(lib().ko_graph_search_digest(self._s))

    def num_nodes(self):
        return int(lib().ko_graph_search_num_nodes(self._s))


class PersistentSearch:
    """Oracle search that keeps the chosen child's subtree between moves (tree re-use)."""

    def __init__(self):
        self._s = lib().ko_search_create()

    def __del__(self):
        if getattr(self, "_s", None):
            lib().ko_search_destroy(self._s)
            self._s = None

    def run(self, game, max_visits, model=None, cpuct=1.0, fpu=0.2, root_fpu=0.2):
        P = 4 * game.HW
        sp = SearchParams(max_visits, 0, 0, 0, 1, 0, cpuct, fpu, root_fpu)
        rv = np.zeros(1, np.int32); rw = np.zeros(1, np.float64)
        ev = np.zeros(P, np.int32); ew = np.zeros(P, np.float64); pol = np.zeros(P, np.float32); order = np.zeros(P, np.uint8)
        cnt = np.zeros(3, np.uint64)
        lib().ko_search_continue(self._s, game._g, game.W, game.H, C.byref(sp), None if model is None else model._m, _p(rv), _p(rw), _p(ev),
                                 _p(ew), _p(pol), _p(order), _p(cnt))
        return {"rootVisits": int(rv[0]), "rootUtilitySum": float(rw[0]), "edgeVisits": ev, "edgeUtilitySum": ew, "policy": pol, "order": order,
                "counters": cnt}

    def advance(self, move_pos):
        lib().ko_search_advance(self._s, int(move_pos))


def search_choose(edge_visits, order, ply, temperature_plies, seed, game_id):
    ev = np.ascontiguousarray(edge_visits, np.int32); od = np.ascontiguousarray(order, np.uint8)
    return lib().ko_search_choose(_p(ev), _p(od), len(ev), ply, temperature_plies, seed, game_id)


def search_choose_values(values, order, board_area, ply, temp_early, temp_late, halflife, subtract, prune, seed, game_id):
    """The move choice under the temperature schedule on full play-selection values (doubles)."""
    v = np.ascontiguousarray(values, np.float64); od = np.ascontiguousarray(order, np.uint8)
    return lib().ko_search_choose_values(_p(v), _p(od), len(v), board_area, ply, temp_early, temp_late, halflife, subtract, prune, seed, game_id)


def search_choose_temperature(edge_visits, order, board_area, ply, temp_early, temp_late, halflife, subtract, prune, seed, game_id):
    ev = np.ascontiguousarray(edge_visits, np.int32); od = np.ascontiguousarray(order, np.uint8)
    return lib().ko_search_choose_temperature(_p(ev), _p(od), len(ev), board_area, ply, temp_early, temp_late, halflife, subtract, prune, seed, game_id)


def training_rows(x, y, k, moves, root_n, root_w, visits, game_id):
    """Oracle training rows of one finished game (dict of the reference's npz arrays, R rows)."""
    R, HW = len(moves), x * y
    P = 4 * HW
    mv = np.ascontiguousarray(moves, np.int32); rn = np.ascontiguousarray(root_n, np.int32); rw = np.ascontiguousarray(root_w, np.float64)
    vs = np.ascontiguousarray(visits, np.int16).reshape(R, P)
    out = dict(binaryInputNCHWPacked=np.zeros((R, 15, (HW + 7) // 8), np.uint8), globalInputNC=np.zeros((R, 1), np.float32),
               policyTargetsNCMove=np.zeros((R, 2, P), np.int16), globalTargetsNC=np.zeros((R, 64), np.float32),
               valueTargetsNCHW=np.zeros((R, 5, y, x), np.int8))
    lib().ko_training_rows(x, y, k, R, _p(mv), _p(rn), _p(rw), _p(vs), game_id, _p(out["binaryInputNCHWPacked"]), _p(out["globalInputNC"]),
                           _p(out["policyTargetsNCMove"]), _p(out["globalTargetsNC"]), _p(out["valueTargetsNCHW"]))
    return out
