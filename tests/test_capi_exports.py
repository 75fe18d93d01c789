"""CPU checks of the drop-in boundary: the library builds, loads, exports every symbol that
include/katacoffee_b200.h declares, the host-only entry points work, and compute entry points fail
loudly (no CPU fallback) when no GPU is present."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "katacoffee_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(kc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(built_lib):
    from katacoffee_b200 import capi
    syms = declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(built_lib, s), f"{s} declared in the header but not exported"
        assert s in capi.PROTOTYPES, f"{s} has no ctypes prototype"
    assert sorted(capi.PROTOTYPES) == syms


def test_no_oracle_in_product():
    """The product must not reference the oracle (tests/bench/smoke may)."""
    for base, _, files in os.walk(os.path.join(ROOT, "katacoffee_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                src = open(os.path.join(base, f), errors="replace").read()
                assert "kc_oracle" not in src and "ko_" not in re.sub(r"[A-Za-z]ko_", "", src).replace("goko_", ""), f


def test_host_only_entry_points(built_lib):
    from katacoffee_b200 import backend
    assert built_lib.kc_abi_version() == 1
    board, player, sx, sy = backend.zobristTables()
    assert player.any() and board[7:36, 1:3].any() and not board[:, 0].any()


def test_compute_fails_loudly_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from katacoffee_b200 import backend, capi
    with pytest.raises(capi.KCError):
        backend.createComputeContext(0)


def test_symmetry_tables_match_oracle(oracle, built_lib):
    """The product's host symmetry map (used to build the device tables) equals copyWithSymmetry."""
    # exercised through the oracle only on CPU; the device tables are checked by the GPU parity tests.
    for h, w in ((5, 5), (6, 6), (3, 4)):
        src = np.arange(h * w, dtype=np.float32)
        for sym in range(8):
            a = oracle.copy_inputs_with_symmetry(src, 1, h, w, 1, False, sym)
            b = oracle.copy_inputs_with_symmetry(src, 1, h, w, 1, True, sym)
            assert (a == b).all()
            inv = oracle.copy_outputs_with_symmetry(a, 1, h, w, sym)
            if h == w or not (sym & 4):
                assert (inv == src).all(), (h, w, sym)   # outputs symmetry undoes inputs symmetry


def test_header_is_c_and_c_program_links(built_lib, tmp_path):
    """The boundary is plain C: the header compiles as C99 with -pedantic, and tests/cpp/test_cabi.c (C99, no C++) links against the library
    and drives the host-only entry points (hashes, the evaluator front end over a C batch function, SGF, the training-data file)."""
    import subprocess
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", os.path.join(inc, "katacoffee_b200.h")],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    libdir = os.path.join(ROOT, "katacoffee_b200")
    exe = str(tmp_path / "test_cabi")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-I" + inc, os.path.join(ROOT, "tests", "cpp", "test_cabi.c"), "-o", exe,
                        "-L" + libdir, "-lkatacoffee_b200", "-Wl,-rpath," + libdir], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "test_cabi: ok" in r.stdout, r.stdout + r.stderr
