import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend, capi
from katacoffee_b200.capi import lib, check, ptr
import ctypes as C
ctx = backend.createComputeContext(0)
for N in (64, 96, 128):
    A = np.zeros((192, 16), np.uint16); B = np.zeros((N, 16), np.uint16)
    D = np.zeros((2, 128, N), np.float32)
    for reps in (2000, 20000):
        check(lib().kc_selftest_umma(ctx._p, ptr(A), ptr(B), ptr(D), 192, N, 16, 8, reps))
        print(f"N={N} reps={reps}: {D.flat[0]:.1f} cycles per 128x{N}x16 MMA (ideal {N/2:.0f})")
