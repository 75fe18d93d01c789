"""Diagnostic (not a test): trunk kernel time at the bench shape. Usage: python tests/diag_trunk_time.py [net] [W] [reps]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend, modeldesc
net = sys.argv[1] if len(sys.argv) > 1 else "b10c128"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 5
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
ctx = backend.createComputeContext(0)
lm = backend.LoadedModel(ctx, modeldesc.Model(net, seed=1))
G = 18944 if W == 5 else 148 * 3 * 32
h = backend.createComputeHandle(ctx, lm, G, W, W)
games = backend.Games(ctx, G, W, W, 4); games.reset(seed=1, autoRefill=True)
games.runTimed(h, 3, 256 << 20)
h.trunkTime()
games.runTimed(h, reps, 256 << 20)
ms, cnt = h.trunkTime()
fl = modeldesc.flops_per_eval(net, W * W)
print(f"{net} {W}x{W} env={ {k: v for k, v in os.environ.items() if k.startswith('KC_')} }: trunk {ms / cnt:.4f} ms/launch -> {G / (ms / cnt) * 1e3 / 1e6:.3f} M evals/s, {G * fl / (ms / cnt * 1e-3) / 1e12:.1f} TFLOP/s")
