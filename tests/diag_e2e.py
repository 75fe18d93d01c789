"""Diagnostic (not a test): kc_forward (NeuralNet::getOutput) throughput with host rows for several chunk schedules."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from katacoffee_b200 import backend, modeldesc
ctx = backend.createComputeContext(0)
G = 18944
lm = backend.LoadedModel(ctx, modeldesc.Model("b10c128", seed=1))
h = backend.createComputeHandle(ctx, lm, G, 5, 5)
games = backend.Games(ctx, G, 5, 5, 4); games.reset(seed=1, autoRefill=True); games.run(None, 6)
p, _ = games.features(nhwc=False)
pin = torch.empty((G, 375), dtype=torch.float32).pin_memory(); pin.numpy()[:] = p
glob = torch.full((G, 1), 4.0).pin_memory(); sym = np.zeros(G, np.int8)
outs = (torch.empty((G, 100)).pin_memory().numpy(), torch.empty((G, 2)).pin_memory().numpy(), torch.empty((G, 2)).pin_memory().numpy(), torch.empty((G, 25)).pin_memory().numpy())
for sched in sys.argv[1:] or ["default"]:
    if sched == "default": os.environ.pop("KC_FORWARD_SCHEDULE", None)
    else: os.environ["KC_FORWARD_SCHEDULE"] = sched
    import ctypes; ctypes.CDLL(None).setenv(b"KC_FORWARD_SCHEDULE", sched.encode(), 1) if sched != "default" else ctypes.CDLL(None).unsetenv(b"KC_FORWARD_SCHEDULE")
    for _ in range(5): backend.getOutput(h, pin.numpy(), glob.numpy(), sym, out=outs)
    best = 0
    for rep in range(3):
        t0 = time.perf_counter()
        for _ in range(40): backend.getOutput(h, pin.numpy(), glob.numpy(), sym, out=outs)
        dt = (time.perf_counter() - t0) / 40
        best = max(best, G / dt)
    print(f"schedule {sched:>16s}: {best/1e6:.3f} M evals/s ({G/best*1e3:.3f} ms per call)")
