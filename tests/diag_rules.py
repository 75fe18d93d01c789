"""Diagnostic (not a test): rules+features kernel alone at 65,536 games (BASELINE config 2), both fused-kernel variants."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend
ctx = backend.createComputeContext(0)
for W in (5, 6):
    g = backend.Games(ctx, 65536, W, W, 4); g.reset(seed=1, autoRefill=True)
    g.runTimed(None, 8, 256 << 20)
    for rep in range(3):
        _, ms = g.runTimed(None, 24, 256 << 20)
        B = {5: 1572, 6: 2236}[W]
        print(f"{W}x{W}: {ms/24*1e3:.1f} us per ply, {65536*24/ms*1e3/1e9:.3f} G steps/s, {65536*24*B/ms*1e3/1e12:.3f} TB/s algorithmic")
    g.close()
