"""Diagnostic (not a test): launches of the rules kernel alone (FEAT 0) and rules + fp32 NCHW planes (FEAT 1) at 65536 games."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend
ctx = backend.createComputeContext(0)
g = backend.Games(ctx, 65536, 5, 5, 4)
g.reset(seed=1, autoRefill=True)
for _ in range(4):
    g.step()
g.run(None, 6)
print("ms/ply", g.lastKernelMs())
