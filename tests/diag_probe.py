"""Diagnostic (not a test): timeline of one layer boundary inside the trunk kernel (KC_TRUNK_PROBE=1)."""
import os, sys
os.environ["KC_TRUNK_PROBE"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend, modeldesc
from katacoffee_b200.capi import lib, check, ptr
ctx = backend.createComputeContext(0)
lm = backend.LoadedModel(ctx, modeldesc.Model("b10c128", seed=1))
G = 18944
h = backend.createComputeHandle(ctx, lm, G, 5, 5)
games = backend.Games(ctx, G, 5, 5, 4); games.reset(seed=1, autoRefill=True)
for rep in range(3):
    games.run(h, 2)
    out = np.zeros(64 + 2 * 48 * 8, np.int64)
    check(lib().kc_handle_trunk_probe(h._p, ptr(out)))
    t0 = out[0]
    print("rel. to last MMA issue of layer 5:", {k: int(out[i] - t0) for k, i in (("issuer starts waiting", 6), ("epilogue sees ACC", 1), ("first tmem_ld done", 2),
          ("chunk 0 published", 3), ("issuer wait over", 4), ("first MMAs of layer 6 issued", 5), ("last chunk published", 7))},
          "layer 6 chunk issue times:", [int(out[i] - t0) for i in range(8, 16)])
    t1 = out[16]
    print("   item boundary rel. to head conv issue:", {k: int(out[i] - t1) for k, i in (("head epilogue starts", 17), ("TMEM released", 18), ("head epilogue ends", 19),
          ("issuer starts item 1", 20), ("layer 0 issued", 21), ("epilogue sees layer 0", 22), ("issuer has chunk 0 of layer 1", 23))},
          "item 0 total (layer 5 ref -> head conv issue)", int(t1 - t0))
    print("   head phases rel. to head conv issue:", {k: int(out[i] - t1) for k, i in (("start", 24), ("pooling done", 25), ("TMEM released, synced", 26),
          ("pooled matmuls done", 27), ("synced", 28), ("v3 done", 29), ("end", 19))})
    print("   head pooling detail rel. to head conv issue:", [int(out[i] - t1) for i in range(30, 38)])
    print("   first pool call: entry, after barrier A, after stores + barrier B:", [int(out[i] - t1) for i in range(40, 43)])
