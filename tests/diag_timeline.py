"""Diagnostic (not a test): whole-item timeline of the trunk kernel (KC_TRUNK_PROBE=1): per layer and tile, when the MMA issuer
reached the layer, when its first input chunk was there, when the layer was issued, when the epilogue saw the accumulator and when
it was done -- CTA 0, second item.  Usage: python tests/diag_timeline.py [net] [W] [out.json]"""
import json, os, sys
os.environ["KC_TRUNK_PROBE"] = "1"
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend, modeldesc
from katacoffee_b200.capi import lib, check, ptr
net = sys.argv[1] if len(sys.argv) > 1 else "b10c128"
W = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ctx = backend.createComputeContext(0)
model = modeldesc.Model(net, seed=1)
lm = backend.LoadedModel(ctx, model)
G = 18944 if W == 5 else 148 * 3 * 32
h = backend.createComputeHandle(ctx, lm, G, W, W)
games = backend.Games(ctx, G, W, W, 4); games.reset(seed=1, autoRefill=True)
games.run(h, 3)
WORDS = 64 + 2 * 48 * 8
out = np.zeros(WORDS, np.int64)
check(lib().kc_handle_trunk_probe(h._p, ptr(out)))
tl = out[64:64 + 768].reshape(2, 48, 8)
nl = 2 * model.num_blocks + 2
t0 = tl[0, 0, 0]
rows = []
print(f"{net} {W}x{W}: per layer (clk rel. to tile 0 reaching layer 0 of its second item)")
print("tile layer | reach  first-chunk  issued | wait-chunk  issue-span | epi-sees-acc  epi-done | acc->epi  epi-span | gap to next layer's first chunk")
for t in range(2):
    for l in range(nl):
        r = tl[t, l]
        if r[0] == 0:
            continue
        reach, chunk, issued, esee, edone = [int(x - t0) for x in r[:5]]
        if l == 0:
            chunk = reach
        nxt = int(tl[t, l + 1, 1] - t0) if l + 1 < nl and tl[t, l + 1, 1] else None
        rows.append(dict(tile=t, layer=l, reach=reach, first_chunk=chunk, issued=issued, epi_sees_acc=esee, epi_done=edone))
        print(f"{t:3d} {l:5d} | {reach:7d} {chunk:7d} {issued:7d} | {chunk - reach:6d} {issued - chunk:7d} | {esee:7d} {edone:7d} | {esee - issued:6d} {edone - esee:6d} | "
              f"{'' if nxt is None else nxt - issued}")
for t in range(2):
    if tl[t, 0, 0] and tl[t, nl - 1, 2]:
        print(f"tile {t}: item span (reach layer 0 -> head conv issued) {int(tl[t, nl - 1, 2] - tl[t, 0, 0])} clk; "
              f"sum of issue spans {sum(r['issued'] - r['first_chunk'] for r in rows if r['tile'] == t)}; sum of chunk waits {sum(r['first_chunk'] - r['reach'] for r in rows if r['tile'] == t)}")
if len(sys.argv) > 3:
    json.dump(rows, open(sys.argv[3], "w"), indent=0)
