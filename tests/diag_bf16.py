"""Diagnostic (not a test): bf16 trunk error statistics vs the fp32 oracle + quick timing."""
import sys, os, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kc_oracle
from katacoffee_b200 import backend, modeldesc

ctx = backend.createComputeContext(0)
for net, W, H, n in (("b2c32", 5, 5, 512), ("b6c96", 5, 5, 512), ("b10c128", 5, 5, 512), ("b6c96", 6, 6, 256)):
    model = modeldesc.Model(net, seed=5)
    om = kc_oracle.Model(model)
    recs, pl, gl = kc_oracle.playout_run(W, H, 4, 3, 0, 64, threads=8)
    sel = np.random.default_rng(3).permutation(len(recs))[:n]
    planes, glob = pl[sel], gl[sel].reshape(-1, 1)
    legal = recs["legal"][sel]
    nextpla = (recs["status"][sel] >> 11) & 3
    ep, ev, em, eo = om.forward(planes, glob, W, H, mode=0, threads=8)
    lm = backend.LoadedModel(ctx, model)
    h = backend.createComputeHandle(ctx, lm, n, W, H)
    p, v, m, o = backend.getOutput(h, planes, glob, None)
    def st(a, b): d = np.abs(a - b); return f"max {d.max():.4f} mean {d.mean():.5f} | ref std {b.std():.3f} absmax {np.abs(b).max():.3f}"
    print(net, W, H, "policy", st(p, ep)); print("   value", st(v, ev)); print("   misc", st(m, em)); print("   own", st(o, eo))
    # post-processed
    worst_p = worst_v = 0
    for i in range(n):
        if (recs["status"][sel][i] >> 8) & 1 or not legal[i].any(): continue
        a = kc_oracle.postprocess(p[i], legal[i], v[i], m[i], int(nextpla[i]))
        b = kc_oracle.postprocess(ep[i], legal[i], ev[i], em[i], int(nextpla[i]))
        worst_p = max(worst_p, np.abs(a[0] - b[0]).max()); worst_v = max(worst_v, np.abs(a[1] - b[1]).max())
    print(f"   post-processed: policy prob max err {worst_p:.5f}, win/loss prob max err {worst_v:.5f}")
    h.close(); lm.close()

# timing: b10c128 through the device-resident path
model = modeldesc.Model("b10c128", seed=1)
lm = backend.LoadedModel(ctx, model)
for G in (1184, 4736, 18944):
    h = backend.createComputeHandle(ctx, lm, G, 5, 5)
    games = backend.Games(ctx, G, 5, 5, 4)
    games.reset(seed=1, autoRefill=True)
    games.run(h, 3)
    h.trunkTime()
    t0 = time.time(); st = games.run(h, 10); dt = time.time() - t0
    ms, cnt = h.trunkTime()
    fl = modeldesc.flops_per_eval("b10c128", 25)
    print(f"G={G}: wall {dt*1e3/10:.3f} ms/ply, trunk kernel {ms/cnt:.3f} ms -> {G/(ms/cnt)*1e3/1e6:.3f} M evals/s, {G*fl/(ms/cnt*1e-3)/1e12:.1f} TFLOP/s")
    games.close(); h.close()
g = backend.Games(ctx, 65536, 5, 5, 4); g.reset(seed=1, autoRefill=True); g.run(None, 5); t0=time.time(); g.run(None, 50); dt=time.time()-t0
print(f"rules+features fp32 NCHW 65536 games: kernel {g.lastKernelMs():.4f} ms/ply -> {65536/g.lastKernelMs()*1e3/1e9:.3f} G steps/s, {65536*1572/g.lastKernelMs()*1e3/1e9:.0f} GB/s algorithmic")
