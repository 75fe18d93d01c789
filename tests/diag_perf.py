"""Diagnostic (not a test): quick device timing of the b10c128 trunk kernel and the rules+features kernel."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend, modeldesc

ctx = backend.createComputeContext(0)
net = sys.argv[1] if len(sys.argv) > 1 else "b10c128"
W = H = int(sys.argv[2]) if len(sys.argv) > 2 else 5
model = modeldesc.Model(net, seed=1)
lm = backend.LoadedModel(ctx, model)
fl = modeldesc.flops_per_eval(net, W * H)
for G in (4736, 18944):
    h = backend.createComputeHandle(ctx, lm, G, W, H)
    games = backend.Games(ctx, G, W, H, 4)
    games.reset(seed=1, autoRefill=True)
    games.run(h, 3)
    h.trunkTime()
    for rep in range(2):
        t0 = time.time(); games.run(h, 20); dt = time.time() - t0
        ms, cnt = h.trunkTime()
        print(f"{net} {W}x{H} G={G}: wall {dt*1e3/20:.3f} ms/ply, trunk kernel {ms/cnt:.3f} ms -> {G/(ms/cnt)*1e3/1e6:.3f} M evals/s, {G*fl/(ms/cnt*1e-3)/1e12:.1f} TFLOP/s")
    games.close(); h.close()
g = backend.Games(ctx, 65536, 5, 5, 4); g.reset(seed=1, autoRefill=True); g.run(None, 5); g.run(None, 50)
print(f"rules+features fp32 NCHW 65536 games: kernel {g.lastKernelMs():.4f} ms/ply -> {65536/g.lastKernelMs()*1e3/1e9:.3f} G steps/s, {65536*1572/g.lastKernelMs()*1e3/1e9:.0f} GB/s algorithmic")
# batched tree search: per-iteration cost on top of the trunk
if net == "b10c128" and W == 5:
    for G, V in ((18944, 64),):
        h = backend.createComputeHandle(ctx, lm, G, W, H)
        s = backend.Search(ctx, h, G, W, H, 4, maxVisits=V, temperaturePlies=30, autoRefill=True)
        s.reset(seed=1)
        s.play(1)
        h.trunkTime()
        st, _, ms = s.play(2)
        tms, tcnt = h.trunkTime()
        print(f"search G={G} V={V}: {ms/2/V:.3f} ms per iteration (trunk {tms/tcnt:.3f} ms), {st.movesPlayed/(ms*1e-3):.0f} moves/s at {V} visits, "
              f"{st.visits/(ms*1e-3)/1e6:.3f} M visits/s, net evals {st.netEvals/st.visits:.3f} of visits, games finished {st.gamesFinished}")
        s.close(); h.close()
