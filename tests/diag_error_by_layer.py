"""Diagnostic (not a test): where the reduced-precision error of the tensor-core path comes from.

CPU only.  Runs the oracle's operand-rounding emulation (oracle/ko_net.cpp, KO_MODE_EMUL) for several
activation / weight formats against the fp32 oracle on random-legal positions and writes, per layer
(trunk after the initial conv, after each block, after trunkTipBN) and for every output, the max and
rms error: profiles/r02_bf16_error_by_layer.json.
Usage: python tests/diag_error_by_layer.py [n_positions]
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import kc_oracle  # noqa: E402
from katacoffee_b200 import modeldesc  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 96
FORMATS = [("bf16", "bf16"), ("fp16", "bf16"), ("bf16", "fp32"), ("fp32", "bf16"), ("fp16", "fp16")]
out = {"how": "oracle emulation: conv operands rounded (activations, weights), fp32 accumulate, vs the fp32 oracle; "
              f"{N} random-legal positions per net", "nets": []}
for net, W, H, cal in (("b10c128", 5, 5, "rms"), ("b15c192", 6, 6, "rms"), ("b10c128", 5, 5, "full"), ("b6c96", 5, 5, "rms")):
    model = modeldesc.Model(net, seed=5, calibrate=cal)
    om = kc_oracle.Model(model)
    recs, pl, gl = kc_oracle.playout_run(W, H, 4, 3, 0, 64, threads=8)
    sel = np.random.default_rng(3).permutation(len(recs))[:N]
    planes, glob = pl[sel], gl[sel].reshape(-1, 1)
    ref = kc_oracle.forward_trace(om, planes, glob, W, H, mode=0)
    entry = {"net": net, "board": f"{W}x{H}", "calibrate": cal, "ref_logit_std": {k: float(np.std(v)) for k, v in zip(("policy", "value", "misc", "own"), ref[:4])},
             "formats": []}
    for af, wf in FORMATS:
        got = kc_oracle.forward_trace(om, planes, glob, W, H, mode=kc_oracle.mode_emul(af, wf))
        layers = []
        for li in range(ref[4].shape[0]):
            d = got[4][li] - ref[4][li]
            name = "initial" if li == 0 else ("tip" if li == ref[4].shape[0] - 1 else f"block{li - 1}")
            layers.append({"layer": name, "max": float(np.abs(d).max()), "rms": float(np.sqrt((d * d).mean())), "ref_rms": float(np.sqrt((ref[4][li] ** 2).mean()))})
        outs = {k: {"max": float(np.abs(a - b).max()), "rms": float(np.sqrt(((a - b) ** 2).mean()))} for k, a, b in zip(("policy", "value", "misc", "own"), got[:4], ref[:4])}
        entry["formats"].append({"activations": af, "weights": wf, "outputs": outs, "layers": layers})
        print(f"{net} {W}x{H} {cal:4s} act {af} w {wf}: " + " ".join(f"{k} {v['max']:.2e}" for k, v in outs.items()) + f" | tip max {layers[-1]['max']:.2e}", flush=True)
    out["nets"].append(entry)
with open(os.path.join(ROOT, "profiles", "r02_bf16_error_by_layer.json"), "w") as f:
    json.dump(out, f, indent=1)
