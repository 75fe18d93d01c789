"""Diagnostic (GPU box): blocking single-row clients against 1 / 2 / 4 / 8 server threads (tests/cpp/bench_evaluator.cpp --single)."""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend, modeldesc  # noqa: E402

net = sys.argv[1] if len(sys.argv) > 1 else "b10c128"
exe = os.path.join(ROOT, "katacoffee_b200", "host", "bench_evaluator")
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, net + ".bin.gz")
    backend.writeModelFile(modeldesc.Model(net, seed=11), path)
    for clients, servers in ((16, 1), (16, 4), (16, 8), (64, 4), (64, 8), (256, 8)):
        p = subprocess.run([exe, path, "--clients", str(clients), "--rows", str(max(2000, 160000 // clients)), "--batch", "64", "--servers", str(servers), "--single"],
                           capture_output=True, text=True, timeout=120)
        rec = json.loads(p.stdout.strip().splitlines()[-1]) if p.stdout.strip() else {"error": p.stderr[-300:]}
        rec["net"] = net
        print(json.dumps(rec), flush=True)
