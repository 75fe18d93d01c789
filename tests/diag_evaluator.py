"""Diagnostic (GPU box): throughput of the evaluator front end with native client threads (tests/cpp/bench_evaluator.cpp) for a few
client / batch / server settings; prints one JSON line per run.  python tests/diag_evaluator.py [net] [out.jsonl]"""
import json
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend, modeldesc  # noqa: E402

net = sys.argv[1] if len(sys.argv) > 1 else "b10c128"
out = sys.argv[2] if len(sys.argv) > 2 else None
exe = os.path.join(ROOT, "katacoffee_b200", "host", "bench_evaluator")
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, net + ".bin.gz")
    backend.writeModelFile(modeldesc.Model(net, seed=11), path)
    runs = [["--clients", "16", "--rows", "400000", "--batch", "18944", "--servers", "2", "--chunk", "2368"],
            ["--clients", "8", "--rows", "400000", "--batch", "4096", "--servers", "2", "--chunk", "512"],
            ["--clients", "8", "--rows", "400000", "--batch", "4096", "--servers", "1", "--chunk", "512"],
            ["--clients", "16", "--rows", "20000", "--batch", "64", "--servers", "2", "--single"],
            ["--clients", "8", "--rows", "400000", "--batch", "4096", "--servers", "2", "--chunk", "512", "--cache", "20", "--repeat", "50000"]]
    for r in runs:
        p = subprocess.run([exe, path] + r, capture_output=True, text=True, timeout=300)
        line = p.stdout.strip().splitlines()[-1] if p.stdout.strip() else json.dumps({"error": p.stderr[-300:]})
        rec = json.loads(line)
        rec["net"] = net
        print(json.dumps(rec), flush=True)
        if out:
            with open(out, "a") as f:
                f.write(json.dumps(rec) + "\n")
