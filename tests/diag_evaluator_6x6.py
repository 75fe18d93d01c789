"""Diagnostic (GPU box): the evaluator front end on 6x6 k=4 with b15c192 (BASELINE configs[4]) through tests/cpp/bench_evaluator.cpp."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from katacoffee_b200 import backend, modeldesc  # noqa: E402

exe = os.path.join(ROOT, "katacoffee_b200", "host", "bench_evaluator")
with tempfile.TemporaryDirectory() as d:
    path = os.path.join(d, "b15c192.bin")
    backend.writeModelFile(modeldesc.Model("b15c192", seed=11), path)
    p = subprocess.run([exe, path, "--size", "6", "--clients", "16", "--rows", "60000", "--batch", "8192", "--servers", "2", "--chunk", "1024"], capture_output=True, text=True, timeout=60)
    print(p.stdout.strip() or p.stderr[-300:])
