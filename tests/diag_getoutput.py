import tempfile, os, subprocess, sys
sys.path.insert(0, ".")
from katacoffee_b200 import backend, modeldesc
d = tempfile.mkdtemp()
path = os.path.join(d, "m.bin.gz")
backend.writeModelFile(modeldesc.Model("b10c128", seed=11), path)
import json
for env in ({}, {"KC_FORWARD_ROWS_THREADS": "2"}, {"KC_FORWARD_ROWS_THREADS": "3"}, {"KC_FORWARD_ROWS_THREADS": "6"}, {"KC_FORWARD_ROWS_THREADS": "8"}):
    e = dict(os.environ); e.update(env)
    p = subprocess.run(["katacoffee_b200/host/bench_getoutput", path, "--size", "5", "--batches", "1024,4736,18944", "--reps", "300"], capture_output=True, text=True, env=e)
    try:
        r = json.loads(p.stdout.strip().splitlines()[-1])
        print(env, [(b["rows_per_call"], round(b["evals_per_s"] / 1e6, 3), b["ms_p50"], b["ms_p99"]) for b in r["batches"]])
    except Exception as ex:
        print(env, "ERR", p.stdout[-300:], p.stderr[-300:])
