"""Builds libkatacoffee_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libkatacoffee_b200.so")
SOURCES = ["zobrist.cpp", "modelfile.cpp", "sgf.cpp", "npzwrite.cpp", "evaluator.cpp", "selfplay.cpp", "games.cu", "net_fp32.cu", "net_bf16.cu", "search.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build(force=False, verbose=False):
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source_mtime():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    logs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc] + NVCC_FLAGS + ["-x", "cu", "-dc" if False else "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        logs.append(f"==== {src} ====\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(logs))
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-lcudart", "-lz", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    with open(os.path.join(CSRC, "build.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


def build_host(force=False):
    """The C++ nninterface backend (host/b200backend.cpp) as a shared library + its C++ test driver."""
    host = os.path.join(HERE, "host")
    lib = os.path.join(HERE, "libkc_b200backend.so")
    exe = os.path.join(host, "test_b200backend")
    src = os.path.join(host, "b200backend.cpp")
    test = os.path.join(HERE, "..", "tests", "cpp", "test_b200backend.cpp")
    bench_src = os.path.join(HERE, "..", "tests", "cpp", "bench_evaluator.cpp")
    bench_exe = os.path.join(host, "bench_evaluator")   # native throughput driver of the evaluator front end (bench.py --evaluator)
    go_src = os.path.join(HERE, "..", "tests", "cpp", "bench_getoutput.cpp")
    go_exe = os.path.join(host, "bench_getoutput")      # native driver of NeuralNet::getOutput itself (bench.py "batch1024")
    nneval_src = os.path.join(host, "b200nneval.cpp")   # class NNEvaluator (nneval.h) over kc_evaluator_*
    nneval_test = os.path.join(HERE, "..", "tests", "cpp", "test_b200nneval.cpp")
    nneval_exe = os.path.join(host, "test_b200nneval")
    newest = max(os.path.getmtime(p) for p in (src, test, bench_src, go_src, nneval_src, nneval_test, os.path.join(host, "b200nneval.h"), os.path.join(host, "reftypes.h"),
                                               os.path.join(HERE, "..", "include", "katacoffee_b200.h")))
    outs = (lib, exe, bench_exe, nneval_exe, go_exe)
    if not force and all(os.path.exists(p) for p in outs) and min(os.path.getmtime(p) for p in outs) >= newest:
        return lib, exe
    inc = ["-I" + os.path.join(HERE, "..", "include"), "-I" + host]
    subprocess.run(["g++", "-std=c++17", "-O2", "-fopenmp", "-fPIC", "-shared", "-Wall"] + inc + [src, nneval_src, "-o", lib, "-L" + HERE, "-lkatacoffee_b200",
                    "-Wl,-rpath,$ORIGIN"], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall"] + inc + [nneval_test, "-o", nneval_exe, "-L" + HERE, "-lkc_b200backend", "-lkatacoffee_b200",
                    "-Wl,-rpath," + HERE, "-lpthread"], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-fopenmp", "-Wall"] + inc + [test, "-o", exe, "-L" + HERE, "-lkc_b200backend", "-lkatacoffee_b200",
                    "-Wl,-rpath," + HERE], check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall"] + inc + [bench_src, "-o", bench_exe, "-L" + HERE, "-lkatacoffee_b200", "-Wl,-rpath," + HERE, "-lpthread"],
                   check=True)
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall"] + inc + [go_src, "-o", go_exe, "-L" + HERE, "-lkc_b200backend", "-lkatacoffee_b200",
                    "-Wl,-rpath," + HERE], check=True)
    return lib, exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
