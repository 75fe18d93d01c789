#!/usr/bin/env python3
"""bench.py -- the hot path's throughput on B200 (and the CPU reference arm beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json: "NN evals/s ... 5x5 k=4, b10c128"): every GPU owns G concurrent 5x5 k=4
Coffee games.  One STEP = one ply of the fused leaf-evaluation hot path for all G games:
  rules kernel (random-legal move, win/draw, sit-hash, legal mask, auto-refill of finished games)
  -> V1 planes written as fp16 straight into the trunk's input tiles
  -> b10c128 forward (one persistent tcgen05 kernel, fp16 operands, fp32 accumulation) -> policy / value / misc / ownership logits,
so a step evaluates G positions.  `value` = positions evaluated per second, whole job, inputs
resident in HBM, timed per step with CUDA events on the launching stream (max over ranks).
`e2e` = the same metric through the C ABI's getOutput entry point (kc_forward, called from Python via ctypes) with HOST rows in
pinned memory: H2D of the planes and D2H of the logits inside the timed region.  `batch1024` = the C++ drop-in itself
(NeuralNet::getOutput of host/b200backend.cpp, native driver in a child process) at BASELINE configs[2]'s 1024 rows and at the
bench batch; `config_6x6` = the same measurement for BASELINE configs[4] (6x6 k=4, b15c192) at reduced steps.
`--impl reference` times the CPU restatement of the reference's own path (oracle: mailbox rules +
fillRowV1 + Winograd/GEMM fp32 forward as the Eigen backend does) on all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NET = "b10c128"
W = H = 5
WINLEN = 4
SEED = 20261018
GAMES_PER_GPU = 148 * 128            # 18944: 8 boards per CTA work item -> 16 items per SM, no tail
L2_FLUSH_BYTES = 256 << 20           # > 126 MB L2, written between timed steps
SELFPLAY_STAGGER = 20                # self-play arm: games start spread over plies 0..19 (a 5x5 game lasts 7..24 plies)
BYTES_PER_STEP_FP32 = 1572           # rules+features algorithmic bytes per game-step (fp32 planes)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks and throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def wait_first(self, timeout_s=3.0):
        """nvidia-smi takes a few hundred ms to print its first row: the timed region (tens of ms) starts only once it is streaming."""
        t = time.time()
        while self.proc is not None and not self.rows and time.time() - t < timeout_s:
            time.sleep(0.01)

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.03] or [r for (t, r) in self.rows if t0 - 0.05 <= t <= t1 + 0.15]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); smax = float(f[1])
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def load_traffic(name):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the same
    bench command (profiles/summarize.py writes the file); None if the capture is absent."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None
    with open(files[-1]) as f:
        t = json.load(f).get(name)
    return None if t is None else t["dram_bytes_read"] + t["dram_bytes_write"]


def dist_setup(n_gpus):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def cpu_reference_run(positions_per_step, steps, warmup, threads):
    """The reference's own CPU path, restated (oracle): random-legal playouts with mailbox rules,
    fillRowV1 (NHWC, as the Eigen backend wants) and the Winograd/GEMM fp32 forward in batches of 4
    per thread (cpp/program/setup.cpp:299), over all host threads.  Returns evals/s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import kc_oracle
    from katacoffee_b200 import modeldesc
    kc_oracle.build()
    model = modeldesc.Model(NET, seed=1)
    om = kc_oracle.Model(model)
    games = max(8, positions_per_step // 16)
    per_step = []
    g0 = 0
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        recs, planes, glob = kc_oracle.playout_run(W, H, WINLEN, SEED, g0, games, planes=True, nhwc=True, threads=threads,
                                                   max_records=positions_per_step)
        n = len(recs)
        om.forward(planes, glob.reshape(-1, 1), W, H, symmetry=None, nhwc=True, mode=1, threads=threads)
        dt = time.perf_counter() - t0
        g0 += games
        if s >= warmup:
            per_step.append((n, dt))
    tot_n = sum(n for n, _ in per_step)
    tot_t = sum(t for _, t in per_step)
    return tot_n / tot_t, tot_t / len(per_step) * 1e3, tot_n // len(per_step)


def cpu_selfplay_run(threads, visits=24):
    """The reference's own search loop restated (oracle/ko_search.cpp: one game, one thread, one position per net call)
    with the fp32 direct-convolution forward, on all host threads at once.  Returns visits/s."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import kc_oracle
    from katacoffee_b200 import modeldesc
    om = kc_oracle.Model(modeldesc.Model(NET, seed=1))
    games = []
    for i in range(threads):
        g = kc_oracle.Game(W, H, WINLEN)
        for _ in range(i % 8):
            g.play(g.choose(SEED, i))
        games.append(g)
    done = [0] * threads

    def work(i):
        done[i] = int(kc_oracle.search_run(games[i], visits, model=om)["counters"][0])
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return sum(done) / (time.perf_counter() - t0)


def evaluator_arm(net, W, H):
    """The evaluator front end (kc_evaluator_*, the NNEvaluator of cpp/neuralnet/nneval.cpp for CPU search threads) with native client
    threads: tests/cpp/bench_evaluator.cpp in its own process (positions in host memory, results in host memory).  Never fatal."""
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "katacoffee_b200", "host", "bench_evaluator")
    try:
        from katacoffee_b200 import backend, modeldesc
        threads = min(os.cpu_count() or 1, 32)
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, net + ".bin.gz")
            backend.writeModelFile(modeldesc.Model(net, seed=11), path)
            p = subprocess.run([exe, path, "--size", str(W), "--clients", str(threads), "--rows", "300000", "--batch", "18944", "--servers", "2", "--chunk", "2368"],
                               capture_output=True, text=True, timeout=180)
        if p.returncode != 0:
            return {"error": (p.stderr or p.stdout)[-300:]}
        rep = json.loads(p.stdout.strip().splitlines()[-1])
        rep["unit"] = "positions/s"
        rep["api"] = "kc_evaluator_evaluate_many (NNEvaluator::evaluate): host positions in, post-processed NNOutput out, native client threads"
        return rep
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)[:300]}


def getoutput_arm(net, size, batches, reps=300):
    """NeuralNet::getOutput of the C++ drop-in (host/b200backend.cpp) driven natively: tests/cpp/bench_getoutput.cpp in its own process,
    NNResultBuf rows in host memory in, NNOutput logits in host memory out.  Never fatal."""
    import tempfile
    exe = os.path.join(ROOT, "katacoffee_b200", "host", "bench_getoutput")
    try:
        from katacoffee_b200 import backend, modeldesc
        with tempfile.TemporaryDirectory() as d:
            path = os.path.join(d, net + ".bin.gz")
            backend.writeModelFile(modeldesc.Model(net, seed=11), path)
            p = subprocess.run([exe, path, "--size", str(size), "--batches", ",".join(str(b) for b in batches), "--reps", str(reps)],
                               capture_output=True, text=True, timeout=240)
        if p.returncode != 0:
            return {"error": (p.stderr or p.stdout)[-300:]}
        return json.loads(p.stdout.strip().splitlines()[-1])
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)[:300]}


def config_6x6_arm(steps, warmup):
    """BASELINE configs[4] (6x6 k=4, b15c192) at reduced steps: this script again in a child process with --config 6x6, without the
    self-play / CPU / evaluator arms.  Returns the child's value / e2e / roofline / clocks.  Never fatal."""
    try:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--config", "6x6", "--steps", str(steps), "--warmup", str(warmup), "--no-cpu-baseline",
                            "--no-selfplay-plain", "--selfplay-moves", "2", "--no-evaluator", "--no-extra-configs"], capture_output=True, text=True, timeout=900)
        if p.returncode != 0:
            return {"error": (p.stderr or p.stdout)[-300:]}
        d = json.loads(p.stdout.strip().splitlines()[-1])
        return {k: d[k] for k in ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "config", "e2e", "getoutput_cpp", "clocks", "roofline",
                                  "roofline_rules_features", "selfplay_graph", "gpu_launches") if k in d}
    except Exception as e:   # noqa: BLE001
        return {"error": repr(e)[:300]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    positions = 768
    evals_s, ms_step, n = cpu_reference_run(positions, args.steps, args.warmup, threads)
    line = {
        "impl": "reference", "metric": "nn_evals_per_s", "value": evals_s, "unit": "evals/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.games),      # identical to our arm's: the CPU arm evaluates a bounded sample of it per step and reports a rate
        "cpu_baseline": {"value": evals_s, "unit": "evals/s", "cores": threads, "kind": "port",
                         "sample": f"{n} positions per step of the same workload: oracle mailbox rules + fillRowV1 + Winograd/GEMM fp32 {NET} forward, batch 4 per thread"},
        "e2e": {"value": evals_s, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


CONFIGS = {
    # --config: net, board, games per GPU, self-play stagger, rules+features bytes per game-step, BASELINE.json config
    "5x5": ("b10c128", 5, 148 * 128, 20, 1572, "configs[2] net on configs[1]-style concurrent random-legal games"),
    # 6x6 k=4 with b15c192: 3 boards per CTA work item -> 148 * 3 * 32 games, no tail
    "6x6": ("b15c192", 6, 148 * 3 * 32, 28, 2236, "configs[4]: non-default board size and the larger trunk"),
}


def select_config(name):
    global NET, W, H, GAMES_PER_GPU, SELFPLAY_STAGGER, BYTES_PER_STEP_FP32, CONFIG_NOTE
    NET, W, GAMES_PER_GPU, SELFPLAY_STAGGER, BYTES_PER_STEP_FP32, CONFIG_NOTE = CONFIGS[name]
    H = W


CONFIG_NOTE = CONFIGS["5x5"][5]


def workload_config(games_per_gpu):
    return {"workload": f"{W}x{H} k=4 Coffee leaf evaluation: rules step + V1 planes + {NET} forward per ply "
                        f"(BASELINE.json {CONFIG_NOTE})",
            "net": NET, "board": f"{W}x{H}", "win_len": WINLEN, "games_per_gpu": games_per_gpu, "seed": SEED,
            "weights": "random-init (modeldesc rms-calibrated)", "l2": f"{L2_FLUSH_BYTES >> 20} MiB scratch overwritten between timed steps"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="5x5", choices=sorted(CONFIGS), help="5x5: b10c128 on 5x5 k=4 (the headline metric); 6x6: b15c192 on 6x6 k=4 (BASELINE configs[4])")
    ap.add_argument("--games", type=int, default=0, help="games per GPU (default: the configuration's)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-selfplay", action="store_true")
    ap.add_argument("--no-selfplay-plain", action="store_true", help="only the selfplay1.cfg self-play arm, not the SearchParams() defaults one")
    ap.add_argument("--no-evaluator", action="store_true", help="skip the evaluator front-end arm (native client threads in a child process)")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the batch1024 (BASELINE configs[2]) and config_6x6 (configs[4]) objects")
    ap.add_argument("--visits", type=int, default=800, help="visits per move of the self-play arm (BASELINE config 4)")
    ap.add_argument("--nn-cache", type=int, default=24, help="nnCacheSizePowerOfTwo of the self-play arms (selfplay1.cfg:121 has 21 for its 128 games; "
                                                              "here 18,944 games share one cache; 0 = no NN cache)")
    ap.add_argument("--selfplay-moves", type=int, default=8, help="consecutive moves per game lane timed in the self-play arm")
    args = ap.parse_args()
    select_config(args.config)
    args.games = args.games or GAMES_PER_GPU
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import __graft_entry__ as ge
    rank, world, local = dist_setup(args.gpus)
    if rank == 0:
        ge.build()
    if world > 1:
        torch.distributed.barrier()
    from katacoffee_b200 import backend, modeldesc, shard
    G = args.games
    peaks = load_peaks()
    ctx = backend.createComputeContext(local)
    model = modeldesc.Model(NET, seed=1)
    lm = backend.LoadedModel(ctx, model)
    handle = backend.createComputeHandle(ctx, lm, G, W, H)
    games = backend.Games(ctx, G, W, H, WINLEN)
    games.reset(seed=SEED, firstGameId=shard.first_game_id(rank), autoRefill=True)   # shard: disjoint game ids per rank

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident arm (`value`) ----------------
    sampler = ClockSampler(local) if rank == 0 else None   # started before the warm-up: nvidia-smi must be streaming when the timed region begins
    if sampler:
        sampler.wait_first()
    barrier()
    games.runTimed(handle, args.warmup, L2_FLUSH_BYTES)
    handle.trunkTime()
    l0 = games.launchCount() + handle.launchCount()
    barrier()
    t0 = time.time()
    stats, ms_total = games.runTimed(handle, args.steps, L2_FLUSH_BYTES)
    t1 = time.time()
    barrier()
    clocks = sampler.stop(t0, t1) if sampler else None
    launches = games.launchCount() + handle.launchCount() - l0
    trunk_ms, trunk_n = handle.trunkTime()
    ms_max = shard.max_over_ranks(ms_total, "cuda")
    # the end-of-run NCCL reduce of statistics -- the only collective of the whole job
    counters = shard.reduce_stats([getattr(stats, f) for f in shard.STAT_FIELDS], "cuda")
    total_evals = G * args.steps * world
    value = total_evals / (ms_max * 1e-3)

    # BASELINE configs[2] as stated: 1024 rows per call.  Device-resident first (1024 games stepped and evaluated per ply) ...
    batch1024 = None
    if rank == 0 and world == 1 and not args.no_extra_configs and args.config == "5x5":
        h1k = backend.createComputeHandle(ctx, lm, 1024, W, H)
        g1k = backend.Games(ctx, 1024, W, H, WINLEN)
        g1k.reset(seed=SEED, autoRefill=True)
        g1k.runTimed(h1k, 10, L2_FLUSH_BYTES)
        h1k.trunkTime()
        _, ms1k = g1k.runTimed(h1k, 50, L2_FLUSH_BYTES)
        t1k_ms, t1k_n = h1k.trunkTime()
        batch1024 = {"rows_per_call": 1024, "net": NET, "board": f"{W}x{H}",
                     "device": {"evals_per_s": 1024 * 50 / (ms1k * 1e-3), "ms_per_step": ms1k / 50, "trunk_kernel_ms": t1k_ms / max(t1k_n, 1),
                                "what": "rules step + fp16 tiles + trunk per ply for 1024 resident games, CUDA events, L2 flushed between steps; "
                                        "1024 rows = 128 work items of 8 boards on 148 SMs: under one wave, latency-bound"}}
        g1k.close(); h1k.close()

    # rules+features alone (the HBM-bound kernel), fp32 NCHW planes + masks + hashes: secondary roofline
    games_rf = backend.Games(ctx, 65536, W, H, WINLEN)
    games_rf.reset(seed=SEED, autoRefill=True)
    games_rf.runTimed(None, 8, L2_FLUSH_BYTES)
    _, rf_ms = games_rf.runTimed(None, 24, L2_FLUSH_BYTES)
    rf_steps_s = 65536 * 24 / (rf_ms * 1e-3)
    games_rf.close()

    # ---------------- host-buffer arm (`e2e`) ----------------
    hw = W * H
    nbuf = 4
    pinned = [torch.empty((G, 15 * hw), dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    glob_h = torch.full((G, 1), float(WINLEN)).pin_memory()
    sym_h = np.zeros(G, np.int8)
    for i in range(nbuf):
        games.run(None, 2)
        p, _ = games.features(nhwc=False)
        pinned[i].numpy()[:] = p
    outs = (torch.empty((G, 4 * hw)).pin_memory().numpy(), torch.empty((G, 2)).pin_memory().numpy(),
            torch.empty((G, 2)).pin_memory().numpy(), torch.empty((G, hw)).pin_memory().numpy())
    for i in range(args.warmup):
        backend.getOutput(handle, pinned[i % nbuf].numpy(), glob_h.numpy(), sym_h, out=outs)
    barrier()
    te0 = time.perf_counter()
    for i in range(args.steps):
        backend.getOutput(handle, pinned[i % nbuf].numpy(), glob_h.numpy(), sym_h, out=outs)
    te = time.perf_counter() - te0
    barrier()
    e2e_value = total_evals / shard.max_over_ranks(te, "cuda")
    h2d = G * (15 * hw * 4 + 4 + 1)
    d2h = G * (4 * hw * 4 + 8 + 8 + hw * 4)

    # ---------------- self-play arm (BASELINE.json: "selfplay moves/s", 800 visits/move) ----------------
    # the same G games under the batched device tree search: per iteration every game descends to one leaf and the G
    # leaves are one batch through the hot path; a move = SELFPLAY_VISITS iterations
    def selfplay_arm(label, **search_kw):
        # kc_selfplay_run (csrc/selfplay.cpp): this rank's device as one native search pool.  Timed on the host clock around the
        # whole loop: `--selfplay-moves` consecutive moves per game lane with refill of finished games, the finished games'
        # training rows emitted on the device (k_emit_rows), read back and written as the reference's .npz files by the pool's
        # writer thread -- all inside the timed region.  One untimed move first: with tree re-use it builds the trees whose
        # chosen subtrees the timed moves re-use, as every move of a running self-play does (Search::makeMove).
        import shutil
        import tempfile
        out_dir = tempfile.mkdtemp(prefix="kc_selfplay_")
        chunk = max(1, min(4, args.selfplay_moves))
        try:
            barrier()
            tot, rep = backend.selfplayRun(model, [local], G, W, H, WINLEN, moves=args.selfplay_moves, movesPerChunk=chunk, warmupMoves=1,
                                           staggerPlies=SELFPLAY_STAGGER, maxRowsPerChunk=G * 3 * chunk, outputDir=out_dir, seed=SEED,
                                           firstGameId=shard.first_game_id(rank), maxVisits=args.visits, autoRefill=1, temperaturePlies=30,
                                           nnCacheSizePowerOfTwo=args.nn_cache, **search_kw)
            barrier()
            files = len(os.listdir(out_dir))
        finally:
            shutil.rmtree(out_dir, ignore_errors=True)
        sec = shard.max_over_ranks(rep.wallSeconds, "cuda")
        dev_ms = shard.max_over_ranks(rep.deviceMsMax, "cuda")
        sp = shard.reduce_stats([tot.movesPlayed, tot.visits, tot.netEvals, tot.terminalVisits, tot.gamesFinished, tot.batchRows, tot.transpositionHits,
                                 tot.catchUpVisits, rep.rowsWritten, rep.rowsDropped, rep.bytesWritten, rep.kernelLaunches, files, tot.nnCacheHits], "cuda")
        return {"metric": "selfplay_moves_per_s", "value": sp[0] / sec, "unit": "moves/s", "visits_per_move": args.visits,
                "games_per_gpu": G, "moves_timed_per_game": args.selfplay_moves,
                "timed_region": "host clock around kc_selfplay_run's move loop: search, move choice, refill of finished games, k_emit_rows, row read-back "
                                "and kc_training_write_npz (writer thread) all inside; max over ranks",
                "seconds": sec, "device_seconds": dev_ms * 1e-3, "moves_per_s_device_time_only": sp[0] / (dev_ms * 1e-3),
                "start_plies": f"game g starts at ply g mod {SELFPLAY_STAGGER} (random-legal prefix)", "visits_per_s": sp[1] / sec,
                "batch_rows_per_s": sp[5] / sec, "net_eval_fraction_of_visits": sp[2] / max(sp[1], 1),
                "transposition_fraction_of_visits": (sp[6] + sp[7]) / max(sp[1], 1),
                "ms_per_move_batch": sec * 1e3 / args.selfplay_moves, "games_finished": int(sp[4]),
                "nn_cache": f"2^{args.nn_cache} entries shared by the games of the GPU (nnCacheSizePowerOfTwo, selfplay1.cfg:121: 21 for 128 games); a hit is bit-identical to the evaluation" if args.nn_cache else "off",
                "nn_cache_hits_per_s": sp[13] / sec,
                "training_rows_written": int(sp[8]), "training_rows_dropped": int(sp[9]), "npz_files": int(sp[12]), "npz_bytes": int(sp[10]),
                "search": label, "kernel_launches": int(sp[11])}

    selfplay = selfplay_graph = None
    if not args.no_selfplay:
        # BASELINE.json configs[3]: graph search + subtree value bias (cpp/configs/training/selfplay1.cfg:180-183)
        selfplay_graph = selfplay_arm("lock-step PUCT per game with the search options of cpp/configs/training/selfplay1.cfg:144-185: useGraphSearch, "
                                      "subtreeValueBiasFactor 0.30 / WeightExponent 0.8 / FreeProp 0.8, shaped Dirichlet root noise 10.83 / 0.25, root policy "
                                      "temperature 1.25 -> 1.1, fpuParentWeightByVisitedPolicy pow 2, rootDesiredPerChildVisitsCoeff 2, cpuct 1.1, root FPU "
                                      "reduction 0, valueWeightExponent 0.5, rootNumSymmetriesToSample 4, useLcbForSelection (lcbStdevs 5, minVisitPropForLCB 0.15, "
                                      "useNonBuggyLcb), move choice from the full getPlaySelectionValues under the temperature schedule 0.75 -> 0.15 with prune 1, "
                                      "nnRandomize; tree re-use.  Nothing of that option set is left out (useNoisePruning / useUncertainty, built too, are off in self-play: setup.cpp:525,543)",
                                      useGraphSearch=True, subtreeValueBiasFactor=0.30, subtreeValueBiasWeightExponent=0.8, reuseTree=True,
                                      cpuctExploration=1.1, rootFpuReductionMax=0.0, rootNoiseEnabled=1, rootDirichletNoiseTotalConcentration=10.83,
                                      rootDirichletNoiseWeight=0.25, rootPolicyTemperature=1.1, rootPolicyTemperatureEarly=1.25,
                                      chosenMoveTemperatureHalflife=19.0, fpuParentWeightByVisitedPolicy=1, fpuParentWeightByVisitedPolicyPow=2.0,
                                      rootDesiredPerChildVisitsCoeff=2.0, valueWeightExponent=0.5,
                                      chosenMoveTemperatureEarly=0.75, chosenMoveTemperature=0.15, chosenMovePrune=1.0, nnRandomize=1,
                                      rootNumSymmetriesToSample=4, useLcbForSelection=1, lcbStdevs=5.0, minVisitPropForLCB=0.15, useNonBuggyLcb=1)
        if not args.no_selfplay_plain:
            selfplay = selfplay_arm("lock-step PUCT per game (SearchParams() defaults, valueWeightExponent 0), tree re-use, visit-proportional move choice",
                                    reuseTree=True)

    if rank == 0:
        flops = modeldesc.flops_per_eval(NET, hw)
        trunk_avg_ms = trunk_ms / max(trunk_n, 1)
        achieved = flops * G / (trunk_avg_ms * 1e-3) / 1e12
        # Which measured peak bounds the trunk kernel: every launch is timed alone between two events with an L2 flush before it, so
        # unless the timed region is long enough for the power cap to settle (>= 2 s) the burst figure is the honest denominator.
        region_s = ms_max * 1e-3
        sustained = region_s >= 2.0
        peak = peaks["bf16_tflops_sustained"] if sustained else peaks["bf16_tflops"]
        # rules+features: bytes one launch of the 8-ply kernel really moves per game-step = per ply planes 15*HW*4 + sit-hash 16 + legal
        # 4*LW + status 4 + move 2 (every ply has its own ring slot) + per launch (state 48 B read + 48 B written + global 4) / 8 plies
        rf_bytes = 15 * hw * 4 + 16 + 4 * ((4 * hw + 31) // 32) + 4 + 2 + (48 + 48 + 4) / 8
        rf_traffic = (load_traffic("prof_games_multi") or load_traffic("prof_games")) if args.config == "5x5" else None
        rf_launch_s = 8 * 65536 / rf_steps_s
        line = {
            "metric": "nn_evals_per_s", "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16",
            "data": "synthetic", "config": workload_config(G),
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "kc_forward, the C ABI's getOutput entry point, called from Python (ctypes) with pinned host rows in the batch layout; "
                           "the C++ NeuralNet::getOutput shim over scattered NNResultBuf rows is `getoutput_cpp`"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"kernel": "trunk_kernel (tcgen05 whole-net forward, fp16 operands, fp32 accumulation)", "bound": "tensor", "achieved": achieved,
                         "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         "frac_of_burst_peak": achieved / peaks["bf16_tflops"], "frac_of_sustained_peak": achieved / peaks["bf16_tflops_sustained"],
                         "traffic": load_traffic("prof_trunk") if args.config == "5x5" else None,
                         "peak_source": peaks["source"] + (" sustained (timed region %.2f s >= 2 s)" % region_s if sustained else
                                                           " burst (timed region %.2f s < 2 s: every launch timed alone after an L2 flush)" % region_s),
                         "timed_region_s": region_s, "sm_mhz_median": clocks["sm_mhz"] if clocks else None,
                         "flops_per_eval": flops, "evals_per_launch": G, "avg_launch_ms": trunk_avg_ms, "launches_timed": trunk_n},
            "roofline_rules_features": {"kernel": "games_multi_split_kernel (rules step + fp32 NCHW planes, 8 plies per launch, producer / consumer warps, "
                                                  "every ply's planes / masks / status / hashes / moves to its own slot of a 4-slot ring) at 65536 games",
                                        "bound": "hbm", "plies_per_launch": 8,
                                        "achieved": rf_steps_s * rf_bytes / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                        "frac": rf_steps_s * rf_bytes / 1e9 / peaks["hbm_gbs"], "game_steps_per_s": rf_steps_s,
                                        "bytes_per_game_step": rf_bytes, "bytes_per_launch": rf_bytes * 8 * 65536, "avg_launch_ms": rf_launch_s * 1e3,
                                        "traffic": rf_traffic,
                                        "frac_traffic": (rf_traffic / rf_launch_s / 1e9 / peaks["hbm_gbs"]) if rf_traffic else None,
                                        "survey_bytes_per_game_step": BYTES_PER_STEP_FP32},
            "selfplay": selfplay,
            "selfplay_graph": selfplay_graph,
            "stats": {"game_steps": int(counters[0]), "evals": int(counters[1]), "games_finished": int(counters[2]),
                      "black_wins": int(counters[3]), "white_wins": int(counters[4]), "draws": int(counters[5])},
        }
        if world == 1 and not args.no_cpu_baseline:
            threads = os.cpu_count() or 1
            # bounded sample: calibrate on a small batch, then ~10 s of CPU work
            r0, _, _ = cpu_reference_run(128, 1, 0, threads)
            sample = int(min(8192, max(256, r0 * 10)))
            r, _, n = cpu_reference_run(sample, 1, 0, threads)
            sp_visits_s = cpu_selfplay_run(threads)
            line["cpu_baseline"] = {"value": r, "unit": "evals/s", "cores": threads, "kind": "port",
                                    "selfplay_moves_per_s": sp_visits_s / args.visits, "selfplay_visits_per_s": sp_visits_s,
                                    "selfplay_sample": f"{threads} games x 24 visits, oracle search + fp32 {NET} forward, one position per call",
                                    "sample": f"{n} positions: oracle rules + fillRowV1 + Winograd/GEMM fp32 {NET} forward (Eigen-algorithm restatement), batch 4 per thread"}
        if world == 1 and not args.no_evaluator and W == H:
            line["evaluator"] = evaluator_arm(NET, W, H)
        if world == 1 and W == H:
            # the C++ drop-in itself at the bench batch (and, for the headline configuration, at BASELINE configs[2]'s 1024 rows)
            go = getoutput_arm(NET, W, [1024, G] if batch1024 is not None else [G])
            if "batches" in go:
                by_rows = {b["rows_per_call"]: b for b in go["batches"]}
                line["getoutput_cpp"] = dict(by_rows[G], api=go["api"])
                if batch1024 is not None:
                    batch1024["e2e"] = dict(by_rows[1024], api=go["api"])
            else:
                line["getoutput_cpp"] = go
        if batch1024 is not None:
            line["batch1024"] = batch1024
    for o in (games, handle, lm, ctx):
        o.close()
    if rank == 0:
        if world == 1 and args.config == "5x5" and not args.no_extra_configs:
            line["config_6x6"] = config_6x6_arm(max(10, min(args.steps, 20)), args.warmup)
        print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
