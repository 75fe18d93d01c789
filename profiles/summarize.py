#!/usr/bin/env python3
"""Turns gpurun_out/{launches.csv, prof_*.ncu-rep} into the committed summaries under profiles/."""
import csv, collections, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = os.path.join(ROOT, "profiles")
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_subpipe_hmma_cycles_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_bytes.sum", "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "dram__cycles_active", "sm__pipe_tensor_cycles_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__cycles_active.avg", "lts__t_sector_hit_rate.pct",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "gpc__cycles_elapsed.avg.per_second", "lts__t_sectors.avg.pct", "l1tex__throughput.avg.pct", "smsp__warp_issue_stalled",
        "smsp__average_warps_issue_stalled", "sm__pipe_fp64_cycles_active.avg", "l1tex__t_sector_hit_rate"]
path = os.path.join(ROOT, "gpurun_out", "launches.csv")
if os.path.exists(path):
    rows = [l for l in open(path) if l.startswith('"')]
    agg = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(rows))):
        k = (r["Kernel Name"].split("(")[0], r["Grid Size"], r["Block Size"])
        agg.setdefault(k, []).append(float(r["Metric Value"]))
    total = sum(sum(v) for v in agg.values())
    with open(os.path.join(out, f"{tag}_launches_summary.txt"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-selfplay\n")
        f.write("# kernel | grid | block | launches | mean us | share of profiled GPU time\n")
        for (name, grid, block), v in agg.items():
            f.write(f"{name} | {grid} | {block} | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/total*100:.1f}%\n")
    print(open(os.path.join(out, f"{tag}_launches_summary.txt")).read())
for rep in sorted(os.listdir(os.path.join(ROOT, "gpurun_out"))):
    if not rep.endswith(".ncu-rep"):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    if len(rd) < 3:
        continue
    hdr, units = rd[0], rd[1]
    with open(os.path.join(out, f"{tag}_{rep[:-8]}_metrics.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none, {rep}; one column per captured launch\n")
        for i, h in enumerate(hdr):
            if any(h.startswith(k) for k in KEYS) or h in ("Kernel Name",):
                f.write(f"{h} [{units[i]}]: " + " | ".join(r[i] for r in rd[2:]) + "\n")
    print(open(os.path.join(out, f"{tag}_{rep[:-8]}_metrics.txt")).read()[:6000])

# per-launch DRAM traffic of each captured kernel -> profiles/<tag>_traffic.json (bench.py's roofline.traffic reads it)
import json
traffic = {}
for rep in sorted(os.listdir(os.path.join(ROOT, "gpurun_out"))):
    if not rep.endswith(".ncu-rep"):
        continue
    raw = subprocess.run(["ncu", "-i", os.path.join(ROOT, "gpurun_out", rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(raw)))
    if len(rd) < 3:
        continue
    hdr, units = rd[0], rd[1]
    def col(name):
        i = hdr.index(name)
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[units[i]]
        return float(rd[2][i]) * scale
    kname = rd[2][hdr.index("Kernel Name")].split("(")[0]
    grid = rd[2][hdr.index("Grid Size")] if "Grid Size" in hdr else ""
    traffic[rep[:-8]] = {"kernel": kname, "grid": grid, "dram_bytes_read": col("dram__bytes_read.sum"), "dram_bytes_write": col("dram__bytes_write.sum"),
                         "gpu_time_ns": float(rd[2][hdr.index("gpu__time_duration.sum")]) * {"ns": 1, "us": 1e3, "ms": 1e6, "msecond": 1e6, "usecond": 1e3, "nsecond": 1}.get(units[hdr.index("gpu__time_duration.sum")], 1)}
with open(os.path.join(out, f"{tag}_traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print(json.dumps(traffic, indent=1))
