"""Summarises gpurun_out/*.ncu-rep and the ncu launch list into profiles/ (run here, no GPU needed):
    python tools/ncu_summary.py r01
writes profiles/<round>_ncu_<kernel>.csv (selected metrics), profiles/<round>_launch_summary.csv and
profiles/traffic.json (DRAM bytes per launch of the dominant kernels, read by bench.py)."""
import csv, glob, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
traffic = {}
fp64_pipe = None
reps = {os.path.basename(r)[len(f"prof_{rnd}_"):-len(".ncu-rep")]: r for r in glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_{rnd}_*.ncu-rep"))}
raws = {os.path.basename(r)[len(f"prof_{rnd}_"):-len(".raw.csv")]: r for r in glob.glob(os.path.join(ROOT, "gpurun_out", f"prof_{rnd}_*.raw.csv"))}
for name in sorted(set(reps) | set(raws)):
    if name in raws and os.path.getsize(raws[name]) > 0:       # the raw page exported on the GPU box
        out = open(raws[name]).read()
    else:
        out = subprocess.run(["ncu", "-i", reps[name], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        continue
    hdr, units, vals = rows[0], rows[1], rows[2]
    kname = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else name
    with open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_{name}.csv"), "w") as f:
        f.write(f"# {kname}\n")
        d = {}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                f.write(f"{h},{u},{v}\n")
                d[h] = (u, v)
    def to_bytes(key):
        u, v = d.get(key, ("byte", "0"))
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        return float(v.replace(",", "")) * mult
    traffic[name] = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
    if name == "pairwise_cells_kernel" and "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active" in d:
        fp64_pipe = float(d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][1].replace(",", ""))
    print(name, kname[:60], "dram bytes/launch", traffic[name])
lst = os.path.join(ROOT, "gpurun_out", f"launches_{rnd}.csv")
if os.path.exists(lst):
    agg = {}
    with open(lst) as f:
        lines = [l for l in f if not l.startswith("==")]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0]
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v_us = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v_us
    tot = sum(a[1] for a in agg.values())
    with open(os.path.join(ROOT, "profiles", f"{rnd}_launch_summary.csv"), "w") as f:
        f.write("kernel,launches,total_us,avg_us,share_pct\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k},{a[0]},{a[1]:.1f},{a[1]/a[0]:.2f},{100*a[1]/tot:.1f}\n")
    print("launch summary:", len(agg), "kernels, total", tot, "us")
# bench.py keys
tj = os.path.join(ROOT, "profiles", "traffic.json")
old = json.load(open(tj)) if os.path.exists(tj) else {}
m = {"spmv": "spmv_tile_kernel", "spmv_solver_order": "spmv_tile_kernel", "pairwise": "pairwise_cells_kernel", "rate_table": "rate_rows_kernel", "pcg_solve": "pcg_persistent_kernel"}
old["tiled_1M"] = {k: traffic[v] for k, v in m.items() if v in traffic}
if fp64_pipe is not None:
    old["pairwise_fp64_pipe_pct"] = fp64_pipe
old["_source"] = f"ncu --set full, one launch each, profiles/{rnd}_ncu_*.csv (dram__bytes_read.sum + dram__bytes_write.sum)"
json.dump(old, open(tj, "w"), indent=1)
