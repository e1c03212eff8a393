"""Turns the ncu reports / launch lists brought back in gpurun_out/ into the tracked summaries under profiles/."""
import csv, io, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "smsp__inst_executed.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "smsp__average_warp_latency_per_inst_issued.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]

def raw(rep, tag):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    with open(os.path.join(OUT, f"{tag}_ncu_raw_selected.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["launch", "kernel", "metric", "unit", "value"])
        for li, r in enumerate(rows[2:]):
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    w.writerow([li, r[kn][:60], k, units[i], r[i]])
    print("wrote", tag)

def launches(path, tag):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = {}
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        v = v / 1000 if r[mu] == "ns" else (v * 1000 if r[mu] == "ms" else v)
        a = agg.setdefault(r[kn].split("(")[0], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    with open(os.path.join(OUT, f"{tag}_launch_summary.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "avg_us", "share_pct"])
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k, v[0], round(v[1], 1), round(v[1] / v[0], 2), round(100 * v[1] / tot, 2)])
    print("wrote", tag, "launch summary")

if __name__ == "__main__":
    g = os.path.join(ROOT, "gpurun_out")
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r2"
    for rep, tag in ((f"prof_{rnd}_frame.ncu-rep", f"{rnd}_register_frame"), (f"prof_{rnd}_nn.ncu-rep", f"{rnd}_nn_query")):
        if os.path.exists(os.path.join(g, rep)):
            raw(os.path.join(g, rep), tag)
    if os.path.exists(os.path.join(g, f"launches_{rnd}.csv")):
        launches(os.path.join(g, f"launches_{rnd}.csv"), f"{rnd}_bench")
