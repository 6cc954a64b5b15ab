"""Turn gpurun_out/*.ncu-rep + launch CSVs into the tracked summaries under profiles/.

    python tools/make_profile_summary.py r1

Reads gpurun_out/<tag>_full.ncu-rep and gpurun_out/<tag>_launches.csv (captured on the GPU box
with the commands quoted in profiles/README.md) and writes
    profiles/<tag>_launches.csv      every launch of our kernels with its device time
    profiles/<tag>_ncu_summary.txt   per-kernel metrics of the `ncu --set full` capture
    profiles/traffic.json            DRAM bytes per launch of the two dominant kernels (bench.py reads it)
"""
import csv, json, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
rep = os.path.join(ROOT, "gpurun_out", tag + "_full.ncu-rep")
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "sm__cycles_elapsed.avg.per_second"]
ik = hdr.index("Kernel Name")
idx = [(w, hdr.index(w)) for w in want if w in hdr]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
traffic = {}
def gb(v, unit):
    v = float(v)
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
with open(os.path.join(ROOT, "profiles", tag + "_ncu_summary.txt"), "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on (B200, per launch; times are cold-cache and\n"
            "# serialised under the profiler -- use bench.py's CUDA-event numbers for rates)\n")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[ik])
        f.write("\n== %s\n" % r[ik][:150])
        for w, i in idx:
            f.write("  %-80s %s %s\n" % (w, r[i], units[i]))
        rd = gb(r[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_read.sum")])
        wr = gb(r[hdr.index("dram__bytes_write.sum")], units[hdr.index("dram__bytes_write.sum")])
        if "rollout_info8_tma_kernel<2, 0, 1, RingStd, 0" in r[ik]:
            traffic["rollout_cfg4"] = rd + wr
        if "rollout_info8_tma_kernel<2, 0, 1, RingStd, 1" in r[ik]:
            traffic["rollout_cfg4_packed"] = rd + wr
        if "sweep_tiled_kernel<float, 3, 0" in r[ik]:
            traffic["sweep_greedy_f32_cfg5"] = rd + wr
        if "sweep_tiled_kernel<double, 3, 0" in r[ik]:
            traffic["sweep_greedy_f64_cfg5"] = rd + wr
tpath = os.path.join(ROOT, "profiles", "traffic.json")
if os.path.exists(tpath):            # keep keys of kernels this capture did not reach (-c limit)
    with open(tpath) as f:
        traffic = dict(json.load(f), **traffic)
with open(tpath, "w") as f:
    json.dump(traffic, f, indent=1)
# launch list: keep kernel name, grid, block, time
src = os.path.join(ROOT, "gpurun_out", tag + "_launches.csv")
if not os.path.exists(src):          # a re-capture of some kernels only: no launch list of its own
    print("traffic", traffic)
    sys.exit(0)
lines = [l for l in open(src) if l.startswith('"')]
rd = list(csv.reader(lines))
h = rd[0]
keep = [h.index(k) for k in ("ID", "Kernel Name", "Block Size", "Grid Size", "Metric Name", "Metric Unit", "Metric Value")]
with open(os.path.join(ROOT, "profiles", tag + "_launches.csv"), "w", newline="") as f:
    w = csv.writer(f)
    for r in rd:
        w.writerow([r[i][:160] for i in keep])
print("traffic", traffic)
