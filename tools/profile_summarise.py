"""Turns the ncu outputs of tools/profile_round.sh (gpurun_out/) into the tracked summaries under profiles/.
usage: python tools/profile_summarise.py r1i"""
import csv, json, sys, collections, os
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# 1. launch list of the bench command: copy + per-kernel summary (share of the step)
rows = list(csv.reader(open(os.path.join(G, f"launches_{tag}.csv"))))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]; ki, vi = h.index("Kernel Name"), h.index("Metric Value")
with open(os.path.join(P, f"{tag}_launch_list.csv"), "w", newline="") as f:
    csv.writer(f).writerows(rows[hdr:])
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > vi:
        name = r[ki].split("(")[0]
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(P, f"{tag}_launch_list_summary.csv"), "w", newline="") as f:
    w = csv.writer(f); w.writerow(["kernel", "launches", "total_us", "mean_us", "share_of_all_gpu_time"])
    for k, (n, t) in agg.items():
        w.writerow([k, n, f"{t/1e3:.1f}", f"{t/n/1e3:.2f}", f"{t/tot:.4f}"])

# 2. --set full capture: the metrics DESIGN.md quotes, one column per launch
raw = list(csv.reader(open(os.path.join(G, f"prof_{tag}_raw.csv"))))
hh, units = raw[0], raw[1]
want = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]
want += [x for x in hh if "average_warps_issue_stalled" in x and "per_issue_active" in x and "not_issued" not in x]
with open(os.path.join(P, f"{tag}_ncu_full_sweep_kernels.csv"), "w", newline="") as f:
    w = csv.writer(f); w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(raw) - 2)])
    for m in want:
        if m in hh:
            i = hh.index(m); w.writerow([m, units[i]] + [r[i] for r in raw[2:]])

# 3. DRAM traffic per launch for bench.py's roofline.traffic
def col(m): return [float(r[hh.index(m)].replace(",", "")) for r in raw[2:]]
un = units[hh.index("dram__bytes_read.sum")]
scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[un]
per = [(a + b) * scale for a, b in zip(col("dram__bytes_read.sum"), col("dram__bytes_write.sum"))]
names = [r[hh.index("Kernel Name")] for r in raw[2:]]
json.dump({"dram_bytes_per_launch_mean": sum(per) / len(per), "per_launch": per, "kernels": names,
           "source": f"profiles/{tag}_ncu_full_sweep_kernels.csv (ncu --set full, {len(per)} consecutive sweep launches, 2048x1024)",
           "note": "bytes counted inside each kernel's own window; part of a kernel's output is written back from L2 after the kernel ends"},
          open(os.path.join(P, "ncu_traffic.json"), "w"), indent=1)
print(open(os.path.join(P, f"{tag}_launch_list_summary.csv")).read())
