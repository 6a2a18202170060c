"""Prints the metrics that matter for the sweep kernels from an `ncu --page raw --csv` export."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__cycles_active.avg']
stall = [x for x in h if 'average_warps_issue_stalled' in x and 'per_issue_active' in x and 'not_issued' not in x]
for w in want + stall:
    if w in h:
        i = h.index(w)
        name = w.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', '')
        vals = [r[i][:14] for r in rows[2:]]
        if name.startswith('stall_') and all(float(v) < 0.05 for v in vals): continue
        print(f"{name[:62]:62s} {units[i]:6s}", vals)
