"""Key metrics of one `ncu -i X.ncu-rep --page raw --csv` export as a small metric,unit,value csv (dev tool):
python tools/ncu_key_metrics.py raw.csv > profiles/rN_kernel_ncu_raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, units, v = rows[0], rows[1], rows[2]
keep = ["Kernel Name", "Block Size", "Grid Size", "dram__bytes_read.sum", "dram__bytes_read.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum", "dram__bytes_write.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "gpu__time_duration.sum",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.per_cycle_active", "smsp__inst_executed.sum",
        "sm__inst_executed.avg.per_cycle_active", "sm__inst_issued.avg.pct_of_peak_sustained_active",
        "sm__instruction_throughput.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
w = csv.writer(sys.stdout, quoting=csv.QUOTE_ALL)
sys.stdout.write("metric,unit,value\n")
for k in keep:
    if k in h:
        i = h.index(k)
        w.writerow([k, units[i], v[i]])
