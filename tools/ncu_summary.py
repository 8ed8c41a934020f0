"""Text summary of an .ncu-rep (raw page): the metrics the rooflines in DESIGN.md / bench.py are argued from.
    python tools/ncu_summary.py gpurun_out/x.ncu-rep "title / command" > profiles/x.txt   (development tool)"""
import csv, subprocess, sys
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tensor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active", "smsp__average_warps_issue_stalled_wait_per_issue_active",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active", "smsp__average_warps_issue_stalled_sleeping_per_issue_active"]
print(title)
print("source: ncu --set full --clock-control none --import-source on (one launch; the same command had exited 0 without ncu first)\n")
for r in rows[2:]:
    for w in want:
        for h, u, v in zip(hdr, units, r):
            if h == w or (w.endswith("pipe_tensor") and h.startswith(w)):
                print("%-90s %-12s %s" % (h, u, v))
    print()
