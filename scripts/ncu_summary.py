"""Print the key ncu metrics of every launch in a .ncu-rep: python scripts/ncu_summary.py file.ncu-rep"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_warps", "launch__grid_size", "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_wait_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
        "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_membar_per_warp_active.pct",
        "smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct",
        "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct", "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct",
        "smsp__warp_issue_stalled_drain_per_warp_active.pct", "smsp__warp_issue_stalled_imc_miss_per_warp_active.pct",
        "smsp__warp_issue_stalled_tex_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_selected_per_warp_active.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max"]
out = subprocess.check_output(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[0]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:100])
    for k in KEYS:
        if k in d:
            print(f"   {k:75s} {d[k]}")
