"""Key metrics of one `ncu --set full` capture: python tools/ncu_keymetrics.py raw.csv  (raw.csv = `ncu -i X.ncu-rep --page raw --csv`)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h, u, v = rows[0], rows[1], rows[2]
want = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread
launch__shared_mem_per_block_dynamic launch__occupancy_limit_registers launch__occupancy_limit_shared_mem
sm__warps_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__thread_inst_executed.sum
smsp__issue_active.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
lts__t_sector_hit_rate.pct l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio""".split()
print("metric,unit,value")
for w in want:
    if w in h:
        i = h.index(w)
        print(f"{w},{u[i]},{v[i]}")
