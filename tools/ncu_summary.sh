#!/bin/bash
# usage: tools/ncu_summary.sh report.ncu-rep "header line" > profiles/xxx.txt  -- the raw-page metrics the design doc quotes
echo "$2"
ncu -i "$1" --page raw --csv 2>/dev/null | python3 -c "
import csv, sys
rows = list(csv.reader(sys.stdin))
names, units, vals = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct', 'smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct']
for w in want:
    if w in names:
        i = names.index(w)
        print('%-95s %s %s' % (w, vals[i], units[i]))
"
