"""Condense an .ncu-rep into the metric,unit,value CSV summaries kept in this directory.
    python profiles/summarise_ncu.py report.ncu-rep "kernel-name substring" "header comment" > summary.csv"""
import csv, io, subprocess, sys
KEEP = ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'gpu__time_duration.sum', 'l1tex__t_sector_hit_rate.pct',
        'launch__block_size', 'launch__grid_size', 'launch__registers_per_thread', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_local_op_st.sum')
rep, want, header = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
names, units = rows[0], rows[1]
pick = [r for r in rows[2:] if want in r[names.index('Kernel Name')]]
if not pick:
    sys.exit(f'no kernel matching {want!r}')
r = pick[-1]                                   # the last matching launch (warm)
print('# ' + header)
print('metric,unit,value')
print(f'Kernel Name,,"{r[names.index("Kernel Name")]}"')
for k in KEEP:
    if k in names:
        i = names.index(k)
        print(f'{k},{units[i]},{r[i]}')
