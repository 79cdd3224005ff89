"""Print selected raw metrics of an ncu report exported with `ncu -i x.ncu-rep --page raw --csv`:  python tools/ncu_pick.py raw.csv [row]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
r = rows[int(sys.argv[2]) if len(sys.argv) > 2 else 2]
d = dict(zip(hdr, r))
KEYS = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'smsp__average_warps_issue_stalled',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fp64', 'sm__inst_executed_pipe_lsu',
        'sm__inst_executed_pipe_alu', 'sm__inst_executed_pipe_fma', 'sm__inst_executed_pipe_xu', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local', 'smsp__inst_executed_op_local', 'smsp__inst_executed_op_shared', 'sm__throughput', 'lts__t_sectors_srcunit_tex_op',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared', 'smsp__pcsamp']
for h in hdr:
    if any(h.startswith(k) for k in KEYS) and 'per_second' not in h and 'pct_of_peak_sustained_elapsed' not in h:
        print(h, '=', d[h])
