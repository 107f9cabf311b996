# raw ncu CSV (ncu -i X.ncu-rep --page raw --csv) → a small JSON per kernel with the metrics DESIGN.md / bench.py quote
import csv, json, sys
src, out = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(src)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__compute_memory_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic', 'launch__shared_mem_per_block_static',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.per_cycle_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
res = []
for r in rows[2:]:
    d = {'kernel': r[idx['Kernel Name']]}
    for k in keep:
        if k in idx and r[idx[k]] not in ('', '-nan', 'nan'):
            try:
                d[k] = float(r[idx[k]].replace(',', ''))
            except ValueError:
                d[k] = r[idx[k]]
            d[k + ' [unit]'] = units[idx[k]]
    res.append(d)
json.dump({'source': src, 'how': 'ncu --set full --clock-control none --import-source on, one launch per kernel, read with ncu -i … --page raw --csv', 'kernels': res}, open(out, 'w'), indent=1)
print(out, len(res), 'kernels')
