"""Summarises `ncu --set full --csv --page raw` logs into the few metrics profiles/README.md quotes."""
import csv, json, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.per_cycle_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'launch__occupancy_limit_registers', 'sm__cycles_elapsed.avg.per_second', 'local_load', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum']


def summarise(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 20]
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        s = {'kernel': d['Kernel Name']}
        for w in WANT:
            if w in d and d[w] not in ('', 'n/a'):
                s[w] = d[w] + ' ' + u[w]
        stalls = []
        for h in hdr:
            if 'issue_stalled' in h and h.endswith('per_issue_active.ratio') and d[h] not in ('', 'n/a'):
                try:
                    stalls.append((float(d[h].replace(',', '')), h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
                except ValueError:
                    pass
        s['stalls_per_issue'] = {k: round(v, 3) for v, k in sorted(stalls, reverse=True)[:6]}
        out.append(s)
    return out


if __name__ == '__main__':
    for p in sys.argv[1:]:
        for s in summarise(p):
            print(json.dumps(s, indent=1))
