#!/usr/bin/env python
"""Pick the metrics quoted in profiles/ out of `ncu -i x.ncu-rep --page raw --csv` files.  Usage: python tools/ncu_raw_summary.py a.csv [b.csv ...]"""
import csv, os, sys
WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__occupancy_limit_registers', 'sm__maximum_warps_per_active_cycle_pct',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'sm__cycles_elapsed.max',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__cycles_active.avg.pct_of_peak_sustained_elapsed']
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print("## " + os.path.basename(path).replace("ncu_raw_", "").replace(".csv", ""))
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print('%-72s %-16s %s' % (w, units[i], vals[i]))
    st = {}
    for i, h in enumerate(hdr):
        if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            try:
                st[h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]] = float(vals[i])
            except ValueError:
                pass
    print('top stall reasons (warps stalled per issue-active cycle): ' + ', '.join('%s=%.2f' % kv for kv in sorted(st.items(), key=lambda kv: -kv[1])[:6]))
    print()
