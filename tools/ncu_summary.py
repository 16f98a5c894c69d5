"""Key metrics + warp-stall mix of every kernel in an `ncu --set full` report, as the text table committed under profiles/.

    ncu -i report.ncu-rep --page raw --csv > raw.csv ;  python tools/ncu_summary.py raw.csv "header line" > profiles/rNN_all_kernels_ncu_full.txt
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.per_cycle_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__inst_executed.sum', 'sm__cycles_active.avg']
want += [h for h in hdr if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h]
ki = hdr.index('Kernel Name')
names = [r[ki].split('(')[0].replace('void ', '')[:58] for r in rows[2:]]
out = [sys.argv[2] if len(sys.argv) > 2 else '# ncu --set full summary', '# (cold-cache, serialised replays: compare shares and percentages, not absolute times)']
out.append('%-92s %-14s ' % ('metric', 'unit') + ' | '.join('%-22s' % n[:22] for n in names))
for w in want:
    if w not in hdr:
        continue
    i = hdr.index(w)
    vals = []
    for r in rows[2:]:
        try:
            vals.append('%-22.6g' % float(r[i].replace(',', '')))
        except ValueError:
            vals.append('%-22s' % r[i][:22])
    out.append('%-92s %-14s ' % (w, units[i][:14]) + ' | '.join(vals))
out.append('')
out.append('kernels (columns): ' + ' ; '.join('%d=%s' % (k + 1, n) for k, n in enumerate(names)))
print('\n'.join(out))
