"""Print the per-launch table from an ncu --csv launch list (see profiles/)."""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi, mi, ii = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name'), hdr.index('ID')
d = collections.OrderedDict()
for r in rows[1:]:
    d.setdefault((r[ii], r[ki][:64]), {})[r[mi]] = r[vi]
f = lambda m, k: float(m.get(k, '0').replace(',', ''))
tot = 0
for (i, k), m in d.items():
    t = f(m, 'gpu__time_duration.sum') / 1e6
    rd, wr = f(m, 'dram__bytes_read.sum'), f(m, 'dram__bytes_write.sum')
    tot += t
    print('%3s %-64s %8.3f ms rd %6.2f GB wr %6.2f GB %6.0f GB/s sm%% %5s regs %3s occ%% %5s' % (
        i, k, t, rd / 1e9, wr / 1e9, (rd + wr) / t / 1e6 if t else 0, m.get('sm__throughput.avg.pct_of_peak_sustained_elapsed', ''),
        m.get('launch__registers_per_thread', ''), m.get('sm__warps_active.avg.pct_of_peak_sustained_active', '')))
print('total %.3f ms' % tot)
