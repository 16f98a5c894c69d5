"""Summarise an `ncu --page source --csv` dump: stall mix, samples by opcode, hottest instructions."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0] != 'Address']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
tot = collections.Counter(); byop = collections.Counter(); execs = collections.Counter(); total = 0
for r in data:
    src = r[idx['Source']]; n = int(r[idx['# Samples']] or 0)
    toks = src.split()
    op = (toks[1] if toks[0].startswith('@') else toks[0]).split('.')[0]
    byop[op] += n; execs[op] += int(r[idx['Instructions Executed']] or 0); total += n
    for h in stalls: tot[h] += int(r[idx[h]] or 0)
print('total samples', total, ' instructions executed', sum(execs.values()))
print({k[6:]: '%.1f%%' % (100 * v / total) for k, v in tot.most_common(9)})
for op, n in byop.most_common(14): print('%-10s samples %7d (%4.1f%%) executed %d' % (op, n, 100 * n / total, execs[op]))
top = sorted(data, key=lambda r: -int(r[idx['# Samples']] or 0))[:int(sys.argv[2]) if len(sys.argv) > 2 else 16]
for r in top:
    st = {h[6:]: int(r[idx[h]] or 0) for h in stalls}
    st = {k: v for k, v in st.items() if v > 0.2 * int(r[idx['# Samples']])}
    print(r[idx['Address']][-5:], '%-58s' % r[idx['Source']][:58], r[idx['# Samples']], st)
