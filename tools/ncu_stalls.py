#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall
samples with the dominant stall reason.  usage: ncu_stalls.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
tot = 0
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] in ('Address', 'Kernel Name'):
        if r and r[0] == 'Kernel Name':
            break          # first section only
        continue
    n = int(r[col['# Samples']] or 0)
    tot += n
    st = {h: int(r[col[h]] or 0) for h in stall_cols}
    data.append((n, r[col['Address']][-5:], r[col['Source']].strip(), st, int(r[col['Instructions Executed']] or 0)))
print('total samples', tot, 'instructions', len(data))
agg = {}
for n, a, s, st, ie in data:
    for k, v in st.items():
        agg[k] = agg.get(k, 0) + v
print('by reason:', sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
order = sorted(range(len(data)), key=lambda i: -data[i][0])[:top]
for i in sorted(order):
    n, a, s, st, ie = data[i]
    main = sorted(((v, k) for k, v in st.items() if v), reverse=True)[:2]
    print('%5d %5.1f%% #%d %s  exec=%d  %s  %s' % (i, 100.0 * n / tot, n, a, ie, s[:70], main))
