#!/usr/bin/env python
"""Hot SASS lines of an ncu source-page CSV (ncu -i X.ncu-rep --page source --csv): samples, executed count, top stall reasons."""
import csv, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]; idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = []
for n, r in enumerate(rows[2:]):
    if len(r) < len(hdr): continue
    try: s = int(r[idx['# Samples']] or 0)
    except ValueError: continue
    data.append((n, r, s))
tot = sum(d[2] for d in data)
print('total samples', tot, 'lines', len(data))
if len(sys.argv) > 3:   # window: print lines a..b in order
    a, b = map(int, sys.argv[3].split(':'))
    sel = [d for d in data if a <= d[0] <= b]
else:
    sel = sorted(data, key=lambda d: -d[2])[:top]
for n, r, s in sel:
    st = sorted(((int(r[idx[h]] or 0), h[6:]) for h in stalls), reverse=True)[:3]
    print(f"{n:6d} {s:7d} {100.0*s/tot:5.2f}% ex={r[idx['Instructions Executed']]:>10s} {r[idx['Source']][:90]:90s} " + ' '.join(f'{h}:{v}' for v, h in st if v))
