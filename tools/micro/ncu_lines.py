"""Not a test: join an `ncu --page source --csv` export with `nvdisasm -g` line info -> per-source-line stall samples.
usage: ncu_lines.py <source.csv> <nvdisasm -c -g output> <mangled kernel name>"""
import csv, re, sys, collections
src_csv, dis, kname = sys.argv[1:4]
# offset -> (file, line)
lines = open(dis).read().split('\n')
start = next(i for i, l in enumerate(lines) if l.startswith('.text.' + kname + ':'))
off2line = {}
cur = None
for l in lines[start + 1:]:
    if l.startswith('//-----') or l.startswith('\t.section'):
        break
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/', l)
    if m:
        off2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) >= len(hdr) and r[ix['# Samples']].isdigit()]
base = min(int(r[ix['Address']], 16) for r in data)
per = collections.defaultdict(lambda: collections.Counter())
for r in data:
    off = int(r[ix['Address']], 16) - base
    key = off2line.get(off, ('?', 0))
    c = per[key]
    c['samples'] += int(r[ix['# Samples']])
    c['inst'] += int(r[ix['Instructions Executed']])
    for s in stalls:
        v = r[ix[s]]
        if v.isdigit():
            c[s] += int(v)
tot = sum(c['samples'] for c in per.values())
toti = sum(c['inst'] for c in per.values())
print('total samples %d, warp instructions %d' % (tot, toti))
for key, c in sorted(per.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if c['samples'] < tot * 0.002 and c['inst'] < toti * 0.002:
        continue
    top = sorted(((s, c[s]) for s in stalls if c[s]), key=lambda x: -x[1])[:3]
    print('%-14s:%4d  samples %6.2f%%  inst %6.2f%%  %s' % (key[0], key[1], 100.0 * c['samples'] / tot, 100.0 * c['inst'] / toti,
          ' '.join('%s=%.1f%%' % (s[6:], 100.0 * v / tot) for s, v in top)))
