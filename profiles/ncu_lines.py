#!/usr/bin/env python
"""Aggregates an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line.
usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass --kernel-name regex:K > f.csv; ncu_lines.py f.csv [N]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hi = [i for i, r in enumerate(rows) if '# Samples' in r][0]
hdr = rows[hi]
si = hdr.index('# Samples')
ie = hdr.index('Instructions Executed')
agg = defaultdict(lambda: [0, 0, ''])
for r in rows[hi + 1:]:
    if len(r) <= si or not r[0].isdigit():
        continue
    try:
        s = int(r[si] or 0); n = int(r[ie] or 0)
    except ValueError:
        continue
    a = agg[int(r[0])]
    a[0] += s; a[1] += n; a[2] = r[1]
tot = sum(a[0] for a in agg.values()) or 1
toti = sum(a[1] for a in agg.values()) or 1
print('samples', tot, 'warp-instructions', toti)
for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%5d %6.1f%% smp %6.1f%% ins  %s' % (ln, 100 * a[0] / tot, 100 * a[1] / toti, a[2].strip()[:110]))
