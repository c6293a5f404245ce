"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel."""
import csv, collections, sys, re
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    name = r[idx['Kernel Name']]
    m = re.match(r'(?:void )?(?:srk::)?([A-Za-z0-9_]+)(<[^(]*>)?', name)
    short = (m.group(1) + (m.group(2) or '')) if m else name[:60]
    short = short.replace('(int)', '')
    v = float(r[idx['Metric Value']]); unit = r[idx['Metric Unit']]
    v = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)
    agg[short][0] += 1; agg[short][1] += v; tot += v
print(f"launches {sum(a[0] for a in agg.values())}  total {tot:.1f} us (cold-cache, serialised: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d}  avg {t/n:8.1f} us  {k}")
