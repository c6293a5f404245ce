"""Summarise an ncu `--metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv`
launch list by kernel: time share, launches, DRAM traffic."""
import csv, collections, sys, re, json
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
per = collections.defaultdict(dict)          # launch id -> {metric: value}
name_of = {}
def to_us(v, unit):
    return v / 1e3 if unit.startswith('n') else (v * 1e3 if unit.startswith('m') else (v * 1e6 if unit in ('s', 'second') else v))
def to_bytes(v, unit):
    u = unit.lower()
    return v * (1e9 if u.startswith('g') else 1e6 if u.startswith('m') else 1e3 if u.startswith('k') else 1.0)
for r in rows[hi + 1:]:
    if len(r) < len(hdr): continue
    i = r[idx['ID']]; name = r[idx['Kernel Name']]
    m = re.match(r'(?:void )?(?:srk::)?([A-Za-z0-9_]+)(<[^(]*>)?', name)
    name_of[i] = ((m.group(1) + (m.group(2) or '')) if m else name[:60]).replace('(int)', '')
    v = float(r[idx['Metric Value']].replace(',', '')); unit = r[idx['Metric Unit']]; met = r[idx['Metric Name']]
    per[i][met] = to_us(v, unit) if 'time' in met else to_bytes(v, unit)
# keep ONE step: from the last launch of the network's first kernel (conv_in_ln / conv_in) to the end
ids = sorted(per, key=lambda x: int(x))
firsts = [i for i in ids if name_of[i].startswith("conv_in")]
if firsts:
    ids = [i for i in ids if int(i) >= int(firsts[-1])]
    per = {i: per[i] for i in ids}
agg = collections.defaultdict(lambda: [0, 0.0, 0.0]); tot = 0.0; totb = 0.0
for i, d in per.items():
    t = d.get('gpu__time_duration.sum', 0.0); b = d.get('dram__bytes_read.sum', 0.0) + d.get('dram__bytes_write.sum', 0.0)
    a = agg[name_of[i]]; a[0] += 1; a[1] += t; a[2] += b; tot += t; totb += b
print(f"launches {len(per)}  total {tot:.1f} us  DRAM traffic {totb/1e9:.3f} GB (under ncu: cold cache, serialised -> compare SHARES)")
for k, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:10.1f} us {100*t/tot:5.1f}%  n={n:4d}  avg {t/n:8.1f} us  dram {b/1e6:10.1f} MB ({b/1e6/n:8.2f} MB/launch)  {k}")
if len(sys.argv) > 2:
    fam = {"gemm": [0, 0.0, 0.0]}
    for k, (n, t, b) in agg.items():
        if k.startswith("gemm_") or k.startswith("mlp_"):
            fam["gemm"][0] += n; fam["gemm"][1] += t; fam["gemm"][2] += b
    famof = lambda k: "attn_block" if k.startswith("attn_block") else "mlp" if k.startswith("mlp_") else "gemm_res_ln" if k.startswith("gemm_tc5_kernel<192, 2") else "gemm" if k.startswith("gemm_") else "other"
    byf = collections.defaultdict(lambda: {"launches": 0, "us": 0.0, "dram_bytes": 0.0})
    for k, (n, t, b) in agg.items():
        f = byf[famof(k)]; f["launches"] += n; f["us"] += t; f["dram_bytes"] += b
    json.dump({"step": {"launches": len(per), "us": tot, "dram_bytes": totb}, "by_family": byf,
               "gemm_family": {"launches": fam["gemm"][0], "us": fam["gemm"][1], "dram_bytes": fam["gemm"][2]},
               "by_kernel": {k: {"launches": n, "us": t, "dram_bytes": b} for k, (n, t, b) in agg.items()},
               "total_us": tot, "total_dram_bytes": totb}, open(sys.argv[2], "w"), indent=1)
