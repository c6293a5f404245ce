"""Stall samples of an `ncu --page source --csv` dump aggregated per warp role.  Roles are delimited by marker
substrings given on the command line as name=substring (first SASS row containing it starts the region)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
def I(r, k):
    try: return int(r[idx[k]])
    except Exception: return 0
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
bounds = []
for spec in sys.argv[2:]:
    name, a, b = spec.split(":")
    bounds.append((name, int(a), int(b)))
tot = sum(I(r, '# Samples') for r in data)
print("total samples", tot, "rows", len(data))
for name, a, b in bounds:
    sub = data[a:b]
    n = sum(I(r, '# Samples') for r in sub)
    ins = sum(I(r, 'Instructions Executed') for r in sub)
    st = sorted(((sum(I(r, h) for r in sub), h) for h in stall), reverse=True)[:5]
    print(f"{name:10s} rows {a}-{b} samples {n} ({100*n/max(tot,1):.1f}%) warp-inst {ins}  " + " ".join(f"{h[6:]}={v}" for v, h in st))
