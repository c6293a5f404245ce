"""Summarise an `ncu --page source --csv` dump: hottest SASS by executed count and by stall samples."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; idx = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[hi + 1:] if len(r) == len(hdr) and r[0].startswith("0x")]
def I(r, k):
    try: return int(r[idx[k]])
    except Exception: return 0
print("kernel:", rows[0][1][:100] if rows[0] else "")
print("sass rows", len(data), "samples", sum(I(r, '# Samples') for r in data), "warp-inst", sum(I(r, 'Instructions Executed') for r in data))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 20
print("--- top by instructions executed")
for r in sorted(data, key=lambda r: -I(r, 'Instructions Executed'))[:n]:
    print(str(I(r, 'Instructions Executed')).rjust(10), str(I(r, '# Samples')).rjust(7), r[idx['Source']].strip()[:100])
print("--- top by stall samples")
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
for r in sorted(data, key=lambda r: -I(r, '# Samples'))[:n]:
    st = sorted([(I(r, h), h) for h in stall], reverse=True)[:2]
    print(str(I(r, '# Samples')).rjust(7), str(I(r, 'Instructions Executed')).rjust(10), r[idx['Source']].strip()[:64].ljust(64), st)
print("--- stall totals")
tot = {h: sum(I(r, h) for r in data) for h in stall}
for h, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]: print(f"  {h:28s} {v}")
print("--- shared conflicts")
for r in sorted(data, key=lambda r: -I(r, 'L1 Wavefronts Shared Excessive'))[:6]:
    print(str(I(r, 'L1 Wavefronts Shared Excessive')).rjust(9), str(I(r, 'L1 Wavefronts Shared')).rjust(9), r[idx['Source']].strip()[:90])
