"""Per-source-line executed-instruction / stall-sample totals of one kernel of an ncu report:
joins `ncu --page source --csv` (SASS order) with `nvdisasm -g` line info of the cubin in libsrk.so.
    python scripts/ncu_lines.py gpurun_out/prof_block_r2.ncu-rep attn_block attn_block_tc5 [top]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kre, cubin_stem = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# several kernels may match: take the first block
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
which = int(os.environ.get("KIDX", "0"))
blk = blocks[which]
h = blk["rows"][0]; ie, si = h.index("Instructions Executed"), h.index("# Samples")
sass = [(r[1].strip(), int(r[ie]), int(r[si])) for r in blk["rows"][1:]]
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "sr_caco_2_b200", "libsrk.so")], cwd=td, capture_output=True)
    cub = [f for f in os.listdir(td) if f.startswith(cubin_stem + ".") and "sm_100a" in f][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=td, capture_output=True, text=True).stdout
# split per function; pick the one whose instruction count matches
funcs, cur, line = {}, None, None
for l in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m: cur = m.group(1); funcs[cur] = []; line = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", l): funcs[cur].append(line)
cands = [f for f, v in funcs.items() if len(v) == len(sass)]
if not cands:
    sys.exit(f"no function with {len(sass)} instructions: {[(f, len(v)) for f, v in funcs.items()]}")
lines = funcs[cands[0]]
print("kernel", blk["name"][:90], "| function", cands[0][:60], "|", len(sass), "SASS instructions")
agg, sagg = collections.Counter(), collections.Counter()
for (src, n, s), ln in zip(sass, lines):
    agg[ln] += n; sagg[ln] += s
tot, stot = sum(agg.values()), sum(sagg.values())
print("total warp instructions", tot, "samples", stot)
srcs = {}
def text(ln):
    if ln is None: return ""
    f = ln[0]
    if f not in srcs:
        p = os.path.join(ROOT, "sr_caco_2_b200", "csrc", f)
        srcs[f] = open(p).read().splitlines() if os.path.exists(p) else []
    return srcs[f][ln[1] - 1].strip()[:100] if ln[1] - 1 < len(srcs[f]) else ""
order = sagg.most_common(top) if os.environ.get("BY_SAMPLES") else agg.most_common(top)
for ln, _ in order:
    n = agg[ln]
    print(f"{n:>10} {100 * n / tot:5.1f}%  smp {100 * sagg[ln] / max(stot, 1):5.1f}%  {ln[0] if ln else '?'}:{ln[1] if ln else 0:<5} {text(ln)}")
# per file
pf, ps = collections.Counter(), collections.Counter()
for ln, n in agg.items(): pf[ln[0] if ln else "?"] += n
for ln, n in sagg.items(): ps[ln[0] if ln else "?"] += n
print({k: f"{100 * v / tot:.1f}% inst / {100 * ps[k] / max(stot, 1):.1f}% samples" for k, v in pf.items()})
if os.environ.get("RANGES"):
    # RANGES="file:lo-hi:name,..." region totals
    for spec in os.environ["RANGES"].split(","):
        f, rg, name = spec.split(":"); lo, hi = map(int, rg.split("-"))
        n = sum(v for k, v in agg.items() if k and k[0] == f and lo <= k[1] <= hi)
        s = sum(v for k, v in sagg.items() if k and k[0] == f and lo <= k[1] <= hi)
        print(f"  {name:28s} {n:>10} {100 * n / tot:5.1f}% inst  {100 * s / max(stot, 1):5.1f}% samples")
