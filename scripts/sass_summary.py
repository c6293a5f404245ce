"""Opcode histogram per kernel of libsrk.so (cuobjdump -sass): the SASS evidence of tcgen05 / TMEM / TMA use.
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "sr_caco_2_b200", "libsrk.so")
out = subprocess.check_output(["cuobjdump", "-sass", so], text=True)
KEYS = ["UTCHMMA", "UTCBAR", "UTMALDG", "UTMAPF", "LDTM", "STTM", "HMMA", "LDGSTS", "SYNCS", "USETMAXREG", "MUFU", "FFMA2", "LDG", "STG", "LDS", "STS", "SHFL", "BAR"]
cur, hist, total = None, collections.OrderedDict(), collections.Counter()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); hist[cur] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and cur:
        op = m.group(1)
        hist[cur]["_all"] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[cur][k] += 1; break
def demangle(n):
    try:
        return subprocess.check_output(["cu++filt", n], text=True).strip()
    except Exception:
        return n
print("# cuobjdump -sass sr_caco_2_b200/libsrk.so (sm_100a): instructions per kernel and the opcodes that prove the Blackwell path")
print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, UTMALDG = TMA load, UTMAPF = TMA L2 prefetch, LDTM/STTM = tcgen05.ld/st, HMMA = mma.sync (legacy path)")
print("kernel".ljust(78) + " ".join(k.rjust(8) for k in ["instr"] + KEYS[:12]))
for fn, h in hist.items():
    name = demangle(fn)
    name = re.sub(r"\((?:[^()]|\([^()]*\))*\)$", "", name)          # drop the argument list, keep the template arguments
    name = name.replace("(int)", "").replace("(bool)", "").replace("void srk::", "").replace("srk::", "")
    print(name[:76].ljust(78) + " ".join(str(h[k]).rjust(8) for k in ["_all"] + KEYS[:12]))
    for k in ["_all"] + KEYS: total[k] += h[k]
print("TOTAL".ljust(78) + " ".join(str(total[k]).rjust(8) for k in ["_all"] + KEYS[:12]))
