"""Here (no GPU): per-source-line stall profile from an ncu SASS source page.
    ncu -i X.ncu-rep --page source --csv > X_sass.csv          (on the box)
    python tools/ncu_sass_lines.py X_sass.csv <dis file from `nvdisasm -g -c mlp_tc.sm_100a.cubin`> <mangled kernel substring> [top]
Maps every SASS instruction (by its offset from the kernel's first instruction) to the CUDA source line nvdisasm reports
and sums the warp-stall samples and their reasons per line."""
import csv
import re
import sys
from collections import defaultdict

sass_csv, dis, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
line_of = {}
cur = None
inside = False
for ln in open(dis, errors="replace"):
    if ln.startswith(".text.") and ln.rstrip().endswith(":"):
        inside = kern in ln
        cur = None
        continue
    if not inside:
        continue
    m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", ln)
    if m:
        line_of[int(m.group(1), 16)] = cur
csv.field_size_limit(10 ** 9)
rows = list(csv.reader(open(sass_csv, errors="replace")))
hi = next(i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address")
hdr = rows[hi]
col = {n: i for i, n in enumerate(hdr)}
samp = col["Warp Stall Sampling (All Samples)"]
inst = col.get("Instructions Executed")
stall_cols = [(n, i) for n, i in col.items() if n.startswith("stall_") and "Not Issued" not in n]
base = None
agg = defaultdict(lambda: [0, 0, defaultdict(int), set()])
total = 0
for r in rows[hi + 1:]:
    if len(r) <= samp or not r[0].startswith("0x"):
        continue
    addr = int(r[0], 16)
    if base is None:
        base = addr
    key = line_of.get(addr - base)
    s = int(float(r[samp] or 0))
    a = agg[key]
    a[0] += s
    total += s
    if inst is not None and r[inst]:
        a[1] += int(float(r[inst]))
    op = r[col["Source"]].split()
    if op:
        a[3].add(op[0] if not op[0].startswith("@") else (op[1] if len(op) > 1 else op[0]))
    for n, i in stall_cols:
        if i < len(r) and r[i]:
            a[2][n] += int(float(r[i]))
src = {}
try:
    for i, l in enumerate(open("/root/repo/scenerf_b200/csrc/mlp_tc.cu"), 1):
        src[i] = l.rstrip()
except OSError:
    pass
print("total samples %d, mapped instructions %d" % (total, len(line_of)))
for key, (s, ins, st, ops) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    reasons = ", ".join("%s %d%%" % (n[6:], 100 * v // max(1, s)) for n, v in sorted(st.items(), key=lambda kv: -kv[1])[:2] if v)
    text = src.get(key[1], "")[:90].strip() if key and key[0] == "mlp_tc.cu" else str(key)
    print("%5.2f%% L%-5s %-34s %-28s | %s" % (100.0 * s / max(1, total), key[1] if key else "?", reasons, ",".join(sorted(ops))[:28], text))
