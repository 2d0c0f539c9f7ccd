"""Stall samples per CUDA source line: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv > f.csv ; python tools/ncu_lines_top.py f.csv [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = ""
agg = {}
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix_s = hdr.index("# Samples")
        stalls = [(i, k[6:]) for i, k in enumerate(hdr) if k.startswith("stall_") and "Not Issued" not in k]
        continue
    if hdr is None or len(r) != len(hdr) or not r[0].isdigit() or not r[ix_s].isdigit():
        continue
    if r[2] != "-":  # SASS rows carry an address; per-line rows have "-"
        continue
    key = (cur_file, int(r[0]))
    a = agg.setdefault(key, [0, r[1].strip()[:100], {}])
    a[0] += int(r[ix_s])
    for i, k in stalls:
        v = int(r[i] or 0)
        if v:
            a[2][k] = a[2].get(k, 0) + v
tot = sum(a[0] for a in agg.values())
print("total samples", tot)
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    top = sorted(a[2].items(), key=lambda kv: -kv[1])[:2]
    print(f"{a[0]:7d} {100 * a[0] / max(tot, 1):5.1f}%  {f}:{ln:<5d} {a[1]:100s} {top}")
