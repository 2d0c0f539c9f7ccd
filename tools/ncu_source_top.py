"""Top stall-sample SASS instructions of one kernel from `ncu -i X.ncu-rep --page source --csv` output.
    ncu -i rep --page source --csv --launch-skip S --launch-count 1 > /tmp/src.csv; python tools/ncu_source_top.py /tmp/src.csv [N]
"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
ix = {k: i for i, k in enumerate(hdr)}
data = [r for r in rows[h + 1:] if len(r) == len(hdr) and r[ix["# Samples"]].isdigit()]
tot = sum(int(r[ix["# Samples"]]) for r in data)
print(rows[0][1][:120] if len(rows[0]) > 1 else "", "total samples", tot)
stalls = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
for i, r in enumerate(data):
    r.append(i)
for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:n]:
    s = sorted(((int(r[ix[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print(f"{int(r[ix['# Samples']]):7d} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  #{r[-1]:4d} {r[ix['Source']][:90]:90s} {s}")
