"""Static evidence for the built library (no GPU needed): per kernel, the resource usage `cuobjdump -res-usage` reports
and how often the Blackwell-specific SASS mnemonics occur (UTCHMMA/UTCQMMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st,
UTMALDG/UTMASTG/UTMAREDG = TMA load/store/reduce, UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = mma.sync, IDP = dp4a,
MUFU.EX2 = exp2 on the XU).  Spill sizes come from a `-Xptxas -v` recompile of each source.
    python tools/sass_summary.py > profiles/r2_sass_resources.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from clip_assisted_data_labeling_b200 import _build  # noqa: E402

MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UTCBAR", "SYNCS", "HMMA", "IDP", "MUFU.EX2",
             "STL", "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    lib = _build.LIB_PATH
    res = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.splitlines():
        m = re.search(r"Function (\S+):", line)
        if m:
            cur = m.group(1)
        elif cur and "REG:" in line:
            usage[cur] = dict(kv.split(":") for kv in line.split() if ":" in kv and not kv.startswith("CONSTANT"))
            cur = None
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur:
            m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m:
                op = m.group(1)
                counts[cur]["_total"] += 1
                for mn in MNEMONICS:
                    if op.startswith(mn):
                        counts[cur][mn] += 1
    spills = {}
    for src in _build.SOURCES:
        r = subprocess.run([_build._nvcc(), *_build.NVCC_FLAGS, "-Xptxas", "-v", "-c", os.path.join(_build.CSRC, src), "-o", os.devnull],
                           capture_output=True, text=True).stderr
        fn = None
        for line in r.splitlines():
            m = re.search(r"Function properties for (\S+)", line)
            if m:
                fn = m.group(1)
            m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
            if m and fn:
                spills[fn] = (int(m.group(2)), int(m.group(3)))
    names = demangle(sorted(usage))
    print(f"# {os.path.relpath(lib, ROOT)}: {len(usage)} kernels, sm_100a; columns: registers / static shared bytes / stack bytes / "
          "spill stores+loads (bytes) / SASS instructions / Blackwell mnemonics")
    for fn in sorted(usage, key=lambda f: names[f]):
        u, c = usage[fn], counts.get(fn, {})
        short = re.sub(r"\(.*", "", names[fn]).replace("b2c::", "")
        mn = " ".join(f"{k}={c[k]}" for k in MNEMONICS if c.get(k))
        sp = spills.get(fn, (0, 0))
        print(f"{short:58s} REG {u.get('REG', '?'):>3s}  SMEM {u.get('SHARED', '?'):>5s}  STACK {u.get('STACK', '?'):>3s}  "
              f"spill {sp[0]}+{sp[1]}  SASS {c.get('_total', 0):>6d}  {mn}")


if __name__ == "__main__":
    main()
