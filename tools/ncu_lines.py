#!/usr/bin/env python
"""Per-source-line instruction counts and stall samples of one kernel from an .ncu-rep captured with
--import-source on: joins the SASS page of the report with nvdisasm's line info of the object file.
usage: tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel-name-substring> [top]"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict
from pathlib import Path

rep, obj, kname = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
# the report may hold several kernels: each section starts with a "Kernel Name" row; take the first whose
# name contains the (demangled) kernel name given as the 5th argument, default = the 3rd argument up to "I"
want = sys.argv[5] if len(sys.argv) > 5 else kname.split("IL")[0]
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
sec = next((i for i in starts if want in rows[i][1]), starts[0] if starts else 0)
end = next((i for i in starts if i > sec), len(rows))
rows = rows[sec:end]
hdr_i = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hdr_i]
ie, ns, src = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
sass = []
for r in rows[hdr_i + 1:]:
    try:
        sass.append((r[src].strip(), int(r[ie]), int(r[ns])))
    except (ValueError, IndexError):
        pass
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", str(Path(obj).resolve())], cwd=td, capture_output=True)
    cub = next(Path(td).glob("*.cubin"))
    dis = subprocess.run(["nvdisasm", "-g", "-c", str(cub)], capture_output=True, text=True).stdout.splitlines()
start = next(i for i, l in enumerate(dis) if l.startswith("_Z") and kname in l and l.rstrip().endswith(":"))
line = 0
lines = []
for l in dis[start + 1:]:
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        line = int(m.group(1))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(line)
    if l.startswith("_Z") or ".section" in l:
        break
n = min(len(lines), len(sass))
agg = defaultdict(lambda: [0, 0])
for k in range(n):
    agg[lines[k]][0] += sass[k][1]
    agg[lines[k]][1] += sass[k][2]
ti, ts = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
srcfile = None
for l in dis[start:start + 400]:
    m = re.search(r'//## File "(.*?)", line', l)
    if m:
        srcfile = m.group(1)
        break
text = Path(srcfile).read_text().splitlines() if srcfile and Path(srcfile).exists() else []
print(f"sass instructions {len(sass)} / disasm {len(lines)}; warp instructions executed {ti:,}; samples {ts:,}")
for ln, (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    code = text[ln - 1].strip()[:100] if 0 < ln <= len(text) else ""
    print(f"{100 * i / max(1, ti):5.1f}% inst {100 * s / max(1, ts):5.1f}% stall  L{ln:<4d} {code}")
