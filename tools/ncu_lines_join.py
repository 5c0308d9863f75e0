"""Join an `ncu --page source --csv` SASS dump (per-instruction executed counts / stall samples) with the line
table of `nvdisasm -g -c` for the same kernel, and print executed warp-instructions and samples per CUDA
source line.  usage: python tools/ncu_lines_join.py src.csv disasm.txt <mangled-name-substring> [N]"""
import csv
import re
import sys
from collections import Counter

src_csv, disasm, name = sys.argv[1:4]
n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(disasm).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith("//---") and name in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith("//---")), len(lines))
cur, per_instr = None, []
for l in lines[start:end]:
    m = re.search(r"//## File \"([^\"]+)\", line (\d+)", l)
    if m:
        cur = int(m.group(2))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        per_instr.append(cur)
rows = list(csv.reader(open(src_csv)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
print(f"SASS rows {len(body)}, disasm instructions {len(per_instr)}")
ex, sm = Counter(), Counter()
for k, r in enumerate(body):
    ln = per_instr[k] if k < len(per_instr) else -1
    ex[ln] += int(r[col["Instructions Executed"]] or 0)
    sm[ln] += int(r[col["# Samples"]] or 0)
te, ts = sum(ex.values()), sum(sm.values())
srcfile = [l.rstrip("\n") for l in open(sys.argv[5])] if len(sys.argv) > 5 else None
for ln, c in ex.most_common(n):
    text = srcfile[ln - 1].strip()[:90] if srcfile and ln and 0 < ln <= len(srcfile) else ""
    print(f"line {ln:5d}  exec {c:11d} {100 * c / te:5.1f}%   samples {sm[ln]:7d} {100 * sm[ln] / max(ts, 1):5.1f}%   {text}")
