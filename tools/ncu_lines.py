"""Aggregate `ncu --page source --print-source cuda,sass --csv` by CUDA source line.
usage: python tools/ncu_lines.py dump.csv [N]"""
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
        S = hdr.index("# Samples")
        I = hdr.index("Instructions Executed")
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[2] != "-":      # SASS rows under a CUDA line: skip (the CUDA row carries the totals)
        continue
    try:
        key = (cur_file, int(r[0]), r[1].strip()[:100])
        agg[key] = (int(r[S] or 0), int(r[I] or 0))
    except ValueError:
        pass
ts = sum(v[0] for v in agg.values())
ti = sum(v[1] for v in agg.values())
print(f"samples {ts}  warp-instructions {ti}")
for (f, ln, src), (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:n]:
    print(f"{100*s/max(ts,1):5.1f}% samp {100*i/max(ti,1):5.1f}% inst  {f}:{ln}  {src}")
