"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples, and
instruction-count share by opcode.  usage: python tools/ncu_hot.py src.csv [N]"""
import csv
import sys
from collections import Counter

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
S, I, SRC = col["# Samples"], col["Instructions Executed"], col["Source"]
tot_s = sum(int(r[S] or 0) for r in body)
tot_i = sum(int(r[I] or 0) for r in body)
print(f"total samples {tot_s}, warp instructions {tot_i}, SASS lines {len(body)}")
print("--- top by samples")
for k, r in sorted(enumerate(body), key=lambda kr: -int(kr[1][S] or 0))[:n]:
    print(f"{k:5d} {int(r[S]):8d} {100*int(r[S])/max(tot_s,1):5.1f}%  exec {int(r[I]):10d}  {r[SRC].strip()[:90]}")
ops = Counter()
for r in body:
    op = r[SRC].strip().split()
    op = [t for t in op if not t.startswith("@")]
    ops[op[0].split(".")[0] if op else "?"] += int(r[I] or 0)
print("--- executed warp-instructions by opcode")
for op, c in ops.most_common(25):
    print(f"{op:12s} {c:12d} {100*c/max(tot_i,1):5.1f}%")
