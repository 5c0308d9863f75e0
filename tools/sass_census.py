"""Per-kernel census of the Blackwell-specific SASS mnemonics in the built library
(B200_PROFILING.md "What proves a Blackwell-native kernel"): tcgen05.mma -> UTC*MMA, tcgen05.ld/st ->
LDTM / STTM, TMA -> UTMALDG / UTMASTG / UBLKCP, mbarrier -> SYNCS, warp reduce -> REDUX, legacy tensor
path -> HMMA (must be absent).
    python tools/sass_census.py > profiles/r02_sass_census.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "librecommender_b200", "libb200reco.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTMACMDFLUSH",
         "SYNCS", "REDUX", "HMMA", "LDGSTS", "FMNMX3", "ATOMG", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            kernels[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + ".") or (w in ("UTCHMMA", "UTCQMMA", "UTCIMMA") and op.startswith(w)):
                    kernels[cur][w] += 1
    demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
    print(f"# SASS census of {os.path.relpath(LIB, ROOT)} (sm_100a), {len(kernels)} kernels")
    print("# columns: instructions | " + " ".join(WATCH))
    totals = collections.Counter()
    for (name, c), dm in zip(kernels.items(), demangle):
        short = re.sub(r"\(.*", "", dm)[:90]
        cols = " ".join(f"{w}={c[w]}" for w in WATCH if c[w])
        print(f"{short:92s} {c['_total']:6d} | {cols}")
        totals.update(c)
    print("# totals: " + " ".join(f"{w}={totals[w]}" for w in WATCH))
    assert totals["HMMA"] == 0, "legacy mma.sync path present"


if __name__ == "__main__":
    sys.exit(main())
