"""Print the key figures of a bench.py JSON line: python tools/bench_print.py file.json"""
import json
import sys

for path in sys.argv[1:]:
    line = [l for l in open(path).read().splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    r = d.get("roofline") or {}
    print(f"{path}: value {d['value']:.4g} {d['unit']}  ms/step {d['ms_per_step']:.4f}  e2e {d['e2e']['value']:.4g}  "
          f"sweep {r.get('avg_launch_ms')} ms  frac {r.get('frac')}  share {r.get('share_of_step')}")
