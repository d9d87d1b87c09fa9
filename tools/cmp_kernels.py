"""Compare per-kernel tables written by tools/quick_bench.py:  python tools/cmp_kernels.py base.txt other.txt [...]"""
import re, sys
def load(f):
    d = {}
    for l in open(f):
        m = re.match(r"(k:\S+)\s+([\d.]+) ms/step\s+(\d+) calls/step\s+([\d.]+) us/call", l)
        if m: d[m.group(1)] = float(m.group(4))
        m = re.match(r"ms/step ([\d.]+)", l)
        if m: d["step(us)"] = float(m.group(1)) * 1000
    return d
tabs = [load(f) for f in sys.argv[1:]]
base = tabs[0]
for k in sorted(base, key=lambda k: -base[k]):
    if k == "step(us)" or any(abs(t.get(k, 0) - base[k]) > 0.03 * base[k] for t in tabs[1:]):
        print(f"{k:28s} " + " ".join(f"{t.get(k, 0):8.1f}" for t in tabs))
