"""Hot SASS instructions of one kernel from `ncu --page source --csv` output (stall samples)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.008
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]; data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
ia = hdr.index("Source"); isamp = hdr.index("Warp Stall Sampling (All Samples)"); iex = hdr.index("Instructions Executed")
num = lambda x: int(x) if x.isdigit() else 0
tot = sum(num(r[isamp]) for r in data)
print("total samples", tot, "static instrs", len(data), "dynamic warp instrs", sum(num(r[iex]) for r in data))
acc = 0
for n, r in enumerate(data):
    s = num(r[isamp]); acc += s
    if s > tot * thr:
        print(f"{n:5d} {s:6d} {100*s/tot:5.1f}% cum {100*acc/tot:5.1f}%  {r[ia].strip()[:72]:72s} ex={r[iex]}")
