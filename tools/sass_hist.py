"""Opcode histogram of one kernel's SASS:  python tools/sass_hist.py <kernel substring> [lib]  (static counts, not dynamic)"""
import collections, re, subprocess, sys
pat = sys.argv[1]; lib = sys.argv[2] if len(sys.argv) > 2 else "mpas_model_b200/csrc/libmpasb.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name = None; hist = collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: name = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name and pat in name:
        hist[name][m.group(1).split(".")[0]] += 1
for n, h in hist.items():
    tot = sum(h.values())
    print(n[:100], tot, "instructions")
    print("   " + "  ".join(f"{k}:{v}" for k, v in h.most_common(30)))
