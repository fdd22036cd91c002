"""Side-by-side per-kernel table from a tools/gpu_r50_ab.py log."""
import re, sys
txt = open(sys.argv[1]).read()
secs = re.split(r"===== ", txt)[1:]
tabs = []
for s in secs:
    rows = []
    for l in s.splitlines():
        m = re.match(r"\s+([\d.]+) us (\S+)\s+(\S+)\s+([\d.]+) TFLOP/s\s+([\d.]+) GB/s\s+(\S+)", l)
        if m:
            rows.append((float(m.group(1)), m.group(3).replace("tcgen05_bf16_", "").replace("_persistent", ""), m.group(6), float(m.group(4))))
    if rows:
        tabs.append(rows)
for r in zip(*tabs):
    print("%-15s" % r[0][2] + " | ".join("%6.1f %4.0fT %-22s" % (c[0], c[3], c[1][:22]) for c in r))
print("%-15s" % "total" + " | ".join("%6.1f %28s" % (sum(c[0] for c in t), "") for t in tabs))
