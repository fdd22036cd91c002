"""Warp-stall samples of one launch of an .ncu-rep aggregated per SOURCE line of the kernel (outermost inlining frame), by correlating the
n-th SASS instruction of ncu's source page with the n-th instruction of `nvdisasm -gi` on the cubin extracted from the built library.
  python tools/stall_lines.py <rep> <launch index> <cubin> <mangled-name substring> [min samples]"""
import csv, io, re, subprocess, sys
from collections import defaultdict

rep, k, cubin, sub = sys.argv[1:5]
thr = int(sys.argv[5]) if len(sys.argv) > 5 else 20
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
lines, cur, inside = [], None, False
pend = []
for l in dis:
    if l.startswith(".text."):
        inside = sub in l
        pend = []
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        pend.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    if re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
        if pend:
            cur = (pend[-1], pend[0])     # (outermost frame, innermost frame)
            pend = []
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", k, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:120])
hdr, data = rows[1], rows[2:]
i_s = hdr.index("# Samples")
stall_cols = [(i, h.replace("stall_", "")) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
print("instructions: ncu %d, nvdisasm %d" % (len(data), len(lines)))
agg = defaultdict(lambda: [0, defaultdict(int), defaultdict(int)])
total = 0
for n, r in enumerate(data):
    c = int(r[i_s] or 0)
    if not c or n >= len(lines) or lines[n] is None:
        continue
    total += c
    outer, inner = lines[n]
    a = agg[outer]
    a[0] += c
    for i, name in stall_cols:
        v = int(r[i] or 0)
        if v:
            a[1][name] += v
    a[2][inner] += c
print("total samples", total)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    if a[0] < thr:
        continue
    st = ", ".join("%s %d" % (n, v) for n, v in sorted(a[1].items(), key=lambda x: -x[1])[:4])
    inner = ", ".join("%s:%d %d" % (f, ln, v) for (f, ln), v in sorted(a[2].items(), key=lambda x: -x[1])[:3])
    print("%s:%-5d %6d %5.1f%%  [%s]  <- %s" % (key[0], key[1], a[0], 100.0 * a[0] / total, st, inner))
