"""Top stall sites of one launch of an .ncu-rep (SASS view): python tools/sass_stalls.py <rep> <launch index> [min samples]"""
import csv, subprocess, sys, io
rep, k = sys.argv[1], sys.argv[2]
thr = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--launch-skip", k, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
print(rows[0][1][:110])
hdr, data = rows[1], rows[2:]
i_s, i_src, i_ex = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
print("total samples", sum(int(r[i_s] or 0) for r in data))
for n, r in enumerate(data):
    c = int(r[i_s] or 0)
    if c >= thr:
        st = {h.replace("stall_", ""): r[hdr.index(h)] for h in hdr if h.startswith("stall_") and "Not Issued" not in h and r[hdr.index(h)] not in ("0", "")}
        print(n, c, r[i_ex], r[i_src][:70], st)
