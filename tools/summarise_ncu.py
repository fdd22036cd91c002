"""Compact text summaries of the ncu CSV logs a GPU round brings back (tools/gpu_round.sh), for profiles/.
  python tools/summarise_ncu.py launches gpurun_out/<tag>_launches.csv        -> per-kernel-name launch count / total / share
  python tools/summarise_ncu.py traffic  gpurun_out/<tag>_conv_traffic.csv     -> one line per launch: time, DRAM bytes, L2->SM bytes, tensor pipe %"""
import csv, re, sys
from collections import OrderedDict, defaultdict


def rows(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    return list(csv.DictReader(lines))


def short(name):
    name = re.sub(r"\(.*", "", name.replace("void ", "").replace("b200::", "").replace("<unnamed>::", ""))
    return name[:70]


def launches(path):
    tot = defaultdict(lambda: [0, 0.0])
    order = []
    for r in rows(path):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else (v if r["Metric Unit"] in ("us", "usecond") else v * 1e3)
        k = short(r["Kernel Name"])
        if k not in tot:
            order.append(k)
        tot[k][0] += 1
        tot[k][1] += us
    total = sum(v[1] for v in tot.values())
    print("%-72s %6s %12s %7s" % ("kernel", "count", "total us", "share"))
    for k in sorted(order, key=lambda k: -tot[k][1]):
        print("%-72s %6d %12.1f %6.1f%%" % (k, tot[k][0], tot[k][1], 100 * tot[k][1] / total))
    print("%-72s %6d %12.1f" % ("all launches (ncu: serialised, cold cache)", sum(v[0] for v in tot.values()), total))


def traffic(path):
    per = OrderedDict()
    for r in rows(path):
        d = per.setdefault(r["ID"], {"name": short(r["Kernel Name"])})
        d[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    print("%-4s %-58s %9s %10s %10s %10s %8s" % ("id", "kernel", "us", "dram rd MB", "dram wr MB", "L2->SM MB", "tensor%"))
    agg = defaultdict(float)
    for i, d in per.items():
        t, tu = d.get("gpu__time_duration.sum", (0, "ns"))
        us = t / 1e3 if tu.startswith("n") else t
        rd = d.get("dram__bytes_read.sum", (0, ""))[0] / 1e6
        wr = d.get("dram__bytes_write.sum", (0, ""))[0] / 1e6
        xb = d.get("l1tex__m_xbar2l1tex_read_bytes.sum", (0, ""))[0] / 1e6
        tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", (0, ""))[0]
        print("%-4s %-58s %9.1f %10.1f %10.1f %10.1f %8.1f" % (i, d["name"][:58], us, rd, wr, xb, tp))
        agg["us"] += us; agg["rd"] += rd; agg["wr"] += wr; agg["xb"] += xb; agg["n"] += 1
    print("sum over %d launches: %.1f us, dram read %.1f MB, dram write %.1f MB, L2->SM %.1f MB" % (agg["n"], agg["us"], agg["rd"], agg["wr"], agg["xb"]))


if __name__ == "__main__":
    {"launches": launches, "traffic": traffic}[sys.argv[1]](sys.argv[2])
