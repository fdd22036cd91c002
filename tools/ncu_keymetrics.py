"""Key metrics per launch from an `ncu --set full` report exported with `ncu -i x.ncu-rep --page raw --csv`:
duration, DRAM / L2 / L1 throughput, L1 hit rate, issue-slot utilisation, achieved occupancy, registers, top warp-stall reasons.
  python tools/ncu_keymetrics.py gpurun_out/<tag>_raw.csv"""
import csv
import sys

KEYS = [("gpu__time_duration.sum", "dur"), ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1%"), ("l1tex__t_sector_hit_rate.pct", "l1hit%"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
        ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "inst"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2->sm B")]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    units = rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    if not stall_cols:
        stall_cols = [h for h in hdr if h.startswith("smsp__average_warp_latency_issue_stalled_") or (h.startswith("smsp__warp_issue_stalled_") and h.endswith("_per_warp_active.pct"))]
    for r in rows[2:]:
        name = r[col["Kernel Name"]][:60]
        out = [name]
        for k, label in KEYS:
            if k in col and r[col[k]] != "":
                out.append("%s=%s%s" % (label, r[col[k]], (" " + units[col[k]]) if label in ("dur", "dram_rd", "dram_wr", "l2->sm B") else ""))
        st = []
        for h in stall_cols:
            try:
                st.append((float(r[col[h]].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "").replace("smsp__warp_issue_stalled_", "").replace("_per_warp_active.pct", "")))
            except ValueError:
                pass
        st.sort(reverse=True)
        print("  ".join(out))
        print("      stalls: " + ", ".join("%s %.2f" % (n, v) for v, n in st[:6]))


if __name__ == "__main__":
    main(sys.argv[1])
