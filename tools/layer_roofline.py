"""Per-layer roofline reading of the compiled ResNet-50 b256 schedule WITHOUT a GPU: algorithmic FLOPs and bytes of every exec node (from
the compiler's own accounting, b200_network_exec_node) against the measured peaks of MEASURED_PEAKS.json, next to the per-launch times of
a committed ncu launch table (profiles/r2_run8_conv_traffic.txt: cold cache, serialised - shares, not absolutes).
usage: python tools/layer_roofline.py [profiles/r2_run8_conv_traffic.txt]"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dldt_b200 import Core, zoo   # noqa: E402


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    tf, gbs = peaks["bf16_tflops_sustained"], peaks["hbm_gbs"]
    net = Core().load_network(zoo.resnet50(batch=256), "B200", {"B200_INFERENCE_PRECISION": "BF16"}, dry=True)
    eg = net.get_exec_graph_info()
    measured = []
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r2_run8_conv_traffic.txt")
    if os.path.exists(path):
        for line in open(path):
            m = re.match(r"\s*(\d+)\s+\S.*?\s+([\d.]+)\s+[\d.]+\s+[\d.]+\s+[\d.]+\s+[\d.]+\s*$", line)
            if m and int(m.group(1)) < 50:
                measured.append(float(m.group(2)))
    conv = [e for e in eg if e["type"] in ("Convolution", "FullyConnected")]
    print("peaks: %.1f TFLOP/s sustained bf16, %.1f GB/s copy   (%d exec nodes, %d tensor-core launches)" % (tf, gbs, len(eg), len(conv)))
    print("%-26s %-34s %8s %8s %8s %8s %6s" % ("layer", "kernel", "flop us", "byte us", "ideal", "ncu us", "bound"))
    tot_ideal = tot_meas = tot_flop = tot_byte = 0.0
    k = 0
    for e in eg:
        ft = e["flops"] / (tf * 1e12) * 1e6
        bt = e["bytes"] / (gbs * 1e9) * 1e6
        ideal = max(ft, bt)
        meas = ""
        if e["type"] in ("Convolution", "FullyConnected") and k < len(measured):
            meas = "%8.1f" % measured[k]
            tot_meas += measured[k]
            k += 1
        tot_ideal += ideal
        tot_flop += ft
        tot_byte += bt
        print("%-26s %-34s %8.1f %8.1f %8.1f %8s %6s" % (e["name"][:26], e["impl"].replace("tcgen05_bf16_", "")[:34], ft, bt, ideal, meas, "TC" if ft >= bt else "HBM"))
    print("%-61s %8.1f %8.1f %8.1f %8.1f" % ("sum (per-layer max(flop, byte) = floor of THIS schedule)", tot_flop, tot_byte, tot_ideal, tot_meas))
    print("whole-step floors: %.3f ms if only FLOPs counted, %.3f ms if only bytes, %.3f ms per-layer max; measured in graph replay: 2.736 ms"
          % (tot_flop / 1e3, tot_byte / 1e3, tot_ideal / 1e3))


if __name__ == "__main__":
    main()
