"""Per-layer table of one of the benchmark configurations on the GPU: kernel, time (event-bracketed eager pass), algorithmic bytes and
FLOPs, achieved GB/s and TFLOP/s.  usage: python tools/gpu_layers.py mobilenet|resnet|bert|ssd [batch]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo

kind = sys.argv[1]
mk = {"mobilenet": (zoo.mobilenet_v2, 64, "BF16"), "resnet": (zoo.resnet50, 256, "BF16"), "bert": (zoo.bert_base, 64, "BF16"),
      "ssd": (lambda batch: zoo.ssd300_resnet34(batch=batch, head_scale=0.02), 32, "FP16")}[kind]
batch = int(sys.argv[2]) if len(sys.argv) > 2 else mk[1]
net = Core().load_network(mk[0](batch=batch), "B200", {"B200_INFERENCE_PRECISION": mk[2]})
req = net.create_infer_request()
rng = np.random.default_rng(0)
for nm, shape, _ in net.inputs:
    if kind == "bert":
        req.set_blob(nm, np.ones(shape, np.int32) if "mask" in nm else (np.zeros(shape, np.int32) if "type" in nm else rng.integers(0, 30000, shape, dtype=np.int32)))
    else:
        req.set_blob(nm, rng.uniform(0, 1, shape).astype(np.float32))
req.infer()
ms = min(req.time_launches(10) / 10 for _ in range(3))
pcs = [req.get_perf_counts() for _ in range(3)]
eg = net.get_exec_graph_info()
tot = 0.0
print("%s batch %d: %.3f ms/step (graph replay)" % (kind, batch, ms))
for e in eg:
    us = min(pc[e["name"]]["real_time"] for pc in pcs)
    tot += us
    print("%8.1f us %8.1f MB %7.0f GB/s %7.1f TF/s  %-16s %-46s %s" % (us, e["bytes"] / 1e6, e["bytes"] / us / 1e3 if us else 0, e["flops"] / us / 1e6 if us else 0,
                                                                   e["type"], e["impl"], e["name"]))
print("sum of eager kernel times %.1f us" % tot)
