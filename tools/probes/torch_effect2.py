"""Diagnostic 2: per-kernel times in the slow (context created by our static cudart) and fast (dynamic cudart first) process states."""
import ctypes, os, sys
sys.path.insert(0, ".")
import numpy as np
mode = sys.argv[1]
if mode == "rt_free0":
    ctypes.CDLL("libcudart.so.12").cudaFree(ctypes.c_void_p(0))
from dldt_b200 import Core, Function, zoo
if mode == "warm_fp32":
    f0 = zoo.resnet50(batch=2, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
    n0 = Core().load_network(f0, "B200.0", {"B200_INFERENCE_PRECISION": "FP32"})
    r0 = n0.create_infer_request(); r0.infer(); del r0, n0
if mode == "warm_many":
    for k in range(12):
        f0 = zoo.resnet50(batch=2, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
        n0 = Core().load_network(f0, "B200.0", {"B200_INFERENCE_PRECISION": ["FP32", "BF16", "FP16"][k % 3]})
        r0 = n0.create_infer_request(); r0.infer(); del r0, n0
f = zoo.resnet50(batch=256)
net = Core().load_network(f, "B200.0", {"B200_INFERENCE_PRECISION": "BF16"})
req = net.create_infer_request(input_precisions={"data": "U8"})
req.input_blobs["data"][...] = np.random.default_rng(0).integers(0, 256, req.input_blobs["data"].shape, dtype=np.uint8)
req.infer()
print(mode, [round(req.time_launches(10) / 10, 3) for _ in range(2)], flush=True)
pc = req.get_perf_counts(); eg = net.get_exec_graph_info()
print(mode, "per-kernel:", " ".join("%s=%.0f" % (e["name"], pc[e["name"]]["real_time"]) for e in eg[:16]), flush=True)
