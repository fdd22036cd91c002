"""FullyConnected [24576,768] x [N,768]^T at BERT-base b64 shapes with different epilogues (not a bench value): is the FFN-up time the GELU?
usage: python tools/gpu_fc_epi.py"""
import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, Function

rng = np.random.default_rng(0)
for (k, n, act, res) in [(768, 3072, "none", 0), (768, 3072, "relu", 0), (768, 3072, "gelu", 0), (3072, 768, "none", 1), (768, 768, "none", 1), (768, 768, "none", 0)]:
    f = Function("fc")
    x = f.parameter((64, 384, k), name="x")
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    y = f.add(f.matmul(x, w, transpose_b=True), f.constant(rng.uniform(-1, 1, (n,)).astype(np.float32)))
    if act == "relu": y = f.relu(y)
    if act == "gelu": y = f.gelu(y, "ERF")
    if res:
        r = f.parameter((64, 384, n), name="r")
        y = f.add(y, r)
    f.result(y)
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    req = net.create_infer_request()
    req.infer()
    req.time_launches(5)
    pc = req.get_perf_counts()
    for e in net.get_exec_graph_info():
        if e["type"] == "FullyConnected":
            us = pc[e["name"]]["real_time"]
            print("K %4d N %4d act %-4s res %d: %7.1f us  %6.1f TFLOP/s  %s" % (k, n, act, res, us, e["flops"] / us / 1e6, e["impl"]), flush=True)
