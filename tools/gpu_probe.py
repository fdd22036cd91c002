"""First-contact GPU probe: per-case parity numbers + a ResNet-50 b256 timing (not a bench value)."""
import sys, time, traceback
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, Function, zoo
from oracle import interp
from tests.util import rel_l2

def conv_case(n, c, h, w, o, k, s, p, d, prec):
    rng = np.random.default_rng(1)
    f = Function("conv")
    x = f.parameter((n, c, h, w))
    y = f.convolution(x, (rng.standard_normal((o, c, k, k)) / np.sqrt(c * k * k)).astype(np.float32), (s, s), (p, p), (p, p), (d, d))
    y = f.relu(f.add(y, rng.standard_normal((1, o, 1, 1)).astype(np.float32)))
    f.result(y)
    xin = rng.standard_normal((n, c, h, w)).astype(np.float32)
    ref = interp.run(f, {"Parameter_0": xin})["Relu_0"]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
    got = net.create_infer_request().infer({"Parameter_0": xin})["Relu_0"]
    return rel_l2(got, ref)

import os
for prec in (() if os.environ.get("PROBE_SKIP_CONV") else ("FP32", "BF16", "FP16")):
    for case in [(2, 64, 14, 14, 64, 1, 1, 0, 1), (2, 64, 14, 14, 128, 3, 1, 1, 1), (1, 128, 15, 13, 256, 3, 2, 1, 1), (2, 3, 32, 32, 64, 7, 2, 3, 1)]:
        try:
            print("conv", prec, case, "rel_l2 %.3e" % conv_case(*case, prec), flush=True)
        except Exception as e:
            print("conv", prec, case, "FAILED", repr(e)[:300], flush=True)

try:
    for prec in (("BF16",) if os.environ.get("PROBE_SKIP_FP32") else ("BF16", "FP32")):
        B = 256 if prec == "BF16" else 32
        f = zoo.resnet50(batch=B)
        t = time.time()
        net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
        req = net.create_infer_request(input_precisions={"data": "U8"})
        print("R50", prec, "compile %.1fs arena %.0f MB kernels %d" % (time.time() - t, net.stat("arena_bytes") / 1e6, net.num_exec_nodes), flush=True)
        req.input_blobs["data"][...] = np.random.default_rng(0).integers(0, 256, req.input_blobs["data"].shape, dtype=np.uint8)
        req.infer()
        for it in range(2):
            ms = req.time_launches(10) / 10
            print("R50 %s b%d: %.3f ms/step -> %.0f img/s (device resident)" % (prec, B, ms, B / ms * 1e3), flush=True)
        ms = req.time_launches(10, end_to_end=True) / 10
        print("R50 %s b%d: %.3f ms/step -> %.0f img/s (H2D+graph+D2H, 1 stream)" % (prec, B, ms, B / ms * 1e3), flush=True)
        if prec == "BF16":
            pc = req.get_perf_counts()
            eg = net.get_exec_graph_info()
            rows = sorted(((pc[e["name"]]["real_time"], e) for e in eg), key=lambda t: -t[0])
            tot = sum(r[0] for r in rows)
            print("per-kernel total %.3f ms" % (tot / 1e3))
            for us, e in rows[:14]:
                print("  %8.1f us %5.1f%% %-14s %-28s %7.1f TFLOP/s %7.1f GB/s  %s" % (us, 100 * us / tot, e["type"], e["impl"], e["flops"] / us / 1e6, e["bytes"] / us / 1e3, e["name"]))
            print("full schedule (execution order):")
            for e in eg:
                us = pc[e["name"]]["real_time"]
                print("  %8.1f us %-14s %-40s %7.1f TFLOP/s %7.1f GB/s  %s" % (us, e["type"], e["impl"], e["flops"] / us / 1e6, e["bytes"] / us / 1e3, e["name"]))
except Exception as e:
    traceback.print_exc()
