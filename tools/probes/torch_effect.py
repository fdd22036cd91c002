"""Diagnostic: what in the process history changes our kernel timings (6.03 ms vs 4.55 ms per ResNet-50 b256 step)?"""
import ctypes, os, sys, time
sys.path.insert(0, ".")
import numpy as np
mode = sys.argv[1]
if mode == "eager_modules":
    os.environ["CUDA_MODULE_LOADING"] = "EAGER"
if mode in ("torch_sync",):
    import torch
    torch.cuda.synchronize(0)
if mode.startswith("rt_"):
    rt = ctypes.CDLL("libcudart.so.12")
    if mode == "rt_free0":
        rt.cudaFree(ctypes.c_void_p(0))            # context creation by the DYNAMIC runtime, nothing else
    if mode == "rt_malloc":
        p = ctypes.c_void_p()
        rt.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(1 << 20)); rt.cudaFree(p)
from dldt_b200 import Core, Function, zoo
if mode == "warm_tiny":
    f0 = zoo.resnet50(batch=2, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
    n0 = Core().load_network(f0, "B200.0", {"B200_INFERENCE_PRECISION": "BF16"})
    r0 = n0.create_infer_request(); r0.infer(); del r0, n0
f = zoo.resnet50(batch=256)
net = Core().load_network(f, "B200.0", {"B200_INFERENCE_PRECISION": "BF16"})
if mode == "second_request":
    tmp = net.create_infer_request(input_precisions={"data": "U8"}); tmp.infer(); del tmp
req = net.create_infer_request(input_precisions={"data": "U8"})
req.input_blobs["data"][...] = np.random.default_rng(0).integers(0, 256, req.input_blobs["data"].shape, dtype=np.uint8)
req.infer()
print(mode, [round(req.time_launches(10) / 10, 3) for _ in range(3)], flush=True)
if mode == "recreate":
    del req
    req = net.create_infer_request(input_precisions={"data": "U8"}); req.infer()
    print(mode, "second request:", [round(req.time_launches(10) / 10, 3) for _ in range(3)], flush=True)
    net2 = Core().load_network(f, "B200.0", {"B200_INFERENCE_PRECISION": "BF16"})
    req2 = net2.create_infer_request(input_precisions={"data": "U8"}); req2.infer()
    print(mode, "second network:", [round(req2.time_launches(10) / 10, 3) for _ in range(3)], flush=True)
