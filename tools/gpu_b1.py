"""Batch-1 ResNet-50 device-only latency under different kernel-selection knobs (not a bench value)."""
import os, subprocess, sys
import numpy as np
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    from dldt_b200 import Core, zoo
    net = Core().load_network(zoo.resnet50(batch=int(os.environ.get("B1_BATCH", "1"))), "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    req = net.create_infer_request(input_precisions={"data": "U8"})
    req.input_blobs["data"][...] = np.random.default_rng(0).integers(0, 256, req.input_blobs["data"].shape, dtype=np.uint8)
    req.infer()
    ms = min(req.time_launches(50) / 50 for _ in range(5))
    impls = {}
    for e in net.get_exec_graph_info():
        k = e["impl"].split("_")[2] if e["impl"].startswith("tcgen05") else e["impl"]
        impls[k] = impls.get(k, 0) + 1
    print("%.4f ms/inference device-only  %s" % (ms, impls))
    sys.exit(0)
for cfg in sys.argv[1:]:
    env = dict(os.environ)
    for kv in cfg.split():
        a, b = kv.split("=")
        env[a] = b
    r = subprocess.run([sys.executable, __file__, "--one"], env=env, capture_output=True, text=True, timeout=300)
    print("[%s] %s %s" % (cfg or "default", r.stdout.strip(), r.stderr.strip()[-200:] if r.returncode else ""), flush=True)
