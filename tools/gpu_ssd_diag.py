import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo
f = zoo.ssd300_resnet34(batch=2)
rng = np.random.default_rng(7)
x = rng.uniform(0, 1, (2, 3, 300, 300)).astype(np.float32)
for prec in ("FP32", "FP16", "BF16"):
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
    out = list(net.create_infer_request().infer({"data": x}).values())[0].reshape(-1, 7)
    valid = out[out[:, 0] >= 0]
    bad = ~np.isfinite(out).all(axis=1)
    print(prec, "rows", out.shape[0], "valid", len(valid), "non-finite rows", int(bad.sum()), "first bad", out[bad][:2] if bad.any() else None, flush=True)
    print("   conf range", valid[:, 2].min() if len(valid) else None, valid[:, 2].max() if len(valid) else None, "box range", np.nanmin(valid[:, 3:]), np.nanmax(valid[:, 3:]))
