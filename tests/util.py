import numpy as np


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def run_b200(func, inputs, precision="FP32", config=None):
    from dldt_b200 import Core
    cfg = {"B200_INFERENCE_PRECISION": precision}
    cfg.update(config or {})
    net = Core().load_network(func, "B200", cfg)
    req = net.create_infer_request()
    return req.infer(inputs), net


def run_oracle(func, inputs):
    from oracle import interp
    return interp.run(func, inputs)
