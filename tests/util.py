import numpy as np


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.linalg.norm(a - b))


def rel_l2_floor(a, b, floor=1e-3):
    """rel-L2 with the reference norm floored at `floor` per element: outputs that are mathematically zero (the mean of mean-centred
    data in a random graph) are compared absolutely instead of noise against noise."""
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), floor * np.sqrt(max(b.size, 1))))


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
