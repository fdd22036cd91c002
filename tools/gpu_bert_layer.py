import sys
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo
net = Core().load_network(zoo.bert_base(batch=64), "B200", {"B200_INFERENCE_PRECISION": "BF16"})
req = net.create_infer_request()
rng = np.random.default_rng(0)
for nm, shape, _ in net.inputs:
    req.set_blob(nm, np.ones(shape, np.int32) if "mask" in nm else (np.zeros(shape, np.int32) if "type" in nm else rng.integers(0, 30000, shape, dtype=np.int32)))
req.infer()
pc = req.get_perf_counts()
for e in net.get_exec_graph_info()[:40]:
    print("%8.1f us %-14s %-36s %s" % (pc[e["name"]]["real_time"], e["type"], e["impl"], e["name"]))
