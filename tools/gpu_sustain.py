"""How the ResNet-50 b256 step time evolves under sustained replay (not a bench value): 30 bursts of 10 graph replays back to back, then the
same with a 200 ms pause before each burst; SM clock / power / throttle reasons sampled by nvidia-smi between bursts."""
import subprocess, sys, time
import numpy as np
sys.path.insert(0, ".")
from dldt_b200 import Core, zoo

def smi():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active", "--format=csv,noheader,nounits"],
                              capture_output=True, text=True, timeout=5).stdout.strip()
    except Exception as e:
        return "n/a"

net = Core().load_network(zoo.resnet50(batch=256), "B200", {"B200_INFERENCE_PRECISION": "BF16"})
req = net.create_infer_request(input_precisions={"data": "U8"})
req.input_blobs["data"][...] = np.random.default_rng(0).integers(0, 256, req.input_blobs["data"].shape, dtype=np.uint8)
req.infer()
for mode in ("back-to-back", "200 ms pause before each burst"):
    ts = []
    for k in range(30):
        if mode != "back-to-back":
            time.sleep(0.2)
        ts.append(req.time_launches(10) / 10)
        if k % 6 == 5:
            print("  [%s] burst %2d: %.3f ms/step   smi: %s" % (mode, k, ts[-1], smi()), flush=True)
    print("[%s] first 3: %s  last 3: %s  min %.3f max %.3f" % (mode, ["%.3f" % t for t in ts[:3]], ["%.3f" % t for t in ts[-3:]], min(ts), max(ts)), flush=True)
