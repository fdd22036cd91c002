"""Generate tests/golden/ref_cases.npz: outputs of the UNMODIFIED reference on every case of
tests/golden_cases.py.  Runs in the build container only (needs oracle/_ref built from /root/reference by
oracle/build_ref.sh):  python tools/make_golden.py

  cpu/<case>/<output>       MKLDNNPlugin, device "CPU", ENFORCE_BF16=NO   (the fp32 oracle of BASELINE.md section 5)
  template/<case>/<output>  templatePlugin, device "TEMPLATE" = nGraph reference kernels (ngraph/core/reference), where the
                            plugin accepts the graph
The file also records which cases each device rejected.  Test infrastructure; never on the product path.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import golden_cases  # noqa: E402
from tools import refrun        # noqa: E402


def main():
    assert refrun.available(), "build the reference first: bash oracle/build_ref.sh"
    out, rejected = {}, {"cpu": {}, "template": {}}
    path = os.path.join(ROOT, "tests", "golden", "ref_cases.npz")
    only = set(sys.argv[1:])            # `python tools/make_golden.py case ...`: (re)generate these cases only, keep the rest of the file
    if only:
        old = np.load(path)
        for k in old.files:
            if k != "__meta__" and k.split("/")[1] not in only:
                out[k] = old[k]
        meta = json.loads(bytes(old["__meta__"]).decode())
        for dev in rejected:
            rejected[dev] = {k: v for k, v in meta["rejected"].get(dev, {}).items() if k not in only}
    for name, fn in golden_cases.CASES.items():
        if only and name not in only:
            continue
        f, inputs = fn()
        for dev, key in (("CPU", "cpu"), ("TEMPLATE", "template")):
            try:
                res, _ = refrun.run(f, inputs, device=dev)
            except Exception as e:  # noqa: BLE001
                rejected[key][name] = str(e).strip().splitlines()[0][:300]
                print("%-28s %-8s REJECTED %s" % (name, dev, rejected[key][name][:120]))
                continue
            for oname, arr in res.items():
                out["%s/%s/%s" % (key, name, oname)] = arr
            print("%-28s %-8s %s" % (name, dev, {k: v.shape for k, v in res.items()}))
    out["__meta__"] = np.frombuffer(json.dumps({"rejected": rejected, "reference": "opencv/dldt @ /root/reference, built by oracle/build_ref.sh",
                                                "cpu_config": {"ENFORCE_BF16": "NO"}}).encode(), dtype=np.uint8)
    np.savez_compressed(path, **out)
    print("wrote %s (%.1f KB, %d arrays)" % (path, os.path.getsize(path) / 1e3, len(out)))


if __name__ == "__main__":
    main()
