"""Calibration of the 16-bit gate on the op-level golden cases: every case of tests/golden_cases.py through the UNMODIFIED reference
CPU plugin's own bf16 mode (ENFORCE_BF16=YES) against its fp32 mode (ENFORCE_BF16=NO) -> tests/golden/ref_bf16_calibration.json
({case: {output: rel-L2}}).  tests/test_gpu_golden.py uses it to justify every tolerance above the north star's 1e-2: a case may
only exceed 1e-2 by what the reference itself loses on it.  Build container only (needs oracle/_ref)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import golden_cases  # noqa: E402
from tools import refrun        # noqa: E402


def main():
    cal = {}
    for name, fn in sorted(golden_cases.CASES.items()):
        f, inputs = fn()
        try:
            o32, _ = refrun.run(f, inputs, "CPU", {"ENFORCE_BF16": "NO"})
            o16, _ = refrun.run(f, inputs, "CPU", {"ENFORCE_BF16": "YES"})
        except Exception as e:  # noqa: BLE001
            print("%-28s skipped: %s" % (name, str(e).splitlines()[0][:100]))
            continue
        cal[name] = {}
        for k in o32:
            if golden_cases.is_index_output(o32[k]):
                continue
            a, b = np.asarray(o16[k], np.float64).ravel(), np.asarray(o32[k], np.float64).ravel()
            cal[name][k] = float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
        print("%-28s %s" % (name, cal[name]), flush=True)
    path = os.path.join(ROOT, "tests", "golden", "ref_bf16_calibration.json")
    json.dump(cal, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
