"""One-off differential run on the B200: seeded random graphs of tests/random_graphs.py as IR v10 through the reference Core +
libb200Plugin.so (nGraph passes, lowering: the attribute mapping of asymmetric pads, per-axis strides / dilations, rounding modes, slice
plans) against the numpy restatement.  usage: python tools/gpu_plugin_fuzz.py [graphs per family]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import random_graphs          # noqa: E402
from tests.util import rel_l2_floor, run_oracle   # noqa: E402
from tools import refrun                 # noqa: E402


def main():
    per = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    bad = total = 0
    for family in ("build_pool", "build_conv_corner", "build_misc", "build_sequence"):
        worst = 0.0
        for seed in range(1, 1 + 7 * per, 7):
            f, inputs = getattr(random_graphs, family)(seed)
            ref = run_oracle(f, inputs)
            got, _ = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"})
            total += 1
            for k, v in ref.items():
                if v.dtype.kind in "iu":
                    ok = np.array_equal(got[k], v)
                else:
                    err = rel_l2_floor(got[k], v)
                    worst = max(worst, err)
                    ok = got[k].shape == v.shape and err <= 1e-4
                if not ok:
                    bad += 1
                    print("MISMATCH", family, seed, k)
        print("%-20s worst rel-L2 %.2e" % (family, worst), flush=True)
    print("graphs %d, mismatches %d" % (total, bad))


if __name__ == "__main__":
    main()
