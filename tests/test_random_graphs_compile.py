"""CPU half of the randomized differential tests (tests/random_graphs.py): every seeded random network must lower and plan in all three
execution precisions without a device, keep one exec node per unfused op at most, and agree with the oracle's shape inference."""
import numpy as np
import pytest

from dldt_b200 import Core
from tests import random_graphs
from tests.util import run_oracle


@pytest.mark.parametrize("seed", list(range(0, 200, 4)) + list(range(1000, 1100, 4)))
def test_random_graph_lowers_in_every_precision(seed):
    f, inputs = random_graphs.build(seed)
    ref = run_oracle(f, inputs)
    for precision in ("FP32", "BF16", "FP16"):
        net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": precision}, dry=True)
        eg = net.get_exec_graph_info()
        assert 0 < len(eg) <= len(f.nodes)
        assert {name: shape for name, shape, _ in net.outputs} == {k: v.shape for k, v in ref.items()}
    for v in ref.values():
        assert np.isfinite(v).all()
