"""CPU half of the randomized differential tests (tests/random_graphs.py): every seeded random network must lower and plan in all three
execution precisions without a device, keep one exec node per unfused op at most, and agree with the oracle's shape inference."""
import numpy as np
import pytest

from dldt_b200 import Core
from tests import random_graphs
from tests.util import rel_l2_floor, run_oracle


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


@pytest.mark.parametrize("family", ["build_sequence", "build_misc"])
@pytest.mark.parametrize("seed", range(0, 120, 3))
def test_random_sequence_and_misc_graphs_lower_in_every_precision(seed, family):
    f, inputs = getattr(random_graphs, family)(seed)
    for precision in ("FP32", "BF16", "FP16"):
        net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": precision}, dry=True)
        assert len(net.get_exec_graph_info()) > 0


@pytest.mark.parametrize("family,seed", [("build", s) for s in list(range(1, 40, 4)) + list(range(1001, 1020, 4))] +
                         [("build_sequence", s) for s in range(0, 30, 3)] + [("build_misc", s) for s in range(0, 60, 3)] +
                         [("build_nms", s) for s in range(0, 60, 2)] + [("build_detection_output", s) for s in range(0, 60, 2)] +
                         [("build_conv_corner", s) for s in range(0, 120, 3)])
def test_oracle_matches_the_reference_cpu_plugin_on_random_graphs(seed, family):
    """Pins the numpy restatement on graphs nobody wrote by hand: the same random networks on the UNMODIFIED MKLDNNPlugin (oracle/_ref,
    fp32) - beyond the fixed golden cases of tests/test_oracle_golden.py."""
    from tools import refrun
    if not refrun.available():
        pytest.skip("oracle/_ref not built")
    f, inputs = getattr(random_graphs, family)(seed)
    mine = run_oracle(f, inputs)
    ref, _ = refrun.run(f, inputs, device="CPU")
    assert set(mine) == set(ref)
    for k, v in ref.items():
        if v.dtype.kind in "iu":
            np.testing.assert_array_equal(mine[k], v)
            continue
        if family == "build_detection_output":      # rows [image, label, conf, box]: ids (and the -1 terminator) exact
            np.testing.assert_array_equal(mine[k].reshape(-1, 7)[:, :2], v.reshape(-1, 7)[:, :2])
        err = rel_l2_floor(mine[k], v)
        assert err <= 1e-5, "seed %d %s: %.3e" % (seed, k, err)


@pytest.mark.parametrize("seed", range(0, 120, 2))
def test_pooling_corner_cases_match_the_ngraph_reference(seed):
    """Non-square kernels / strides, asymmetric padding, both rounding modes, include / exclude padding: the restatement against the nGraph
    reference kernels (device TEMPLATE of the unmodified reference).  600 such graphs were checked while the generator was written
    (<= 1.4e-7); MKLDNNPlugin agrees on all but four of them: three single-channel cases (its C = 1 path divides an include-pad window by
    the clipped size) and one two-pool chain where it disagrees with the nGraph reference on its own."""
    from tools import refrun
    if not refrun.available():
        pytest.skip("oracle/_ref not built")
    f, inputs = random_graphs.build_pool(seed)
    mine = run_oracle(f, inputs)
    for precision in ("FP32", "BF16"):
        Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": precision}, dry=True)
    ref, _ = refrun.run(f, inputs, device="TEMPLATE")
    for k, v in ref.items():
        assert mine[k].shape == v.shape
        assert rel_l2_floor(mine[k], v) <= 1e-5, "seed %d %s" % (seed, k)
