"""Randomized differential test on the B200: seeded random small networks (tests/random_graphs.py) through the C ABI against the numpy
oracle.  fp32 mode at the north star's 1e-4; bf16 / fp16 as a sanity bound (3e-2 / 1e-2: depth and activation mix are random, the
calibrated 16-bit gates live in test_gpu_golden.py / test_gpu_configs.py)."""
import numpy as np
import pytest

from tests import random_graphs
from tests.util import rel_l2, rel_l2_floor, run_b200, run_oracle

pytestmark = pytest.mark.gpu
TOL = {"FP32": 1e-4, "BF16": 3e-2, "FP16": 1e-2}


def _check(f, inputs, precisions, tag, tol=TOL, indices_exact=("FP32",)):
    ref = run_oracle(f, inputs)
    for precision in precisions:
        got, _ = run_b200(f, inputs, precision)
        assert set(got) == set(ref)
        for k in ref:
            assert got[k].shape == ref[k].shape
            if ref[k].dtype.kind in "iu":
                if precision in indices_exact:
                    np.testing.assert_array_equal(got[k], ref[k], err_msg="%s %s %s" % (tag, precision, k))
                continue
            err = rel_l2_floor(got[k], ref[k])
            assert err <= tol[precision], "%s %s %s: rel-L2 %.3e" % (tag, precision, k, err)


@pytest.mark.parametrize("seed", list(range(80)) + list(range(1000, 1080)))
def test_random_graph_matches_oracle(seed):
    f, inputs = random_graphs.build(seed)
    _check(f, inputs, ("FP32", "BF16", "FP16") if seed % 3 == 0 else ("FP32", "BF16"), "conv seed %d" % seed)


@pytest.mark.parametrize("seed", range(80))
def test_random_sequence_model_matches_oracle(seed):
    """Encoder-style networks (Gather, LayerNorm, dense layers, multi-head attention with sequence lengths / head sizes that are not tile
    multiples, GELU FFN): the MatMul + Softmax + LayerNorm path, including the fused-attention fold where the shapes allow it."""
    f, inputs = random_graphs.build_sequence(seed)
    _check(f, inputs, ("FP32", "BF16", "FP16") if seed % 3 == 0 else ("FP32", "BF16"), "sequence seed %d" % seed)


@pytest.mark.parametrize("seed", range(120))
def test_random_elementwise_and_data_movement_matches_oracle(seed):
    """Broadcast binaries, activations, PReLU, transposes, split / concat, reductions over arbitrary axes, reshapes, softmax / MVN over any
    axis, TopK (indices bit-exact in fp32 mode)."""
    f, inputs = random_graphs.build_misc(seed)
    # bf16: a sanity bound only - random chains of x^3, 1 / x, MVN over a handful of elements amplify the 2^-9 rounding of every intermediate
    _check(f, inputs, ("FP32", "BF16"), "misc seed %d" % seed, tol={"FP32": 1e-4, "BF16": 1.5e-1})


@pytest.mark.parametrize("seed", list(range(0, 60, 6)) + list(range(1000, 1030, 6)))
def test_random_graph_through_the_plugin_matches_the_reference_cpu_plugin(seed):
    """Same networks as IR v10 through the reference Core: device B200 (libb200Plugin.so: nGraph passes + lowering + C ABI) against device
    CPU (the UNMODIFIED MKLDNNPlugin, fp32) - the differential test with the real reference on both sides of the drop-in boundary."""
    import os
    from tools import refrun
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not (refrun.available() and os.path.exists(os.path.join(root, "plugin", "lib", "libb200Plugin.so"))):
        pytest.skip("oracle/_ref or plugin/lib did not travel")
    f, inputs = random_graphs.build(seed)
    ref, _ = refrun.run(f, inputs, device="CPU")
    got, _ = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"})
    assert set(got) == set(ref)
    for k in ref:
        assert got[k].shape == ref[k].shape
        err = rel_l2(got[k], ref[k])
        assert err <= 1e-4, "seed %d %s: rel-L2 %.3e" % (seed, k, err)


@pytest.mark.parametrize("seed", range(120))
def test_random_nms_is_bit_exact(seed):
    """NonMaxSuppression-5 with random geometry / thresholds / soft-NMS sigma / score ties / duplicate boxes: selected indices and the
    valid count bit-exact, scores <= 1e-4 (fp32 mode, as the north star asks for index outputs)."""
    f, inputs = random_graphs.build_nms(seed)
    _check(f, inputs, ("FP32",), "nms seed %d" % seed)


@pytest.mark.parametrize("seed", range(120))
def test_random_detection_output_ids_are_exact(seed):
    """DetectionOutput with random prior counts, class counts, top_k / keep_top_k, code types, shared / per-class locations, variance
    placement, clipping: the (image, label) sequence incl. the terminator row exact, confidences and boxes <= 1e-4 (fp32 mode)."""
    f, inputs = random_graphs.build_detection_output(seed)
    ref = run_oracle(f, inputs)
    got, _ = run_b200(f, inputs, "FP32")
    (name, r), = ref.items()
    g, r = got[name].reshape(-1, 7), r.reshape(-1, 7)
    np.testing.assert_array_equal(g[:, :2], r[:, :2], err_msg="detout seed %d" % seed)
    assert rel_l2(g[:, 2:], r[:, 2:]) <= 1e-4


@pytest.mark.parametrize("seed", list(range(0, 60, 2)) + list(range(1000, 1030, 2)))
@pytest.mark.parametrize("blob", ["U8", "FP16", "BF16"])
def test_random_graph_edge_conversions(seed, blob):
    """User-side blob precisions and layouts converted at the edge (mkldnn_infer_request.cpp:76-127): the same random networks fed U8 / FP16 /
    BF16 blobs, NCHW or NHWC, changing the layout again AFTER the first inference (the I/O format of an idle request may change)."""
    from dldt_b200 import Core
    f, inputs = random_graphs.build(seed)
    name, x = next(iter(inputs.items()))
    if blob == "U8":
        x = np.clip(np.round(x * 40 + 128), 0, 255).astype(np.uint8)
        xf = x.astype(np.float32)
    elif blob == "FP16":
        x = x.astype(np.float16)
        xf = x.astype(np.float32)
    else:      # bf16 blobs travel as their 16-bit patterns
        bits = (x.view(np.uint32) + 0x7FFF + ((x.view(np.uint32) >> 16) & 1)) >> 16
        xf = (bits.astype(np.uint32) << 16).view(np.float32)
        x = bits.astype(np.uint16)
    ref = run_oracle(f, {name: xf})
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "FP32"})
    from dldt_b200 import _capi as capi
    first = "NCHW" if seed % 4 else "NHWC"
    req = net.create_infer_request(input_precisions={name: blob}, input_layouts={name: first})
    nhwc = np.ascontiguousarray(x.transpose(0, 2, 3, 1))
    got = req.infer({name: x if first == "NCHW" else nhwc})
    for k in ref:
        assert rel_l2(got[k], ref[k]) <= 1e-4, "seed %d %s %s %s" % (seed, blob, first, k)
    # the same (idle) request, the other layout: the captured graph is dropped and recorded again
    second = "NHWC" if first == "NCHW" else "NCHW"
    capi.check(capi.lib.b200_request_set_io_format(req._h, 0, 0, {"U8": capi.U8, "FP16": capi.F16, "BF16": capi.BF16}[blob],
                                                   capi.LAYOUT_NHWC if second == "NHWC" else capi.LAYOUT_NCHW))
    req.input_blobs[name].reshape(-1)[:] = (nhwc if second == "NHWC" else x).reshape(-1)
    got = req.infer()
    for k in ref:
        assert rel_l2(got[k], ref[k]) <= 1e-4, "seed %d %s %s after %s: %s" % (seed, blob, second, first, k)


@pytest.mark.parametrize("seed", range(150))
def test_random_pooling_corner_cases_match_oracle(seed):
    """Pooling with non-square kernels / strides, asymmetric padding, ceil rounding, include / exclude padding (tests/random_graphs.py
    build_pool; the restatement is pinned to the nGraph reference kernels on 600 of these)."""
    f, inputs = random_graphs.build_pool(seed)
    _check(f, inputs, ("FP32", "BF16"), "pool seed %d" % seed)


@pytest.mark.parametrize("seed", range(240))
def test_random_convolution_geometry_matches_oracle(seed):
    """One convolution with non-square kernels (1x7, 7x1, 3x5 ...), per-axis strides / dilations, asymmetric padding, dense / grouped /
    depthwise, 1 - 320 channels (tests/random_graphs.py build_conv_corner; the restatement agrees with MKLDNNPlugin on 500 of these)."""
    f, inputs = random_graphs.build_conv_corner(seed)
    _check(f, inputs, ("FP32", "BF16", "FP16") if seed % 3 == 0 else ("FP32", "BF16"), "conv-corner seed %d" % seed)
