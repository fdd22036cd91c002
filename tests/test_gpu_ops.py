"""Parity of the CUDA path (through the C ABI) against the numpy oracle on seeded inputs.
fp32 mode: rel-L2 <= 1e-4; bf16/fp16 mode: <= 1e-2."""
import numpy as np
import pytest

from dldt_b200 import Function, zoo
from tests.util import rel_l2, run_b200, run_oracle

pytestmark = pytest.mark.gpu
TOL = {"FP32": 1e-4, "BF16": 1e-2, "FP16": 3e-3}


def _check(func, inputs, precision, tol=None):
    got, net = run_b200(func, inputs, precision)
    ref = run_oracle(func, inputs)
    assert set(got) == set(ref)
    for k in ref:
        if ref[k].dtype.kind in "iu":
            np.testing.assert_array_equal(got[k], ref[k])
        else:
            err = rel_l2(got[k], ref[k])
            assert err <= (tol or TOL[precision]), "%s: rel-L2 %.3e (%s)" % (k, err, precision)
    return net


CONV_CASES = [
    # N, C, H, W, O, k, stride, pad, dil
    (2, 64, 14, 14, 64, 1, 1, 0, 1),
    (2, 64, 14, 14, 128, 3, 1, 1, 1),
    (1, 128, 15, 13, 256, 3, 2, 1, 1),
    (2, 3, 32, 32, 64, 7, 2, 3, 1),
    (3, 24, 9, 9, 40, 3, 1, 1, 1),
    (1, 16, 12, 12, 8, 3, 1, 2, 2),
    (2, 256, 7, 7, 512, 1, 2, 0, 1),
    (1, 72, 10, 10, 1000, 1, 1, 0, 1),
    (5, 512, 7, 7, 512, 3, 1, 1, 1),
]


@pytest.mark.parametrize("precision", ["FP32", "BF16", "FP16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_bias_relu(case, precision):
    n, c, h, w, o, k, s, p, d = case
    rng = np.random.default_rng(hash(case) % 2**31)
    f = Function("conv")
    x = f.parameter((n, c, h, w))
    y = f.convolution(x, (rng.standard_normal((o, c, k, k)) / np.sqrt(c * k * k)).astype(np.float32), (s, s), (p, p), (p, p), (d, d))
    y = f.relu(f.add(y, rng.standard_normal((1, o, 1, 1)).astype(np.float32)))
    f.result(y)
    net = _check(f, {"Parameter_0": rng.standard_normal((n, c, h, w)).astype(np.float32)}, precision)
    eg = net.get_exec_graph_info()
    assert len(eg) == 1 and eg[0]["type"] == "Convolution" and len(eg[0]["fused"]) == 3


@pytest.mark.parametrize("precision", ["FP32", "BF16"])
def test_conv_bn_residual(precision):
    rng = np.random.default_rng(7)
    f = Function("res")
    x = f.parameter((2, 64, 10, 10))
    w = (rng.standard_normal((64, 64, 3, 3)) / 24).astype(np.float32)
    y = f.convolution(x, w, (1, 1), (1, 1), (1, 1))
    y = f.batch_norm_inference(y, rng.uniform(0.5, 1.5, 64), rng.uniform(-.1, .1, 64), rng.uniform(-.1, .1, 64), rng.uniform(0.5, 1.5, 64), 1e-5)
    y = f.relu(f.add(y, x))
    f.result(y)
    net = _check(f, {"Parameter_0": rng.standard_normal((2, 64, 10, 10)).astype(np.float32)}, precision)
    assert len(net.get_exec_graph_info()) == 1


@pytest.mark.parametrize("precision", ["FP32", "BF16"])
def test_depthwise_pool_clamp(precision):
    rng = np.random.default_rng(8)
    f = Function("dw")
    x = f.parameter((2, 24, 13, 13))
    y = f.group_convolution(x, rng.standard_normal((24, 1, 1, 3, 3)).astype(np.float32) / 3, (2, 2), (1, 1), (1, 1))
    y = f.clamp(f.add(y, rng.standard_normal((1, 24, 1, 1)).astype(np.float32)), 0.0, 6.0)
    y = f.max_pool(y, (3, 3), (2, 2), (1, 1), (1, 1))
    y = f.avg_pool(y, (2, 2), (1, 1), (0, 0), (1, 1), exclude_pad=False, rounding="ceil")
    f.result(y)
    _check(f, {"Parameter_0": rng.standard_normal((2, 24, 13, 13)).astype(np.float32)}, precision)


@pytest.mark.parametrize("precision", ["FP32", "BF16"])
def test_fc_softmax_topk(precision):
    rng = np.random.default_rng(9)
    f = Function("fc")
    x = f.parameter((5, 64))
    y = f.matmul(x, rng.standard_normal((40, 64)).astype(np.float32) / 8, transpose_b=True)
    y = f.add(y, rng.standard_normal((1, 40)).astype(np.float32))
    p = f.softmax(y, 1)
    f.result(p)
    if precision == "FP32":
        v, i = f.topk(p, 5, 1)
        f.result(v)
        f.result(i)
    _check(f, {"Parameter_0": rng.standard_normal((5, 64)).astype(np.float32)}, precision)


@pytest.mark.parametrize("precision", ["FP32", "BF16", "FP16"])
@pytest.mark.parametrize("m,k,n", [(1, 84, 10), (7, 30, 16), (33, 100, 72)])
def test_fc_k_not_multiple_of_8(m, k, n, precision):
    """FullyConnected whose reduction length is not a multiple of 8 (LeNet's 84 -> 10; found by the reference's behaviour suite,
    core_integration.hpp:578-624): the activation rows are zero-padded to the 16-byte granule of the kernels first."""
    rng = np.random.default_rng(m * 1000 + k)
    f = Function("fc_k")
    x = f.parameter((m, k))
    y = f.matmul(f.relu(x), rng.standard_normal((n, k)).astype(np.float32) / np.sqrt(k), transpose_b=True)
    y = f.add(y, rng.standard_normal((1, n)).astype(np.float32))
    f.result(y)
    net = _check(f, {"Parameter_0": rng.standard_normal((m, k)).astype(np.float32)}, precision)
    assert any(e["impl"] == "copy2d" for e in net.get_exec_graph_info())


@pytest.mark.parametrize("precision", ["FP32", "BF16"])
def test_tiny_resnet(precision):
    f = zoo.resnet50(batch=2, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
    rng = np.random.default_rng(11)
    _check(f, {"data": rng.uniform(0, 1, (2, 3, 64, 64)).astype(np.float32)}, precision, tol=None if precision == "FP32" else 1e-2)


@pytest.mark.parametrize("precision", ["FP32", "BF16"])
def test_tiny_mobilenet(precision):
    f = zoo.mobilenet_v2(batch=2, image=64, width_mult=0.25, num_classes=16,
                         settings=[(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 2, 2), (6, 64, 1, 1)])
    rng = np.random.default_rng(12)
    _check(f, {"data": rng.uniform(0, 1, (2, 3, 64, 64)).astype(np.float32)}, precision, tol=None if precision == "FP32" else 1e-2)


@pytest.mark.parametrize("precision", ["FP32", "BF16"])
def test_tiny_bert(precision):
    f = zoo.bert_base(batch=2, seq=32, hidden=64, heads=4, layers=2, ffn=128, vocab=100, max_pos=64)
    rng = np.random.default_rng(13)
    ins = {"input_ids": rng.integers(0, 100, (2, 32)).astype(np.int32), "attention_mask": np.ones((2, 32), np.int32),
           "token_type_ids": rng.integers(0, 2, (2, 32)).astype(np.int32)}
    ins["attention_mask"][1, 20:] = 0
    _check(f, ins, precision, tol=None if precision == "FP32" else 1e-2)


def test_tiny_ssd_fp32():
    f = zoo.ssd300_resnet34(batch=2, image=96, width=16, num_classes=6, trunk_blocks=(1, 1, 1), keep_top_k=20, top_k=10)
    rng = np.random.default_rng(14)
    got, _ = run_b200(f, {"data": rng.uniform(0, 1, (2, 3, 96, 96)).astype(np.float32)}, "FP32")
    ref = run_oracle(f, {"data": rng.uniform(0, 1, (2, 3, 96, 96)).astype(np.float32)})
    # recompute with the same input
    x = np.random.default_rng(14).uniform(0, 1, (2, 3, 96, 96)).astype(np.float32)
    got, _ = run_b200(f, {"data": x}, "FP32")
    ref = run_oracle(f, {"data": x})
    g, r = got["detection_out"].reshape(-1, 7), ref["detection_out"].reshape(-1, 7)
    np.testing.assert_array_equal(g[:, :2], r[:, :2])          # image ids + labels (and the -1 terminator)
    assert rel_l2(g[:, 2:], r[:, 2:]) < 1e-4


@pytest.mark.parametrize("priors,classes,top_k,keep_top_k,n", [(8732, 21, 200, 200, 2), (3000, 40, 400, 100, 1), (700, 5, 64, -1, 3)])
def test_detection_output_select_path_equals_sort_path(priors, classes, top_k, keep_top_k, n, monkeypatch):
    """The radix-select + overlap-bitmask DetectionOutput (SSD sizes) against the full-sort fall-back of the same library and the
    numpy restatement, with heavily quantised scores so that equal scores are ordered by prior index on every path."""
    from dldt_b200 import Function
    rng = np.random.default_rng(priors + classes)
    f = Function("detout_big")
    loc = f.parameter((n, priors * 4), name="loc")
    conf = f.parameter((n, priors * classes), name="conf")
    cx, cy = rng.uniform(0.1, 0.9, priors), rng.uniform(0.1, 0.9, priors)
    w, h = rng.uniform(0.05, 0.3, priors), rng.uniform(0.05, 0.3, priors)
    pb = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
    var = np.tile(np.asarray([0.1, 0.1, 0.2, 0.2], np.float32), (priors, 1))
    pri = np.stack([pb.reshape(-1), var.reshape(-1)])[None]
    det = f.detection_output(loc, conf, pri, classes, 0, top_k, keep_top_k, "caffe.PriorBoxParameter.CENTER_SIZE", True, 0.45, 0.02, name="det")
    f.result(det)
    logits = rng.standard_normal((n, priors, classes)) * 2
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    p = np.round(p * 64) / 64            # many exact ties
    ins = {"loc": (rng.standard_normal((n, priors * 4)) * 0.5).astype(np.float32), "conf": p.reshape(n, -1).astype(np.float32)}
    fast, _ = run_b200(f, ins, "FP32")
    monkeypatch.setenv("B200_DETOUT_FAST", "0")
    slow, _ = run_b200(f, ins, "FP32")
    monkeypatch.delenv("B200_DETOUT_FAST")
    np.testing.assert_array_equal(fast["det"], slow["det"])
    ref = run_oracle(f, ins)["det"].reshape(-1, 7)
    g = fast["det"].reshape(-1, 7)
    np.testing.assert_array_equal(g[:, :2], ref[:, :2])
    assert rel_l2(g[:, 2:], ref[:, 2:]) < 1e-5
