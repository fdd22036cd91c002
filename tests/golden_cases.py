"""Catalog of seeded parity cases shared by
  tools/make_golden.py        runs every case on the UNMODIFIED reference (oracle/_ref: MKLDNNPlugin "CPU" with
                              ENFORCE_BF16=NO, and the nGraph-reference-backed "TEMPLATE" plugin) -> tests/golden/ref_cases.npz
  tests/test_oracle_golden.py the numpy oracle against those reference outputs (CPU, no GPU)
  tests/test_gpu_golden.py    the CUDA path against the same reference outputs (-m gpu)
Each case is a function returning (Function, inputs).  Inputs are regenerated from the seed; only the
reference OUTPUTS are stored.  Shapes follow the reference's own single-layer instantiations
(inference-engine/tests/functional/plugin/cpu/shared_tests_instances/single_layer_tests/*.cpp) scaled down.
"""
import zlib

import numpy as np

from dldt_b200 import Function, zoo

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _rng(name):
    return np.random.default_rng(zlib.crc32(name.encode()))


def _f32(a):
    return np.asarray(a, np.float32)


# ---- convolution family -----------------------------------------------------------------------------
def _conv_case(name, n, c, h, w, o, k, s, p, d, act="relu", bias=True):
    r = _rng(name)
    f = Function(name)
    x = f.parameter((n, c, h, w), name="x")
    y = f.convolution(x, _f32(r.standard_normal((o, c, k, k)) / np.sqrt(c * k * k)), (s, s), (p, p), (p, p), (d, d), name="conv")
    if bias:
        y = f.add(y, _f32(r.standard_normal((1, o, 1, 1))), name="bias")
    if act == "relu":
        y = f.relu(y, name="act")
    elif act == "swish":
        y = f.swish(y, name="act")
    elif act == "clamp":
        y = f.clamp(y, 0.0, 6.0, name="act")
    f.result(y)
    return f, {"x": _f32(r.standard_normal((n, c, h, w)))}


@case
def conv_1x1(): return _conv_case("conv_1x1", 2, 64, 14, 14, 64, 1, 1, 0, 1)
@case
def conv_3x3(): return _conv_case("conv_3x3", 2, 64, 14, 14, 128, 3, 1, 1, 1)
@case
def conv_3x3_s2(): return _conv_case("conv_3x3_s2", 1, 128, 15, 13, 256, 3, 2, 1, 1)
@case
def conv_7x7_s2_c3(): return _conv_case("conv_7x7_s2_c3", 2, 3, 32, 32, 64, 7, 2, 3, 1)
@case
def conv_ragged_channels(): return _conv_case("conv_ragged_channels", 3, 24, 9, 9, 40, 3, 1, 1, 1, act="swish")
@case
def conv_dilated(): return _conv_case("conv_dilated", 1, 16, 12, 12, 8, 3, 1, 2, 2, act="clamp")
@case
def conv_1x1_s2(): return _conv_case("conv_1x1_s2", 2, 256, 7, 7, 512, 1, 2, 0, 1, act=None)
@case
def conv_wide_out(): return _conv_case("conv_wide_out", 1, 72, 10, 10, 1000, 1, 1, 0, 1, act=None, bias=False)
@case
def conv_7x7_spatial(): return _conv_case("conv_7x7_spatial", 5, 512, 7, 7, 512, 3, 1, 1, 1)


@case
def conv_bn_residual():
    r = _rng("conv_bn_residual")
    f = Function("conv_bn_residual")
    x = f.parameter((2, 64, 10, 10), name="x")
    y = f.convolution(x, _f32(r.standard_normal((64, 64, 3, 3)) / 24), (1, 1), (1, 1), (1, 1), name="conv")
    y = f.batch_norm_inference(y, r.uniform(0.5, 1.5, 64), r.uniform(-.1, .1, 64), r.uniform(-.1, .1, 64), r.uniform(0.5, 1.5, 64), 1e-5, name="bn")
    y = f.relu(f.add(y, x, name="sum"), name="out")
    f.result(y)
    return f, {"x": _f32(r.standard_normal((2, 64, 10, 10)))}


@case
def group_conv_g4():
    r = _rng("group_conv_g4")
    f = Function("group_conv_g4")
    x = f.parameter((2, 32, 9, 9), name="x")
    y = f.group_convolution(x, _f32(r.standard_normal((4, 12, 8, 3, 3)) / 8), (1, 1), (1, 1), (1, 1), name="gconv")
    f.result(f.relu(y, name="out"))
    return f, {"x": _f32(r.standard_normal((2, 32, 9, 9)))}


@case
def depthwise_s2_clamp_pools():
    r = _rng("depthwise_s2_clamp_pools")
    f = Function("depthwise")
    x = f.parameter((2, 24, 13, 13), name="x")
    y = f.group_convolution(x, _f32(r.standard_normal((24, 1, 1, 3, 3)) / 3), (2, 2), (1, 1), (1, 1), name="dw")
    y = f.clamp(f.add(y, _f32(r.standard_normal((1, 24, 1, 1))), name="bias"), 0.0, 6.0, name="relu6")
    y = f.max_pool(y, (3, 3), (2, 2), (1, 1), (1, 1), name="mp")
    y = f.avg_pool(y, (2, 2), (1, 1), (0, 0), (1, 1), exclude_pad=False, rounding="ceil", name="ap")
    f.result(y)
    return f, {"x": _f32(r.standard_normal((2, 24, 13, 13)))}


# ---- pooling ----------------------------------------------------------------------------------------
def _pool_case(name, kind, shape, k, s, pb, pe, rounding="floor", exclude_pad=False):
    r = _rng(name)
    f = Function(name)
    x = f.parameter(shape, name="x")
    if kind == "max":
        y = f.max_pool(x, k, s, pb, pe, rounding=rounding, name="pool")
    else:
        y = f.avg_pool(x, k, s, pb, pe, exclude_pad=exclude_pad, rounding=rounding, name="pool")
    f.result(y)
    return f, {"x": _f32(r.standard_normal(shape))}


@case
def maxpool_3x3_s2_p1(): return _pool_case("maxpool_3x3_s2_p1", "max", (2, 16, 15, 15), (3, 3), (2, 2), (1, 1), (1, 1))
@case
def maxpool_ceil(): return _pool_case("maxpool_ceil", "max", (1, 8, 14, 14), (3, 3), (2, 2), (0, 0), (0, 0), rounding="ceil")
@case
def avgpool_include_pad(): return _pool_case("avgpool_include_pad", "avg", (2, 8, 9, 9), (3, 3), (2, 2), (1, 1), (1, 1), exclude_pad=False)
@case
def avgpool_exclude_pad(): return _pool_case("avgpool_exclude_pad", "avg", (2, 8, 9, 9), (3, 3), (2, 2), (1, 1), (1, 1), exclude_pad=True)
@case
def avgpool_global(): return _pool_case("avgpool_global", "avg", (3, 40, 7, 7), (7, 7), (1, 1), (0, 0), (0, 0))
@case
def avgpool_ceil_nopad(): return _pool_case("avgpool_ceil_nopad", "avg", (1, 8, 10, 10), (3, 3), (2, 2), (0, 0), (0, 0), rounding="ceil", exclude_pad=True)

# ceil rounding pushes the last window PAST the padded extent (rows / cols 3..5 of a 4 x 4 input padded by 1): include-pad still divides by
# kh * kw (reference/avg_pool.hpp:215-232; MKLDNNPlugin's jit pooling).  Found by the randomized differential test (seed 43).
@case
def avgpool_ceil_include_pad_overflow():
    return _pool_case("avgpool_ceil_include_pad_overflow", "avg", (2, 24, 4, 6), (3, 3), (2, 2), (1, 1), (1, 1), rounding="ceil", exclude_pad=False)


# ---- matmul / FC ----------------------------------------------------------------------------------------
@case
def fc_bias_gelu():
    r = _rng("fc_bias_gelu")
    f = Function("fc")
    x = f.parameter((6, 64), name="x")
    y = f.matmul(x, _f32(r.standard_normal((48, 64)) / 8), transpose_b=True, name="mm")
    y = f.gelu(f.add(y, _f32(r.standard_normal((1, 48))), name="bias"), name="gelu")
    f.result(y)
    return f, {"x": _f32(r.standard_normal((6, 64)))}


@case
def fc_no_transpose_3d():
    r = _rng("fc_no_transpose_3d")
    f = Function("fc3d")
    x = f.parameter((2, 5, 32), name="x")
    y = f.matmul(x, _f32(r.standard_normal((32, 40)) / 6), name="mm")
    f.result(f.tanh(y, name="out"))
    return f, {"x": _f32(r.standard_normal((2, 5, 32)))}


@case
def matmul_batched_activations():
    r = _rng("matmul_batched_activations")
    f = Function("bmm")
    a = f.parameter((2, 3, 16, 24), name="a")
    b = f.parameter((2, 3, 16, 24), name="b")
    s = f.matmul(a, b, transpose_b=True, name="qk")          # [2,3,16,16]
    p = f.softmax(s, 3, name="sm")
    c = f.matmul(p, b, name="pv")                            # [2,3,16,24]
    f.result(c)
    return f, {"a": _f32(r.standard_normal((2, 3, 16, 24))), "b": _f32(r.standard_normal((2, 3, 16, 24)))}


# ---- softmax / mvn --------------------------------------------------------------------------------------
def _softmax_case(name, shape, axis):
    r = _rng(name)
    f = Function(name)
    x = f.parameter(shape, name="x")
    f.result(f.softmax(x, axis, name="sm"))
    return f, {"x": _f32(r.standard_normal(shape) * 3)}


@case
def softmax_last(): return _softmax_case("softmax_last", (4, 1000), 1)
@case
def softmax_mid(): return _softmax_case("softmax_mid", (2, 21, 37), 1)
@case
def softmax_ssd(): return _softmax_case("softmax_ssd", (2, 50, 9), 2)


def _mvn_case(name, shape, axes, nv, eps, mode, affine):
    r = _rng(name)
    f = Function(name)
    x = f.parameter(shape, name="x")
    y = f.mvn(x, axes, nv, eps, mode, name="mvn")
    if affine:
        y = f.multiply(y, _f32(r.uniform(0.5, 1.5, shape[-1])), name="gamma")
        y = f.add(y, _f32(r.uniform(-.5, .5, shape[-1])), name="beta")
    f.result(y)
    return f, {"x": _f32(r.standard_normal(shape) * 2 + 0.5)}


@case
def mvn_layernorm(): return _mvn_case("mvn_layernorm", (2, 7, 96), [2], True, 1e-12, "INSIDE_SQRT", True)
@case
def mvn_outside_sqrt(): return _mvn_case("mvn_outside_sqrt", (3, 5, 64), [-1], True, 1e-5, "OUTSIDE_SQRT", False)
@case
def mvn_mean_only(): return _mvn_case("mvn_mean_only", (4, 40), [1], False, 1e-9, "INSIDE_SQRT", False)


# ---- eltwise / activations ------------------------------------------------------------------------------
@case
def unary_activations():
    r = _rng("unary_activations")
    f = Function("unary")
    x = f.parameter((2, 12, 5, 5), name="x")
    outs = [f.relu(x, name="relu"), f.sigmoid(x, name="sigmoid"), f.tanh(x, name="tanh"), f.clamp(x, -0.5, 0.75, name="clamp"),
            f.swish(x, name="swish"), f.gelu(x, "ERF", name="gelu_erf"), f.gelu(x, "TANH", name="gelu_tanh"), f.hswish(x, name="hswish"),
            f.hsigmoid(x, name="hsigmoid"), f.mish(x, name="mish"), f.exp(x, name="exp"), f.abs(x, name="abs"), f.negative(x, name="neg"),
            f.erf(x, name="erf"), f.elu(x, 0.7, name="elu"), f.sqrt(f.abs(x, name="abs2"), name="sqrt")]
    for o in outs:
        f.result(o)
    return f, {"x": _f32(r.standard_normal((2, 12, 5, 5)) * 2)}


@case
def round_modes():
    # opset5 Round, both modes, with exact .5 ties in the data (what FakeQuantize decomposes to: fq_decomposition.cpp)
    r = _rng("round_modes")
    f = Function("round")
    x = f.parameter((3, 40), name="x")
    f.result(f.round(x, "half_to_even", name="even"))
    f.result(f.round(x, "half_away_from_zero", name="away"))
    v = _f32(r.standard_normal((3, 40)) * 4)
    v[0, :20] = np.arange(-10, 10, dtype=np.float32) + 0.5
    return f, {"x": v}


@case
def fake_quantize_per_tensor():
    # opset1 FakeQuantize, 256 levels, one range for the tensor; values below / above the input range saturate to out_low / out_high
    r = _rng("fake_quantize_per_tensor")
    f = Function("fq_tensor")
    x = f.parameter((2, 16, 9, 9), name="x")
    f.result(f.fake_quantize(x, _f32([-1.5]), _f32([2.0]), _f32([-1.0]), _f32([3.0]), 256, name="fq"))
    return f, {"x": _f32(r.standard_normal((2, 16, 9, 9)) * 1.5)}


@case
def fake_quantize_per_channel_conv():
    # the quantised-model pattern: Convolution -> FakeQuantize with per-channel ranges [1, C, 1, 1] -> Convolution (activations in NHWC8 here)
    r = _rng("fake_quantize_per_channel_conv")
    f = Function("fq_channel")
    x = f.parameter((2, 24, 10, 10), name="x")
    y = f.relu(f.add(f.convolution(x, _f32(r.standard_normal((40, 24, 3, 3)) / np.sqrt(216.0)), (1, 1), (1, 1), (1, 1), (1, 1), name="conv1"),
                     _f32(r.standard_normal((1, 40, 1, 1)) * 0.2), name="bias1"), name="relu1")
    hi = _f32(r.uniform(1.0, 3.0, (1, 40, 1, 1)))
    y = f.fake_quantize(y, _f32(np.zeros((1, 40, 1, 1))), hi, _f32(np.zeros((1, 40, 1, 1))), hi, 256, name="fq")
    y = f.convolution(y, _f32(r.standard_normal((16, 40, 1, 1)) / np.sqrt(40.0)), (1, 1), (0, 0), (0, 0), (1, 1), name="conv2")
    f.result(y)
    return f, {"x": _f32(r.standard_normal((2, 24, 10, 10)))}


@case
def binary_broadcast():
    r = _rng("binary_broadcast")
    f = Function("binary")
    a = f.parameter((2, 6, 4, 5), name="a")
    b = f.parameter((6, 1, 5), name="b")
    outs = [f.add(a, b, name="add"), f.subtract(a, b, name="sub"), f.multiply(a, b, name="mul"),
            f.divide(a, f.add(f.abs(b, name="babs"), _f32([1.0]), name="bden"), name="div"),
            f.maximum(a, b, name="max"), f.minimum(a, b, name="min"), f.squared_difference(a, b, name="sqd"),
            f.power(f.add(f.abs(a, name="aabs"), _f32([0.5]), name="apos"), _f32([1.5]), name="pow")]
    for o in outs:
        f.result(o)
    return f, {"a": _f32(r.standard_normal((2, 6, 4, 5))), "b": _f32(r.standard_normal((6, 1, 5)))}


# ---- data movement ---------------------------------------------------------------------------------------
@case
def concat_split_transpose():
    r = _rng("concat_split_transpose")
    f = Function("move")
    x = f.parameter((2, 16, 6, 6), name="x")
    y = f.parameter((2, 8, 6, 6), name="y")
    c = f.concat([x, y], 1, name="cat")                      # [2,24,6,6]
    t = f.transpose(c, (0, 2, 3, 1), name="perm")            # NHWC
    fl = f.reshape(t, (2, -1), name="flat")
    s0, s1 = f.split(fl, 1, 2, name="split")
    f.result(f.relu(s0, name="o0"))
    f.result(f.negative(s1, name="o1"))
    f.result(f.transpose(x, (1, 0, 3, 2), name="perm2"))
    return f, {"x": _f32(r.standard_normal((2, 16, 6, 6))), "y": _f32(r.standard_normal((2, 8, 6, 6)))}


@case
def gather_embedding():
    r = _rng("gather_embedding")
    f = Function("gather")
    ids = f.parameter((3, 9), np.int32, name="ids")
    e = f.gather(_f32(r.standard_normal((50, 32))), ids, 0, name="emb")
    f.result(f.add(e, _f32(r.standard_normal((1, 9, 32))), name="out"))
    return f, {"ids": r.integers(0, 50, (3, 9)).astype(np.int32)}


@case
def reduce_mean_spatial():
    r = _rng("reduce_mean_spatial")
    f = Function("reduce")
    x = f.parameter((2, 24, 5, 7), name="x")
    c = f.relu(f.convolution(x, _f32(r.standard_normal((16, 24, 1, 1)) / 5), name="conv"), name="act")
    f.result(f.reduce_mean(c, [2, 3], keep_dims=False, name="gap"))
    f.result(f.reduce_max(x, [3], keep_dims=True, name="rmax"))
    f.result(f.reduce_sum(x, [1, 2], keep_dims=False, name="rsum"))
    return f, {"x": _f32(r.standard_normal((2, 24, 5, 7)))}


# ---- TopK / NMS / DetectionOutput (index outputs must be bit-exact) -------------------------------------
def _topk_case(name, shape, k, axis, mode, sort, ties):
    r = _rng(name)
    f = Function(name)
    x = f.parameter(shape, name="x")
    v, i = f.topk(x, k, axis, mode, sort, name="topk")
    f.result(v)
    f.result(i)
    data = r.integers(-6, 6, shape).astype(np.float32) if ties else _f32(r.standard_normal(shape))
    return f, {"x": data}


@case
def topk_imagenet_top5(): return _topk_case("topk_imagenet_top5", (8, 1000), 5, 1, "max", "value", False)
@case
def topk_ties_max_value(): return _topk_case("topk_ties_max_value", (6, 40), 7, 1, "max", "value", True)
@case
def topk_ties_min_index(): return _topk_case("topk_ties_min_index", (3, 17, 5), 4, 1, "min", "index", True)
@case
def topk_axis0(): return _topk_case("topk_axis0", (12, 9), 3, 0, "max", "value", True)


def _boxes(r, b, n, center):
    xy = r.uniform(0, 0.8, (b, n, 2))
    wh = r.uniform(0.05, 0.4, (b, n, 2))
    if center:
        return _f32(np.concatenate([xy + wh / 2, wh], axis=2))
    return _f32(np.concatenate([xy, xy + wh], axis=2))


def _nms_case(name, b, c, n, max_out, iou, score_thr, center=False, sort=True, ties=False, soft=0.0):
    r = _rng(name)
    f = Function(name)
    boxes = f.parameter((b, n, 4), name="boxes")
    scores = f.parameter((b, c, n), name="scores")
    i, s, v = f.non_max_suppression(boxes, scores, max_out, iou, score_thr, soft, "center" if center else "corner", sort, name="nms")
    f.result(i)
    f.result(s)
    f.result(v)
    sc = r.uniform(0, 1, (b, c, n))
    if ties:
        sc = np.round(sc * 8) / 8
    return f, {"boxes": _boxes(r, b, n, center), "scores": _f32(sc)}


@case
def nms_corner_sorted(): return _nms_case("nms_corner_sorted", 2, 3, 60, 10, 0.5, 0.2)
@case
def nms_center_unsorted(): return _nms_case("nms_center_unsorted", 1, 2, 45, 6, 0.4, 0.1, center=True, sort=False)
@case
def nms_score_ties(): return _nms_case("nms_score_ties", 2, 2, 50, 8, 0.5, 0.25, ties=True)
@case
def nms_keep_nothing(): return _nms_case("nms_keep_nothing", 1, 1, 20, 5, 0.5, 2.0)
@case
def nms_soft_sigma(): return _nms_case("nms_soft_sigma", 2, 2, 40, 12, 0.6, 0.15, soft=0.5)
@case
def nms_soft_center_unsorted(): return _nms_case("nms_soft_center_unsorted", 1, 3, 30, 8, 0.5, 0.1, center=True, sort=False, soft=0.25)


def _detout_case(name, n, priors, classes, top_k, keep_top_k, center=True, share=True, var_in_target=False, clip=False, conf_thr=0.05):
    r = _rng(name)
    f = Function(name)
    nloc = 1 if share else classes
    loc = f.parameter((n, priors * nloc * 4), name="loc")
    conf = f.parameter((n, priors * classes), name="conf")
    pb = _boxes(r, 1, priors, False)[0]
    var = np.tile(_f32([0.1, 0.1, 0.2, 0.2]), (priors, 1))
    pri = pb.reshape(1, 1, -1) if var_in_target else np.stack([pb.reshape(-1), var.reshape(-1)])[None]
    det = f.detection_output(loc, conf, _f32(pri), classes, 0, top_k, keep_top_k,
                             "caffe.PriorBoxParameter.CENTER_SIZE" if center else "caffe.PriorBoxParameter.CORNER", share, 0.45, conf_thr,
                             var_in_target, True, clip, False, name="det")
    f.result(det)
    logits = r.standard_normal((n, priors, classes)) * 2
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    return f, {"loc": _f32(r.standard_normal((n, priors * nloc * 4)) * 0.5), "conf": _f32(p.reshape(n, -1))}


@case
def detout_ssd_like(): return _detout_case("detout_ssd_like", 2, 120, 6, 20, 15)
@case
def detout_corner_clip(): return _detout_case("detout_corner_clip", 1, 80, 4, 10, 50, center=False, clip=True)
@case
def detout_not_shared(): return _detout_case("detout_not_shared", 2, 40, 3, 8, 6, share=False)
@case
def detout_var_in_target(): return _detout_case("detout_var_in_target", 1, 64, 5, -1, 12, var_in_target=True)


@case
def prelu_per_channel_slope():
    """PRelu with the slope Model Optimizer emits: rank 1, one value per channel ([C], not [1,C,1,1]); C == W on purpose, so a plain numpy
    broadcast would silently scale along W.  Also after a convolution (the conv output stays NHWC on the device)."""
    r = _rng("prelu_per_channel_slope")
    f = Function("prelu_per_channel_slope")
    x = f.parameter((2, 8, 5, 8), name="x")
    f.result(f.prelu(x, _f32(r.uniform(0.05, 0.5, 8)), name="prelu"))
    y = f.convolution(x, _f32(r.standard_normal((16, 8, 3, 3)) / np.sqrt(72)), (1, 1), (1, 1), (1, 1), name="conv")
    f.result(f.prelu(y, _f32(r.uniform(0.05, 0.5, 16)), name="conv_prelu"))
    f.result(f.prelu(x, _f32([0.25]), name="prelu_scalar"))
    return f, {"x": _f32(r.standard_normal((2, 8, 5, 8)))}


# ---- whole (tiny) networks of the four model families -----------------------------------------------------
@case
def net_tiny_resnet():
    f = zoo.resnet50(batch=2, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
    return f, {"data": _f32(_rng("net_tiny_resnet").uniform(0, 1, (2, 3, 64, 64)))}


@case
def net_tiny_mobilenet():
    f = zoo.mobilenet_v2(batch=2, image=64, width_mult=0.25, num_classes=16, settings=[(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 2, 2), (6, 64, 1, 1)])
    return f, {"data": _f32(_rng("net_tiny_mobilenet").uniform(0, 1, (2, 3, 64, 64)))}


@case
def net_tiny_bert():
    r = _rng("net_tiny_bert")
    f = zoo.bert_base(batch=2, seq=32, hidden=64, heads=4, layers=2, ffn=128, vocab=100, max_pos=64)
    ins = {"input_ids": r.integers(0, 100, (2, 32)).astype(np.int32), "attention_mask": np.ones((2, 32), np.int32),
           "token_type_ids": r.integers(0, 2, (2, 32)).astype(np.int32)}
    ins["attention_mask"][1, 20:] = 0
    return f, ins


@case
def net_tiny_ssd():
    f = zoo.ssd300_resnet34(batch=2, image=96, width=16, num_classes=6, trunk_blocks=(1, 1, 1), keep_top_k=20, top_k=10)
    return f, {"data": _f32(_rng("net_tiny_ssd").uniform(0, 1, (2, 3, 96, 96)))}


# outputs holding indices / ids that must match bit-exactly (north star: "TopK/NMS indices must be bit-exact in fp32 mode")
def is_index_output(arr):
    return np.asarray(arr).dtype.kind in "iu"
