"""ORACLE - TEST INFRASTRUCTURE ONLY (see oracle/np_ops.py header; parity status: PINNED there).

Evaluates a recorded `dldt_b200.graph.Function` node by node with the numpy restatements in
np_ops.py - the role nGraph's INTERPRETER backend plays for the reference's own functional tests
(inference-engine/tests/functional/shared_test_classes/src/base/layer_test_utils.cpp:351-404).
"""
import numpy as np

from . import np_ops as ops

_UNARY = {"Relu": "Relu", "Sigmoid": "Sigmoid", "Tanh": "Tanh", "HSwish": "HSwish", "HSigmoid": "HSigmoid", "Mish": "Mish",
          "Exp": "Exp", "Sqrt": "Sqrt", "Abs": "Abs", "Negative": "Negative", "Erf": "Erf"}
_BINARY = {"Add", "Subtract", "Multiply", "Divide", "Maximum", "Minimum", "Power", "SquaredDifference", "PRelu"}


def run(func, inputs, return_all=False):
    """inputs: dict parameter-name -> ndarray.  Returns dict output-name -> ndarray (fp32 / int32)."""
    env = {}
    for n in func.nodes:
        a = n.attrs
        x = [env[i.vid] for i in n.inputs]
        if n.op == "Parameter":
            v = np.asarray(inputs[n.name])
            assert tuple(v.shape) == tuple(n.outputs[0].shape), (n.name, v.shape, n.outputs[0].shape)
            out = [v.astype(np.int32) if v.dtype.kind in "iu" and v.dtype != np.uint8 else v.astype(np.float32)]
        elif n.op == "Constant":
            out = [n.data.astype(np.int32) if n.data.dtype.kind in "iu" else n.data.astype(np.float32)]
        elif n.op == "Result":
            out = [x[0]]
        elif n.op == "Convolution":
            out = [ops.convolution(x[0], x[1], a["strides"], a["pads_begin"], a["pads_end"], a["dilations"])]
        elif n.op == "GroupConvolution":
            out = [ops.group_convolution(x[0], x[1], a["strides"], a["pads_begin"], a["pads_end"], a["dilations"])]
        elif n.op == "BatchNormInference":
            out = [ops.batch_norm_inference(x[0], x[1], x[2], x[3], x[4], a["epsilon"])]
        elif n.op in _BINARY:
            out = [ops.binary(x[0], x[1], n.op)]
        elif n.op in _UNARY:
            out = [ops.activation(x[0], _UNARY[n.op])]
        elif n.op == "Clamp":
            out = [ops.activation(x[0], "Clamp", a["min"], a["max"])]
        elif n.op == "Swish":
            beta = float(n.ir_consts[0]) if n.ir_consts else 1.0
            out = [ops.activation(x[0], "Swish", beta)]
        elif n.op == "Gelu":
            out = [ops.activation(x[0], "GeluErf" if a["approximation_mode"] == "ERF" else "GeluTanh")]
        elif n.op == "Elu":
            out = [ops.activation(x[0], "Elu", a["alpha"])]
        elif n.op == "Round":
            out = [ops.activation(x[0], "RoundEven" if a["mode"] == "half_to_even" else "RoundAway")]
        elif n.op == "FakeQuantize":
            out = [ops.fake_quantize(x[0], x[1], x[2], x[3], x[4], a["levels"])]
        elif n.op == "MaxPool":
            out = [ops.max_pool(x[0], a["kernel"], a["strides"], a["pads_begin"], a["pads_end"], a["rounding_type"])]
        elif n.op == "AvgPool":
            out = [ops.avg_pool(x[0], a["kernel"], a["strides"], a["pads_begin"], a["pads_end"], a["exclude-pad"] == "true", a["rounding_type"])]
        elif n.op in ("ReduceMean", "ReduceSum", "ReduceMax"):
            out = [ops.reduce(x[0], [int(v) for v in n.ir_consts[0]], a["keep_dims"] == "true", n.op)]
        elif n.op == "MatMul":
            out = [ops.matmul(x[0], x[1], a["transpose_a"] == "true", a["transpose_b"] == "true")]
        elif n.op == "Softmax":
            out = [ops.softmax(x[0], a["axis"])]
        elif n.op == "MVN":
            out = [ops.mvn(x[0], [int(v) for v in n.ir_consts[0]], a["normalize_variance"] == "true", a["eps"], a["eps_mode"])]
        elif n.op == "Reshape":
            out = [x[0].reshape(n.outputs[0].shape)]
        elif n.op == "Transpose":
            out = [np.ascontiguousarray(np.transpose(x[0], [int(v) for v in n.ir_consts[0]]))]
        elif n.op == "StridedSlice":
            # x[b:e:s, ...] with the masks of opset1 StridedSlice (bit set = the value is ignored); ngraph reference/strided_slice.hpp
            b, e, st = ([int(v) for v in c] for c in n.ir_consts)
            sl = tuple(slice(None if a["begin_mask"][d] else b[d], None if a["end_mask"][d] else e[d], st[d]) for d in range(len(b)))
            out = [np.ascontiguousarray(x[0][sl])]
        elif n.op == "Concat":
            out = [np.concatenate(x, axis=a["axis"])]
        elif n.op == "Split":
            out = list(np.split(x[0], a["num_splits"], axis=int(n.ir_consts[0])))
        elif n.op == "Gather":
            out = [ops.gather(x[0], x[1], int(n.ir_consts[0]))]
        elif n.op == "Convert":
            out = [x[0].astype(np.int32) if a["destination_type"] in ("i32", "i64") else x[0].astype(np.float32)]
        elif n.op == "TopK":
            v, i = ops.topk(x[0], int(n.ir_consts[0]), a["axis"], a["mode"], a["sort"])
            out = [v, i]
        elif n.op == "NonMaxSuppression":
            c = n.ir_consts
            out = list(ops.non_max_suppression(x[0], x[1], int(c[0]), float(c[1]), float(c[2]), a["box_encoding"] == "center",
                                               a["sort_result_descending"] == "true", float(c[3]) if len(c) > 3 else 0.0))
        elif n.op == "DetectionOutput":
            t = lambda k: a[k] == "true"
            out = [ops.detection_output(x[0], x[1], x[2], a["num_classes"], a["background_label_id"], a["top_k"], a["keep_top_k"][0],
                                        a["code_type"].endswith("CENTER_SIZE"), t("share_location"), a["nms_threshold"],
                                        a["confidence_threshold"], t("variance_encoded_in_target"), t("normalized"), t("clip_before_nms"),
                                        t("clip_after_nms"), a["input_height"], a["input_width"])]
        else:
            raise NotImplementedError(n.op)
        for o, v in zip(n.outputs, out):
            env[o.vid] = v
    res = {}
    for r in func.results:
        src = r.inputs[0]
        name = src.node.name if len(src.node.outputs) == 1 else "%s.%d" % (src.node.name, src.index)
        res[name] = env[src.vid]
    if return_all:
        return res, {n.name: env[n.outputs[0].vid] for n in func.nodes if n.outputs}
    return res
