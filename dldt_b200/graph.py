"""Host-side mirror of the nGraph function builder for the ops on the hot path.

`Function` records an op graph with nGraph opset semantics (names, attribute meaning and input
order follow ngraph/core/include/ngraph/op/*.hpp of the reference) and feeds it node by node into
the C ABI (`b200_graph_add`), which does shape inference.  The same recorded graph can be
  * compiled for the GPU (`dldt_b200.ie.Core().load_network(func, "B200")`),
  * serialised to IR v10 for the reference `benchmark_app` / CPU plugin (`dldt_b200.ir.serialize`),
  * interpreted by the numpy oracle (`oracle/interp.py`, tests only).
"""
import ctypes as C

import numpy as np

from . import _capi as capi

_NP2DT = {np.dtype(np.float32): capi.F32, np.dtype(np.float16): capi.F16, np.dtype(np.int32): capi.I32,
          np.dtype(np.uint8): capi.U8, np.dtype(np.int64): capi.I64, np.dtype(np.int8): capi.I8, np.dtype(np.bool_): capi.BOOL}
DT_NAME = {capi.F32: "f32", capi.F16: "f16", capi.BF16: "bf16", capi.I32: "i32", capi.U8: "u8", capi.I64: "i64",
           capi.I8: "i8", capi.BOOL: "boolean"}
DT_NP = {capi.F32: np.float32, capi.F16: np.float16, capi.I32: np.int32, capi.U8: np.uint8, capi.I64: np.int64,
         capi.I8: np.int8, capi.BOOL: np.bool_}


class Output:
    """One output port of a node (ngraph::Output<Node>)."""

    def __init__(self, node, index, vid, shape, dtype):
        self.node, self.index, self.vid, self.shape, self.dtype = node, index, vid, tuple(shape), dtype

    def __repr__(self):
        return "<%s.%d %s %s>" % (self.node.name, self.index, DT_NAME[self.dtype], list(self.shape))


class Node:
    def __init__(self, op, name, inputs, attrs, version="opset1"):
        self.op, self.name, self.inputs, self.attrs, self.version = op, name, list(inputs), dict(attrs), version
        self.outputs = []
        self.data = None      # Constant payload (numpy)
        self.ir_consts = []   # extra constant inputs the IR form of this op carries (shape, axes, k ...)


class Function:
    def __init__(self, name="network"):
        self.name = name
        self.nodes, self.parameters, self.results = [], [], []
        self._h = C.c_void_p()
        capi.check(capi.lib.b200_graph_create(name.encode(), C.byref(self._h)))
        self._count = {}

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None and capi.lib is not None:
            capi.lib.b200_graph_destroy(self._h)
            self._h = None

    # ---- plumbing ---------------------------------------------------------------------------
    def _uname(self, op, name):
        if name:
            return name
        k = self._count.get(op, 0)
        self._count[op] = k + 1
        return "%s_%d" % (op, k)

    def _add(self, op, inputs, name=None, i=(), f=(), dtype=0, shape=(), data=None, attrs=None, version="opset1", ir_consts=()):
        node = Node(op, self._uname(op, name), inputs, attrs or {}, version)
        node.ir_consts = list(ir_consts)
        d = capi.NodeDesc()
        d.op = capi.OP_ID[op]
        d.name = node.name.encode()
        d.n_inputs = len(inputs)
        for k, x in enumerate(inputs):
            d.inputs[k] = x.vid
        for k, v in enumerate(i):
            d.i[k] = int(v)
        for k, v in enumerate(f):
            d.f[k] = float(v)
        d.dtype = dtype
        d.rank = len(shape)
        for k, v in enumerate(shape):
            d.shape[k] = int(v)
        if data is not None:
            node.data = np.ascontiguousarray(data)
            d.data = node.data.ctypes.data_as(C.c_void_p)
            d.data_bytes = node.data.nbytes
        out_ids = (C.c_int32 * capi.MAX_INPUTS)()
        n_out = C.c_int32(0)
        capi.check(capi.lib.b200_graph_add(self._h, C.byref(d), out_ids, C.byref(n_out)))
        info = capi.TensorInfo()
        for k in range(n_out.value):
            capi.check(capi.lib.b200_graph_value_info(self._h, out_ids[k], C.byref(info)))
            node.outputs.append(Output(node, k, out_ids[k], [info.shape[r] for r in range(info.rank)], info.dtype))
        self.nodes.append(node)
        return node

    # ---- sources / sinks ----------------------------------------------------------------------
    def parameter(self, shape, dtype=np.float32, name=None):
        dt = _NP2DT[np.dtype(dtype)]
        n = self._add("Parameter", [], name, dtype=dt, shape=shape, attrs={"shape": shape, "element_type": DT_NAME[dt]})
        self.parameters.append(n)
        return n.outputs[0]

    def constant(self, value, dtype=None, name=None):
        arr = np.asarray(value)
        if dtype is not None:
            arr = arr.astype(dtype)
        if arr.dtype == np.float64:
            arr = arr.astype(np.float32)
        dt = _NP2DT[arr.dtype]
        n = self._add("Constant", [], name, dtype=dt, shape=arr.shape, data=arr, attrs={"element_type": DT_NAME[dt], "shape": arr.shape})
        return n.outputs[0]

    def result(self, x, name=None):
        n = self._add("Result", [x], name)
        self.results.append(n)
        return n

    # ---- convolutions -------------------------------------------------------------------------
    def convolution(self, x, w, strides=(1, 1), pads_begin=(0, 0), pads_end=(0, 0), dilations=(1, 1), name=None):
        w = w if isinstance(w, Output) else self.constant(w)
        attrs = {"strides": strides, "dilations": dilations, "pads_begin": pads_begin, "pads_end": pads_end, "auto_pad": "explicit"}
        i = [strides[0], strides[1], dilations[0], dilations[1], pads_begin[0], pads_begin[1], pads_end[0], pads_end[1]]
        return self._add("Convolution", [x, w], name, i=i, attrs=attrs).outputs[0]

    def group_convolution(self, x, w, strides=(1, 1), pads_begin=(0, 0), pads_end=(0, 0), dilations=(1, 1), name=None):
        w = w if isinstance(w, Output) else self.constant(w)
        attrs = {"strides": strides, "dilations": dilations, "pads_begin": pads_begin, "pads_end": pads_end, "auto_pad": "explicit"}
        i = [strides[0], strides[1], dilations[0], dilations[1], pads_begin[0], pads_begin[1], pads_end[0], pads_end[1]]
        return self._add("GroupConvolution", [x, w], name, i=i, attrs=attrs).outputs[0]

    def batch_norm_inference(self, x, gamma, beta, mean, variance, epsilon=1e-5, name=None):
        ins = [x] + [v if isinstance(v, Output) else self.constant(np.asarray(v, np.float32)) for v in (gamma, beta, mean, variance)]
        return self._add("BatchNormInference", ins, name, f=[epsilon], attrs={"epsilon": epsilon}, version="opset5").outputs[0]

    # ---- elementwise --------------------------------------------------------------------------
    def _binary(self, op, a, b, name):
        a = a if isinstance(a, Output) else self.constant(np.asarray(a, np.float32))
        b = b if isinstance(b, Output) else self.constant(np.asarray(b, np.float32))
        return self._add(op, [a, b], name, attrs={"auto_broadcast": "numpy"}).outputs[0]

    def add(self, a, b, name=None): return self._binary("Add", a, b, name)
    def subtract(self, a, b, name=None): return self._binary("Subtract", a, b, name)
    def multiply(self, a, b, name=None): return self._binary("Multiply", a, b, name)
    def divide(self, a, b, name=None): return self._binary("Divide", a, b, name)
    def maximum(self, a, b, name=None): return self._binary("Maximum", a, b, name)
    def minimum(self, a, b, name=None): return self._binary("Minimum", a, b, name)
    def power(self, a, b, name=None): return self._binary("Power", a, b, name)
    def squared_difference(self, a, b, name=None): return self._binary("SquaredDifference", a, b, name)

    def prelu(self, x, slope, name=None):
        """opset1 PRelu; a rank-1 slope of size C is per-channel (what Model Optimizer emits)."""
        slope = slope if isinstance(slope, Output) else self.constant(np.asarray(slope, np.float32))
        return self._add("PRelu", [x, slope], name).outputs[0]

    def relu(self, x, name=None): return self._add("Relu", [x], name).outputs[0]
    def sigmoid(self, x, name=None): return self._add("Sigmoid", [x], name).outputs[0]
    def tanh(self, x, name=None): return self._add("Tanh", [x], name).outputs[0]
    def exp(self, x, name=None): return self._add("Exp", [x], name).outputs[0]
    def sqrt(self, x, name=None): return self._add("Sqrt", [x], name).outputs[0]
    def abs(self, x, name=None): return self._add("Abs", [x], name).outputs[0]
    def negative(self, x, name=None): return self._add("Negative", [x], name).outputs[0]
    def erf(self, x, name=None): return self._add("Erf", [x], name).outputs[0]
    def hswish(self, x, name=None): return self._add("HSwish", [x], name, version="opset4").outputs[0]
    def hsigmoid(self, x, name=None): return self._add("HSigmoid", [x], name, version="opset5").outputs[0]
    def mish(self, x, name=None): return self._add("Mish", [x], name, version="opset4").outputs[0]

    def clamp(self, x, lo, hi, name=None):
        return self._add("Clamp", [x], name, f=[lo, hi], attrs={"min": lo, "max": hi}).outputs[0]

    def swish(self, x, beta=1.0, name=None):
        return self._add("Swish", [x], name, f=[beta], attrs={}, version="opset4",
                         ir_consts=[] if beta == 1.0 else [np.asarray(beta, np.float32)]).outputs[0]

    def gelu(self, x, approximation="ERF", name=None):
        return self._add("Gelu", [x], name, i=[0 if approximation.upper() == "ERF" else 1],
                         attrs={"approximation_mode": approximation.upper()}, version="opset7").outputs[0]

    def round(self, x, mode="half_to_even", name=None):
        return self._add("Round", [x], name, i=[0 if mode == "half_to_even" else 1], attrs={"mode": mode}, version="opset5").outputs[0]

    def fake_quantize(self, x, in_low, in_high, out_low, out_high, levels, name=None):
        """opset1 FakeQuantize; the four ranges are constants holding one value or one value per channel (shape [1, C, 1, 1] or [C, 1, 1])."""
        r = [v if isinstance(v, Output) else self.constant(np.asarray(v, np.float32)) for v in (in_low, in_high, out_low, out_high)]
        return self._add("FakeQuantize", [x] + r, name, i=[int(levels)], attrs={"levels": int(levels), "auto_broadcast": "numpy"}).outputs[0]

    def elu(self, x, alpha=1.0, name=None):
        return self._add("Elu", [x], name, f=[alpha], attrs={"alpha": alpha}).outputs[0]

    # ---- pooling / reductions -------------------------------------------------------------------
    def max_pool(self, x, kernel, strides, pads_begin=(0, 0), pads_end=(0, 0), rounding="floor", name=None):
        i = [kernel[0], kernel[1], strides[0], strides[1], pads_begin[0], pads_begin[1], pads_end[0], pads_end[1], 1 if rounding == "ceil" else 0]
        attrs = {"strides": strides, "pads_begin": pads_begin, "pads_end": pads_end, "kernel": kernel, "rounding_type": rounding, "auto_pad": "explicit"}
        return self._add("MaxPool", [x], name, i=i, attrs=attrs).outputs[0]

    def avg_pool(self, x, kernel, strides, pads_begin=(0, 0), pads_end=(0, 0), exclude_pad=False, rounding="floor", name=None):
        i = [kernel[0], kernel[1], strides[0], strides[1], pads_begin[0], pads_begin[1], pads_end[0], pads_end[1],
             1 if rounding == "ceil" else 0, 1 if exclude_pad else 0]
        attrs = {"strides": strides, "pads_begin": pads_begin, "pads_end": pads_end, "kernel": kernel,
                 "exclude-pad": "true" if exclude_pad else "false", "rounding_type": rounding, "auto_pad": "explicit"}
        return self._add("AvgPool", [x], name, i=i, attrs=attrs).outputs[0]

    def _reduce(self, op, x, axes, keep_dims, name):
        axes = [int(a) for a in axes]
        return self._add(op, [x], name, i=[1 if keep_dims else 0, len(axes)] + axes,
                         attrs={"keep_dims": "true" if keep_dims else "false"}, ir_consts=[np.asarray(axes, np.int64)]).outputs[0]

    def reduce_mean(self, x, axes, keep_dims=False, name=None): return self._reduce("ReduceMean", x, axes, keep_dims, name)
    def reduce_sum(self, x, axes, keep_dims=False, name=None): return self._reduce("ReduceSum", x, axes, keep_dims, name)
    def reduce_max(self, x, axes, keep_dims=False, name=None): return self._reduce("ReduceMax", x, axes, keep_dims, name)

    # ---- matmul / softmax / mvn -------------------------------------------------------------------
    def matmul(self, a, b, transpose_a=False, transpose_b=False, name=None):
        b = b if isinstance(b, Output) else self.constant(np.asarray(b, np.float32))
        return self._add("MatMul", [a, b], name, i=[int(transpose_a), int(transpose_b)],
                         attrs={"transpose_a": "true" if transpose_a else "false", "transpose_b": "true" if transpose_b else "false"}).outputs[0]

    def softmax(self, x, axis, name=None):
        axis = axis if axis >= 0 else axis + len(x.shape)
        return self._add("Softmax", [x], name, i=[axis], attrs={"axis": axis}).outputs[0]

    def mvn(self, x, axes, normalize_variance=True, eps=1e-9, eps_mode="INSIDE_SQRT", name=None):
        axes = [int(a) for a in axes]
        i = [int(normalize_variance), 0 if eps_mode.upper() == "INSIDE_SQRT" else 1, len(axes)] + axes
        attrs = {"normalize_variance": "true" if normalize_variance else "false", "eps": eps, "eps_mode": eps_mode.upper()}
        return self._add("MVN", [x], name, i=i, f=[eps], attrs=attrs, version="opset6", ir_consts=[np.asarray(axes, np.int64)]).outputs[0]

    # ---- data movement ------------------------------------------------------------------------------
    def reshape(self, x, shape, name=None):
        shape = list(shape)
        total = int(np.prod(x.shape))
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1]))
            shape[shape.index(-1)] = total // max(known, 1)
        return self._add("Reshape", [x], name, shape=shape, attrs={"special_zero": "false"}, ir_consts=[np.asarray(shape, np.int64)]).outputs[0]

    def transpose(self, x, order, name=None):
        order = [int(o) for o in order]
        return self._add("Transpose", [x], name, i=order, ir_consts=[np.asarray(order, np.int64)]).outputs[0]

    def strided_slice(self, x, begin, end, strides=None, name=None):
        """opset1 StridedSlice with numpy semantics and no new-axis / shrink / ellipsis masks: x[b0:e0:s0, b1:e1:s1, ...].  `None` in begin /
        end sets the corresponding bit of begin_mask / end_mask (value ignored: the real start / end in the direction of the stride)."""
        rank = len(x.shape)
        begin = list(begin) + [None] * (rank - len(begin))
        end = list(end) + [None] * (rank - len(end))
        strides = list(strides or []) + [1] * (rank - len(strides or []))
        starts, counts = [], []
        for d in range(rank):
            rng = range(*slice(begin[d], end[d], strides[d]).indices(x.shape[d]))
            if len(rng) == 0:
                raise ValueError("strided_slice: empty result along dimension %d" % d)
            starts.append(rng.start)
            counts.append(len(rng))
        i = [rank] + list(x.shape) + starts + [int(v) for v in strides] + counts
        attrs = {"begin_mask": [1 if v is None else 0 for v in begin], "end_mask": [1 if v is None else 0 for v in end],
                 "new_axis_mask": [0] * rank, "shrink_axis_mask": [0] * rank, "ellipsis_mask": [0] * rank}
        consts = [np.asarray([0 if v is None else v for v in begin], np.int64), np.asarray([0 if v is None else v for v in end], np.int64),
                  np.asarray(strides, np.int64)]
        return self._add("StridedSlice", [x], name, i=i, shape=counts, attrs=attrs, ir_consts=consts).outputs[0]

    def concat(self, xs, axis, name=None):
        return self._add("Concat", list(xs), name, i=[axis], attrs={"axis": axis}).outputs[0]

    def split(self, x, axis, num_splits, name=None):
        n = self._add("Split", [x], name, i=[axis, num_splits], attrs={"num_splits": num_splits}, ir_consts=[np.asarray(axis, np.int64)])
        return n.outputs

    def gather(self, data, indices, axis=0, name=None):
        data = data if isinstance(data, Output) else self.constant(np.asarray(data, np.float32))
        return self._add("Gather", [data, indices], name, i=[axis], ir_consts=[np.asarray(axis, np.int64)]).outputs[0]

    def convert(self, x, dtype, name=None):
        dt = _NP2DT[np.dtype(dtype)]
        return self._add("Convert", [x], name, dtype=dt, attrs={"destination_type": DT_NAME[dt]}).outputs[0]

    # ---- selection ----------------------------------------------------------------------------------
    def topk(self, x, k, axis, mode="max", sort="value", name=None):
        axis = axis if axis >= 0 else axis + len(x.shape)
        i = [k, axis, 0 if mode == "max" else 1, {"none": 0, "value": 1, "index": 2}[sort]]
        attrs = {"axis": axis, "mode": mode, "sort": sort, "index_element_type": "i32"}
        return self._add("TopK", [x], name, i=i, attrs=attrs, version="opset3", ir_consts=[np.asarray(k, np.int64)]).outputs

    def non_max_suppression(self, boxes, scores, max_output_boxes_per_class, iou_threshold, score_threshold, soft_nms_sigma=0.0,
                            box_encoding="corner", sort_result_descending=True, name=None):
        i = [max_output_boxes_per_class, 1 if box_encoding == "center" else 0, int(sort_result_descending)]
        attrs = {"box_encoding": box_encoding, "sort_result_descending": "true" if sort_result_descending else "false", "output_type": "i32"}
        consts = [np.asarray(max_output_boxes_per_class, np.int64), np.asarray(iou_threshold, np.float32),
                  np.asarray(score_threshold, np.float32), np.asarray(soft_nms_sigma, np.float32)]
        return self._add("NonMaxSuppression", [boxes, scores], name, i=i, f=[iou_threshold, score_threshold, soft_nms_sigma], attrs=attrs,
                         version="opset5", ir_consts=consts).outputs

    def detection_output(self, loc, conf, priors, num_classes, background_label_id=0, top_k=-1, keep_top_k=-1,
                         code_type="caffe.PriorBoxParameter.CENTER_SIZE", share_location=True, nms_threshold=0.45,
                         confidence_threshold=0.01, variance_encoded_in_target=False, normalized=True, clip_before_nms=False,
                         clip_after_nms=False, input_height=1, input_width=1, decrease_label_id=False, name=None):
        priors = priors if isinstance(priors, Output) else self.constant(np.asarray(priors, np.float32))
        center = code_type.endswith("CENTER_SIZE")
        i = [num_classes, background_label_id, top_k, keep_top_k, int(center), int(share_location), int(variance_encoded_in_target),
             int(normalized), int(clip_before_nms), int(clip_after_nms), input_height, input_width, int(decrease_label_id)]
        b = lambda v: "true" if v else "false"
        attrs = {"num_classes": num_classes, "background_label_id": background_label_id, "top_k": top_k,
                 "variance_encoded_in_target": b(variance_encoded_in_target), "keep_top_k": [keep_top_k], "code_type": code_type,
                 "share_location": b(share_location), "nms_threshold": nms_threshold, "confidence_threshold": confidence_threshold,
                 "clip_after_nms": b(clip_after_nms), "clip_before_nms": b(clip_before_nms), "decrease_label_id": b(decrease_label_id),
                 "normalized": b(normalized), "input_height": input_height, "input_width": input_width, "objectness_score": 0}
        return self._add("DetectionOutput", [loc, conf, priors], name, i=i, f=[nms_threshold, confidence_threshold], attrs=attrs).outputs[0]
