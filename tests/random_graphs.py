"""Seeded random small networks over the ops of the hot path (convolutions of every flavour the compiler distinguishes, pooling, element-wise
chains, residual adds, concats, a FullyConnected head) for the differential tests: the CUDA path against the numpy oracle
(tests/test_gpu_random_graphs.py) and, without a GPU, the compiler against itself (tests/test_random_graphs_compile.py: every graph must
lower, plan and keep its output names).  Shapes are drawn to hit the awkward cases on purpose: channel counts that are not multiples of 8,
odd spatial sizes, strides / dilations / asymmetric padding, reduction lengths that are not multiples of 8 or 64."""
import numpy as np

from dldt_b200 import Function

CHANNELS = [1, 3, 4, 5, 8, 13, 16, 24, 32, 40, 48, 64, 72, 96, 128]
UNARY = ["relu", "sigmoid", "tanh", "hswish", "hsigmoid", "mish", "abs", "swish", "gelu", "clamp", "elu"]


def _act(f, rng, y):
    kind = UNARY[rng.integers(len(UNARY))]
    if kind == "clamp":
        return f.clamp(y, 0.0, 6.0)
    if kind == "swish":
        return f.swish(y, 1.0)
    if kind == "gelu":
        return f.gelu(y, "ERF" if rng.integers(2) else "TANH")
    if kind == "elu":
        return f.elu(y, 0.5)
    return getattr(f, kind)(y)


def _conv(f, rng, x, shape):
    n, c, h, w = shape
    k = int(rng.choice([1, 1, 3, 3, 5]))
    s = int(rng.choice([1, 1, 2]))
    d = int(rng.choice([1, 1, 1, 2])) if k > 1 else 1
    pb = (int(rng.integers(0, k // 2 + 1)), int(rng.integers(0, k // 2 + 1)))
    pe = (int(rng.integers(0, k // 2 + 1)), int(rng.integers(0, k // 2 + 1)))
    eff = d * (k - 1) + 1
    if h + pb[0] + pe[0] < eff or w + pb[1] + pe[1] < eff:
        k, d, eff, pb, pe = 1, 1, 1, (0, 0), (0, 0)
    mode = rng.integers(4)
    if mode == 0 and c > 1:                        # depthwise
        wt = (rng.standard_normal((c, 1, 1, k, k)) / k).astype(np.float32)
        y = f.group_convolution(x, wt, (s, s), pb, pe, (d, d))
        o = c
    elif mode == 1 and c % 4 == 0 and c >= 8:      # grouped
        g = int(rng.choice([2, 4]))
        o = g * int(rng.choice([1, 2, 3, 4]))
        wt = (rng.standard_normal((g, o // g, c // g, k, k)) / np.sqrt(c // g * k * k)).astype(np.float32)
        y = f.group_convolution(x, wt, (s, s), pb, pe, (d, d))
    else:                                          # dense
        o = int(rng.choice(CHANNELS))
        wt = (rng.standard_normal((o, c, k, k)) / np.sqrt(c * k * k)).astype(np.float32)
        y = f.convolution(x, wt, (s, s), pb, pe, (d, d))
    ho = (h + pb[0] + pe[0] - eff) // s + 1
    wo = (w + pb[1] + pe[1] - eff) // s + 1
    if rng.integers(4):
        y = f.add(y, rng.standard_normal((1, o, 1, 1)).astype(np.float32) * 0.1)
    if rng.integers(3):
        y = _act(f, rng, y)
    return y, (n, o, ho, wo)


def _pool(f, rng, x, shape):
    n, c, h, w = shape
    k = int(rng.choice([2, 3]))
    s = int(rng.choice([1, 2]))
    p = int(rng.integers(0, 2))
    if h + 2 * p < k or w + 2 * p < k:
        return x, shape
    rounding = "ceil" if rng.integers(3) == 0 else "floor"

    def out(v, mode):
        num = v + 2 * p - k
        return (-(-num // s) if mode == "ceil" else num // s) + 1
    # ceil rounding may add a window that lies entirely in the padding: with exclude-pad that is 0 / 0 - the nGraph reference refuses it
    # ("AvgPool elements == 0"), MKLDNNPlugin and this plugin return different garbage.  Not a defined case: fall back to floor.
    if rounding == "ceil" and any((out(v, "ceil") - 1) * s - p >= v for v in (h, w)):
        rounding = "floor"
    if rng.integers(2):
        y = f.max_pool(x, (k, k), (s, s), (p, p), (p, p), rounding=rounding)
    else:
        y = f.avg_pool(x, (k, k), (s, s), (p, p), (p, p), exclude_pad=bool(rng.integers(2)), rounding=rounding)
    return y, (n, c, out(h, rounding), out(w, rounding))


def _inverted_residual(f, rng, x, shape):
    """MobileNetV2 block (expand 1x1 + ReLU6 -> depthwise 3x3 + ReLU6 -> project 1x1 [+ input]): the shape the fused block kernel takes."""
    n, c, h, w = shape
    t = int(rng.choice([1, 2, 4, 6]))
    s = int(rng.choice([1, 1, 2]))
    o = c if rng.integers(2) else int(rng.choice(CHANNELS[4:]))
    hid = c * t
    y = x
    if t > 1:
        y = f.clamp(f.add(f.convolution(y, (rng.standard_normal((hid, c, 1, 1)) / np.sqrt(c)).astype(np.float32), (1, 1), (0, 0), (0, 0)),
                          rng.standard_normal((1, hid, 1, 1)).astype(np.float32) * 0.1), 0.0, 6.0)
    y = f.clamp(f.add(f.group_convolution(y, (rng.standard_normal((hid, 1, 1, 3, 3)) / 3).astype(np.float32), (s, s), (1, 1), (1, 1)),
                      rng.standard_normal((1, hid, 1, 1)).astype(np.float32) * 0.1), 0.0, 6.0)
    y = f.add(f.convolution(y, (rng.standard_normal((o, hid, 1, 1)) / np.sqrt(hid)).astype(np.float32), (1, 1), (0, 0), (0, 0)),
              rng.standard_normal((1, o, 1, 1)).astype(np.float32) * 0.1)
    ho, wo = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
    if s == 1 and o == c:
        y = f.add(y, x)
    return y, (n, o, ho, wo)


def _bottleneck(f, rng, x, shape):
    """ResNet bottleneck with a projection shortcut (1x1 reduce -> 3x3 -> 1x1 expand, + 1x1 strided shortcut, ReLU): the K-concatenated
    shortcut fold and the strided second source of the pair kernel."""
    n, c, h, w = shape
    mid = int(rng.choice([8, 16, 32, 64]))
    o = int(rng.choice([32, 64, 128, 256]))
    s = int(rng.choice([1, 2]))

    def cbr(v, wshape, stride, pad, relu=True):
        fan = wshape[1] * wshape[2] * wshape[3]
        v = f.add(f.convolution(v, (rng.standard_normal(wshape) / np.sqrt(fan)).astype(np.float32), (stride, stride), (pad, pad), (pad, pad)),
                  rng.standard_normal((1, wshape[0], 1, 1)).astype(np.float32) * 0.1)
        return f.relu(v) if relu else v
    a = cbr(x, (mid, c, 1, 1), 1, 0)
    b = cbr(a, (mid, mid, 3, 3), s, 1)
    cc = cbr(b, (o, mid, 1, 1), 1, 0, relu=False)
    sc = cbr(x, (o, c, 1, 1), s, 0, relu=False)
    ho, wo = (h + 2 - 3) // s + 1, (w + 2 - 3) // s + 1
    return f.relu(f.add(cc, sc)), (n, o, ho, wo)


def build(seed):
    """Returns (Function, {input name: array}).  Seeds >= 1000 draw wider tensors (64-512 channels, up to 28 x 28, batch up to 4): several
    k-blocks and several tiles per tcgen05 launch."""
    rng = np.random.default_rng(seed)
    f = Function("random_%d" % seed)
    if seed >= 1000:
        shape = (int(rng.integers(1, 5)), int(rng.choice([64, 96, 128, 192, 256, 320, 512])), int(rng.integers(7, 29)), int(rng.integers(7, 29)))
    else:
        shape = (int(rng.integers(1, 4)), int(rng.choice(CHANNELS[:9])), int(rng.integers(6, 21)), int(rng.integers(6, 21)))
    x = f.parameter(shape)
    inputs = {"Parameter_0": rng.standard_normal(shape).astype(np.float32)}
    y = x
    for _ in range(int(rng.integers(2, 6))):
        op = rng.integers(8)
        if op == 6:
            y, shape = _inverted_residual(f, rng, y, shape)
        elif op == 7:
            y, shape = _bottleneck(f, rng, y, shape)
        elif op <= 2:
            y, shape = _conv(f, rng, y, shape)
        elif op == 3:
            y, shape = _pool(f, rng, y, shape)
        elif op == 4:                               # residual block: y + act(conv1x1(y))
            c = shape[1]
            wt = (rng.standard_normal((c, c, 1, 1)) / np.sqrt(c)).astype(np.float32)
            z = f.convolution(y, wt, (1, 1), (0, 0), (0, 0))
            y = f.add(_act(f, rng, z), y)
            if rng.integers(2):
                y = f.relu(y)
        else:                                       # two branches concatenated on channels
            c = shape[1]
            o = int(rng.choice(CHANNELS[1:8]))
            wt = (rng.standard_normal((o, c, 1, 1)) / np.sqrt(c)).astype(np.float32)
            y = f.concat([y, f.relu(f.convolution(y, wt, (1, 1), (0, 0), (0, 0)))], 1)
            shape = (shape[0], c + o, shape[2], shape[3])
        if min(shape[2], shape[3]) < 2:
            break
    f.result(y)
    # head: global average pool -> flatten -> FullyConnected (reduction length = whatever channel count came out) -> softmax
    n, c = shape[0], shape[1]
    g = f.reshape(f.reduce_mean(y, (2, 3), keep_dims=False), (n, c))
    classes = int(rng.choice([10, 17, 100]))
    logits = f.add(f.matmul(g, (rng.standard_normal((classes, c)) / np.sqrt(c)).astype(np.float32), transpose_b=True),
                   rng.standard_normal((1, classes)).astype(np.float32) * 0.1)
    f.result(f.softmax(logits, 1))
    return f, inputs


# ---- second family: sequence models (the MatMul + Softmax + LayerNorm path) -------------------------------------------------------------
def build_sequence(seed):
    """Random encoder-style network on [B, S, H] tensors: embedding Gather, LayerNorm (MVN-6 + scale + shift), dense layers with awkward
    sizes, multi-head attention (head split / merge by Reshape + Transpose, QK^T, scale, additive mask, Softmax, PV), GELU FFN, residuals."""
    rng = np.random.default_rng(seed + 77777)
    f = Function("random_seq_%d" % seed)
    batch = int(rng.integers(1, 4))
    seq = int(rng.choice([5, 8, 16, 24, 33, 64, 128]))
    heads = int(rng.choice([1, 2, 4]))
    dh = int(rng.choice([8, 16, 32, 64]))
    hidden = heads * dh
    vocab = int(rng.integers(20, 200))
    ids = f.parameter((batch, seq), np.int32, name="ids")
    inputs = {"ids": rng.integers(0, vocab, (batch, seq)).astype(np.int32)}

    def vec(n, lo=-0.1, hi=0.1):
        return rng.uniform(lo, hi, n).astype(np.float32)

    def dense(x, o, i):
        y = f.matmul(x, (rng.standard_normal((o, i)) / np.sqrt(i)).astype(np.float32), transpose_b=True)
        return f.add(y, vec(o))

    def layer_norm(x):
        y = f.mvn(x, [len(x.shape) - 1], True, float(rng.choice([1e-12, 1e-5])), "INSIDE_SQRT")
        return f.add(f.multiply(y, rng.uniform(0.8, 1.2, hidden).astype(np.float32)), vec(hidden))

    x = f.gather((rng.standard_normal((vocab, hidden)) * 0.5).astype(np.float32), ids, 0)
    x = layer_norm(f.add(x, f.constant((rng.standard_normal((1, seq, hidden)) * 0.1).astype(np.float32))))
    use_mask = bool(rng.integers(2))
    if use_mask:
        mask = f.parameter((batch, seq), np.int32, name="mask")
        mv = np.ones((batch, seq), np.int32)
        for b in range(batch):
            mv[b, int(rng.integers(seq // 2 + 1, seq + 1)):] = 0
        inputs["mask"] = mv
        m = f.reshape(f.multiply(f.subtract(np.float32(1.0), f.convert(mask, np.float32)), np.float32(-10000.0)), (batch, 1, 1, seq))
    for _ in range(int(rng.integers(1, 3))):
        def split_heads(t):
            return f.transpose(f.reshape(t, (batch, seq, heads, dh)), (0, 2, 1, 3))
        q, k, v = split_heads(dense(x, hidden, hidden)), split_heads(dense(x, hidden, hidden)), split_heads(dense(x, hidden, hidden))
        s = f.multiply(f.matmul(q, k, transpose_b=True), np.float32(1.0 / np.sqrt(dh)))
        if use_mask:
            s = f.add(s, m)
        c = f.matmul(f.softmax(s, 3), v)
        c = f.reshape(f.transpose(c, (0, 2, 1, 3)), (batch, seq, hidden))
        x = layer_norm(f.add(dense(c, hidden, hidden), x))
        ffn = int(rng.choice([hidden, 2 * hidden, 4 * hidden, 3 * hidden + 8]))
        act = rng.integers(3)
        h = dense(x, ffn, hidden)
        h = f.gelu(h, "ERF") if act == 0 else (f.gelu(h, "TANH") if act == 1 else f.relu(h))
        x = layer_norm(f.add(dense(h, hidden, ffn), x))
    f.result(dense(x, int(rng.choice([2, 8, 11])), hidden))
    return f, inputs


# ---- third family: element-wise / data-movement / selection ops around a small feature map ---------------------------------------------------
def build_misc(seed):
    rng = np.random.default_rng(seed + 424242)
    f = Function("random_misc_%d" % seed)
    n, c, h, w = int(rng.integers(1, 4)), int(rng.choice([2, 4, 6, 8, 12, 16, 20])), int(rng.integers(3, 12)), int(rng.integers(3, 12))
    x = f.parameter((n, c, h, w))
    inputs = {"Parameter_0": rng.standard_normal((n, c, h, w)).astype(np.float32)}
    y = x
    shape = [n, c, h, w]
    for _ in range(int(rng.integers(2, 6))):
        op = int(rng.integers(12))
        if op == 11:      # StridedSlice: random begin / end / stride per axis (negative strides, ignored bounds)
            begin, end, strides = [], [], []
            for dim in shape:
                st = int(rng.choice([1, 1, 1, 2, 3, -1, -2])) if dim > 1 else 1
                lo = int(rng.integers(0, dim))
                hi = int(rng.integers(lo + 1, dim + 1))
                if st > 0:
                    begin.append(None if rng.integers(4) == 0 else lo)
                    end.append(None if rng.integers(4) == 0 else hi)
                else:          # backwards from hi - 1 down to lo (exclusive lo - 1: None when lo == 0)
                    begin.append(None if rng.integers(4) == 0 else hi - 1)
                    end.append(None if (lo == 0 or rng.integers(4) == 0) else lo - 1)
                strides.append(st)
            y = f.strided_slice(y, begin, end, strides)
        elif op == 0:       # binary op with a broadcast constant of a random broadcastable shape
            bshape = [d if rng.integers(2) else 1 for d in shape]
            cst = rng.uniform(0.5, 1.5, bshape).astype(np.float32)
            kind = ["add", "subtract", "multiply", "divide", "maximum", "minimum", "squared_difference"][rng.integers(7)]
            y = getattr(f, kind)(y, cst) if rng.integers(2) else getattr(f, kind)(f.constant(cst), y)
        elif op == 1:
            y = _act(f, rng, y)
        elif op == 2 and len(shape) == 4:     # per-channel PReLU with the 1-D slope Model Optimizer emits
            y = f.prelu(y, rng.uniform(0.05, 0.3, shape[1]).astype(np.float32))
        elif op == 3:
            perm = list(rng.permutation(len(shape)))
            y = f.transpose(y, perm)
            shape = [shape[p] for p in perm]
        elif op == 4:
            axis = int(rng.integers(len(shape)))
            parts = [k for k in (2, 3, 4) if shape[axis] % k == 0]
            if parts:
                outs = f.split(y, axis, int(rng.choice(parts)))
                y = f.concat(list(outs[::-1]), axis)
        elif op == 5 and len(shape) > 2:
            axes = sorted(set(int(a) for a in rng.integers(1, len(shape), int(rng.integers(1, 3)))))
            keep = bool(rng.integers(2))
            y = [f.reduce_mean, f.reduce_sum, f.reduce_max][rng.integers(3)](y, axes, keep_dims=keep)
            shape = [1 if k in axes else d for k, d in enumerate(shape)] if keep else [d for k, d in enumerate(shape) if k not in axes]
        elif op == 6:
            total = int(np.prod(shape))
            cands = [s for s in ([shape[0], total // shape[0]], [total], [shape[0], -1, shape[-1]]) if total % max(abs(int(np.prod(s))), 1) == 0]
            tgt = cands[rng.integers(len(cands))]
            y = f.reshape(y, tgt)
            shape = list(y.shape)
        elif op == 7:
            y = f.softmax(y, int(rng.integers(len(shape))))
        elif op == 8 and len(shape) >= 2:
            axes = [len(shape) - 1] if rng.integers(2) else list(range(max(1, len(shape) - 2), len(shape)))
            y = f.mvn(y, axes, bool(rng.integers(2)), 1e-6, "INSIDE_SQRT" if rng.integers(2) else "OUTSIDE_SQRT")
        elif op == 9:
            y = f.power(f.add(f.abs(y), np.float32(0.5)), np.float32(rng.choice([0.5, 2.0, 3.0, -1.0])))
        elif op == 10:
            y = f.concat([y, f.relu(y)], int(rng.integers(len(shape))))
            shape = list(y.shape)
        shape = list(y.shape)
    f.result(y)
    if shape and shape[-1] >= 2 and rng.integers(2):
        k = int(rng.integers(1, min(shape[-1], 5) + 1))
        vals, idx = f.topk(y, k, len(shape) - 1, mode="max" if rng.integers(2) else "min", sort="value")
        f.result(vals)
        f.result(idx)
    return f, inputs


# ---- fourth family: selection ops with bit-exact outputs (NonMaxSuppression-5, DetectionOutput) ----------------------------------------------
def _rand_boxes(rng, b, n, center):
    xy = rng.uniform(0, 0.8, (b, n, 2))
    wh = rng.uniform(0.05, 0.4, (b, n, 2))
    return (np.concatenate([xy + wh / 2, wh], axis=2) if center else np.concatenate([xy, xy + wh], axis=2)).astype(np.float32)


def build_nms(seed):
    rng = np.random.default_rng(seed + 9091)
    f = Function("random_nms_%d" % seed)
    b, c, n = int(rng.integers(1, 4)), int(rng.integers(1, 5)), int(rng.integers(4, 150))
    center = bool(rng.integers(2))
    boxes = f.parameter((b, n, 4), name="boxes")
    scores = f.parameter((b, c, n), name="scores")
    soft = float(rng.choice([0.0, 0.0, 0.25, 0.5, 1.5]))
    i, s, v = f.non_max_suppression(boxes, scores, int(rng.choice([1, 2, 3, 10, 200])), float(rng.choice([0.0, 0.3, 0.5, 0.7, 1.0])),
                                    float(rng.choice([0.0, 0.1, 0.4, 0.95])), soft, "center" if center else "corner", bool(rng.integers(2)), name="nms")
    f.result(i)
    f.result(s)
    f.result(v)
    sc = rng.uniform(0, 1, (b, c, n))
    if rng.integers(2):
        sc = np.round(sc * 8) / 8          # score ties: the index order decides
    bx = _rand_boxes(rng, b, n, center)
    if rng.integers(3) == 0:
        bx[:, n // 2:] = bx[:, : n - n // 2]   # duplicate boxes: IoU exactly 1
    return f, {"boxes": bx, "scores": sc.astype(np.float32)}


def build_detection_output(seed):
    rng = np.random.default_rng(seed + 5150)
    f = Function("random_detout_%d" % seed)
    n, priors, classes = int(rng.integers(1, 4)), int(rng.integers(8, 200)), int(rng.integers(2, 8))
    share = bool(rng.integers(2))
    var_in_target = bool(rng.integers(2))
    center = bool(rng.integers(2))
    nloc = 1 if share else classes
    loc = f.parameter((n, priors * nloc * 4), name="loc")
    conf = f.parameter((n, priors * classes), name="conf")
    pb = _rand_boxes(rng, 1, priors, False)[0]
    var = np.tile(np.asarray([0.1, 0.1, 0.2, 0.2], np.float32), (priors, 1))
    pri = pb.reshape(1, 1, -1) if var_in_target else np.stack([pb.reshape(-1), var.reshape(-1)])[None]
    top_k = int(rng.choice([-1, 5, 20, 400]))
    keep_top_k = int(rng.choice([-1, 1, 10, 50]))
    det = f.detection_output(loc, conf, pri.astype(np.float32), classes, int(rng.integers(0, classes)) if rng.integers(2) else -1, top_k, keep_top_k,
                             "caffe.PriorBoxParameter.CENTER_SIZE" if center else "caffe.PriorBoxParameter.CORNER", share,
                             float(rng.choice([0.3, 0.45, 0.6])), float(rng.choice([0.01, 0.05, 0.2])), var_in_target, True,
                             bool(rng.integers(2)), bool(rng.integers(2)), name="det")
    f.result(det)
    logits = rng.standard_normal((n, priors, classes)) * 2
    p = np.exp(logits - logits.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    return f, {"loc": (rng.standard_normal((n, priors * nloc * 4)) * 0.5).astype(np.float32), "conf": p.reshape(n, -1).astype(np.float32)}


# ---- fifth family: pooling corner cases -------------------------------------------------------------------------------------------------------
def build_pool(seed):
    """One or two pooling layers with non-square kernels / strides, asymmetric padding, both rounding modes, include / exclude padding,
    after an optional 1x1 convolution (so the pooled tensor is channels-last on the device) - the divisor and window-clipping rules of
    reference/avg_pool.hpp and reference/max_pool.hpp.  Windows that lie entirely in the padding are not drawn (undefined: 0 / 0)."""
    rng = np.random.default_rng(seed + 31337)
    f = Function("random_pool_%d" % seed)
    n, c, h, w = int(rng.integers(1, 3)), int(rng.choice([1, 3, 8, 13, 16, 24])), int(rng.integers(4, 18)), int(rng.integers(4, 18))
    x = f.parameter((n, c, h, w))
    inputs = {"Parameter_0": rng.standard_normal((n, c, h, w)).astype(np.float32)}
    y = x
    if rng.integers(2):
        o = int(rng.choice([4, 8, 12, 16]))
        y = f.relu(f.convolution(y, (rng.standard_normal((o, c, 1, 1)) / np.sqrt(c)).astype(np.float32), (1, 1), (0, 0), (0, 0)))
        c = o
    for _ in range(int(rng.integers(1, 3))):
        for _try in range(20):
            k = (int(rng.integers(1, 6)), int(rng.integers(1, 6)))
            s = (int(rng.integers(1, 4)), int(rng.integers(1, 4)))
            pb = (int(rng.integers(0, k[0] // 2 + 1)), int(rng.integers(0, k[1] // 2 + 1)))
            pe = (int(rng.integers(0, k[0] // 2 + 1)), int(rng.integers(0, k[1] // 2 + 1)))
            rounding = "ceil" if rng.integers(2) else "floor"
            dims = []
            ok = True
            for d, v in enumerate((h, w)):
                num = v + pb[d] + pe[d] - k[d]
                if num < 0:
                    ok = False
                    break
                o_ = (-(-num // s[d]) if rounding == "ceil" else num // s[d]) + 1
                if (o_ - 1) * s[d] - pb[d] >= v:          # last window starts past the input: only padding inside
                    ok = False
                    break
                dims.append(o_)
            if ok:
                break
        else:
            break
        if rng.integers(2):
            y = f.max_pool(y, k, s, pb, pe, rounding=rounding)
        else:
            y = f.avg_pool(y, k, s, pb, pe, exclude_pad=bool(rng.integers(2)), rounding=rounding)
        h, w = dims
        if min(h, w) < 2:
            break
    f.result(y)
    return f, inputs


# ---- sixth family: convolution geometry corner cases ----------------------------------------------------------------------------------------
def build_conv_corner(seed):
    """ONE convolution (+ bias, + activation, sometimes + residual) with everything non-square: kernel (1..7 x 1..7, e.g. the 1x7 / 7x1 pairs of
    Inception), strides and dilations per axis, asymmetric padding, dense / grouped / depthwise, channel counts on both sides of the
    tensor-core tile sizes (the implicit-GEMM, halo, small-Cin, depthwise and fp32 kernels all get reached)."""
    rng = np.random.default_rng(seed + 271828)
    f = Function("random_conv_%d" % seed)
    n = int(rng.integers(1, 4))
    c = int(rng.choice([1, 3, 4, 8, 16, 24, 32, 64, 64, 96, 128, 160, 256]))
    h, w = int(rng.integers(5, 30)), int(rng.integers(5, 30))
    x = f.parameter((n, c, h, w))
    inputs = {"Parameter_0": rng.standard_normal((n, c, h, w)).astype(np.float32)}
    for _try in range(50):
        k = (int(rng.choice([1, 1, 2, 3, 3, 5, 7])), int(rng.choice([1, 1, 2, 3, 3, 5, 7])))
        s = (int(rng.choice([1, 1, 2, 3])), int(rng.choice([1, 1, 2, 3])))
        d = (int(rng.choice([1, 1, 2, 3])), int(rng.choice([1, 1, 2, 3])))
        pb = (int(rng.integers(0, k[0])), int(rng.integers(0, k[1])))
        pe = (int(rng.integers(0, k[0])), int(rng.integers(0, k[1])))
        if all(v + pb[a] + pe[a] >= d[a] * (k[a] - 1) + 1 for a, v in enumerate((h, w))):
            break
    else:
        k, s, d, pb, pe = (1, 1), (1, 1), (1, 1), (0, 0), (0, 0)
    mode = int(rng.integers(5))
    if mode == 0 and c > 1:
        wt = (rng.standard_normal((c, 1, 1) + k) / np.sqrt(k[0] * k[1])).astype(np.float32)
        y = f.group_convolution(x, wt, s, pb, pe, d)
        o = c
    elif mode == 1 and c % 4 == 0 and c >= 8:
        g = int(rng.choice([2, 4]))
        o = g * int(rng.choice([1, 2, 4, 8, 16]))
        wt = (rng.standard_normal((g, o // g, c // g) + k) / np.sqrt(c // g * k[0] * k[1])).astype(np.float32)
        y = f.group_convolution(x, wt, s, pb, pe, d)
    else:
        o = int(rng.choice([1, 5, 8, 16, 32, 64, 64, 72, 128, 192, 256, 320]))
        wt = (rng.standard_normal((o, c) + k) / np.sqrt(c * k[0] * k[1])).astype(np.float32)
        y = f.convolution(x, wt, s, pb, pe, d)
    if rng.integers(4):
        y = f.add(y, rng.standard_normal((1, o, 1, 1)).astype(np.float32) * 0.1)
    if rng.integers(3):
        y = _act(f, rng, y)
    if rng.integers(4) == 0:      # a second convolution on top: 1x1 with a residual back to the first one's output
        wt2 = (rng.standard_normal((o, o, 1, 1)) / np.sqrt(o)).astype(np.float32)
        y = f.relu(f.add(f.convolution(y, wt2, (1, 1), (0, 0), (0, 0)), y))
    f.result(y)
    return f, inputs
