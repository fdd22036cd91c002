"""nGraph-style builders of the five benchmark configurations (BASELINE.json `configs`), with
deterministic random-init weights (there are no checkpoints offline).  Initialisation follows
SURVEY.md 8d: conv ~ N(0, sqrt(2/fan_in)), BN gamma U[0.5,1.5], beta/mean U[-0.1,0.1], var U[0.5,1.5],
FC ~ N(0, 0.01).  Every builder takes size knobs so the tests can run reduced variants of the
same topology through the oracle in seconds.
"""
import numpy as np

from .graph import Function


class _Init:
    def __init__(self, seed):
        self.rng = np.random.default_rng(seed)

    def conv(self, o, i, kh, kw):
        return (self.rng.standard_normal((o, i, kh, kw)) * np.sqrt(2.0 / (i * kh * kw))).astype(np.float32)

    def bn(self, c):
        r = self.rng
        return (r.uniform(0.5, 1.5, c).astype(np.float32), r.uniform(-0.1, 0.1, c).astype(np.float32),
                r.uniform(-0.1, 0.1, c).astype(np.float32), r.uniform(0.5, 1.5, c).astype(np.float32))

    def fc(self, o, i, std=0.01):
        return (self.rng.standard_normal((o, i)) * std).astype(np.float32)

    def vec(self, n, lo=-0.1, hi=0.1):
        return self.rng.uniform(lo, hi, n).astype(np.float32)


def _conv_bn(f, init, x, cout, k, stride=1, pad=None, act=None, name=""):
    cin = x.shape[1]
    pad = (k // 2) if pad is None else pad
    y = f.convolution(x, init.conv(cout, cin, k, k), (stride, stride), (pad, pad), (pad, pad), name=name + "_conv")
    g, b, m, v = init.bn(cout)
    y = f.batch_norm_inference(y, g, b, m, v, 1e-5, name=name + "_bn")
    if act == "relu":
        y = f.relu(y, name=name + "_relu")
    elif act == "relu6":
        y = f.clamp(y, 0.0, 6.0, name=name + "_relu6")
    return y


# ---------------------------------------------------------------------------------------------------
def resnet50(batch=1, image=224, width=64, blocks=(3, 4, 6, 3), num_classes=1000, seed=1, input_dtype=np.float32, softmax=False):
    """ResNet-50 v1.5 (stride 2 on the 3x3 of each downsampling bottleneck)."""
    f = Function("resnet50_v1.5")
    init = _Init(seed)
    x = f.parameter((batch, 3, image, image), input_dtype, name="data")
    if np.dtype(input_dtype) != np.float32:
        x = f.convert(x, np.float32, name="data_f32")
    y = _conv_bn(f, init, x, width, 7, 2, 3, "relu", "conv1")
    y = f.max_pool(y, (3, 3), (2, 2), (1, 1), (1, 1), name="pool1")
    cin = width
    for si, nb in enumerate(blocks):
        mid = width * (2 ** si)
        cout = mid * 4
        for bi in range(nb):
            stride = 2 if (bi == 0 and si > 0) else 1
            nm = "res%d_%d" % (si + 2, bi)
            sc = y
            if bi == 0:
                sc = _conv_bn(f, init, y, cout, 1, stride, 0, None, nm + "_sc")
            z = _conv_bn(f, init, y, mid, 1, 1, 0, "relu", nm + "_a")
            z = _conv_bn(f, init, z, mid, 3, stride, 1, "relu", nm + "_b")
            z = _conv_bn(f, init, z, cout, 1, 1, 0, None, nm + "_c")
            y = f.relu(f.add(z, sc, name=nm + "_add"), name=nm + "_out")
            cin = cout
    y = f.reduce_mean(y, (2, 3), keep_dims=False, name="pool5")
    y = f.matmul(y, init.fc(num_classes, cin), transpose_b=True, name="fc")
    y = f.add(y, init.vec(num_classes).reshape(1, -1), name="fc_bias")
    if softmax:
        y = f.softmax(y, 1, name="prob")
    f.result(y)
    return f


RESNET50_GMAC_PER_IMAGE = 4.0892  # conv + FC MACs at 224x224


# ---------------------------------------------------------------------------------------------------
def mobilenet_v2(batch=1, image=224, width_mult=1.0, num_classes=1000, seed=2, settings=None):
    f = Function("mobilenet_v2")
    init = _Init(seed)
    settings = settings or [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]
    ch = lambda c: max(8, int(c * width_mult + 4) // 8 * 8)
    x = f.parameter((batch, 3, image, image), np.float32, name="data")
    y = _conv_bn(f, init, x, ch(32), 3, 2, 1, "relu6", "conv_first")
    cin = ch(32)
    for bi, (t, c, n, s) in enumerate(settings):
        cout = ch(c)
        for i in range(n):
            stride = s if i == 0 else 1
            nm = "ir%d_%d" % (bi, i)
            z = y
            hidden = cin * t
            if t != 1:
                z = _conv_bn(f, init, z, hidden, 1, 1, 0, "relu6", nm + "_expand")
            wdw = (init.rng.standard_normal((hidden, 1, 1, 3, 3)) * np.sqrt(2.0 / 9)).astype(np.float32)
            z = f.group_convolution(z, wdw, (stride, stride), (1, 1), (1, 1), name=nm + "_dw_conv")
            g, b, m, v = init.bn(hidden)
            z = f.clamp(f.batch_norm_inference(z, g, b, m, v, 1e-5, name=nm + "_dw_bn"), 0.0, 6.0, name=nm + "_dw_relu6")
            z = _conv_bn(f, init, z, cout, 1, 1, 0, None, nm + "_project")
            y = f.add(y, z, name=nm + "_add") if (stride == 1 and cin == cout) else z
            cin = cout
    last = max(1280, int(1280 * width_mult)) if width_mult >= 1.0 else ch(1280)
    y = _conv_bn(f, init, y, last, 1, 1, 0, "relu6", "conv_last")
    y = f.reduce_mean(y, (2, 3), keep_dims=False, name="pool")
    y = f.matmul(y, init.fc(num_classes, last), transpose_b=True, name="fc")
    y = f.add(y, init.vec(num_classes).reshape(1, -1), name="fc_bias")
    f.result(y)
    return f


# ---------------------------------------------------------------------------------------------------
def bert_base(batch=1, seq=384, hidden=768, heads=12, layers=12, ffn=3072, vocab=30522, max_pos=512, seed=3):
    """BERT-base encoder: Gather embeddings, LayerNorm as MVN-6 + Mul + Add (eps 1e-12), GELU(erf)."""
    f = Function("bert_base")
    init = _Init(seed)
    rng = init.rng
    dh = hidden // heads
    ids = f.parameter((batch, seq), np.int32, name="input_ids")
    mask = f.parameter((batch, seq), np.int32, name="attention_mask")
    types = f.parameter((batch, seq), np.int32, name="token_type_ids")

    def emb(n):
        return (rng.standard_normal((n, hidden)) * 0.02).astype(np.float32)

    def layer_norm(x, nm):
        y = f.mvn(x, [len(x.shape) - 1], True, 1e-12, "INSIDE_SQRT", name=nm + "_mvn")
        y = f.multiply(y, rng.uniform(0.8, 1.2, hidden).astype(np.float32), name=nm + "_gamma")
        return f.add(y, init.vec(hidden), name=nm + "_beta")

    def dense(x, o, i, nm, std=0.02):
        y = f.matmul(x, init.fc(o, i, std), transpose_b=True, name=nm)
        return f.add(y, init.vec(o, -0.02, 0.02), name=nm + "_bias")

    x = f.gather(emb(vocab), ids, 0, name="word_emb")
    pos = f.constant(emb(max_pos)[:seq].reshape(1, seq, hidden), name="pos_emb")
    x = f.add(f.add(x, pos, name="emb_add_pos"), f.gather(emb(2), types, 0, name="type_emb"), name="emb_add_type")
    x = layer_norm(x, "emb_ln")
    m = f.convert(mask, np.float32, name="mask_f32")
    m = f.multiply(f.subtract(np.float32(1.0), m, name="mask_inv"), np.float32(-10000.0), name="mask_scale")
    m = f.reshape(m, (batch, 1, 1, seq), name="mask_4d")
    for li in range(layers):
        nm = "l%d" % li

        def split_heads(t, tag):
            t = f.reshape(t, (batch, seq, heads, dh), name="%s_%s_r" % (nm, tag))
            return f.transpose(t, (0, 2, 1, 3), name="%s_%s_t" % (nm, tag))

        q = split_heads(dense(x, hidden, hidden, nm + "_q"), "q")
        k = split_heads(dense(x, hidden, hidden, nm + "_k"), "k")
        v = split_heads(dense(x, hidden, hidden, nm + "_v"), "v")
        s = f.matmul(q, k, transpose_b=True, name=nm + "_qk")
        s = f.multiply(s, np.float32(1.0 / np.sqrt(dh)), name=nm + "_scale")
        s = f.add(s, m, name=nm + "_mask")
        p = f.softmax(s, 3, name=nm + "_softmax")
        c = f.matmul(p, v, name=nm + "_av")
        c = f.reshape(f.transpose(c, (0, 2, 1, 3), name=nm + "_ctx_t"), (batch, seq, hidden), name=nm + "_ctx_r")
        a = dense(c, hidden, hidden, nm + "_o")
        x = layer_norm(f.add(a, x, name=nm + "_res1"), nm + "_ln1")
        h = f.gelu(dense(x, ffn, hidden, nm + "_ffn1"), "ERF", name=nm + "_gelu")
        h = dense(h, hidden, ffn, nm + "_ffn2")
        x = layer_norm(f.add(h, x, name=nm + "_res2"), nm + "_ln2")
    y = dense(x, 8, hidden, "qa_outputs")
    f.result(y)
    return f


def bert_gflop_per_seq(seq=384, hidden=768, layers=12, ffn=3072, heads=12):
    mac = layers * (4 * seq * hidden * hidden + 2 * seq * hidden * ffn + 2 * seq * seq * hidden)
    return 2.0 * mac / 1e9


# ---------------------------------------------------------------------------------------------------
def _ssd_priors(fmaps, image, aspect, rng):
    """Deterministic SSD-style prior boxes [1,2,P*4] (boxes, variances); PriorBox is constant-folded
    in the reference anyway (ngraph/core/src/op/prior_box.cpp:149-177)."""
    boxes = []
    nf = len(fmaps)
    for k, (fm, na) in enumerate(zip(fmaps, aspect)):
        sk = 0.1 + 0.8 * k / max(nf - 1, 1)
        sk1 = 0.1 + 0.8 * (k + 1) / max(nf - 1, 1)
        sizes = [(sk, sk), (np.sqrt(sk * sk1), np.sqrt(sk * sk1))]
        for ar in (2.0, 3.0)[: (na - 2) // 2]:
            sizes += [(sk * np.sqrt(ar), sk / np.sqrt(ar)), (sk / np.sqrt(ar), sk * np.sqrt(ar))]
        for i in range(fm):
            for j in range(fm):
                cx, cy = (j + 0.5) / fm, (i + 0.5) / fm
                for w, h in sizes[:na]:
                    boxes.append([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2])
    boxes = np.clip(np.asarray(boxes, np.float32), 0.0, 1.0)
    var = np.tile(np.asarray([0.1, 0.1, 0.2, 0.2], np.float32), (boxes.shape[0], 1))
    return np.stack([boxes.reshape(-1), var.reshape(-1)])[None]


def ssd300_resnet34(batch=1, image=300, width=64, num_classes=81, seed=4, trunk_blocks=(3, 4, 6), keep_top_k=200, top_k=200,
                    input_dtype=np.float32, head_scale=0.5, conf_scale=None):
    """SSD300 with a ResNet-34 trunk to conv4_x (stride-1 variant), 5 extra blocks, 6 heads and
    DetectionOutput.  head_scale multiplies the He-initialised head weights: at full width the default 0.5 (what
    the committed golden cases were generated with) lets the random-init location head overflow exp() in the box decoder; the full-size
    runs of tools/gpu_configs.py use 0.02.  conf_scale (default: head_scale) scales the confidence heads separately so that the full-size
    parity fixtures (tools/make_config_golden.py) get a softmax that is neither saturated nor uniform."""
    f = Function("ssd300_resnet34")
    init = _Init(seed)
    x = f.parameter((batch, 3, image, image), input_dtype, name="data")
    if np.dtype(input_dtype) != np.float32:
        x = f.convert(x, np.float32, name="data_f32")
    y = _conv_bn(f, init, x, width, 7, 2, 3, "relu", "conv1")
    y = f.max_pool(y, (3, 3), (2, 2), (1, 1), (1, 1), name="pool1")
    cin = width
    for si, nb in enumerate(trunk_blocks):
        cout = width * (2 ** si)
        for bi in range(nb):
            stride = 2 if (bi == 0 and si == 1) else 1  # conv4_x keeps stride 1 (38x38 map)
            nm = "r34_%d_%d" % (si + 2, bi)
            sc = y
            if bi == 0 and (stride != 1 or cin != cout):
                sc = _conv_bn(f, init, y, cout, 1, stride, 0, None, nm + "_sc")
            z = _conv_bn(f, init, y, cout, 3, stride, 1, "relu", nm + "_a")
            z = _conv_bn(f, init, z, cout, 3, 1, 1, None, nm + "_b")
            y = f.relu(f.add(z, sc, name=nm + "_add"), name=nm + "_out")
            cin = cout
    feats = [y]
    extra = [(256, 512, 2, 1), (256, 512, 2, 1), (128, 256, 2, 1), (128, 256, 1, 0), (128, 256, 1, 0)]
    sc = width / 64.0
    for ei, (m, o, s, p) in enumerate(extra):
        if feats[-1].shape[2] < 3 and p == 0:
            break
        z = f.relu(f.add(f.convolution(feats[-1], init.conv(int(m * sc), feats[-1].shape[1], 1, 1), name="extra%d_a" % ei),
                         init.vec(int(m * sc)).reshape(1, -1, 1, 1), name="extra%d_a_bias" % ei), name="extra%d_a_relu" % ei)
        z = f.relu(f.add(f.convolution(z, init.conv(int(o * sc), int(m * sc), 3, 3), (s, s), (p, p), (p, p), name="extra%d_b" % ei),
                         init.vec(int(o * sc)).reshape(1, -1, 1, 1), name="extra%d_b_bias" % ei), name="extra%d_b_relu" % ei)
        feats.append(z)
    anchors = [4, 6, 6, 6, 4, 4][: len(feats)]
    locs, confs = [], []
    for hi, (ft, na) in enumerate(zip(feats, anchors)):
        c = ft.shape[1]
        for kind, ch, lst in (("loc", na * 4, locs), ("conf", na * num_classes, confs)):
            hsc = head_scale if (kind == "loc" or conf_scale is None) else conf_scale
            h = f.add(f.convolution(ft, init.conv(ch, c, 3, 3) * hsc, (1, 1), (1, 1), (1, 1), name="head%d_%s" % (hi, kind)),
                      init.vec(ch).reshape(1, -1, 1, 1), name="head%d_%s_bias" % (hi, kind))
            h = f.transpose(h, (0, 2, 3, 1), name="head%d_%s_perm" % (hi, kind))
            lst.append(f.reshape(h, (batch, -1), name="head%d_%s_flat" % (hi, kind)))
    loc = f.concat(locs, 1, name="loc_concat")
    conf = f.concat(confs, 1, name="conf_concat")
    num_priors = loc.shape[1] // 4
    conf = f.reshape(conf, (batch, num_priors, num_classes), name="conf_3d")
    conf = f.softmax(conf, 2, name="conf_softmax")
    conf = f.reshape(conf, (batch, num_priors * num_classes), name="conf_flat")
    priors = _ssd_priors([ft.shape[2] for ft in feats], image, anchors, init.rng)
    assert priors.shape[2] == num_priors * 4, (priors.shape, num_priors)
    det = f.detection_output(loc, conf, priors, num_classes, 0, top_k, keep_top_k, "caffe.PriorBoxParameter.CENTER_SIZE", True, 0.45, 0.01,
                             name="detection_out")
    f.result(det)
    return f
