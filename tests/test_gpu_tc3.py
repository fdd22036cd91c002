"""CTA-pair tcgen05 kernel (dldt_b200/csrc/conv_tc3.cuh) against the numpy oracle, through the C ABI.

At the sizes the oracle finishes in seconds the static policy would pick the single-CTA kernel (too few tiles for 74 SM pairs), so
B200_TC3=1 forces the pair kernel wherever it is supported; the exec graph is checked to make sure it is the kernel that ran.
Covers the three A-operand modes (activation matrix, im2col TMA, 3x3 halo box), weight-stationary and streamed weights, residual
+ second activation, ragged M tails, N tiles of 64/128/256, bf16 and fp16, and the load-time autotuner's bookkeeping."""
import numpy as np
import pytest

from dldt_b200 import Core, Function
from oracle import interp
from tests.util import rel_l2

pytestmark = pytest.mark.gpu

# name: (N, C, H, W, O, k, stride, pad, residual, act, precision, extra env, expected mode)
CASES = {
    "mat_ws_bn256_tail": (4, 64, 28, 28, 256, 1, 1, 0, False, "relu", "BF16", {}, "tma"),
    "mat_ws_bn64": (4, 256, 28, 28, 64, 1, 1, 0, False, "relu", "BF16", {}, "tma"),
    "mat_res_ws_bn256": (3, 256, 20, 20, 256, 1, 1, 0, True, "relu", "BF16", {}, "tma"),
    "mat_nows_bn256_x2": (2, 512, 14, 14, 512, 1, 1, 0, False, "relu", "BF16", {}, "tma"),
    "mat_bn128_nows": (2, 128, 14, 14, 128, 1, 1, 0, False, "relu", "BF16", {"B200_NO_WS": "1"}, "tma"),
    "mat_res_nows": (2, 512, 14, 14, 512, 1, 1, 0, True, "relu", "BF16", {}, "tma"),
    "mat_ws_two_n_tiles": (4, 64, 28, 28, 512, 1, 1, 0, False, "relu", "BF16", {}, "tma"),
    "mat_ws_four_n_tiles_res": (3, 256, 20, 20, 256, 1, 1, 0, True, "relu", "BF16", {"B200_TC3_FORCE_BN": "64"}, "tma"),
    "mat_swish_fp16": (2, 64, 16, 16, 192, 1, 1, 0, False, "swish", "FP16", {}, "tma"),
    "mat_k2048_n1000": (3, 2048, 7, 7, 1000, 1, 1, 0, False, "none", "BF16", {}, "tma"),
    "im2col_s2": (2, 128, 30, 30, 256, 3, 2, 1, False, "relu", "BF16", {}, "im2col"),
    "im2col_3x3_7x7": (5, 512, 7, 7, 512, 3, 1, 1, False, "relu", "BF16", {}, "im2col"),
    "im2col_nohalo_56": (2, 64, 56, 56, 64, 3, 1, 1, False, "relu", "BF16", {"B200_TC3_HALO": "0"}, "im2col"),
    "halo_56_ws": (2, 64, 56, 56, 64, 3, 1, 1, False, "relu", "BF16", {}, "halo"),
    "halo_56_c128_o256": (1, 128, 56, 56, 256, 3, 1, 1, False, "relu", "BF16", {}, "halo"),
    "halo_28": (3, 128, 28, 28, 128, 3, 1, 1, False, "relu", "BF16", {}, "halo"),
    "halo_28_res_nows": (3, 128, 28, 28, 128, 3, 1, 1, True, "relu", "BF16", {"B200_NO_WS": "1"}, "halo"),
    "halo_28_res_fp16": (3, 128, 28, 28, 128, 3, 1, 1, True, "relu", "FP16", {}, "halo"),
    "halo_14_partial_tile": (5, 256, 14, 14, 256, 3, 1, 1, False, "relu", "BF16", {"B200_TC3_HALO": "2"}, "halo"),
    "halo_14_partial_res": (3, 128, 14, 14, 128, 3, 1, 1, True, "relu", "BF16", {"B200_NO_WS": "1", "B200_TC3_HALO": "2"}, "halo"),
    "s2d_stem_64": (2, 3, 64, 64, 64, 7, 2, 3, False, "relu", "BF16", {}, "s2d_s2d_stem"),
    "s2d_stem_w124_widest": (1, 3, 40, 248, 64, 7, 2, 3, False, "relu", "BF16", {}, "s2d_s2d_stem"),
    "s2d_stem_too_wide_falls_back": (1, 3, 40, 520, 64, 7, 2, 3, False, "relu", "BF16", {}, "im2col_s2d_stem"),
    "s2d_stem_fp16_partial_rows": (3, 3, 50, 40, 64, 7, 2, 3, False, "relu", "FP16", {}, "s2d_s2d_stem"),
    "s2d_stem_ssd300_falls_back": (1, 3, 300, 300, 64, 7, 2, 3, False, "relu", "BF16", {}, "im2col_s2d_stem"),
    "s2d_stem_3x3_c32_relu6": (2, 3, 64, 64, 32, 3, 2, 1, False, "clamp", "BF16", {}, "s2d_s2d_stem"),      # MobileNet's first convolution
    "s2d_stem_5x5_c48_fp16": (3, 3, 48, 80, 48, 5, 2, 2, False, "relu", "FP16", {}, "s2d_s2d_stem"),
    "halo_w30_clip": (2, 64, 16, 30, 128, 3, 1, 1, False, "clamp", "BF16", {}, "halo"),
}


def build(name):
    n, c, h, w, o, k, s, p, res, act, prec, env, mode = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    f = Function("conv")
    x = f.parameter((n, c, h, w))
    y = f.convolution(x, (rng.standard_normal((o, c, k, k)) / np.sqrt(c * k * k)).astype(np.float32), (s, s), (p, p), (p, p), (1, 1))
    y = f.add(y, rng.standard_normal((1, o, 1, 1)).astype(np.float32))
    if res:
        y = f.relu(f.add(y, x))
    elif act == "relu":
        y = f.relu(y)
    elif act == "swish":
        y = f.swish(y)
    elif act == "clamp":
        y = f.clamp(y, 0.0, 6.0)
    f.result(y)
    return f, rng.standard_normal((n, c, h, w)).astype(np.float32)


@pytest.mark.parametrize("name", sorted(CASES))
def test_pair_kernel_matches_oracle(name, monkeypatch):
    prec, env, mode = CASES[name][10], CASES[name][11], CASES[name][12]
    monkeypatch.setenv("B200_TC3", "1")
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    f, xin = build(name)
    ref = list(interp.run(f, {"Parameter_0": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
    got = list(net.create_infer_request().infer({"Parameter_0": xin}).values())[0]
    impl = [e["impl"] for e in net.get_exec_graph_info() if e["type"] == "Convolution"]
    assert impl and all("pair256" in i and i.endswith(mode) for i in impl), impl
    tol = 1e-2 if prec == "BF16" else 2e-3      # north star: bf16/fp16 within 1e-2 relative L2
    assert rel_l2(got, ref) < tol, (name, rel_l2(got, ref))


# name: (N, H, W, precision, pads_end, fused?)
STEM_POOL = {
    "pool_64": (2, 64, 64, "BF16", (1, 1), True),
    "pool_odd_tiles_fp16": (3, 48, 80, "FP16", (1, 1), True),          # 3 x 6 CTA tiles: the last pair tile has one valid half
    "pool_widest_pads_end0": (1, 40, 248, "BF16", (0, 0), True),       # 124 output pixels = 31 pixel groups, 62 pooled columns
    "pool_rows_not_x4_stays_separate": (3, 50, 40, "BF16", (1, 1), False),
}


@pytest.mark.parametrize("name", sorted(STEM_POOL))
def test_stem_with_max_pool_in_the_epilogue(name, monkeypatch):
    # ResNet's conv1 -> ReLU -> MaxPool 3x3 / 2 / pad 1: the pool is taken in the stem's epilogue (TMA max-reduce into a zero-filled tensor)
    n, h, w, prec, pe, fused = STEM_POOL[name]
    monkeypatch.setenv("B200_TC3", "1")
    rng = np.random.default_rng(sum(map(ord, name)))
    f = Function("stem")
    x = f.parameter((n, 3, h, w))
    y = f.convolution(x, (rng.standard_normal((64, 3, 7, 7)) / np.sqrt(147.0)).astype(np.float32), (2, 2), (3, 3), (3, 3), (1, 1))
    y = f.relu(f.add(y, rng.standard_normal((1, 64, 1, 1)).astype(np.float32)))
    f.result(f.max_pool(y, (3, 3), (2, 2), (1, 1), pe))
    xin = rng.standard_normal((n, 3, h, w)).astype(np.float32)
    ref = list(interp.run(f, {"Parameter_0": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
    req = net.create_infer_request()
    got = list(req.infer({"Parameter_0": xin}).values())[0]
    eg = net.get_exec_graph_info()
    assert any(e["impl"].endswith("_s2d_stem_maxpool") for e in eg) == fused, [e["impl"] for e in eg]
    assert any(e["type"] == "Pooling" for e in eg) == (not fused)
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < (1e-2 if prec == "BF16" else 2e-3), (name, rel_l2(got, ref))
    # the pooled tensor is zero-filled by every run: a second inference gives the same bits
    assert np.array_equal(got, list(req.infer({"Parameter_0": xin}).values())[0])


def test_autotuned_selection_is_one_of_the_candidates_and_matches_oracle(monkeypatch):
    # large enough that both persistent kernels apply (>= 74 pair tiles): the first request times them and keeps one
    monkeypatch.delenv("B200_TC3", raising=False)
    rng = np.random.default_rng(5)
    f = Function("conv")
    x = f.parameter((24, 64, 28, 28))
    y = f.relu(f.add(f.convolution(x, (rng.standard_normal((64, 64, 3, 3)) / 24.0).astype(np.float32), (1, 1), (1, 1), (1, 1), (1, 1)),
                     rng.standard_normal((1, 64, 1, 1)).astype(np.float32)))
    f.result(y)
    xin = rng.standard_normal((24, 64, 28, 28)).astype(np.float32)
    ref = list(interp.run(f, {"Parameter_0": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    req = net.create_infer_request()
    got = list(req.infer({"Parameter_0": xin}).values())[0]
    impl = [e["impl"] for e in net.get_exec_graph_info() if e["type"] == "Convolution"][0]
    assert "pair256" in impl or "128x" in impl, impl
    assert rel_l2(got, ref) < 1e-2
    # a second request of the same network reuses the decision and gives the same bits
    got2 = list(net.create_infer_request().infer({"Parameter_0": xin}).values())[0]
    assert np.array_equal(got, got2)


def _bottleneck(f, x, cin, mid, cout, stride, rng, bn_on_shortcut):
    def conv(v, o, c, k, s, p, name):
        w = (rng.standard_normal((o, c, k, k)) / np.sqrt(c * k * k)).astype(np.float32)
        return f.add(f.convolution(v, w, (s, s), (p, p), (p, p), (1, 1)), rng.standard_normal((1, o, 1, 1)).astype(np.float32) * 0.1)
    a = f.relu(conv(x, mid, cin, 1, 1, 0, "a"))
    b = f.relu(conv(a, mid, mid, 3, stride, 1, "b"))
    c = conv(b, cout, mid, 1, 1, 0, "c")
    sc = conv(x, cout, cin, 1, stride, 0, "proj")
    if bn_on_shortcut:   # a per-channel scale after the shortcut conv is folded into its weights before the K concatenation
        sc = f.multiply(sc, rng.uniform(0.5, 1.5, (1, cout, 1, 1)).astype(np.float32))
    return f.relu(f.add(c, sc))


@pytest.mark.parametrize("cin,mid,cout,stride,hw,bn", [(64, 64, 256, 1, 16, False), (128, 64, 256, 2, 16, True), (256, 128, 512, 2, 14, False)])
def test_projection_shortcut_is_folded_into_the_last_conv(cin, mid, cout, stride, hw, bn, monkeypatch):
    # ResNet downsampling block: conv_c(b) + conv_proj(x) runs as ONE GEMM over concatenated K (compile.cu fold_shortcuts)
    monkeypatch.delenv("B200_TC3", raising=False)
    monkeypatch.delenv("B200_FUSE_NEXT", raising=False)
    rng = np.random.default_rng(cin + stride)
    f = Function("block0")
    x = f.parameter((3, cin, hw, hw))
    f.result(_bottleneck(f, x, cin, mid, cout, stride, rng, bn))
    xin = rng.standard_normal((3, cin, hw, hw)).astype(np.float32)
    ref = list(interp.run(f, {"Parameter_0": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    got = list(net.create_infer_request().infer({"Parameter_0": xin}).values())[0]
    convs = [e["impl"] for e in net.get_exec_graph_info() if e["type"] == "Convolution"]
    assert len(convs) == 3 and convs[-1].endswith("tma_kcat"), convs
    assert rel_l2(got, ref) < 1e-2
    # and with the folding off the same graph gives four kernels and the same answer within bf16 rounding of the shortcut tensor
    monkeypatch.setenv("B200_FOLD_SHORTCUT", "0")
    net2 = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    got2 = list(net2.create_infer_request().infer({"Parameter_0": xin}).values())[0]
    assert len([e for e in net2.get_exec_graph_info() if e["type"] == "Convolution"]) == 4
    assert rel_l2(got2, ref) < 1e-2 and rel_l2(got, got2) < 1e-2


def _identity_block(f, x, c, mid, rng):
    def conv(v, o, ci, k, p):
        w = (rng.standard_normal((o, ci, k, k)) / np.sqrt(ci * k * k)).astype(np.float32)
        return f.add(f.convolution(v, w, (1, 1), (p, p), (p, p), (1, 1)), rng.standard_normal((1, o, 1, 1)).astype(np.float32) * 0.1)
    a = f.relu(conv(x, mid, c, 1, 0))
    b = f.relu(conv(a, mid, mid, 3, 1))
    return f.relu(f.add(conv(b, c, mid, 1, 0), x))


# the last case gives some of the 74 clusters two M tiles (the Z accumulator is then drained one tile late, while the next
# M tile's first Y slabs wait in the staging ring)
@pytest.mark.parametrize("cin,mid,cout,stride,hw,n", [(64, 64, 256, 1, 16, 3), (256, 128, 512, 2, 14, 5), (512, 256, 1024, 2, 12, 2),
                                                      (64, 64, 256, 1, 56, 9)])
def test_next_reduce_conv_rides_on_the_expand_conv(cin, mid, cout, stride, hw, n, monkeypatch):
    # block 0 (projection shortcut folded as extra K) -> identity block -> identity block: the 1x1 reduce conv of each following block is
    # computed inside the kernel that produces its input (compile.cu fold_next_reduce, conv_tc3f.cu); odd M tails, 1-8 N tiles, Ca 64-256
    monkeypatch.delenv("B200_TC3", raising=False)
    monkeypatch.setenv("B200_FUSE_NEXT", "1")        # opt-in: conv_tc3f
    rng = np.random.default_rng(cin * 3 + stride)
    f = Function("blocks")
    x = f.parameter((n, cin, hw, hw))
    y = _bottleneck(f, x, cin, mid, cout, stride, rng, False)
    y = _identity_block(f, y, cout, mid, rng)
    y = _identity_block(f, y, cout, mid, rng)
    f.result(y)
    xin = rng.standard_normal((n, cin, hw, hw)).astype(np.float32)
    ref = list(interp.run(f, {"Parameter_0": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    got = list(net.create_infer_request().infer({"Parameter_0": xin}).values())[0]
    convs = [e["impl"] for e in net.get_exec_graph_info() if e["type"] == "Convolution"]
    # a, b, out0(+a1) | b1, out1(+a2) | b2, out2
    assert len(convs) == 7 and sum("+1x1x%d" % mid in c for c in convs) == 2 and "kcat+1x1" in convs[2], convs
    assert rel_l2(got, ref) < 1e-2          # three bottlenecks deep in bf16
    monkeypatch.setenv("B200_FUSE_NEXT", "0")
    net2 = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    got2 = list(net2.create_infer_request().infer({"Parameter_0": xin}).values())[0]
    assert len([e for e in net2.get_exec_graph_info() if e["type"] == "Convolution"]) == 9
    assert rel_l2(got, got2) < 1e-2


@pytest.mark.parametrize("cin,cc,ca", [(512, 1024, 256), (128, 512, 128)])
def test_chained_pointwise_convs_with_several_m_tiles_per_cluster(cin, cc, ca, monkeypatch):
    # 20 480 rows = 80 pair M tiles on 74 clusters, Ca = 256 -> two Z slabs per epilogue group behind a pending Y slab of the next M tile
    monkeypatch.delenv("B200_TC3", raising=False)
    monkeypatch.setenv("B200_FUSE_NEXT", "1")        # opt-in: conv_tc3f
    rng = np.random.default_rng(cc)
    f = Function("chain")
    x = f.parameter((20, cin, 32, 32))
    y = f.relu(f.add(f.convolution(x, (rng.standard_normal((cc, cin, 1, 1)) / np.sqrt(cin)).astype(np.float32), (1, 1), (0, 0), (0, 0), (1, 1)),
                     rng.standard_normal((1, cc, 1, 1)).astype(np.float32) * 0.1))
    z = f.relu(f.add(f.convolution(y, (rng.standard_normal((ca, cc, 1, 1)) / np.sqrt(cc)).astype(np.float32), (1, 1), (0, 0), (0, 0), (1, 1)),
                     rng.standard_normal((1, ca, 1, 1)).astype(np.float32) * 0.1))
    f.result(z)
    xin = rng.standard_normal((20, cin, 32, 32)).astype(np.float32)
    ref = list(interp.run(f, {"Parameter_0": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    got = list(net.create_infer_request().infer({"Parameter_0": xin}).values())[0]
    convs = [e["impl"] for e in net.get_exec_graph_info() if e["type"] == "Convolution"]
    assert len(convs) == 1 and convs[0].endswith("+1x1x%d" % ca), convs
    assert rel_l2(got, ref) < 1e-2


@pytest.mark.parametrize("prec,tol", [("BF16", 1e-2), ("FP16", 3e-3)])
@pytest.mark.parametrize("n,seq,heads,k", [(2, 64, 3, 128), (3, 384, 12, 768), (1, 32, 2, 64)])
def test_head_split_is_written_by_the_fc_stores(prec, tol, n, seq, heads, k, monkeypatch):
    """FullyConnected -> Reshape [N,S,H,64] -> Transpose (0,2,1,3): the pair kernel's 32-token x one-head store boxes land at
    [n][h][s][:], no permute kernel (B200_FOLD_HEADS=0 keeps the three-step form, same numbers)"""
    rng = np.random.default_rng(seq + heads)
    f = Function("heads")
    x = f.parameter((n, seq, k), name="x")
    w = (rng.standard_normal((heads * 64, k)) / np.sqrt(k)).astype(np.float32)
    b = rng.uniform(-0.5, 0.5, (heads * 64,)).astype(np.float32)
    y = f.add(f.matmul(x, w, transpose_b=True), f.constant(b))
    f.result(f.transpose(f.reshape(y, (n, seq, heads, 64)), (0, 2, 1, 3)))
    xin = rng.standard_normal((n, seq, k)).astype(np.float32)
    ref = list(interp.run(f, {"x": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
    got = list(net.create_infer_request().infer({"x": xin}).values())[0]
    eg = net.get_exec_graph_info()
    assert [e["type"] for e in eg if e["type"] in ("FullyConnected", "Transpose")] == ["FullyConnected"], eg
    assert "tma_heads" in [e["impl"] for e in eg if e["type"] == "FullyConnected"][0]
    assert got.shape == (n, heads, seq, 64) and rel_l2(got, ref) < tol
    monkeypatch.setenv("B200_FOLD_HEADS", "0")
    net0 = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
    got0 = list(net0.create_infer_request().infer({"x": xin}).values())[0]
    assert any(e["type"] == "Transpose" for e in net0.get_exec_graph_info())
    assert rel_l2(got, got0) < 2e-3
