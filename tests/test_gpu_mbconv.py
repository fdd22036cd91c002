"""Fused inverted-residual block kernel (dldt_b200/csrc/mbconv.cu: expand 1x1 -> depthwise 3x3 -> project 1x1 (+ block input) in one tcgen05
kernel) against the numpy oracle, through the C ABI.  Covers stride 1 / 2, residual, 1-3 input k-blocks, padded hidden widths, partial
tiles at the image edge, several tiles per CTA (cross-tile barrier phases), the expand-less first block, asymmetric stride-2 padding and
fp16; the exec graph is checked so that the fused kernel is what ran."""
import numpy as np
import pytest

from dldt_b200 import Core, Function
from oracle import interp
from tests.util import rel_l2

pytestmark = pytest.mark.gpu

# name: (N, Cin, expansion, Cout, H, W, stride, pads_begin, pads_end, precision)
CASES = {
    "s1_res_c24_hidden144_padded": (2, 24, 6, 24, 28, 28, 1, (1, 1), (1, 1), "BF16"),
    "s1_res_partial_tiles_14": (3, 32, 6, 32, 14, 14, 1, (1, 1), (1, 1), "BF16"),
    "s1_c64_to_96_no_res": (2, 64, 6, 96, 14, 14, 1, (1, 1), (1, 1), "BF16"),
    "s1_res_two_k_blocks_c96": (2, 96, 6, 96, 14, 14, 1, (1, 1), (1, 1), "BF16"),
    "s1_res_many_tiles_per_cta": (20, 32, 6, 32, 28, 28, 1, (1, 1), (1, 1), "BF16"),
    "s1_res_56_tiles_224": (8, 24, 6, 24, 56, 56, 1, (1, 1), (1, 1), "BF16"),
    "s1_res_five_tiles_per_cta": (24, 24, 6, 24, 56, 56, 1, (1, 1), (1, 1), "BF16"),      # 672 tiles on 148 CTAs: X-tile ring wraps, deferred epilogues
    "s1_noexpand_many_tiles": (16, 32, 1, 16, 56, 56, 1, (1, 1), (1, 1), "BF16"),         # 448 tiles
    "s2_three_tiles_per_cta": (16, 16, 6, 24, 64, 64, 2, (1, 1), (1, 1), "BF16"),         # 512 tiles
    "s2_c16_to_24": (2, 16, 6, 24, 32, 32, 2, (1, 1), (1, 1), "BF16"),
    "s2_odd_size_many_tiles": (12, 24, 6, 32, 57, 57, 2, (1, 1), (1, 1), "BF16"),
    "s2_pads_begin0_end1": (2, 24, 6, 32, 30, 30, 2, (0, 0), (1, 1), "BF16"),
    "s1_noexpand_c32_to_16": (3, 32, 1, 16, 32, 32, 1, (1, 1), (1, 1), "BF16"),
    "s2_noexpand_c64": (2, 64, 1, 24, 28, 28, 2, (1, 1), (1, 1), "BF16"),
    "s1_res_fp16": (2, 32, 6, 32, 20, 24, 1, (1, 1), (1, 1), "FP16"),
    "s2_fp16": (3, 32, 4, 64, 28, 28, 2, (1, 1), (1, 1), "FP16"),
}


def build(name):
    n, cin, t, cout, h, w, stride, pb, pe, prec = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    f = Function("block")
    x = f.parameter((n, cin, h, w))
    hidden = cin * t
    z = x
    if t != 1:
        z = f.convolution(z, (rng.standard_normal((hidden, cin, 1, 1)) * np.sqrt(2.0 / cin)).astype(np.float32), (1, 1), (0, 0), (0, 0), (1, 1))
        z = f.clamp(f.add(z, rng.standard_normal((1, hidden, 1, 1)).astype(np.float32) * 0.2), 0.0, 6.0)
    z = f.group_convolution(z, (rng.standard_normal((hidden, 1, 1, 3, 3)) * np.sqrt(2.0 / 9)).astype(np.float32), (stride, stride), pb, pe)
    z = f.clamp(f.add(z, rng.standard_normal((1, hidden, 1, 1)).astype(np.float32) * 0.2), 0.0, 6.0)
    z = f.convolution(z, (rng.standard_normal((cout, hidden, 1, 1)) / np.sqrt(hidden)).astype(np.float32), (1, 1), (0, 0), (0, 0), (1, 1))
    z = f.add(z, rng.standard_normal((1, cout, 1, 1)).astype(np.float32) * 0.2)
    if stride == 1 and cin == cout:
        z = f.add(x, z)
    f.result(z)
    return f, rng.standard_normal((n, cin, h, w)).astype(np.float32)


@pytest.mark.parametrize("name", sorted(CASES))
def test_fused_block_matches_oracle(name):
    n, cin, t, cout, h, w, stride, pb, pe, prec = CASES[name]
    f, xin = build(name)
    ref = list(interp.run(f, {"Parameter_0": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
    req = net.create_infer_request()
    got = list(req.infer({"Parameter_0": xin}).values())[0]
    impl = [e["impl"] for e in net.get_exec_graph_info() if e["type"] == "Convolution"]
    want = "_mbconv_s%d%s%s" % (stride, "_noexpand" if t == 1 else "", "_res" if (stride == 1 and cin == cout and t != 1) else "")
    assert len(impl) == 1 and impl[0].endswith(want), impl
    assert got.shape == ref.shape
    assert rel_l2(got, ref) < (1e-2 if prec == "BF16" else 2e-3), (name, rel_l2(got, ref))
    assert np.array_equal(got, list(req.infer({"Parameter_0": xin}).values())[0])      # a second run gives the same bits


def test_fusion_can_be_turned_off(monkeypatch):
    monkeypatch.setenv("B200_FUSE_MBCONV", "0")
    f, xin = build("s1_res_partial_tiles_14")
    ref = list(interp.run(f, {"Parameter_0": xin}).values())[0]
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": "BF16"})
    got = list(net.create_infer_request().infer({"Parameter_0": xin}).values())[0]
    assert not any("mbconv" in e["impl"] for e in net.get_exec_graph_info())
    assert rel_l2(got, ref) < 1e-2
