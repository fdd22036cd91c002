"""Fast paths of the memory-bound kernels (kernels_misc.cu) at the shapes where they apply, against the numpy oracle through the C ABI:
fused scale + mask + softmax (BERT attention), plain warp-per-row softmax, warp-per-row MVN with gamma/beta, 16-byte permute, depthwise
3x3 stride 1/2 with odd widths, packed 3x3/s2 max-pool, warp-per-channel-group global average pool.  The impl strings of the exec graph
are checked so that a silent fall back to the generic kernels fails the test."""
import numpy as np
import pytest

from dldt_b200 import Core, Function
from oracle import interp
from tests.util import rel_l2

pytestmark = pytest.mark.gpu


def run(f, inputs, prec):
    ref = interp.run(f, inputs)
    net = Core().load_network(f, "B200", {"B200_INFERENCE_PRECISION": prec})
    got = net.create_infer_request().infer(inputs)
    return ref, got, net.get_exec_graph_info()


@pytest.mark.parametrize("prec,tol", [("FP32", 1e-5), ("BF16", 1e-2)])
@pytest.mark.parametrize("seq,mask_batch,heads", [(384, True, 3), (128, False, 3), (40, True, 3), (384, True, 12), (200, False, 24)])
def test_attention_softmax_is_one_kernel(prec, tol, seq, mask_batch, heads):
    # heads = 12 / 24: >= 8192 rows, where 16-bit rows go four to a warp (softmax_warp_multirow_kernel)
    rng = np.random.default_rng(seq)
    n = 2
    f = Function("attn")
    x = f.parameter((n, heads, seq, seq), name="scores")
    m = f.parameter((n if mask_batch else 1, 1, 1, seq), name="mask")
    y = f.softmax(f.add(f.multiply(x, f.constant(np.float32(0.125))), m), 3)
    f.result(y)
    xin = rng.standard_normal((n, heads, seq, seq)).astype(np.float32) * 4
    mask = np.where(rng.uniform(size=(n if mask_batch else 1, 1, 1, seq)) < 0.2, -10000.0, 0.0).astype(np.float32)
    ref, got, eg = run(f, {"scores": xin, "mask": mask}, prec)
    assert [e["impl"] for e in eg if e["type"] in ("Softmax", "Eltwise")] == ["softmax_scale_mask_warp_rows"], eg
    a, b = list(got.values())[0], list(ref.values())[0]
    assert rel_l2(a, b) < tol
    assert np.allclose(a.sum(-1), 1.0, atol=2e-2 if prec == "BF16" else 1e-5)


@pytest.mark.parametrize("prec,tol", [("FP32", 1e-5), ("BF16", 1e-2)])
def test_plain_softmax_and_layernorm_rows(prec, tol):
    rng = np.random.default_rng(11)
    f = Function("ln")
    x = f.parameter((6, 50, 768), name="x")
    g = rng.uniform(0.5, 1.5, (768,)).astype(np.float32)
    b = rng.uniform(-0.2, 0.2, (768,)).astype(np.float32)
    y = f.add(f.multiply(f.mvn(x, [2], True, 1e-12, "INSIDE_SQRT"), g), b)
    f.result(f.softmax(y, 2))
    xin = rng.standard_normal((6, 50, 768)).astype(np.float32) * 3 + 1
    ref, got, eg = run(f, {"x": xin}, prec)
    impls = [e["impl"] for e in eg]
    assert "softmax_warp_rows" in impls, impls
    assert rel_l2(list(got.values())[0], list(ref.values())[0]) < tol


@pytest.mark.parametrize("prec", ["FP32", "BF16"])
def test_head_transpose_moves_16_byte_vectors(prec):
    rng = np.random.default_rng(5)
    f = Function("perm")
    x = f.parameter((3, 40, 12, 64), name="x")
    f.result(f.transpose(f.relu(x), (0, 2, 1, 3)))
    xin = rng.standard_normal((3, 40, 12, 64)).astype(np.float32)
    ref, got, eg = run(f, {"x": xin}, prec)
    a, b = list(got.values())[0], list(ref.values())[0]
    assert a.shape == (3, 12, 40, 64) and rel_l2(a, b) < (1e-6 if prec == "FP32" else 5e-3)


@pytest.mark.parametrize("prec", ["BF16", "FP16"])
@pytest.mark.parametrize("stride,h,w,c,pad", [(1, 20, 23, 48, 1), (2, 21, 30, 48, 1), (2, 56, 56, 48, 1), (1, 7, 7, 40, 1), (1, 9, 5, 24, 0), (2, 14, 14, 96, 0),
                                              (1, 3, 3, 8, 1), (2, 5, 6, 20, 1)])
def test_depthwise_3x3_fast_path(stride, h, w, c, pad, prec):
    # strips of 4 / 2 output rows, two output columns per thread: odd widths, heights that are not a multiple of the strip, channel counts
    # that are not a multiple of 8 (padded lanes), no padding, maps smaller than one strip
    rng = np.random.default_rng(stride * 100 + w)
    f = Function("dw")
    x = f.parameter((2, c, h, w), name="x")
    wdw = (rng.standard_normal((c, 1, 1, 3, 3)) * 0.4).astype(np.float32)
    y = f.group_convolution(x, wdw, (stride, stride), (pad, pad), (pad, pad), (1, 1))
    y = f.clamp(f.add(y, rng.standard_normal((1, c, 1, 1)).astype(np.float32) * 0.1), 0.0, 6.0)
    f.result(y)
    xin = rng.standard_normal((2, c, h, w)).astype(np.float32)
    ref, got, eg = run(f, {"x": xin}, prec)
    assert any(e["impl"].startswith("dw_") for e in eg)
    assert rel_l2(list(got.values())[0], list(ref.values())[0]) < (1e-2 if prec == "BF16" else 2e-3)


@pytest.mark.parametrize("cin,cout,k,stride,pad,dil,h,w", [(3, 32, 3, 2, 1, 1, 64, 64), (3, 32, 3, 2, 1, 1, 33, 47), (1, 16, 5, 1, 2, 1, 20, 20), (4, 24, 3, 1, 1, 2, 17, 19),
                                                        (2, 64, 1, 1, 0, 1, 9, 9)])
def test_first_convolution_over_few_input_channels(cin, cout, k, stride, pad, dil, h, w):
    """Dense convolution over <= 4 input channels (MobileNet's C = 3 first layer) runs as the direct CUDA-core kernel, BN and ReLU6 folded."""
    rng = np.random.default_rng(cin * 1000 + cout)
    f = Function("first")
    x = f.parameter((2, cin, h, w), name="x")
    y = f.convolution(x, (rng.standard_normal((cout, cin, k, k)) / np.sqrt(cin * k * k)).astype(np.float32), (stride, stride), (pad, pad), (pad, pad), (dil, dil))
    y = f.clamp(f.add(y, rng.standard_normal((1, cout, 1, 1)).astype(np.float32) * 0.1), 0.0, 6.0)
    f.result(y)
    xin = rng.uniform(0, 1, (2, cin, h, w)).astype(np.float32)
    for prec, tol in (("BF16", 1e-2), ("FP16", 2e-3)):
        ref, got, eg = run(f, {"x": xin}, prec)
        assert [e["impl"] for e in eg if e["type"] == "Convolution"] == ["direct_small_cin_cuda_core"]
        assert rel_l2(list(got.values())[0], list(ref.values())[0]) < tol


@pytest.mark.parametrize("h,w", [(28, 28), (15, 21), (7, 7)])
def test_maxpool_3x3_s2_and_global_average_pool(h, w):
    rng = np.random.default_rng(h * w)
    c = 64
    f = Function("pools")
    x = f.parameter((3, c, h, w), name="x")
    p = f.max_pool(f.relu(x), (3, 3), (2, 2), (1, 1), (1, 1))
    gmean = f.reduce_mean(x, [2, 3], keep_dims=True)
    f.result(p)
    f.result(gmean)
    xin = rng.standard_normal((3, c, h, w)).astype(np.float32)
    ref, got, eg = run(f, {"x": xin}, "BF16")
    for k in ref:
        assert got[k].shape == ref[k].shape and rel_l2(got[k], ref[k]) < 1e-2, k


@pytest.mark.parametrize("shape", [(3, 4, 40, 64), (2, 3, 384, 64), (2, 2, 70, 34), (1, 2, 37, 50)])
def test_swap_of_the_last_two_dims(shape):
    # [N, heads, S, 64] -> [N, heads, 64, S] (V^T of BERT's attention): tiled shared-memory transpose for even sizes, generic permute otherwise
    rng = np.random.default_rng(shape[2])
    f = Function("t")
    x = f.parameter(shape, name="x")
    f.result(f.transpose(f.relu(x), (0, 1, 3, 2)))
    xin = rng.standard_normal(shape).astype(np.float32)
    ref, got, eg = run(f, {"x": xin}, "BF16")
    a, b = list(got.values())[0], list(ref.values())[0]
    assert a.shape == b.shape and rel_l2(a, b) < 5e-3


@pytest.mark.parametrize("prec,tol", [("BF16", 1e-2), ("FP16", 4e-3)])
@pytest.mark.parametrize("seq,dh", [(128, 64), (384, 64), (256, 128)])
def test_attention_matmuls_run_as_one_grouped_launch(prec, tol, seq, dh, monkeypatch):
    """scores = Q K^T and context = P V with tile-aligned sequence lengths: every (batch, head) pair is a weight group of one flat
    persistent launch (ConvParams::w_grp_m), not a blockIdx.z batch of single-tile CTAs; same result as the batched launch"""
    rng = np.random.default_rng(seq + dh)
    n, heads = 2, 3
    f = Function("attn_mm")
    q = f.parameter((n, heads, seq, dh), name="q")
    k = f.parameter((n, heads, seq, dh), name="k")
    v = f.parameter((n, heads, seq, dh), name="v")
    s = f.matmul(q, k, transpose_b=True)
    f.result(s)
    f.result(f.matmul(f.softmax(f.multiply(s, f.constant(np.float32(0.125))), 3), v))
    ins = {nm: rng.standard_normal((n, heads, seq, dh)).astype(np.float32) for nm in "qkv"}
    ref, got, eg = run(f, ins, prec)
    mm = [e["impl"] for e in eg if e["type"] == "MatMul"]
    assert len(mm) == 2 and all("persistent_tma" in i for i in mm), mm
    for a, b in zip(got.values(), ref.values()):
        assert rel_l2(a, b) < tol
    monkeypatch.setenv("B200_MATMUL_FLAT", "0")
    _, got0, eg0 = run(f, ins, prec)
    assert not any("persistent" in e["impl"] for e in eg0 if e["type"] == "MatMul")
    for a, b in zip(got.values(), got0.values()):
        assert rel_l2(a, b) < 2e-3


@pytest.mark.parametrize("prec,tol", [("BF16", 1e-2), ("FP16", 4e-3)])
@pytest.mark.parametrize("n,heads,seq", [(2, 3, 128), (2, 12, 384), (1, 2, 256)])
def test_context_matmul_writes_the_merged_heads_layout(prec, tol, n, heads, seq, monkeypatch):
    """context = P V followed by Transpose (0,2,1,3) + Reshape [N,S,H*64]: the grouped launch's 128-row stores land at [n][s][h][:]"""
    rng = np.random.default_rng(seq + heads)
    f = Function("ctx")
    pr = f.parameter((n, heads, seq, seq), name="p")
    v = f.parameter((n, heads, seq, 64), name="v")
    c = f.matmul(f.softmax(pr, 3), v)
    f.result(f.reshape(f.transpose(c, (0, 2, 1, 3)), (n, seq, heads * 64)))
    ins = {"p": rng.standard_normal((n, heads, seq, seq)).astype(np.float32), "v": rng.standard_normal((n, heads, seq, 64)).astype(np.float32)}
    ref, got, eg = run(f, ins, prec)
    mm = [e["impl"] for e in eg if e["type"] == "MatMul"]
    assert len(mm) == 1 and mm[0].endswith("_bmn_heads"), mm               # V read MN-major (no V^T), merged-heads stores
    assert len([e for e in eg if e["type"] == "Transpose"]) == 0, eg
    a, b = list(got.values())[0], list(ref.values())[0]
    assert a.shape == (n, seq, heads * 64) and rel_l2(a, b) < tol
    monkeypatch.setenv("B200_FOLD_HEADS", "0")
    _, got0, eg0 = run(f, ins, prec)
    assert len([e for e in eg0 if e["type"] == "Transpose"]) == 1          # the context transpose
    assert rel_l2(a, list(got0.values())[0]) < 2e-3
    monkeypatch.setenv("B200_MN_MAJOR", "0")
    _, got1, eg1 = run(f, ins, prec)
    assert len([e for e in eg1 if e["type"] == "Transpose"]) == 2          # + V^T
    assert rel_l2(a, list(got1.values())[0]) < 2e-3


@pytest.mark.parametrize("prec,tol", [("BF16", 1e-2), ("FP16", 2e-3)])
@pytest.mark.parametrize("n,heads,seq,mask_batch", [(2, 3, 384, True), (1, 2, 128, False), (3, 4, 256, True), (2, 12, 384, True)])
def test_fused_attention_kernel(prec, tol, n, heads, seq, mask_batch, monkeypatch):
    """softmax(Q K^T / 8 + mask) V + merged-heads transpose as ONE tcgen05 kernel (attention.cu): scores in TMEM, P in shared memory, V read
    MN-major.  Checked against the numpy restatement and against the three-kernel schedule of the same library; one padded sequence per
    batch exercises the -10000 mask (whole key ranges with probability 0)."""
    rng = np.random.default_rng(seq * heads + n)
    f = Function("attention")
    q = f.parameter((n, heads, seq, 64), name="q")
    k = f.parameter((n, heads, seq, 64), name="k")
    v = f.parameter((n, heads, seq, 64), name="v")
    m = f.parameter((n if mask_batch else 1, 1, 1, seq), name="mask")
    s = f.add(f.multiply(f.matmul(q, k, transpose_b=True), np.float32(0.125)), m)
    c = f.matmul(f.softmax(s, 3), v)
    f.result(f.reshape(f.transpose(c, (0, 2, 1, 3)), (n, seq, heads * 64)))
    mask = np.zeros((n if mask_batch else 1, 1, 1, seq), np.float32)
    mask[-1, 0, 0, seq - seq // 3:] = -10000.0
    ins = {"q": rng.standard_normal((n, heads, seq, 64)).astype(np.float32) * 1.5, "k": rng.standard_normal((n, heads, seq, 64)).astype(np.float32) * 1.5,
           "v": rng.standard_normal((n, heads, seq, 64)).astype(np.float32), "mask": mask}
    ref, got, eg = run(f, ins, prec)
    kinds = [e["type"] for e in eg]
    assert kinds.count("Attention") == 1 and "MatMul" not in kinds and "Softmax" not in kinds and "Transpose" not in kinds, eg
    a, b = list(got.values())[0], list(ref.values())[0]
    assert a.shape == (n, seq, heads * 64) and np.isfinite(a).all()
    assert rel_l2(a, b) < tol, rel_l2(a, b)
    monkeypatch.setenv("B200_FUSED_ATTENTION", "0")
    _, got0, eg0 = run(f, ins, prec)
    assert [e["type"] for e in eg0].count("MatMul") == 2
    # the fused kernel keeps the scores in fp32 until P is formed: it must not be further from the fp32 oracle than the unfused schedule
    assert rel_l2(a, b) <= rel_l2(list(got0.values())[0], b) * 1.25 + 1e-4
