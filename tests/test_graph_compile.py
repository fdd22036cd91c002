"""Graph compiler (b200_graph_add / b200_compile) without a device: shape inference, the fusion
legality rules mirrored from mkldnn_graph_optimizer.cpp, layout selection, exec-graph info, the
arena plan, and the host mirror's config / metric / error behaviour (ie_plugin_config.hpp)."""
from collections import Counter

import numpy as np
import pytest

from dldt_b200 import Core, Function, _capi, ie, zoo


def dry(f, precision="BF16", **cfg):
    c = {"B200_INFERENCE_PRECISION": precision}
    c.update(cfg)
    return Core().load_network(f, "B200", c, dry=True)


def kinds(net):
    return Counter(e["type"] for e in net.get_exec_graph_info())


def test_resnet50_schedule():
    net = dry(zoo.resnet50(batch=256))
    eg = net.get_exec_graph_info()
    k = kinds(net)
    # 53 convs with BN/ReLU/residual folded in, of which the 4 projection shortcuts ride on the last 1x1 of their block as extra K
    # (fold_shortcuts); the max pool rides in the stem's epilogue (max-reduce into a zero-filled tensor); global avg pool, FC (+bias)
    assert k == {"Fill": 1, "Convolution": 49, "Pooling": 1, "FullyConnected": 1}
    stem = [e for e in eg if e["impl"].endswith("_s2d_stem_maxpool")]
    assert len(stem) == 1 and any("pool" in n.lower() for n in stem[0]["fused"])
    assert sum(1 for e in eg if "tma_kcat" in e["impl"]) == 4
    assert all(e["impl"].startswith("tcgen05_bf16") for e in eg if e["type"] in ("Convolution", "FullyConnected"))
    assert sum(1 for e in eg if e["type"] == "Convolution" and any("add" in n.lower() for n in e["fused"])) >= 16  # residual sums
    assert abs(net.stat("flops") / 256 / 1e9 - 8.178) < 0.01
    assert net.stat("arena_bytes") < 2.0e9                              # conv1 out (411 MB) + pool + block buffers, reused
    assert [i[0] for i in net.inputs] == ["data"] and net.inputs[0][1] == (256, 3, 224, 224)
    assert net.outputs[0][1] == (256, 1000)


def test_resnet50_schedule_with_chained_reduce_convs(monkeypatch):
    # opt-in (B200_FUSE_NEXT=1): 12 of the 1x1 reduce convs (Ca <= 256) are computed inside the kernel that produces their input
    monkeypatch.setenv("B200_FUSE_NEXT", "1")
    net = dry(zoo.resnet50(batch=256))
    eg = net.get_exec_graph_info()
    assert kinds(net) == {"Fill": 1, "Convolution": 37, "Pooling": 1, "FullyConnected": 1}
    assert sum(1 for e in eg if "+1x1x" in e["impl"]) == 12
    assert abs(net.stat("flops") / 256 / 1e9 - 8.178) < 0.01


def test_resnet50_schedule_without_shortcut_folding(monkeypatch):
    monkeypatch.setenv("B200_FOLD_SHORTCUT", "0")
    monkeypatch.setenv("B200_FUSE_POOL", "0")
    net = dry(zoo.resnet50(batch=256))
    assert kinds(net) == {"Convolution": 53, "Pooling": 2, "FullyConnected": 1}
    assert abs(net.stat("flops") / 256 / 1e9 - 8.178) < 0.01


def test_resnet50_fp32_uses_cuda_core_conv():
    net = dry(zoo.resnet50(batch=2), "FP32")
    assert all(e["impl"].startswith("simt_fp32") for e in net.get_exec_graph_info() if e["type"] == "Convolution")
    assert all(e["out_precision"] == "FP32" for e in net.get_exec_graph_info())


def test_mobilenet_schedule():
    net = dry(zoo.mobilenet_v2(batch=64))
    eg = net.get_exec_graph_info()
    # 14 of the 17 inverted-residual blocks run as ONE kernel each (expand -> depthwise -> project (+ add), mbconv.cu); the three 7x7 blocks
    # (160 -> 960 channels: the X tile, the weight chunks and the staging do not fit 227 KB) keep their three kernels
    mb = [e for e in eg if "_mbconv_" in e["impl"]]
    dw = [e for e in eg if e["impl"].startswith("dw_")]
    assert len(mb) == 14 and len(dw) == 3 and kinds(net)["Convolution"] == 26
    assert sum(1 for e in mb if e["impl"].endswith("_res")) == 8        # residual adds taken from the block input tile in shared memory
    assert all(len(e["fused"]) >= 2 for e in dw)                        # BN (+Clamp) folded into the depthwise kernel
    assert not any(e["type"] == "Eltwise" for e in eg)                  # the 10 residual adds ride on the project convs
    assert net.get_exec_graph_info()[0]["impl"].endswith("_s2d_stem")   # 3x3 / stride-2 first convolution on the space-to-depth stem path


def test_mobilenet_schedule_without_block_fusion(monkeypatch):
    monkeypatch.setenv("B200_FUSE_MBCONV", "0")
    net = dry(zoo.mobilenet_v2(batch=64))
    eg = net.get_exec_graph_info()
    dw = [e for e in eg if e["impl"].startswith("dw_")]
    assert len(dw) == 17 and kinds(net)["Convolution"] == 52


def test_bert_schedule():
    net = dry(zoo.bert_base(batch=2, seq=64, layers=2))
    k = kinds(net)
    assert k["FullyConnected"] == 2 * 6 + 1 or k["FullyConnected"] == 2 * 6
    assert k["MatMul"] == 4 and k["Softmax"] == 2 and k["MVN"] == 5
    mvn = [e for e in net.get_exec_graph_info() if e["type"] == "MVN"]
    assert all(len(e["fused"]) == 3 for e in mvn)                       # MVN + gamma Multiply + beta Add (mkldnn_graph_optimizer.cpp:1132-1168)
    gelu = [e for e in net.get_exec_graph_info() if e["type"] == "FullyConnected" and any("gelu" in n.lower() for n in e["fused"])]
    assert len(gelu) == 2


def test_ssd_schedule():
    net = dry(zoo.ssd300_resnet34(batch=2), "FP16")
    k = kinds(net)
    assert k["DetectionOutput"] == 1 and k["Softmax"] == 1
    assert net.outputs[0][1][-1] == 7


def test_conv_bias_act_residual_act_is_one_kernel():
    rng = np.random.default_rng(0)
    f = Function("t")
    x = f.parameter((1, 64, 8, 8))
    y = f.convolution(x, rng.standard_normal((64, 64, 3, 3)).astype(np.float32), (1, 1), (1, 1), (1, 1))
    y = f.add(y, rng.standard_normal((1, 64, 1, 1)).astype(np.float32))
    y = f.relu(f.add(y, x))
    f.result(y)
    eg = dry(f).get_exec_graph_info()
    assert len(eg) == 1 and len(eg[0]["fused"]) == 4 and eg[0]["out_layout"] == "nhwc8"


def test_conv_with_two_consumers_is_not_fused_with_the_add():
    # the sum may only be fused when the conv has exactly one consumer (mkldnn_graph_optimizer.cpp:958-1130)
    rng = np.random.default_rng(0)
    f = Function("t")
    x = f.parameter((1, 16, 8, 8))
    y = f.convolution(x, rng.standard_normal((16, 16, 1, 1)).astype(np.float32))
    z = f.add(y, x)
    f.result(z)
    f.result(f.relu(y))
    assert kinds(dry(f))["Eltwise"] >= 1


def test_output_naming_rule():
    # result name = producer friendly name, ".port" only for multi-output producers (mkldnn_graph.cpp:230-241)
    f = Function("t")
    x = f.parameter((2, 10), name="in0")
    v, i = f.topk(x, 3, 1, name="tk")
    f.result(v)
    f.result(i)
    f.result(f.relu(x, name="act"))
    net = dry(f, "FP32")
    assert [o[0] for o in net.outputs] == ["tk.0", "tk.1", "act"]
    assert net.outputs[1][2] in (_capi.I32, _capi.I64)


def test_shape_inference_conv_pool():
    f = Function("t")
    x = f.parameter((1, 3, 224, 224))
    y = f.convolution(x, np.zeros((64, 3, 7, 7), np.float32), (2, 2), (3, 3), (3, 3))
    assert y.shape == (1, 64, 112, 112)
    p = f.max_pool(y, (3, 3), (2, 2), (1, 1), (1, 1))
    assert p.shape == (1, 64, 56, 56)
    q = f.avg_pool(p, (3, 3), (2, 2), (0, 0), (0, 0), rounding="ceil")
    assert q.shape == (1, 64, 28, 28)


def test_malformed_graphs_are_rejected():
    f = Function("t")
    x = f.parameter((1, 3, 8, 8))
    with pytest.raises(_capi.B200Error) as e:
        f.convolution(x, np.zeros((4, 5, 3, 3), np.float32))     # channel mismatch
    assert e.value.code == _capi.ERR_INVALID
    with pytest.raises(_capi.B200Error):
        f.matmul(x, f.constant(np.zeros((9, 2), np.float32)))    # K mismatch
    g = Function("empty")
    g.parameter((1, 4))
    with pytest.raises(_capi.B200Error):
        dry(g)                                                   # no Result


def test_op_support_query():
    assert all(_capi.lib.b200_op_supported(k) == 1 for k in range(len(_capi.OPS)))
    assert _capi.lib.b200_op_supported(10_000) == 0


def test_config_keys_and_metrics():
    core = Core()
    with pytest.raises(ie.NotFound):
        core.set_config({"CPU_THREADS_NUM": "4"})               # unknown key -> NotFound (config.cpp:112-114)
    with pytest.raises(ie.NotFound):
        core.load_network(Function("x"), "GPU")
    with pytest.raises(ValueError):
        core.set_config({"B200_INFERENCE_PRECISION": "INT4"})
    assert core.get_metric("B200", "OPTIMIZATION_CAPABILITIES") == ["FP32", "BF16", "FP16"]
    assert "PERF_COUNT" in core.get_metric("B200", "SUPPORTED_CONFIG_KEYS")
    f = Function("n1")
    f.result(f.relu(f.parameter((1, 8))))
    net = core.load_network(f, "B200.0", {"B200_THROUGHPUT_STREAMS": "2"}, dry=True)
    assert net.get_metric("NETWORK_NAME") == "n1"
    assert net.get_metric("OPTIMAL_NUMBER_OF_INFER_REQUESTS") == 4
    assert net.get_config("DEVICE_ID") == "0"
    with pytest.raises(ie.NotFound):
        net.get_config("NOPE")
