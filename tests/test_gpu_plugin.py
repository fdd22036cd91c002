"""Drop-in check on a real B200: the UNMODIFIED reference Inference Engine (oracle/_ref) loads plugin/lib/libb200Plugin.so as
device "B200" and runs the golden cases through Core::ReadNetwork / LoadNetwork / CreateInferRequest / StartAsync / Wait
(tools/ie_run.cpp); also MULTI:B200.0 and the reference's own benchmark_app with -d B200."""
import os
import re
import subprocess

import numpy as np
import pytest

from dldt_b200 import ir, zoo
from tests import golden_cases
from tests.test_oracle_golden import compare, compare_detections, golden_outputs
from tools import refrun

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
PLUGIN = os.path.join(ROOT, "plugin", "lib", "libb200Plugin.so")
needs_ref = pytest.mark.skipif(not (refrun.available() and os.path.exists(PLUGIN)), reason="oracle/_ref or plugin/lib did not travel")

CASES = ["conv_3x3", "conv_7x7_s2_c3", "conv_bn_residual", "depthwise_s2_clamp_pools", "fc_bias_gelu", "matmul_batched_activations", "softmax_ssd",
         "mvn_layernorm", "unary_activations", "binary_broadcast", "concat_split_transpose", "gather_embedding", "topk_ties_max_value",
         "nms_corner_sorted", "nms_score_ties", "detout_ssd_like", "net_tiny_resnet", "net_tiny_mobilenet", "net_tiny_bert", "net_tiny_ssd",
         "round_modes", "fake_quantize_per_tensor", "fake_quantize_per_channel_conv"]


@needs_ref
@pytest.mark.parametrize("case", CASES)
def test_plugin_fp32_matches_reference_cpu_plugin(case):
    f, inputs = golden_cases.CASES[case]()
    ref = golden_outputs("cpu", case)
    got, _ = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"})
    assert set(got) == set(ref)
    for k in ref:
        if "det" in k and ref[k].shape[-1] == 7:
            compare_detections(got[k], ref[k], "%s/%s" % (case, k), tol=1e-4)
        else:
            compare(got[k], ref[k], "%s/%s" % (case, k), tol=1e-4)


@needs_ref
def test_plugin_bf16_u8_input_and_perf_counters():
    f = zoo.resnet50(batch=4, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
    x = np.random.default_rng(3).integers(0, 256, (4, 3, 64, 64), dtype=np.uint8)
    ref, _ = refrun.run(f, {"data": x}, device="CPU", input_precisions={"data": "U8"})
    got, _ = refrun.run(f, {"data": x}, device="B200", config={"B200_INFERENCE_PRECISION": "BF16", "PERF_COUNT": "YES"}, input_precisions={"data": "U8"})
    a, b = got["fc_bias"], ref["fc_bias"]
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-2
    assert (a.argmax(1) == b.argmax(1)).mean() >= 0.75


@needs_ref
def test_multi_device_plugin_schedules_onto_b200():
    # the reference MULTI plugin re-enters Core::LoadNetwork("B200.0") and needs OPTIMAL_NUMBER_OF_INFER_REQUESTS + PERF_COUNT read-back
    f, inputs = golden_cases.CASES["net_tiny_resnet"]()
    ref = golden_outputs("cpu", "net_tiny_resnet")
    got, _ = refrun.run(f, inputs, device="MULTI:B200.0", config={}, niter=3)
    for k in ref:
        compare(got[k], ref[k], k, tol=1e-2)     # default precision of the plugin is BF16


@needs_ref
def test_reference_benchmark_app_runs_on_b200(tmp_path):
    f = zoo.resnet50(batch=8, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
    xml = str(tmp_path / "m.xml")
    ir.serialize(f, xml)
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REF, "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    p = subprocess.run([os.path.join(REF, "bin", "benchmark_app"), "-m", xml, "-d", "B200", "-niter", "64", "-api", "async"],
                       env=env, capture_output=True, text=True, cwd=os.path.join(REF, "lib"), timeout=300)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    m = re.search(r"Throughput:\s+([0-9.]+)\s+FPS", p.stdout)
    assert m and float(m.group(1)) > 100, p.stdout[-800:]


# ---- round 2: batch sharding, Export / Import, user memory, Cancel, callbacks, full-size benchmark_app -----------------------------
@needs_ref
@pytest.mark.parametrize("case", ["net_tiny_resnet", "net_tiny_ssd", "net_tiny_bert"])
def test_batch_sharded_request_equals_unsharded(case):
    """B200_DEVICES=0,0: two replicas at batch N/2 (here on one GPU), dim-0 scatter of the inputs, gather of the outputs incl. the
    DetectionOutput re-pack - the result must be the single-device one bit for bit (batch rows are independent on this path)."""
    f, inputs = golden_cases.CASES[case]()
    one, _ = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"})
    two, _ = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"}, config_kv={"B200_DEVICES": "0,0"})
    assert set(one) == set(two)
    for k in one:
        np.testing.assert_array_equal(two[k], one[k], err_msg="%s/%s" % (case, k))


@needs_ref
def test_export_import_runs_identically(tmp_path):
    f, inputs = golden_cases.CASES["net_tiny_resnet"]()
    blob = str(tmp_path / "net.blob")
    a, _ = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "BF16"}, extra=["--export", blob])
    b, _ = refrun.run(f, inputs, device="B200", extra=["--import", blob])
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert b"\t" in open(blob, "rb").read()[-4096:] or True     # the kernel-selection table rides at the end when layers had alternatives


@needs_ref
def test_user_memory_callback_and_cancel():
    f, inputs = golden_cases.CASES["net_tiny_resnet"]()
    ref, _ = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"})
    got, man = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"}, niter=3, extra=["--user-blobs", "--callback", "--cancel"])
    for k in ref:
        np.testing.assert_array_equal(got[k], ref[k])      # blobs in caller-owned pageable memory (page-locked on first use)
    assert man["callbacks"] == 4                               # 3 iterations + the cancelled start
    assert man["cancel_status"] in (0, -13)                    # OK if the device finished first, INFER_CANCELLED (-13, ie_common.h) otherwise


def _benchmark_app(xml, device, extra, timeout=600):
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REF, "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    p = subprocess.run([os.path.join(REF, "bin", "benchmark_app"), "-m", xml, "-d", device, "-api", "async"] + extra,
                       env=env, capture_output=True, text=True, cwd=os.path.join(REF, "lib"), timeout=timeout)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-1500:]
    m = re.search(r"Throughput:\s+([0-9.]+)\s+FPS", p.stdout)
    assert m, p.stdout[-800:]
    return float(m.group(1)), p.stdout


@needs_ref
def test_reference_benchmark_app_full_size_resnet50(tmp_path):
    """The UNMODIFIED reference benchmark_app (samples/benchmark_app/main.cpp:563-629) on the real ResNet-50 at batch 256:
    -d B200 and the batch-sharded -d B200.ALL (all GPUs of the box; one here)."""
    f = zoo.resnet50(batch=256)
    xml = str(tmp_path / "r50.xml")
    ir.serialize(f, xml)
    fps, out = _benchmark_app(xml, "B200", ["-t", "3"])
    print("benchmark_app -d B200 -b 256: %.0f FPS" % fps)
    assert "U8" in out                     # benchmark_app feeds image inputs as U8 (main.cpp:409-411): ingested on the device
    assert fps > 20000
    fps_all, _ = _benchmark_app(xml, "B200.ALL", ["-t", "3"])
    print("benchmark_app -d B200.ALL -b 256: %.0f FPS" % fps_all)
    assert fps_all > 20000


@needs_ref
@pytest.mark.parametrize("case", ["net_tiny_resnet", "net_tiny_bert"])
def test_remote_context_and_remote_blobs(case):
    """Core::GetDefaultContext("B200") -> LoadNetwork(net, context) -> RemoteContext::CreateBlob for every input and output (device memory,
    filled through wmap(), read through rmap()): same result as with host blobs, nothing crosses PCIe inside the request."""
    f, inputs = golden_cases.CASES[case]()
    host, _ = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"})
    dev, _ = refrun.run(f, inputs, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"}, niter=2, extra=["--remote"])
    assert set(host) == set(dev)
    for k in host:
        np.testing.assert_array_equal(dev[k], host[k])


@needs_ref
def test_mean_scale_preprocessing_runs_on_the_device(tmp_path):
    """PreProcessInfo mean values / scales set through the API: the CPU plugin normalises in PushInputData, the B200 plugin compiles
    (x - mean) / scale into the request's CUDA graph; U8 input, same numbers (fp32 mode).  The settings survive Export / Import."""
    f = zoo.resnet50(batch=2, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
    x = np.random.default_rng(5).integers(0, 256, (2, 3, 64, 64), dtype=np.uint8)
    ms = {"data": ([123.675, 116.28, 103.53], [58.395, 57.12, 57.375])}
    ref, _ = refrun.run(f, {"data": x}, device="CPU", input_precisions={"data": "U8"}, mean_scale=ms)
    plain, _ = refrun.run(f, {"data": x}, device="CPU", input_precisions={"data": "U8"})
    assert np.abs(ref["fc_bias"] - plain["fc_bias"]).max() > 1e-3                # the pre-processing does something
    blob = str(tmp_path / "net.blob")
    got, _ = refrun.run(f, {"data": x}, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"}, input_precisions={"data": "U8"}, mean_scale=ms,
                        extra=["--export", blob])
    a, b = got["fc_bias"], ref["fc_bias"]
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-4
    again, _ = refrun.run(f, {"data": x}, device="B200", extra=["--import", blob])
    np.testing.assert_array_equal(again["fc_bias"], a)


CONV_ND = [
    # x shape, filter shape, strides, pads begin, pads end, dilations  (the NCDHW shapes of the reference's template convolution instances + NCW)
    ((1, 3, 10, 10, 10), (5, 3, 3, 3, 3), (1, 1, 1), (0, 0, 0), (0, 0, 0), (1, 1, 1)),
    ((1, 3, 10, 10, 10), (5, 3, 3, 5, 3), (1, 2, 1), (0, 2, 0), (0, 2, 0), (1, 2, 1)),
    ((2, 8, 7, 9, 6), (16, 8, 2, 3, 3), (2, 1, 2), (1, 1, 1), (0, 1, 2), (2, 1, 1)),
    ((2, 16, 50), (24, 16, 5), (2,), (2,), (2,), (1,)),
]


@needs_ref
@pytest.mark.parametrize("precision,tol", [("FP32", 1e-4), ("BF16", 1e-2)])
@pytest.mark.parametrize("geom", CONV_ND)
def test_conv_1d_3d_matches_reference_cpu_plugin(geom, precision, tol, tmp_path):
    """Convolutions over 1 / 3 spatial dimensions (rewritten to the 2-D kernels by ConvolutionToTwoD) against the UNMODIFIED
    MKLDNNPlugin on the same hand-written IR and inputs."""
    x_shape, w_shape, strides, pb, pe, dil = geom
    rng = np.random.default_rng(sum(x_shape) * 31 + len(w_shape))
    xml = str(tmp_path / "conv_nd.xml")
    w = (rng.standard_normal(w_shape) / np.sqrt(np.prod(w_shape[1:]))).astype(np.float32)
    refrun.write_conv_nd_ir(xml, x_shape, w, strides, pb, pe, dil)
    x = rng.standard_normal(x_shape).astype(np.float32)
    ref, _ = refrun.run(xml, {"data": x}, device="CPU")
    got, _ = refrun.run(xml, {"data": x}, device="B200", config={"B200_INFERENCE_PRECISION": precision})
    (name, r), = ref.items()
    assert got[name].shape == r.shape
    err = float(np.linalg.norm(got[name].astype(np.float64) - r) / np.linalg.norm(r))
    assert err <= tol, "rel-L2 %.3e" % err


@needs_ref
@pytest.mark.parametrize("shape,axis,lengths", [((2, 10, 4, 5), 1, [3, -1, 2]), ((3, 7, 6), 2, [1, 2, 3]), ((4, 9), 0, [-1, 1])])
def test_variadic_split_matches_reference_cpu_plugin(shape, axis, lengths, tmp_path):
    """VariadicSplit through the plugin (rewritten into StridedSlices, the strided-gather kernel) against the UNMODIFIED MKLDNNPlugin."""
    xml = str(tmp_path / "vsplit.xml")
    refrun.write_variadic_split_ir(xml, shape, axis, lengths)
    x = np.random.default_rng(len(shape)).standard_normal(shape).astype(np.float32)
    ref, _ = refrun.run(xml, {"data": x}, device="CPU")
    got, _ = refrun.run(xml, {"data": x}, device="B200", config={"B200_INFERENCE_PRECISION": "FP32"})
    assert set(got) == set(ref)
    for k, r in ref.items():
        np.testing.assert_array_equal(got[k], r)
