"""The C++ device plugin (plugin/lib/libb200Plugin.so) loaded by the UNMODIFIED reference Inference Engine
(oracle/_ref/lib/libinference_engine.so) through plugins.xml - host-side behaviour that needs no GPU:
registration under the device name "B200", QueryNetwork coverage, metrics, config-key errors, the lowering of an
ngraph::Function through the C ABI (dry compile) and the loud failure when no device exists."""
import json
import os
import subprocess

import numpy as np
import pytest

from dldt_b200 import ir, zoo
from tests import golden_cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
IE_RUN = os.path.join(REF, "bin", "ie_run")
PLUGIN = os.path.join(ROOT, "plugin", "lib", "libb200Plugin.so")

pytestmark = pytest.mark.skipif(not (os.path.exists(IE_RUN) and os.path.exists(PLUGIN)),
                                reason="needs oracle/_ref (oracle/build_ref.sh) and plugin/lib (make -C plugin), both built by __graft_entry__.build()")


def ie_run(xml, *args):
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(REF, "lib") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    cmd = [IE_RUN, "--model", xml, "--device", "B200", "--plugins", os.path.join(REF, "lib", "plugins.xml")] + list(args)
    return subprocess.run(cmd, env=env, capture_output=True, text=True)


@pytest.fixture(scope="module")
def tiny_resnet_xml(tmp_path_factory):
    d = tmp_path_factory.mktemp("ir")
    f = zoo.resnet50(batch=2, image=64, width=16, blocks=(1, 1, 1, 1), num_classes=24)
    xml = str(d / "tiny_resnet.xml")
    ir.serialize(f, xml)
    return xml


def test_plugin_exports_only_the_factory_symbol():
    out = subprocess.run(["nm", "-D", "--defined-only", PLUGIN], capture_output=True, text=True).stdout
    names = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    assert names == ["CreatePluginEngine"]        # ie_iplugin_internal.hpp:319-331


def test_query_network_and_metrics(tiny_resnet_xml):
    p = ie_run(tiny_resnet_xml, "--query")
    assert p.returncode == 0, p.stderr
    q = json.loads(p.stdout)
    assert q["supported_layers"] == q["total_layers"] > 100
    assert q["capabilities"] == ["FP32", "BF16", "FP16"]
    assert {"DEVICE_ID", "PERF_COUNT", "B200_INFERENCE_PRECISION", "B200_THROUGHPUT_STREAMS"} <= set(q["config_keys"])
    assert q["architecture"] == "sm_100a"


def test_load_network_lowers_and_fuses(tiny_resnet_xml):
    p = ie_run(tiny_resnet_xml, "--dry", "--config", "B200_DRY_COMPILE=YES,B200_THROUGHPUT_STREAMS=3")
    assert p.returncode == 0, p.stderr
    d = json.loads(p.stdout)
    assert d["network_name"] == "resnet50_v1.5"
    assert d["optimal_requests"] == 7                        # 2 per stream + 1
    # 17 convs with BN/ReLU/residual folded; the two projection shortcuts with >= 64 input channels ride on their block's last 1x1 (K concat)
    assert d["exec_convolutions"] == 15 and d["exec_nodes"] <= 24


def test_unknown_config_key_is_not_found(tiny_resnet_xml):
    p = ie_run(tiny_resnet_xml, "--dry", "--config", "B200_DRY_COMPILE=YES,CPU_THREADS_NUM=4")
    assert p.returncode != 0 and "NOT_FOUND" in p.stderr and "CPU_THREADS_NUM" in p.stderr


def test_bad_precision_value_is_rejected(tiny_resnet_xml):
    p = ie_run(tiny_resnet_xml, "--dry", "--config", "B200_DRY_COMPILE=YES,B200_INFERENCE_PRECISION=INT4")
    assert p.returncode != 0 and "B200_INFERENCE_PRECISION" in p.stderr


def test_no_device_fails_loudly(tiny_resnet_xml):
    from dldt_b200 import _capi
    if _capi.device_count() > 0:
        pytest.skip("GPU present")
    p = ie_run(tiny_resnet_xml, "--niter", "1")
    assert p.returncode != 0 and "no CUDA device" in p.stderr and "no CPU fallback" in p.stderr


@pytest.mark.parametrize("case", ["net_tiny_mobilenet", "net_tiny_bert", "net_tiny_ssd", "unary_activations", "topk_ties_min_index", "nms_corner_sorted",
                                  "concat_split_transpose", "mvn_layernorm", "matmul_batched_activations", "fake_quantize_per_channel_conv", "round_modes"])
def test_every_model_family_lowers_through_the_plugin(case, tmp_path):
    f, _ = golden_cases.CASES[case]()
    xml = str(tmp_path / "m.xml")
    ir.serialize(f, xml)
    q = ie_run(xml, "--query")
    assert q.returncode == 0, q.stderr
    qq = json.loads(q.stdout)
    assert qq["supported_layers"] == qq["total_layers"], qq
    p = ie_run(xml, "--dry", "--config", "B200_DRY_COMPILE=YES,B200_INFERENCE_PRECISION=FP32")
    assert p.returncode == 0, p.stderr


# ---- batch sharding / Export / Import: host logic that needs no device (dry compile) --------------------------------------------
def test_sharded_load_divides_the_batch(tmp_path):
    """B200_DEVICES=0,1: the replica is the same function at batch N/2 (ShardBatch re-validates the clone incl. the Reshape targets that
    spell the batch out - the BERT head split); a malformed list and DEVICE_ID=ALL without devices are handled at LoadNetwork."""
    f = zoo.bert_base(batch=4, seq=32, hidden=64, heads=2, layers=1, ffn=128, vocab=100, max_pos=64)
    xml = str(tmp_path / "bert.xml")
    ir.serialize(f, xml)
    p = ie_run(xml, "--dry", "--config", "B200_DRY_COMPILE=YES", "--config-kv", "B200_DEVICES=0,1")
    assert p.returncode == 0, p.stderr
    assert json.loads(p.stdout)["network_name"] == "bert_base"
    p = ie_run(xml, "--dry", "--config", "B200_DRY_COMPILE=YES", "--config-kv", "B200_DEVICES=0+1")
    assert p.returncode != 0 and "bad device id" in p.stderr           # the separator is a comma
    p = ie_run(xml, "--dry", "--config", "B200_DRY_COMPILE=YES,DEVICE_ID=ALL")   # no visible device: ALL resolves to one replica
    assert p.returncode == 0, p.stderr


def test_sharded_dry_compile_two_devices(tiny_resnet_xml):
    p = ie_run(tiny_resnet_xml, "--dry", "--config", "B200_DRY_COMPILE=YES", "--config-kv", "B200_DEVICES=0,1")
    assert p.returncode == 0, p.stderr
    d = json.loads(p.stdout)
    assert d["exec_convolutions"] == 15                      # same schedule, batch 1 per device
    p = ie_run(tiny_resnet_xml, "--dry", "--config", "B200_DRY_COMPILE=YES", "--config-kv", "B200_DEVICES=0,1,2")
    assert p.returncode != 0 and "does not divide" in p.stderr


def test_export_import_round_trip(tiny_resnet_xml, tmp_path):
    blob = str(tmp_path / "net.blob")
    p = ie_run(tiny_resnet_xml, "--dry", "--export", blob, "--config", "B200_DRY_COMPILE=YES,B200_INFERENCE_PRECISION=FP16", "--input-precision", "data=U8")
    assert p.returncode == 0, p.stderr
    first = json.loads(p.stdout)
    assert open(blob, "rb").read(8) == b"B200NET2"
    p = ie_run(tiny_resnet_xml, "--dry", "--import", blob, "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode == 0, p.stderr
    again = json.loads(p.stdout)
    assert again == first
    assert again["input_precisions"] == {"data": "U8"}       # user-side precision travels with the export
    assert again["config_precision"] == "FP16"               # and so does the arithmetic the network was compiled for
    # a foreign stream is refused with NetworkNotRead
    bad = str(tmp_path / "bad.blob")
    open(bad, "wb").write(b"not a network")
    p = ie_run(tiny_resnet_xml, "--dry", "--import", bad, "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode != 0 and "not an exported B200 network" in p.stderr


def test_core_cache_dir_drives_export_and_import(tiny_resnet_xml, tmp_path):
    """IMPORT_EXPORT_SUPPORT is advertised, so Core::LoadNetwork with CACHE_DIR exports on the first load and imports on the second
    (inference-engine/src/inference_engine/ie_core.cpp:265-340)."""
    cache = tmp_path / "cache"
    cache.mkdir()
    p = ie_run(tiny_resnet_xml, "--dry", "--cache-dir", str(cache), "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode == 0, p.stderr
    blobs = list(cache.iterdir())
    assert len(blobs) == 1 and blobs[0].stat().st_size > 100000
    stamp = blobs[0].stat().st_mtime_ns
    p = ie_run(tiny_resnet_xml, "--dry", "--cache-dir", str(cache), "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode == 0, p.stderr
    assert [b.stat().st_mtime_ns for b in cache.iterdir()] == [stamp]      # second load read the entry, did not rewrite it


def test_unsupported_geometry_fails_at_load_not_at_first_infer(tmp_path):
    """DetectionOutput over 20000 priors with top_k = -1 has no kernel (the fall-back sorts at most 16384 keys in shared memory):
    QueryNetwork leaves the layer out (so HETERO could place it elsewhere) and LoadNetwork refuses it - the reference fails at
    LoadNetwork too, never inside the first Infer."""
    import numpy as np
    from dldt_b200 import Function
    priors, classes = 20000, 3
    f = Function("detout_too_big")
    loc = f.parameter((1, priors * 4), name="loc")
    conf = f.parameter((1, priors * classes), name="conf")
    pri = np.zeros((1, 2, priors * 4), np.float32)
    det = f.detection_output(loc, conf, pri, classes, 0, -1, 10, "caffe.PriorBoxParameter.CENTER_SIZE", True, 0.45, 0.01, name="det")
    f.result(det)
    xml = str(tmp_path / "d.xml")
    ir.serialize(f, xml)
    q = ie_run(xml, "--query")
    assert q.returncode == 0, q.stderr
    q = json.loads(q.stdout)
    assert q["supported_layers"] < q["total_layers"]
    p = ie_run(xml, "--dry", "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode != 0 and "priors" in p.stderr
    # the same layer with top_k = 200 takes the radix-select path (up to 32768 priors)
    f2 = Function("detout_select")
    loc = f2.parameter((1, priors * 4), name="loc")
    conf = f2.parameter((1, priors * classes), name="conf")
    f2.result(f2.detection_output(loc, conf, pri, classes, 0, 200, 10, "caffe.PriorBoxParameter.CENTER_SIZE", True, 0.45, 0.01, name="det"))
    xml2 = str(tmp_path / "d2.xml")
    ir.serialize(f2, xml2)
    p = ie_run(xml2, "--dry", "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode == 0, p.stderr


@pytest.mark.parametrize("x_shape,w_shape,strides,pb,pe,dil", [
    ((1, 3, 10, 10, 10), (5, 3, 3, 5, 3), (1, 2, 1), (1, 2, 0), (0, 2, 1), (1, 2, 1)),
    ((2, 16, 50), (24, 16, 5), (2,), (2,), (2,), (1,)),
])
def test_convolutions_over_1_and_3_spatial_dims_are_rewritten_to_2d(tmp_path, x_shape, w_shape, strides, pb, pe, dil):
    """NCW / NCDHW convolutions (single_layer_tests/convolution.cpp of the reference's template instances) reach the 2-D kernels through
    ConvolutionToTwoD (plugin/src/b200_lowering.cpp): QueryNetwork accepts every layer, the compiled schedule has ONE convolution."""
    from tools import refrun
    xml = str(tmp_path / "conv_nd.xml")
    refrun.write_conv_nd_ir(xml, x_shape, np.ones(w_shape, np.float32), strides, pb, pe, dil)
    q = ie_run(xml, "--query")
    assert q.returncode == 0, q.stderr
    assert json.loads(q.stdout)["supported_layers"] == json.loads(q.stdout)["total_layers"]
    p = ie_run(xml, "--dry", "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode == 0, p.stderr
    assert json.loads(p.stdout)["exec_convolutions"] == 1


@pytest.mark.parametrize("family,seed", [("build", s) for s in (0, 7, 19, 43, 1003, 1012, 1040)] + [("build_sequence", s) for s in (2, 8, 19, 31)] +
                         [("build_misc", s) for s in (5, 37, 61, 90)] + [("build_nms", 3), ("build_detection_output", 4)])
def test_random_graphs_load_through_the_plugin(tmp_path, family, seed):
    """The seeded random networks of tests/random_graphs.py as IR v10 through Core::ReadNetwork + LoadNetwork("B200") (nGraph passes,
    lowering, compile, plan) on a machine without a GPU: every layer supported, the network compiles in fp32 and bf16."""
    from tests import random_graphs
    f, _ = getattr(random_graphs, family)(seed)
    xml = str(tmp_path / "random.xml")
    ir.serialize(f, xml)
    q = ie_run(xml, "--query")
    assert q.returncode == 0, q.stderr
    assert json.loads(q.stdout)["supported_layers"] == json.loads(q.stdout)["total_layers"]
    for precision in ("FP32", "BF16"):
        p = ie_run(xml, "--dry", "--config", "B200_DRY_COMPILE=YES,B200_INFERENCE_PRECISION=" + precision)
        assert p.returncode == 0, p.stderr


def test_variadic_split_becomes_strided_slices(tmp_path):
    """VariadicSplit (one length -1) is rewritten into one StridedSlice per output AFTER CommonOptimizations (whose StridedSliceOptimization
    merges neighbouring slices into a VariadicSplit): every layer supported, output names kept."""
    from tools import refrun
    xml = str(tmp_path / "vsplit.xml")
    refrun.write_variadic_split_ir(xml, (2, 10, 4, 5), 1, [3, -1, 2])
    q = ie_run(xml, "--query")
    assert q.returncode == 0, q.stderr
    assert json.loads(q.stdout)["supported_layers"] == json.loads(q.stdout)["total_layers"]
    p = ie_run(xml, "--dry", "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode == 0, p.stderr


def test_mean_image_survives_export_and_import(tiny_resnet_xml, tmp_path):
    """PreProcessInfo MEAN_IMAGE is compiled into the network as one element-wise Subtract (AddMeanScale) and travels with the exported
    stream (per-channel planes after the mean / scale values): the imported network has the same schedule as the loaded one, which is
    longer than the schedule without a mean image."""
    plain = json.loads(ie_run(tiny_resnet_xml, "--dry", "--config", "B200_DRY_COMPILE=YES").stdout)
    blob = str(tmp_path / "mean.blob")
    p = ie_run(tiny_resnet_xml, "--dry", "--export", blob, "--mean-image", "data=0.5", "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode == 0, p.stderr
    loaded = json.loads(p.stdout)
    assert loaded["exec_nodes"] > plain["exec_nodes"]     # the Subtract (and the layout step in front of the first convolution)
    p = ie_run(tiny_resnet_xml, "--dry", "--import", blob, "--config", "B200_DRY_COMPILE=YES")
    assert p.returncode == 0, p.stderr
    assert json.loads(p.stdout) == loaded
    assert os.path.getsize(blob) > 3 * 64 * 64 * 4        # the three 64 x 64 fp32 planes are in the stream
