"""The reference's OWN device-plugin conformance binary against the B200 plugin.

`oracle/_ref/bin/templateFuncTests` is docs/template_plugin/tests/functional of the reference, built UNMODIFIED by oracle/build_ref.sh
(the behaviour suites of inference-engine/tests/functional/plugin/shared: config, core integration, callbacks, Cancel / Wait states,
SetBlob / layouts / pre-processing, caching (Export / Import through CACHE_DIR), exec-graph, HETERO / MULTI integration, plus the
convolution / reshape / softmax / split single-layer instances and 684 synthetic HETERO splits).  tools/run_conformance.sh registers
libb200Plugin.so under the library / device name the binary asks for, so every test lands in the B200 plugin.

Every test must pass except the three classes below, which cannot apply to this plugin; they are matched by NAME so that a new failure
anywhere else fails this test.  The `nightly_*` instances (597 HETERO splits drawn from std::random_device at start-up, 16 more NCDHW
convolutions) are left to tools/run_conformance.sh (profiles/r2_conformance_run4_fp32.txt: all pass): a test suite should not depend on
a random draw.  (File name: last in the collection order.)"""
import os
import re
import subprocess
import xml.etree.ElementTree as ET

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "bin", "templateFuncTests")

NOT_APPLICABLE = [
    # the binary's device is called TEMPLATE: it sets the TEMPLATE plugin's private key (ours is B200_THROUGHPUT_STREAMS; unknown keys
    # must be refused, which IncorrectConfigTests checks and passes)
    (r"CorrectConfigTests\..*TEMPLATE_THROUGHPUT_STREAMS", "TEMPLATE-private configuration key"),
    # behavior/exec_graph_info.hpp:61-63,116-119 hard-codes that the device named TEMPLATE throws NotImplemented from GetExecGraphInfo;
    # the B200 plugin implements it (tests/test_plugin_host.py, tests/test_gpu_plugin.py check its contents)
    (r"ExecGraphTests\.", "expects NotImplemented from the device named TEMPLATE"),
    # op_reference/*: known-answer tests of the nGraph INTERPRETER kernels behind the TEMPLATE plugin over every element type
    # (u1 / u4 / i4 / i8 / i16 / u16 / u32 / i64 / u64 / boolean blobs, Sign, Select); the B200 plugin's blob precisions are
    # FP32 / FP16 / BF16 / I32 / U8 - the cases inside that list (19 Convert cases, all 24 GRN cases) pass
    (r"Reference(Convert|Select|Sign)LayerTest\.", "nGraph interpreter known-answer test outside the plugin's precision / op list"),
]


@pytest.mark.skipif(not os.path.exists(BIN), reason="oracle/_ref/bin/templateFuncTests not built (oracle/build_ref.sh)")
def test_reference_conformance_binary(tmp_path):
    prefix = str(tmp_path / "conformance")
    env = dict(os.environ, CONFORMANCE_TIMEOUT="600")
    p = subprocess.run([os.path.join(ROOT, "tools", "run_conformance.sh"), "*-nightly_*", prefix], env=env, capture_output=True, text=True, timeout=900)
    assert os.path.exists(prefix + ".xml"), "no gtest report: the binary died or timed out\n" + p.stdout[-3000:] + p.stderr[-2000:]
    passed, failed, na = 0, [], {}
    for case in ET.parse(prefix + ".xml").getroot().iter("testcase"):
        if case.get("status", "run") != "run" or case.get("result") == "skipped" or case.find("skipped") is not None:
            continue
        name = case.get("classname") + "." + case.get("name")
        if case.find("failure") is None:
            passed += 1
            continue
        for pattern, why in NOT_APPLICABLE:
            if re.search(pattern, name):
                na[why] = na.get(why, 0) + 1
                break
        else:
            failed.append(name + ": " + (case.find("failure").get("message") or "")[:300])
    print("reference conformance: %d passed, %d not applicable %s, %d failed" % (passed, sum(na.values()), na, len(failed)))
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):      # keep the summary next to the other evidence of the run
        with open(os.path.join(out, "conformance_pytest.txt"), "w") as f:
            f.write("passed %d\nnot applicable %d %s\nfailed %d\n%s\n" % (passed, sum(na.values()), na, len(failed), "\n".join(failed)))
    assert not failed, "%d conformance failures:\n%s" % (len(failed), "\n".join(failed[:40]))
    assert passed >= 800, "only %d conformance tests passed" % passed
