"""Run a recorded Function on the UNMODIFIED reference (oracle/_ref: Inference Engine + MKLDNNPlugin)
through tools/ie_run: serialise to IR v10, write raw inputs, read raw outputs back.
Test infrastructure (golden generation + GPU-box cross checks); never on the product path."""
import json
import os
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
IE_RUN = os.path.join(REF, "bin", "ie_run")
_NP = {"FP32": np.float32, "I32": np.int32, "U8": np.uint8, "FP16": np.float16, "I64": np.int64}


def available():
    return os.path.exists(IE_RUN) and os.path.exists(os.path.join(REF, "lib", "libMKLDNNPlugin.so"))


def run(func, inputs, device="CPU", config=None, input_precisions=None, niter=1, workdir=None, config_kv=None, extra=None, mean_scale=None):
    """Returns (outputs dict, manifest).  fp32 oracle runs must pass ENFORCE_BF16=NO."""
    from dldt_b200 import ir
    config = dict(config or {})
    if device == "CPU":
        config.setdefault("ENFORCE_BF16", "NO")
    tmp = workdir or tempfile.mkdtemp(prefix="refrun_")
    xml = os.path.join(tmp, "model.xml")
    ir.serialize(func, xml)
    cmd = [IE_RUN, "--model", xml, "--device", device, "--output-dir", tmp, "--niter", str(niter),
           "--plugins", os.path.join(REF, "lib", "plugins.xml")]
    if config:
        cmd += ["--config", ",".join("%s=%s" % kv for kv in config.items())]
    for kv in (config_kv or {}).items():       # values that contain commas (B200_DEVICES=0,1)
        cmd += ["--config-kv", "%s=%s" % kv]
    for name, (means, scales) in (mean_scale or {}).items():     # PreProcessInfo mean values / scales per channel
        cmd += ["--mean-scale", "%s=%s/%s" % (name, ":".join(repr(float(v)) for v in means), ":".join(repr(float(v)) for v in scales))]
    cmd += list(extra or [])
    for name, arr in inputs.items():
        path = os.path.join(tmp, "in_%s.bin" % name)
        np.ascontiguousarray(arr).tofile(path)
        cmd += ["--input", "%s=%s" % (name, path)]
        prec = (input_precisions or {}).get(name)
        if prec:
            cmd += ["--input-precision", "%s=%s" % (name, prec)]
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.join(REF, "lib") + ":" + os.path.join(ROOT, "dldt_b200", "lib") + ":" + env.get("LD_LIBRARY_PATH", "")
    p = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError("ie_run failed (%d): %s\n%s" % (p.returncode, p.stderr[-2000:], " ".join(cmd)))
    man = json.load(open(os.path.join(tmp, "manifest.json")))
    outs = {}
    for name, o in man["outputs"].items():
        outs[name] = np.fromfile(os.path.join(tmp, o["file"]), dtype=_NP[o["precision"]]).reshape(o["shape"])
    return outs, man
