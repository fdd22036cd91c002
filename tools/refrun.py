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
    if isinstance(func, str):      # an IR that already exists (graphs the recorded-Function front end cannot express, e.g. NCDHW convolutions)
        xml = func
    else:
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


def write_conv_nd_ir(xml_path, x_shape, weights, strides, pads_begin, pads_end, dilations, relu=True):
    """IR v10 of Parameter -> Convolution (1, 2 or 3 spatial dimensions) [-> ReLU] -> Result, written by hand: the recorded-Function
    front end (and the C ABI behind it) is 2-D only, the plugin accepts the others by rewriting them (ConvolutionToTwoD)."""
    w = np.ascontiguousarray(weights, np.float32)
    nd = len(x_shape) - 2
    out = [x_shape[0], w.shape[0]] + [(x_shape[2 + k] + pads_begin[k] + pads_end[k] - (dilations[k] * (w.shape[2 + k] - 1) + 1)) // strides[k] + 1
                                      for k in range(nd)]

    def port(pid, shape):
        return '<port id="%d" precision="FP32">%s</port>' % (pid, "".join("<dim>%d</dim>" % d for d in shape))

    j = lambda v: ", ".join(str(int(a)) for a in v)
    layers = [
        '<layer id="0" name="data" type="Parameter" version="opset1"><data element_type="f32" shape="%s"/><output>%s</output></layer>' % (j(x_shape), port(0, x_shape)),
        '<layer id="1" name="w" type="Const" version="opset1"><data element_type="f32" shape="%s" offset="0" size="%d"/><output>%s</output></layer>'
        % (j(w.shape), w.nbytes, port(0, w.shape)),
        '<layer id="2" name="conv" type="Convolution" version="opset1"><data strides="%s" pads_begin="%s" pads_end="%s" dilations="%s" auto_pad="explicit"/>'
        '<input>%s%s</input><output>%s</output></layer>' % (j(strides), j(pads_begin), j(pads_end), j(dilations), port(0, x_shape), port(1, w.shape), port(2, out)),
    ]
    edges = ['<edge from-layer="0" from-port="0" to-layer="2" to-port="0"/>', '<edge from-layer="1" from-port="0" to-layer="2" to-port="1"/>']
    last = 2
    if relu:
        layers.append('<layer id="3" name="relu" type="ReLU" version="opset1"><input>%s</input><output>%s</output></layer>' % (port(0, out), port(1, out)))
        edges.append('<edge from-layer="2" from-port="2" to-layer="3" to-port="0"/>')
        last = 3
    layers.append('<layer id="%d" name="result" type="Result" version="opset1"><input>%s</input></layer>' % (last + 1, port(0, out)))
    edges.append('<edge from-layer="%d" from-port="%d" to-layer="%d" to-port="0"/>' % (last, 2 if last == 2 else 1, last + 1))
    with open(xml_path, "w") as f:
        f.write('<?xml version="1.0"?>\n<net name="conv_nd" version="10"><layers>%s</layers><edges>%s</edges></net>\n' % ("".join(layers), "".join(edges)))
    w.tofile(xml_path[:-4] + ".bin")
    return out


def write_variadic_split_ir(xml_path, x_shape, axis, lengths):
    """IR v10 of Parameter -> VariadicSplit(axis, lengths; one length may be -1) -> one ReLU + Result per part (the recorded-Function front
    end has no VariadicSplit: the plugin rewrites it into StridedSlices).  Returns the output shapes."""
    lengths = list(lengths)
    rest = x_shape[axis] - sum(v for v in lengths if v >= 0)
    sizes = [rest if v < 0 else v for v in lengths]
    outs = [list(x_shape[:axis]) + [sz] + list(x_shape[axis + 1:]) for sz in sizes]
    consts = np.asarray([axis] + lengths, np.int64)

    def port(pid, shape, prec="FP32"):
        return '<port id="%d" precision="%s">%s</port>' % (pid, prec, "".join("<dim>%d</dim>" % d for d in shape))

    j = lambda v: ", ".join(str(int(a)) for a in v)
    n = len(lengths)
    layers = [
        '<layer id="0" name="data" type="Parameter" version="opset1"><data element_type="f32" shape="%s"/><output>%s</output></layer>' % (j(x_shape), port(0, x_shape)),
        '<layer id="1" name="axis" type="Const" version="opset1"><data element_type="i64" shape="" offset="0" size="8"/><output><port id="0" precision="I64"/></output></layer>',
        '<layer id="2" name="lengths" type="Const" version="opset1"><data element_type="i64" shape="%d" offset="8" size="%d"/><output>%s</output></layer>'
        % (n, 8 * n, port(0, [n], "I64")),
        '<layer id="3" name="vsplit" type="VariadicSplit" version="opset1"><input>%s<port id="1" precision="I64"/>%s</input><output>%s</output></layer>'
        % (port(0, x_shape), port(2, [n], "I64"), "".join(port(3 + k, outs[k]) for k in range(n))),
    ]
    edges = ['<edge from-layer="0" from-port="0" to-layer="3" to-port="0"/>', '<edge from-layer="1" from-port="0" to-layer="3" to-port="1"/>',
             '<edge from-layer="2" from-port="0" to-layer="3" to-port="2"/>']
    lid = 4
    for k in range(n):
        layers.append('<layer id="%d" name="relu%d" type="ReLU" version="opset1"><input>%s</input><output>%s</output></layer>' % (lid, k, port(0, outs[k]), port(1, outs[k])))
        layers.append('<layer id="%d" name="result%d" type="Result" version="opset1"><input>%s</input></layer>' % (lid + 1, k, port(0, outs[k])))
        edges.append('<edge from-layer="3" from-port="%d" to-layer="%d" to-port="0"/>' % (3 + k, lid))
        edges.append('<edge from-layer="%d" from-port="1" to-layer="%d" to-port="0"/>' % (lid, lid + 1))
        lid += 2
    with open(xml_path, "w") as f:
        f.write('<?xml version="1.0"?>\n<net name="vsplit" version="10"><layers>%s</layers><edges>%s</edges></net>\n' % ("".join(layers), "".join(edges)))
    consts.tofile(xml_path[:-4] + ".bin")
    return outs
