"""Serialise a recorded Function to OpenVINO IR v10 (model.xml + model.bin) so that the UNMODIFIED
reference tools (benchmark_app, the IR reader, MKLDNNPlugin) can load exactly the graph the B200
engine compiles.  Layout of the XML follows what the reference's own serializer emits
(inference-engine/src/transformations/src/transformations/serialize.cpp; reader:
inference-engine/src/readers/ir_reader/ie_ir_parser.cpp): layers with <data> attributes, numbered
ports (inputs first), Const layers pointing into the .bin by offset/size, an <edges> table.
Attributes that the C ABI carries as node attributes but nGraph carries as constant inputs
(Reshape target shape, Transpose order, reduction axes, TopK k, ...) become Const layers here.
"""
import numpy as np

from .graph import DT_NAME
from . import _capi as capi

_PREC = {capi.F32: "FP32", capi.F16: "FP16", capi.BF16: "BF16", capi.I32: "I32", capi.U8: "U8", capi.I64: "I64", capi.I8: "I8",
         capi.BOOL: "BOOL"}
_NP_PREC = {"float32": ("f32", "FP32"), "int64": ("i64", "I64"), "int32": ("i32", "I32"), "uint8": ("u8", "U8"), "float16": ("f16", "FP16")}
_TYPE = {"Constant": "Const", "Relu": "ReLU"}


def _fmt(v):
    if isinstance(v, (list, tuple, np.ndarray)):
        return ", ".join(_fmt(x) for x in v)
    if isinstance(v, (bool, np.bool_)):
        return "true" if v else "false"
    if isinstance(v, (float, np.floating)):
        return repr(float(v))
    return str(v)


def _port(pid, prec, shape):
    dims = "".join("\n\t\t\t\t\t<dim>%d</dim>" % d for d in shape)
    return '\t\t\t\t<port id="%d" precision="%s">%s\n\t\t\t\t</port>\n' % (pid, prec, dims) if len(shape) else \
        '\t\t\t\t<port id="%d" precision="%s" />\n' % (pid, prec)


def serialize(func, xml_path, bin_path=None):
    bin_path = bin_path or xml_path[:-4] + ".bin"
    blob = bytearray()
    layers, edges = [], []
    next_id = [0]
    out_port = {}   # (value id) -> (layer id, port id)

    def add_const(arr, name):
        arr = np.asarray(arr, order="C")   # keeps rank 0 (ascontiguousarray would promote scalars to rank 1)
        et, prec = _NP_PREC[str(arr.dtype)]
        off = len(blob)
        blob.extend(arr.tobytes())
        while len(blob) % 8:
            blob.append(0)
        lid = next_id[0]
        next_id[0] += 1
        shape = arr.shape
        s = '\t\t<layer id="%d" name="%s" type="Const" version="opset1">\n' % (lid, name)
        s += '\t\t\t<data element_type="%s" shape="%s" offset="%d" size="%d" />\n' % (et, _fmt(list(shape)), off, arr.nbytes)
        s += "\t\t\t<output>\n" + _port(0, prec, shape) + "\t\t\t</output>\n\t\t</layer>\n"
        layers.append(s)
        return lid, 0, prec, shape

    for n in func.nodes:
        if n.op == "Constant":
            lid, pid, _, _ = add_const(n.data, n.name)
            out_port[n.outputs[0].vid] = (lid, pid)
            continue
        # synthesise constant inputs nGraph expects
        extra = [add_const(c, "%s/const%d" % (n.name, k)) for k, c in enumerate(n.ir_consts)]
        lid = next_id[0]
        next_id[0] += 1
        ltype = _TYPE.get(n.op, n.op)
        s = '\t\t<layer id="%d" name="%s" type="%s" version="%s">\n' % (lid, n.name, ltype, n.version)
        attrs = {k: v for k, v in n.attrs.items()}
        if attrs:
            s += "\t\t\t<data " + " ".join('%s="%s"' % (k, _fmt(v)) for k, v in attrs.items()) + " />\n"
        pid = 0
        if n.inputs or extra:
            s += "\t\t\t<input>\n"
            for x in n.inputs:
                s += _port(pid, _PREC[x.dtype], x.shape)
                src = out_port[x.vid]
                edges.append((src[0], src[1], lid, pid))
                pid += 1
            for (clid, cpid, cprec, cshape) in extra:
                s += _port(pid, cprec, cshape)
                edges.append((clid, cpid, lid, pid))
                pid += 1
            s += "\t\t\t</input>\n"
        if n.op != "Result":
            s += "\t\t\t<output>\n"
            for o in n.outputs:
                s += _port(pid, _PREC[o.dtype], o.shape)
                out_port[o.vid] = (lid, pid)
                pid += 1
            s += "\t\t\t</output>\n"
        s += "\t\t</layer>\n"
        layers.append(s)

    xml = '<?xml version="1.0"?>\n<net name="%s" version="10">\n\t<layers>\n' % func.name
    xml += "".join(layers)
    xml += "\t</layers>\n\t<edges>\n"
    xml += "".join('\t\t<edge from-layer="%d" from-port="%d" to-layer="%d" to-port="%d" />\n' % e for e in edges)
    xml += "\t</edges>\n</net>\n"
    with open(xml_path, "w") as fh:
        fh.write(xml)
    with open(bin_path, "wb") as fh:
        fh.write(bytes(blob) if blob else b"\0" * 8)
    return xml_path, bin_path
