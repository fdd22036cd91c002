"""Python mirror of the Inference-Engine objects the B200 plugin implements
(Core::LoadNetwork -> ExecutableNetwork -> InferRequest::StartAsync/Wait), driving the same C ABI
the C++ plugin (plugin/src) drives.  Names, config keys, metrics and error behaviour follow
inference-engine/include/ie_core.hpp, cpp/ie_executable_network.hpp, cpp/ie_infer_request.hpp and
ie_plugin_config.hpp of the reference.  All compute happens in libb200kernels.so.
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from .graph import DT_NP, Function

DEVICE_NAME = "B200"
_PREC = {"FP32": capi.PREC_FP32, "BF16": capi.PREC_BF16, "FP16": capi.PREC_FP16}
_BLOB_DT = {"FP32": capi.F32, "FP16": capi.F16, "BF16": capi.BF16, "I32": capi.I32, "U8": capi.U8}
_DT_BLOB_NP = {capi.F32: np.float32, capi.F16: np.float16, capi.I32: np.int32, capi.U8: np.uint8, capi.BF16: np.uint16}

SUPPORTED_CONFIG_KEYS = ["DEVICE_ID", "PERF_COUNT", "B200_INFERENCE_PRECISION", "B200_THROUGHPUT_STREAMS"]


class NotFound(KeyError):
    """IE_THROW(NotFound): unknown config key / device (config.cpp:112-114)."""


class RequestBusy(RuntimeError):
    """StatusCode::REQUEST_BUSY (ie_common.h:231-249)."""


class PinnedArray:
    """numpy view over cudaHostAlloc'ed memory (the plugin's default blob allocator)."""

    def __init__(self, shape, dtype):
        self.nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self._ptr = C.c_void_p()
        capi.check(capi.lib.b200_host_alloc(C.byref(self._ptr), max(self.nbytes, 16)))
        buf = (C.c_char * max(self.nbytes, 16)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    @property
    def ptr(self):
        return self._ptr.value

    def __del__(self):
        if getattr(self, "_ptr", None) and self._ptr.value and capi is not None and capi.lib is not None:
            self.array = None
            capi.lib.b200_host_free(self._ptr)
            self._ptr = C.c_void_p()


class InferRequest:
    """IInferRequestInternal + AsyncInferRequestThreadSafeDefault for one request."""

    def __init__(self, exec_net, input_precisions=None, input_layouts=None):
        self._net = exec_net
        self._h = C.c_void_p()
        capi.check(capi.lib.b200_request_create(exec_net._h, C.byref(self._h)))
        self._busy = False
        self.input_blobs, self.output_blobs = {}, {}
        self._in_pinned, self._out_pinned = [], []
        for idx, (name, shape, gdt) in enumerate(exec_net.inputs):
            prec = (input_precisions or {}).get(name)
            dt = _BLOB_DT[prec] if prec else (capi.I32 if gdt in (capi.I32, capi.I64) else (capi.U8 if gdt == capi.U8 else capi.F32))
            layout = {"NHWC": capi.LAYOUT_NHWC}.get((input_layouts or {}).get(name, "NCHW"), capi.LAYOUT_NCHW)
            capi.check(capi.lib.b200_request_set_io_format(self._h, 0, idx, dt, layout))
            bshape = shape if layout != capi.LAYOUT_NHWC else (shape[0], shape[2], shape[3], shape[1])
            p = PinnedArray(bshape, _DT_BLOB_NP[dt])
            self._in_pinned.append(p)
            self.input_blobs[name] = p.array
        for idx, (name, shape, gdt) in enumerate(exec_net.outputs):
            dt = capi.I32 if gdt in (capi.I32, capi.I64) else capi.F32
            capi.check(capi.lib.b200_request_set_io_format(self._h, 1, idx, dt, capi.LAYOUT_NCHW))
            p = PinnedArray(shape, _DT_BLOB_NP[dt])
            self._out_pinned.append(p)
            self.output_blobs[name] = p.array
        self._in_ptrs = (C.c_void_p * max(len(self._in_pinned), 1))(*[p.ptr for p in self._in_pinned])
        self._out_ptrs = (C.c_void_p * max(len(self._out_pinned), 1))(*[p.ptr for p in self._out_pinned])

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None and capi.lib is not None:
            capi.lib.b200_request_destroy(self._h)
            self._h = None

    def set_blob(self, name, array):
        dst = self.input_blobs[name]
        if array.shape != dst.shape:
            raise ValueError("blob '%s': shape %s does not match %s" % (name, array.shape, dst.shape))
        np.copyto(dst, array.astype(dst.dtype, copy=False))

    def start_async(self):
        if self._busy:
            raise RequestBusy("request is busy")
        capi.check(capi.lib.b200_request_start(self._h, self._in_ptrs, self._out_ptrs))
        self._busy = True

    def wait(self, timeout_ms=-1):
        rc = capi.check(capi.lib.b200_request_wait(self._h, timeout_ms))
        if rc == capi.NOT_READY:
            return "RESULT_NOT_READY"
        self._busy = False
        return "OK"

    def infer(self, inputs=None):
        for k, v in (inputs or {}).items():
            self.set_blob(k, np.asarray(v))
        self.start_async()
        self.wait()
        return {k: v.copy() for k, v in self.output_blobs.items()}

    # device-resident path used by bench.py `value` and the profiler runs
    def launch_device(self):
        capi.check(capi.lib.b200_request_launch(self._h))

    def device_io(self, is_output, index):
        p, n = C.c_void_p(), C.c_uint64()
        capi.check(capi.lib.b200_request_device_io(self._h, int(is_output), index, C.byref(p), C.byref(n)))
        return p.value, n.value

    def time_launches(self, iters, end_to_end=False):
        ms = C.c_float()
        capi.check(capi.lib.b200_request_time(self._h, iters, 1 if end_to_end else 0, self._in_ptrs, self._out_ptrs, C.byref(ms)))
        return ms.value

    def get_perf_counts(self):
        n = self._net.num_exec_nodes
        us = (C.c_float * max(n, 1))()
        capi.check(capi.lib.b200_request_perf_counts(self._h, us, n))
        out = {}
        for k, info in enumerate(self._net.get_exec_graph_info()):
            out[info["name"]] = {"status": "EXECUTED", "real_time": us[k], "cpu_time": us[k], "exec_type": info["impl"],
                                 "layer_type": info["type"], "execution_index": k}
        return out

    def num_launches(self):
        n = C.c_int()
        capi.check(capi.lib.b200_request_num_launches(self._h, C.byref(n)))
        return n.value

    def dump_node(self, index, count):
        out = np.empty(count, np.float32)
        capi.check(capi.lib.b200_request_dump_node(self._h, index, out.ctypes.data_as(C.POINTER(C.c_float)), count))
        return out


class ExecutableNetwork:
    def __init__(self, func, device_id, precision, config, conv_impl=0, dry=False):
        self.function = func
        self.config = dict(config)
        self._h = C.c_void_p()
        opts = capi.CompileOptions()
        opts.abi_version = 1
        opts.device = -1 if dry else device_id
        opts.precision = precision
        opts.conv_impl = conv_impl
        capi.check(capi.lib.b200_compile(func._h, C.byref(opts), C.byref(self._h)))
        info = capi.TensorInfo()
        self.inputs, self.outputs = [], []
        for is_out, lst, n in ((0, self.inputs, capi.lib.b200_network_num_inputs(self._h)), (1, self.outputs, capi.lib.b200_network_num_outputs(self._h))):
            for k in range(n):
                capi.check(capi.lib.b200_network_io_info(self._h, is_out, k, C.byref(info)))
                lst.append((info.name.decode(), tuple(info.shape[r] for r in range(info.rank)), info.dtype))
        self.num_exec_nodes = capi.lib.b200_network_num_exec_nodes(self._h)

    def __del__(self):
        if getattr(self, "_h", None) and capi is not None and capi.lib is not None:
            capi.lib.b200_network_destroy(self._h)
            self._h = None

    def create_infer_request(self, input_precisions=None, input_layouts=None):
        return InferRequest(self, input_precisions, input_layouts)

    def stat(self, key):
        v = C.c_double()
        capi.check(capi.lib.b200_network_stat(self._h, key.encode(), C.byref(v)))
        return v.value

    def get_metric(self, name):
        if name == "NETWORK_NAME":
            return self.function.name
        if name == "OPTIMAL_NUMBER_OF_INFER_REQUESTS":
            return 2 * int(self.config.get("B200_THROUGHPUT_STREAMS", 1))
        if name == "SUPPORTED_METRICS":
            return ["NETWORK_NAME", "SUPPORTED_METRICS", "SUPPORTED_CONFIG_KEYS", "OPTIMAL_NUMBER_OF_INFER_REQUESTS"]
        if name == "SUPPORTED_CONFIG_KEYS":
            return list(SUPPORTED_CONFIG_KEYS)
        raise NotFound(name)

    def get_config(self, key):
        if key not in SUPPORTED_CONFIG_KEYS:
            raise NotFound(key)
        return self.config.get(key)

    def get_exec_graph_info(self):
        out, info = [], capi.ExecNodeInfo()
        for k in range(self.num_exec_nodes):
            capi.check(capi.lib.b200_network_exec_node(self._h, k, C.byref(info)))
            out.append({"name": info.name.decode(), "type": info.type.decode(), "impl": info.impl.decode(),
                        "fused": info.fused.decode().split(","), "out_precision": info.out_precision.decode(),
                        "out_layout": info.out_layout.decode(), "exec_order": info.exec_order, "flops": info.flops, "bytes": info.bytes})
        return out


class Core:
    """InferenceEngine::Core restricted to the one device this repository implements."""

    def __init__(self):
        self._config = {"DEVICE_ID": "0", "PERF_COUNT": "NO", "B200_INFERENCE_PRECISION": "BF16", "B200_THROUGHPUT_STREAMS": "1"}

    def get_available_devices(self):
        n = capi.device_count()
        return ["%s.%d" % (DEVICE_NAME, i) for i in range(n)]

    def set_config(self, config, device_name=DEVICE_NAME):
        for k, v in config.items():
            if k not in SUPPORTED_CONFIG_KEYS:
                raise NotFound("Unsupported property %s by %s plugin" % (k, DEVICE_NAME))
            if k == "B200_INFERENCE_PRECISION" and str(v).upper() not in _PREC:
                raise ValueError("B200_INFERENCE_PRECISION must be one of FP32, BF16, FP16")
            self._config[k] = str(v)

    def get_metric(self, device_name, name):
        if name == "AVAILABLE_DEVICES":
            return [str(i) for i in range(capi.device_count())]
        if name == "FULL_DEVICE_NAME":
            buf = C.create_string_buffer(256)
            capi.check(capi.lib.b200_device_name(int(self._config["DEVICE_ID"]), buf, 256))
            return buf.value.decode()
        if name == "OPTIMIZATION_CAPABILITIES":
            return ["FP32", "BF16", "FP16"]
        if name == "SUPPORTED_CONFIG_KEYS":
            return list(SUPPORTED_CONFIG_KEYS)
        if name == "RANGE_FOR_ASYNC_INFER_REQUESTS":
            return (1, 16, 1)
        if name == "DEVICE_ARCHITECTURE":
            return "sm_100a"
        raise NotFound(name)

    def load_network(self, func, device_name=DEVICE_NAME, config=None, dry=False):
        """device_name: "B200" or "B200.<id>" (ie_core.cpp:45-66 splits the id into DEVICE_ID)."""
        if not isinstance(func, Function):
            raise TypeError("load_network expects a dldt_b200.graph.Function")
        base, _, dev = device_name.partition(".")
        if base != DEVICE_NAME:
            raise NotFound('Device with "%s" name is not registered in the InferenceEngine' % base)
        cfg = dict(self._config)
        for k, v in (config or {}).items():
            if k not in SUPPORTED_CONFIG_KEYS and k != "B200_CONV_IMPL":
                raise NotFound("Unsupported property %s by %s plugin" % (k, DEVICE_NAME))
            cfg[k] = str(v)
        if dev:
            cfg["DEVICE_ID"] = dev
        prec = _PREC[cfg["B200_INFERENCE_PRECISION"].upper()]
        return ExecutableNetwork(func, int(cfg["DEVICE_ID"]), prec, cfg, conv_impl=int(cfg.get("B200_CONV_IMPL", 0)), dry=dry)
