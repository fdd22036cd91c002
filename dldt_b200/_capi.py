"""ctypes binding of the C ABI declared in include/b200.h (libb200kernels.so).

The library is the product: there is no Python/CPU fallback.  Importing this module without the
built shared object raises ImportError with the build command.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200kernels.so")

MAX_RANK, MAX_INPUTS, MAX_IATTR, MAX_FATTR = 8, 16, 40, 8

# b200_status
OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_NO_DEVICE, ERR_BUSY, ERR_OOM, NOT_READY = 0, -1, -2, -3, -4, -5, -6, 1
# b200_dtype
F32, F16, BF16, I32, U8, I64, I8, BOOL = range(8)
# b200_precision
PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2
LAYOUT_ANY, LAYOUT_NCHW, LAYOUT_NHWC = 0, 1, 2

OPS = ["Parameter", "Constant", "Result", "Convolution", "GroupConvolution", "BatchNormInference", "Add", "Subtract",
       "Multiply", "Divide", "Maximum", "Minimum", "Relu", "Sigmoid", "Tanh", "Clamp", "Swish", "Gelu", "HSwish",
       "HSigmoid", "Mish", "Exp", "Sqrt", "Abs", "Negative", "Erf", "Elu", "MaxPool", "AvgPool", "ReduceMean", "MatMul",
       "Softmax", "MVN", "Reshape", "Transpose", "Concat", "Split", "Gather", "TopK", "NonMaxSuppression",
       "DetectionOutput", "Convert", "Power", "PRelu", "ReduceSum", "ReduceMax", "SquaredDifference", "Round", "FakeQuantize", "StridedSlice"]
OP_ID = {n: i for i, n in enumerate(OPS)}


class NodeDesc(C.Structure):
    _fields_ = [("op", C.c_int32), ("name", C.c_char_p), ("n_inputs", C.c_int32), ("inputs", C.c_int32 * MAX_INPUTS),
                ("i", C.c_int32 * MAX_IATTR), ("f", C.c_float * MAX_FATTR), ("dtype", C.c_int32), ("rank", C.c_int32),
                ("shape", C.c_int64 * MAX_RANK), ("data", C.c_void_p), ("data_bytes", C.c_uint64)]


class TensorInfo(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("dtype", C.c_int32), ("rank", C.c_int32), ("shape", C.c_int64 * MAX_RANK)]


class CompileOptions(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("precision", C.c_int32), ("conv_impl", C.c_int32),
                ("reserved", C.c_int32 * 12)]


class ExecNodeInfo(C.Structure):
    _fields_ = [("name", C.c_char * 256), ("type", C.c_char * 64), ("impl", C.c_char * 64), ("fused", C.c_char * 1024),
                ("out_precision", C.c_char * 16), ("out_layout", C.c_char * 16), ("exec_order", C.c_int32),
                ("flops", C.c_double), ("bytes", C.c_double)]


# every symbol include/b200.h declares: (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "b200_abi_version": (C.c_int, []),
    "b200_last_error": (C.c_char_p, []),
    "b200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "b200_device_name": (C.c_int, [C.c_int, C.c_char_p, C.c_size_t]),
    "b200_device_props": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_uint64)]),
    "b200_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "b200_host_free": (C.c_int, [_P]),
    "b200_host_register": (C.c_int, [_P, C.c_size_t]),
    "b200_host_unregister": (C.c_int, [_P]),
    "b200_graph_create": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "b200_graph_destroy": (None, [_P]),
    "b200_graph_add": (C.c_int, [_P, C.POINTER(NodeDesc), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "b200_graph_value_info": (C.c_int, [_P, C.c_int32, C.POINTER(TensorInfo)]),
    "b200_op_supported": (C.c_int, [C.c_int32]),
    "b200_compile": (C.c_int, [_P, C.POINTER(CompileOptions), C.POINTER(_P)]),
    "b200_network_destroy": (None, [_P]),
    "b200_network_num_inputs": (C.c_int, [_P]),
    "b200_network_num_outputs": (C.c_int, [_P]),
    "b200_network_io_info": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(TensorInfo)]),
    "b200_network_num_exec_nodes": (C.c_int, [_P]),
    "b200_network_exec_node": (C.c_int, [_P, C.c_int, C.POINTER(ExecNodeInfo)]),
    "b200_network_stat": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_double)]),
    "b200_device_alloc": (C.c_int, [C.c_int, C.POINTER(_P), C.c_size_t]),
    "b200_device_free": (C.c_int, [C.c_int, _P]),
    "b200_device_copy": (C.c_int, [C.c_int, _P, _P, C.c_size_t, C.c_int]),
    "b200_network_tune": (C.c_int, [_P]),
    "b200_network_get_tuning": (C.c_int, [_P, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "b200_network_set_tuning": (C.c_int, [_P, C.c_char_p]),
    "b200_plan_arena": (C.c_int, [C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int64),
                                  C.POINTER(C.c_int64)]),
    "b200_request_create": (C.c_int, [_P, C.POINTER(_P)]),
    "b200_request_destroy": (None, [_P]),
    "b200_request_set_io_format": (C.c_int, [_P, C.c_int, C.c_int, C.c_int32, C.c_int32]),
    "b200_request_io_bytes": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_uint64)]),
    "b200_request_start": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    "b200_request_wait": (C.c_int, [_P, C.c_int64]),
    "b200_request_device_io": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "b200_request_launch": (C.c_int, [_P]),
    "b200_request_time": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_float)]),
    "b200_request_perf_counts": (C.c_int, [_P, C.POINTER(C.c_float), C.c_int]),
    "b200_request_num_launches": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "b200_request_stream": (C.c_int, [_P, C.POINTER(_P)]),
    "b200_request_dump_node": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), C.c_uint64]),
}


class B200Error(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("b200 error %d: %s" % (code, msg))
        self.code = code


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `make -C dldt_b200/csrc` (or __graft_entry__.build()); "
                          "there is no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    if lib.b200_abi_version() != 1:
        raise ImportError("libb200kernels.so ABI version mismatch")
    return lib


lib = _load()


def check(code):
    if code < 0:
        raise B200Error(code, lib.b200_last_error().decode("utf-8", "replace"))
    return code


def device_count():
    n = C.c_int(0)
    check(lib.b200_device_count(C.byref(n)))
    return n.value
