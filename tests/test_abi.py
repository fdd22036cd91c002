"""The C-ABI boundary: libb200kernels.so loads without a GPU and exports every symbol include/b200.h
declares; the ctypes binding covers exactly that set; device-free entry points behave."""
import ctypes as C
import os
import re
import subprocess

import pytest

from dldt_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", src)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert len(syms) >= 35 and "b200_compile" in syms and "b200_request_start" in syms and "b200_plan_arena" in syms


def test_library_exports_every_declared_symbol():
    lib = C.CDLL(_capi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_ctypes_binding_matches_header():
    assert sorted(_capi.SYMBOLS) == declared_symbols()


def test_only_c_abi_symbols_are_exported():
    out = subprocess.run(["nm", "-D", "--defined-only", _capi.LIB_PATH], capture_output=True, text=True).stdout
    names = [ln.split()[-1] for ln in out.splitlines() if " T " in ln]
    foreign = [n for n in names if not n.startswith("b200_")]
    assert not foreign, foreign[:10]


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "b200.h"\nint main(void) { b200_node_desc d; (void)d; return B200_ABI_VERSION == 1 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_struct_sizes_match_header(tmp_path):
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "b200.h"\nint main(void) { printf("%zu %zu %zu %zu\\n", sizeof(b200_node_desc), '
                   'sizeof(b200_tensor_info), sizeof(b200_compile_options), sizeof(b200_exec_node_info)); return 0; }\n')
    exe = tmp_path / "s"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes == [C.sizeof(_capi.NodeDesc), C.sizeof(_capi.TensorInfo), C.sizeof(_capi.CompileOptions), C.sizeof(_capi.ExecNodeInfo)]


def test_version_and_device_count_without_gpu():
    assert _capi.lib.b200_abi_version() == 1
    assert _capi.device_count() >= 0


def test_no_cpu_fallback_on_missing_device():
    """Creating a request on a network compiled without a device must fail loudly (no CPU path)."""
    from dldt_b200 import Core, Function
    import numpy as np
    if _capi.device_count() > 0:
        pytest.skip("GPU present")
    f = Function("t")
    x = f.parameter((1, 8, 4, 4))
    f.result(f.relu(x))
    with pytest.raises(_capi.B200Error) as e:
        Core().load_network(f, "B200")
    assert e.value.code == _capi.ERR_NO_DEVICE
    net = Core().load_network(f, "B200", dry=True)
    with pytest.raises(_capi.B200Error) as e:
        net.create_infer_request()
    assert e.value.code == _capi.ERR_NO_DEVICE


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "dldt_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dp, fn), errors="ignore").read()
                assert "import oracle" not in txt and "from oracle" not in txt, os.path.join(dp, fn)
