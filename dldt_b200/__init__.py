"""dldt_b200: host-side mirror of the B200 Inference-Engine device plugin.

Everything that computes lives in lib/libb200kernels.so (hand-written sm_100a CUDA behind the C
ABI of include/b200.h); this package only builds graphs, binds the ABI and moves numpy buffers.
"""
from . import _capi  # noqa: F401  (fails loudly when the CUDA library is missing)
from .graph import Function  # noqa: F401
from .ie import Core, ExecutableNetwork, InferRequest  # noqa: F401

__all__ = ["Function", "Core", "ExecutableNetwork", "InferRequest"]
