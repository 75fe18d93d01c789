"""KataCoffee-B200: B200-native batched Coffee rules, V1 features and residual-net forward.

The product is the C-ABI CUDA library (include/katacoffee_b200.h, built from csrc/ into
libkatacoffee_b200.so); this package is the thin Python host side used by tests and bench.py:
ctypes bindings (capi), the model-description builder (modeldesc) and small wrappers that mirror
the reference's nninterface.h operator names (backend).  There is no CPU fallback anywhere in it.
"""
from . import capi, modeldesc, backend  # noqa: F401

__all__ = ["capi", "modeldesc", "backend"]
