"""liloc_b200 -- B200-native (sm_100a) scan-to-map registration path of LiLoc behind a C ABI.

The compute lives in ``libliloc_gpu.so`` (hand-written CUDA, see ``csrc/``); this package is the thin
Python binding used by the tests and ``bench.py``.  There is no CPU fallback: importing
``liloc_b200.abi`` raises if the shared library has not been built (``python -m liloc_b200._build``).
"""
from .abi import (  # noqa: F401
    Context, LilocError, S2mOpts, S2mResult, NdtOpts, lib, lib_path, POINT_DTYPE,
    LILOC_OK, LILOC_SKIPPED,
)

__all__ = ["Context", "LilocError", "S2mOpts", "S2mResult", "NdtOpts", "lib", "lib_path", "POINT_DTYPE",
           "LILOC_OK", "LILOC_SKIPPED"]
