"""liloc_b200 -- B200-native (sm_100a) scan-to-map registration path of LiLoc behind a C ABI.

The compute lives in ``libliloc_gpu.so`` (hand-written CUDA, see ``csrc/``); this package is the thin
Python binding used by the tests and ``bench.py``.  There is no CPU fallback: touching any of the names
below loads ``liloc_b200.abi``, which raises if the shared library has not been built
(``python -m liloc_b200._build``).

The binding is loaded on first use, not at package import: ``bench.py --impl reference`` (the CPU arm) imports
``liloc_b200.parallel`` / ``liloc_b200.synth`` for launcher plumbing and input generation and must not map the
product's shared libraries into its process.
"""
__all__ = ["Context", "LilocError", "S2mOpts", "S2mResult", "NdtOpts", "lib", "lib_path", "POINT_DTYPE",
           "LILOC_OK", "LILOC_SKIPPED"]


def __getattr__(name):
    if name in __all__:
        from . import abi
        return getattr(abi, name)
    raise AttributeError(f"module 'liloc_b200' has no attribute {name!r}")
