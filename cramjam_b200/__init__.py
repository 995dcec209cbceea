"""cramjam_b200 — B200-native batched block-codec engine behind the cramjam snappy / lz4 / zstd API.

  cramjam_b200.cramjam   C++ (pybind11) host binding with the reference's Python surface for this path:
                         snappy / lz4 / zstd submodules, Buffer, File, CompressionError, DecompressionError
  cramjam_b200._capi     ctypes binding of the extern "C" boundary (include/cramjam_cuda.h), incl. the batch API

The codec kernels are hand-written sm_100a CUDA in cramjam_b200/csrc/.  No CPU fallback exists on this path:
a missing library raises ImportError, a missing CUDA device raises when the first codec call is made.
"""
__version__ = "0.1.0"


def __getattr__(name):
    if name == "cramjam":
        import importlib
        import os
        import sys
        here = os.path.dirname(os.path.abspath(__file__))
        if not any(f.startswith("cramjam.") and f.endswith(".so") for f in os.listdir(here)):
            raise ImportError("cramjam_b200/cramjam.*.so is missing: run `python -m cramjam_b200.build` "
                              "(builds libcramjam_cuda.so with nvcc for sm_100a and the C++ host binding). No CPU fallback exists.")
        mod = importlib.import_module("cramjam_b200.cramjam")
        for sub in ("snappy", "lz4", "zstd"):
            sys.modules.setdefault(f"cramjam_b200.cramjam.{sub}", getattr(mod, sub))
        return mod
    raise AttributeError(name)
