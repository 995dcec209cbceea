"""cramjam_b200 — B200-native batched block-codec engine behind the cramjam snappy / lz4 / zstd API.

`cramjam_b200._capi` binds the extern "C" boundary (include/cramjam_cuda.h); the codec kernels are
hand-written sm_100a CUDA in cramjam_b200/csrc/.  No CPU fallback exists on this path.
"""
__version__ = "0.1.0"
