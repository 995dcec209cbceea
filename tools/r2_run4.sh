#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_host_api.py tests/test_gpu_frames.py -m gpu -x -q > gpurun_out/r2_tests4.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_tests4.log
python - <<'PY'
import time, sys
sys.path.insert(0, "tests")
import corpus
from cramjam_b200 import cramjam as cj
import numpy as np
from cramjam_b200 import _capi as capi
buf = capi.synth_host(4096, 65536).tobytes()
for name, mod in (("snappy", cj.snappy), ("lz4", cj.lz4), ("zstd", cj.zstd)):
    c = bytes(mod.compress(buf)); mod.decompress(c)
    t0 = time.perf_counter(); c = bytes(mod.compress(buf)); t1 = time.perf_counter(); d = mod.decompress(c); t2 = time.perf_counter()
    print(f"{name} 256MiB compress {1e3*(t1-t0):.1f} ms decompress {1e3*(t2-t1):.1f} ms", flush=True)
    src = cj.Buffer(buf, pinned=True); comp = cj.Buffer(pinned=True); comp.reserve(len(buf) + (len(buf) >> 3)); back = cj.Buffer(pinned=True); back.reserve(len(buf))
    mod.compress_into(src, comp); mod.decompress_into(comp, back)
    comp.truncate(); back.truncate()
    t0 = time.perf_counter(); mod.compress_into(src, comp); t1 = time.perf_counter(); mod.decompress_into(comp, back); t2 = time.perf_counter()
    assert bytes(back) == buf
    print(f"{name} 256MiB pinned Buffers: compress_into {1e3*(t1-t0):.1f} ms decompress_into {1e3*(t2-t1):.1f} ms", flush=True)
PY
