#!/bin/bash
mkdir -p gpurun_out
python -c "
import torch
p=torch.cuda.get_device_properties(0); print('L2', p.L2_cache_size)
import ctypes; rt=ctypes.CDLL('libcudart.so.12'); v=ctypes.c_int(); rt.cudaDeviceGetAttribute(ctypes.byref(v), 108, 0); print('persistingL2CacheMaxSize', v.value)"
for mb in 32 64 96; do for pol in 4 5; do echo "POL=$pol persist=$mb MB"; CJ_L2_PERSIST_MB=$mb CJ_LIB_PATH=$PWD/variants/lib_pol$pol.so SWEEP_ONLY=7:3 timeout 300 python tools/g7_sweep.py 65536 snappy --oracle 2>&1 | tail -1; done; done
