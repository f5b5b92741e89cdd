#!/bin/bash
mkdir -p gpurun_out
python - <<'PY'
import ctypes
rt=ctypes.CDLL('libcudart.so.12') if False else None
import torch
torch.cuda.init()
lib=ctypes.CDLL([l for l in open('/proc/self/maps').read().split() if 'libcudart' in l][0]) if any('libcudart' in l for l in open('/proc/self/maps').read().split()) else None
print('cudart', lib)
if lib:
    v=ctypes.c_int()
    for name,a in (('MaxPersistingL2CacheSize',108),('MaxAccessPolicyWindowSize',109),('L2CacheSize',38)):
        lib.cudaDeviceGetAttribute(ctypes.byref(v),a,0); print(name, v.value)
PY
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_ -s 600 -c 4 -f -o gpurun_out/prof_cg2 \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host > gpurun_out/ncu2.log 2>&1
echo rc=$?
