#!/bin/bash
# A/B of environment switches on the bench (kernel-level numbers only):  VAR=name VALS="a b" bash tools/gpu_ab.sh
mkdir -p gpurun_out
timeout 300 python -m pytest tests -q -m gpu -x 2>&1 | tail -3
for v in $VALS; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single-gpu-ref > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err; echo rc=$?; tail -2 gpurun_out/ab_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/ab_$v.json'))
print('ms/step',d['ms_per_step'],'value',d['value'],'its',d['config']['pcg_iterations_per_step'])
for k,x in d['roofline']['all_kernels'].items(): print('  ',k,x)
PY
done
