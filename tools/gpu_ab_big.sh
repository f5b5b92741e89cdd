#!/bin/bash
mkdir -p gpurun_out
for v in $VALS; do
  echo "== $VAR=$v"
  env $VAR=$v timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload ${WL:-16384x4096} > gpurun_out/abb_$v.json 2> gpurun_out/abb_$v.err; echo rc=$?; tail -2 gpurun_out/abb_$v.err
  python - <<PY
import json
d=json.load(open('gpurun_out/abb_$v.json'))
print('ms/step',d['ms_per_step'],'value',d['value'],'its',d['config']['pcg_iterations_per_step'])
for k,x in d['roofline']['all_kernels'].items(): print('  ',k,x)
PY
done
