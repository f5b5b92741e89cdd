#!/bin/bash
# V-cycle shape sweep on the N=1 bench workload: "levels coarse_sweeps omega" per line
mkdir -p gpurun_out
for cfg in ${CFGS:-"6 2 0.7" "7 2 0.7" "8 2 0.7" "7 1 0.7" "7 3 0.7" "6 3 0.7" "6 1 0.7" "7 2 0.8"}; do
set -- $cfg
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-single-gpu-ref --mg-levels $1 --mg-sweeps $2 --mg-omega $3 > gpurun_out/tune.json 2> gpurun_out/tune.err || tail -3 gpurun_out/tune.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/tune.json') if l.startswith('{')][0])
print('levels $1 sweeps $2 omega $3: ms/step %.3f value %.3e its %.1f launches %d'%(d['ms_per_step'],d['value'],d['config']['pcg_iterations_per_step'],d['gpu_launches']))
PY
done
