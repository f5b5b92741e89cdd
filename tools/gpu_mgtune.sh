#!/bin/bash
mkdir -p gpurun_out
for cfg in "6 2 0.8" "5 2 0.8" "4 2 0.8" "4 4 0.8" "5 1 0.8" "6 2 0.7" "6 2 0.9" "7 2 0.8"; do
set -- $cfg
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --mg-levels $1 --mg-sweeps $2 --mg-omega $3 > gpurun_out/tune.json 2> gpurun_out/tune.err || tail -3 gpurun_out/tune.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/tune.json') if l.startswith('{')][0])
k=d['roofline']['all_kernels']
print('levels $1 sweeps $2 omega $3: ms/step %.2f value %.3e its %.1f'%(d['ms_per_step'],d['value'],d['config']['pcg_iterations_per_step']), {n:(k[n]['launches'], round(k[n]['avg_ms']*1e3,1)) for n in k if n.startswith('mg') or n in ('cg_A','cg_B')})
PY
done
