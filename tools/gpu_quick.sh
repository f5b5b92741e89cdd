#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
for e in 1 0; do
echo "== bench extrapolate=$e"; HP_EXTRAP=$e timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline > gpurun_out/bench_e$e.json 2> gpurun_out/bench_e$e.err; echo rc=$?; tail -3 gpurun_out/bench_e$e.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_e$e.json') if l.startswith('{')][0])
print('ms/step',d['ms_per_step'],'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'its',d['config']['pcg_iterations_per_step'])
print([ (s['iters'], '%.2e'%s['residual_l2']) for s in d['last_sweep_stats']])
PY
done
