#!/bin/bash
mkdir -p gpurun_out
echo "== pytest mgline"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "mgline" 2>&1 | tail -25
echo "== pytest all gpu"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -6
for pc in mgline rline; do
echo "== bench precond=$pc"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --precond $pc > gpurun_out/bench_$pc.json 2> gpurun_out/bench_$pc.err; echo rc=$?; tail -3 gpurun_out/bench_$pc.err
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_$pc.json') if l.startswith('{')][0])
print('ms/step',d['ms_per_step'],'value %.4e'%d['value'],'e2e %.4e'%d['e2e']['value'],'its',d['config']['pcg_iterations_per_step'])
print([ (s['iters'], '%.2e'%s['residual_l2'], '%.1e'%s['crit']) for s in d['last_sweep_stats']])
for k,x in d['roofline']['all_kernels'].items(): print('  ',k,x)
PY
done
