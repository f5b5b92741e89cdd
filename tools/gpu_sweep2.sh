#!/bin/bash
# A/B of the predicted start of the second sweep (HP_SWEEP2=0 disables the c2 tracking)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > gpurun_out/s2_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/s2_tests.log
for steps in 20 60; do
for v in 0 1; do
  HP_SWEEP2=$v timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/s2_bench_${v}_$steps.json 2> gpurun_out/s2_bench_${v}_$steps.err; echo "bench sweep2=$v steps=$steps rc=$?"
  python - <<PY
import json
j=json.loads([x for x in open('gpurun_out/s2_bench_${v}_$steps.json') if x.startswith('{')][-1])
print($v, $steps, 'ms/step %.3f value %.4e e2e %.4e its/step %.2f'%(j['ms_per_step'], j['value'], j['e2e']['value'], j['config']['pcg_iterations_per_step']), [(s['iters'], '%.1e'%s['residual_l2']) for s in j['last_sweep_stats']])
PY
done; done
