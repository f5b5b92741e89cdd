#!/bin/bash
# A/B of the extrapolation order of the PCG start (HP_EXTRAP = 0..3) on the N=1 bench workload, 20 and 60 steps
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "extrap or checkpoint or transient or run_host" > gpurun_out/ex_tests.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/ex_tests.log
for steps in 20 60; do
for v in 1 2 3; do
  HP_EXTRAP=$v timeout 300 python bench.py --steps $steps --warmup 3 --no-cpu-baseline > gpurun_out/ex_bench_${v}_$steps.json 2> gpurun_out/ex_bench_${v}_$steps.err; echo "bench order=$v steps=$steps rc=$?"
  python - <<PY
import json
j=json.loads([x for x in open('gpurun_out/ex_bench_${v}_$steps.json') if x.startswith('{')][-1])
print($v, $steps, 'ms/step %.3f value %.4e e2e %.4e its/step %.2f'%(j['ms_per_step'], j['value'], j['e2e']['value'], j['config']['pcg_iterations_per_step']))
PY
done; done
