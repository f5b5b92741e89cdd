#!/bin/bash
# A/B of the fused coarse up-leg kernel (HP_MG_UPFUSE=0 disables it)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "mgline or extrap or full or transient" > gpurun_out/up_tests.log 2>&1; echo "tests rc=$?" 
tail -3 gpurun_out/up_tests.log
for v in 0 1; do
  HP_MG_UPFUSE=$v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/up_bench_$v.json 2> gpurun_out/up_bench_$v.err; echo "bench upfuse=$v rc=$?"
  python - <<PY
import json
l=[x for x in open('gpurun_out/up_bench_$v.json') if x.startswith('{')]
j=json.loads(l[-1]); print($v, j['ms_per_step'], j['value'], j['e2e']['value'], j['gpu_launches'], j['config'].get('iters_per_step'))
PY
done
