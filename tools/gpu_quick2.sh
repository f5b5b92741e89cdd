#!/bin/bash
# quick validation: parity tests (subset via $K) + one bench line with the per-kernel table
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu ${K:+-k "$K"} > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/q_tests.log
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline ${BENCH_ARGS} > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads([x for x in open('gpurun_out/q_bench.json') if x.startswith('{')][-1])
print(j['ms_per_step'], j['value'], j['e2e']['value'], j['gpu_launches'])
for k,d in j['roofline']['all_kernels'].items():
    print(f"  {k:10s} n={d['launches']:5d} avg={d['avg_ms']*1e3:8.2f}us total={d['launches']*d['avg_ms']:8.3f}ms")
PY
