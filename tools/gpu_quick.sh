#!/bin/bash
# quick validation on one B200: parity tests (subset via $K), one bench line with the per-kernel table, [ensemble line]
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu ${K:+-k "$K"} > gpurun_out/q_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/q_tests.log
timeout 300 python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --no-single-gpu-ref ${BENCH_ARGS} > gpurun_out/q_bench.json 2> gpurun_out/q_bench.err; echo "bench rc=$?"
python - <<PY
import json
j=json.loads([x for x in open('gpurun_out/q_bench.json') if x.startswith('{')][-1])
print('ms/step %.3f value %.4e e2e %.4e its/step %.2f launches %d' % (j['ms_per_step'], j['value'], j['e2e']['value'], j['config']['pcg_iterations_per_step'], j['gpu_launches']))
for k,d in j['roofline']['all_kernels'].items():
    print(f"  {k:10s} launches={d['launches']:5d} full={d['full_launches']:5d} avg(full)={d['avg_ms']*1e3:8.2f}us total={d['total_ms']:8.3f}ms GB/s={d['GBps']:.0f}")
PY
if [ "${ENSEMBLE:-0}" = "1" ]; then
timeout 300 python bench.py --workload ensemble:1024 --steps 10 --warmup 3 --precond rline --no-cpu-baseline 2> gpurun_out/q2.err | cut -c1-400
fi
