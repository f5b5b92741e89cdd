#!/bin/bash
mkdir -p gpurun_out
for f in 1 0; do
HP_PDL=$f timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/pdl_$f.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('HP_PDL=$f ms/step %.3f value %.4e its %.2f' % (d['ms_per_step'], d['value'], d['config']['pcg_iterations_per_step']))" || tail -5 gpurun_out/pdl_$f.err
done
HP_PDL=1 timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --precond rline 2> gpurun_out/pdl_r.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('rline HP_PDL=1 ms/step %.3f value %.4e its %.2f' % (d['ms_per_step'], d['value'], d['config']['pcg_iterations_per_step']))"
echo "== all gpu tests (PDL on)"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
