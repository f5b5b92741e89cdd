#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
echo "== bench"; python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/q.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['all_kernels']
print('ms/step %.3f value %.4e e2e %.4e its %.2f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['pcg_iterations_per_step']))
print({n:(k[n]['launches'], round(k[n]['avg_ms']*1e3,1)) for n in k})
print({x:d['roofline'][x] for x in ('kernel','achieved','frac','traffic','share_of_step')})" || tail -5 gpurun_out/q.err
echo "== ensemble"; python bench.py --workload ensemble:1024 --steps 10 --warmup 3 --precond rline 2> gpurun_out/q2.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ensemble1024: ms/step %.3f value %.4e its %.1f' % (d['ms_per_step'], d['value'], d['config']['pcg_iterations_per_step']))" || tail -5 gpurun_out/q2.err
