#!/bin/bash
mkdir -p gpurun_out
echo "== mgline tests (fused)"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "mgline or full_size" 2>&1 | tail -8
for f in 1 0; do
HP_MG_FUSED=$f timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/fused_$f.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); k=d['roofline']['all_kernels']
print('HP_MG_FUSED=$f ms/step %.3f value %.4e its %.2f' % (d['ms_per_step'], d['value'], d['config']['pcg_iterations_per_step']), {n:(k[n]['launches'], round(k[n]['avg_ms']*1e3,1)) for n in k if n.startswith('mg')})" || tail -5 gpurun_out/fused_$f.err
done
echo "== all gpu tests"; timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
