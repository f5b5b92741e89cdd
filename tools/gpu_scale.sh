#!/bin/bash
# scaling sweep on an 8-GPU box: 4-GPU mgline slab tests, then the 16384x4096 slab bench at N = 8 (mgline, rline), 4
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== pytest slabs (world 4, mgline)"; timeout 600 python -m pytest tests/test_gpu_slabs.py -q -m gpu -x -k "mgline and 4-" 2>&1 | tail -5
run() { N=$1; pc=$2; extra=$3
  echo "== bench N=$N precond=$pc"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --precond $pc $extra > gpurun_out/scale_n${N}_$pc.json 2> gpurun_out/scale_n${N}_$pc.err
  echo rc=$?; grep -v "^\*\|OMP_NUM\|^$\|NCCL version" gpurun_out/scale_n${N}_$pc.err | tail -4
  python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/scale_n${N}_$pc.json') if l.startswith('{')][0])
    print('N',d['n_gpus'],'ms/step %.1f'%d['ms_per_step'],'value %.3e'%d['value'],'e2e %.3e'%d['e2e']['value'],'its',d['config']['pcg_iterations_per_step'],'clocks',d['clocks'])
    print('   strong', d.get('strong_scaling'))
except Exception as e: print('no json',e)
PY
}
run 8 mgline ""; run 8 rline "--no-single-gpu-ref"; run 4 mgline "--no-single-gpu-ref"
