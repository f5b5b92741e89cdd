#!/bin/bash
# scaling sweep on an 8-GPU box: slab tests, then the 16384x4096 slab bench at N = 8 (p2p, nccl), 4, 2
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== pytest slabs"; timeout 900 python -m pytest tests/test_gpu_slabs.py -q -m gpu -x 2>&1 | tail -5
run() { N=$1; comm=$2
  echo "== bench N=$N comm=$comm"
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --comm $comm > gpurun_out/scale_n${N}_$comm.json 2> gpurun_out/scale_n${N}_$comm.err
  echo rc=$?; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/scale_n${N}_$comm.err | tail -4
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/scale_n${N}_$comm.json'))
    print('N',d['n_gpus'],'ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'its',d['config']['pcg_iterations_per_step'],'clocks',d['clocks'])
except Exception as e: print('no json',e)
PY
}
run 8 p2p; run 8 nccl; run 4 p2p; run 2 p2p
