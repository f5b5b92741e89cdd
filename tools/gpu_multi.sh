#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ "${TESTS:-1}" = "1" ]; then
echo "== pytest slabs"; timeout 900 python -m pytest tests/test_gpu_slabs.py -q -m gpu -x 2>&1 | tail -15
fi
for comm in ${COMMS:-p2p nccl}; do
echo "== bench N=$N comm=$comm"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --comm $comm > gpurun_out/bench_n${N}_$comm.json 2> gpurun_out/bench_n${N}_$comm.err
echo rc=$?; grep -v "^\*\|OMP_NUM\|^$" gpurun_out/bench_n${N}_$comm.err | tail -5; cut -c1-1100 gpurun_out/bench_n${N}_$comm.json
done
if [ "${ONE:-0}" = "1" ]; then
echo "== bench N=1 same workload"
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --workload 16384x4096 --no-cpu-baseline > gpurun_out/bench_big_n1.json 2> gpurun_out/bench_big_n1.err
echo rc=$?; tail -3 gpurun_out/bench_big_n1.err; cut -c1-700 gpurun_out/bench_big_n1.json
fi
