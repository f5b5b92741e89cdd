#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest gpu (all)"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -15
echo "== bench N=$N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 --loop-mode host > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo rc=$?; tail -5 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json | cut -c1-1500
if [ "${ONE:-1}" = "1" ]; then
echo "== bench N=1 same workload"
timeout 900 python bench.py --steps ${STEPS:-5} --warmup 3 --workload 16384x4096 --no-cpu-baseline > gpurun_out/bench_big_n1.json 2> gpurun_out/bench_big_n1.err
echo rc=$?; tail -3 gpurun_out/bench_big_n1.err; cat gpurun_out/bench_big_n1.json | cut -c1-1500
fi
