#!/bin/bash
# pytest -m gpu, bench, ncu launch list + one full capture of the two PCG kernels. Each stage under a timeout.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -25
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host > gpurun_out/ncu1.log 2>&1
echo rc=$?
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_ -s 600 -c 4 -f -o gpurun_out/prof_cg \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host > gpurun_out/ncu2.log 2>&1
echo rc=$?
fi
