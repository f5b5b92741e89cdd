#!/bin/bash
# pytest -m gpu, bench, ncu launch list + one full capture of the two PCG kernels. Each stage under a timeout.
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?; tail -3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; cut -c1-400 gpurun_out/bench_ref.json
if [ "${NCU:-1}" = "1" ]; then
echo "== ncu launch list (stream-launch mode so that every kernel is a separate launch)"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host > gpurun_out/ncu1.log 2>&1
echo rc=$?
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cg_ -s 600 -c 4 -f -o gpurun_out/prof_cg \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host > gpurun_out/ncu2.log 2>&1
echo rc=$?
fi
