#!/bin/bash
# pytest -m gpu, smoke, bench (both arms), ncu launch list + one full capture of the PCG / V-cycle kernels.
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
fi
if [ "${BENCH:-1}" = "1" ]; then
echo "== bench"; timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo rc=$?; tail -3 gpurun_out/bench.err; cut -c1-700 gpurun_out/bench.json
echo "== bench rline"; timeout 900 python bench.py --steps ${STEPS:-20} --warmup 3 --precond rline --no-cpu-baseline > gpurun_out/bench_rline.json 2> gpurun_out/bench_rline.err; echo rc=$?; cut -c1-300 gpurun_out/bench_rline.json
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo rc=$?; cut -c1-300 gpurun_out/bench_ref.json
fi
if [ "${NCU:-list}" = "list" ]; then
echo "== ncu launch list (stream-launch mode so that every kernel is a separate launch)"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-gpu-ref --loop-mode host > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-gpu-ref --loop-mode host > gpurun_out/ncu1.log 2>&1
echo rc=$?
fi
# one profiler pass per gpurun call: NCU=full (alone, with TESTS=0) after the plain run of the same command passed
if [ "${NCU:-list}" = "full" ]; then
echo "== ncu full"
timeout 600 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-gpu-ref --loop-mode host > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cg_A|k_cg_B|k_mg_res|k_mg_line|k_mg_fused' -s 500 -c 40 -f -o gpurun_out/prof_cg \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-single-gpu-ref --loop-mode host > gpurun_out/ncu2.log 2>&1
echo rc=$?
fi
