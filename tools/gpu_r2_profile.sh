#!/bin/bash
# One GPU: parity suite, ensemble lines, then the ncu launch list and one full capture of a complete PCG iteration
# (stream-launch mode so that every kernel is a separate launch).  Outputs gpurun_out/<tag>_*.
tag=${1:-r2h}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --workload ensemble:1024 --precond rline --steps 20 --warmup 5 > gpurun_out/${tag}_ensemble_rline.json 2> gpurun_out/${tag}_ens1.err; echo "ens rline rc=$?"
timeout 600 python bench.py --workload ensemble:1024 --precond mgline --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_ensemble_mgline.json 2> gpurun_out/${tag}_ens2.err; echo "ens mgline rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host"
timeout 600 $CMD > gpurun_out/${tag}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cg_A|k_cg_B|k_mg_res|k_mg_line|k_mg_fused|k_diag|k_factor' -s 1400 -c 24 -f -o gpurun_out/${tag}_prof $CMD > gpurun_out/${tag}_ncu2.log 2>&1; echo "ncu full rc=$?"
python - <<PY
import json
for f in ("${tag}_ensemble_rline.json","${tag}_ensemble_mgline.json","${tag}_bench.json"):
    try:
        d=json.load(open("gpurun_out/"+f))
        print(f, "value %.3e e2e %.3e ms/step %.3f it/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["run"]["pcg_iterations_per_step"]))
        for k,v in d["roofline"]["all_kernels"].items(): print("   ", k, v)
    except Exception as e: print(f, "ERR", e)
PY
