#!/bin/bash
# One-GPU round: GPU parity suite, smoke, default bench line, ensemble line.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_tests.log
tail -n 5 gpurun_out/${tag}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_bench_100steps.json 2> gpurun_out/${tag}_bench_100.err; echo "bench100 rc=$?"
timeout 600 python bench.py --workload ensemble:1024 --precond rline --steps 20 --warmup 5 > gpurun_out/${tag}_ensemble_n1.json 2> gpurun_out/${tag}_ens.err; echo "ens rc=$?"
python - <<PY
import json
for f in ("${tag}_bench.json","${tag}_bench_100steps.json","${tag}_ensemble_n1.json"):
    try:
        d=json.load(open("gpurun_out/"+f))
        print(f, "value %.3e e2e %.3e ms/step %.3f it/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["run"]["pcg_iterations_per_step"]))
        r=d.get("roofline")
        if r:
            print("  roofline", r["kernel"], "frac %.3f share %.3f" % (r["frac"], r["share_of_step"]))
            for k,v in r["all_kernels"].items(): print("   ", k, v)
    except Exception as e: print(f, "ERR", e)
PY
