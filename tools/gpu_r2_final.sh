#!/bin/bash
# Final one-GPU evidence of the round: parity suite, smoke, both bench arms, protocol variants, ensemble, ncu launch list.
tag=${1:-r2z}
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "pytest rc=$?"; tail -n 3 gpurun_out/${tag}_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${tag}_bench_100steps.json 2> gpurun_out/${tag}_b100.err; echo "bench100 rc=$?"
timeout 600 python bench.py --precond rline --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_bench_rline.json 2> gpurun_out/${tag}_brl.err; echo "bench rline rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --workload ensemble:1024 --precond rline --steps 20 --warmup 5 > gpurun_out/${tag}_ensemble_n1.json 2> gpurun_out/${tag}_ens.err; echo "ens rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --loop-mode host"
timeout 600 $CMD > gpurun_out/${tag}_plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu1.log 2>&1; echo "ncu list rc=$?"
python - <<PY
import json
for f in ("bench","bench_100steps","bench_rline","bench_ref","ensemble_n1"):
    try:
        d=json.load(open("gpurun_out/${tag}_%s.json" % f))
        print(f, "value %.3e e2e %.3e ms/step %.3f it/step %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["run"]["pcg_iterations_per_step"]))
        r=d.get("roofline")
        if r: print("   roofline", r["kernel"], "frac %.3f share %.3f" % (r["frac"], r["share_of_step"]))
    except Exception as e: print(f, "ERR", e)
PY
