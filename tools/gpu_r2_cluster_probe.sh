#!/bin/bash
# per-rank shape of the 8-GPU run on ONE GPU (no peers): per-kernel table + ncu capture of the cluster-team stages
tag=${1:-r2e}
timeout 300 python bench.py --workload 2048x4096 --steps 10 --warmup 5 --no-cpu-baseline > gpurun_out/${tag}_2048x4096.json 2> gpurun_out/${tag}_2048x4096.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_2048x4096.json"))
print("ms/step %.3f it/step %.2f" % (d["ms_per_step"], d["run"]["pcg_iterations_per_step"]))
for k,v in d["roofline"]["all_kernels"].items(): print("   ", k, v)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_mg_fused_cl4|k_mg_line|k_mg_res' -s 40 -c 16 -f -o gpurun_out/${tag}_cluster \
   python bench.py --workload 2048x4096 --steps 1 --warmup 3 --no-cpu-baseline --loop-mode host > gpurun_out/${tag}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/${tag}_cluster.ncu-rep
