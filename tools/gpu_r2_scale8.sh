#!/bin/bash
# 8-GPU box: per-kernel table of the slab run, variants, ensemble.  Outputs gpurun_out/<tag>_*.
tag=${1:-r2d}
run() { # name, nproc, args...
  name=$1; n=$2; shift 2
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; echo "$name rc=$?"
}
run n8_prof 8 --steps 20 --warmup 5 --profile-slabs
run n8_cs1 8 --steps 20 --warmup 5 --mg-sweeps 1 --no-single-gpu-ref
run n8_unroll2 8 --steps 20 --warmup 5 --while-unroll 2 --no-single-gpu-ref
run ens_n8 8 --steps 20 --warmup 5 --workload ensemble:1024 --precond rline
python - <<PY
import json
for name in ("n8_prof","n8_cs1","n8_unroll2","n8_rline","ens_n8"):
    try:
        d=json.load(open("gpurun_out/${tag}_%s.json" % name))
        print(name, "ms/step %.2f value %.3e e2e %.3e it/step %.2f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["run"]["pcg_iterations_per_step"]), (d.get("strong_scaling") or {}).get("speedup_vs_single_gpu"))
        r=d.get("roofline")
        if r:
            for k,v in r["all_kernels"].items(): print("   ", k, v)
    except Exception as e: print(name, "ERR", e)
PY
