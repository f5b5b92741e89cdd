#!/bin/bash
# 8-GPU box: slab parity tests (2 and 4 ranks), strong-scaling lines at 8/4/2, ensemble at 8.
tag=${1:-r2g}
timeout 900 python -m pytest tests/test_gpu_slabs.py -x -q -m gpu -k "distributed_vcycle or main_run or (p2p and graph) or (nccl and mgline)" > gpurun_out/${tag}_slabs.log 2>&1; echo "slab tests rc=$?"; tail -n 3 gpurun_out/${tag}_slabs.log
run() { name=$1; n=$2; shift 2
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@" > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; echo "$name rc=$?"
}
run scale_n8 8 --steps 20 --warmup 5 --profile-slabs
run scale_n4 4 --steps 20 --warmup 5
run scale_n2 2 --steps 20 --warmup 5
run ens_n8 8 --steps 20 --warmup 5 --workload ensemble:1024 --precond rline
python - <<PY
import json
for name in ("scale_n8","scale_n4","scale_n2","ens_n8"):
    try:
        d=json.load(open("gpurun_out/${tag}_%s.json" % name))
        print(name, "ms/step %.2f value %.3e e2e %.3e it/step %.2f" % (d["ms_per_step"], d["value"], d["e2e"]["value"], d["run"]["pcg_iterations_per_step"]), d.get("strong_scaling"))
        r=d.get("roofline")
        if r:
            for k,v in r["all_kernels"].items(): print("   ", k, v)
    except Exception as e: print(name, "ERR", e)
PY
