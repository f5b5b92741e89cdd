#!/bin/bash
# First-contact GPU script: every stage under its own timeout so a hang cannot eat the box.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== host-mode sweep" ; timeout 120 python - <<'PY' 2>&1 | tail -20
import numpy as np, torch, sys
sys.path.insert(0,'.')
from heat_pipe_solver_b200 import ConductionSolver, generate_composite_mesh, params
from oracle import fv_oracle as fo
cfg=params.load_config()
for mode in ('host','graph'):
    mesh,_=generate_composite_mesh(dict(cfg['mesh']),cfg['dimensions'])
    s=ConductionSolver(mesh,cfg['dimensions'],cfg['parameters'],cfg['constants'],loop_mode=mode)
    o=fo.Oracle(fo.build_grid(cfg['mesh'],cfg['dimensions']),fo.case_from_config())
    for k in range(3):
        r=s.sweep(0.1); ro=o.sweep(0.1,solver='banded')
        print(mode,k,'res',r,ro,'stats',s.stats(1)[0]['iters'],'err',np.abs(s.value.reshape(o.T.shape)-o.T).max(), flush=True)
    s.close()
PY
echo "== smoke"; timeout 120 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -40
