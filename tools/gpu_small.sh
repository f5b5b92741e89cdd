#!/bin/bash
for pc in mgline rline; do
python bench.py --workload ensemble:1024 --steps 10 --warmup 3 --precond $pc 2> gpurun_out/ens_$pc.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ensemble1024 $pc: ms/step %.3f value %.4e e2e %.4e its %.1f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['config']['pcg_iterations_per_step']))" || tail -5 gpurun_out/ens_$pc.err
done
python - <<'PY'
import time, numpy as np
from heat_pipe_solver_b200.main import run
for pc in ('mgline','rline'):
    for kw in (dict(sweeps=1, has_old=False), dict(sweeps=5, has_old=True)):
        t=time.time(); out=run(t_end=101.0, dt=0.1, precond=pc, capture_properties=False, **kw); dt=time.time()-t
        print('native 101 s run', pc, kw, 'wall %.2f s' % dt, 'elapsed(loop) %.2f' % out['elapsed'], 'iters', out['iterations'], 'peak T %.3f' % out['profiles'][-1].max(), 'cell-steps/s %.3e' % (4500*1010/out['elapsed']))
PY
