#!/bin/bash
echo "== tests touching small nr"; timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x 2>&1 | tail -4
python bench.py --workload ensemble:1024 --steps 10 --warmup 3 --precond rline 2> gpurun_out/ens.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('ensemble1024 rline: ms/step %.3f value %.4e its %.1f' % (d['ms_per_step'], d['value'], d['config']['pcg_iterations_per_step']))" || tail -5 gpurun_out/ens.err
python - <<'PY'
import time
from heat_pipe_solver_b200.main import run
for kw in (dict(sweeps=1, has_old=False), dict(sweeps=5, has_old=True)):
    out=run(t_end=101.0, dt=0.1, capture_properties=False, **kw)
    print('native 101 s run', kw, 'elapsed(loop) %.2f' % out['elapsed'], 'iters', out['iterations'], 'peak T %.3f' % out['profiles'][-1].max(), 'cell-steps/s %.3e' % (4500*1010/out['elapsed']))
PY
