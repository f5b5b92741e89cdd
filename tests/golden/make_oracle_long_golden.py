"""Full-length transients of BASELINE configs[0] and configs[1], produced by the oracle with direct (banded) solves.

    python tests/golden/make_oracle_long_golden.py [simple|tdep]     # rewrites tests/golden/oracle_long_<case>.npz

configs[0]: config/faghri.json, native 500 x 9 mesh, constant properties (main.py:41-46), dt = 0.1 s, one sweep per
            step with hasOld=False (main.py:34,200-202), to t = 3000 s (30 000 steps).
configs[1]: the same mesh with temperature-dependent k / cp / rho (main.py:103-117), `T.updateOld()` + 5 sweeps per step
            (main.py:267-271), dt = 1 s (the quick-parity variant of SURVEY 8d), to t = 12 000 s (12 000 steps).
Whole fields are kept at a few step counts so that the GPU run can be compared along the transient, not only at its end;
the outer-wall row (main.py:165-169 `top_wall_mask`) is kept every 1000 s for the read-offs of the reference's committed
plot (plots/12000s/top_wall_axial_profiles.png).

Like oracle_golden.npz these snapshots pin the oracle and the CUDA path to each other, not to FiPy.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fv_oracle as fo  # noqa: E402

CASES = {
    'simple': dict(modes=dict(properties='simple'), dt=0.1, steps=30000, sweeps=1, has_old=False,
                   keep=(1000, 10000, 20000, 30000), wall_every=10000),
    'tdep': dict(modes=dict(), dt=1.0, steps=12000, sweeps=5, has_old=True,
                 keep=(100, 1000, 3000, 6000, 12000), wall_every=1000),
}


def run_case(spec, progress=None):
    cfg = fo.load_config()
    grid = fo.build_grid(dict(cfg['mesh']), cfg['dimensions'])
    o = fo.Oracle(grid, fo.case_from_config(**spec['modes']), has_old=spec['has_old'])
    out = {'dt': np.float64(spec['dt']), 'sweeps': np.int64(spec['sweeps']), 'has_old': np.int64(spec['has_old']),
           'keep_steps': np.array(spec['keep'], dtype=np.int64)}
    walls, wall_steps, last_res = [], [], []
    keep = set(spec['keep'])
    for step in range(1, spec['steps'] + 1):
        if spec['has_old']:
            o.update_old()
        for _ in range(spec['sweeps']):
            res = o.sweep(spec['dt'], solver='banded')
        if step in keep:
            out[f'T_{step}'] = o.T.copy()
            last_res.append(res)
        if step % spec['wall_every'] == 0:
            walls.append(o.T[:, -1].copy())
            wall_steps.append(step)
        if progress and step % 1000 == 0:
            progress(step, o.T)
    out['wall'] = np.array(walls)
    out['wall_steps'] = np.array(wall_steps, dtype=np.int64)
    out['residual_at_keep'] = np.array(last_res)
    return out


if __name__ == '__main__':
    which = sys.argv[1:] or list(CASES)
    for name in which:
        t0 = time.time()
        out = run_case(CASES[name], lambda s, T: print(f'{name}: step {s}, wall T(z=0) {T[0, -1]:.3f} K, max {T.max():.3f} K, '
                                                       f'{time.time() - t0:.0f} s', flush=True))
        np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', f'oracle_long_{name}.npz'), **out)
        print(name, 'done in', time.time() - t0, 's')
