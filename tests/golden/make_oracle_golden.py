"""Field snapshots produced by the oracle (oracle/fv_oracle.py), committed as regression fixtures.

The reference holds no golden vectors for the assembled solve (parity unpinned; FiPy is not installable), so
these snapshots do NOT pin the oracle to the reference -- they pin the oracle (and, in the -m gpu tests, the CUDA path)
to itself across changes.  Native Faghri mesh (500 x 9 = 4500 cells), direct banded solves.

    python tests/golden/make_oracle_golden.py        # rewrites tests/golden/oracle_golden.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import fv_oracle as fo  # noqa: E402

CASES = {
    # BASELINE configs[0]: constant properties, one sweep per step, hasOld=False (main.py:41-46,200-202)
    'simple_1sweep': dict(modes=dict(properties='simple'), dt=0.1, steps=40, sweeps=1, has_old=False),
    # BASELINE configs[1]: T-dependent literal mode, updateOld + 5 sweeps (main.py:267-271), quick-parity dt = 1 s
    'tdep_5sweeps': dict(modes=dict(), dt=1.0, steps=12, sweeps=5, has_old=True),
    # SURVEY 8f switches together: lazy Gamma, q source, harmonic face k, radiative condenser
    'intended_physics': dict(modes=dict(gamma='lazy', source='q', face_k='harmonic', radiation=True), dt=0.5, steps=10,
                             sweeps=3, has_old=True),
}


def run_case(spec):
    cfg = fo.load_config()
    grid = fo.build_grid(dict(cfg['mesh']), cfg['dimensions'])
    o = fo.Oracle(grid, fo.case_from_config(**spec['modes']), has_old=spec['has_old'])
    res = []
    for _ in range(spec['steps']):
        if spec['has_old']:
            o.update_old()
        for _ in range(spec['sweeps']):
            res.append(o.sweep(spec['dt'], solver='banded'))
    return o.T.copy(), np.array(res)


if __name__ == '__main__':
    out = {}
    for name, spec in CASES.items():
        T, res = run_case(spec)
        out[name + '_T'] = T
        out[name + '_res'] = res
        print(name, T.shape, 'T range', T.min(), T.max(), 'last residual', res[-1])
    np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'oracle_golden.npz'), **out)
