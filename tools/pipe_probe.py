"""The first 60 steps of the native-mesh run (what `ncu --metrics gpu__time_duration.sum` times k_pipe_step on).

    python tools/pipe_probe.py        # the reference's active loop: 1 sweep per step, hasOld=False
    python tools/pipe_probe.py 5      # the sweep variant: updateOld + 5 sweeps (main.py:267-271)
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from heat_pipe_solver_b200.main import run  # noqa: E402

kw = dict(sweeps=5, has_old=True) if len(sys.argv) > 1 and sys.argv[1] == '5' else {}
r = run(t_end=6.0, dt=0.1, capture_properties=False, **kw)
print(r['iterations'], r['elapsed'])
