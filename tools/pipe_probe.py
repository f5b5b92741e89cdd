import sys
sys.path.insert(0, '/root/repo')
from heat_pipe_solver_b200.main import run
kw = dict(sweeps=5, has_old=True) if len(sys.argv) > 1 and sys.argv[1] == '5' else {}
r = run(t_end=6.0, dt=0.1, capture_properties=False, **kw)
print(r['iterations'], r['elapsed'])
