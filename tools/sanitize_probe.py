"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck  python tools/sanitize_probe.py
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py pipe
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def main(which):
    from heat_pipe_solver_b200 import ConductionSolver, generate_composite_mesh, params
    from oracle import fv_oracle as fo
    cfg = params.load_config()
    cases = {
        'pipe': (None, dict(precond='rline'), 3, 3),                                   # k_pipe_step / k_pipe_finalize
        'rows': ((96, 12, 24, 64), dict(precond='mgline', extrapolate=3), 3, 3),        # row kernels, fused V-cycle, k_factor_team
        'tma': ((24, 120, 250, 630), dict(precond='rline', tma_rows=True), 2, 2),       # k_cg_B_tma
        'cluster': ((64, 512, 1024, 2560), dict(precond='mgline'), 1, 2),               # cluster-team stages (4 CTAs per row)
        'subwarp': ((40, 2, 4, 10), dict(precond='mgline', pipe_kernel=False), 2, 2),   # 16-lane teams
    }
    for name, (shape, kw, steps, sweeps) in cases.items():
        if which and name not in which:
            continue
        mp = dict(cfg['mesh']) if shape is None else fo.synthetic_mesh_params(*shape)
        mesh, _ = generate_composite_mesh(dict(mp), cfg['dimensions'])
        with ConductionSolver(mesh, cfg['dimensions'], cfg['parameters'], cfg['constants'], **kw) as s:
            s.run(0.1, steps, sweeps=sweeps, has_old=True)
            T = s.value
            st = s.stats(steps * sweeps)
        assert np.isfinite(T).all() and all(x['converged'] for x in st), (name, st)
        print(name, 'ok', mesh.shape, [x['iters'] for x in st], flush=True)


if __name__ == '__main__':
    main(sys.argv[1:])
