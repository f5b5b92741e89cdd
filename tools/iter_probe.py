"""PCG iterations per sweep along a transient, for a list of solver option sets (tuning / diagnosis tool).

    python tools/iter_probe.py NZxNR steps "extrapolate=0" "extrapolate=3,mg_omega=0.6" ...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    import torch
    from heat_pipe_solver_b200 import ConductionSolver, generate_composite_mesh, params
    nz, nr = bench.parse_workload(sys.argv[1])
    steps = int(sys.argv[2])
    cfg = params.load_config()
    mesh, _ = generate_composite_mesh(bench.mesh_params_for(nz, nr), cfg['dimensions'])
    for spec in sys.argv[3:]:
        kw = {}
        for item in spec.split(','):
            if not item:
                continue
            k, v = item.split('=')
            kw[k] = v if k in ('precond', 'loop_mode') else (float(v) if ('.' in v or 'e' in v or k == 'ext_guard_tols') else int(v))
        opts = dict(precond='mgline', refactor_every=5, extrapolate=3)
        opts.update(kw)
        with ConductionSolver(mesh, cfg['dimensions'], cfg['parameters'], cfg['constants'], **opts) as s:
            rows = []
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.run(0.1, 5, sweeps=5, has_old=True)
            ev0.record(s._stream)
            for k in range(5, steps):
                s.run(0.1, 1, sweeps=5, has_old=True)
                if k % 5 == 4 or k == steps - 1:
                    rows.append((k + 1, [st['iters'] for st in s.stats(5)]))
            ev1.record(s._stream)
            s.synchronize()
            tot = s.total_iters
        print(f"{spec or 'default'}: total iters {tot} over {steps} steps, {ev0.elapsed_time(ev1) / (steps - 5):.3f} ms/step (incl. stats reads); "
              + ' '.join(f"step{k}:{it}" for k, it in rows), flush=True)


if __name__ == '__main__':
    main()
