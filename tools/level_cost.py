"""In-situ cost of the V-cycle per level: pins the PCG at exactly max_iter iterations per sweep (tol=0) and times steps
for several hierarchy depths / loop modes.  per-iteration cost = (t[2K] - t[K]) / (SWEEPS*K).  GPU only."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from heat_pipe_solver_b200 import generate_composite_mesh, params
from heat_pipe_solver_b200.distributed import make_solver_for_rank

nz, nr = bench.parse_workload(sys.argv[1] if len(sys.argv) > 1 else '4096x1024')
cfg = params.load_config()
mesh, _ = generate_composite_mesh(bench.mesh_params_for(nz, nr), cfg['dimensions'])
stream = torch.cuda.Stream()
K = 6
out = []
with torch.cuda.stream(stream):
    for loop_mode in ('graph', 'host'):
        for precond, levels in (('rline', 0), ('mgline', 2), ('mgline', 3), ('mgline', 4), ('mgline', 5), ('mgline', 6)):
            t = {}
            for mi in (K, 2 * K):
                s = make_solver_for_rank(mesh, cfg, 0, 1, tol=1e-300, max_iter=mi, loop_mode=loop_mode, precond=precond,
                                         mg_levels=max(levels, 2), refactor_every=5)
                s.run(bench.DT, 2, sweeps=bench.SWEEPS, has_old=bench.HAS_OLD)
                stream.synchronize()
                i0 = s.total_iters
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                s.run(bench.DT, 5, sweeps=bench.SWEEPS, has_old=bench.HAS_OLD)
                e1.record(stream)
                stream.synchronize()
                t[mi] = (e0.elapsed_time(e1) / 5, (s.total_iters - i0) / 5)
                s.close()
            per_it = (t[2 * K][0] - t[K][0]) / max(t[2 * K][1] - t[K][1], 1) * 1e3
            fixed = t[K][0] - per_it * 1e-3 * t[K][1]
            rec = dict(loop_mode=loop_mode, precond=precond, levels=levels, us_per_iter=per_it, fixed_ms_per_step=fixed,
                       its=(t[K][1], t[2 * K][1]), ms=(t[K][0], t[2 * K][0]))
            print(json.dumps(rec), flush=True)
            out.append(rec)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/level_cost.json', 'w'), indent=1)
