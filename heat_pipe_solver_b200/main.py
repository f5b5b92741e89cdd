"""Time loop of the reference driver (`/root/reference/src/2d/main.py:14-339`) as a function.

The reference is a module-level script that ends in blocking plot calls; here the same sequence --
parameters, property tables, composite mesh, region masks, equation, implicit-Euler loop with
optional ``T.updateOld()`` + sweeps, profile capture every ``measure_every`` seconds of loop time --
is ``run(...)`` and returns the temperature field and the captured profiles.  ``python -m
heat_pipe_solver_b200.main`` prints the reference's console lines (main.py:32,161,197,207,337-339)
without plotting.

All arithmetic of the loop body runs on the GPU (libhpsolver.so); steps between two capture times are
enqueued back to back without host synchronisation.
"""
from __future__ import annotations

import time

import numpy as np

from . import params as params_mod
from .mesh import generate_composite_mesh
from .solver import ConductionSolver
from .regions import BAND_VC, BAND_WICK, BAND_WALL


def measurement_schedule(t_end, dt, measure_times):
    """Loop times, and for each captured profile (step index, loop time) -- the bookkeeping of main.py:200-259."""
    times = np.arange(0, t_end, dt)                                        # main.py:200
    captures = []
    nxt = 0
    last = None
    for step, t in enumerate(times):
        if nxt < len(measure_times) and t >= measure_times[nxt]:            # main.py:204
            if last is None or last < t:                                    # main.py:206
                captures.append((step, float(t)))
                last = float(t)
            nxt += 1
            while nxt < len(measure_times) and measure_times[nxt] <= t:     # main.py:258-259
                nxt += 1
    return times, captures


def run(config_file=None, *, t_end=101.0, dt=0.1, sweeps=1, has_old=False, h=5.0,
        properties='temperature_dependent', gamma='snapshot', source='q_over_k', face_k='masked_at_Tface',
        radiation=False, measure_every=5, measure_times=None, mesh_params=None, T0=None,
        precond='auto', criterion='precond_inf', tol=1e-9, max_iter=5000, loop_mode='graph',
        capture_properties=True, verbose=False, device=None, sweep_tol=None, max_sweeps=50,
        checkpoint=None, checkpoint_every=None, resume=None, extrapolate=True, reuse_converged=False, comm='p2p'):
    """Transient conduction run.  Defaults = the reference's active code path (main.py:119-120,200-202:
    t_end=101 s, dt=0.1 s, one sweep per step with hasOld=False, h=5 W/m2K, temperature-dependent laws).
    The commented "sweep" variant (main.py:267-271) is ``sweeps=5, has_old=True``.

    Extras beyond the reference: ``sweeps='auto'`` (or ``sweep_tol``) re-sweeps every step until FiPy's
    pre-solve residual drops below ``sweep_tol`` (default 1e-8) or ``max_sweeps``; ``checkpoint=path`` writes an ``.npz``
    (field, step index, loop time and EVERY capture list so far) every ``checkpoint_every`` steps counted from the start of
    this call and at the end; ``resume=path`` continues such a run (T_old is rebuilt by ``updateOld`` at the next step, so the
    result is identical to solver tolerance; bit-identical with ``extrapolate=False``).

    Under ``torch.distributed`` with world size N > 1 (``python -m torch.distributed.run --nproc-per-node N -m
    heat_pipe_solver_b200.main``: the reference's ``mpirun -np N python main.py``, main.py:11,339) every rank owns an axial
    slab of the mesh (FiPy partitions ``Grid2D`` the same way); captured profiles and the returned field are gathered, so
    every rank returns the same global arrays.  ``comm='p2p'|'nccl'`` picks the slab exchange.

    Returns a dict: ``T`` (nz, nr) final field, ``times`` / ``profiles`` (top-wall axial temperature profiles,
    main.py:208-209), ``residuals`` (FiPy-style pre-solve residual at each capture), property profiles
    (``rho_*``, ``cp_*``, ``k_*`` for top wall / wick mean / vapor core, main.py:212-254), ``n_steps``,
    ``iterations`` (total PCG iterations), ``elapsed``.
    """
    cfg = params_mod.load_config(config_file)
    parameters, dimensions, constants = dict(cfg['parameters']), dict(cfg['dimensions']), dict(cfg['constants'])
    mparams = dict(cfg['mesh']) if mesh_params is None else dict(mesh_params)
    mesh, _cell_types = generate_composite_mesh(mparams, dimensions)       # main.py:29
    if verbose:
        print(f"Number of cells: {mesh.numberOfCells}")                    # main.py:32
    if measure_times is None:
        measure_times = list(range(measure_every, int(t_end), measure_every))   # main.py:158-159
    if verbose:
        print(f"Measuring at: {measure_times} seconds")                    # main.py:161
    times, captures = measurement_schedule(t_end, dt, measure_times)

    rank, world = 0, 1
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            rank, world = dist.get_rank(), dist.get_world_size()
    except ImportError:
        pass
    kw = dict(h=h, properties=properties, gamma=gamma, source=source, face_k=face_k, radiation=radiation, precond=precond,
              criterion=criterion, tol=tol, max_iter=max_iter, extrapolate=extrapolate, reuse_converged=reuse_converged)
    if world > 1:
        from .distributed import gather_field, make_solver_for_rank
        solver = make_solver_for_rank(mesh, cfg, rank, world, comm=comm,
                                      loop_mode='host' if comm == 'nccl' else loop_mode, **kw)

        def whole(t2d):                      # owned rows of every rank -> the global (nz, k) array, on every rank
            return gather_field(t2d.contiguous())
    else:
        solver = ConductionSolver(mesh, dimensions, parameters, constants, loop_mode=loop_mode, device=device, **kw)

        def whole(t2d):
            return t2d
    j0, j1 = solver.layout.j0, solver.layout.j1
    try:
        if T0 is not None:
            solver.set_value(np.asarray(T0, dtype=np.float64).reshape(mesh.shape)[j0:j1] if np.ndim(T0) else T0)
        reg = solver.regions
        top_is_wall = reg.cell_band[-1] == BAND_WALL                       # main.py:165-169
        vc_cols = np.nonzero(reg.cell_band == BAND_VC)[0]
        wick_cols = np.nonzero(reg.cell_band == BAND_WICK)[0]
        out = dict(times=[], profiles=[], residuals=[],
                   rho_top_wall=[], rho_wick=[], rho_vc=[], cp_top_wall=[], cp_wick=[], cp_vc=[],
                   k_top_wall=[], k_wick=[], k_vc=[])

        def capture(t):
            T = whole(solver.value_tensor())
            st = solver.stats(1)[0]
            if verbose:
                print(f"Time {t:.2f}, Residual: {st['residual_l2']}")      # main.py:207
            out['times'].append(t)
            out['residuals'].append(st['residual_l2'])
            out['profiles'].append(T[:, -1].cpu().numpy() if top_is_wall else np.empty(0))
            if capture_properties:
                f = solver.cell_fields()
                for key, name in (('rho', 'rho'), ('cp', 'cp'), ('k', 'k')):
                    a = whole(f[key])
                    out[f'{name}_top_wall'].append(a[:, -1].cpu().numpy() if top_is_wall else np.empty(0))
                    out[f'{name}_vc'].append(a[:, vc_cols].reshape(-1).cpu().numpy())
                    out[f'{name}_wick'].append(a[:, wick_cols].mean(dim=1).cpu().numpy() if wick_cols.size else np.empty(0))

        if verbose:
            print(f"Simulating for {t_end} seconds...")                    # main.py:197
        start = time.time()
        done = 0
        auto = (sweeps == 'auto') or (sweep_tol is not None)
        stol = 1e-8 if sweep_tol is None else float(sweep_tol)
        sweep_counts = []
        if resume is not None:
            ck = np.load(resume, allow_pickle=False)
            if tuple(ck['shape']) != tuple(mesh.shape) or float(ck['dt']) != float(dt):
                raise ValueError("checkpoint does not match this mesh / dt")
            solver.set_value(ck['T'][j0:j1])
            done = int(ck['step'])
            for key in out:
                if 'ck_' + key in ck.files:
                    out[key] = list(ck['ck_' + key])

        def global_field():
            return whole(solver.value_tensor()).cpu().numpy().reshape(mesh.shape)

        def write_checkpoint(step_index):
            Tg = global_field()              # collective under slabs: every rank takes part, rank 0 writes
            if rank == 0:
                np.savez(checkpoint, T=Tg, step=step_index, dt=dt, shape=np.array(mesh.shape),
                         t=float(times[step_index - 1]) + dt if step_index else 0.0,
                         **{'ck_' + k: np.array(v) for k, v in out.items()})

        def advance(n):
            """n time steps; fixed sweep count = back-to-back graph launches, 'auto' = residual-controlled sweeps."""
            if not auto:
                solver.run(dt, n, sweeps=sweeps, has_old=has_old)
                return
            for _ in range(n):
                solver.update_old()
                k = 0
                while True:
                    res = solver.sweep(dt, has_old=True)
                    k += 1
                    if res <= stol or k >= max_sweeps:
                        break
                sweep_counts.append(k)

        ck_stops = set(range(done + checkpoint_every, len(times), checkpoint_every)) if checkpoint and checkpoint_every else set()
        stops = sorted(set([s + 1 for s, _ in captures if s + 1 > done] + [len(times)]) | ck_stops)
        cap_at = {s + 1: t for s, t in captures}
        for stop in stops:
            if stop > done:
                advance(stop - done)
                done = stop
            if stop in cap_at and (not out['times'] or out['times'][-1] < cap_at[stop]):
                capture(cap_at[stop])
            if stop in ck_stops:
                write_checkpoint(stop)
        solver.synchronize()
        if checkpoint:
            write_checkpoint(done)
        elapsed = time.time() - start
        if verbose:
            print(f"Simulated time: {t_end} seconds")                      # main.py:337
            print(f"Actual execution time: {elapsed:.2f} seconds")         # main.py:338
            print("%d cells on processor %d of %d" % (solver.n_cells, rank, world))   # main.py:339
        result = {k: (np.array(v) if k in ('times', 'residuals') else v) for k, v in out.items()}
        result['profiles'] = np.array(out['profiles'])
        result.update(T=global_field(), n_steps=len(times), iterations=solver.total_iters,
                      sweeps=solver.total_sweeps, elapsed=elapsed, mesh=mesh, kernel_launches=solver.kernel_launches,
                      sweeps_per_step=np.array(sweep_counts))
        return result
    finally:
        solver.close()



def run_adaptive(config_file=None, *, t_end=101.0, dt0=0.1, dt_min=1e-3, dt_max=10.0, target_dT=0.5, reject_factor=2.0,
                 grow=1.5, sweeps=5, h=5.0, measure_times=(), mesh_params=None, T0=None, verbose=False, max_steps=1000000,
                 **solver_kw):
    """Implicit Euler with a step-size controller (not in the reference, whose dt is fixed at main.py:37).

    The controller keeps the largest temperature change of a step near ``target_dT`` [K]: a step that moved some cell by
    more than ``reject_factor * target_dT`` is undone (field restored) and retried with half the dt; after an accepted step
    the next dt is ``dt * clip(target_dT / max|dT|, 0.5, grow)``, kept inside [dt_min, dt_max] and shortened to land exactly on
    the next entry of ``measure_times`` and on ``t_end``.  Every step is ``T.updateOld()`` + ``sweeps`` re-sweeps
    (main.py:267-271) on the GPU; only max|dT| (one device reduction) comes back to the host per step.

    Returns ``T`` (nz, nr), ``t`` and ``dt`` of every accepted step, ``max_dT``, ``n_rejected``, ``times`` / ``profiles``
    (top-wall profiles at ``measure_times``), ``iterations``, ``elapsed``, ``mesh``.
    """
    if not (0.0 < dt_min <= dt0 <= dt_max):
        raise ValueError("need 0 < dt_min <= dt0 <= dt_max")
    if not (target_dT > 0.0 and reject_factor > 1.0 and grow > 1.0):
        raise ValueError("need target_dT > 0, reject_factor > 1, grow > 1")
    cfg = params_mod.load_config(config_file)
    parameters, dimensions, constants = dict(cfg['parameters']), dict(cfg['dimensions']), dict(cfg['constants'])
    mparams = dict(cfg['mesh']) if mesh_params is None else dict(mesh_params)
    mesh, _ = generate_composite_mesh(mparams, dimensions)
    stops = sorted(float(m) for m in measure_times if 0.0 < float(m) < t_end) + [float(t_end)]
    solver = ConductionSolver(mesh, dimensions, parameters, constants, h=h, **solver_kw)
    try:
        if T0 is not None:
            solver.set_value(T0)
        top_is_wall = solver.regions.cell_band[-1] == BAND_WALL
        t, dt = 0.0, float(dt0)
        ts, dts, dTs, times, profiles = [], [], [], [], []
        rejected = 0
        start = time.time()
        nxt = 0
        while nxt < len(stops) and len(ts) < max_steps:
            room = stops[nxt] - t
            step = min(dt, room)
            if room - step < 0.5 * dt_min:               # do not leave a sliver before the stop
                step = room
            before = solver.value_tensor()
            solver.step(step, sweeps=sweeps, has_old=True)
            moved = float((solver.value_tensor() - before).abs().max().item())
            if moved > reject_factor * target_dT and step > dt_min * (1 + 1e-12):
                solver.set_value(before)                 # undo: the step is repeated from the same field with half the dt
                dt = max(dt_min, 0.5 * step)
                rejected += 1
                continue
            t = stops[nxt] if step == room else t + step
            ts.append(t); dts.append(step); dTs.append(moved)
            if verbose:
                print(f"t = {t:.4f} s, dt = {step:.4g} s, max|dT| = {moved:.3e} K")
            if step == room:
                if nxt < len(stops) - 1 or stops[nxt] in [float(m) for m in measure_times]:
                    times.append(t)
                    profiles.append(solver.value_tensor()[:, -1].cpu().numpy() if top_is_wall else np.empty(0))
                nxt += 1
            factor = grow if moved <= 0.0 else min(grow, max(0.5, target_dT / moved))
            dt = min(dt_max, max(dt_min, (dt if step < dt else step) * factor))
        solver.synchronize()
        return dict(T=solver.value.reshape(mesh.shape), t=np.array(ts), dt=np.array(dts), max_dT=np.array(dTs),
                    n_rejected=rejected, times=np.array(times), profiles=np.array(profiles), iterations=solver.total_iters,
                    sweeps=solver.total_sweeps, elapsed=time.time() - start, mesh=mesh, n_steps=len(ts))
    finally:
        solver.close()


if __name__ == '__main__':
    import argparse
    ap = argparse.ArgumentParser(description="heat-pipe transient conduction (B200)")
    ap.add_argument('--config', default=None)
    ap.add_argument('--t-end', type=float, default=101.0)
    ap.add_argument('--dt', type=float, default=0.1)
    ap.add_argument('--sweeps', type=int, default=1)
    ap.add_argument('--has-old', action='store_true')
    ap.add_argument('--simple', action='store_true')
    ap.add_argument('--save', default=None, help='write <path>.npz (fields, profiles) and <path>_profiles.csv')
    ap.add_argument('--experiment', action='append', default=[], metavar='TIME_S=CSV',
                    help='measured wall temperatures (x,y columns) to compare the snapshot at TIME_S with (main.py:372-405)')
    ap.add_argument('--mesh', default=None, metavar='NZ,NR_VC,NR_WICK,NR_WALL', help='synthetic refinement of the composite mesh')
    ap.add_argument('--comm', default='p2p', choices=['p2p', 'nccl'])
    a = ap.parse_args()
    import os
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world > 1:                                  # launched by torch.distributed.run: one process per GPU
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        dist.init_process_group('nccl', device_id=torch.device('cuda', int(os.environ.get('LOCAL_RANK', '0'))))
    mp_ = None
    if a.mesh:
        nz_, v_, w_, wl_ = (int(x) for x in a.mesh.split(','))
        mp_ = dict(nx_wall=nz_, nr_wall=wl_, nx_wick=nz_, nr_wick=w_, nx_vc=nz_, nr_vc=v_)
    r = run(a.config, t_end=a.t_end, dt=a.dt, sweeps=a.sweeps, has_old=a.has_old, mesh_params=mp_, comm=a.comm,
            properties='simple' if a.simple else 'temperature_dependent', verbose=True)
    if world > 1:
        dist.barrier()
        if dist.get_rank() != 0:
            dist.destroy_process_group()
            raise SystemExit(0)
    print(f"peak top-wall temperature: {r['profiles'][-1].max() if len(r['profiles']) else float('nan'):.3f} K")
    if a.save:
        from . import report
        print("wrote", report.save_results(r, a.save), report.write_profiles_csv(r, a.save + '_profiles.csv'))
    if a.experiment:
        from . import report
        files = {float(e.split('=', 1)[0]): e.split('=', 1)[1] for e in a.experiment}
        print(report.format_comparison(report.compare_with_experiment(r, files, time_tolerance=a.dt)))
    if world > 1:
        dist.destroy_process_group()
