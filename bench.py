#!/usr/bin/env python
"""bench.py -- cell-timesteps/s (all sweeps and PCG iterations included) of the transient conduction hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NZxNR]

One "step" = one implicit-Euler time step of the reference's sweep loop (main.py:267-271: updateOld + 5 sweeps)
on a synthetic refined composite mesh with temperature-dependent properties (BASELINE.json configs[2] at N=1,
configs[3] -- axial slabs -- at N>1).  Prints ONE JSON line (contract in the task statement):
value = cells * K / device seconds (CUDA events, max over ranks); e2e = the same through the public API with
pinned HOST buffers (H2D of the field before and D2H after every step inside the timed region);
roofline = dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM copy peak;
cpu_baseline = the oracle (numpy/scipy restatement of the FiPy sweep) timed on this box's host cores.

--impl reference times that oracle alone (FiPy itself is not installable: DESIGN.md) and prints the same line
with "impl": "reference".
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout must carry exactly ONE JSON line, but libraries chat on it (NCCL prints its version banner there at
# NCCL_DEBUG=VERSION/WARN): keep a private copy of the real stdout for the result and point fd 1 at stderr meanwhile
_REAL_STDOUT = None


def guard_stdout():
    """Called once by main(): fd 1 -> stderr for everything but the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        _REAL_STDOUT = os.fdopen(os.dup(1), 'w')
        os.dup2(2, 1)


def emit_line(obj):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(obj) + '\n')
    out.flush()


RADIAL_SPLIT = {1024: (128, 256, 640), 4096: (512, 1024, 2560), 512: (64, 128, 320), 256: (32, 64, 160), 128: (16, 32, 80)}
DT, SWEEPS, HAS_OLD = 0.1, 5, True
# algorithmic bytes per cell and launch (fp64; DESIGN.md "Kernels"): reads + writes of full-size fields
BYTES_PER_CELL = {'cg_A': 56, 'cg_B': 72, 'cg_init': 32, 'diag': 48, 'factor': 24, 'coef': 24,
                  'mg_res': 62, 'mg_line': 48}      # fine-level V-cycle kernels (down 64 incl. the fused level-1 pre-smoothing / up 60
                                                    # B per cell; final line solve 48)


def parse_workload(s):
    nz, nr = (int(v) for v in s.lower().split('x'))
    if nr not in RADIAL_SPLIT:
        raise SystemExit(f"unsupported radial size {nr}; choose from {sorted(RADIAL_SPLIT)}")
    return nz, nr


def mesh_params_for(nz, nr):
    a, b, c = RADIAL_SPLIT[nr]
    return dict(nx_wall=nz, nr_wall=c, nx_wick=nz, nr_wick=b, nx_vc=nz, nr_vc=a)


def workload_config(kind, nz, nr, world, members=0, sample=None):
    """`config` of the JSON line: built from the workload alone, so the B200 arm and the reference arm print the SAME dict
    when they run the same workload (run-specific facts -- iterations, loop mode, solver options -- go under `run`)."""
    if kind == 'ensemble':
        w = (f'ensemble of {members} independent Faghri pipes (native {nz}x{nr} mesh each), T-dependent properties, literal mode, '
             f'dt={DT}s, updateOld + {SWEEPS} sweeps/step, member m: q~U(0.5,1.5)*26770 W/m2, h~U(2,20), T_amb~U(280,300) K, rng(1234)')
    else:
        w = (f'synthetic composite mesh {nz}x{nr} (axial x radial), T-dependent properties, literal mode, dt={DT}s, '
             f'updateOld + {SWEEPS} sweeps/step, linear solves to 1e-9 K')
    if sample:
        w += f'; {sample}'
    return {'workload': w,
            'l2': 'working set (>= 11 fp64 fields) exceeds the 126 MB L2; no flush needed' if kind != 'ensemble' or members * nz * nr * 88 > 126e6
                  else 'working set fits the L2 (latency-bound workload)',
            'time_steps': 'steps W+1..W+K of the transient from the ambient field'}


class ClockSampler:
    Q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
        'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.proc, self.index = None, index
        self.path = f'/tmp/hp_clocks_{os.getpid()}.csv'

    def start(self):
        exe = shutil.which('nvidia-smi')
        if not exe:
            return
        self.fh = open(self.path, 'w')
        self.proc = subprocess.Popen([exe, '-i', str(self.index), f'--query-gpu={self.Q}', '--format=csv,noheader,nounits',
                                      '-lms', '100'], stdout=self.fh, stderr=subprocess.DEVNULL)

    def stop(self):
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.fh.close()
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        os.unlink(self.path)
        if not sm:
            return None
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'samples': len(sm), 'reasons': sorted(reasons)}


def measured_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        try:
            return float(json.load(open(p))['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
        except Exception:
            pass
    return 6650.0, 'fallback (B200_PROFILING.md 6.65 TB/s)'


# ------------------------------------------------------------------------------------------------------------
# oracle timing (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------------------------
def oracle_steps(nz, nr, n_steps, warmup=0, precond='mgline', solver=None):
    """Time n_steps implicit-Euler steps (updateOld + 5 sweeps) of the oracle: by default the same PCG + preconditioner +
    tolerance as the GPU arm; solver='splu' = SuperLU direct solve per sweep (closest to FiPy's scipy LinearLUSolver)."""
    from oracle import fv_oracle as fo
    cfg = fo.load_config()
    grid = fo.build_grid(mesh_params_for(nz, nr), cfg['dimensions'])
    o = fo.Oracle(grid, fo.case_from_config(), has_old=HAS_OLD)
    if solver is None:
        solver = 'banded' if nr <= 64 else ('pcg_mgline' if precond == 'mgline' else 'pcg_rline')
    its = []

    def one():
        o.update_old()
        for _ in range(SWEEPS):
            if solver.startswith('pcg'):
                o.sweep(DT, solver=solver, tol=1e-9, criterion='precond_inf')
                its.append(o.last_solver_info.get('iterations', 0))
            else:
                o.sweep(DT, solver=solver)
    for _ in range(warmup):
        one()
    its.clear()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        one()
    dt = time.perf_counter() - t0
    return grid.n_cells * n_steps / dt, dt, (sum(its) / max(n_steps, 1) if its else None)


def oracle_ensemble_steps(members, n_steps, warmup=0):
    """The oracle on `members` independent native-mesh pipes (banded direct solves), one after the other."""
    import numpy as np
    from oracle import fv_oracle as fo
    cfg = fo.load_config()
    grid = fo.build_grid(dict(cfg['mesh']), cfg['dimensions'])
    rng = np.random.default_rng(1234)
    q = cfg['parameters']['Q_input_flux'] * rng.uniform(0.5, 1.5, 1024)
    hh = rng.uniform(2.0, 20.0, 1024)
    Ta = rng.uniform(280.0, 300.0, 1024)
    orcs = []
    for m in range(members):
        case = fo.case_from_config(h=float(hh[m]))
        case.parameters['Q_input_flux'] = float(q[m])
        case.parameters['T_amb'] = float(Ta[m])
        orcs.append(fo.Oracle(grid, case, has_old=HAS_OLD))

    def one():
        for o in orcs:
            o.update_old()
            for _ in range(SWEEPS):
                o.sweep(DT, solver='banded')
    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        one()
    dt = time.perf_counter() - t0
    return members * grid.n_cells * n_steps / dt, dt


REF_SAMPLE = (4096, 1024)      # largest mesh of the workload's recipe the oracle steps through in ~10 s per step


def run_reference(args, rank, world):
    """The reference's CPU path on this box's host cores.  FiPy is not installable, so this is the oracle port.
    N = 1: the SAME workload as the B200 arm (4096x1024, same steps / warm-up).  N > 1: the 16384x4096 workload does not fit
    the time box on a CPU (>= 5 min per step), so each step is a 4096x1024 sample of it -- named in config.workload."""
    if rank != 0:
        return
    ensemble = (args.workload or '').startswith('ensemble')
    if ensemble:
        M = int(args.workload.split(':')[1]) if ':' in args.workload else 1024
        sm = min(M, 8)
        value, secs = oracle_ensemble_steps(sm, args.steps, args.warmup)
        from oracle import fv_oracle as fo
        mcfg = fo.load_config()['mesh']
        nz, nr = int(mcfg['nx_wall']), int(mcfg['nr_vc'] + mcfg['nr_wick'] + mcfg['nr_wall'])
        config = workload_config('ensemble', nz, nr, world, members=M,
                                 sample=None if sm == M else f'reference arm: {sm} of the {M} members per step (members are independent: per-cell rate)')
        sample = f'{args.steps} steps (after {args.warmup} warm-up) of the oracle (banded direct solves) on {sm} members, one after the other'
        its, splu = None, None
    else:
        nz, nr = parse_workload(args.workload or ('4096x1024' if world == 1 else '16384x4096'))
        snz, snr = (nz, nr) if nz * nr <= REF_SAMPLE[0] * REF_SAMPLE[1] else REF_SAMPLE
        value, secs, its = oracle_steps(snz, snr, args.steps, args.warmup, args.precond)
        config = workload_config('mesh', nz, nr, world,
                                 sample=None if (snz, snr) == (nz, nr) else
                                 f'reference arm: each step is a {snz}x{snr} sample of it (same recipe and physics, 1/{nz * nr // (snz * snr)} of the cells)')
        sample = (f'{args.steps} steps (after {args.warmup} warm-up) of the numpy/scipy oracle (FiPy restatement) on the {snz}x{snr} synthetic '
                  f'composite mesh, PCG/{args.precond} to 1e-9 K like the B200 arm; FiPy itself is not installable')
        # SuperLU direct solve per sweep = FiPy's scipy LinearLUSolver: one step on a 1024x256 mesh
        # of the same recipe (a factorisation at 4096x1024 takes minutes)
        splu = None
        if not args.no_splu:
            v2, s2, _ = oracle_steps(1024, 256, 1, 0, args.precond, solver='splu')
            splu = {'value': v2, 'unit': 'cell-timesteps/s', 'cores': 1, 'kind': 'port',
                    'sample': f'1 step (5 SuperLU factorisations + solves) on a 1024x256 mesh of the same recipe, {s2:.1f} s'}
    line = {
        'impl': 'reference', 'metric': 'cell-timesteps/s (incl. sweeps)', 'value': value, 'unit': 'cell-timesteps/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * secs / max(args.steps, 1),
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': config,
        'run': {'pcg_iterations_per_step': its, 'host_cores': os.cpu_count(), 'threads_used': 1},
        'cpu_baseline': {'value': value, 'unit': 'cell-timesteps/s', 'cores': 1, 'kind': 'port', 'sample': sample},
        'cpu_baseline_splu': splu,
        'e2e': {'value': value, 'unit': 'cell-timesteps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    emit_line(line)


# ------------------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------------------
def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the conduction path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    from heat_pipe_solver_b200 import generate_composite_mesh, params
    from heat_pipe_solver_b200.distributed import make_solver_for_rank
    cfg = params.load_config()
    ensemble = (args.workload or '').startswith('ensemble')
    stream = torch.cuda.Stream()
    if ensemble:
        # BASELINE configs[4]: M independent Faghri-geometry pipes (native 500x9 mesh), members sharded over the ranks,
        # no data-path collective; member m: q ~ U(0.5,1.5)*26770, h ~ U(2,20), T_amb ~ U(280,300), rng(1234)
        import numpy as np
        from heat_pipe_solver_b200.ensemble import EnsembleSolver, shard_members
        M = int(args.workload.split(':')[1]) if ':' in args.workload else 1024
        rng = np.random.default_rng(1234)
        q = cfg['parameters']['Q_input_flux'] * rng.uniform(0.5, 1.5, M)
        hh = rng.uniform(2.0, 20.0, M)
        Ta = rng.uniform(280.0, 300.0, M)
        lo, hi = shard_members(M, rank, world)
        nz, nr = int(cfg['mesh']['nx_wall']), int(cfg['mesh']['nr_vc'] + cfg['mesh']['nr_wick'] + cfg['mesh']['nr_wall'])
        loop_mode = args.loop_mode
        with torch.cuda.stream(stream):
            ens = EnsembleSolver(dict(cfg['mesh']), cfg['dimensions'], cfg['parameters'], cfg['constants'], q=q[lo:hi], h=hh[lo:hi],
                                 T_amb=Ta[lo:hi], tol=args.tol, loop_mode=loop_mode, precond=args.precond,
                                 refactor_every=args.refactor_every, mg_levels=args.mg_levels,
                                 mg_coarse_sweeps=args.mg_sweeps, mg_omega=args.mg_omega)
        solver = ens.solver

        class _M:                      # what the rest of the function needs from a mesh
            numberOfCells = M * nz * nr
        mesh = _M()
    else:
        nz, nr = parse_workload(args.workload or ('4096x1024' if world == 1 else '16384x4096'))
        mesh, _ = generate_composite_mesh(mesh_params_for(nz, nr), cfg['dimensions'])
    with torch.cuda.stream(stream):
        loop_mode = 'host' if (world > 1 and args.comm == 'nccl' and not ensemble) else args.loop_mode
        if not ensemble:
            solver = make_solver_for_rank(mesh, cfg, rank, world, comm=args.comm, tol=args.tol, loop_mode=loop_mode,
                                          extrapolate=args.extrapolate, precond=args.precond,
                                          mg_levels=args.mg_levels, mg_coarse_sweeps=args.mg_sweeps, mg_omega=args.mg_omega,
                                          refactor_every=args.refactor_every, reuse_converged=args.reuse_converged, while_unroll=args.while_unroll, tma_rows=bool(args.tma_rows))

        def barrier():
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        # ---- device-resident throughput ------------------------------------------------------------
        T_init = solver.value_tensor().clone()        # the transient restarts from here for the end-to-end region
        solver.run(DT, args.warmup, sweeps=SWEEPS, has_old=HAS_OLD)
        barrier()
        it0, l0 = solver.total_iters, solver.kernel_launches
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        solver.run(DT, args.steps, sweeps=SWEEPS, has_old=HAS_OLD)
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        clk = clocks.stop() if rank == 0 else None
        iters, launches = solver.total_iters - it0, solver.kernel_launches - l0
        stats = solver.stats(min(SWEEPS * args.steps, 64))
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        n_cells = mesh.numberOfCells
        value = n_cells * args.steps / (ms * 1e-3)

        # ---- end to end through the public API with host buffers ---------------------------------------
        n_own = solver.n_cells
        a = torch.empty(n_own, dtype=torch.float64).pin_memory()
        b = torch.empty(n_own, dtype=torch.float64).pin_memory()
        # same time steps as the device-resident region: the iteration count per step falls as the transient smooths
        # (and the predicted PCG start gains history), so continuing from where that region stopped would flatter e2e
        solver.set_value(T_init)
        solver.run(DT, args.warmup, sweeps=SWEEPS, has_old=HAS_OLD)
        a.copy_(solver.value_tensor().reshape(-1).cpu())
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            solver.run_host(a, b, DT, 1, sweeps=SWEEPS, has_old=HAS_OLD)     # H2D + step + D2H + sync
            a, b = b, a
        e1.record(stream)
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))
        if world > 1:
            t = torch.tensor([e2e_ms], dtype=torch.float64, device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms = float(t.item())
        e2e_value = n_cells * args.steps / (e2e_ms * 1e-3)

        slab_field = None
        if world > 1 and not ensemble and not args.no_single_gpu_ref:
            # field after the end-to-end region = W + K steps from the ambient field, the state the one-GPU run ends in
            from heat_pipe_solver_b200.distributed import gather_field
            full = gather_field(solver.value_tensor())
            slab_field = full if rank == 0 else None
            del full
        # ---- per-kernel CUDA-event times (stream-launch mode, same workload continuing) ----------------------
        roof = None
        if (world == 1 and not (ensemble and args.precond == 'rline')) or (args.profile_slabs and not ensemble and args.comm == 'p2p'):
            # (ensembles with the line preconditioner run each pipe's step inside one CTA, k_pipe_step: on chip, no per-kernel
            #  HBM roofline to speak of -- BASELINE.md section 4)
            # (slabs: every rank runs the same stream-mode pass -- the flag protocol needs all of them; rank 0's table is
            #  printed, its per-kernel times include the waits for the neighbours)
            solver.profile_begin(16384)
            solver.run(DT, max(1, min(3, args.steps)), sweeps=SWEEPS, has_old=HAS_OLD)
            kt = solver.profile_end()
            peak, peak_src = measured_peak()
            tot_ms = sum(v[1] for v in kt.values())
            dom = max((k for k in kt if k in BYTES_PER_CELL), key=lambda k: kt[k][1])
            # per-launch duration over the launches that did the kernel's work ("full": >= half the longest launch of the
            # kind) -- launches that return at once in an already converged sweep do not move the algorithmic bytes
            cnt, tms, fcnt, fms = kt[dom]
            n_loc = solver.n_cells                 # cells this rank's kernels stream (the whole mesh on one GPU)
            bytes_per_launch = BYTES_PER_CELL.get(dom, 0) * n_loc
            achieved = bytes_per_launch / (fms / fcnt * 1e-3) / 1e9 if fcnt else 0.0
            traffic = None
            tp = os.path.join(ROOT, 'profiles', 'ncu_traffic.json')
            if os.path.exists(tp):
                try:
                    traffic = json.load(open(tp)).get(dom)
                except Exception:
                    traffic = None
            names = {'mg_res': 'k_mg_fused<MODE 2> + k_mg_res (the two fine-level residual passes of the V-cycle)',
                     'mg_line': 'k_mg_line (final line solve of the V-cycle)', 'cg_init': 'k_cg_B<INIT>'}
            roof = {'bound': 'hbm', 'kernel': names.get(dom, 'k_' + dom), 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                    'frac': achieved / peak, 'traffic': traffic, 'peak_source': peak_src,
                    'algorithmic_bytes_per_launch': bytes_per_launch, 'avg_launch_ms': fms / fcnt if fcnt else None,
                    'launches_timed': fcnt, 'launches_returning_at_once': cnt - fcnt,
                    'share_of_step': tms / tot_ms if tot_ms else None,
                    'all_kernels': {k: {'launches': c, 'total_ms': m, 'full_launches': fc, 'avg_ms': (fm / fc if fc else None),
                                        'GBps': (BYTES_PER_CELL.get(k, 0) * n_loc / (fm / fc * 1e-3) / 1e9 if fc else None)}
                                    for k, (c, m, fc, fm) in kt.items() if c}}
        solver.close()
        # strong-scaling reference: the same 16384x4096 workload on ONE GPU (rank 0, after the slab handles are gone), over
        # the SAME time steps (the cost of a step falls along the transient, so the step indices must match)
        def single_gpu_run(big_mesh):
            one = make_solver_for_rank(big_mesh, cfg, 0, 1, tol=args.tol, loop_mode=args.loop_mode, precond=args.precond,
                                       extrapolate=args.extrapolate, refactor_every=args.refactor_every,
                                       mg_levels=args.mg_levels, mg_coarse_sweeps=args.mg_sweeps, mg_omega=args.mg_omega,
                                       reuse_converged=args.reuse_converged, while_unroll=args.while_unroll, tma_rows=bool(args.tma_rows))
            try:
                one.run(DT, args.warmup, sweeps=SWEEPS, has_old=HAS_OLD)
                stream.synchronize()
                s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                it1 = one.total_iters
                s0.record(stream)
                one.run(DT, args.steps, sweeps=SWEEPS, has_old=HAS_OLD)
                s1.record(stream)
                stream.synchronize()
                diff = None
                if slab_field is not None and tuple(slab_field.shape) == (one.nz, one.nr):
                    diff = float((one.value_tensor() - slab_field).abs().max().item())
                return s0.elapsed_time(s1) / args.steps, (one.total_iters - it1) / args.steps, diff
            finally:
                one.close()

        strong = None
        if world > 1 and not ensemble:
            barrier()
            if rank == 0 and not args.no_single_gpu_ref:
                try:
                    ms1, its1, fdiff = single_gpu_run(mesh)
                    strong = {'single_gpu_ms_per_step': ms1, 'single_gpu_value': n_cells / (ms1 * 1e-3),
                              'max_abs_T_slabs_minus_single_gpu_K': fdiff,
                              'single_gpu_pcg_iterations_per_step': its1,
                              'speedup_vs_single_gpu': ms1 / (ms / args.steps),
                              'note': f'same {nz}x{nr} workload on one B200 over the same time steps ({args.warmup} warm-up + '
                                      f'{args.steps} timed), measured in this run'}
                except Exception as exc:      # the other ranks wait at the barrier below: never leave them there
                    strong = {'error': repr(exc)[:300]}
            barrier()
        # N = 1 with --scaling-ref: the line of the 4096x1024 workload also carries the one-GPU number of the multi-GPU
        # workload (the N > 1 lines measure that denominator themselves: `strong_scaling`)
        scaling_ref = None
        if world == 1 and not ensemble and args.workload is None and args.scaling_ref:
            try:
                bnz, bnr = 16384, 4096
                big, _ = generate_composite_mesh(mesh_params_for(bnz, bnr), cfg['dimensions'])
                ms1, its1, _ = single_gpu_run(big)
                scaling_ref = {'workload': f'synthetic composite mesh {bnz}x{bnr}, same physics / sweeps / solver, one B200',
                               'ms_per_step': ms1, 'value': bnz * bnr / (ms1 * 1e-3), 'unit': 'cell-timesteps/s',
                               'pcg_iterations_per_step': its1, 'steps': args.steps, 'warmup': args.warmup}
            except Exception as exc:          # never let the extra measurement cost the headline line
                scaling_ref = {'error': repr(exc)[:300]}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        if ensemble:
            sm, n_cpu = 8, 4
            v, secs = oracle_ensemble_steps(sm, n_cpu, 0)
            cpu = {'value': v, 'unit': 'cell-timesteps/s', 'cores': 1, 'kind': 'port',
                   'sample': f'{n_cpu} steps of the numpy/scipy oracle (banded direct solves) on {sm} of the members, one after the other '
                             f'({secs:.1f} s; members are independent, so the per-cell rate carries over); host has {os.cpu_count()} cores, the port uses 1'}
        else:
            snz, snr = 2048, 512
            n_cpu = 5
            v, secs, its_cpu = oracle_steps(snz, snr, n_cpu, 0, args.precond)
            cpu = {'value': v, 'unit': 'cell-timesteps/s', 'cores': 1, 'kind': 'port',
                   'sample': f'{n_cpu} steps of the numpy/scipy oracle (FiPy restatement, PCG/{args.precond} tol 1e-9 K, {its_cpu:.1f} it/step) on a {snz}x{snr} synthetic '
                             f'composite mesh, same physics and sweeps ({secs:.1f} s); host has {os.cpu_count()} cores, the port uses 1; the full-size run is `--impl reference`'}
    if rank == 0:
        line = {
            'metric': 'cell-timesteps/s (incl. sweeps)', 'value': value, 'unit': 'cell-timesteps/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak' if (world == 1 or ensemble) else 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config('ensemble' if ensemble else 'mesh', nz, nr, world,
                                      members=(int(args.workload.split(':')[1]) if ensemble and ':' in args.workload else 1024)),
            'run': {'parallelism': 'single GPU' if world == 1 else (f'members sharded over {world} ranks, no collective' if ensemble else f'axial slabs x{world}, ' + ('NVLink peer-memory halo rows + CG scalars fused into the kernels, one CUDA graph per step' if args.comm == 'p2p' else 'NCCL send/recv halo + all-gathered CG scalars, host-driven loop')),
                    'solver': f'PCG ({args.precond}) tol {args.tol:g} K', 'pcg_iterations_per_step': iters / args.steps, 'loop_mode': loop_mode,
                    'refactor_every_sweeps': args.refactor_every, 'tma_rows': bool(args.tma_rows), 'reuse_converged_sweeps': bool(args.reuse_converged),
                    'pcg_start': f"extrapolated from the step history, max order {args.extrapolate}",
                    'e2e_note': 'hp_run_host uploads the field every step but keeps the solver-side step history (T_old, increments) for the '
                                'predicted PCG start; a caller uploading an unrelated field pays first-step iteration counts'},
            'e2e': {'value': e2e_value, 'unit': 'cell-timesteps/s', 'h2d_bytes_per_step': 8 * n_cells,
                    'd2h_bytes_per_step': 8 * n_cells, 'ms_per_step': e2e_ms / args.steps},
            'gpu_launches': int(launches), 'clocks': clk, 'strong_scaling': strong if world > 1 else None,
            'multi_gpu_workload_on_one_gpu': scaling_ref,
            'roofline': roof, 'cpu_baseline': cpu,
            'last_sweep_stats': stats[-SWEEPS:],
        }
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default=None, help='NZxNR, e.g. 4096x1024 (default at N=1) or 16384x4096 (default at N>1), or ensemble[:M]')
    ap.add_argument('--tol', type=float, default=1e-9)
    ap.add_argument('--loop-mode', default='graph', choices=['graph', 'host'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-splu', action='store_true', help='--impl reference: skip the SuperLU variant (cpu_baseline_splu)')
    ap.add_argument('--profile-slabs', action='store_true', help='N>1: per-kernel CUDA-event table of rank 0 (stream-launch pass on every rank)')
    ap.add_argument('--while-unroll', type=int, default=1)
    ap.add_argument('--tma-rows', type=int, default=0, help='1: k_cg_B with the cp.async.bulk / mbarrier row pipeline (hp_solve_opts.tma_rows)')
    ap.add_argument('--extrapolate', type=int, default=3, help='maximum order of the extrapolated PCG start (0 = off)')
    ap.add_argument('--scaling-ref', action='store_true',
                    help='N=1: also time the 16384x4096 multi-GPU workload on this one GPU (extra key multi_gpu_workload_on_one_gpu)')
    ap.add_argument('--no-single-gpu-ref', action='store_true', help='N>1: skip the one-GPU run of the same workload (strong-scaling denominator)')
    ap.add_argument('--precond', default=None, choices=['rline', 'mgline', 'jacobi'],
                    help='PCG preconditioner: mgline (V-cycle over radial-line smoothing, default for meshes), rline (the radial-line '
                         'tridiagonal preconditioner alone), jacobi')
    ap.add_argument('--refactor-every', type=int, default=SWEEPS,
                    help='re-factor the line pivots / coarse operators every k-th sweep of a step (default: once per step; '
                         'they only precondition, the answer is unchanged)')
    ap.add_argument('--mg-levels', type=int, default=6)
    ap.add_argument('--mg-sweeps', type=int, default=2)
    ap.add_argument('--mg-omega', type=float, default=0.6)
    ap.add_argument('--reuse-converged', action='store_true',
                    help='sweeps that exactly repeat a converged sweep only redo the bookkeeping (off: every sweep assembles, like the reference)')
    ap.add_argument('--comm', default='p2p', choices=['p2p', 'nccl'], help='slab exchange: NVLink peer stores fused into the kernels, or NCCL calls')
    args = ap.parse_args()
    if args.precond is None:        # meshes: multigrid over lines; ensembles of small pipes: the line preconditioner (-> k_pipe_step)
        args.precond = 'rline' if (args.workload or '').startswith('ensemble') else 'mgline'
    guard_stdout()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if args.warmup < 3 and args.impl != 'reference':
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, world, local_rank)


if __name__ == '__main__':
    main()
