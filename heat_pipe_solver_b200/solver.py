"""Host-side driver object of the conduction hot path.

``ConductionSolver`` plays the role of the reference's ``(T, eq)`` pair
(`/root/reference/src/2d/main.py:34,148-150`): ``sweep(dt)`` is ``eq.sweep(var=T, dt=dt)``
(main.py:202), ``update_old()`` is ``T.updateOld()`` (main.py:269), ``value`` is ``T.value``.
All arithmetic happens in libhpsolver.so (CUDA, sm_100a) through the C ABI of
``include/hp_solver.h``; torch is only used for device buffers and the stream.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .regions import Regions

PRECOND = {'jacobi': 0, 'rline': 1, 'mgline': 2}
CRITERION = {'precond_inf': 0, 'rowscaled_inf': 1, 'l2_rel_rhs': 2}
LOOP_MODE = {'host': 0, 'graph': 1}
PROPERTIES = {'temperature_dependent': 0, 'simple': 1}
GAMMA = {'snapshot': 0, 'lazy': 1}
SOURCE = {'q_over_k': 0, 'q': 1}
FACE_K = {'masked_at_Tface': 0, 'harmonic': 1}

# hp::PropId (csrc/hp_props.cuh)
PROP_IDS = {n: i for i, n in enumerate([
    'na_rho_s', 'na_rho_l', 'na_rho_v', 'na_h_vap', 'na_k_s', 'na_k_l', 'na_cp_s', 'na_cp_l', 'na_cp_v',
    'na_cv_v', 'na_h_l', 'na_h_v', 'na_mu_v', 'na_psat', 'steel_rho_eff', 'steel_rho', 'steel_k', 'steel_cp',
    'wick_k', 'wick_cp', 'wick_rho', 'vc_k_evap_cond', 'vc_k_adiabatic', 'na_k_i', 'na_cp_i', 'na_rho_i',
    'na_dpdt'])}


def _choice(table, key, what):
    try:
        return table[key]
    except KeyError:
        raise ValueError(f"{what} must be one of {sorted(table)}, got {key!r}")


def make_phys(parameters, dimensions, constants, *, properties='temperature_dependent', gamma='snapshot',
              source='q_over_k', face_k='masked_at_Tface', radiation=False,
              simple_rho=7200.0, simple_cp=440.5, simple_k=55.0):
    ph = _lib.HpPhys()
    ph.R_vc = dimensions['R_vc']
    ph.M_g = parameters['M_g_sodium']
    ph.R_gas = constants['R']
    ph.porosity = parameters['porosity_wick']
    ph.T_melt = parameters['T_melting_sodium']
    ph.dT_melt = parameters['delta_T_sodium']
    ph.sigma = constants.get('sigma', 5.670374419e-8)
    ph.simple_rho, ph.simple_cp, ph.simple_k = simple_rho, simple_cp, simple_k      # main.py:41-46
    ph.properties = _choice(PROPERTIES, properties, 'properties')
    ph.gamma = _choice(GAMMA, gamma, 'gamma')
    ph.source = _choice(SOURCE, source, 'source')
    ph.face_k = _choice(FACE_K, face_k, 'face_k')
    ph.radiation = 1 if radiation else 0
    return ph


class SlabLayout:
    """Ghost-inclusive 1-D arrays of the rows [j0, j1) of a global mesh (what hp_mesh_desc wants)."""

    def __init__(self, mesh, regions, j0, j1, line_values, axial_break=None):
        nzg = mesh.nz
        if not (0 <= j0 < j1 <= nzg):
            raise ValueError(f"bad slab rows [{j0},{j1}) of {nzg}")
        self.j0, self.j1, self.nz, self.nr = j0, j1, j1 - j0, mesh.nr
        zv = mesh.z_v
        lo = zv[j0 - 1] if j0 > 0 else zv[j0] - (zv[j0 + 1] - zv[j0])
        hi = zv[j1 + 1] if j1 < nzg else zv[j1] + (zv[j1] - zv[j1 - 1])
        self.z_v = np.ascontiguousarray(np.concatenate(([lo], zv[j0:j1 + 1], [hi])), dtype=np.float64)
        rows = np.clip(np.arange(j0 - 1, j1 + 1), 0, nzg - 1)              # ghost rows clamp at physical ends
        self.zc_flags = np.ascontiguousarray(regions.zc_flags[rows], dtype=np.uint8)
        # face between ghost-inclusive row g and g+1 is the global vertex j0+g
        verts = np.arange(j0, j1 + 2)
        self.zv_flags = np.ascontiguousarray(regions.zv_flags[np.clip(verts, 0, nzg)], dtype=np.uint8)
        is_open = (verts > 0) & (verts < nzg)
        if axial_break is not None:
            ab = np.asarray(axial_break, dtype=bool)                       # [nzg-1]: break between row j and j+1
            if ab.shape != (nzg - 1,):
                raise ValueError("axial_break must have nz-1 entries")
            inner = is_open.copy()
            inner[is_open] &= ~ab[verts[is_open] - 1]
            is_open = inner
        self.axial_open = np.ascontiguousarray(is_open, dtype=np.uint8)
        self.lines = {k: np.ascontiguousarray(np.asarray(v, dtype=np.float64)[rows]) for k, v in line_values.items()}


def _as_line(value, nz, name):
    arr = np.asarray(value, dtype=np.float64)
    if arr.ndim == 0:
        return np.full(nz, float(arr))
    if arr.shape != (nz,):
        raise ValueError(f"{name} must be a scalar or have one entry per axial row ({nz}), got shape {arr.shape}")
    return arr.copy()


class ConductionSolver:
    def __init__(self, mesh, dimensions, parameters, constants, *, h=5.0, q=None, T_amb=None, emissivity=None,
                 properties='temperature_dependent', gamma='snapshot', source='q_over_k',
                 face_k='masked_at_Tface', radiation=False, axial_break=None, rows=None,
                 precond='auto', criterion='precond_inf', tol=1e-9, max_iter=5000, loop_mode='graph',
                 poll_chunk=8, refactor_every=1, extrapolate=True, mg_levels=6, mg_coarse_sweeps=2, mg_omega=0.6,
                 mg_zscale=0.5, reuse_converged=False, while_unroll=1, peer_timeout_ms=60000, tma_rows=False, pipe_kernel=True, ext_guard_tols=300.0,
                 device=None, simple=(7200.0, 440.5, 55.0),
                 layout=None, regions=None):
        import torch
        if not torch.cuda.is_available():
            raise _lib.HpError("ConductionSolver needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
        self._torch = torch
        self.lib = _lib.load()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device) \
            if not isinstance(device, torch.device) else device
        self.mesh, self.dimensions, self.parameters, self.constants = mesh, dimensions, parameters, constants
        self.regions = Regions(mesh, dimensions) if regions is None else regions
        nzg = mesh.nz
        j0, j1 = (0, nzg) if rows is None else rows
        if layout is None:
            lines = dict(
                q=_as_line(parameters['Q_input_flux'] if q is None else q, nzg, 'q'),               # main.py:130
                h=_as_line(h, nzg, 'h'),                                                            # main.py:37
                Tamb=_as_line(parameters['T_amb'] if T_amb is None else T_amb, nzg, 'T_amb'),       # main.py:36
                eps=_as_line(parameters.get('emissivity', 0.0) if emissivity is None else emissivity, nzg, 'emissivity'))
            layout = SlabLayout(mesh, self.regions, j0, j1, lines, axial_break)
        self.layout = lay = layout
        self.nr, self.nz = lay.nr, lay.nz
        self.phys = make_phys(parameters, dimensions, constants, properties=properties, gamma=gamma, source=source,
                              face_k=face_k, radiation=radiation, simple_rho=simple[0], simple_cp=simple[1],
                              simple_k=simple[2])
        r_v = np.ascontiguousarray(mesh.r_v, dtype=np.float64)
        keep = [r_v, lay.z_v, np.ascontiguousarray(self.regions.cell_band, dtype=np.int8),
                np.ascontiguousarray(self.regions.rface_band, dtype=np.int8),
                np.ascontiguousarray(self.regions.zface_band, dtype=np.int8),
                lay.zc_flags, lay.zv_flags, lay.axial_open, lay.lines['q'], lay.lines['h'], lay.lines['Tamb'], lay.lines['eps']]
        md = _lib.HpMeshDesc()
        md.nr, md.nz, md.cartesian = self.nr, self.nz, 0 if mesh.cylindrical else 1
        md.r_v = keep[0].ctypes.data_as(_lib.c_double_p)
        md.z_v = keep[1].ctypes.data_as(_lib.c_double_p)
        md.cell_band = keep[2].ctypes.data_as(_lib.c_int8_p)
        md.rface_band = keep[3].ctypes.data_as(_lib.c_int8_p)
        md.zface_band = keep[4].ctypes.data_as(_lib.c_int8_p)
        md.zc_flags = keep[5].ctypes.data_as(_lib.c_uint8_p)
        md.zv_flags = keep[6].ctypes.data_as(_lib.c_uint8_p)
        md.axial_open = keep[7].ctypes.data_as(_lib.c_uint8_p)
        md.q_line = keep[8].ctypes.data_as(_lib.c_double_p)
        md.h_line = keep[9].ctypes.data_as(_lib.c_double_p)
        md.Tamb_line = keep[10].ctypes.data_as(_lib.c_double_p)
        md.eps_line = keep[11].ctypes.data_as(_lib.c_double_p)
        self._keep = keep
        self._h = C.c_void_p()
        with torch.cuda.device(self.device):
            self._stream = torch.cuda.current_stream(self.device)
            _lib.check(self.lib.hp_create(C.byref(self._h), C.byref(md), C.byref(self.phys),
                                          C.c_void_p(self._stream.cuda_stream)), 'hp_create')
        if precond == 'auto':
            # multigrid pays off once the mesh is refined axially (mesh-independent iteration counts); on small meshes
            # the solve is launch-latency bound and the plain radial-line preconditioner has fewer launches per iteration
            precond = 'mgline' if (self.nz >= 256 and self.nz * self.nr >= 262144) else 'rline'
        self.set_options(precond=precond, criterion=criterion, tol=tol, max_iter=max_iter, loop_mode=loop_mode,
                         poll_chunk=poll_chunk, refactor_every=refactor_every, extrapolate=extrapolate,
                         mg_levels=mg_levels, mg_coarse_sweeps=mg_coarse_sweeps, mg_omega=mg_omega, mg_zscale=mg_zscale,
                         reuse_converged=reuse_converged, while_unroll=while_unroll, peer_timeout_ms=peer_timeout_ms,
                         tma_rows=tma_rows, pipe_kernel=pipe_kernel, ext_guard_tols=ext_guard_tols)

    # ------------------------------------------------------------------ options
    def set_options(self, **kw):
        cur = getattr(self, 'options', dict(precond='rline', criterion='precond_inf', tol=1e-9, max_iter=5000,
                                            loop_mode='graph', poll_chunk=8, refactor_every=1, extrapolate=True,
                                            mg_levels=6, mg_coarse_sweeps=2, mg_omega=0.6, mg_zscale=0.5,
                                            reuse_converged=False, while_unroll=1, peer_timeout_ms=60000, tma_rows=False, pipe_kernel=True, ext_guard_tols=300.0))
        cur = dict(cur)
        for k, v in kw.items():
            if k not in cur:
                raise TypeError(f"unknown solver option {k!r}")
            cur[k] = v
        o = _lib.HpSolveOpts()
        o.precond = _choice(PRECOND, cur['precond'], 'precond')
        o.criterion = _choice(CRITERION, cur['criterion'], 'criterion')
        o.tol, o.max_iter = float(cur['tol']), int(cur['max_iter'])
        o.loop_mode = _choice(LOOP_MODE, cur['loop_mode'], 'loop_mode')
        o.poll_chunk, o.refactor_every = int(cur['poll_chunk']), int(cur['refactor_every'])
        e = cur['extrapolate']                      # True = highest order (cubic), False = off, 1..3 = maximum order
        o.extrapolate = (3 if e else 0) if isinstance(e, bool) else int(e)
        o.mg_levels, o.mg_coarse_sweeps, o.mg_omega = int(cur['mg_levels']), int(cur['mg_coarse_sweeps']), float(cur['mg_omega'])
        o.reuse_converged = 1 if cur['reuse_converged'] else 0
        o.mg_zscale = float(cur['mg_zscale'])
        o.while_unroll, o.peer_timeout_ms = int(cur['while_unroll']), int(cur['peer_timeout_ms'])
        o.tma_rows = 1 if cur['tma_rows'] else 0
        o.pipe_kernel = 1 if cur['pipe_kernel'] else 0
        o.ext_guard_tols = float(cur['ext_guard_tols'])
        _lib.check(self.lib.hp_set_opts(self._h, C.byref(o)), 'hp_set_opts')
        self.options = cur

    @property
    def effective_precond(self):
        """The preconditioner in force (the library replaces 'mgline' by 'rline' on a mesh too short to coarsen)."""
        o = _lib.HpSolveOpts()
        _lib.check(self.lib.hp_get_opts(self._h, C.byref(o)), 'hp_get_opts')
        return {v: k for k, v in PRECOND.items()}[o.precond]

    # ------------------------------------------------------------------ field access
    @property
    def n_cells(self):
        return self.nr * self.nz

    # The library enqueues on the stream the handle was created with (self._stream); torch temporaries of the helpers
    # below are produced / consumed on whatever stream is current when they are called.  _on_handle_stream orders the two.
    def _enter(self):
        """The handle's stream waits for the caller's current stream (inputs prepared there are complete)."""
        cur = self._torch.cuda.current_stream(self.device)
        if cur != self._stream:
            self._stream.wait_stream(cur)
        return cur

    def _leave(self, cur, *tensors):
        """The caller's current stream waits for the handle's stream; the tensors stay alive until both are done."""
        if cur != self._stream:
            cur.wait_stream(self._stream)
            for t in tensors:
                t.record_stream(self._stream)

    def set_value(self, T):
        """`T.setValue(...)`: scalar, numpy array or torch tensor of nz*nr values (FiPy order)."""
        torch = self._torch
        if np.isscalar(T):
            buf = torch.full((self.n_cells,), float(T), dtype=torch.float64, device=self.device)
        elif isinstance(T, torch.Tensor):
            buf = T.to(device=self.device, dtype=torch.float64).reshape(-1).contiguous()
        else:
            buf = torch.from_numpy(np.ascontiguousarray(T, dtype=np.float64).reshape(-1)).to(self.device)
        if buf.numel() != self.n_cells:
            raise ValueError(f"expected {self.n_cells} values, got {buf.numel()}")
        cur = self._enter()
        _lib.check(self.lib.hp_set_T_dev(self._h, C.c_void_p(buf.data_ptr())), 'hp_set_T_dev')
        self._leave(cur, buf)
        self._stream.synchronize()            # buf may go out of scope: the copy has consumed it

    def value_tensor(self):
        """Copy of the field as a torch.float64 CUDA tensor of shape (nz, nr)."""
        out = self._torch.empty((self.nz, self.nr), dtype=self._torch.float64, device=self.device)
        cur = self._enter()
        _lib.check(self.lib.hp_get_T_dev(self._h, C.c_void_p(out.data_ptr())), 'hp_get_T_dev')
        self._leave(cur, out)
        return out

    @property
    def value(self):
        """`T.value`: numpy array in FiPy cell order (main.py:208)."""
        return self.value_tensor().cpu().numpy().ravel()

    # ------------------------------------------------------------------ FiPy-like stepping
    def update_old(self):
        _lib.check(self.lib.hp_update_old(self._h), 'hp_update_old')

    def sweep(self, dt, has_old=False, want_residual=True):
        """`res = eq.sweep(var=T, dt=dt)`; returns FiPy's pre-solve L2 residual (None if not wanted: no sync)."""
        if want_residual:
            res = C.c_double()
            _lib.check(self.lib.hp_sweep(self._h, float(dt), int(has_old), C.byref(res)), 'hp_sweep')
            return res.value
        _lib.check(self.lib.hp_sweep(self._h, float(dt), int(has_old), None), 'hp_sweep')
        return None

    def step(self, dt, sweeps=1, has_old=False):
        _lib.check(self.lib.hp_step(self._h, float(dt), int(sweeps), int(has_old)), 'hp_step')

    def run(self, dt, n_steps, sweeps=1, has_old=False):
        _lib.check(self.lib.hp_run(self._h, float(dt), int(n_steps), int(sweeps), int(has_old)), 'hp_run')

    def run_host(self, T_in_pinned, T_out_pinned, dt, n_steps, sweeps=1, has_old=False):
        """End-to-end call with HOST buffers (pinned torch CPU tensors): H2D + steps + D2H + sync."""
        for t in (T_in_pinned, T_out_pinned):
            if t.device.type != 'cpu' or t.dtype != self._torch.float64 or t.numel() != self.n_cells or not t.is_contiguous():
                raise ValueError("host buffers must be contiguous float64 CPU tensors of nz*nr elements")
        _lib.check(self.lib.hp_run_host(self._h, C.c_void_p(T_in_pinned.data_ptr()), C.c_void_p(T_out_pinned.data_ptr()),
                                        float(dt), int(n_steps), int(sweeps), int(has_old)), 'hp_run_host')

    def synchronize(self):
        self._stream.synchronize()

    # ------------------------------------------------------------------ statistics / hooks
    def stats(self, n=1):
        arr = (_lib.HpSweepStat * n)()
        got = self.lib.hp_get_stats(self._h, arr, n)
        if got < 0:
            _lib.check(got, 'hp_get_stats')
        return [dict(residual_l2=a.residual_l2, rhs_l2=a.rhs_l2, crit=a.crit, iters=a.iters,
                     converged=bool(a.flags & 1), hit_max_iter=bool(a.flags & 2), breakdown=bool(a.flags & 4),
                     peer_timeout=bool(a.flags & 8))
                for a in arr[:got]]

    @property
    def total_iters(self):
        return int(self.lib.hp_total_iters(self._h))

    @property
    def total_sweeps(self):
        return int(self.lib.hp_total_sweeps(self._h))

    @property
    def kernel_launches(self):
        return int(self.lib.hp_kernel_launches(self._h))

    @property
    def peer_fault(self):
        """True once a kernel of this handle gave up waiting for a neighbouring slab; every call then raises HpError."""
        return bool(self.lib.hp_comm_fault(self._h))

    def clear_peer_fault(self):
        """Collective recovery (all ranks: synchronise, barrier, call this, barrier)."""
        _lib.check(self.lib.hp_comm_clear_fault(self._h), 'hp_comm_clear_fault')

    def init_comm(self, rank, world, unique_id: bytes):
        """Join the library's NCCL communicator (axial slabs); `unique_id` = 128 bytes from rank 0."""
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _lib.check(self.lib.hp_comm_init(self._h, int(rank), int(world), buf), 'hp_comm_init')
        self.rank, self.world = rank, world

    def peer_export(self, nz_of_rank) -> bytes:
        """256 bytes of CUDA-IPC handles (mailbox, (q,z) block, T, global coarse levels) for `init_peer` on the other ranks."""
        buf = (C.c_uint8 * 256)()
        nzs = (C.c_int32 * len(nz_of_rank))(*[int(v) for v in nz_of_rank])
        _lib.check(self.lib.hp_peer_export(self._h, nzs, buf), 'hp_peer_export')
        return bytes(buf)

    def init_peer(self, all_handles: bytes, nz_of_rank):
        """Switch the slab exchange from NCCL calls to peer-memory stores fused into the kernels."""
        n = len(nz_of_rank)
        if len(all_handles) != 256 * n:
            raise ValueError("need 256 bytes of handles per rank")
        buf = (C.c_uint8 * len(all_handles)).from_buffer_copy(all_handles)
        nzs = (C.c_int32 * n)(*[int(v) for v in nz_of_rank])
        _lib.check(self.lib.hp_peer_import(self._h, buf, nzs), 'hp_peer_import')

    def profile_begin(self, max_launches=8192):
        """Start per-kernel CUDA-event timing; sweeps run in stream-launch mode until profile_end()."""
        _lib.check(self.lib.hp_profile_begin(self._h, int(max_launches)), 'hp_profile_begin')

    def profile_end(self):
        """-> {kind: (launches, total_ms, full_launches, full_ms)}; "full" = launches that did the work of the kind
        (duration >= half the longest one), i.e. without those that return at once because the sweep is converged."""
        kt = _lib.HpKernelTimes()
        _lib.check(self.lib.hp_profile_end(self._h, C.byref(kt)), 'hp_profile_end')
        return {name: (int(kt.count[i]), float(kt.total_ms[i]), int(kt.full_count[i]), float(kt.full_ms[i]))
                for i, name in enumerate(_lib.KERNEL_KINDS)}

    def assemble_debug(self, dt, has_old=False):
        """The assembled 5-point system at the current iterate, as (nz, nr) CUDA tensors."""
        torch = self._torch
        out = {k: torch.empty((self.nz, self.nr), dtype=torch.float64, device=self.device) for k in ('cR', 'cZ', 'diag', 'rhs')}
        self._enter()                        # hp_assemble_debug synchronises the handle's stream itself
        _lib.check(self.lib.hp_assemble_debug(self._h, float(dt), int(has_old), C.c_void_p(out['cR'].data_ptr()),
                                              C.c_void_p(out['cZ'].data_ptr()), C.c_void_p(out['diag'].data_ptr()),
                                              C.c_void_p(out['rhs'].data_ptr())), 'hp_assemble_debug')
        return out

    def cell_fields(self):
        """rho, cp and face-averaged k per cell at the current iterate (main.py:212-254)."""
        torch = self._torch
        out = {k: torch.empty((self.nz, self.nr), dtype=torch.float64, device=self.device) for k in ('rho', 'cp', 'k')}
        cur = self._enter()
        _lib.check(self.lib.hp_eval_cell_fields(self._h, C.c_void_p(out['rho'].data_ptr()), C.c_void_p(out['cp'].data_ptr()),
                                                C.c_void_p(out['k'].data_ptr())), 'hp_eval_cell_fields')
        self._leave(cur, *out.values())
        return out

    def eval_property(self, name, T):
        """Elementwise material law on the device; T: torch CUDA float64 tensor."""
        torch = self._torch
        Tc = T.to(device=self.device, dtype=torch.float64).contiguous()
        out = torch.empty_like(Tc)
        cur = self._enter()
        _lib.check(self.lib.hp_eval_property(self._h, _choice(PROP_IDS, name, 'property'), C.c_void_p(Tc.data_ptr()),
                                             C.c_void_p(out.data_ptr()), Tc.numel()), 'hp_eval_property')
        self._leave(cur, Tc, out)
        return out

    # ------------------------------------------------------------------ life cycle
    def close(self):
        if getattr(self, '_h', None) is not None and self._h:
            self.lib.hp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
