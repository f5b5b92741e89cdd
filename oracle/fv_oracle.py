"""ORACLE (test infrastructure only -- never imported by the product path).

CPU fp64 restatement (numpy + scipy/LAPACK) of the reference's hot path:
`eq.sweep(var=T, dt=dt)` of `/root/reference/src/2d/main.py:148-150,200-202,267-271`
on the composite cylindrical mesh of `/root/reference/src/2d/mesh.py:70-109`, with the
lazy property composites of `main.py:103-117` and the Robin / flux boundary algebra
of `main.py:123-147`.

The arithmetic of that path lives in the third-party package FiPy (un-pinned in
`/root/reference/environment.yml:6`, absent from /root/reference and not installable
here), so what is restated is FiPy's documented finite-volume behaviour (SURVEY.md
section 9): `CylindricalGrid2D` geometry per radian, interior-face-only
`DiffusionTerm`, arithmetic `faceValue`, `FaceVariable(value=<Variable>)` snapshot
semantics, `.divergence` of masked exterior-face fields as cell sources, the pre-solve
L2 residual returned by `sweep`.

PARITY STATUS: **parity unpinned** for the assembled solve -- the reference holds no
test, fixture or golden vector for this path (`tests/test_fipy.py`,
`tests/test_parallel.py` assert nothing) and FiPy cannot run here.  What *is* pinned:
the property laws (against the reference's own `material_properties.py`, see
`oracle/properties.py`), the mesh vertex recipe (against the reference's own
`mesh.py` run with a recording stub), analytic known-answer tests, and the coarse
read-offs of the reference's committed plots.

Arrays are 2-D `(nz, nr)` C-ordered, i.e. flattened cell id = j*nr + i with the
radial index fastest -- FiPy's ordering for `CylindricalGrid2D(dr->x, dz->y)`.
"""
from __future__ import annotations

import copy
import json
import os
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import properties as P

# Optional row-block threading of the heavy array passes (numpy / LAPACK release the GIL on large blocks): the reference
# arm of bench.py uses the host's cores with it.  Results are bit-identical to the serial path -- every block computes
# exactly the entries the serial expression computes.
_POOL = None
_NTHREADS = 1


def set_threads(n):
    """Row-block threads for matvec / line solves / property evaluation (1 = serial)."""
    global _POOL, _NTHREADS
    from concurrent.futures import ThreadPoolExecutor
    n = max(1, int(n))
    if _POOL is not None:
        _POOL.shutdown()
        _POOL = None
    _NTHREADS = n
    if n > 1:
        _POOL = ThreadPoolExecutor(max_workers=n)


def _row_blocks(nz, min_rows=64):
    nb = min(_NTHREADS, max(1, nz // min_rows))
    edges = [nz * k // nb for k in range(nb + 1)]
    return [(a, b) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def _par_rows(fn, nz, min_rows=64):
    """fn(a, b) on row blocks [a, b) covering [0, nz), threaded when set_threads(n > 1)."""
    blocks = _row_blocks(nz, min_rows)
    if _POOL is None or len(blocks) == 1:
        for a, b in blocks:
            fn(a, b)
        return
    list(_POOL.map(lambda ab: fn(*ab), blocks))

# z-flag bits (cells / radial faces use z_c[j]; axial faces use z_v[j+1])
ZF_VC_EVAPCOND = 1   # z < L_e  or  z > L_e+L_a                 main.py:65,87
ZF_VC_ADIAB = 2      # L_e < z < L_e+L_a                         main.py:66,88
ZF_BEFORE_COND = 4   # z < L_e+L_a                               main.py:68,71
ZF_IN_COND = 8       # z > L_e+L_a                               main.py:69,72
ZF_BC_EVAP = 16      # L_input_left < z < L_input_right          main.py:123
ZF_BC_COND = 32      # L_e+L_a < z < L_total                     main.py:124

BAND_VC, BAND_WICK, BAND_WALL, BAND_NONE = 0, 1, 2, 3


def default_config_path():
    return os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'config', 'faghri.json')


def load_config(path=None):
    with open(path or default_config_path()) as f:
        return json.load(f)


# ---------------------------------------------------------------------------
# grid (mesh.py:70-109 + FiPy CylindricalGrid2D geometry)
# ---------------------------------------------------------------------------
@dataclass
class Grid:
    r_v: np.ndarray
    z_v: np.ndarray

    def __post_init__(self):
        self.r_v = np.asarray(self.r_v, dtype=np.float64)
        self.z_v = np.asarray(self.z_v, dtype=np.float64)
        self.nr = self.r_v.size - 1
        self.nz = self.z_v.size - 1
        self.r_c = 0.5 * (self.r_v[:-1] + self.r_v[1:])
        self.z_c = 0.5 * (self.z_v[:-1] + self.z_v[1:])
        self.dr = np.diff(self.r_v)
        self.dz = np.diff(self.z_v)

    @property
    def n_cells(self):
        return self.nr * self.nz


def build_grid(mesh_params, dimensions):
    """Vertex recipe of mesh.py:91-106 followed by FiPy's cumulative-sum vertices."""
    nz = int(mesh_params['nx_wall'])
    counts = (int(mesh_params['nr_vc']), int(mesh_params['nr_wick']), int(mesh_params['nr_wall']))
    radii = (dimensions['R_vc'], dimensions['R_wick'], dimensions['R_wall'])
    widths = (radii[0] / counts[0], (radii[1] - radii[0]) / counts[1], (radii[2] - radii[1]) / counts[2])
    verts = [0.0]
    for w, c in zip(widths, counts):                       # mesh.py:98-101 python running sum
        for _ in range(c):
            verts.append(verts[-1] + w)
    dr = np.diff(verts)                                    # mesh.py:102
    L_total = sum([dimensions['L_e'], dimensions['L_a'], dimensions['L_c']])
    dz = np.diff(np.linspace(0, L_total, nz + 1))          # mesh.py:105-106
    r_v = np.concatenate(([0.0], np.cumsum(dr)))           # FiPy vertices = cumsum of spacings
    z_v = np.concatenate(([0.0], np.cumsum(dz)))
    return Grid(r_v, z_v)


# ---------------------------------------------------------------------------
# case description (physics + mode switches)
# ---------------------------------------------------------------------------
@dataclass
class Case:
    dimensions: dict
    parameters: dict
    constants: dict
    h: float = 5.0                        # main.py:37
    properties: str = 'temperature_dependent'   # | 'simple' (main.py:41-46)
    gamma: str = 'snapshot'               # | 'lazy'   (main.py:133)
    source: str = 'q_over_k'              # | 'q'      (main.py:150)
    face_k: str = 'masked_at_Tface'       # | 'harmonic'
    radiation: bool = False               # verification_examples.py:81-91
    # optional per-axial-station overrides (ensembles): arrays of length nz or None
    q_line: np.ndarray | None = None
    h_line: np.ndarray | None = None
    Tamb_line: np.ndarray | None = None
    eps_line: np.ndarray | None = None
    axial_break: np.ndarray | None = None   # bool[nz-1]: True -> no face between j and j+1
    simple_rho: float = 7200.0
    simple_cp: float = 440.5
    simple_k: float = 55.0

    @property
    def T_amb(self):
        return float(self.parameters['T_amb'])

    @property
    def q(self):
        return float(self.parameters['Q_input_flux'])


def case_from_config(path=None, **kw):
    cfg = load_config(path)
    return Case(dimensions=dict(cfg['dimensions']), parameters=dict(cfg['parameters']),
                constants=dict(cfg['constants']), **kw)


# ---------------------------------------------------------------------------
# masks  (main.py:64-72, 86-90, 123-124)
# ---------------------------------------------------------------------------
def cell_band(r, d):
    """Cells: strict inequalities of main.py:64,67,70."""
    b = np.full(r.shape, BAND_NONE, dtype=np.int8)
    b[r < d['R_vc']] = BAND_VC
    b[(d['R_vc'] < r) & (r < d['R_wick'])] = BAND_WICK
    b[r > d['R_wick']] = BAND_WALL
    return b


def face_band(r, d):
    """Faces: main.py:86,89,90 (note `<=` for the vapor core)."""
    b = np.full(r.shape, BAND_NONE, dtype=np.int8)
    b[r <= d['R_vc']] = BAND_VC
    b[(d['R_vc'] < r) & (r < d['R_wick'])] = BAND_WICK
    b[r > d['R_wick']] = BAND_WALL
    return b


def z_flags(z, d):
    Le, La, Lc = d['L_e'], d['L_a'], d['L_c']
    L_total = Le + La + Lc                                   # main.py:25
    f = np.zeros(z.shape, dtype=np.int8)
    f[(z < Le) | (z > (Le + La))] |= ZF_VC_EVAPCOND
    f[(z > Le) & (z < Le + La)] |= ZF_VC_ADIAB
    f[z < Le + La] |= ZF_BEFORE_COND
    f[z > (Le + La)] |= ZF_IN_COND
    f[(z < d['L_input_right']) & (z > d['L_input_left'])] |= ZF_BC_EVAP
    f[(z > (Le + La)) & (z < L_total)] |= ZF_BC_COND
    return f


@dataclass
class Masks:
    cell_band: np.ndarray      # [nr]   band of cell centres
    rface_band: np.ndarray     # [nr]   band of the radial face at r_v[i+1] (i = nr-1: outer wall surface)
    zface_band: np.ndarray     # [nr]   band of axial faces (face centre r = r_c[i], face inequalities)
    zc_flags: np.ndarray       # [nz]   flags at z_c[j]   (cells, radial faces, outer-surface faces)
    zv_flags: np.ndarray       # [nz]   flags at z_v[j+1] (axial face above cell row j; last entry unused)


def build_masks(grid: Grid, d: dict) -> Masks:
    return Masks(cell_band=cell_band(grid.r_c, d), rface_band=face_band(grid.r_v[1:], d),
                 zface_band=face_band(grid.r_c, d), zc_flags=z_flags(grid.z_c, d),
                 zv_flags=z_flags(grid.z_v[1:], d))


# ---------------------------------------------------------------------------
# properties on the grid
# ---------------------------------------------------------------------------
def _k_by_band(case: Case, T, band, zflag):
    """k(T) for a 2-D block whose columns carry `band[i]` and rows `zflag[j]` (main.py:103-106)."""
    T = np.asarray(T, dtype=np.float64)
    if case.properties == 'simple':
        return np.full(T.shape, case.simple_k)
    pr, dm, cs = case.parameters, case.dimensions, case.constants
    out = np.zeros(T.shape)
    cols = np.nonzero(band == BAND_VC)[0]
    if cols.size:
        Tv = T[:, cols]
        ec = ((zflag & ZF_VC_EVAPCOND) != 0).astype(np.float64)[:, None]
        ad = ((zflag & ZF_VC_ADIAB) != 0).astype(np.float64)[:, None]
        out[:, cols] = (P.vc_k(Tv, 'evap_cond', dm['R_vc'], pr['M_g_sodium'], cs['R']) * ec
                        + P.vc_k(Tv, 'adiabatic', dm['R_vc'], pr['M_g_sodium'], cs['R']) * ad)
    cols = np.nonzero(band == BAND_WICK)[0]
    if cols.size:
        out[:, cols] = P.wick_k(T[:, cols], pr['porosity_wick'], pr['T_melting_sodium'], pr['delta_T_sodium'])
    cols = np.nonzero(band == BAND_WALL)[0]
    if cols.size:
        out[:, cols] = P.steel_k(T[:, cols])
    return out


def rho_cp_cells(case: Case, masks: Masks, T):
    """(rho, cp) per cell, main.py:108-117."""
    T = np.asarray(T, dtype=np.float64)
    if case.properties == 'simple':
        return np.full(T.shape, case.simple_rho), np.full(T.shape, case.simple_cp)
    pr = case.parameters
    rho = np.zeros(T.shape)
    cp = np.zeros(T.shape)
    zf = masks.zc_flags
    cols = np.nonzero(masks.cell_band == BAND_VC)[0]
    if cols.size:
        # cp = f*evap_cond_mask + f*adiabatic_mask  (main.py:108-109,113-114)
        w = (((zf & ZF_VC_EVAPCOND) != 0).astype(np.float64) + ((zf & ZF_VC_ADIAB) != 0).astype(np.float64))[:, None]
        cp[:, cols] = P.vc_cp(T[:, cols]) * w
        rho[:, cols] = P.vc_rho(T[:, cols]) * w
    cols = np.nonzero(masks.cell_band == BAND_WICK)[0]
    if cols.size:
        a = (pr['porosity_wick'], pr['T_melting_sodium'], pr['delta_T_sodium'])
        cp[:, cols] = P.wick_cp(T[:, cols], *a)
        rho[:, cols] = P.wick_rho(T[:, cols], *a)
    cols = np.nonzero(masks.cell_band == BAND_WALL)[0]
    if cols.size:
        Tw = T[:, cols]
        cp[:, cols] = P.steel_cp(Tw)
        ea = ((zf & ZF_BEFORE_COND) != 0).astype(np.float64)[:, None]
        cd = ((zf & ZF_IN_COND) != 0).astype(np.float64)[:, None]
        rho[:, cols] = P.steel_rho_eff(Tw) * ea + P.steel_rho(Tw) * cd    # main.py:116-117
    return rho, cp


def k_cells(case: Case, masks: Masks, T):
    """Cell-centred k (only used by face_k='harmonic')."""
    return _k_by_band(case, T, masks.cell_band, masks.zc_flags)


def face_conductivities(case: Case, grid: Grid, masks: Masks, T):
    """k on radial faces [nz,nr-1], axial faces [nz-1,nr] and the outer surface [nz]."""
    T = np.asarray(T, dtype=np.float64)
    rc, zc, rv, zv = grid.r_c, grid.z_c, grid.r_v, grid.z_v
    if case.face_k == 'masked_at_Tface':
        ar = (rv[1:-1] - rc[:-1]) / (rc[1:] - rc[:-1])          # FiPy _faceToCellDistanceRatio
        az = (zv[1:-1] - zc[:-1]) / (zc[1:] - zc[:-1])
        Tr = T[:, :-1] + (T[:, 1:] - T[:, :-1]) * ar[None, :]   # arithmeticFaceValue
        Tz = T[:-1, :] + (T[1:, :] - T[:-1, :]) * az[:, None]
        k_r = _k_by_band(case, Tr, masks.rface_band[:-1], masks.zc_flags)
        k_z = _k_by_band(case, Tz, masks.zface_band, masks.zv_flags[:-1])
        k_out = _k_by_band(case, T[:, -1:], masks.rface_band[-1:], masks.zc_flags)[:, 0]
    elif case.face_k == 'harmonic':
        kc = k_cells(case, masks, T)
        d1 = (rv[1:-1] - rc[:-1])[None, :]
        d2 = (rc[1:] - rv[1:-1])[None, :]
        with np.errstate(divide='ignore', invalid='ignore'):
            k_r = (d1 + d2) / (d1 / kc[:, :-1] + d2 / kc[:, 1:])
            e1 = (zv[1:-1] - zc[:-1])[:, None]
            e2 = (zc[1:] - zv[1:-1])[:, None]
            k_z = (e1 + e2) / (e1 / kc[:-1, :] + e2 / kc[1:, :])
        k_r = np.where(np.isfinite(k_r), k_r, 0.0)
        k_z = np.where(np.isfinite(k_z), k_z, 0.0)
        k_out = kc[:, -1].copy()
    else:
        raise ValueError(case.face_k)
    return k_r, k_z, k_out


def _properties(case: Case, grid: Grid, masks: Masks, T):
    """(k_r, k_z, k_out, rho, cp) at T; evaluated per row block on the thread pool when set_threads(n > 1) (each block is
    the same expression on a sub-grid of whole rows, so the values are those of the serial evaluation)."""
    if _POOL is None:
        k_r, k_z, k_out = face_conductivities(case, grid, masks, T)
        rho, cp = rho_cp_cells(case, masks, T)
        return k_r, k_z, k_out, rho, cp
    nz, nr = T.shape
    k_r, k_z, k_out = np.empty((nz, nr - 1)), np.empty((nz - 1, nr)), np.empty(nz)
    rho, cp = np.empty((nz, nr)), np.empty((nz, nr))

    def block(a, b):
        b2 = min(b + 1, nz)
        sub = Grid(grid.r_v, grid.z_v[a:b2 + 1])
        sm = Masks(masks.cell_band, masks.rface_band, masks.zface_band, masks.zc_flags[a:b2], masks.zv_flags[a:b2])
        kr, kz, ko = face_conductivities(case, sub, sm, T[a:b2])
        k_r[a:b] = kr[:b - a]
        k_out[a:b] = ko[:b - a]
        k_z[a:b2 - 1] = kz
        smc = Masks(masks.cell_band, masks.rface_band, masks.zface_band, masks.zc_flags[a:b], masks.zv_flags[a:b])
        rho[a:b], cp[a:b] = rho_cp_cells(case, smc, T[a:b])
    _par_rows(block, nz, min_rows=16)
    return k_r, k_z, k_out, rho, cp


# ---------------------------------------------------------------------------
# assembly of one sweep
# ---------------------------------------------------------------------------
@dataclass
class System:
    cR: np.ndarray     # [nz,nr]  coupling (j,i)<->(j,i+1); last column 0
    cZ: np.ndarray     # [nz,nr]  coupling (j,i)<->(j+1,i); last row 0
    diag: np.ndarray   # [nz,nr]
    rhs: np.ndarray    # [nz,nr]

    def matvec(self, x):
        x = x.reshape(self.diag.shape)
        nz = x.shape[0]
        if _POOL is None:
            y = self.diag * x
            y[:, :-1] -= self.cR[:, :-1] * x[:, 1:]
            y[:, 1:] -= self.cR[:, :-1] * x[:, :-1]
            y[:-1, :] -= self.cZ[:-1, :] * x[1:, :]
            y[1:, :] -= self.cZ[:-1, :] * x[:-1, :]
            return y
        y = np.empty_like(x)
        cR, cZ, dg = self.cR, self.cZ, self.diag

        def block(a, b):
            yb = dg[a:b] * x[a:b]
            yb[:, :-1] -= cR[a:b, :-1] * x[a:b, 1:]
            yb[:, 1:] -= cR[a:b, :-1] * x[a:b, :-1]
            hi = min(b, nz - 1)
            yb[:hi - a, :] -= cZ[a:hi, :] * x[a + 1:hi + 1, :]
            lo = max(a, 1)
            yb[lo - a:, :] -= cZ[lo - 1:b - 1, :] * x[lo - 1:b - 1, :]
            y[a:b] = yb
        _par_rows(block, nz)
        return y

    def to_csr(self):
        nz, nr = self.diag.shape
        n = nz * nr
        offd_r = (-self.cR).ravel()[:-1]                 # entries coupling id, id+1 (zero at line ends)
        offd_z = (-self.cZ).ravel()[:n - nr]
        return sp.diags([self.diag.ravel(), offd_r, offd_r, offd_z, offd_z],
                        [0, 1, -1, nr, -nr], shape=(n, n), format='csc')


class Oracle:
    """Holds grid, masks, snapshots and the iterate; `sweep()` mirrors FiPy's `eq.sweep`."""

    def __init__(self, grid: Grid, case: Case, T0=None, has_old=False):
        self.grid, self.case = grid, case
        self.masks = build_masks(grid, case.dimensions)
        nz, nr = grid.nz, grid.nr
        T_init = case.T_amb if T0 is None else T0
        self.T = np.broadcast_to(np.asarray(T_init, dtype=np.float64), (nz, nr)).copy() \
            if np.ndim(T_init) < 2 else np.array(T_init, dtype=np.float64).reshape(nz, nr)
        self.has_old = has_old
        self.T_old = self.T.copy()
        g = grid
        self.vol = g.r_c[None, :] * g.dr[None, :] * g.dz[:, None]              # per radian
        self.geo_r = g.r_v[1:-1][None, :] * g.dz[:, None] / (g.r_c[1:] - g.r_c[:-1])[None, :]   # A/d radial
        self.geo_z = (g.r_c * g.dr)[None, :] / (g.z_c[1:] - g.z_c[:-1])[:, None]                # A/d axial
        self.A_out = g.r_v[-1] * g.dz
        self.d_out = g.r_v[-1] - g.r_c[-1]
        # snapshots taken while T == T_amb everywhere (main.py:34,133,145)
        Tsnap = np.full((nz, nr), case.T_amb)
        if case.Tamb_line is not None:
            Tsnap = np.broadcast_to(np.asarray(case.Tamb_line, dtype=np.float64)[:, None], (nz, nr)).copy()
        self._snap = face_conductivities(case, grid, self.masks, Tsnap)
        self.n_sweeps = 0
        self.last_solver_info = {}

    # per-line BC parameters
    def _line(self, arr, scalar):
        nz = self.grid.nz
        if arr is None:
            return np.full(nz, float(scalar))
        return np.asarray(arr, dtype=np.float64).reshape(nz)

    def update_old(self):
        self.T_old = self.T.copy()

    def assemble(self, dt) -> System:
        g, c, m = self.grid, self.case, self.masks
        T = self.T
        T_old = self.T_old if self.has_old else T
        k_r, k_z, k_out, rho, cp = _properties(c, g, m, T)
        if c.gamma == 'snapshot':
            G_r, G_z, k0_out = self._snap
        elif c.gamma == 'lazy':
            G_r, G_z, k0_out = k_r, k_z, k_out
        else:
            raise ValueError(c.gamma)
        nz, nr = g.nz, g.nr
        cR = np.zeros((nz, nr))
        cZ = np.zeros((nz, nr))
        cR[:, :-1] = G_r * self.geo_r
        cZ[:-1, :] = G_z * self.geo_z
        if c.axial_break is not None:
            cZ[:-1, :][np.asarray(c.axial_break, dtype=bool), :] = 0.0
        cap = cp * rho * self.vol / dt
        diag = cap.copy()
        diag[:, :-1] += cR[:, :-1]
        diag[:, 1:] += cR[:, :-1]
        diag[:-1, :] += cZ[:-1, :]
        diag[1:, :] += cZ[:-1, :]
        rhs = cap * T_old
        # outer surface: Robin (condenser) + flux (evaporator), main.py:123-150
        h = self._line(c.h_line, c.h)
        q = self._line(c.q_line, c.q)
        Ta = self._line(c.Tamb_line, c.T_amb)
        cond = ((m.zc_flags & ZF_BC_COND) != 0).astype(np.float64)
        evap = ((m.zc_flags & ZF_BC_EVAP) != 0).astype(np.float64)
        robin = cond * k_out * self.A_out / (h * self.d_out + k0_out)     # RobinCoeff . n A
        diag[:, -1] += robin * h                                         # ImplicitSourceTerm
        g_src = h * Ta
        if c.radiation:
            eps = self._line(c.eps_line, c.parameters['emissivity'])
            sigma = c.constants['sigma']
            g_src = g_src + sigma * eps * (Ta ** 4 - T[:, -1] ** 4)       # verification_examples.py:86-88
        rhs[:, -1] += robin * g_src
        if c.source == 'q_over_k':
            with np.errstate(divide='ignore', invalid='ignore'):
                flux = np.where(evap != 0, q / k_out, 0.0)
            rhs[:, -1] += evap * flux * self.A_out
        elif c.source == 'q':
            rhs[:, -1] += evap * q * self.A_out
        else:
            raise ValueError(c.source)
        return System(cR, cZ, diag, rhs)

    def sweep(self, dt, solver='auto', **solver_kw):
        """Assemble at the current iterate, return the PRE-solve ||L T - b||_2, overwrite T."""
        S = self.assemble(dt)
        res = float(np.linalg.norm((S.matvec(self.T) - S.rhs).ravel()))
        self.T = solve_system(S, x0=self.T, solver=solver, info=self.last_solver_info, **solver_kw)
        self.n_sweeps += 1
        return res


# ---------------------------------------------------------------------------
# linear solvers
# ---------------------------------------------------------------------------
def solve_system(S: System, x0=None, solver='auto', info=None, **kw):
    nz, nr = S.diag.shape
    if solver == 'auto':
        solver = 'banded' if nr <= 160 else 'pcg_rline'
    if solver == 'banded':
        # symmetric positive definite, half-bandwidth nr (r fastest): LAPACK dpbsv
        n = nz * nr
        ab = np.zeros((nr + 1, n))
        ab[0] = S.diag.ravel()
        ab[1, :-1] = (-S.cR).ravel()[:-1]
        ab[nr, :n - nr] = (-S.cZ).ravel()[:n - nr]
        x = sla.solveh_banded(ab, S.rhs.ravel(), lower=True, check_finite=False)
        return x.reshape(nz, nr)
    if solver == 'splu':
        return spla.splu(S.to_csr()).solve(S.rhs.ravel()).reshape(nz, nr)
    if solver in ('pcg_rline', 'pcg_jacobi', 'pcg_mgline'):
        return pcg(S, x0=x0, precond=solver[4:], info=info, **kw)
    raise ValueError(solver)


def rline_factor(S: System):
    """L D L^T of the radial-line tridiagonal part (LAPACK dpttrf on the block-diagonal chain)."""
    d = S.diag.ravel().copy()
    e = (-S.cR).ravel()[:-1].copy()
    d, e, inf = sla.lapack.dpttrf(d, e)
    if inf != 0:
        raise np.linalg.LinAlgError(f'dpttrf info={inf}')
    return d, e



# ---------------------------------------------------------------------------
# 'mgline' preconditioner twin: V(1,1), axial semi-coarsening by row pairs (piecewise-constant aggregation; radial couplings
# Galerkin, axial coupling scaled by `zscale`: 1 = Galerkin, 1/2 = a discretisation on the coarse rows), damped radial-line
# Jacobi smoother -- the same cycle as csrc/hp_kernels.cu (k_mg_coarsen / k_mg_res / k_mg_line)
# ---------------------------------------------------------------------------
class _MgLevel:
    def __init__(self, cR, cZ, diag):
        self.S = System(cR, cZ, diag, np.zeros_like(diag))
        self.fd, self.fe = rline_factor(self.S)

    def line(self, v):
        if _POOL is None:
            y, _ = sla.lapack.dpttrs(self.fd, self.fe, v.ravel().copy())
            return y.reshape(v.shape)
        # the chain decouples at every row end (cR = 0 there): row blocks are independent tridiagonal solves
        nz, nr = v.shape
        out = np.empty_like(v)
        fd, fe = self.fd, self.fe

        def block(a, b):
            y, _ = sla.lapack.dpttrs(fd[a * nr:b * nr], fe[a * nr:b * nr - 1], v[a:b].ravel().copy())
            out[a:b] = y.reshape(b - a, nr)
        _par_rows(block, nz)
        return out

    def coarsen(self, zscale=1.0):
        cR, cZ, dg = self.S.cR, self.S.cZ, self.S.diag
        nz = dg.shape[0]
        nzc = (nz + 1) // 2
        a = slice(0, nz, 2)
        cRc, cZc, dgc = cR[a].copy(), np.zeros((nzc, cR.shape[1])), dg[a].copy()
        nb = nz // 2                                    # complete pairs
        cRc[:nb] += cR[1::2]
        dgc[:nb] += dg[1::2] - 2.0 * cZ[0:2 * nb:2]
        top = np.where(np.arange(nzc) < nb, 2 * np.arange(nzc) + 1, 2 * np.arange(nzc))   # top fine row of every pair
        cZc[:] = cZ[top]
        cZc[-1, :] = 0.0
        if zscale != 1.0:
            # weaker axial coupling, row sums (capacity + boundary terms) kept: diag_c -= (1 - zs) (c_up + c_down)
            dgc[:-1] -= (1.0 - zscale) * cZc[:-1]
            dgc[1:] -= (1.0 - zscale) * cZc[:-1]
            cZc *= zscale
        return _MgLevel(cRc, cZc, dgc)


def mg_hierarchy(S: System, max_levels=6, zscale=0.5):
    cZ = S.cZ.copy()
    cZ[-1, :] = 0.0
    levels = [_MgLevel(S.cR, cZ, S.diag)]
    nzl = S.diag.shape[0]
    while len(levels) < max_levels and nzl >= 32:
        levels.append(levels[-1].coarsen(zscale))
        nzl = (nzl + 1) // 2
    return levels


def _restrict(r):
    nz = r.shape[0]
    out = r[0::2].copy()
    out[:nz // 2] += r[1::2]
    return out


def _prolong(e, nz):
    return np.repeat(e, 2, axis=0)[:nz]


def mg_vcycle(levels, r, zline, omega=0.6, coarse_sweeps=2):
    """z = V(r); `zline` = M_line^-1 r of the fine level (what k_cg_B leaves in z)."""
    c = len(levels) - 1
    xs, rs = [omega * zline], [r]
    rs.append(_restrict(rs[0] - levels[0].S.matvec(xs[0])))
    for l in range(1, c):
        xs.append(omega * levels[l].line(rs[l]))
        rs.append(_restrict(rs[l] - levels[l].S.matvec(xs[l])))
    xc = omega * levels[c].line(rs[c])
    for _ in range(coarse_sweeps):
        xc = xc + omega * levels[c].line(rs[c] - levels[c].S.matvec(xc))
    e = xc
    for l in range(c - 1, -1, -1):
        xt = xs[l] + _prolong(e, xs[l].shape[0])
        e = xt + omega * levels[l].line(rs[l] - levels[l].S.matvec(xt))
    return e


def pcg(S: System, x0=None, precond='rline', tol=1e-9, criterion='precond_inf', max_iter=10000, info=None, mg=None):
    """Preconditioned CG restating the GPU algorithm (same preconditioners, same stopping rules, same order
    of operations: x/r update, z = M^-1 r, test, direction update).

    criterion 'precond_inf'  : stop when max_i |(M^-1 r)_i| <= tol   (Kelvin units; the GPU default)
    criterion 'rowscaled_inf': stop when max_i |r_i / a_ii| <= tol   (Kelvin units)
    criterion 'l2_rel_rhs'   : stop when ||r||_2 <= tol * ||b||_2
    """
    shape = S.diag.shape
    b = S.rhs
    x = np.zeros(shape) if x0 is None else np.array(x0, dtype=np.float64).reshape(shape)
    r = b - S.matvec(x)
    if precond == 'rline':
        fd, fe = rline_factor(S)

        def apply_M(v):
            y, inf = sla.lapack.dpttrs(fd, fe, v.ravel().copy())
            return y.reshape(shape)
    elif precond == 'jacobi':
        dinv = 1.0 / S.diag

        def apply_M(v):
            return v * dinv
    elif precond == 'mgline':
        return _pcg_mgline(S, x, r, tol, criterion, max_iter, info, **(mg or {}))
    else:
        raise ValueError(precond)
    bnorm = float(np.linalg.norm(b.ravel()))

    def measure(rr, zz):
        if criterion == 'precond_inf':
            return float(np.max(np.abs(zz)))
        if criterion == 'rowscaled_inf':
            return float(np.max(np.abs(rr / S.diag)))
        if criterion == 'l2_rel_rhs':
            return float(np.linalg.norm(rr.ravel())) / bnorm if bnorm > 0 else 0.0
        raise ValueError(criterion)

    it = 0
    z = apply_M(r)
    crit = measure(r, z)
    if crit > tol:
        p = z.copy()
        rz = float(np.vdot(r, z))
        while it < max_iter:
            q = S.matvec(p)
            alpha = rz / float(np.vdot(p, q))
            x += alpha * p
            r -= alpha * q
            it += 1
            z = apply_M(r)
            crit = measure(r, z)
            if crit <= tol:
                break
            rz_new = float(np.vdot(r, z))
            p = z + (rz_new / rz) * p
            rz = rz_new
    if info is not None:
        info['iterations'] = it
        info['crit'] = crit
    return x


def _pcg_mgline(S, x, r, tol, criterion, max_iter, info, max_levels=6, omega=0.6, coarse_sweeps=2, zscale=0.5):
    """PCG with the V-cycle as preconditioner; the stopping measure stays max|M_line^-1 r| (same rule as 'rline')."""
    if criterion != 'precond_inf':
        raise ValueError("mgline twin implements the precond_inf criterion")
    levels = mg_hierarchy(S, max_levels, zscale)
    if len(levels) < 2:
        raise ValueError("mesh too short to coarsen")
    it = 0
    zl = levels[0].line(r)
    crit = float(np.max(np.abs(zl)))
    if crit > tol:
        z = mg_vcycle(levels, r, zl, omega, coarse_sweeps)
        p = z.copy()
        rz = float(np.vdot(r, z))
        while it < max_iter:
            q = S.matvec(p)
            alpha = rz / float(np.vdot(p, q))
            x += alpha * p
            r -= alpha * q
            it += 1
            zl = levels[0].line(r)
            crit = float(np.max(np.abs(zl)))
            if crit <= tol:
                break
            z = mg_vcycle(levels, r, zl, omega, coarse_sweeps)
            rz_new = float(np.vdot(r, z))
            p = z + (rz_new / rz) * p
            rz = rz_new
    if info is not None:
        info['iterations'] = it
        info['crit'] = crit
        info['levels'] = len(levels)
    return x


# ---------------------------------------------------------------------------
# time loop (main.py:119-120,157-159,200-259,267-271)
# ---------------------------------------------------------------------------
def run_transient(grid: Grid, case: Case, t_end=101.0, dt=0.1, sweeps=1, has_old=False, measure_times=None,
                  solver='auto', T0=None, keep_fields_at=(), **solver_kw):
    orc = Oracle(grid, case, T0=T0, has_old=has_old)
    if measure_times is None:
        measure_times = list(range(5, int(t_end), 5))                   # main.py:158-159
    times = np.arange(0, t_end, dt)                                     # main.py:200
    profiles, prof_times, residuals, fields = [], [], [], {}
    nxt = 0
    keep = set(int(s) for s in keep_fields_at)
    for step, t in enumerate(times):
        if has_old:
            orc.update_old()                                            # main.py:269
        for _ in range(sweeps):
            res = orc.sweep(dt, solver=solver, **solver_kw)             # main.py:202 / :271
        residuals.append(res)
        if step in keep:
            fields[step] = orc.T.copy()
        if nxt < len(measure_times) and t >= measure_times[nxt]:        # main.py:204-209
            if not prof_times or prof_times[-1] < t:
                profiles.append(orc.T[:, -1].copy())                    # top-wall cells main.py:165-169
                prof_times.append(float(t))
            nxt += 1
            while nxt < len(measure_times) and measure_times[nxt] <= t:
                nxt += 1
    return dict(T=orc.T, profile_times=np.array(prof_times), profiles=np.array(profiles),
                residuals=np.array(residuals), fields=fields, n_steps=len(times), oracle=orc)


def synthetic_mesh_params(nz, nr_vc, nr_wick, nr_wall):
    return dict(nx_wall=nz, nr_wall=nr_wall, nx_wick=nz, nr_wick=nr_wick, nx_vc=nz, nr_vc=nr_vc)


def clone_case(case: Case, **kw) -> Case:
    c = copy.deepcopy(case)
    for k, v in kw.items():
        setattr(c, k, v)
    return c
