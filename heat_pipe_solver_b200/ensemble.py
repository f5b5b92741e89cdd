"""Batched parameter ensembles of independent heat pipes (BASELINE config 5).

M pipes of identical geometry but different heat input / condenser conditions are stacked along the
axial index into one (M*nz) x nr system whose member blocks are uncoupled (`axial_open = 0` at member
boundaries, per-row q / h / T_amb / emissivity).  The same kernels as for one pipe then stream the whole
ensemble HBM-bound instead of latency-bound; the PCG scalars are shared, which is still CG on the
block-diagonal system and converges when the slowest member does.  Members shard over GPUs with no
collective (`shard_members`).
"""
from __future__ import annotations

import numpy as np

from .mesh import StructuredGrid2D, composite_spacings
from .regions import Regions
from .solver import ConductionSolver


def shard_members(n_members, rank, world):
    """Contiguous block of members for `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_members, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class _StackedLayout:
    """Ghost-inclusive arrays of M stacked copies of one pipe (what hp_mesh_desc wants)."""

    def __init__(self, mesh, regions, members, lines):
        nz1, M = mesh.nz, members
        self.nr, self.nz = mesh.nr, nz1 * M
        self.j0, self.j1 = 0, self.nz
        dz = np.tile(mesh.dy, M)
        zv = np.concatenate(([0.0], np.cumsum(dz)))
        self.z_v = np.ascontiguousarray(np.concatenate(([-dz[0]], zv, [zv[-1] + dz[-1]])))
        zc = np.tile(regions.zc_flags, M)
        self.zc_flags = np.ascontiguousarray(np.concatenate((zc[:1], zc, zc[-1:])), dtype=np.uint8)
        # face between ghost-inclusive row g and g+1: bottom face of member row (g mod nz1)
        zvf = np.tile(regions.zv_flags[:-1], M)
        self.zv_flags = np.ascontiguousarray(np.concatenate((zvf, regions.zv_flags[-1:], [0])), dtype=np.uint8)
        is_open = np.ones(self.nz + 2, dtype=np.uint8)
        is_open[np.arange(0, self.nz + 1, nz1)] = 0          # pipe ends
        is_open[self.nz + 1] = 0
        self.axial_open = is_open
        self.lines = {}
        for k, v in lines.items():
            full = np.repeat(np.asarray(v, dtype=np.float64), nz1)
            self.lines[k] = np.ascontiguousarray(np.concatenate((full[:1], full, full[-1:])))


class EnsembleSolver:
    """M independent Faghri-geometry pipes advanced together on one GPU."""

    def __init__(self, mesh_params, dimensions, parameters, constants, *, q, h, T_amb, emissivity=None, **solver_kw):
        q, h, T_amb = (np.atleast_1d(np.asarray(a, dtype=np.float64)) for a in (q, h, T_amb))
        M = q.size
        if h.size != M or T_amb.size != M:
            raise ValueError("q, h and T_amb need one entry per member")
        eps = np.full(M, parameters.get('emissivity', 0.0)) if emissivity is None else np.asarray(emissivity, dtype=np.float64)
        mp = dict(mesh_params)
        dr, dz = composite_spacings(mp, dimensions)
        self.member_mesh = StructuredGrid2D(dx=dr, dy=dz, cylindrical=True)
        self.members = M
        regions = Regions(self.member_mesh, dimensions)
        layout = _StackedLayout(self.member_mesh, regions, M, dict(q=q, h=h, Tamb=T_amb, eps=eps))
        self.solver = ConductionSolver(self.member_mesh, dimensions, parameters, constants, layout=layout,
                                       regions=regions, **solver_kw)

    def run(self, dt, n_steps, sweeps=1, has_old=False):
        self.solver.run(dt, n_steps, sweeps=sweeps, has_old=has_old)

    def value(self):
        """(M, nz, nr) numpy array."""
        m = self.member_mesh
        return self.solver.value.reshape(self.members, m.nz, m.nr)

    def stats(self, n=1):
        return self.solver.stats(n)

    def close(self):
        self.solver.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
