"""Mesh builders with the reference's surface (`/root/reference/src/2d/mesh.py`).

``generate_mesh_2d`` (mesh.py:12-32) and ``generate_composite_mesh`` (mesh.py:42-149)
return a :class:`StructuredGrid2D` instead of a FiPy mesh: a tensor-product grid kept as
two 1-D vertex arrays (that is all the CUDA kernels need) which also answers the FiPy
attribute queries the reference's driver makes (main.py:30-32,123-143,165-169):
``cellCenters``, ``faceCenters``, ``numberOfCells``, ``facesRight`` ..., ``faceCellIDs``,
``cellFaceIDs``, ``faceNormals``, ``_faceToCellDistanceRatio``, ``_cellCenters``,
``_faceCenters``, ``cellVolumes``.  Those are materialised lazily in FiPy's ordering:

* cells: id = j*nx + i, first coordinate (x = r) fastest  -- `CylindricalGrid2D(dr->x, dz->y)`
* faces: all y-normal ("horizontal") faces first, nx*(ny+1) of them, id = j*nx + i;
  then the x-normal ("vertical") faces, (nx+1)*ny, id = nH + j*(nx+1) + i
* cellFaceIDs rows: bottom, right, top, left

Cylindrical geometry is per radian (no 2*pi): V = r_c dr dz, A_r = r_f dz, A_z = r_c dr.
"""
from __future__ import annotations

import numpy as np


class StructuredGrid2D:
    """Tensor-product grid. ``dx``/``nx`` is the fast (radial) axis, ``dy``/``ny`` the slow (axial) one."""

    def __init__(self, dx, dy, nx=None, ny=None, cylindrical=False):
        dx = np.atleast_1d(np.asarray(dx, dtype=np.float64))
        dy = np.atleast_1d(np.asarray(dy, dtype=np.float64))
        if dx.size == 1 and nx is not None:
            dx = np.full(int(nx), float(dx[0]))
        if dy.size == 1 and ny is not None:
            dy = np.full(int(ny), float(dy[0]))
        if nx is not None and int(nx) != dx.size or ny is not None and int(ny) != dy.size:
            raise ValueError("spacing arrays do not match nx / ny")
        if np.any(dx <= 0) or np.any(dy <= 0):
            raise ValueError("cell sizes must be positive")
        self.dx, self.dy = dx, dy
        self.nx, self.ny = dx.size, dy.size
        self.cylindrical = bool(cylindrical)
        # FiPy builds vertices as cumulative sums of the spacings
        self.x_v = np.concatenate(([0.0], np.cumsum(dx)))
        self.y_v = np.concatenate(([0.0], np.cumsum(dy)))
        self.x_c = 0.5 * (self.x_v[:-1] + self.x_v[1:])
        self.y_c = 0.5 * (self.y_v[:-1] + self.y_v[1:])
        self._cache = {}

    # --- the names the kernels use -------------------------------------------------
    @property
    def nr(self):
        return self.nx

    @property
    def nz(self):
        return self.ny

    r_v = property(lambda self: self.x_v)
    z_v = property(lambda self: self.y_v)
    r_c = property(lambda self: self.x_c)
    z_c = property(lambda self: self.y_c)

    @property
    def shape(self):
        """(nz, nr): fields reshape to this, radial index fastest."""
        return (self.ny, self.nx)

    # --- FiPy-like attribute surface -----------------------------------------------
    @property
    def numberOfCells(self):
        return self.nx * self.ny

    @property
    def numberOfHorizontalFaces(self):
        return self.nx * (self.ny + 1)

    @property
    def numberOfVerticalFaces(self):
        return (self.nx + 1) * self.ny

    @property
    def numberOfFaces(self):
        return self.numberOfHorizontalFaces + self.numberOfVerticalFaces

    def _lazy(self, key, fn):
        if key not in self._cache:
            self._cache[key] = fn()
        return self._cache[key]

    @property
    def cellCenters(self):
        def make():
            X, Y = np.meshgrid(self.x_c, self.y_c)       # (ny, nx), x fastest when ravelled
            return np.array([X.ravel(), Y.ravel()])
        return self._lazy('cc', make)

    _cellCenters = cellCenters

    @property
    def faceCenters(self):
        def make():
            Xh, Yh = np.meshgrid(self.x_c, self.y_v)     # horizontal faces
            Xv, Yv = np.meshgrid(self.x_v, self.y_c)     # vertical faces
            return np.array([np.concatenate((Xh.ravel(), Xv.ravel())),
                             np.concatenate((Yh.ravel(), Yv.ravel()))])
        return self._lazy('fc', make)

    _faceCenters = faceCenters

    def _hmask(self, rows):
        m = np.zeros((self.ny + 1, self.nx), dtype=bool)
        m[rows, :] = True
        return np.concatenate((m.ravel(), np.zeros(self.numberOfVerticalFaces, dtype=bool)))

    def _vmask(self, cols):
        m = np.zeros((self.ny, self.nx + 1), dtype=bool)
        m[:, cols] = True
        return np.concatenate((np.zeros(self.numberOfHorizontalFaces, dtype=bool), m.ravel()))

    @property
    def facesBottom(self):
        return self._lazy('fB', lambda: self._hmask(0))

    @property
    def facesTop(self):
        return self._lazy('fT', lambda: self._hmask(self.ny))

    @property
    def facesLeft(self):
        return self._lazy('fL', lambda: self._vmask(0))

    @property
    def facesRight(self):
        return self._lazy('fR', lambda: self._vmask(self.nx))

    @property
    def exteriorFaces(self):
        return self.facesBottom | self.facesTop | self.facesLeft | self.facesRight

    @property
    def interiorFaces(self):
        return ~self.exteriorFaces

    @property
    def cellFaceIDs(self):
        def make():
            j, i = np.divmod(np.arange(self.numberOfCells), self.nx)
            nH = self.numberOfHorizontalFaces
            bottom = j * self.nx + i
            top = (j + 1) * self.nx + i
            left = nH + j * (self.nx + 1) + i
            right = left + 1
            return np.array([bottom, right, top, left])
        return self._lazy('cf', make)

    @property
    def faceCellIDs(self):
        """(2, nFaces) masked array; exterior faces have their second entry masked (FiPy convention)."""
        def make():
            nx, ny = self.nx, self.ny
            ids = np.arange(nx * ny).reshape(ny, nx)
            h1 = np.vstack((ids[:1, :], ids))            # below face j is cell j-1 (bottom boundary: cell 0 row)
            h2 = np.vstack((ids, ids[-1:, :]))
            hm = np.zeros((ny + 1, nx), dtype=bool)
            hm[0, :] = hm[-1, :] = True
            v1 = np.hstack((ids[:, :1], ids))
            v2 = np.hstack((ids, ids[:, -1:]))
            vm = np.zeros((ny, nx + 1), dtype=bool)
            vm[:, 0] = vm[:, -1] = True
            first = np.concatenate((h1.ravel(), v1.ravel()))
            second = np.concatenate((h2.ravel(), v2.ravel()))
            mask = np.concatenate((hm.ravel(), vm.ravel()))
            return np.ma.array(np.array([first, second]), mask=np.array([np.zeros_like(mask), mask]))
        return self._lazy('fci', make)

    @property
    def faceNormals(self):
        def make():
            nH = self.numberOfHorizontalFaces
            n = np.zeros((2, self.numberOfFaces))
            n[1, :nH] = 1.0
            n[0, nH:] = 1.0
            n[1, :self.nx] = -1.0                                      # bottom boundary points outwards
            n[0, nH + np.arange(self.ny) * (self.nx + 1)] = -1.0       # left boundary points outwards
            return n
        return self._lazy('fn', make)

    @property
    def _cellDistances(self):
        def make():
            dyc = np.concatenate(([self.y_c[0] - self.y_v[0]], np.diff(self.y_c), [self.y_v[-1] - self.y_c[-1]]))
            dxc = np.concatenate(([self.x_c[0] - self.x_v[0]], np.diff(self.x_c), [self.x_v[-1] - self.x_c[-1]]))
            H = np.repeat(dyc[:, None], self.nx, axis=1)
            V = np.repeat(dxc[None, :], self.ny, axis=0)
            return np.concatenate((H.ravel(), V.ravel()))
        return self._lazy('cd', make)

    @property
    def _faceToCellDistanceRatio(self):
        """distance(face, first cell) / distance(cell1, cell2); 1 on exterior faces."""
        def make():
            fy = np.concatenate(([self.y_c[0] - self.y_v[0]], self.y_v[1:-1] - self.y_c[:-1], [self.y_v[-1] - self.y_c[-1]]))
            fx = np.concatenate(([self.x_c[0] - self.x_v[0]], self.x_v[1:-1] - self.x_c[:-1], [self.x_v[-1] - self.x_c[-1]]))
            H = np.repeat(fy[:, None], self.nx, axis=1)
            V = np.repeat(fx[None, :], self.ny, axis=0)
            return np.concatenate((H.ravel(), V.ravel())) / self._cellDistances
        return self._lazy('fr', make)

    @property
    def cellVolumes(self):
        def make():
            v = np.outer(self.dy, self.dx)
            if self.cylindrical:
                v = v * self.x_c[None, :]
            return v.ravel()
        return self._lazy('cv', make)

    @property
    def _faceAreas(self):
        def make():
            H = np.repeat(self.dx[None, :], self.ny + 1, axis=0)
            V = np.repeat(self.dy[:, None], self.nx + 1, axis=1)
            a = np.concatenate((H.ravel(), V.ravel()))
            if self.cylindrical:
                a = a * self.faceCenters[0]
            return a
        return self._lazy('fa', make)

    def __repr__(self):
        kind = 'CylindricalGrid2D' if self.cylindrical else 'Grid2D'
        return f"StructuredGrid2D<{kind} nr={self.nx} nz={self.ny}>"


def generate_mesh_2d(L_x, L_y, nx, ny):
    """Uniform Cartesian grid of L_x by L_y (reference mesh.py:12-32, FiPy ``Grid2D``)."""
    dx = L_x / nx
    dy = L_y / ny
    return StructuredGrid2D(dx=dx, dy=dy, nx=nx, ny=ny, cylindrical=False)


def composite_spacings(mesh_params, dimensions):
    """Spacing arrays the reference hands to ``CylindricalGrid2D`` (mesh.py:70-106)."""
    n_axial = mesh_params["nx_wall"]
    n_band = (mesh_params["nr_vc"], mesh_params["nr_wick"], mesh_params["nr_wall"])
    R = (dimensions["R_vc"], dimensions["R_wick"], dimensions["R_wall"])
    L_total = sum([dimensions["L_e"], dimensions["L_a"], dimensions["L_c"]])
    widths = (R[0] / n_band[0], (R[1] - R[0]) / n_band[1], (R[2] - R[1]) / n_band[2])
    ladder = [0.0]                                   # python running sum, then np.diff (mesh.py:98-102)
    for w, n in zip(widths, n_band):
        for _ in range(n):
            ladder.append(ladder[-1] + w)
    dr = np.diff(ladder)
    dz = np.diff(np.linspace(0, L_total, n_axial + 1))   # mesh.py:105-106
    return dr, dz


def generate_composite_mesh(mesh_params: dict, dimensions: dict):
    """Composite cylindrical heat-pipe mesh: uniform axial cells, three radial bands (mesh.py:42-149).

    Returns ``(mesh, cell_types)`` like the reference.  Side effect kept: ``mesh_params['nx_vc']`` and
    ``['nx_wick']`` are forced to ``nx_wall`` (mesh.py:86-88).

    ``cell_types`` reproduces the reference LITERALLY: it unpacks ``x, y = mesh.cellCenters`` and then
    treats ``x`` as axial and ``y`` as radial (mesh.py:113-122) although FiPy's first coordinate is the
    radius, so the codes it yields are 0 / 10 / 20 by *axial* position.  ``main.py`` never reads them.
    The intended codes (0/1 vapor core, 10/11 wick, 20/21 wall; +1 = adiabatic section) are available as
    ``mesh.region_codes``.
    """
    for key in ("nx_vc", "nx_wick"):
        if mesh_params[key] != mesh_params["nx_wall"]:
            mesh_params[key] = mesh_params["nx_wall"]
    dr, dz = composite_spacings(mesh_params, dimensions)
    mesh = StructuredGrid2D(dx=dr, dy=dz, nx=dr.size, ny=dz.size, cylindrical=True)

    R_vc, R_wick = dimensions["R_vc"], dimensions["R_wick"]
    L_e, L_a = dimensions["L_e"], dimensions["L_a"]

    def codes(radial, axial):
        vc = radial <= R_vc
        wick = (radial > R_vc) & (radial <= R_wick)
        wall = radial > R_wick
        adiabatic = (axial > L_e) & (axial < L_e + L_a)
        out = np.zeros(radial.shape, dtype=int)
        out[wick] = 10
        out[wall] = 20
        out[vc & adiabatic] = 1
        out[wick & adiabatic] = 11
        out[wall & adiabatic] = 21
        return out

    if mesh.numberOfCells <= 2_000_000:
        first, second = mesh.cellCenters
        cell_types = codes(radial=second, axial=first)        # the reference's swapped reading
        mesh.region_codes = codes(radial=first, axial=second)
    else:                                                     # keep huge synthetic meshes 1-D
        cell_types = None
        mesh.region_codes = None
    return mesh, cell_types
