"""The oracle's structured 5-point assembly against an INDEPENDENT face-list assembly of the same equation.

The oracle (oracle/fv_oracle.py) builds the system from separable band codes and structured (j, i) arrays.  Here the
equation of `/root/reference/src/2d/main.py:64-150` is assembled the way FiPy does it -- one pass over the flat FiPy-ordered
face list (`faceCellIDs`, `_faceAreas`, `_cellDistances`, `_faceToCellDistanceRatio`, `facesRight`, face / cell centres with
the coordinate masks of main.py:64-72,86-90,123-124 evaluated per face and per cell), with the property callables of the
package's host mirror (pinned to the reference's own functions by tests/test_properties.py).  Two derivations that share
neither the mask recipe nor the geometry formulas must give the same matrix and right-hand side.  CPU only.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from heat_pipe_solver_b200 import generate_composite_mesh, material_properties as mp, params
from oracle import fv_oracle as fo


def face_list_system(mesh, cfg, T, T_old, dt, h=5.0, Gamma_at=None):
    """(L, b) of one sweep, literal reference mode: Gamma and the Robin denominator's k are snapshots at `Gamma_at`
    (the ambient field, main.py:133,145), everything else is evaluated at the iterate T (flat, FiPy cell order)."""
    dims, par, con = cfg['dimensions'], cfg['parameters'], cfg['constants']
    na, steel, wick, vc = mp.get_sodium_properties(), mp.get_steel_properties(), mp.get_wick_properties(), mp.get_vc_properties()
    r_c, z_c = np.asarray(mesh.cellCenters)
    r_f, z_f = np.asarray(mesh.faceCenters)
    L_e, L_a = dims['L_e'], dims['L_a']
    L_total = L_e + L_a + dims['L_c']
    first = np.asarray(mesh.faceCellIDs[0])
    second = np.ma.filled(mesh.faceCellIDs[1], -1)
    interior = np.asarray(mesh.interiorFaces)
    ratio, dist, area, vol = mesh._faceToCellDistanceRatio, mesh._cellDistances, mesh._faceAreas, mesh.cellVolumes

    def face_k(Tc):
        Tf = Tc[first].copy()                                            # exterior faces: value of the adjacent cell
        Tf[interior] += ratio[interior] * (Tc[second[interior]] - Tc[first[interior]])
        in_vc = r_f <= dims['R_vc']
        k = np.zeros_like(Tf)
        sel = in_vc & ((z_f < L_e) | (z_f > L_e + L_a))
        k[sel] += vc['thermal_conductivity'](Tf[sel], mesh, 'evap_cond', na, dims, par, con)
        sel = in_vc & (z_f > L_e) & (z_f < L_e + L_a)
        k[sel] += vc['thermal_conductivity'](Tf[sel], mesh, 'adiabatic', na, dims, par, con)
        sel = (dims['R_vc'] < r_f) & (r_f < dims['R_wick'])
        k[sel] += wick['thermal_conductivity'](Tf[sel], na, steel, par)
        sel = r_f > dims['R_wick']
        k[sel] += steel['thermal_conductivity'](Tf[sel])
        return k

    k_now = face_k(T)
    k_snap = face_k(np.full_like(T, par['T_amb']) if Gamma_at is None else Gamma_at)

    in_vc = r_c < dims['R_vc']
    in_wick = (dims['R_vc'] < r_c) & (r_c < dims['R_wick'])
    in_wall = r_c > dims['R_wick']
    vc_live = in_vc & ((z_c < L_e) | (z_c > L_e + L_a) | ((z_c > L_e) & (z_c < L_e + L_a)))
    cp = np.zeros_like(T)
    rho = np.zeros_like(T)
    cp[vc_live] = vc['specific_heat'](T[vc_live])
    rho[vc_live] = vc['density'](T[vc_live])
    cp[in_wick] = wick['specific_heat'](T[in_wick], na, steel, par)
    rho[in_wick] = wick['density'](T[in_wick], na, steel, par)
    cp[in_wall] = steel['specific_heat'](T[in_wall])
    sel = in_wall & (z_c < L_e + L_a)
    rho[sel] += steel['density_evap_adia'](T[sel])
    sel = in_wall & (z_c > L_e + L_a)
    rho[sel] += steel['density_cond'](T[sel])

    n = mesh.numberOfCells
    L = sp.lil_matrix((n, n))
    b = np.zeros(n)
    cap = cp * rho * vol / dt
    L.setdiag(cap)
    b += cap * T_old

    outer = np.asarray(mesh.facesRight)
    evap = outer & (z_f < dims['L_input_right']) & (z_f > dims['L_input_left'])
    cond = outer & (z_f > L_e + L_a) & (z_f < L_total)
    for f in range(mesh.numberOfFaces):
        a_, b_ = first[f], second[f]
        if interior[f]:
            c = k_snap[f] * area[f] / dist[f]                            # Gamma A / d, Gamma frozen at the snapshot
            L[a_, a_] += c
            L[b_, b_] += c
            L[a_, b_] -= c
            L[b_, a_] -= c
            continue
        if cond[f]:
            # Robin: k n / (h dPf.n + k0); dPf = ratio * (face - cell) = the cell-centre-to-face distance on an exterior face
            coeff = k_now[f] / (h * ratio[f] * dist[f] + k_snap[f]) * area[f]
            L[a_, a_] += coeff * h
            b[a_] += coeff * h * par['T_amb']
        if evap[f]:
            b[a_] += par['Q_input_flux'] / k_now[f] * area[f]
    return L.tocsr(), b


def hot_field(mesh):
    r, z = np.asarray(mesh.cellCenters)
    zz, rr = z / z.max(), r / r.max()
    return 290.0 + 500.0 * np.exp(-((zz - 0.1) / 0.35) ** 2) * (0.8 + 0.2 * rr) + 3.0 * np.sin(40 * zz) * rr


@pytest.mark.parametrize('shape', ['native', 's40x24'])
@pytest.mark.parametrize('state', ['ambient', 'hot'])
def test_structured_oracle_equals_face_list_assembly(shape, state):
    cfg = params.load_config()
    mparams = dict(cfg['mesh']) if shape == 'native' else fo.synthetic_mesh_params(40, 4, 8, 12)
    if shape == 'native':
        mparams.update(nx_wall=60, nx_wick=60, nx_vc=60)                   # 60 x 9: the python face loop stays fast
    mesh, _ = generate_composite_mesh(dict(mparams), cfg['dimensions'])
    grid = fo.build_grid(dict(mparams), cfg['dimensions'])
    T = np.full(mesh.numberOfCells, cfg['parameters']['T_amb']) if state == 'ambient' else hot_field(mesh)
    T_old = T - 1.25
    o = fo.Oracle(grid, fo.case_from_config(), T0=T.reshape(grid.nz, grid.nr), has_old=True)
    o.T_old = T_old.reshape(grid.nz, grid.nr).copy()
    S = o.assemble(0.1)
    L, b = face_list_system(mesh, cfg, T, T_old, 0.1)
    A = S.to_csr().tocsr()
    scale = abs(A).max()
    assert abs(A - L).max() <= 1e-12 * scale
    # row-wise: the vapor-core rows are ~1e-12 of the wall rows, check them relative to their own diagonal
    d = A.diagonal()
    assert np.max(np.abs((A - L).diagonal()) / d) <= 1e-11
    assert np.max(np.abs(S.rhs.ravel() - b) / np.maximum(np.abs(b), 1e-300)) <= 1e-11
    # and the off-diagonal pattern is exactly FiPy's: interior faces only
    assert (A != 0).sum() == (L != 0).sum()


@pytest.mark.parametrize('has_old', [False, True])
def test_sweep_sequence_equals_face_list_solves(has_old):
    """`res = eq.sweep(var=T, dt=dt)` (main.py:202) and the `T.updateOld()` + re-sweep variant (main.py:267-271): the oracle's
    sweeps against direct solves of the face-list system -- pre-solve residual and field, step after step."""
    import scipy.sparse.linalg as spla
    cfg = params.load_config()
    mparams = dict(cfg['mesh'])
    mparams.update(nx_wall=40, nx_wick=40, nx_vc=40)
    mesh, _ = generate_composite_mesh(dict(mparams), cfg['dimensions'])
    grid = fo.build_grid(dict(mparams), cfg['dimensions'])
    o = fo.Oracle(grid, fo.case_from_config(), has_old=has_old)
    T = np.full(mesh.numberOfCells, cfg['parameters']['T_amb'])
    T_old = T.copy()
    dt = 2.0
    for step in range(3):
        if has_old:
            o.update_old()
            T_old = T.copy()
        for _ in range(2 if has_old else 1):
            L, b = face_list_system(mesh, cfg, T, T_old if has_old else T, dt)
            res = np.linalg.norm(L @ T - b)
            T = spla.spsolve(L.tocsc(), b)
            res_o = o.sweep(dt, solver='banded')
            assert abs(res - res_o) <= 1e-9 * max(res_o, 1e-30) + 1e-12
            assert np.max(np.abs(T - o.T.ravel())) <= 1e-9
    assert T.max() > cfg['parameters']['T_amb'] + 0.5            # the evaporator did heat the wall
