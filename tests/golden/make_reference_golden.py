#!/usr/bin/env python
"""Generate golden vectors from the REFERENCE's own Python modules.

Run in the build container only (it reads /root/reference, which does not exist on
the GPU box):   python tests/golden/make_reference_golden.py

FiPy is not installable here, so the reference modules are imported with a small
recording stub of `fipy`:
  * `material_properties.py` only needs `fipy.tools.numerix` (a numpy alias) for its
    live code paths -> every property law is evaluated by the reference's OWN code;
  * `mesh.py` calls `CylindricalGrid2D(dz=..., dr=..., nz=..., nr=...)` -> the stub
    records the spacing arrays the reference hands to FiPy and lays out cell centres
    in FiPy's documented order (dr -> x fastest, dz -> y), so the vertex recipe
    (`mesh.py:91-106`) and the side effects (`mesh.py:86-88`) are the reference's own;
  * `params.py` is stdlib-only and imported verbatim.

Outputs (committed): tests/golden/reference_golden.npz, tests/golden/reference_params.json
"""
import importlib
import json
import os
import sys
import types

import numpy as np

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))


class _RecordingGrid:
    def __init__(self, dz=None, dr=None, nz=None, nr=None, **kw):
        self.dz = np.array(dz, dtype=np.float64)
        self.dr = np.array(dr, dtype=np.float64)
        self.nz, self.nr = int(nz), int(nr)
        self.numberOfCells = self.nz * self.nr
        r_v = np.concatenate(([0.0], np.cumsum(self.dr)))
        z_v = np.concatenate(([0.0], np.cumsum(self.dz)))
        rc = 0.5 * (r_v[:-1] + r_v[1:])
        zc = 0.5 * (z_v[:-1] + z_v[1:])
        R, Z = np.meshgrid(rc, zc)            # shape (nz, nr): r fastest when ravelled
        self.cellCenters = np.array([R.ravel(), Z.ravel()])


class _CellVariable:
    def __init__(self, name=None, mesh=None, value=None, **kw):
        self.name, self.mesh, self.value = name, mesh, np.array(value)


def install_stub():
    fipy = types.ModuleType('fipy')
    fipy.CellVariable = _CellVariable
    fipy.Grid2D = _RecordingGrid
    fipy.Viewer = object
    tools = types.ModuleType('fipy.tools')
    tools.numerix = np
    meshes = types.ModuleType('fipy.meshes')
    meshes.CylindricalGrid2D = _RecordingGrid
    nu = types.ModuleType('fipy.meshes.nonUniformGrid2D')
    nu.NonUniformGrid2D = _RecordingGrid
    fipy.tools, fipy.meshes = tools, meshes
    sys.modules.update({'fipy': fipy, 'fipy.tools': tools, 'fipy.meshes': meshes,
                        'fipy.meshes.nonUniformGrid2D': nu})


def main():
    install_stub()
    sys.path.insert(0, os.path.join(REF, 'src', '2d'))
    params = importlib.import_module('params')
    mp = importlib.import_module('material_properties')
    mesh_mod = importlib.import_module('mesh')

    out = {}
    flat = params.get_all_params()
    groups = {g: params.get_param_group(g) for g in ('parameters', 'dimensions', 'mesh', 'constants', 'simulation')}
    with open(os.path.join(HERE, 'reference_params.json'), 'w') as f:
        json.dump({'flat': flat, 'groups': groups, 'unknown_group': params.get_param_group('nope'),
                   'none_group': params.get_param_group(None)}, f, indent=1, sort_keys=True)

    prm, dims, cons = groups['parameters'], groups['dimensions'], groups['constants']
    Tm, dT = prm['T_melting_sodium'], prm['delta_T_sodium']
    rng = np.random.default_rng(20251019)
    T = np.concatenate([
        np.linspace(280.0, 1500.0, 489),
        np.array([290.0, 370.98, 500.0, 900.0, 1200.0]),
        np.array([Tm - dT, Tm + dT, np.nextafter(Tm - dT, 0), np.nextafter(Tm + dT, 1e9), Tm]),
        rng.uniform(285.0, 1400.0, 512),
    ])
    out['T'] = T
    na, st, wk, vc = (mp.get_sodium_properties(), mp.get_steel_properties(),
                      mp.get_wick_properties(), mp.get_vc_properties())
    for name, fn in na.items():
        out['na_' + name] = fn(T)
    for name, fn in st.items():
        out['steel_' + name] = fn(T)
    for name, fn in wk.items():
        out['wick_' + name] = fn(T, na, st, prm)
    out['vc_k_evap_cond'] = vc['thermal_conductivity'](T, None, 'evap_cond', na, dims, prm, cons)
    out['vc_k_adiabatic'] = vc['thermal_conductivity'](T, None, 'adiabatic', na, dims, prm, cons)
    out['vc_specific_heat'] = vc['specific_heat'](T)
    out['vc_density'] = vc['density'](T)
    out['na_k_interp'] = mp.get_k_Na_i(T, prm, na['thermal_conductivity_liquid'](T), na['thermal_conductivity_solid'](T))
    out['na_cp_interp'] = mp.get_c_p_Na_i(T, na, prm, na['specific_heat_liquid'](T), na['specific_heat_solid'](T))
    out['na_rho_interp'] = mp.get_rho_Na_i(T, prm, na['density_liquid'](T), na['density_solid'](T))

    # meshes: native + the synthetic shapes of BASELINE.json (radial spacings only need 1-D arrays)
    shapes = {'native': None, 's64x32': (64, 4, 8, 20), 's4096x1024': (4096, 128, 256, 640),
              's16384x4096': (16384, 512, 1024, 2560)}
    for tag, shp in shapes.items():
        mparams = dict(groups['mesh'])
        if shp is not None:
            nz, a, b, c = shp
            mparams.update(nx_wall=nz, nr_vc=a, nr_wick=b, nr_wall=c, nx_vc=7, nx_wick=9)
        m, ct = mesh_mod.generate_composite_mesh(mparams, dims)
        out[f'mesh_{tag}_dr'] = m.dr
        out[f'mesh_{tag}_dz'] = m.dz
        out[f'mesh_{tag}_shape'] = np.array([m.nz, m.nr])
        out[f'mesh_{tag}_nx_side_effect'] = np.array([mparams['nx_vc'], mparams['nx_wick']])
        if m.numberOfCells <= 5000:
            out[f'mesh_{tag}_cell_types'] = np.array(ct.value)
    np.savez_compressed(os.path.join(HERE, 'reference_golden.npz'), **out)
    print('wrote', len(out), 'arrays;', {k: v.shape for k, v in out.items() if k.startswith('mesh')})


if __name__ == '__main__':
    main()
