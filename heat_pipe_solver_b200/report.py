"""Results on disk and the comparison with measured wall temperatures -- the numbers behind the reference's figures
(`/root/reference/src/2d/main.py:364-405`: top-wall axial profiles with the Faghri thermocouple read-offs overlaid),
without matplotlib.  Host-side only: works on the dict returned by `heat_pipe_solver_b200.main.run`.
"""
from __future__ import annotations

import csv

import numpy as np

_ARRAY_KEYS = ('T', 'times', 'profiles', 'residuals', 'rho_top_wall', 'rho_wick', 'rho_vc', 'cp_top_wall', 'cp_wick', 'cp_vc',
               'k_top_wall', 'k_wick', 'k_vc', 'sweeps_per_step')
_SCALAR_KEYS = ('n_steps', 'iterations', 'sweeps', 'elapsed', 'kernel_launches')


def axial_stations(result):
    """z of the top-wall cells the profiles are sampled at (`z_cell[top_wall_mask]`, main.py:165-169,365)."""
    mesh = result['mesh']
    if hasattr(mesh, 'y_c'):                     # StructuredGrid2D: y is the axial direction (dz -> y)
        return np.asarray(mesh.y_c, dtype=np.float64)
    return np.asarray(mesh.cellCenters[1], dtype=np.float64).reshape(mesh.shape)[:, -1]


def save_results(result, path):
    """Write the run to `<path>.npz` (every array of the result dict + the axial stations + the scalars)."""
    out = {k: np.asarray(result[k]) for k in _ARRAY_KEYS if k in result and len(np.shape(result[k])) > 0}
    out.update({k: np.asarray(result[k]) for k in _SCALAR_KEYS if k in result})
    if 'mesh' in result:
        out['z'] = axial_stations(result)
    target = path if str(path).endswith('.npz') else str(path) + '.npz'
    np.savez_compressed(target, **out)
    return target


def load_results(path):
    with np.load(path, allow_pickle=False) as f:
        return {k: (f[k].item() if f[k].ndim == 0 else f[k]) for k in f.files}


def write_profiles_csv(result, path, z=None):
    """One row per axial station, one column per captured time: the table behind `plots/top_wall_axial_profiles.png`
    (main.py:381-383,408)."""
    z = axial_stations(result) if z is None else np.asarray(z)
    prof = np.asarray(result['profiles'])
    times = np.asarray(result['times'])
    order = np.argsort(z)                                               # main.py:368-369
    with open(path, 'w', newline='') as f:
        w = csv.writer(f)
        w.writerow(['z [m]'] + [f'T(t={t:g} s) [K]' for t in times])
        for j in order:
            w.writerow([repr(float(z[j]))] + [repr(float(prof[k, j])) for k in range(len(times))])
    return path


def read_xy_csv(path):
    """The reference's digitised measurement files: header `x,y`, x = axial position [m], y = temperature [K]
    (`Faghri_Data/*.csv`, read with pandas at main.py:388-390).  Raises KeyError if a column is missing (main.py:396)."""
    with open(path, newline='') as f:
        rows = list(csv.DictReader(f, skipinitialspace=True))
    if rows and ('x' not in rows[0] or 'y' not in rows[0]):
        raise KeyError(f"{path}: columns 'x' and 'y' expected")
    return (np.array([float(r['x']) for r in rows], dtype=np.float64),
            np.array([float(r['y']) for r in rows], dtype=np.float64))


def compare_with_experiment(result, experiment_files, time_tolerance, z=None):
    """For every captured profile whose time lies within `time_tolerance` of a key of `experiment_files`
    ({time_s: csv_path}; the reference uses {1038, 1998, 2958} s and tolerance = dt, main.py:372-386), interpolate the
    simulated top-wall profile at the measured positions and report the differences.

    -> list of dicts: t_sim, t_exp, n_points, rms_K, max_abs_K, mean_K (simulation - measurement), x, T_exp, T_sim.
    The first matching file per snapshot wins (main.py:401-403); missing files are skipped with a warning entry
    (`error` key), like the reference's `except FileNotFoundError` branch."""
    z = axial_stations(result) if z is None else np.asarray(z, dtype=np.float64)
    order = np.argsort(z)
    zs = z[order]
    table = []
    for prof, t_sim in zip(np.asarray(result['profiles']), np.asarray(result['times'])):
        for t_exp, path in experiment_files.items():
            if abs(float(t_sim) - float(t_exp)) < time_tolerance:
                try:
                    x, T_exp = read_xy_csv(path)
                except FileNotFoundError:
                    table.append(dict(t_sim=float(t_sim), t_exp=float(t_exp), error=f'file not found: {path}'))
                    break
                T_sim = np.interp(x, zs, np.asarray(prof)[order])
                diff = T_sim - T_exp
                table.append(dict(t_sim=float(t_sim), t_exp=float(t_exp), n_points=int(x.size),
                                  rms_K=float(np.sqrt(np.mean(diff ** 2))), max_abs_K=float(np.max(np.abs(diff))),
                                  mean_K=float(np.mean(diff)), x=x, T_exp=T_exp, T_sim=T_sim))
                break
    return table


def format_comparison(table):
    lines = [f"{'t_sim [s]':>10s} {'t_exp [s]':>10s} {'points':>6s} {'rms [K]':>9s} {'max [K]':>9s} {'mean [K]':>9s}"]
    for row in table:
        if 'error' in row:
            lines.append(f"{row['t_sim']:10.2f} {row['t_exp']:10.2f}   {row['error']}")
        else:
            lines.append(f"{row['t_sim']:10.2f} {row['t_exp']:10.2f} {row['n_points']:6d} {row['rms_K']:9.3f} "
                         f"{row['max_abs_K']:9.3f} {row['mean_K']:9.3f}")
    return '\n'.join(lines)
