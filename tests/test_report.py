"""Host-side result files and the comparison table with measured wall temperatures (SURVEY 8f-3;
reference: src/2d/main.py:364-405).  Synthetic result dicts: nothing here needs a GPU."""
import numpy as np
import pytest

from heat_pipe_solver_b200 import generate_composite_mesh, params, report


@pytest.fixture()
def result():
    cfg = params.load_config()
    mesh, _ = generate_composite_mesh(dict(cfg['mesh']), cfg['dimensions'])
    z = mesh.y_c
    times = np.array([5.0, 1038.0, 1998.0])
    profiles = np.stack([290.0 + 0.01 * t * np.exp(-z / 0.3) for t in times])
    return dict(T=np.full(mesh.shape, 290.0), times=times, profiles=profiles, residuals=np.array([1e-3, 1e-4, 1e-5]),
                n_steps=3, iterations=42, sweeps=3, elapsed=0.5, mesh=mesh, kernel_launches=99,
                rho_top_wall=[np.ones(mesh.nz)] * 3, sweeps_per_step=np.array([]))


def test_axial_stations_are_the_top_wall_cell_centres(result):
    z = report.axial_stations(result)
    mesh = result['mesh']
    assert z.shape == (mesh.nz,)
    assert np.array_equal(z, np.asarray(mesh.cellCenters[1]).reshape(mesh.shape)[:, -1])


def test_save_and_load_round_trip(result, tmp_path):
    path = report.save_results(result, tmp_path / 'run')
    assert path.endswith('.npz')
    back = report.load_results(path)
    assert np.array_equal(back['profiles'], result['profiles'])
    assert np.array_equal(back['times'], result['times'])
    assert back['iterations'] == 42 and back['n_steps'] == 3
    assert np.array_equal(back['z'], report.axial_stations(result))
    assert 'mesh' not in back


def test_profiles_csv(result, tmp_path):
    p = report.write_profiles_csv(result, tmp_path / 'profiles.csv')
    rows = open(p).read().strip().splitlines()
    assert rows[0].startswith('z [m],T(t=5 s) [K],T(t=1038 s) [K]')
    assert len(rows) == 1 + result['mesh'].nz
    first = [float(v) for v in rows[1].split(',')]
    assert first[0] == report.axial_stations(result).min()
    assert first[1:] == [float(v) for v in result['profiles'][:, 0]]


def test_comparison_with_measurements(result, tmp_path):
    z = report.axial_stations(result)
    x = np.array([0.011, 0.12, 0.27, 0.41, 0.9])
    exact = np.interp(x, z, result['profiles'][1])
    f1 = tmp_path / 'm1038.csv'
    f1.write_text('x,y\n' + '\n'.join(f'{a}, {b + 2.0}' for a, b in zip(x, exact)) + '\n')      # "x, y" with a blank, like the reference's files
    table = report.compare_with_experiment(result, {1038: str(f1), 1998: str(tmp_path / 'missing.csv'), 2958: str(f1)}, time_tolerance=0.1)
    assert [row['t_exp'] for row in table] == [1038.0, 1998.0]            # 5 s matches nothing, 2958 s was never captured
    ok, missing = table
    assert ok['n_points'] == 5
    assert ok['rms_K'] == pytest.approx(2.0, abs=1e-9) and ok['max_abs_K'] == pytest.approx(2.0, abs=1e-9)
    assert ok['mean_K'] == pytest.approx(-2.0, abs=1e-9)
    assert 'file not found' in missing['error']
    text = report.format_comparison(table)
    assert 'rms [K]' in text and 'file not found' in text and len(text.splitlines()) == 3


def test_measurement_file_without_xy_columns(tmp_path):
    f = tmp_path / 'bad.csv'
    f.write_text('a,b\n1,2\n')
    with pytest.raises(KeyError):
        report.read_xy_csv(f)
