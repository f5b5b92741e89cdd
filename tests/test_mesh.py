"""mesh.py surface: vertex recipe pinned to the reference's own mesh.py (recorded through a stub FiPy)."""
import numpy as np
import pytest

from heat_pipe_solver_b200 import mesh as M
from heat_pipe_solver_b200.regions import Regions
from oracle import fv_oracle as fo

SHAPES = {'native': None, 's64x32': (64, 4, 8, 20), 's4096x1024': (4096, 128, 256, 640),
          's16384x4096': (16384, 512, 1024, 2560)}


def _mparams(cfg, shp):
    mp = dict(cfg['mesh'])
    if shp is not None:
        nz, a, b, c = shp
        mp.update(nx_wall=nz, nr_vc=a, nr_wick=b, nr_wall=c, nx_vc=7, nx_wick=9)
    return mp


@pytest.mark.parametrize('tag', list(SHAPES))
def test_spacings_bit_exact_vs_reference(cfg, golden, tag):
    mp = _mparams(cfg, SHAPES[tag])
    m, ct = M.generate_composite_mesh(mp, cfg['dimensions'])
    assert np.array_equal(m.dx, golden[f'mesh_{tag}_dr'])
    assert np.array_equal(m.dy, golden[f'mesh_{tag}_dz'])
    assert [m.nz, m.nr] == list(golden[f'mesh_{tag}_shape'])
    # side effect of reference mesh.py:86-88
    assert [mp['nx_vc'], mp['nx_wick']] == list(golden[f'mesh_{tag}_nx_side_effect'])
    if f'mesh_{tag}_cell_types' in golden.files:
        assert np.array_equal(ct, golden[f'mesh_{tag}_cell_types'])
    # the oracle builds the same vertices
    g = fo.build_grid(mp, cfg['dimensions'])
    assert np.array_equal(g.r_v, m.r_v) and np.array_equal(g.z_v, m.z_v)


def test_intended_region_codes(cfg):
    m, _ = M.generate_composite_mesh(dict(cfg['mesh']), cfg['dimensions'])
    codes = m.region_codes.reshape(m.shape)
    assert set(np.unique(codes)) == {0, 1, 10, 11, 20, 21}
    assert (codes[:, 0] < 10).all() and (codes[:, 1:3] // 10 == 1).all() and (codes[:, 3:] // 10 == 2).all()


def test_interface_face_classification_follows_float_recipe(cfg):
    # native mesh: face at R_vc is 'vc', face at R_wick is 'wick' (0.011199999999999998 < R_wick)
    m, _ = M.generate_composite_mesh(dict(cfg['mesh']), cfg['dimensions'])
    reg = Regions(m, cfg['dimensions'])
    assert list(reg.cell_band) == [0, 1, 1, 2, 2, 2, 2, 2, 2]
    assert list(reg.rface_band) == [0, 1, 1, 2, 2, 2, 2, 2, 2]
    assert m.r_v[3] < cfg['dimensions']['R_wick']
    # and the oracle classifies identically on every synthetic shape
    for shp in SHAPES.values():
        mp = _mparams(cfg, shp)
        m, _ = M.generate_composite_mesh(mp, cfg['dimensions'])
        reg = Regions(m, cfg['dimensions'])
        om = fo.build_masks(fo.build_grid(mp, cfg['dimensions']), cfg['dimensions'])
        assert np.array_equal(reg.cell_band, om.cell_band)
        assert np.array_equal(reg.rface_band, om.rface_band)
        assert np.array_equal(reg.zface_band, om.zface_band)
        assert np.array_equal(reg.zc_flags, om.zc_flags.astype(np.uint8))
        assert np.array_equal(reg.zv_flags[1:], om.zv_flags.astype(np.uint8))


def test_fipy_like_attributes_small_grid():
    m = M.StructuredGrid2D(dx=[1.0, 2.0, 1.0], dy=[0.5, 0.5], cylindrical=True)
    assert m.numberOfCells == 6 and m.numberOfFaces == 3 * 3 + 4 * 2
    r, z = m.cellCenters
    assert np.allclose(r, [0.5, 2.0, 3.5] * 2) and np.allclose(z, [0.25] * 3 + [0.75] * 3)
    fr, fz = m.faceCenters
    assert np.allclose(fr[:9], [0.5, 2.0, 3.5] * 3) and np.allclose(fz[:9], [0] * 3 + [0.5] * 3 + [1.0] * 3)
    assert np.allclose(fr[9:], [0, 1, 3, 4] * 2)
    assert m.facesRight.sum() == 2 and m.facesLeft.sum() == 2 and m.facesTop.sum() == 3 and m.facesBottom.sum() == 3
    assert np.allclose(fr[m.facesRight], 4.0)
    # cylindrical volumes / areas per radian
    assert np.allclose(m.cellVolumes, np.array([0.5, 2.0 * 2, 3.5] * 2) * 0.5)
    assert np.allclose(m._faceAreas[m.facesRight], 4.0 * 0.5)
    assert np.allclose(m._faceAreas[m.facesLeft], 0.0)
    # cellFaceIDs: bottom, right, top, left; top-wall detection of main.py:165-169
    cf = m.cellFaceIDs
    assert cf.shape == (4, 6)
    top_cells = np.any(m.facesRight[cf], axis=0)
    assert list(np.nonzero(top_cells)[0]) == [2, 5]
    # faceCellIDs + distance ratio reproduce arithmetic face interpolation between cell centres
    fc = m.faceCellIDs
    interior = m.interiorFaces
    assert (~fc.mask[1][interior]).all() and fc.mask[1][~interior].all()
    ratio = m._faceToCellDistanceRatio
    assert np.allclose(ratio[~interior], 1.0)
    vidx = np.nonzero(interior & (np.arange(m.numberOfFaces) >= 9))[0]
    c1, c2 = fc[0][vidx], fc[1][vidx]
    assert np.allclose(r[c1] + ratio[vidx] * (r[c2] - r[c1]), fr[vidx])
    n = m.faceNormals
    assert np.allclose(n[:, m.facesRight], [[1, 1], [0, 0]]) and np.allclose(n[:, m.facesLeft], [[-1, -1], [0, 0]])


def test_generate_mesh_2d():
    m = M.generate_mesh_2d(0.6, 1.0, 6, 4)
    assert not m.cylindrical and m.numberOfCells == 24
    assert np.allclose(m.cellVolumes, 0.1 * 0.25)
