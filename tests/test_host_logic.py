"""Host-side logic that needs no GPU: loop bookkeeping, slab layouts, member sharding, and the slab partition
under a real 2-process gloo group."""
import os
import socket

import numpy as np
import pytest

from heat_pipe_solver_b200 import generate_composite_mesh
from heat_pipe_solver_b200.distributed import neighbours, slab_rows
from heat_pipe_solver_b200.ensemble import _StackedLayout, shard_members
from heat_pipe_solver_b200.main import measurement_schedule
from heat_pipe_solver_b200.regions import Regions
from heat_pipe_solver_b200.solver import SlabLayout


def test_measurement_schedule_matches_reference_loop():
    times, caps = measurement_schedule(101.0, 0.1, list(range(5, 101, 5)))       # main.py:119-120,158-159
    assert len(times) == 1010
    assert [t for _, t in caps][:3] == [5.0, 10.0, 15.0] and len(caps) == 20
    assert all(times[s] == t for s, t in caps)
    # a capture happens on the first loop time >= the scheduled time, once
    times, caps = measurement_schedule(12.0, 0.7, [1, 2, 3, 10])
    assert [round(t, 10) for _, t in caps] == [1.4, 2.1, 3.5, 10.5]
    assert measurement_schedule(3.0, 1.0, [])[1] == []


def test_slab_rows_cover_and_balance():
    for nz, world in ((16384, 8), (4096, 3), (500, 7), (5, 5)):
        rows = [slab_rows(nz, r, world) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == nz
        assert all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        sizes = [b - a for a, b in rows]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab_rows(3, 0, 4)
    assert neighbours(0, 4) == (None, 1) and neighbours(3, 4) == (2, None) and neighbours(0, 1) == (None, None)


def test_slab_layout_is_a_window_of_the_global_arrays(cfg):
    mesh, _ = generate_composite_mesh(dict(cfg['mesh']), cfg['dimensions'])
    reg = Regions(mesh, cfg['dimensions'])
    lines = dict(q=np.arange(500.0), h=np.full(500, 5.0), Tamb=np.full(500, 290.0), eps=np.zeros(500))
    whole = SlabLayout(mesh, reg, 0, 500, lines)
    assert whole.z_v.size == 503 and np.array_equal(whole.z_v[1:-1], mesh.z_v)
    assert whole.axial_open[0] == 0 and whole.axial_open[500] == 0 and whole.axial_open[1:500].all()
    a, b = SlabLayout(mesh, reg, 0, 250, lines), SlabLayout(mesh, reg, 250, 500, lines)
    assert a.axial_open[250] == 1 and b.axial_open[0] == 1 and b.axial_open[250] == 0
    # ghost rows are the neighbour's rows
    assert np.array_equal(a.z_v[-3:], mesh.z_v[249:252]) and np.array_equal(b.z_v[:3], mesh.z_v[249:252])
    assert a.zc_flags[-1] == reg.zc_flags[250] and b.zc_flags[0] == reg.zc_flags[249]
    assert a.lines['q'][-1] == 250.0 and b.lines['q'][0] == 249.0 and b.lines['q'][-1] == 499.0
    # the face between the two slabs carries the same flags on both sides
    assert a.zv_flags[250] == b.zv_flags[0] == reg.zv_flags[250]
    ab = np.zeros(499, dtype=bool)
    ab[99] = True
    cut = SlabLayout(mesh, reg, 0, 500, lines, axial_break=ab)
    assert cut.axial_open[100] == 0 and cut.axial_open[99] == 1 and cut.axial_open[101] == 1


def test_member_sharding_and_stacked_layout(cfg):
    parts = [shard_members(1024, r, 8) for r in range(8)]
    assert parts[0] == (0, 128) and parts[-1] == (896, 1024)
    parts = [shard_members(10, r, 4) for r in range(4)]
    assert [b - a for a, b in parts] == [3, 3, 2, 2] and parts[-1][1] == 10
    mesh, _ = generate_composite_mesh(dict(cfg['mesh']), cfg['dimensions'])
    reg = Regions(mesh, cfg['dimensions'])
    lay = _StackedLayout(mesh, reg, 3, dict(q=[1.0, 2.0, 3.0], h=[5.0] * 3, Tamb=[290.0] * 3, eps=[0.0] * 3))
    assert lay.nz == 1500 and lay.z_v.size == 1503 and np.all(np.diff(lay.z_v) > 0)
    assert list(np.nonzero(lay.axial_open == 0)[0]) == [0, 500, 1000, 1500, 1501]
    assert np.array_equal(lay.zc_flags[1:501], reg.zc_flags) and np.array_equal(lay.zc_flags[501:1001], reg.zc_flags)
    assert lay.lines['q'][1] == 1.0 and lay.lines['q'][501] == 2.0 and lay.lines['q'][1500] == 3.0


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _gloo_worker(rank, world, port, nz, nr, out):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from heat_pipe_solver_b200.distributed import gather_field, slab_rows
    j0, j1 = slab_rows(nz, rank, world)
    full = torch.arange(nz * nr, dtype=torch.float64).reshape(nz, nr)
    got = gather_field(full[j0:j1].clone())
    # halo exchange pattern of the slab solver: my first row goes down, my last row goes up
    lo, hi = (rank - 1 if rank > 0 else None), (rank + 1 if rank < world - 1 else None)
    recv_lo, recv_hi = torch.zeros(nr, dtype=torch.float64), torch.zeros(nr, dtype=torch.float64)
    reqs = []
    if lo is not None:
        reqs += [dist.isend(full[j0].clone(), lo), dist.irecv(recv_lo, lo)]
    if hi is not None:
        reqs += [dist.isend(full[j1 - 1].clone(), hi), dist.irecv(recv_hi, hi)]
    for r in reqs:
        r.wait()
    ok = bool(torch.equal(got, full))
    if lo is not None:
        ok &= bool(torch.equal(recv_lo, full[j0 - 1]))
    if hi is not None:
        ok &= bool(torch.equal(recv_hi, full[j1]))
    # CG scalars: sum / max over ranks
    v = torch.tensor([float(rank + 1), 10.0 * (rank + 1)], dtype=torch.float64)
    dist.all_reduce(v)
    ok &= bool(v[0].item() == world * (world + 1) / 2)
    out[rank] = ok
    dist.destroy_process_group()


def test_slab_partition_under_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    mgr = ctx.Manager()
    out = mgr.dict()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, 11, 4, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(out) == {0: True, 1: True}


def test_aligned_slab_cuts_and_global_precond_resolution():
    """Peer-memory slabs with 'mgline' cut the mesh on multiples of 2^(levels-1) rows (the distributed V-cycle is then the
    single-domain one), and 'auto' / the mgline->rline fallback are decided from the GLOBAL mesh: every rank gets the same
    answer even when the slabs differ by a row (nz=1023, nr=512, world=2 gave mgline on one rank, rline on the other)."""
    from heat_pipe_solver_b200.distributed import mg_level_count, resolve_precond, slab_alignment, slab_rows
    assert [mg_level_count(n) for n in (16, 31, 32, 64, 500, 4096, 16384)] == [1, 1, 2, 3, 6, 6, 6]
    for nz, world in ((16384, 8), (16384, 4), (4096, 2), (1000, 3), (100, 2), (64, 2), (70, 4)):
        a = slab_alignment(nz, world)
        cuts = [slab_rows(nz, r, world, a) for r in range(world)]
        assert cuts[0][0] == 0 and cuts[-1][1] == nz
        assert all(c0[1] == c1[0] for c0, c1 in zip(cuts, cuts[1:]))
        assert all((b - a_) % a == 0 for a_, b in cuts[:-1]) and all(b > a_ for a_, b in cuts)
        sizes = [b - a_ for a_, b in cuts[:-1]]
        assert max(sizes) - min(sizes) <= a
    assert slab_alignment(16384, 8) == 32 and slab_rows(16384, 3, 8, 32) == (6144, 8192)
    assert slab_alignment(64, 2) == 4 and slab_alignment(70, 4) == 4
    # the same decision on every rank, whatever its own slab looks like
    assert resolve_precond('auto', 1023, 512, 2, 'p2p') == 'mgline'
    assert resolve_precond('auto', 1023, 512, 2, 'nccl') == 'mgline'
    assert resolve_precond('auto', 63, 64, 2, 'p2p') == 'rline'
    assert resolve_precond('mgline', 63, 64, 2, 'nccl') == 'rline'          # 31-row slab cannot coarsen: all ranks use rline
    assert resolve_precond('mgline', 40, 64, 2, 'p2p') == 'mgline' and resolve_precond('mgline', 20, 64, 2, 'p2p') == 'rline'
    assert resolve_precond('jacobi', 4096, 1024, 8, 'p2p') == 'jacobi'
    with pytest.raises(ValueError):
        slab_rows(64, 0, 9, 8)


def _coarsen_slab(cR, cZ, dg, cZ_below, zs):
    """numpy mirror of k_mg_coarsen on one slab: cZ[j] couples row j to row j+1 (the last row's entry is the coupling to
    the slab above, 0 at the physical top); cZ_below couples the first row to the slab below."""
    nz = dg.shape[0]
    nzc = (nz + 1) // 2
    cRc, cZc, dgc = np.zeros((nzc,) + cR.shape[1:]), np.zeros((nzc,) + cR.shape[1:]), np.zeros((nzc,) + cR.shape[1:])
    for J in range(nzc):
        g0, two = 2 * J, 2 * J + 1 < nz
        inner = cZ[g0] if two else 0.0
        up = cZ[g0 + 1] if two else cZ[g0]
        dn = cZ[g0 - 1] if g0 > 0 else cZ_below
        cRc[J] = cR[g0] + (cR[g0 + 1] if two else 0.0)
        cZc[J] = zs * up
        dgc[J] = dg[g0] + ((dg[g0 + 1] - 2.0 * inner) if two else 0.0) - (1.0 - zs) * (up + dn)
    return cRc, cZc, dgc, zs * cZ_below


@pytest.mark.parametrize('nz,world', [(256, 2), (512, 4), (1000, 3), (96, 4)])
def test_distributed_hierarchy_equals_the_single_domain_one(cfg, nz, world):
    """With the slab cuts aligned to 2^(levels-1) rows (distributed.slab_alignment) every coarse coefficient of the V-cycle
    is a function of the slab's own rows: coarsening slab by slab -- what every rank's k_mg_coarsen does, no communication --
    gives exactly the rows of the whole-mesh hierarchy (the oracle's), level after level."""
    from heat_pipe_solver_b200.distributed import mg_level_count, slab_alignment, slab_rows
    from oracle import fv_oracle as fo
    grid = fo.build_grid(fo.synthetic_mesh_params(nz, 4, 8, 20), cfg['dimensions'])
    z = grid.z_c[:, None] / grid.z_v[-1]
    T0 = 290.0 + 300.0 * np.exp(-((z - 0.1) / 0.35) ** 2) * np.ones((1, grid.nr))
    S = fo.Oracle(grid, fo.case_from_config(), T0=T0).assemble(0.1)
    align = slab_alignment(nz, world)
    levels = 1 + int(np.log2(align))
    assert levels <= mg_level_count(nz)
    ref = fo.mg_hierarchy(S, max_levels=levels, zscale=0.5)
    assert len(ref) == levels
    cuts = [slab_rows(nz, r, world, align) for r in range(world)]
    slabs = []
    for (a, b) in cuts:
        below = S.cZ[a - 1] if a > 0 else np.zeros(grid.nr)
        slabs.append((S.cR[a:b].copy(), S.cZ[a:b].copy(), S.diag[a:b].copy(), below))
    for l in range(1, levels):
        slabs = [_coarsen_slab(cR, cZ, dg, below, 0.5) for cR, cZ, dg, below in slabs]
        for name, k in (('cR', 0), ('cZ', 1), ('diag', 2)):
            got = np.concatenate([s[k] for s in slabs], axis=0)
            want = getattr(ref[l].S, name)
            assert got.shape == want.shape, (l, name)
            scale = np.abs(want).max()
            assert np.max(np.abs(got - want)) <= 1e-13 * scale, (l, name)
