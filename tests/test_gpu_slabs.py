"""Axial-slab decomposition over >= 2 GPUs (NCCL halo rows + combined CG scalars) against the single-domain
oracle.  Skipped on a 1-GPU box; the host-side partition logic is covered on CPU in test_host_logic.py."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize('comm,loop_mode,precond', [('nccl', 'host', 'rline'), ('p2p', 'host', 'rline'), ('p2p', 'graph', 'rline'),
                                                    ('p2p', 'graph', 'mgline'), ('nccl', 'host', 'mgline')])
@pytest.mark.parametrize('world,shape', [(2, (64, 4, 8, 20)), (2, (50, 32, 64, 160)), (4, (96, 12, 24, 64))])
def test_slabs_match_single_domain_oracle(cfg, tmp_path, world, shape, comm, loop_mode, precond):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    from oracle import fv_oracle as fo
    out = str(tmp_path / 'slab.npy')
    port = _free_port()
    n_steps, sweeps = 2, 3
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, 'tests', 'slab_worker.py'), str(r), str(world), str(port),
                               ','.join(map(str, shape)), str(n_steps), str(sweeps), out, comm, loop_mode, precond]) for r in range(world)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    T = np.load(out)
    grid = fo.build_grid(fo.synthetic_mesh_params(*shape), cfg['dimensions'])
    z = grid.z_c[:, None] / grid.z_v[-1]
    T0 = 290.0 + 250.0 * np.exp(-((z - 0.45) / 0.3) ** 2) * np.ones((1, grid.nr))
    o = fo.Oracle(grid, fo.case_from_config(), T0=T0, has_old=True)
    for _ in range(n_steps):
        o.update_old()
        for _ in range(sweeps):
            o.sweep(0.1, solver='banded' if grid.nr <= 160 else 'splu')
    assert np.max(np.abs(T - o.T)) <= 1e-6 * n_steps * sweeps
