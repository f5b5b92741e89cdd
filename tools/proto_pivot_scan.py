"""Prototype (numpy, CPU): the pivots of the radial-line tridiagonal, w_i = d_i - c_{i-1}^2 / w_{i-1}, as a parallel scan.

w_i = p_i / q_i with (p_i, q_i)^T = M_i (p_{i-1}, q_{i-1})^T, M_i = [[d_i, -c_{i-1}^2], [1, 0]]: a product of 2x2 matrices,
i.e. an associative scan.  The entries of the partial products are continuants (leading minors) and overflow quickly, so
every combine step renormalises by the largest entry (the ratio p/q is scale-free).  This script measures the accuracy of
that scan (Hillis-Steele over the whole row, the shape a warp-shuffle scan has) against the sequential recurrence on the
rows of the real operator.  It answers "is k_factor_rline safe to parallelise this way" before any kernel is written.

    python tools/proto_pivot_scan.py [nz nr]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import fv_oracle as fo  # noqa: E402


def pivots_sequential(d, c):
    w = np.empty_like(d)
    w[:, 0] = d[:, 0]
    for i in range(1, d.shape[1]):
        w[:, i] = d[:, i] - c[:, i - 1] ** 2 / w[:, i - 1]
    return w


def pivots_scan(d, c):
    rows, n = d.shape
    M = np.zeros((rows, n, 2, 2))
    M[:, :, 0, 0] = d
    M[:, 1:, 0, 1] = -c[:, :-1] ** 2
    M[:, :, 1, 0] = 1.0
    M[:, 0, 1, 0] = 0.0
    M[:, 0, 1, 1] = 1.0                     # first element: (p, q) = (d_0, 1)
    M[:, 0, 0, 1] = 0.0
    step = 1
    while step < n:
        left = M[:, :-step]                 # prefix ending at i - step
        right = M[:, step:]                 # block (i - step, i]
        prod = np.einsum('rnij,rnjk->rnik', right, left)
        prod /= np.max(np.abs(prod), axis=(2, 3), keepdims=True)
        M = np.concatenate([M[:, :step], prod], axis=1)
        step *= 2
    # apply to the start vector (p, q) = (1, 0) -> first column... the prefix product applied to e = (1, 0)^T scaled
    p = M[:, :, 0, 0] * 1.0 + M[:, :, 0, 1] * 1.0
    q = M[:, :, 1, 0] * 1.0 + M[:, :, 1, 1] * 1.0
    return p / q


if __name__ == '__main__':
    nz, nr = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 1024)
    cfg = fo.load_config()
    grid = fo.build_grid(bench.mesh_params_for(nz, nr) if nr in bench.RADIAL_SPLIT else dict(cfg['mesh']), cfg['dimensions'])
    for label, T0 in (('ambient', None), ('hot', 290.0 + 500.0 * np.exp(-((grid.z_c[:, None] / grid.z_v[-1] - 0.1) / 0.35) ** 2) * np.ones((1, grid.nr)))):
        o = fo.Oracle(grid, fo.case_from_config(), T0=T0)
        S = o.assemble(0.1)
        d, c = S.diag, S.cR
        ws = pivots_sequential(d, c)
        wp = pivots_scan(d, c)
        rel = np.abs(wp / ws - 1.0)
        print(f"{label:8s} {grid.nz}x{grid.nr}: max rel. error of the scanned pivots {rel.max():.3e} (median {np.median(rel):.1e}), "
              f"min pivot/diag {np.min(ws / d):.3e}")
