"""Region / boundary classification of the composite mesh with the reference's exact inequalities.

The reference builds boolean masks over all cells and faces (main.py:64-72 cells, :86-90 faces,
:123-124 boundary faces).  On a tensor-product grid they are separable, so they are kept as 1-D
codes: a radial band per column and a bit set per axial station.  Which side of a band interface a
face falls on is decided by floating-point luck in the reference (`r_face <= R_vc`,
`r_face < R_wick` on running-sum vertices), hence the masks are evaluated here on the host with the
very same float operations and handed to the kernels as codes.
"""
from __future__ import annotations

import numpy as np

BAND_VC, BAND_WICK, BAND_WALL, BAND_NONE = 0, 1, 2, 3
ZF_VC_EVAPCOND, ZF_VC_ADIAB, ZF_BEFORE_COND, ZF_IN_COND, ZF_BC_EVAP, ZF_BC_COND = 1, 2, 4, 8, 16, 32


def cell_band(r, dims):
    """main.py:64,67,70 -- strict inequalities; a centre exactly on an interface belongs to no band."""
    band = np.full(np.shape(r), BAND_NONE, dtype=np.int8)
    band[r < dims['R_vc']] = BAND_VC
    band[(dims['R_vc'] < r) & (r < dims['R_wick'])] = BAND_WICK
    band[r > dims['R_wick']] = BAND_WALL
    return band


def face_band(r, dims):
    """main.py:86,89,90 -- `<=` on the vapor-core side only."""
    band = np.full(np.shape(r), BAND_NONE, dtype=np.int8)
    band[r <= dims['R_vc']] = BAND_VC
    band[(dims['R_vc'] < r) & (r < dims['R_wick'])] = BAND_WICK
    band[r > dims['R_wick']] = BAND_WALL
    return band


def z_flags(z, dims):
    L_e, L_a = dims['L_e'], dims['L_a']
    L_total = dims['L_e'] + dims['L_a'] + dims['L_c']                       # main.py:25
    f = np.zeros(np.shape(z), dtype=np.uint8)
    f[(z < L_e) | (z > (L_e + L_a))] |= ZF_VC_EVAPCOND                      # main.py:65,87
    f[(z > L_e) & (z < L_e + L_a)] |= ZF_VC_ADIAB                           # main.py:66,88
    f[z < L_e + L_a] |= ZF_BEFORE_COND                                      # main.py:68,71
    f[z > (L_e + L_a)] |= ZF_IN_COND                                        # main.py:69,72
    f[(z < dims['L_input_right']) & (z > dims['L_input_left'])] |= ZF_BC_EVAP   # main.py:123
    f[(z > (L_e + L_a)) & (z < L_total)] |= ZF_BC_COND                      # main.py:124
    return f


class Regions:
    """1-D region codes of a grid (global, not ghost-inclusive)."""

    def __init__(self, mesh, dims):
        self.cell_band = cell_band(mesh.r_c, dims)
        self.rface_band = face_band(mesh.r_v[1:], dims)      # face at r_v[i+1]; the last one is the outer surface
        self.zface_band = face_band(mesh.r_c, dims)          # axial faces sit at r_c
        self.zc_flags = z_flags(mesh.z_c, dims)              # cells, radial faces, outer-surface faces
        self.zv_flags = z_flags(mesh.z_v, dims)              # [nz+1]: axial faces at every vertex z_v[j]

    # the reference's 2-D masks, for API parity / tests (main.py:64-72)
    def cell_mask(self, name, mesh):
        band = {'vc': BAND_VC, 'wick': BAND_WICK, 'wall': BAND_WALL}[name]
        return np.repeat((self.cell_band == band)[None, :], mesh.nz, axis=0).ravel()
