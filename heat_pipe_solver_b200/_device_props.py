"""Evaluate a material law on a CUDA tensor with the library kernel (`hp_eval_property_raw`)."""
from __future__ import annotations

import ctypes as C

from . import _lib
from .params import load_config
from .solver import PROP_IDS, make_phys

_default_phys = None


def _phys(consts):
    global _default_phys
    if _default_phys is None:
        cfg = load_config()
        _default_phys = (cfg['parameters'], cfg['dimensions'], cfg['constants'])
    prm, dims, cst = (dict(d) for d in _default_phys)
    if 'porosity' in consts:
        prm['porosity_wick'] = consts['porosity']
    if 'T_melt' in consts:
        prm['T_melting_sodium'] = consts['T_melt']
    if 'dT_melt' in consts:
        prm['delta_T_sodium'] = consts['dT_melt']
    if 'M_g' in consts:
        prm['M_g_sodium'] = consts['M_g']
    if 'R_vc' in consts:
        dims['R_vc'] = consts['R_vc']
    if 'R_gas' in consts:
        cst['R'] = consts['R_gas']
    return make_phys(prm, dims, cst)


def evaluate(name, T, consts):
    import torch
    lib = _lib.load()
    ph = _phys(consts)
    Tc = T.to(torch.float64).contiguous()
    out = torch.empty_like(Tc)
    stream = torch.cuda.current_stream(Tc.device).cuda_stream
    with torch.cuda.device(Tc.device):
        _lib.check(lib.hp_eval_property_raw(C.byref(ph), PROP_IDS[name], C.c_void_p(Tc.data_ptr()),
                                            C.c_void_p(out.data_ptr()), Tc.numel(), C.c_void_p(stream)),
                   'hp_eval_property_raw')
    return out.to(T.dtype) if T.dtype != torch.float64 else out
