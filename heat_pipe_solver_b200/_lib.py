"""ctypes binding of libhpsolver.so (the C ABI declared in include/hp_solver.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is
visible, the product path raises.  `load()` only dlopens the library (works on a
CPU-only box, used by the symbol-export test); anything that computes needs a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, 'libhpsolver.so')

c_double_p = C.POINTER(C.c_double)
c_int8_p = C.POINTER(C.c_int8)
c_uint8_p = C.POINTER(C.c_uint8)


class HpPhys(C.Structure):
    _fields_ = [(n, C.c_double) for n in ('R_vc', 'M_g', 'R_gas', 'porosity', 'T_melt', 'dT_melt', 'sigma',
                                          'simple_rho', 'simple_cp', 'simple_k')] + \
               [(n, C.c_int32) for n in ('properties', 'gamma', 'source', 'face_k', 'radiation', '_pad')]


class HpMeshDesc(C.Structure):
    _fields_ = [('nr', C.c_int32), ('nz', C.c_int32), ('cartesian', C.c_int32), ('_pad', C.c_int32),
                ('r_v', c_double_p), ('z_v', c_double_p),
                ('cell_band', c_int8_p), ('rface_band', c_int8_p), ('zface_band', c_int8_p),
                ('zc_flags', c_uint8_p), ('zv_flags', c_uint8_p), ('axial_open', c_uint8_p),
                ('q_line', c_double_p), ('h_line', c_double_p), ('Tamb_line', c_double_p), ('eps_line', c_double_p)]


class HpSolveOpts(C.Structure):
    _fields_ = [('precond', C.c_int32), ('criterion', C.c_int32), ('tol', C.c_double), ('max_iter', C.c_int32),
                ('loop_mode', C.c_int32), ('poll_chunk', C.c_int32), ('refactor_every', C.c_int32),
                ('extrapolate', C.c_int32), ('mg_levels', C.c_int32), ('mg_coarse_sweeps', C.c_int32), ('reuse_converged', C.c_int32),
                ('mg_omega', C.c_double), ('while_unroll', C.c_int32), ('peer_timeout_ms', C.c_int32),
                ('mg_zscale', C.c_double), ('tma_rows', C.c_int32), ('pipe_kernel', C.c_int32),
                ('_pad', C.c_int32), ('ext_guard_tols', C.c_double)]


class HpSweepStat(C.Structure):
    _fields_ = [('residual_l2', C.c_double), ('rhs_l2', C.c_double), ('crit', C.c_double),
                ('iters', C.c_int32), ('flags', C.c_int32)]


class HpKernelTimes(C.Structure):
    _fields_ = [('count', C.c_int64 * 16), ('total_ms', C.c_double * 16),
                ('full_count', C.c_int64 * 16), ('full_ms', C.c_double * 16)]


KERNEL_KINDS = ('other', 'coef', 'diag', 'factor', 'cg_init', 'cg_A', 'cg_B', 'mg_res', 'mg_line', 'mg_coarse', 'mg_setup',
                'mg_l1', 'mg_l2', 'mg_l3', 'mg_l4', 'mg_l5')

# every symbol include/hp_solver.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SIGNATURES = {
    'hp_last_error': (C.c_char_p, []),
    'hp_version': (C.c_int, []),
    'hp_create': (C.c_int, [C.POINTER(_H), C.POINTER(HpMeshDesc), C.POINTER(HpPhys), C.c_void_p]),
    'hp_destroy': (C.c_int, [_H]),
    'hp_set_stream': (C.c_int, [_H, C.c_void_p]),
    'hp_set_opts': (C.c_int, [_H, C.POINTER(HpSolveOpts)]),
    'hp_get_opts': (C.c_int, [_H, C.POINTER(HpSolveOpts)]),
    'hp_set_T_dev': (C.c_int, [_H, C.c_void_p]),
    'hp_get_T_dev': (C.c_int, [_H, C.c_void_p]),
    'hp_set_T_host': (C.c_int, [_H, C.c_void_p]),
    'hp_get_T_host': (C.c_int, [_H, C.c_void_p]),
    'hp_T_ptr': (C.c_void_p, [_H]),
    'hp_update_old': (C.c_int, [_H]),
    'hp_sweep': (C.c_int, [_H, C.c_double, C.c_int, c_double_p]),
    'hp_step': (C.c_int, [_H, C.c_double, C.c_int, C.c_int]),
    'hp_run': (C.c_int, [_H, C.c_double, C.c_int, C.c_int, C.c_int]),
    'hp_run_host': (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_int]),
    'hp_get_stats': (C.c_int, [_H, C.POINTER(HpSweepStat), C.c_int]),
    'hp_total_sweeps': (C.c_int64, [_H]),
    'hp_total_iters': (C.c_int64, [_H]),
    'hp_kernel_launches': (C.c_int64, [_H]),
    'hp_assemble_debug': (C.c_int, [_H, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    'hp_eval_property': (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    'hp_eval_property_raw': (C.c_int, [C.POINTER(HpPhys), C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'hp_eval_cell_fields': (C.c_int, [_H, C.c_void_p, C.c_void_p, C.c_void_p]),
    'hp_profile_begin': (C.c_int, [_H, C.c_int]),
    'hp_profile_end': (C.c_int, [_H, C.POINTER(HpKernelTimes)]),
    'hp_comm_unique_id': (C.c_int, [C.c_void_p]),
    'hp_comm_init': (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p]),
    'hp_peer_export': (C.c_int, [_H, C.POINTER(C.c_int32), C.c_void_p]),
    'hp_peer_import': (C.c_int, [_H, C.c_void_p, C.POINTER(C.c_int32)]),
    'hp_comm_fault': (C.c_int, [_H]),
    'hp_comm_clear_fault': (C.c_int, [_H]),
}

_lib = None


class HpError(RuntimeError):
    pass


def load():
    """dlopen libhpsolver.so and attach prototypes.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise HpError(f'{LIB_PATH} is missing: build it with `python -m heat_pipe_solver_b200.build` '
                      '(nvcc, sm_100a).  There is no CPU fallback for the conduction hot path.')
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=''):
    if rc != 0:
        msg = load().hp_last_error()
        raise HpError(f'{what}: {msg.decode() if msg else "error"} (rc={rc})')
