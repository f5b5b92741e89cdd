"""The C-ABI library loads on a CPU-only box and exports every symbol include/hp_solver.h declares;
computing without a GPU fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, 'include', 'hp_solver.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(hp_[a-z_A-Z0-9]+)\s*\(', src)))


def test_header_symbols_are_exported_and_bound(built_lib):
    from heat_pipe_solver_b200 import _lib
    lib = _lib.load()
    names = _declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/hp_solver.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    assert set(_lib.SIGNATURES) == set(names)
    assert lib.hp_version() >= 100


def test_struct_layouts_match_header():
    from heat_pipe_solver_b200 import _lib
    assert C.sizeof(_lib.HpPhys) == 10 * 8 + 6 * 4
    assert C.sizeof(_lib.HpMeshDesc) == 4 * 4 + 12 * 8
    assert C.sizeof(_lib.HpSolveOpts) == 96
    assert C.sizeof(_lib.HpSweepStat) == 32
    assert C.sizeof(_lib.HpKernelTimes) == 512


def test_no_cpu_fallback(built_lib, cfg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from heat_pipe_solver_b200 import ConductionSolver, generate_composite_mesh, _lib
    mesh, _ = generate_composite_mesh(dict(cfg['mesh']), cfg['dimensions'])
    with pytest.raises(_lib.HpError, match='no CPU fallback'):
        ConductionSolver(mesh, cfg['dimensions'], cfg['parameters'], cfg['constants'])
    # and the C entry point itself refuses
    lib = _lib.load()
    h = C.c_void_p()
    md, ph = _lib.HpMeshDesc(), _lib.HpPhys()
    arrs = [np.zeros(8) for _ in range(12)]
    md.nr, md.nz = 2, 2
    for name, a in zip(['r_v', 'z_v', 'q_line', 'h_line', 'Tamb_line', 'eps_line'], arrs):
        setattr(md, name, a.ctypes.data_as(_lib.c_double_p))
    b = np.zeros(8, dtype=np.int8)
    u = np.zeros(8, dtype=np.uint8)
    for name in ('cell_band', 'rface_band', 'zface_band'):
        setattr(md, name, b.ctypes.data_as(_lib.c_int8_p))
    for name in ('zc_flags', 'zv_flags', 'axial_open'):
        setattr(md, name, u.ctypes.data_as(_lib.c_uint8_p))
    assert lib.hp_create(C.byref(h), C.byref(md), C.byref(ph), None) != 0
    assert b'no CPU fallback' in lib.hp_last_error()


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, 'heat_pipe_solver_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_struct_field_offsets_match_the_c_compiler(tmp_path):
    """Every field of every ABI struct sits where a C compiler puts it for include/hp_solver.h (name, offset and size),
    so the ctypes mirror in _lib.py cannot drift from the header unnoticed."""
    import shutil
    import subprocess
    from heat_pipe_solver_b200 import _lib
    cc = shutil.which('gcc') or shutil.which('cc')
    if cc is None:
        pytest.skip("no C compiler")
    structs = {'hp_phys': _lib.HpPhys, 'hp_mesh_desc': _lib.HpMeshDesc, 'hp_solve_opts': _lib.HpSolveOpts,
               'hp_sweep_stat': _lib.HpSweepStat, 'hp_kernel_times': _lib.HpKernelTimes}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "hp_solver.h"', 'int main(void) {']
    for cname, ct in structs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu %zu\\n", offsetof({cname}, {fname}), sizeof((({cname}*)0)->{fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / 'layout.c'
    src.write_text('\n'.join(lines))
    exe = tmp_path / 'layout'
    subprocess.run([cc, '-std=c11', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)], check=True)
    got = dict((ln.split()[0], tuple(int(v) for v in ln.split()[1:])) for ln in
               subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert got[cname] == (C.sizeof(ct),), cname
        for fname, ftype in ct._fields_:
            fld = getattr(ct, fname)
            assert got[f'{cname}.{fname}'] == (fld.offset, fld.size), f'{cname}.{fname}'
