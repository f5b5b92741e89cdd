"""heat_pipe_solver_b200 -- B200-native transient r-z conduction hot path of kubaniak/heat-pipe-solver.

Python surface mirrors the reference's modules (`params`, `mesh`, `material_properties`, `main`);
the arithmetic lives in `libhpsolver.so` (hand-written sm_100a CUDA behind the C ABI of
`include/hp_solver.h`).  Importing this package does not need a GPU; computing does.
"""
from . import params, mesh, material_properties, regions, report  # noqa: F401
from .mesh import generate_composite_mesh, generate_mesh_2d, StructuredGrid2D  # noqa: F401

__all__ = ['params', 'mesh', 'material_properties', 'regions', 'report', 'generate_composite_mesh', 'generate_mesh_2d',
           'StructuredGrid2D', 'ConductionSolver', 'run', 'run_adaptive']


def __getattr__(name):
    if name == 'ConductionSolver':
        from .solver import ConductionSolver
        return ConductionSolver
    if name == 'run':
        from .main import run
        return run
    if name == 'run_adaptive':
        from .main import run_adaptive
        return run_adaptive
    raise AttributeError(name)
