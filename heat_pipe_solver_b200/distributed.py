"""Axial-slab decomposition across the GPUs of one box (one process per GPU).

Replaces FiPy's `parallelComm` / PETSc ghost scatter (reference main.py:11,339, tests/test_parallel.py):
FiPy partitions `Grid2D` along y -- the axial index here -- and so do we, because rows are contiguous in
memory (radial index fastest) and the radial-line preconditioner stays rank-local.  Each rank owns rows
[j0, j1) plus one ghost row per neighbour; per PCG iteration one row of z travels to each neighbour and
the CG scalars are combined over all ranks (NCCL, see csrc/hp_api.cu).

The partitioning helpers are pure host logic (tested on CPU with the gloo backend, world_size 2).
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def slab_rows(nz, rank, world, align=1):
    """Contiguous block of axial rows for `rank`.  The cuts fall on multiples of `align` rows (sizes differ by at most one
    block of `align` rows; the rows beyond the last whole block go to the last rank)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    blocks = nz // align
    if world > blocks:
        raise ValueError(f"cannot split {nz} rows over {world} ranks in blocks of {align}")
    base, extra = divmod(blocks, world)
    b0 = rank * base + min(rank, extra)
    b1 = b0 + base + (1 if rank < extra else 0)
    return b0 * align, (nz if rank == world - 1 else b1 * align)


def mg_level_count(nz, max_levels=6):
    """Levels of the 'mgline' hierarchy of a mesh with nz rows (rows halve per level; a level is added while the one above
    has >= 32 rows) -- the rule of csrc/hp_api.cu::local_level_count."""
    n, rows = 1, nz
    while n < max_levels and rows >= 32:
        n += 1
        rows = (rows + 1) // 2
    return n


def slab_alignment(nz, world, max_levels=6):
    """Row alignment of the slab cuts under which the distributed V-cycle is the single-domain one: every level pairs
    rows 2J, 2J+1, so every slab but the last must hold a multiple of 2^(levels-1) rows.  Fewer levels if the mesh is too
    short for that many whole blocks per rank."""
    n = mg_level_count(nz, max_levels)
    while n > 1 and nz // (1 << (n - 1)) < world:
        n -= 1
    return 1 << (n - 1)


def resolve_precond(precond, nz, nr, world, comm, max_levels=6):
    """'auto' decided from the GLOBAL mesh so that every rank runs the same kernel sequence (a per-rank decision can
    differ between slabs that differ by one row).  Without peer memory the V-cycle is rank-local: every slab must be tall
    enough to coarsen, else all ranks use the line preconditioner."""
    if precond == 'auto':
        precond = 'mgline' if (nz >= 256 and nz * nr >= 262144) else 'rline'
    if precond == 'mgline' and world > 1:
        if comm == 'p2p':
            if mg_level_count(nz, max_levels) < 2 or slab_alignment(nz, world, max_levels) < 2:
                precond = 'rline'
        elif min(b - a for a, b in (slab_rows(nz, r, world) for r in range(world))) < 32:
            precond = 'rline'
    return precond


def neighbours(rank, world):
    return (rank - 1 if rank > 0 else None), (rank + 1 if rank < world - 1 else None)


def gather_field(local, group=None):
    """All-gather the owned rows of every rank into the global (nz, nr) field (torch.distributed plumbing)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    sizes = [torch.zeros(1, dtype=torch.int64, device=local.device) for _ in range(world)]
    dist.all_gather(sizes, torch.tensor([local.shape[0]], dtype=torch.int64, device=local.device), group=group)
    rows = [int(s.item()) for s in sizes]
    most = max(rows)                       # all_gather wants equal shapes: pad to the tallest slab, trim after
    padded = torch.zeros((most, local.shape[1]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:n] for p, n in zip(parts, rows)], dim=0)


def make_solver_for_rank(mesh, cfg, rank, world, comm='p2p', **solver_kw):
    """ConductionSolver for this rank's slab of `mesh` (the whole mesh when world == 1).

    comm = 'nccl': ghost rows by ncclSend/Recv and CG scalars by ncclAllGather between the kernels (host-driven loop).
    comm = 'p2p' : the same exchange as NVLink peer-memory stores fused into the PCG kernels (CUDA-IPC mapped
                   buffers); the time step is one CUDA graph per rank.  NCCL is then only used for the first ghost
                   refresh after a field upload."""
    from .solver import ConductionSolver
    if world == 1:
        return ConductionSolver(mesh, cfg['dimensions'], cfg['parameters'], cfg['constants'], **solver_kw)
    import torch
    import torch.distributed as dist
    precond = resolve_precond(solver_kw.pop('precond', 'auto'), mesh.nz, mesh.nr, world, comm, solver_kw.get('mg_levels', 6))
    align = slab_alignment(mesh.nz, world, solver_kw.get('mg_levels', 6)) if (precond == 'mgline' and comm == 'p2p') else 1
    rows = slab_rows(mesh.nz, rank, world, align)
    solver = ConductionSolver(mesh, cfg['dimensions'], cfg['parameters'], cfg['constants'], rows=rows, precond=precond, **solver_kw)
    # NCCL communicator of the library itself: rank 0 makes the id, torch.distributed carries it
    ident = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        buf = (C.c_uint8 * 128)()
        from . import _lib
        _lib.check(solver.lib.hp_comm_unique_id(buf), 'hp_comm_unique_id')
        ident = torch.tensor(list(buf), dtype=torch.uint8)
    ident = ident.cuda()
    dist.broadcast(ident, src=0)
    solver.init_comm(rank, world, bytes(ident.cpu().numpy().tobytes()))
    if comm == 'p2p':
        nzs = [b - a for a, b in (slab_rows(mesh.nz, r, world, align) for r in range(world))]
        mine = torch.frombuffer(bytearray(solver.peer_export(nzs)), dtype=torch.uint8).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        handles = b''.join(bytes(g.cpu().numpy().tobytes()) for g in gathered)
        solver.init_peer(handles, nzs)
        dist.barrier()
        if precond == 'mgline' and solver.effective_precond != 'mgline':
            raise RuntimeError("mgline was resolved for the slabs but the library fell back: inconsistent level count")
    elif comm != 'nccl':
        raise ValueError("comm must be 'p2p' or 'nccl'")
    return solver
