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


def slab_rows(nz, rank, world):
    """Contiguous block of axial rows for `rank`; sizes differ by at most one row."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    if world > nz:
        raise ValueError(f"cannot split {nz} rows over {world} ranks")
    base, extra = divmod(nz, world)
    j0 = rank * base + min(rank, extra)
    return j0, j0 + base + (1 if rank < extra else 0)


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
    rows = slab_rows(mesh.nz, rank, world)
    solver = ConductionSolver(mesh, cfg['dimensions'], cfg['parameters'], cfg['constants'], rows=rows, **solver_kw)
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
        nzs = [b - a for a, b in (slab_rows(mesh.nz, r, world) for r in range(world))]
        mine = torch.frombuffer(bytearray(solver.peer_export(nzs)), dtype=torch.uint8).cuda()
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        handles = b''.join(bytes(g.cpu().numpy().tobytes()) for g in gathered)
        solver.init_peer(handles, nzs)
        dist.barrier()
    elif comm != 'nccl':
        raise ValueError("comm must be 'p2p' or 'nccl'")
    return solver
