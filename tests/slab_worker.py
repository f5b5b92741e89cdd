"""Worker of the multi-GPU slab test (one process per GPU; launched by tests/test_gpu_slabs.py or torchrun).
Runs a few steps on an axial-slab decomposition and writes the gathered field of rank 0 to an .npy file."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(rank, world, port, shape, n_steps, sweeps, out_path, comm='nccl', loop_mode='host', precond='rline'):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from heat_pipe_solver_b200 import generate_composite_mesh, params
    from heat_pipe_solver_b200.distributed import gather_field, make_solver_for_rank
    from oracle import fv_oracle as fo
    cfg = params.load_config()
    mesh, _ = generate_composite_mesh(fo.synthetic_mesh_params(*shape), cfg['dimensions'])
    solver = make_solver_for_rank(mesh, cfg, rank, world, comm=comm, loop_mode=loop_mode, poll_chunk=4, precond=precond)
    j0, j1 = solver.layout.j0, solver.layout.j1
    z = mesh.z_c[:, None] / mesh.z_v[-1]
    T0 = 290.0 + 250.0 * np.exp(-((z - 0.45) / 0.3) ** 2) * np.ones((1, mesh.nr))     # hot spot across the slab cut
    solver.set_value(T0[j0:j1])
    solver.run(0.1, n_steps, sweeps=sweeps, has_old=True)
    full = gather_field(solver.value_tensor())
    stats = solver.stats(sweeps * n_steps)
    assert not any(s.get('peer_timeout') for s in stats), stats
    if rank == 0:
        np.save(out_path, full.cpu().numpy())
        np.save(out_path + '.iters.npy', np.array([s['iters'] for s in stats]))
    solver.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), tuple(int(v) for v in sys.argv[4].split(',')),
         int(sys.argv[5]), int(sys.argv[6]), sys.argv[7], *(sys.argv[8:11]))
