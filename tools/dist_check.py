"""Multi-GPU correctness (run under torchrun on >= 2 GPUs): the sample-range-sharded render + NCCL film reduce equals the
single-GPU render of the same Sobol indices.   torchrun --nproc-per-node 2 tools/dist_check.py"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from ptina_b200 import scenes, worker, _native, dist as pdist

sc = scenes.cornell_monkey(nx=256, ny=256, spp=8)
worker.init(device=local)
ctx = _native.context()
scenes.apply(worker, sc)
spp = 8 * world
nx, ny = sc['size']
# two frames without a clear in between: the reduce is out of place, so the second one must not count anything twice
films = []
for frame in range(2):
    pdist.render_sharded(_native.ENGINE_PATH, spp)
    total = pdist.reduce_film_pass(0, 0)
    torch.cuda.synchronize()
    films.append(total.cpu().numpy().reshape(nx, ny, 4) if rank == 0 else None)
t_after = ctx.sobol_time
if rank == 0:
    worker.clear(); ctx.sobol_reset()
    for frame in range(2):
        ctx.render(_native.ENGINE_PATH, spp)
        film_one = ctx.get_film()
        film_dist = films[frame]
        assert np.array_equal(film_dist[..., 3], film_one[..., 3]) and (film_one[..., 3] == spp * (frame + 1)).all()
        err = np.abs(film_dist[..., :3] - film_one[..., :3]).max() / film_one[..., :3].max()
        print(f'dist_check world={world} frame {frame}: max |sharded+reduced - single| / max = {err:.3e}')
        assert err < 1e-5
    assert t_after == ctx.sobol_time == 64 + 2 * spp
    print('dist_check ok')
dist.barrier()
dist.destroy_process_group()
