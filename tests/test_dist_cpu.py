"""Multi-process logic on CPU (gloo, world_size 2): the sample-range sharding covers every Sobol index exactly once and
the film reduce reproduces the single-process sum."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ptina_b200.dist import shard_range, reduce_film, reduce_out_of_place


def test_shard_range_partitions():
    for world in (1, 2, 3, 4, 8):
        for count in (0, 1, 5, 32, 1024):
            got = []
            for r in range(world):
                first, n, stride = shard_range(65, count, r, world)
                got += [first + i * stride for i in range(n)]
            assert sorted(got) == list(range(65, 65 + count))
            sizes = [shard_range(65, count, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    # a fake per-sample "film contribution" that depends only on the Sobol index: each rank accumulates its shard
    nx, ny, count = 8, 6, 13
    film = torch.zeros(nx * ny, 4)
    first, n, stride = shard_range(65, count, rank, world)
    for i in range(n):
        k = first + i * stride
        g = torch.Generator().manual_seed(k)
        film[:, :3] += torch.rand(nx * ny, 3, generator=g)
        film[:, 3] += 1
    # out of place, twice (a frame, more samples, a second frame): the per-rank film must keep only its own samples
    total = reduce_out_of_place(film, dst=0)
    own = film.clone()
    total2 = reduce_out_of_place(film, dst=0)
    assert torch.equal(film, own)
    if rank == 0:
        assert torch.equal(total2, total)
        torch.save(total.clone(), out)
    else:
        assert total is None and total2 is None
    reduce_film(film, dst=0)          # the in-place primitive still sums onto dst
    if rank == 0:
        assert torch.allclose(film, torch.load(out), rtol=1e-6, atol=1e-6)
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_film_reduce_world2(tmp_path):
    out = str(tmp_path / 'film.pt')
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    film = torch.load(out)
    want = torch.zeros(48, 4)
    for k in range(65, 78):
        g = torch.Generator().manual_seed(k)
        want[:, :3] += torch.rand(48, 3, generator=g)
        want[:, 3] += 1
    assert torch.allclose(film, want, rtol=1e-6, atol=1e-6) and torch.equal(film[:, 3], want[:, 3])
