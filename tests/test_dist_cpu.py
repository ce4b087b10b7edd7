"""Multi-process logic on CPU (gloo, world_size 2): the sample-range sharding covers every Sobol index exactly once and
the film reduce reproduces the single-process sum."""
import os
import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ptina_b200.dist import shard_range, shard_chains, reduce_film, reduce_out_of_place


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


def test_shard_chains_partitions():
    for world in (1, 2, 3, 4, 8):
        for n in (1, 7, 2**14, 2**18):
            parts = [shard_chains(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


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
    # MLT sharding (mltpath.py:47-52 splats): every chain adds (L, 1) at its own pixel; chains are split over the ranks
    nch = 1000
    first, cnt = shard_chains(nch, rank, world)
    splat = torch.zeros(nx * ny, 4)
    for c in range(first, first + cnt):
        g = torch.Generator().manual_seed(10_000 + c)
        px = int(torch.randint(0, nx * ny, (1,), generator=g))
        splat[px, :3] += torch.rand(3, generator=g)
        splat[px, 3] += 1
    tot = reduce_out_of_place(splat, dst=0)
    if rank == 0:
        torch.save(tot.clone(), out + '.mlt')
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
    mlt = torch.load(out + '.mlt')
    want = torch.zeros(48, 4)
    for c in range(1000):
        g = torch.Generator().manual_seed(10_000 + c)
        px = int(torch.randint(0, 48, (1,), generator=g))
        want[px, :3] += torch.rand(3, generator=g)
        want[px, 3] += 1
    assert torch.allclose(mlt, want, rtol=1e-5, atol=1e-6) and torch.equal(mlt[:, 3], want[:, 3]) and mlt[:, 3].sum() == 1000
