"""Multi-GPU: one process per GPU, sample ranges sharded across ranks, film combined with one NCCL reduce.

The reference is single-device.  Samples are independent and the film is additive (filmtable.py:14, path.py:93), so
the natural partition is the SAMPLE RANGE: rank g of G renders Sobol point indices k_first+g, k_first+g+G, ... for every
pixel (perfect load balance; the device evaluates any Gray-code point in closed form), with the scene and BVH
replicated (the build is deterministic).  The one exchange step is a sum-reduce of film pass 0 over NVLink
(nx*ny*16 bytes; 33 MB at 1080p); resolve (rgb / w) then happens on the destination rank.
"""
import torch
import torch.distributed as dist

from . import _native


def shard_range(k_first, count, rank, world):
    """Interleaved split of `count` consecutive sample indices starting at k_first: -> (first, count, stride)."""
    mine = (count - rank + world - 1) // world if count > rank else 0
    return k_first + rank, mine, world


def render_sharded(engine, nsamples, rank=None, world=None):
    """Every rank calls this with the same arguments; advances the Sobol time by `nsamples` on all ranks."""
    ctx = _native.context()
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    k_first = ctx.sobol_time + 1
    first, count, stride = shard_range(k_first, nsamples, rank, world)
    if count:
        ctx.render_range(engine, first, count, stride)
    ctx.sobol_time = ctx.sobol_time + nsamples


def reduce_film(film, dst=0, group=None):
    """Sum `film` (a tensor view of film pass 0, or any tensor in CPU tests) onto rank `dst`."""
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return film


def reduce_film_pass(id=0, dst=0):
    """In-place reduce of the native film pass onto rank dst (ordered on torch's current stream)."""
    return reduce_film(_native.context().film_tensor(id), dst)
