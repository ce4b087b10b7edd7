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


def shard_chains(nchains, rank, world):
    """Contiguous split of the MLT chains (engine/mltpath.py:12: 2^18 of them): -> (first, count); every chain on exactly one rank.
    Chain c draws from the Philox stream keyed (seed, c, iteration) wherever it runs, so the split does not change what a chain does."""
    base, extra = divmod(nchains, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


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


_staging = {}


def reduce_film_pass(id=0, dst=0, group=None):
    """Sum film pass `id` of every rank onto rank `dst`, OUT OF PLACE: each rank's own film keeps only its own samples, so the call
    is idempotent -- render more, reduce again, and nothing is counted twice (an in-place reduce would leave rank dst's film holding
    everybody's samples and re-add them on the next call).  Returns the summed film [nx*ny, 4] on rank dst (a staging tensor that the
    next call reuses) and None elsewhere.  Ordered on torch's current stream."""
    return reduce_out_of_place(_native.context().film_tensor(id), dst, group)


def reduce_out_of_place(film, dst=0, group=None):
    """The out-of-place reduce of reduce_film_pass on any tensor (CPU tensors in the gloo tests)."""
    key = (film.device, tuple(film.shape))
    buf = _staging.get(key)
    if buf is None:
        _staging.clear()
        buf = _staging[key] = torch.empty_like(film)
    buf.copy_(film)
    reduce_film(buf, dst, group)
    return buf if (not dist.is_initialized() or dist.get_rank(group) == dst) else None


def resolve(total):
    """filmtable.py:47-63 on a reduced film [nx*ny, 4] -> image [nx, ny, 4] (rgb / w, a = 1; magenta where w == 0)."""
    nx, ny = _native.context().get_size()
    w = total[:, 3:4]
    img = torch.where(w != 0, torch.cat([total[:, :3] / w, torch.ones_like(w)], 1),
                      torch.tensor([0.9, 0.4, 0.9, 0.0], device=total.device).expand_as(total))
    return img.reshape(nx, ny, 4)
