"""Multi-GPU host logic: one process per GPU, the frame partitioned by image TILES x SAMPLE SETS (SURVEY.md 8e;
every (pixel, sample) of tracer.fs main() :436-518 is independent, and the reference's README.md:26-28 lists
"Tiled rendering" as a TODO).

The collectives themselves run INSIDE the library, behind the C ABI, on the context's own stream over NCCL
(include/fspt_b200.h: fspt_comm_init / fspt_reduce_accum / fspt_scene_broadcast), so a Node host reaches them through
the N-API shim exactly like this Python host does.  torch.distributed only carries the 128-byte NCCL unique id from
rank 0 to the other processes (and the gloo tests of the partition logic).
"""
import os

import numpy as np


def share_host_threads(local_world=None):
    """One process per GPU: divide the node's cores between the processes' scene-upload staging threads."""
    if local_world is None:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    n = max(4, (os.cpu_count() or 4) // max(1, local_world))
    os.environ.setdefault("FSPT_UPLOAD_THREADS", str(n))
    return n


def shard_ticks(n_total, rank, world):
    """Sample-set sharding: the tick indices rank `rank` renders (r, r+G, ...).  Unequal when world does not divide
    n_total -- the library keeps every pixel's own sample count, so the resolved image is still exact."""
    return np.arange(rank, n_total, world)


def tile_grid(world, width, height, n_tiles=None):
    """tiles x sample-sets factorisation of `world` ranks.  Tiles are horizontal bands whose height is a multiple of 4
    and whose width is the frame's (so the 8x4 pixel ordering of the traversal kernel applies inside every tile);
    default: as many tiles as keeps >= 64 samples in flight per wave at this resolution (64 Mi paths per wave)."""
    if n_tiles is None:
        n_tiles = 1
        while n_tiles * 2 <= world and world % (n_tiles * 2) == 0 and (width * height) // n_tiles * 64 > (64 << 20):
            n_tiles *= 2
    if world % n_tiles:
        raise ValueError("n_tiles %d does not divide world %d" % (n_tiles, world))
    return n_tiles, world // n_tiles


def partition(rank, world, width, height, n_samples, n_tiles=None):
    """What rank `rank` renders: ((x0, y0, w, h), tick indices).  rank = tile * n_sets + sample_set."""
    n_tiles, n_sets = tile_grid(world, width, height, n_tiles)
    tile, sset = rank // n_sets, rank % n_sets
    rows4 = (height + 3) // 4                        # bands in units of 4 rows
    r0, r1 = rows4 * tile // n_tiles * 4, min(height, rows4 * (tile + 1) // n_tiles * 4)
    if tile == n_tiles - 1:
        r1 = height
    return (0, r0, width, r1 - r0), shard_ticks(n_samples, sset, n_sets)


def init_comm(ctx, rank=None, world=None):
    """fspt_comm_init on every rank; the NCCL unique id travels through torch.distributed's object broadcast (any
    backend).  Plumbing only."""
    import torch.distributed as dist
    from . import capi
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    box = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    ctx.comm_init(box[0], rank, world)


def reduce_accum(ctx, dst=0, **_ignored):
    """sum-mode accumulation targets of all ranks -> rank dst, inside the library (ncclReduce on the context's
    stream, ordered after the renders already enqueued).  The alpha channel carries per-pixel sample counts."""
    ctx.reduce_accum(dst)


def reduce_arrays_cpu(local_sum, dst=0):
    """gloo path used by the CPU tests of the sharding logic: numpy sum buffer in, reduced buffer out on dst."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_sum, np.float32).copy())
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t.numpy()
