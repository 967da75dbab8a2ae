"""Multi-GPU plumbing: one process per GPU, frame sharded by SAMPLE SETS (rank r renders ticks r, r+G, ... of
the same frame with its own rand-base entries), per-GPU f32 sum buffers combined by one NCCL reduce(sum) over
NVLink / NVSwitch per batch, root divides by the total sample count and runs the post-pass.
torch.distributed is plumbing only; the accumulation buffer stays in the library's device memory and is exposed
to torch zero-copy through __cuda_array_interface__.  gloo (CPU tests) goes through a host copy."""
import numpy as np


class _DevBuf:
    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (ptr, False), "version": 3}


def accum_tensor(ctx, device_index):
    import torch
    ptr, n, _ = ctx.accum_device_ptr()
    return torch.as_tensor(_DevBuf(ptr, n), device=torch.device("cuda", device_index))


def share_host_threads(local_world=None):
    """One process per GPU: divide the node's cores between the processes' scene-upload staging threads."""
    import os
    if local_world is None:
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", os.environ.get("WORLD_SIZE", "1")))
    n = max(4, (os.cpu_count() or 4) // max(1, local_world))
    os.environ.setdefault("FSPT_UPLOAD_THREADS", str(n))
    return n


def shard_ticks(n_total, rank, world):
    """Sample-set sharding: the tick indices rank `rank` renders."""
    return np.arange(rank, n_total, world)


def reduce_accum(ctx, dst, n_local_samples, world, device_index=None):
    """sum-mode accumulation buffers -> rank dst (NCCL reduce); dst then holds sum over world*n_local samples."""
    import torch
    import torch.distributed as dist
    if device_index is None:
        device_index = torch.cuda.current_device()
    t = accum_tensor(ctx, device_index)  # synchronises the library stream
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    if dist.get_rank() == dst:
        ctx.set_accum_samples(n_local_samples * world)


def reduce_arrays_cpu(local_sum, dst=0):
    """gloo path used by the CPU tests of the sharding logic: numpy sum buffer in, reduced buffer out on dst."""
    import torch
    import torch.distributed as dist
    t = torch.from_numpy(np.ascontiguousarray(local_sum, np.float32).copy())
    dist.reduce(t, dst=dst, op=dist.ReduceOp.SUM)
    return t.numpy()
