#!/usr/bin/env python
"""Times fspt_scene_upload (rank 0) and fspt_scene_broadcast (all ranks) on the bench scene:
   torchrun --nproc-per-node N tools/bcast_time.py    (NCCL_DEBUG=INFO shows the transport NCCL picked)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from fspt_b200 import capi, scenes, dist as fdist

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("gloo")
fdist.share_host_threads()
sa, cam = scenes.bunny_class(subdiv=6, atlas_res=2048)
ctx = capi.Context(1280, 720, lr)
fdist.init_comm(ctx, rank, world)
for it in range(5):
    dist.barrier()
    t0 = time.perf_counter()
    if rank == 0:
        ctx.scene_upload(sa)
        ctx.synchronize()
    t1 = time.perf_counter()
    ctx.scene_broadcast(0)
    ctx.synchronize()
    t2 = time.perf_counter()
    dist.barrier()
    t3 = time.perf_counter()
    print("rank %d iter %d: upload %.2f ms  broadcast %.2f ms  (+barrier %.2f ms)" % (rank, it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3), flush=True)
# and a reduce of the accumulation target
ctx.set_accum_mode(1)
for it in range(3):
    dist.barrier()
    t0 = time.perf_counter()
    ctx.reduce_accum(0)
    ctx.synchronize()
    print("rank %d reduce %.2f ms" % (rank, (time.perf_counter() - t0) * 1e3), flush=True)
ctx.close()
dist.destroy_process_group()
