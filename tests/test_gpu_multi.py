"""Multi-GPU parity (needs >= 2 B200s on one box: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`;
skipped on a single-GPU box).  Sample-set sharding + NCCL reduce(sum) over NVLink must equal the sum of the
oracle's per-sample colours, and its mean must match the reference running mean within f32 summation-order error."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
W, H, N = 96, 64, 8


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from fspt_b200 import capi, dist as fdist, scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sa, cam = scenes.bunny_class(subdiv=3, atlas_res=32, env_size=(128, 64))
    ctx = capi.Context(W, H, rank)
    ctx.scene_upload(sa)
    ctx.set_accum_mode(1)
    rc, rt = scenes.rand_bases(N, 21)
    ticks = fdist.shard_ticks(N, rank, world)
    fr = ctx.frame(cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), cam["env_theta"])
    ctx.clear()
    ctx.render(fr, 0, rc[ticks], rt[ticks])
    fdist.reduce_accum(ctx, dst=0, n_local_samples=len(ticks), world=world, device_index=rank)
    torch.cuda.synchronize()
    if rank == 0:
        np.save(os.path.join(out_dir, "sum.npy"), ctx.read_accum())
        np.save(os.path.join(out_dir, "rgba.npy"), ctx.resolve())
    ctx.close()
    dist.destroy_process_group()


def test_two_gpu_sample_sharding_matches_oracle(tmp_path, oracle_mod):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from fspt_b200 import scenes
    mp.spawn(_worker, args=(2, 29650 + os.getpid() % 300, str(tmp_path)), nprocs=2, join=True)
    sa, cam = scenes.bunny_class(subdiv=3, atlas_res=32, env_size=(128, 64))
    O = oracle_mod.Oracle(sa)
    rc, rt = scenes.rand_bases(N, 21)
    cols, mean = [], None
    for k in range(N):
        pos, d = oracle_mod.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k])
        _, col, _ = O.trace(pos, d, W, H, 0, rt[k], cam["env_theta"], want_color=True)
        cols.append(col[..., :3])
        mean = cols[k] if mean is None else (cols[k] + mean * np.float32(k)) / np.float32(k + 1)
    z = np.zeros_like(cols[0])
    ref = ((((z + cols[0]) + cols[2]) + cols[4]) + cols[6]) + ((((z + cols[1]) + cols[3]) + cols[5]) + cols[7])
    got = np.load(tmp_path / "sum.npy")[..., :3]
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32))
    assert np.allclose(got / N, mean, rtol=3e-6, atol=1e-7)
    rgba = np.load(tmp_path / "rgba.npy")
    full = np.zeros((H, W, 4), np.float32)
    full[..., :3] = got / np.float32(N)
    assert np.abs(rgba.astype(int) - oracle_mod.draw(full).astype(int)).max() <= 1
