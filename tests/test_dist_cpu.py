"""Multi-GPU host logic on CPU: world_size-2 gloo.  Sample-set sharding (rank r renders ticks r, r+G, ...) and
the reduce(sum) of per-rank f32 sum buffers must equal the single-process sum over all ticks."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from fspt_b200 import dist as fdist, scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sa, cam = scenes.bunny_class(subdiv=2, atlas_res=16, env_size=(64, 32))
    O = oracle.Oracle(sa)
    W, H, N = 32, 24, 6
    rc, rt = scenes.rand_bases(N, 9)
    ticks = fdist.shard_ticks(N, rank, world)
    local = np.zeros((H, W, 4), np.float32)
    for k in ticks:  # per-rank sum buffer (accumulation mode 1 of the library), colours from the CPU oracle
        pos, d = oracle.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k], nthreads=1)
        _, col, _ = O.trace(pos, d, W, H, 0, rt[k], cam["env_theta"], want_color=True, nthreads=1)
        local += col
    total = fdist.reduce_arrays_cpu(local, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "reduced.npy"), total)
    np.save(os.path.join(out_dir, "ticks%d.npy" % rank), ticks)
    dist.destroy_process_group()


def test_sample_set_sharding_reduce_matches_single_process(tmp_path, oracle_mod):
    world, port = 2, 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from fspt_b200 import scenes
    t0, t1 = np.load(tmp_path / "ticks0.npy"), np.load(tmp_path / "ticks1.npy")
    assert sorted(np.concatenate([t0, t1]).tolist()) == list(range(6)) and not set(t0) & set(t1)
    sa, cam = scenes.bunny_class(subdiv=2, atlas_res=16, env_size=(64, 32))
    O = oracle_mod.Oracle(sa)
    W, H, N = 32, 24, 6
    rc, rt = scenes.rand_bases(N, 9)
    cols = []
    for k in range(N):
        pos, d = oracle_mod.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k])
        _, col, _ = O.trace(pos, d, W, H, 0, rt[k], cam["env_theta"], want_color=True)
        cols.append(col)
    # same association as the sharded run: (sum of even ticks) + (sum of odd ticks)
    ref = (cols[0] + cols[2] + cols[4]) + (cols[1] + cols[3] + cols[5])
    got = np.load(tmp_path / "reduced.npy")
    assert np.array_equal(got[..., :3], ref[..., :3])
    # and it is the running mean of the reference up to f32 summation order
    mean = None
    for k in range(N):
        mean = cols[k][..., :3] if mean is None else (cols[k][..., :3] + mean * np.float32(k)) / np.float32(k + 1)
    assert np.allclose(got[..., :3] / N, mean, rtol=2e-6, atol=1e-7)


def test_share_host_threads(monkeypatch):
    import os
    from fspt_b200 import dist as fdist
    monkeypatch.delenv("FSPT_UPLOAD_THREADS", raising=False)
    monkeypatch.setenv("LOCAL_WORLD_SIZE", "8")
    n = fdist.share_host_threads()
    assert n == max(4, (os.cpu_count() or 4) // 8) and os.environ["FSPT_UPLOAD_THREADS"] == str(n)
    monkeypatch.setenv("FSPT_UPLOAD_THREADS", "5")  # an explicit setting wins
    fdist.share_host_threads()
    assert os.environ["FSPT_UPLOAD_THREADS"] == "5"


def test_partition_covers_every_pixel_sample_exactly_once():
    """tiles x sample sets (fspt_b200.dist.partition): whatever the world size, every (pixel, tick) belongs to exactly
    one rank; tile heights are multiples of 4 rows (the 8x4 path ordering of k_trace) except possibly the last."""
    from fspt_b200 import dist as fdist
    for world, W, H, N, tiles in [(1, 64, 48, 5, None), (2, 64, 48, 5, 2), (4, 96, 50, 7, 2), (8, 384, 216, 24, 8),
                                  (8, 128, 72, 64, None), (8, 128, 70, 3, 8), (6, 64, 36, 10, 3)]:
        cover = np.zeros((H, W, N), np.int32)
        n_tiles, n_sets = fdist.tile_grid(world, W, H, tiles)
        assert n_tiles * n_sets == world
        for r in range(world):
            (x0, y0, w, h), ticks = fdist.partition(r, world, W, H, N, n_tiles=tiles)
            assert w == W and x0 == 0 and h > 0 and (y0 % 4 == 0)
            assert h % 4 == 0 or y0 + h == H
            cover[y0:y0 + h, x0:x0 + w][:, :, ticks] += 1
        assert np.all(cover == 1), (world, W, H, N)
    # 4K with 8 GPUs: the automatic grid picks tiles, so that a wave holds 64 samples again (8 when untiled)
    assert fdist.tile_grid(8, 3840, 2160)[0] == 8 and fdist.tile_grid(8, 1280, 720) == (1, 8)
    rows = np.zeros(2160, np.int32)
    for r in range(8):
        (x0, y0, w, h), ticks = fdist.partition(r, 8, 3840, 2160, 1024)
        rows[y0:y0 + h] += 1
        assert (x0, w) == (0, 3840) and len(ticks) == 1024 and h % 4 == 0
    assert np.all(rows == 1)


def _tile_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import oracle
    from fspt_b200 import dist as fdist, scenes
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sa, cam = scenes.bunny_class(subdiv=2, atlas_res=16, env_size=(64, 32))
    O = oracle.Oracle(sa)
    W, H, N = 32, 24, 5
    rc, rt = scenes.rand_bases(N, 9)
    (x0, y0, w, h), ticks = fdist.partition(rank, world, W, H, N, n_tiles=2)
    local = np.zeros((H, W, 4), np.float32)     # what the library's sum mode holds: rgb sum + sample count in alpha
    for k in ticks:
        pos, d = oracle.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k], nthreads=1)
        _, col, _ = O.trace(pos, d, W, H, 0, rt[k], cam["env_theta"], want_color=True, nthreads=1)
        local[y0:y0 + h, x0:x0 + w, :3] += col[y0:y0 + h, x0:x0 + w, :3]
        local[y0:y0 + h, x0:x0 + w, 3] += 1.0
    total = fdist.reduce_arrays_cpu(local, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "tiles.npy"), total)
    dist.destroy_process_group()


def test_tile_sharding_reduce_carries_per_pixel_counts(tmp_path, oracle_mod):
    """world 2 = 2 tiles x 1 sample set: the reduced buffer is the frame, every pixel with its own sample count (the
    divisor fspt_resolve uses), equal to the single-process sum."""
    world, port = 2, 29100 + (os.getpid() % 500)
    mp.spawn(_tile_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    from fspt_b200 import scenes
    sa, cam = scenes.bunny_class(subdiv=2, atlas_res=16, env_size=(64, 32))
    O = oracle_mod.Oracle(sa)
    W, H, N = 32, 24, 5
    rc, rt = scenes.rand_bases(N, 9)
    ref = np.zeros((H, W, 3), np.float32)
    for k in range(N):
        pos, d = oracle_mod.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), rc[k])
        _, col, _ = O.trace(pos, d, W, H, 0, rt[k], cam["env_theta"], want_color=True)
        ref = ref + col[..., :3]
    got = np.load(tmp_path / "tiles.npy")
    assert np.array_equal(got[..., :3], ref) and np.all(got[..., 3] == N)
