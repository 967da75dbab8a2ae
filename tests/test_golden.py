"""Committed golden fixtures (tests/golden/*.npz).  bunny_small / pbr_refractive_small / quad_kat hold OUTPUTS OF THE
REFERENCE'S OWN SHADERS (camera.fs, bvh_test.fs, tracer.fs, draw.fs run on the CPU through oracle/glsl_cpu where the
reference tree exists; generator tests/golden/make_golden.py, provenance recorded inside each file).
CPU: the oracle reproduces them bit for bit;  GPU: the CUDA path reproduces them bit for bit."""
import glob
import os
import types

import numpy as np
import pytest

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def load(path):
    z = np.load(path)
    sa = types.SimpleNamespace(**{k: z[k] for k in ("bvh", "tris", "mats", "norms", "uvs", "atlas", "env", "bins")})
    sa.leaf_size = 4
    return z, sa


def beq(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.dtype.kind == "f":
        return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))
    return bool(np.array_equal(a, b))


def test_shader_fixtures_say_where_they_come_from():
    made = [p for p in GOLDEN if os.path.basename(p) in ("bunny_small.npz", "pbr_refractive_small.npz", "quad_kat.npz")]
    assert len(made) == 3
    for p in made:
        assert "outputs of /root/reference/shader" in str(np.load(p)["provenance"])


def post_kwargs(z):
    e, s, m, d = [float(x) for x in z["post"]]
    return dict(exposure=e, saturation=s, max_sigma=m, denoise=bool(d))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_oracle_reproduces_golden(path, oracle_mod):
    z, sa = load(path)
    W, H = int(z["width"]), int(z["height"])
    O = oracle_mod.Oracle(sa)
    pos, d = oracle_mod.camera(W, H, z["eye"], z["dir"], float(z["fov_scale"]), z["lens"], float(z["rand_cam"][0]))
    assert beq(pos, z["cam_pos"]) and beq(d, z["cam_dir"])
    idx, t, cnt, st = O.bvh_test(pos, d)
    assert beq(idx, z["hit_index"]) and beq(t, z["hit_t"]) and beq(cnt, z["hit_count"])
    assert [st["rays"], st["node_visits"], st["leaf_visits"]] == list(z["visits"])
    fb = None
    for k in range(len(z["rand_cam"])):
        pos, d = oracle_mod.camera(W, H, z["eye"], z["dir"], float(z["fov_scale"]), z["lens"], float(z["rand_cam"][k]))
        fb, _ = O.trace(pos, d, W, H, k, float(z["rand_trace"][k]), float(z["env_theta"]), fb_prev=fb)
    assert beq(fb[..., :3], z["accum"][..., :3])
    assert beq(oracle_mod.draw(fb, **post_kwargs(z)), z["rgba8"])


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p) for p in GOLDEN])
def test_cuda_reproduces_golden(path):
    from fspt_b200 import capi
    z, sa = load(path)
    W, H = int(z["width"]), int(z["height"])
    ctx = capi.Context(W, H)
    try:
        ctx.scene_upload(sa)
        fr = ctx.frame(z["eye"], z["dir"], float(z["fov_scale"]), z["lens"], float(z["env_theta"]))
        idx, t, cnt, pos, d = ctx.debug_primary(fr, float(z["rand_cam"][0]))
        assert beq(pos, z["cam_pos"]) and beq(d, z["cam_dir"])
        assert beq(idx, z["hit_index"]) and beq(t, z["hit_t"]) and beq(cnt, z["hit_count"])
        ctx.clear()
        ctx.render(fr, 0, z["rand_cam"], z["rand_trace"])
        assert beq(ctx.read_accum()[..., :3], z["accum"][..., :3])
        assert beq(ctx.resolve(**post_kwargs(z)), z["rgba8"])
    finally:
        ctx.close()
