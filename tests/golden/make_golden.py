#!/usr/bin/env python
"""Generates tests/golden/*.npz from OUTPUTS OF THE REFERENCE ITSELF: camera.fs / bvh_test.fs / tracer.fs / draw.fs of
/root/reference/shader, compiled for the CPU by `make -C oracle ref` (oracle/glsl_cpu/, oracle/reference_shaders.py)
and run here, where the reference tree exists.  The reference ships no golden vectors of its own.  The oracle
restatement must give the same bits (asserted below before anything is written), and the fixtures then travel: on the
GPU box, where neither the reference nor a compiler for it is guaranteed, tests/test_golden.py checks the oracle and
the CUDA path against them.  Inputs (scene arrays in the reference's layout, camera, seeds) are stored next to the
outputs so nothing is regenerated at test time.  Visit statistics (rays / V / L) come from the oracle's counters:
the shaders do not export them, their per-ray `count` (bvh_test.fs:184) is compared instead.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from oracle import reference_shaders as ref  # noqa: E402
from fspt_b200 import scenes  # noqa: E402

PROVENANCE = ("cam_pos cam_dir hit_index hit_t hit_count accum rgba8 = outputs of /root/reference/shader/*.fs run on the CPU "
              "through oracle/glsl_cpu (oracle/_ref/libfspt_ref.so); the oracle restatement reproduced every one bit for bit "
              "when this file was written")


def beq(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.dtype.kind == "f":
        return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))
    return bool(np.array_equal(a, b))


def make(name, sa, cam, W, H, n_samples, seed, post):
    O, R = oracle.Oracle(sa), ref.Reference(sa)
    rc, rt = scenes.rand_bases(n_samples, seed)
    lens = np.asarray(scenes.lens_features(cam), np.float32)
    pos0, dir0 = ref.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], lens, rc[0])
    assert all(beq(a, b) for a, b in zip((pos0, dir0), oracle.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], lens, rc[0])))
    idx, t, cnt = R.bvh_test(pos0, dir0)
    o_idx, o_t, o_cnt, st = O.bvh_test(pos0, dir0)
    assert beq(idx, o_idx) and beq(t, o_t) and beq(cnt, o_cnt)
    fb = None
    for k in range(n_samples):
        pos, d = ref.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], lens, rc[k])
        o_fb, _ = O.trace(pos, d, W, H, k, rt[k], cam["env_theta"], fb_prev=fb, sanitize=0)
        fb = R.trace(pos, d, W, H, k, rt[k], cam["env_theta"], fb_prev=fb)
        assert beq(fb, o_fb)
    rgba = ref.draw(fb, **post)
    assert beq(rgba, oracle.draw(fb, **post))
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        bvh=sa.bvh, tris=sa.tris, mats=sa.mats, norms=sa.norms, uvs=sa.uvs, atlas=sa.atlas, env=sa.env, bins=sa.bins,
        eye=np.asarray(cam["eye"], np.float32), dir=np.asarray(cam["dir"], np.float32),
        fov_scale=np.float32(cam["fov_scale"]), env_theta=np.float32(cam["env_theta"]), lens=lens,
        width=W, height=H, rand_cam=rc, rand_trace=rt,
        cam_pos=pos0, cam_dir=dir0, hit_index=idx, hit_t=t, hit_count=cnt,
        visits=np.array([st["rays"], st["node_visits"], st["leaf_visits"]], np.int64),
        accum=fb, rgba8=rgba, provenance=np.array(PROVENANCE),
        post=np.array([post["exposure"], post["saturation"], post["max_sigma"], 1.0 if post["denoise"] else 0.0], np.float32))
    print(name, "tris", sa.n_tris, "hit frac %.2f" % (idx >= 0).mean(), "mean", fb[..., :3].mean())


if __name__ == "__main__":
    sa, cam = scenes.bunny_class(subdiv=2, atlas_res=16, env_size=(64, 32))
    make("bunny_small", sa, cam, 48, 32, 3, 17, dict(exposure=1.0, saturation=1.0, max_sigma=2.0, denoise=False))
    sa, cam = scenes.pbr_scene(atlas_res=16, subdiv=1, env_size=(64, 32))
    make("pbr_refractive_small", sa, cam, 40, 24, 3, 23, dict(exposure=1.3, saturation=0.8, max_sigma=2.0, denoise=True))
    sa, cam = scenes.quad_scene()
    make("quad_kat", sa, cam, 16, 16, 2, 5, dict(exposure=1.0, saturation=1.0, max_sigma=2.0, denoise=False))
