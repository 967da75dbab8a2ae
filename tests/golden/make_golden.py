#!/usr/bin/env python
"""Generates tests/golden/*.npz with the CPU oracle (oracle/).  The reference ships no golden vectors and cannot
run in this image, so these pin the ORACLE's outputs (and the scene compilers' outputs) across compilers,
platforms and refactors; the inputs travel inside the fixture so nothing is regenerated at test time.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle  # noqa: E402
from fspt_b200 import scenes  # noqa: E402


def make(name, sa, cam, W, H, n_samples, seed, post):
    O = oracle.Oracle(sa)
    rc, rt = scenes.rand_bases(n_samples, seed)
    lens = np.asarray(scenes.lens_features(cam), np.float32)
    pos0, dir0 = oracle.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], lens, rc[0])
    idx, t, cnt, st = O.bvh_test(pos0, dir0)
    fb = None
    for k in range(n_samples):
        pos, d = oracle.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], lens, rc[k])
        fb, st2 = O.trace(pos, d, W, H, k, rt[k], cam["env_theta"], fb_prev=fb)
    rgba = oracle.draw(fb, **post)
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        bvh=sa.bvh, tris=sa.tris, mats=sa.mats, norms=sa.norms, uvs=sa.uvs, atlas=sa.atlas, env=sa.env, bins=sa.bins,
        eye=np.asarray(cam["eye"], np.float32), dir=np.asarray(cam["dir"], np.float32),
        fov_scale=np.float32(cam["fov_scale"]), env_theta=np.float32(cam["env_theta"]), lens=lens,
        width=W, height=H, rand_cam=rc, rand_trace=rt,
        cam_pos=pos0, cam_dir=dir0, hit_index=idx, hit_t=t, hit_count=cnt,
        visits=np.array([st["rays"], st["node_visits"], st["leaf_visits"]], np.int64),
        accum=fb, rgba8=rgba,
        post=np.array([post["exposure"], post["saturation"], post["max_sigma"], 1.0 if post["denoise"] else 0.0], np.float32))
    print(name, "tris", sa.n_tris, "hit frac %.2f" % (idx >= 0).mean(), "mean", fb[..., :3].mean())


if __name__ == "__main__":
    sa, cam = scenes.bunny_class(subdiv=2, atlas_res=16, env_size=(64, 32))
    make("bunny_small", sa, cam, 48, 32, 3, 17, dict(exposure=1.0, saturation=1.0, max_sigma=2.0, denoise=False))
    sa, cam = scenes.pbr_scene(atlas_res=16, subdiv=1, env_size=(64, 32))
    make("pbr_refractive_small", sa, cam, 40, 24, 3, 23, dict(exposure=1.3, saturation=0.8, max_sigma=2.0, denoise=True))
    sa, cam = scenes.quad_scene()
    make("quad_kat", sa, cam, 16, 16, 2, 5, dict(exposure=1.0, saturation=1.0, max_sigma=2.0, denoise=False))
