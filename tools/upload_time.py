import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fspt_b200 import scenes, capi
ASYNC = "--async" in sys.argv   # fspt_scene_upload_async: time to return, then the time until everything has landed
sa, cam = scenes.sphere_soup() if "--soup" in sys.argv else scenes.bunny_class(subdiv=6, atlas_res=2048)
ctx = capi.Context(1280, 720)
if "--pinned" in sys.argv:
    capi.host_register(sa.atlas); capi.host_register(sa.env)
for i in range(4):
    t0 = time.perf_counter(); n = ctx.scene_upload(sa, wait=not ASYNC); t1 = time.perf_counter(); ctx.synchronize(); dt = time.perf_counter() - t0
    print("upload returned after %.2f ms, landed after %.1f ms for %.1f MB -> %.1f GB/s" % ((t1 - t0) * 1e3, dt * 1e3, n / 1e6, n / dt / 1e9))
import numpy as np
out = np.empty((720, 1280, 4), np.uint8)
rc, rt = scenes.rand_bases(1, 1)
ctx.render(ctx.frame(cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), cam["env_theta"]), 0, rc, rt)
for i in range(3):
    t0 = time.perf_counter(); ctx.resolve(out=out); dt = time.perf_counter() - t0
    print("resolve+readback %.2f ms" % (dt * 1e3))
