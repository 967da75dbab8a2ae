#!/bin/bash
# compute-sanitizer pass over one small frame of every kernel (run on the GPU box: gpurun -- bash tools/sanitize.sh)
set -e
cat > /tmp/fspt_san.py <<PY
import sys; sys.path.insert(0, ".")
from fspt_b200 import scenes, capi
sa, cam = scenes.pbr_scene(atlas_res=64, subdiv=2, env_size=(128, 64))
ctx = capi.Context(64, 48)
ctx.scene_upload(sa)
fr = ctx.frame(cam["eye"], cam["dir"], cam["fov_scale"], scenes.lens_features(cam), cam["env_theta"])
rc, rt = scenes.rand_bases(3, 1)
ctx.clear(); ctx.render(fr, 0, rc, rt); img = ctx.resolve(denoise=True)
ctx.scene_upload(sa, wait=False)   # asynchronous upload: atlas thread + ring reuse (FSPT_RING_SLOTS=2 below)
ctx.clear(); ctx.render(fr, 0, rc, rt); img2 = ctx.resolve(denoise=True)
assert (img == img2).all()
idx, t, cnt, pos, d = ctx.debug_primary(fr, 5.0)
print("ok", img.mean(), (idx >= 0).mean(), ctx.stats()["rays"])
ctx.close()
PY
for tool in memcheck racecheck initcheck; do
  echo "== $tool"; FSPT_RING_SLOTS=2 compute-sanitizer --tool $tool --error-exitcode 3 python /tmp/fspt_san.py 2>&1 | tail -2
done
