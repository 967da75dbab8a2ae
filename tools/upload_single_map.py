"""Upload time of a scene whose textured materials have ONE image map each (the common FSPT case: diffuse texture +
colour defaults), host-side vs GPU-side atlas interleave (FSPT_ATLAS_INTERLEAVE)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fspt_b200 import scenes, capi, procedural as pr

assets = {}
props = []
for i in range(6):
    assets["T%d" % i] = pr.pbr_maps(2048, 3 + i, "T%d" % i)["baseColor"]
    props.append(dict(mesh=(pr.QUAD_VERTS, pr.QUAD_FACES, pr.QUAD_FACE_UVS), scale=2, rotate=[], translate=[i - 3, 0, 0],
                      emittance=[0, 0, 0], normals="flat", diffuse="T%d" % i))
sa = scenes.compile_props(props, assets, 2048, scenes._env((512, 256)))
print("atlas layers", sa.atlas.shape)
for mode in ("cpu", "gpu", None):
    if mode:
        os.environ["FSPT_ATLAS_INTERLEAVE"] = mode
    else:
        os.environ.pop("FSPT_ATLAS_INTERLEAVE", None)
    ctx = capi.Context(640, 360)
    best = 1e9
    for i in range(4):
        t0 = time.perf_counter(); n = ctx.scene_upload(sa); ctx.synchronize(); best = min(best, time.perf_counter() - t0)
    print("interleave=%s: upload %.1f ms for %.1f MB" % (mode or "auto", best * 1e3, n / 1e6))
    ctx.close()
