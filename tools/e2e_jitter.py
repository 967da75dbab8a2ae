"""Per-step wall time of the e2e leg (upload + clear + render + resolve), with and without torch in the process;
--async: fspt_scene_upload_async (the atlas transfer overlaps the primary traversal); --pinned: atlas and environment
page-locked in place (fspt_host_register)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
if "--torch" in sys.argv:
    import torch
    torch.cuda.set_device(0)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    if "--one-thread" in sys.argv:
        torch.set_num_threads(1)
from fspt_b200 import scenes
from fspt_b200.path_tracer import PathTracer
sa, cam = scenes.bunny_class(subdiv=6, atlas_res=2048)
pt = PathTracer(sa, (1280, 720), cam, device=0)
if "--pinned" in sys.argv:
    from fspt_b200 import capi
    pt.ctx.synchronize(); capi.host_register(sa.atlas); capi.host_register(sa.env)
rc, rt = scenes.rand_bases(64, 1)
out8 = np.empty((720, 1280, 4), np.uint8)
ts = []
for i in range(24):
    t0 = time.perf_counter()
    pt.ctx.scene_upload(sa, wait="--async" not in sys.argv)
    t1 = time.perf_counter()
    pt.clear(); pt.ctx.render(pt._frame(), 0, rc, rt); pt.stats()
    t2 = time.perf_counter()
    pt.drawQuad(out8)
    t3 = time.perf_counter()
    ts.append(((t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3))
a = np.array(ts[2:])
print(" ".join(sys.argv[1:]) or "plain", "| upload ms: median %.1f max %.1f | render %.1f max %.1f | resolve %.2f max %.2f | step median %.1f mean %.1f" % (
    np.median(a[:, 0]), a[:, 0].max(), np.median(a[:, 1]), a[:, 1].max(), np.median(a[:, 2]), a[:, 2].max(), np.median(a.sum(1)), a.sum(1).mean()))
print("  uploads:", " ".join("%.1f" % x for x in a[:, 0]))
