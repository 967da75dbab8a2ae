#!/usr/bin/env python
"""SwiftShader pin for the oracle (SURVEY.md 8c pin 7) -- TEST INFRASTRUCTURE.

Runs the UNMODIFIED reference (index.html's DOM + main.js + shader/*.fs, served from the read-only reference tree) in
headless Chromium on SwiftShader software WebGL2, with
  * inject.js (this directory) loaded ahead of the reference's modules: seeded Math.random, float read-back of the
    camera targets and of the accumulation target after every pass, optional timing;
  * `mode=test` swapping tracer.fs for shader/bvh_test.fs (main.js:882-884), whose last two lines (:230-231, which write
    the visit-count heat map) are replaced ON THE FLY by the server with a tail that writes what the shader computed:
    `fragColor = vec4(float(result.index), result.t, float(count), 1.0)`;
  * the four entries of .MISSING_LARGE_BLOBS substituted by procedural assets (the same ones
    tests/golden/make_bunny_json_fixture.py uses) so that scene/bunny.json loads.
and compares what comes back with the CPU oracle on the same scene, camera and rand bases:
  camera pos/dir and (index, t, count): bit-exact expected up to the platform's sin/sqrt/div (reported as mismatch
  counts and max ulp); radiance: per-pixel statistics (the sin-hash RNG is platform-dependent).

    python oracle/swiftshader/run_harness.py --reference /root/reference [--browser /path/to/chromium] \\
        [--res 160x96] [--ticks 2] [--mode test|trace|timing] [--out gpurun_out/swiftshader.npz]

Needs a Chromium-family browser; this image has none (probed: chromium chromium-browser google-chrome chrome
headless_shell) and there is no network, so the script exits 3 with that message here.  Nothing under tests/, smoke()
or the default bench path depends on it; `bench.py --impl reference` calls time_reference() when a browser and the
reference tree are both present at run time.
"""
import argparse
import http.server
import io
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile
import threading
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

BROWSERS = ("chromium", "chromium-browser", "google-chrome", "google-chrome-stable", "chrome", "headless_shell")
BVH_TEST_TAIL_OLD = ("  vec3 color = vec3(float(count) * 0.001);\n"
                     "  fragColor = vec4((color + (tcolor * float(tick)))/(float(tick)+1.0),1.0);")
BVH_TEST_TAIL_NEW = "  fragColor = vec4(float(result.index), result.t, float(count), 1.0);"


def find_browser(explicit=None):
    if explicit:
        return explicit if os.path.exists(explicit) or shutil.which(explicit) else None
    for n in BROWSERS:
        p = shutil.which(n)
        if p:
            return p
    return os.environ.get("FSPT_BROWSER")


def harness_page(ref_root):
    """index.html of the reference with inject.js ahead of its modules (the DOM ids main.js reads stay untouched)."""
    html = open(os.path.join(ref_root, "index.html")).read()
    assert "<head>" in html
    return html.replace("<head>", "<head>\n    <script src=\"/__harness__/inject.js\"></script>", 1)


def substitutes(atlas_res):
    """The four .MISSING_LARGE_BLOBS entries -> bytes (OBJ text / PNG files), same content as the golden fixture."""
    from PIL import Image
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_bunny_json_fixture as mk
    from fspt_b200 import procedural as pr

    def png(arr):
        b = io.BytesIO()
        Image.fromarray(arr, "RGBA").save(b, "PNG")
        return b.getvalue()
    return {
        "asset_packs/misc/bunny_big.obj": mk.substitute_obj().encode(),
        "environment/autumn_meadow_2k.RGBE.PNG": png(pr.environment(256, 128)),
        "asset_packs/dungeon/RootNode_normal.png": png(pr.pbr_maps(atlas_res, 7, "A")["normal"]["pixels"]),
        "asset_packs/dungeon/Scene_-_Root_normal.png": png(pr.pbr_maps(atlas_res, 11, "B")["normal"]["pixels"]),
    }


class Handler(http.server.SimpleHTTPRequestHandler):
    server_version = "fspt-harness"

    def log_message(self, *a):
        pass

    def do_GET(self):
        S = self.server
        path = self.path.split("?")[0].lstrip("/")
        body, ctype = None, "application/octet-stream"
        if path in ("", "harness.html"):
            body, ctype = harness_page(S.ref_root).encode(), "text/html"
        elif path == "__harness__/inject.js":
            body, ctype = open(os.path.join(HERE, "inject.js"), "rb").read(), "text/javascript"
        elif path == "shader/bvh_test.fs":
            src = open(os.path.join(S.ref_root, path)).read()
            if BVH_TEST_TAIL_OLD not in src:
                self.send_error(500, "bvh_test.fs tail not found: the reference changed")
                return
            body, ctype = src.replace(BVH_TEST_TAIL_OLD, BVH_TEST_TAIL_NEW).encode(), "text/plain"
        elif path == "scene/fspt_harness.json":
            scene = json.load(open(os.path.join(S.ref_root, "scene", "bunny.json")))
            scene["samples"] = S.ticks + 1          # tick() renders while pingpong <= max (main.js:841)
            scene["atlasRes"] = S.atlas_res
            body, ctype = json.dumps(scene).encode(), "application/json"
        elif path in S.subs:
            body = S.subs[path]
        if body is None:
            full = os.path.normpath(os.path.join(S.ref_root, path))
            if not full.startswith(os.path.abspath(S.ref_root)) or not os.path.isfile(full):
                self.send_error(404)
                return
            body = open(full, "rb").read()
            ctype = {".js": "text/javascript", ".html": "text/html", ".css": "text/css", ".json": "application/json",
                     ".png": "image/png", ".jpeg": "image/jpeg", ".jpg": "image/jpeg"}.get(os.path.splitext(full)[1].lower(), "text/plain")
        self.send_response(200)
        self.send_header("Content-Type", ctype)
        self.send_header("Content-Length", str(len(body)))
        self.end_headers()
        self.wfile.write(body)

    def do_POST(self):
        n = int(self.headers.get("Content-Length", "0"))
        data = self.rfile.read(n)
        if self.path.startswith("/result"):
            self.server.result = data
        else:
            self.server.errors.append(data.decode(errors="replace"))
        self.send_response(200)
        self.send_header("Content-Length", "0")
        self.end_headers()


def run_browser(browser, ref_root, W, H, ticks, mode, seed=1, atlas_res=128, timeout=1800):
    """Serve, launch, wait for the POST.  Returns the decoded result dict."""
    import numpy as np
    srv = http.server.ThreadingHTTPServer(("127.0.0.1", 0), Handler)
    srv.ref_root, srv.ticks, srv.atlas_res = os.path.abspath(ref_root), ticks, atlas_res
    srv.subs, srv.result, srv.errors = substitutes(atlas_res), None, []
    threading.Thread(target=srv.serve_forever, daemon=True).start()
    url = "http://127.0.0.1:%d/harness.html?scene=fspt_harness&res=%dx%d&seed=%d&ticks=%d%s%s" % (
        srv.server_address[1], W, H, seed, ticks, "&mode=test" if mode == "test" else "", "&timing=1" if mode == "timing" else "")
    prof = tempfile.mkdtemp(prefix="fspt_chromium_")
    cmd = [browser, "--headless=new", "--use-angle=swiftshader", "--enable-unsafe-swiftshader", "--ignore-gpu-blocklist",
           "--no-sandbox", "--disable-gpu-sandbox", "--user-data-dir=" + prof, "--window-size=%d,%d" % (W, H), url]
    p = subprocess.Popen(cmd, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    try:
        t0 = time.time()
        while srv.result is None and time.time() - t0 < timeout and p.poll() is None:
            time.sleep(0.2)
    finally:
        p.terminate()
        srv.shutdown()
        shutil.rmtree(prof, ignore_errors=True)
    if srv.result is None:
        raise RuntimeError("no result from the browser (%s)" % ("; ".join(srv.errors) or "timeout / browser exited"))
    raw = srv.result
    hlen, n = struct.unpack_from("<II", raw, 0)
    hdr = json.loads(raw[8:8 + hlen].decode())
    w, h = hdr["width"], hdr["height"]
    arr = np.frombuffer(raw, np.float32, count=n * 3 * w * h * 4, offset=8 + hlen).reshape(n, 3, h, w, 4) if n else None
    hdr["arrays"] = arr
    hdr["errors"] = srv.errors
    return hdr


def time_reference(browser, ref_root, W, H, steps, warmup):
    """bench.py --impl reference: gl.finish-bracketed drawCamera + drawTracer passes, no vsync cap."""
    r = run_browser(browser, ref_root, W, H, steps + warmup, "timing", atlas_res=2048)
    ms = r["tick_ms"][warmup:warmup + steps]
    return {"seconds": sum(ms) / 1e3, "renderer": r.get("renderer"), "user_agent": r.get("user_agent")}


def compare(res, W, H, mode):
    """Browser output vs the CPU oracle on the same scene / camera / rand bases.  Prints a report, returns a dict."""
    import numpy as np
    import oracle
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_bunny_json_fixture as mk
    from fspt_b200 import scene_json, scenes
    oracle.build()
    sa, cam = mk.compile_bunny_json()
    dist = scene_json.autofocus_distance(sa.verts64, cam["eye"], cam["dir"])
    cam = dict(cam, focal_depth=dist)
    O = oracle.Oracle(sa)
    lens = scenes.lens_features(cam)
    rand = np.asarray(res["rand"], np.float64) * 10000.0
    # main.js discards its first pass (clear() after the first tick): captured pass k used draws 2(k+1), 2(k+1)+1
    out = {"passes": int(res["passes"]), "renderer": res.get("renderer")}
    arr = res["arrays"]

    def ulp(a, b):
        ia, ib = a.view(np.int32).astype(np.int64), b.view(np.int32).astype(np.int64)
        return np.abs(ia - ib)
    fb = None
    for k in range(arr.shape[0]):
        rc, rt = np.float32(rand[2 * (k + 1)]), np.float32(rand[2 * (k + 1) + 1])
        pos, d = oracle.camera(W, H, cam["eye"], cam["dir"], cam["fov_scale"], lens, rc)
        bpos, bdir, bacc = arr[k, 0], arr[k, 1], arr[k, 2]
        out["pass%d_cam_pos_max_ulp" % k] = int(ulp(np.ascontiguousarray(bpos[..., :3]), np.ascontiguousarray(pos[..., :3])).max())
        out["pass%d_cam_dir_max_ulp" % k] = int(ulp(np.ascontiguousarray(bdir[..., :3]), np.ascontiguousarray(d[..., :3])).max())
        if mode == "test":
            # traverse the BROWSER's rays so that platform differences in camera.fs do not leak into the hit records
            idx, t, cnt, _ = O.bvh_test(np.ascontiguousarray(bpos), np.ascontiguousarray(bdir))
            bi, bt, bc = bacc[..., 0].astype(np.int32).ravel(), bacc[..., 1].ravel(), bacc[..., 2].astype(np.int32).ravel()
            out["pass%d_index_mismatch" % k] = int((bi != idx).sum())
            out["pass%d_count_mismatch" % k] = int((bc != cnt).sum())
            out["pass%d_t_max_ulp" % k] = int(ulp(np.ascontiguousarray(bt), np.ascontiguousarray(t)).max())
            out["pixels"] = int(idx.size)
        else:
            fb, _ = O.trace(pos, d, W, H, k, rt, cam["env_theta"], fb_prev=fb)
            diff = bacc[..., :3] - fb[..., :3]
            out["pass%d_radiance_rmse" % k] = float(np.sqrt((diff ** 2).mean()))
            out["pass%d_radiance_mean_browser" % k] = float(bacc[..., :3].mean())
            out["pass%d_radiance_mean_oracle" % k] = float(fb[..., :3].mean())
    print(json.dumps(out, indent=1))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default=os.environ.get("FSPT_REFERENCE_ROOT", "/root/reference"))
    ap.add_argument("--browser", default=None)
    ap.add_argument("--res", default="160x96")
    ap.add_argument("--ticks", type=int, default=2)
    ap.add_argument("--mode", default="test", choices=["test", "trace", "timing"])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    W, H = [int(x) for x in a.res.split("x")]
    browser = find_browser(a.browser)
    if not browser:
        print("no Chromium-family browser found (%s); install one or pass --browser.  This image has none and no network: "
              "the pin that exists is the shaders compiled for the CPU, oracle/glsl_cpu (DESIGN.md section 2)." % ", ".join(BROWSERS))
        return 3
    if not os.path.isfile(os.path.join(a.reference, "main.js")):
        print("reference tree not found at %s" % a.reference)
        return 3
    if a.mode == "timing":
        print(json.dumps(time_reference(browser, a.reference, W, H, a.ticks, 1)))
        return 0
    res = run_browser(browser, a.reference, W, H, a.ticks, a.mode)
    if a.out:
        import numpy as np
        np.savez_compressed(a.out, arrays=res["arrays"], rand=np.asarray(res["rand"]), width=W, height=H)
    compare(res, W, H, a.mode)
    return 0


if __name__ == "__main__":
    sys.exit(main())
