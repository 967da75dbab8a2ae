"""The reference's own JavaScript scene compilers, executed -- TEST INFRASTRUCTURE, never imported by fspt_b200/.

bvh.js (+ vector.js), env_sampler.js, texture_packer.js, obj_loader.js (+ mtl_loader.js) of /root/reference are
loaded as the ES modules they are into Qt's ECMAScript engine (oracle/js_engine.py) and driven from Python, so that the
restatements (oracle/fspt_oracle_host.cpp) and the product's native / numpy versions can be compared with what the
reference's code really computes.  Only where the reference tree exists (this container).

What is NOT the reference's code here, and why:
  * the few lines of main.js around each call (main.js cannot be imported: it is one DOM + WebGL closure): the loop
    that flattens serializeTree() into the node buffer and maskBVHBuffer (main.js:272-282, 352-392) are restated in the
    driver scripts below, next to the line numbers they follow;
  * browser objects the modules touch: `document.createElement('canvas')` + `getImageData` of env_sampler.js:55-61
    (stand-in: the image's RGBA bytes as they are), `console`;
  * obj_loader.js is `async` for one `await Utility.getText(...)` and this engine is ES2016: the module is loaded from a
    copy under oracle/_ref/js/ with the two keywords removed and utility.js replaced by a table lookup; its closing
    `new Promise(resolve => ...)` is resolved synchronously.  Nothing else of any module is touched.
"""
import json
import os
import re

import numpy as np

from . import js_engine

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get("FSPT_REFERENCE_ROOT", "/root/reference")
_JS_DIR = os.path.join(_HERE, "_ref", "js")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "bvh.js")) and js_engine.available()


_eng = None

# Array.prototype.sort must be stable (ECMAScript 2019, section 22.1.3.27; V8 since Chrome 70, i.e. every browser that
# runs the reference's WebGL2 code today).  This engine predates that and is not (probed: tests/test_reference_js_pin.py),
# and bvh.js:80-91 sorts centroids that tie on every symmetric mesh, so a stable merge sort is installed in its
# place.  A stable sort's result is unique for a consistent comparator, so which stable algorithm does not matter.
_STABLE_SORT = """
(function () {
  var probe = [];
  for (var i = 0; i < 64; i++) probe.push([i % 3, i]);
  probe.sort(function (a, b) { return a[0] - b[0]; });
  var stable = true;
  for (var i = 1; i < probe.length; i++) if (probe[i][0] === probe[i - 1][0] && probe[i][1] < probe[i - 1][1]) stable = false;
  this.FSPT_NATIVE_SORT_STABLE = stable;
  if (stable) return;
  Object.defineProperty(Array.prototype, "sort", { configurable: true, writable: true, enumerable: false,
    value: function (cmp) {
      if (cmp === undefined) cmp = function (a, b) { var x = String(a), y = String(b); return x < y ? -1 : (x > y ? 1 : 0); };
      var n = this.length, src = this.slice(0, n), dst = new Array(n);
      for (var w = 1; w < n; w *= 2) {
        for (var lo = 0; lo < n; lo += 2 * w) {
          var mid = Math.min(lo + w, n), hi = Math.min(lo + 2 * w, n), i = lo, j = mid, k = lo;
          while (i < mid && j < hi) dst[k++] = cmp(src[j], src[i]) < 0 ? src[j++] : src[i++];
          while (i < mid) dst[k++] = src[i++];
          while (j < hi) dst[k++] = src[j++];
        }
        var t = src; src = dst; dst = t;
      }
      for (var i = 0; i < n; i++) this[i] = src[i];
      return this;
    } });
}).call(this)
"""


def _prepare_copies():
    """oracle/_ref/js/: the modules obj_loader.js needs, next to each other (ES imports are relative)."""
    os.makedirs(_JS_DIR, exist_ok=True)
    for name in ("bvh.js", "vector.js", "mtl_loader.js", "obj_loader.js"):
        src = open(os.path.join(REFERENCE_ROOT, name)).read()
        if name == "obj_loader.js":
            n_async, n_await = len(re.findall(r"\basync\s+function\b", src)), len(re.findall(r"\bawait\s+", src))
            assert (n_async, n_await) == (1, 1), "obj_loader.js changed: expected one async function with one await"
            src = re.sub(r"\basync\s+function\b", "function", src)
            src = re.sub(r"\bawait\s+", "", src)
        with open(os.path.join(_JS_DIR, name), "w") as f:
            f.write(src)
    with open(os.path.join(_JS_DIR, "utility.js"), "w") as f:  # stand-in for the XHR helpers: texts come from a table
        f.write("export function getText(path) {\n  if (!(path in FSPT_TEXTS)) throw new Error('no text for ' + path);\n"
                "  return FSPT_TEXTS[path];\n}\n")


def engine():
    global _eng
    if _eng is None:
        e = js_engine.JSEngine()
        e.evaluate(_STABLE_SORT)
        e.import_module(os.path.join(REFERENCE_ROOT, "bvh.js"), "REF_BVH")            # unmodified, from where it lies
        e.import_module(os.path.join(REFERENCE_ROOT, "env_sampler.js"), "REF_ENV")
        e.import_module(os.path.join(REFERENCE_ROOT, "texture_packer.js"), "REF_PACKER")
        _prepare_copies()
        e.evaluate("var FSPT_TEXTS = {};")
        e.import_module(os.path.join(_JS_DIR, "obj_loader.js"), "REF_OBJ")
        e.import_module(os.path.join(_JS_DIR, "mtl_loader.js"), "REF_MTL")   # byte-identical copy (it imports ./utility.js)
        _eng = e
    return _eng


def _js(x):
    """numpy / python -> a JS literal that parses back to the same doubles (repr round-trips)."""
    if isinstance(x, np.ndarray):
        x = x.tolist()
    return json.dumps(x)


def bvh_build(verts, max_tris=4, box_verts=None):
    """new BVH(geometry, leafSize) + serializeTree() (bvh.js) and the node-buffer loop of main.js:352-392 with
    maskBVHBuffer (main.js:272-282).  verts: (T,3,3) f64.  box_verts: the vertices each Triangle was CONSTRUCTED with,
    when scene.normalize rescaled `verts` afterwards and left Triangle.boundingBox stale (main.js:335-347).
    Returns (nodes[N,9] f32 with int bits in [0..2], order[T] i32 = input index of each emitted triangle, depth)."""
    verts = np.asarray(verts, np.float64).reshape(-1, 3, 3)
    bv = None if box_verts is None else np.asarray(box_verts, np.float64).reshape(-1, 3, 3)
    src = """
(function () {
  let V = %s, BV = %s;
  let geometry = V.map((v, i) => {
    let t = new REF_BVH.Triangle(BV ? BV[i] : v, null, null, null);
    t.verts = v;          // scene.normalize rewrites verts after construction (main.js:341-345)
    t.fsptId = i;
    return t;
  });
  let bvh = new REF_BVH.BVH(geometry, %d);                  // main.js:352
  let bvhArray = bvh.serializeTree();                       // main.js:355
  let bvhBuffer = [], order = [];
  for (let i = 0; i < bvhArray.length; i++) {               // main.js:363-392
    let e = bvhArray[i];
    let node = e.node;
    let triIndex = node.leaf ? order.length : -1;           // = trianglesBuffer.length / 3 / 3
    let bufferNode = [e.left, e.right, triIndex].concat(node.boundingBox.min, node.boundingBox.max);
    if (node.leaf) {
      let tris = node.getTriangles();
      for (let j = 0; j < tris.length; j++) order.push(tris[j].fsptId);
    }
    for (let j = 0; j < bufferNode.length; j++) bvhBuffer.push(bufferNode[j]);
  }
  let masked = new Float32Array(new Int32Array(bvhBuffer).buffer);   // maskBVHBuffer, main.js:272-282
  for (let i = 0; i < bvhBuffer.length; i += 9)
    for (let j = 3; j < 9; j++) masked[i + j] = bvhBuffer[i + j];
  return JSON.stringify({nodes: Array.from(new Uint32Array(masked.buffer)), order: order, depth: bvh.depth});
})()""" % (_js(verts), "null" if bv is None else _js(bv), int(max_tris))
    out = json.loads(engine().evaluate(src))
    nodes = np.asarray(out["nodes"], np.uint32).view(np.float32).reshape(-1, 9)
    return nodes, np.asarray(out["order"], np.int32), int(out["depth"])


def env_bins(rgba8):
    """ProcessEnvRadiance(img) (env_sampler.js) on an RGBA8 image (H,W,4), row 0 = top.  Returns (n,4) uint16."""
    rgba8 = np.ascontiguousarray(rgba8, np.uint8)
    H, W = rgba8.shape[0], rgba8.shape[1]
    src = """
(function () {
  let data = new Uint8ClampedArray(%s);
  var document = { createElement: function () { return {               // the canvas of env_sampler.js:55-61
    getContext: function () { return { drawImage: function () {}, getImageData: function () { return {data: data}; } }; } }; } };
  this.document = document;
  let bins = REF_ENV.ProcessEnvRadiance({width: %d, height: %d});
  return JSON.stringify(Array.from(bins));
}).call(this)""" % (_js(rgba8.reshape(-1)), W, H)
    return np.asarray(json.loads(engine().evaluate(src)), np.uint16).reshape(-1, 4)


def packer_indices(ops, atlas_res):
    """TexturePacker.addTexture / addColor / setAndGetResolution (texture_packer.js:5-42) for a sequence of
    ("tex", currentSrc, height, corrected) / ("color", [r, g, b]) operations.  Returns (indices, resolution,
    [kind of every layer])."""
    src = """
(function () {
  let p = new REF_PACKER.TexturePacker(%d);
  let ops = %s, out = [];
  for (let i = 0; i < ops.length; i++) {
    if (ops[i][0] === "tex") out.push(p.addTexture({currentSrc: ops[i][1], height: ops[i][2]}, ops[i][3]));
    else out.push(p.addColor(ops[i][1]));
  }
  let res = p.setAndGetResolution();
  return JSON.stringify({idx: out, res: res, layers: p.imageSet.map(x => Array.isArray(x) ? x : x.currentSrc)});
})()""" % (int(atlas_res), _js([list(o) for o in ops]))
    out = json.loads(engine().evaluate(src))
    return out["idx"], out["res"], out["layers"]


def parse_mesh(obj_text, transforms, world_transforms=None, base_path="", texts=None):
    """parseMesh(objText, transforms, worldTransforms, basePath) (obj_loader.js:6-215).  texts: {url: text} served to
    Utility.getText (the mtllib).  Returns {groups: {name: {verts, normals, tangents, bitangents, uvs (T,3,*) f64,
    material}}, urls, bounds}."""
    e = engine()
    src = """
(function () {
  FSPT_TEXTS = %s;
  let SavedPromise = Promise, box = {};
  Promise = function (executor) { executor(function (v) { box.value = v; }); };   // the closing `new Promise` (obj_loader.js:214)
  let r;
  try { REF_OBJ.parseMesh(%s, %s, %s, %s); r = box.value; } finally { Promise = SavedPromise; }
  let groups = {};
  Object.keys(r.groups).forEach(function (k) {
    let g = r.groups[k];
    groups[k] = { material: g.material,
      verts: g.triangles.map(t => t.verts), normals: g.triangles.map(t => t.normals),
      tangents: g.triangles.map(t => t.tangents.slice(0, 3)), bitangents: g.triangles.map(t => t.bitangents.slice(0, 3)),
      n_tangents: g.triangles.map(t => t.tangents.length), uvs: g.triangles.map(t => t.uvs) };
  });
  function enc(x) { return JSON.stringify(x, function (k, v) {
    return (typeof v === "number" && !isFinite(v)) ? (isNaN(v) ? "NaN" : (v > 0 ? "Infinity" : "-Infinity")) : v; }); }
  return enc({groups: groups, order: Object.keys(r.groups), urls: r.urls ? Array.from(r.urls) : null, bounds: r.bounds});
}).call(this)""" % (_js(texts or {}), _js(obj_text), _js(transforms), _js(world_transforms), _js(base_path))

    def dec(v):
        if isinstance(v, list):
            return [dec(x) for x in v]
        if isinstance(v, dict):
            return {k: dec(x) for k, x in v.items()}
        if v in ("NaN", "Infinity", "-Infinity"):
            return float(v.replace("Infinity", "inf"))
        return v
    out = dec(json.loads(e.evaluate(src)))
    for g in out["groups"].values():
        for k in ("verts", "normals", "tangents", "bitangents", "uvs"):
            g[k] = np.asarray(g[k], np.float64)
    return out


def parse_materials(mtl_text, base_path):
    """ParseMaterials(mtlText, basePath) (mtl_loader.js:3-40).  Returns (materials dict, sorted url list)."""
    out = json.loads(engine().evaluate("(function () { let r = REF_MTL.ParseMaterials(%s, %s); "
                                       "return JSON.stringify({materials: r.materials, urls: Array.from(r.urls)}); })()"
                                       % (_js(mtl_text), _js(base_path))))
    return out["materials"], sorted(out["urls"])
