"""The reference's own JavaScript scene compilers, executed by a real ECMAScript engine, against the repo's versions.

oracle/reference_js.py loads /root/reference/{bvh,vector,env_sampler,texture_packer,obj_loader,mtl_loader}.js as the ES
modules they are into Qt 6's QJSEngine (shipped inside Nsight Compute's host directory; driven through ctypes,
oracle/js_engine.py).  Compared bit for bit:
  * bvh.js  vs  oracle/fspt_oracle_host.cpp (the literal restatement)  vs  the product's native builder (C ABI);
  * env_sampler.js  vs  the same two;
  * texture_packer.js's dedup / index / resolution rules  vs  fspt_b200/texture_packer.py;
  * obj_loader.js + mtl_loader.js  vs  fspt_b200/scene_json.py + geometry.py (numpy, float64).
Needs the reference tree and Qt's libraries, i.e. this container; skipped elsewhere.
"""
import numpy as np
import pytest

from fspt_b200 import capi, procedural as pr, scene_json as SJ
from fspt_b200.texture_packer import TexturePacker
from oracle import reference_js as J

pytestmark = pytest.mark.skipif(not J.available(), reason="no reference tree or no Qt QJSEngine in this image")


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype == np.float64:
        return bool(np.array_equal(a.view(np.uint64), b.view(np.uint64)))
    if a.dtype == np.float32:
        return bool(np.array_equal(a.view(np.uint32), b.view(np.uint32)))
    return bool(np.array_equal(a, b))


def test_engine_is_a_real_one_and_its_sort_needed_the_es2019_fix():
    e = J.engine()
    assert e.evaluate("0.1 + 0.2") == "0.30000000000000004"
    assert e.evaluate("Math.pow(2, -1074) > 0 && Number.EPSILON === Math.pow(2, -52)") == "true"
    assert e.evaluate("Object.keys(REF_BVH).sort().join()") == "BVH,BoundingBox,Node,Triangle"
    # stable after the fix of oracle/reference_js.py (ES2019); the flag records what the engine did natively
    assert e.evaluate("(function(){let a=[];for(let i=0;i<999;i++)a.push([i%5,i]);a.sort((x,y)=>x[0]-y[0]);"
                      "for(let i=1;i<a.length;i++)if(a[i][0]===a[i-1][0]&&a[i][1]<a[i-1][1])return 'unstable';return 'stable'})()") == "stable"
    assert e.evaluate("FSPT_NATIVE_SORT_STABLE") in ("true", "false")


def _grid(n):
    """2 n^2 axis-aligned right triangles: every centroid coordinate ties with many others"""
    g = np.stack(np.meshgrid(np.arange(float(n)), np.arange(float(n)), indexing="ij"), -1).reshape(-1, 2)
    tri = np.zeros((len(g) * 2, 3, 3))
    tri[0::2, :, 0] = g[:, None, 0] + np.array([0, 1, 0]); tri[0::2, :, 1] = g[:, None, 1] + np.array([0, 0, 1])
    tri[1::2, :, 0] = g[:, None, 0] + np.array([1, 1, 0]); tri[1::2, :, 1] = g[:, None, 1] + np.array([0, 1, 1])
    return tri


def _meshes():
    rng = np.random.default_rng(7)
    v2, f2 = pr.icosphere(2)
    v3, f3 = pr.icosphere(3)
    soup = rng.uniform(-1, 1, (1500, 1, 3)) + rng.normal(0, 0.05, (1500, 3, 3))
    grid = _grid(14)
    return {
        "quad_root_leaf": (pr.QUAD_VERTS[np.asarray(pr.QUAD_FACES)], None),
        "five_triangles": (soup[:5], None),
        "icosphere_lumpy": (pr.lumpy(v2)[f2], None),
        "icosphere_symmetric_ties": (v3[f3], None),
        "soup": (soup, None),
        "grid_all_ties": (grid, None),
        "stale_boxes_after_normalize": (grid * 0.37 + 0.1, grid),   # main.js:335-347
        "duplicates": (np.concatenate([soup[:40], soup[:40]]), None),
        "four_point_triangles_one_leaf": (np.zeros((4, 3, 3)), None),
    }


@pytest.mark.parametrize("name", list(_meshes()))
def test_bvh_js(name, oracle_mod):
    verts, box = _meshes()[name]
    rn, ro, rd = J.bvh_build(verts, 4, box_verts=box)
    on, oo, od = oracle_mod.bvh_build(verts, 4, box_verts=box)
    assert bits_equal(on, rn) and bits_equal(oo, ro) and od == rd, "restatement (oracle/fspt_oracle_host.cpp)"
    pn, po, pd = capi.bvh_build(verts, 4, box_verts=box)
    assert bits_equal(pn, rn) and bits_equal(po, ro) and pd == rd, "product (fspt_b200/csrc/bvh_builder.cpp)"
    assert sorted(ro.tolist()) == list(range(len(ro)))


@pytest.mark.parametrize("verts", [np.zeros((8, 3, 3)), np.ones((5, 3, 3)),
                                   np.concatenate([np.random.default_rng(1).normal(size=(40, 3, 3)), np.zeros((8, 3, 3))])],
                         ids=["zeros", "one_point", "soup_plus_zeros"])
def test_bvh_js_crashes_exactly_where_the_builders_refuse(verts, oracle_mod):
    """A node whose triangles all have zero surface area gives NaN costs, no split is chosen and bvh.js dereferences an
    undefined child list (bvh.js:186-196, :22): the reference throws; the restatement and the product report it."""
    from oracle.js_engine import JSError
    with pytest.raises(JSError):
        J.bvh_build(verts, 4)
    with pytest.raises(RuntimeError):
        oracle_mod.bvh_build(verts, 4)
    with pytest.raises(capi.FsptError):
        capi.bvh_build(verts, 4)


@pytest.mark.parametrize("leaf", [1, 2, 8])
def test_bvh_js_other_leaf_sizes(leaf, oracle_mod):
    verts, _ = _meshes()["icosphere_lumpy"]
    rn, ro, rd = J.bvh_build(verts, leaf)
    on, oo, od = oracle_mod.bvh_build(verts, leaf)
    pn, po, pd = capi.bvh_build(verts, leaf)
    assert bits_equal(on, rn) and bits_equal(oo, ro) and od == rd
    assert bits_equal(pn, rn) and bits_equal(po, ro) and pd == rd


def test_env_sampler_js(oracle_mod):
    rng = np.random.default_rng(3)
    noisy = rng.integers(0, 256, (48, 96, 4), dtype=np.uint8)
    noisy[..., 3] = rng.integers(120, 136, (48, 96))
    black = np.zeros((16, 32, 4), np.uint8)
    for env in (pr.environment(64, 32), pr.environment(128, 64), pr.environment(96, 40), noisy, black,
                pr.constant_environment(32, 16, 1.0)):
        ref = J.env_bins(env)
        assert bits_equal(oracle_mod.env_bins(env), ref)
        assert bits_equal(capi.env_bins(env), ref)
        assert len(ref) >= 1


def test_texture_packer_js_indices_and_resolution():
    img = lambda src, h: {"src": src, "pixels": np.zeros((h, h, 4), np.uint8)}
    ops = [("tex", "a.png", 512, True), ("color", [1, 1, 1]), ("tex", "b.png", 2048, False), ("tex", "a.png", 512, True),
           ("color", [1, 1, 1]), ("color", [0, 0, 0]), ("tex", "b.png", 2048, False), ("color", [0.5, 0.25, 1]),
           ("color", [0.1, 1e-7, 1e21])]
    for atlas_res in (4096, 1024, 256):
        r_idx, r_res, r_layers = J.packer_indices(ops, atlas_res)
        p = TexturePacker(atlas_res, pack_layer=lambda *a, **k: None)
        idx = [p.addTexture(img(o[1], o[2]), o[3]) if o[0] == "tex" else p.addColor(o[1]) for o in ops]
        assert idx == r_idx
        assert p.setAndGetResolution() == r_res
        assert [(x["src"] if isinstance(x, dict) else list(x)) for x in p.imageSet] == r_layers
    assert r_idx[0] == 0 and r_idx[3] == 3  # the reference's index-0 layer is never deduplicated (texture_packer.js:14)


OBJ = """# quads, relative indices between vertex blocks, two groups, a group name with a blank
mtllib m.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0.25
vt 0 0
vt 1 0
vt 1 1
vt 0 1 0.5
vn 0 0 1
vn 0 0.6 0.8
usemtl red
f 1/1/1 2/2/1 3/3/2 4/4/2
v 2 0 0
v 3 0.5 1
v 2.5 1 -1
usemtl blue stuff
f -3/1/1 -2/2/2 -1/3/1
f 5/1/-1 3/2/-2 2/3/1
v 0 0 2
v 1 0 2
v 0 1 2
f -3/4/1 -2/3/1   -1/2/2
"""
MTL = """newmtl red
Kd 1 0 0
map_Kd red.png
Ns 12
newmtl blue stuff
Kd 0 0 1
ior 1.5
map_bump  n.png
"""
BASE = dict(scale=0.5, rotate=[{"angle": 0.3, "axis": [0, 0, 1]}, {"angle": -1.1, "axis": [0.6, 0, 0.8]}],
            translate=[0.1, -0.4, 2.0], path="x.obj")
WORLD = [{"rotate": [{"angle": 0.5, "axis": [0, 1, 0]}]}, {"translate": [1, 2, 3]}]


@pytest.mark.parametrize("case", ["mesh", "smooth_world", "flat_skip", "no_uv_spherical"])
def test_obj_loader_js(case):
    obj, tf, wt = OBJ, dict(BASE), None
    if case == "mesh":
        tf["normals"] = "mesh"
    elif case == "smooth_world":
        tf["normals"], wt = "smooth", WORLD
    elif case == "flat_skip":
        tf["normals"], tf["skips"] = "flat", ["red"]
    else:
        tf["normals"] = "flat"
        obj = "\n".join(l for l in OBJ.split("\n") if not l.startswith("vt"))
        for k in "1234":
            obj = obj.replace("/%s/" % k, "//")
    ref = J.parse_mesh(obj, tf, wt, "base", {"base/m.mtl": MTL})
    parsed = SJ.parse_obj(obj, lambda p: {"base/m.mtl": MTL}[p], "base", tf.get("skips"))
    sets = SJ.obj_to_triangle_sets(parsed, tf, wt)
    assert [g for g, _ in sets] == ref["order"]
    for g, ts in sets:
        r = ref["groups"][g]
        for k in ("verts", "normals", "tangents", "bitangents", "uvs"):
            assert bits_equal(getattr(ts, k), r[k]), "%s.%s" % (g, k)
        assert r["material"] == parsed["materials"].get(g, {})
    assert sorted(ref["urls"]) == sorted(parsed["urls"])


def test_mtl_loader_js():
    rm, ru = J.parse_materials(MTL, "base")
    pm, pu = SJ.parse_materials(MTL, "base")
    assert rm == pm and ru == sorted(pu)
