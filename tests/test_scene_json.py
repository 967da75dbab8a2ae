"""Scene-JSON / OBJ / MTL front end (row f4): the file-based compile path must produce the same flattened arrays
as the in-memory path, and follow obj_loader.js / mtl_loader.js semantics."""
import json
import os

import numpy as np
import pytest

from fspt_b200 import procedural as pr, scene_json, scenes

QUAD_OBJ = """v  0.5  0.0 0.5
v  0.5  0.0 -0.5
v  -0.5  0.0 -0.5
v  -0.5  0.0 0.5

vt  0.0  0.0
vt  0.0  1.0
vt  1.0  1.0
vt  1.0  0.0

f 1/1 3/3 2/2
f 3/3 1/1 4/4"""  # == asset_packs/misc/top_mono.obj


def _write_assets(tmp_path, extra_obj=None):
    from PIL import Image
    (tmp_path / "scene").mkdir()
    (tmp_path / "asset_packs" / "misc").mkdir(parents=True)
    (tmp_path / "environment").mkdir()
    (tmp_path / "asset_packs" / "misc" / "top_mono.obj").write_text(QUAD_OBJ)
    if extra_obj:
        for name, text in extra_obj.items():
            (tmp_path / "asset_packs" / "misc" / name).write_text(text)
    maps = pr.pbr_maps(16, 7, "A")
    for k, img in maps.items():
        Image.fromarray(img["pixels"], "RGBA").save(tmp_path / "asset_packs" / "misc" / ("%s.png" % k))
    Image.fromarray(pr.environment(64, 32), "RGBA").save(tmp_path / "environment" / "env.RGBE.PNG")
    return maps


def test_scene_json_matches_in_memory_compile(tmp_path):
    maps = _write_assets(tmp_path)
    scene = {
        "environment": "environment/env.RGBE.PNG", "environmentTheta": 1.66, "cameraPos": [-0.751, 0.665, 1.82],
        "cameraDir": [0.304, -0.489, -0.818], "atlasRes": 16,
        "props": [{"path": "asset_packs/misc/top_mono.obj", "scale": 4, "rotate": [{"angle": 3.1415, "axis": [0, 0, 1]}],
                   "translate": [0, -0.75, 0], "emittance": [0, 0, 0], "normals": "flat",
                   "diffuse": "asset_packs/misc/baseColor.png", "metallicRoughness": "asset_packs/misc/metallicRoughness.png",
                   "normal": "asset_packs/misc/normal.png"}]}
    (tmp_path / "scene" / "q.json").write_text(json.dumps(scene))
    sa, cam = scene_json.load_scene(str(tmp_path / "scene" / "q.json"), emulate_canvas=False)
    # the same prop through the in-memory path
    p = dict(scene["props"][0])
    p["mesh"] = (pr.QUAD_VERTS, pr.QUAD_FACES, pr.QUAD_FACE_UVS)
    p["diffuse"], p["metallicRoughness"], p["normal"] = "A/baseColor", "A/metallicRoughness", "A/normal"
    assets = {"A/" + k: v for k, v in maps.items()}
    env = pr.environment(64, 32)
    from fspt_b200 import capi
    ref = scenes.compile_props([p], assets, 16, (env, capi.env_bins(env)))
    for k in ("bvh", "tris", "mats", "norms", "uvs", "atlas", "env", "bins"):
        a, b = getattr(sa, k), getattr(ref, k)
        assert a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8)), k
    assert cam["env_theta"] == 1.66 and cam["fov_scale"] == 0.5 and cam["samples"] == 2000


def test_obj_semantics(tmp_path):
    obj = """mtllib m.mtl
v 0 0 0
v 1 0 0
v 1 1 0
v 0 1 0
v 0 0 1
usemtl red
f 1 2 3 4
usemtl glass
f -1 1 2
usemtl skipme
f 1 2 5
"""
    mtl = """newmtl red
Kd 1 0 0
Pmr 0.5 0.25 0
newmtl glass
ior 1.5
dielectric 0.2
map_Kd tex.png
"""
    parsed = scene_json.parse_obj(obj, lambda p: mtl, "base", skips=["skipme"])
    assert parsed["groups"] == ["red", "glass"]
    assert len(parsed["faces"]) == 3  # the quad fans into 2 triangles (obj_loader.js:54-60), skipped group dropped
    assert parsed["materials"]["red"]["kd"] == [1.0, 0.0, 0.0] and parsed["materials"]["red"]["pmr"] == [0.5, 0.25, 0.0]
    assert parsed["materials"]["glass"]["ior"] == 1.5 and parsed["materials"]["glass"]["dielectric"] == 0.2
    assert parsed["urls"] == {"base/tex.png"}
    prop = {"scale": 2, "rotate": [], "translate": [0, 0, 1], "normals": "flat"}
    sets = dict(scene_json.obj_to_triangle_sets(parsed, prop, None))
    assert sets["red"].count == 2 and sets["glass"].count == 1
    # negative index -1 = last vertex (0,0,1) -> scaled and translated to (0,0,3)
    assert np.allclose(sets["glass"].verts[0, 0], [0, 0, 3])
    # no vt -> spherical mapping from the vertex direction (obj_loader.js:63-70) + Number.EPSILON offsets
    v = sets["red"].verts[0, 1]
    d = v / np.linalg.norm(v)
    assert np.isclose(sets["red"].uvs[0, 1, 0], np.arctan2(d[2], d[0]) / (2 * np.pi) + 2 * 2.0 ** -52)


def test_autofocus_distance_matches_hand_value():
    verts = np.array([[[0.5, 0, 0.5], [-0.5, 0, -0.5], [0.5, 0, -0.5]], [[-0.5, 0, -0.5], [0.5, 0, 0.5], [-0.5, 0, 0.5]]])
    assert scene_json.autofocus_distance(verts, [0.1, 2.0, 0.2], [0, -1, 0]) == 2.0
    assert scene_json.autofocus_distance(verts, [0.1, 2.0, 0.2], [0, 1, 0]) == 1e6  # miss -> maxT (main.js:44,543)


def test_black_colour_stop_environment(tmp_path):
    _write_assets(tmp_path)
    scene = {"environment": [[0, 0, 0], [0, 0, 0]],
             "props": [{"path": "asset_packs/misc/top_mono.obj", "scale": 1, "rotate": [], "translate": [0, 0, 0],
                        "emittance": [0, 0, 0], "diffuse": [0.8, 0.8, 0.8]}]}
    sa, _ = scene_json.compile_scene(scene, str(tmp_path))
    assert sa.env.shape == (2048, 1, 4) and sa.bins.tolist() == [[0, 0, 1, 2048]]
    with pytest.raises(NotImplementedError):
        scene_json.compile_scene(dict(scene, environment=[[1, 1, 1], [0, 0, 0]]), str(tmp_path))
    with pytest.raises(ValueError):
        scene_json.compile_scene({k: v for k, v in scene.items() if k != "environment"}, str(tmp_path))


def test_relative_obj_indices_resolve_at_the_face_line():
    """obj_loader.js parseTriangle (:107-113): `vertices.length + idx + 1` uses the vertices read SO FAR, so two
    v/f blocks with `f -3 -2 -1` make two different triangles (an OBJ exporter's usual layout)."""
    obj = "\n".join(["v 0 0 0", "v 1 0 0", "v 0 1 0", "vn 0 0 1", "f -3//-1 -2//-1 -1//-1",
                     "v 5 5 5", "v 6 5 5", "v 5 6 5", "vn 1 0 0", "f -3//-1 -2//-1 -1//-1", "f 1 2 3"])
    p = scene_json.parse_obj(obj, lambda path: "", "")
    got = [[int(c[0]) for c in tri] for _, tri in p["faces"]]
    assert got == [[1, 2, 3], [4, 5, 6], [1, 2, 3]]
    assert [[int(c[2]) for c in tri] for _, tri in p["faces"][:2]] == [[1, 1, 1], [2, 2, 2]]
    sets = scene_json.obj_to_triangle_sets(p, {"rotate": [], "normals": "mesh"}, None)
    v = sets[0][1].verts.reshape(-1, 3, 3)
    assert np.array_equal(v[0][0], [0, 0, 0]) and np.array_equal(v[1][0], [5, 5, 5])


def test_shared_image_takes_the_last_swizzle_like_the_reference():
    """getMaterial (main.js:226-237) assigns `img.swizzle` on the SHARED Image object and the packer blits with the
    object's final state (texture_packer.js:159-176): a map reused with another (or no) swizzle changes the
    layer that was added first."""
    from fspt_b200.scene import get_material
    from fspt_b200.texture_packer import TexturePacker
    px = np.zeros((4, 4, 4), np.uint8)
    px[..., 0], px[..., 1], px[..., 2], px[..., 3] = 10, 20, 30, 255
    base = {"emittance": [0, 0, 0], "diffuse": [1, 1, 1]}

    def run(second_swizzle):
        assets = {"mr.png": {"src": "mr.png", "pixels": px.copy()}}
        packer = TexturePacker(4)
        get_material(dict(base, metallicRoughness="mr.png", mrSwizzle=[2, 1, 0, 3]), {}, packer, assets)
        second = dict(base, metallicRoughness="mr.png")
        if second_swizzle is not None:
            second["mrSwizzle"] = second_swizzle
        m2 = get_material(second, {}, packer, assets)
        packer.setAndGetResolution()
        return packer.getPixels()[m2["roughnessIndex"]][0, 0, :3].tolist()
    assert run([2, 1, 0, 3]) == [30, 20, 10]
    assert run(None) == [10, 20, 30]        # `img.swizzle = undefined` resets the shared object
    assert run([1, 1, 1, 3]) == [20, 20, 20]


REF = os.environ.get("FSPT_REFERENCE_ROOT", "/root/reference")
needs_reference = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "asset_packs", "dungeon")),
                                     reason="the reference tree (bundled dungeon maps) is only present in the build container")


@needs_reference
def test_bunny_json_with_the_bundled_assets_reproduces_the_fixture():
    """The reference's scene/bunny.json through scene_json with the bundled JPEG / PNG maps and top_mono.obj (only the
    four .MISSING_LARGE_BLOBS entries substituted): the arrays are the ones frozen in tests/golden/bunny_json_assets.npz,
    which the GPU box replays through the CUDA path (tests/test_golden.py)."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_bunny_json_fixture as mk
    sa, cam = mk.compile_bunny_json(REF)
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bunny_json_assets.npz"))
    for k in ("bvh", "tris", "mats", "norms", "uvs", "atlas", "env", "bins"):
        assert np.array_equal(np.ascontiguousarray(getattr(sa, k)).view(np.uint8), np.ascontiguousarray(z[k]).view(np.uint8)), k
    assert sa.atlas.shape == (11, 128, 128, 4)       # 4 colours + 7 image maps; the emission colour [0,0,0] deduplicates
    # layer order = first addTexture/addColor call (main.js:212-261): prop 0 colours, then the two quads' maps
    assert np.all(sa.atlas[0, 0, 0] == [255, 255, 255, 255]) and np.all(sa.atlas[2, 0, 0] == [0, 0, 0, 255])
    assert sa.atlas[4].std() > 5 and sa.atlas[9].std() > 5   # RootNode_baseColor.png, Scene_-_Root_emissive.jpeg
    assert [float(x) for x in z["eye"]] == pytest.approx(cam["eye"]) and float(z["env_theta"]) == pytest.approx(1.66)
    # the third prop's `"ior": "10"` (a string in the JSON) reaches the material record as the number 10
    assert set(np.unique(sa.mats[:, 9]).tolist()) == {np.float32(1.4), np.float32(10.0)}


@needs_reference
def test_native_blit_of_a_real_2048_map_matches_the_per_fragment_restatement(oracle_mod):
    """fspt_pack_layer on real decoded data (an sRGB JPEG and an RGBA PNG with varying alpha), at the real 2048 px
    resolution for a band of rows and resampled to 512: bit-identical to oracle_pack_layer."""
    from fspt_b200 import capi
    for name, corrected, swz in (("Scene_-_Root_baseColor.jpeg", True, None), ("RootNode_baseColor.png", True, None),
                                 ("RootNode_metallicRoughness.png", False, [2, 1, 0, 3])):
        px = scene_json.load_image(os.path.join(REF, "asset_packs", "dungeon", name))["pixels"]
        assert px.shape == (2048, 2048, 4)
        got = capi.pack_layer(px, 512, corrected, swz)
        ref = oracle_mod.pack_layer(px, 512, corrected, swz)
        assert np.array_equal(got, ref), name
        band = px[:96]   # full-resolution texels, a band of rows keeps the per-fragment oracle quick
        assert np.array_equal(capi.pack_layer(band, 2048, corrected, swz)[:, :64], oracle_mod.pack_layer(band, 2048, corrected, swz)[:, :64])
