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
