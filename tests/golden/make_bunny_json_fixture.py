#!/usr/bin/env python
"""tests/golden/bunny_json_assets.npz: the reference's own scene/bunny.json, compiled from the BUNDLED assets that are
present in the reference tree (five 2048x2048 dungeon maps as JPEG / PNG, asset_packs/misc/top_mono.obj), with only the
four entries of .MISSING_LARGE_BLOBS substituted:

    asset_packs/misc/bunny_big.obj                 -> OBJ text of a lumpy icosphere (5,120 triangles), generated here
    environment/autumn_meadow_2k.RGBE.PNG          -> procedural 256x128 RGBE environment
    asset_packs/dungeon/RootNode_normal.png        -> procedural normal map
    asset_packs/dungeon/Scene_-_Root_normal.png    -> procedural normal map

and `atlasRes` set to 128 so that the fixture stays small (the JSON has no atlasRes; main.js:948 would default to 2048).
Needs /root/reference (build container only); the resulting arrays + oracle outputs travel to the GPU box inside the
fixture, where tests/test_golden.py replays them through the CUDA path.  tests/test_scene_json.py re-derives the scene
arrays from the reference tree when it is present and compares them with the fixture byte for byte.

    python tests/golden/make_bunny_json_fixture.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

REF = os.environ.get("FSPT_REFERENCE_ROOT", "/root/reference")
ATLAS_RES = 128
MISSING = ("asset_packs/misc/bunny_big.obj", "environment/autumn_meadow_2k.RGBE.PNG",
           "asset_packs/dungeon/RootNode_normal.png", "asset_packs/dungeon/Scene_-_Root_normal.png")


def substitute_obj():
    from fspt_b200 import procedural as pr
    v, f = pr.icosphere(4)
    v = pr.lumpy(v)
    lines = ["v %r %r %r" % tuple(float(x) for x in p) for p in v]
    lines += ["f %d %d %d" % tuple(int(i) + 1 for i in t) for t in f]
    return "\n".join(lines) + "\n"


def compile_bunny_json(ref_root=REF, atlas_res=ATLAS_RES):
    """scene/bunny.json -> (SceneArrays, camera) through fspt_b200.scene_json, the four missing blobs substituted."""
    from fspt_b200 import procedural as pr, scene_json
    with open(os.path.join(ref_root, "scene", "bunny.json")) as f:
        scene = json.load(f)
    missing = [l.strip() for l in open(os.path.join(ref_root, ".MISSING_LARGE_BLOBS")) if l.strip()]
    assert sorted(missing) == sorted(MISSING), missing
    scene["atlasRes"] = atlas_res
    obj_text = substitute_obj()

    def read_text(path):
        if path == "asset_packs/misc/bunny_big.obj":
            return obj_text
        return open(os.path.join(ref_root, path)).read()

    def load_img(path):
        if path == "environment/autumn_meadow_2k.RGBE.PNG":
            return {"src": path, "pixels": pr.environment(256, 128)}
        if path == "asset_packs/dungeon/RootNode_normal.png":
            return dict(pr.pbr_maps(atlas_res, 7, "A")["normal"], src=path)
        if path == "asset_packs/dungeon/Scene_-_Root_normal.png":
            return dict(pr.pbr_maps(atlas_res, 11, "B")["normal"], src=path)
        return scene_json.load_image(os.path.join(ref_root, path))
    sa, cam = scene_json.compile_scene(scene, ref_root, read_text=read_text, load_img=load_img)
    return sa, cam


if __name__ == "__main__":
    import make_golden
    from fspt_b200 import scene_json
    sa, cam = compile_bunny_json()
    # shootAutoFocusRay (main.js:447-546) -> lensFeatures[0]
    dist = scene_json.autofocus_distance(sa.verts64, cam["eye"], cam["dir"])
    cam = dict(cam, focal_depth=dist)
    print("layers", sa.atlas.shape, "tris", sa.n_tris, "autofocus distance", dist)
    make_golden.make("bunny_json_assets", sa, cam, 160, 96, 3, 31, dict(exposure=1.0, saturation=1.0, max_sigma=2.0, denoise=True))
