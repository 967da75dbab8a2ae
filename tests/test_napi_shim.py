"""The N-API shim (fspt_b200/napi/fspt_napi.cc) cannot be built here -- no Node toolchain, no node_api.h -- but it must
keep type-checking against include/fspt_b200.h: compiled with -fsyntax-only against a declaration-only stub header."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_napi_shim_type_checks_against_the_c_abi():
    src = os.path.join(ROOT, "fspt_b200", "napi", "fspt_napi.cc")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", os.path.join(ROOT, "tests", "stubs"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_napi_shim_exposes_the_multi_gpu_entry_points():
    text = open(os.path.join(ROOT, "fspt_b200", "napi", "fspt_napi.cc")).read()
    for c_name in ("fspt_set_tile", "fspt_comm_unique_id", "fspt_comm_init", "fspt_reduce_accum", "fspt_scene_broadcast",
                   "fspt_set_accum_mode", "fspt_scene_upload", "fspt_render", "fspt_resolve"):
        assert re.search(r"\b%s\(" % c_name, text), c_name
    host = open(os.path.join(ROOT, "fspt_b200", "napi", "main_multi.mjs")).read()
    for js_name in ("commUniqueId", "commInit", "sceneBroadcast", "setTile", "reduceAccum"):
        assert "fspt." + js_name in host and '"%s"' % js_name in text
