"""The host worker pool of fspt_scene_upload (fspt_b200/csrc/host_pool.h) under stress, on the CPU: every item of
every region runs exactly once, for any region shape and with two calling threads; also under ThreadSanitizer when the
toolchain has it."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_pool_stress.cpp")


def _build_and_run(tmp_path, extra, name):
    exe = str(tmp_path / name)
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread"] + extra + [SRC, "-o", exe], capture_output=True, text=True)
    if r.returncode != 0:
        return None, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    return r, r.stdout + r.stderr


def test_host_pool_runs_every_item_exactly_once(tmp_path):
    assert shutil.which("g++")
    r, out = _build_and_run(tmp_path, [], "pool_stress")
    assert r is not None, out
    assert r.returncode == 0 and out.startswith("ok"), out


def test_host_pool_under_thread_sanitizer(tmp_path):
    r, out = _build_and_run(tmp_path, ["-fsanitize=thread", "-g"], "pool_stress_tsan")
    if r is None:
        pytest.skip("ThreadSanitizer runtime not available: " + out.strip().splitlines()[-1][:120])
    if "FATAL: ThreadSanitizer" in out and "unexpected memory mapping" in out:
        pytest.skip("ThreadSanitizer cannot map its shadow memory in this container")
    assert r.returncode == 0 and "WARNING: ThreadSanitizer" not in out, out[-3000:]
