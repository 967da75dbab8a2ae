"""fspt_b200/csrc/fastdiv.h -- the launch-invariant integer divisions of the kernels (slot -> pixel, slot -> sample) --
is exact: every divisor class against the hardware division on the CPU, exhaustively for the bench frame's divisors."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fastdiv_is_exact(tmp_path):
    exe = str(tmp_path / "fastdiv_check")
    r = subprocess.run(["g++", "-O2", "-std=c++17", os.path.join(ROOT, "tests", "fastdiv_check.cpp"), "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr
