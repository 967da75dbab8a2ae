"""The reference's own fragment shaders executed on the CPU -- TEST INFRASTRUCTURE, never imported by fspt_b200/.

ctypes front-end of oracle/_ref/libfspt_ref.so = /root/reference/shader/{camera,tracer,bvh_test,draw}.fs compiled by
g++ behind a GLSL subset (oracle/glsl_cpu/; recipe: `make -C oracle ref`).  Same call signatures as the restatement in
oracle/__init__.py (Oracle.bvh_test / Oracle.trace / camera / draw), so tests put the two side by side.

The library can only be BUILT where the reference tree exists (this container); the built file is git-ignored but
travels to the GPU box with the snapshot.  `available()` says whether it can be used.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import OScene, Oracle, _f32, _p

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libfspt_ref.so")
REFERENCE_ROOT = os.environ.get("FSPT_REFERENCE_ROOT", "/root/reference")


def can_build():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "shader", "tracer.fs"))


def build(force=False):
    """(Re)builds oracle/_ref/libfspt_ref.so when the reference tree is present; returns the path or None."""
    if can_build():
        cmd = ["make", "-C", _HERE, "-s", "ref", "REFERENCE=" + REFERENCE_ROOT] + (["-B"] if force else [])
        subprocess.check_call(cmd)
    return _LIB if os.path.exists(_LIB) else None


def available():
    return build() is not None


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("oracle/_ref/libfspt_ref.so is not built and %s does not exist" % REFERENCE_ROOT)
        _lib = C.CDLL(path)
    return _lib


class Reference(Oracle):
    """Oracle's scene holder, with the traversal / path-tracing passes run by the reference's shader sources."""

    def bvh_test(self, pos4, dir4, nthreads=0, want_heat=False):
        pos4, dir4 = _f32(pos4).reshape(-1, 4), _f32(dir4).reshape(-1, 4)
        n = pos4.shape[0]
        idx = np.empty(n, np.int32); t = np.empty(n, np.float32); cnt = np.empty(n, np.int32)
        heat = np.empty((n, 4), np.float32) if want_heat else None
        lib().ref_bvh_test(C.byref(self.s), _p(pos4), _p(dir4), C.c_int(n), _p(idx), _p(t), _p(cnt), _p(heat),
                           C.c_int(nthreads))
        return (idx, t, cnt, heat) if want_heat else (idx, t, cnt)

    def trace(self, pos4, dir4, W, H, tick, rand_base, env_theta, fb_prev=None, nthreads=0):
        pos4, dir4 = _f32(pos4).reshape(-1, 4), _f32(dir4).reshape(-1, 4)
        assert pos4.shape[0] == W * H
        fb_prev = _f32(fb_prev).reshape(-1, 4) if fb_prev is not None else None
        out = np.empty((H, W, 4), np.float32)
        lib().ref_trace(C.byref(self.s), _p(pos4), _p(dir4), C.c_int(W), C.c_int(H), C.c_uint32(tick),
                        C.c_float(rand_base), C.c_float(env_theta), _p(fb_prev), _p(out), C.c_int(nthreads))
        return out


def camera(W, H, P, I, fov_scale, lens, rand_base, nthreads=0):
    P, I, lens = _f32(P), _f32(I), _f32(lens)
    pos = np.empty((H, W, 4), np.float32); d = np.empty((H, W, 4), np.float32)
    lib().ref_camera(C.c_int(W), C.c_int(H), _p(P), _p(I), C.c_float(fov_scale), _p(lens), C.c_float(rand_base),
                     _p(pos), _p(d), C.c_int(nthreads))
    return pos, d


def draw(fb, exposure=1.0, saturation=1.0, denoise=False, max_sigma=2.0, scale=1.0, nthreads=0, want_float=False):
    fb = _f32(fb)
    H, W = fb.shape[0], fb.shape[1]
    out = np.empty((H, W, 4), np.uint8)
    outf = np.empty((H, W, 4), np.float32) if want_float else None
    lib().ref_draw(_p(fb), C.c_int(W), C.c_int(H), C.c_float(exposure), C.c_float(saturation),
                   C.c_int(1 if denoise else 0), C.c_float(max_sigma), C.c_float(scale), _p(out), _p(outf), C.c_int(nthreads))
    return (out, outf) if want_float else out


def pack_layer(pixels, res, corrected=False, swizzle=None):
    """One atlas layer by the blit shader inside texture_packer.js (:103-121), GL state of :88-95,159-176."""
    pixels = np.ascontiguousarray(pixels, dtype=np.uint8)
    h, w = pixels.shape[0], pixels.shape[1]
    out = np.empty((res, res, 4), np.uint8)
    sw = np.asarray(swizzle, np.int32) if swizzle is not None else None
    lib().ref_pack_layer(_p(pixels), C.c_int(w), C.c_int(h), C.c_int(res), C.c_int(1 if corrected else 0), _p(sw), _p(out))
    return out
