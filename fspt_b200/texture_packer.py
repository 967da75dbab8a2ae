"""texture_packer.js without WebGL: dedup + resample every material map into one res x res RGBA8 layer.

Reference: TexturePacker (texture_packer.js:5-63) and the blit program of WebGLTextureWriter (:103-121,
:152-176).  Images are dicts {"src": str, "pixels": (h,w,4) uint8 with row 0 = image top, "swizzle": [4]}.
"""
import numpy as np


def _srgb_to_linear(c):
    c = c.astype(np.float32) / np.float32(255.0)
    lo = c / np.float32(12.92)
    hi = np.power((c + np.float32(0.055)) / np.float32(1.055), np.float32(2.4)).astype(np.float32)
    return np.where(c <= np.float32(0.04045), lo, hi).astype(np.float32)


class TexturePacker:
    def __init__(self, atlas_res):
        self.res = atlas_res          # texture_packer.js:7
        self.imageSet = []
        self.imageKeys = {}
        self.maxRes = 1

    def addTexture(self, image, corrected=False):
        key = image["src"]
        if self.imageKeys.get(key):   # `if (this.imageKeys[key])`: index 0 is never deduplicated (:14)
            return self.imageKeys[key]
        self.maxRes = max(self.maxRes, image["pixels"].shape[0])
        image = dict(image)
        image["corrected"] = corrected
        self.imageSet.append(image)
        self.imageKeys[key] = len(self.imageSet) - 1
        return self.imageKeys[key]

    def addColor(self, color):
        key = " ".join(_js_num(c) for c in color)  # color.join(' ') (:26)
        if self.imageKeys.get(key):
            return self.imageKeys[key]
        self.imageSet.append(list(color))
        self.imageKeys[key] = len(self.imageSet) - 1
        return self.imageKeys[key]

    def setAndGetResolution(self):
        if self.maxRes < self.res:
            self.res = self.maxRes
        return self.res

    def getPixels(self):
        """-> (layers, res, res, 4) uint8; row 0 of a layer = GL row 0 (bottom of the blit = image bottom)."""
        res = self.res
        out = np.zeros((len(self.imageSet), res, res, 4), np.uint8)
        for i, img in enumerate(self.imageSet):
            if isinstance(img, list):
                # clearColor(r,g,b,1) on an RGBA8 canvas (:152-157): unorm conversion, round to nearest
                c = np.clip(np.asarray(img[:3], np.float32), 0.0, 1.0)
                out[i, :, :, :3] = np.floor(c * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
                out[i, :, :, 3] = 255
            else:
                out[i] = _blit(img, res)
        return out


def _js_num(x):
    if float(x) == int(x):
        return str(int(x))
    return repr(float(x))


def _blit(img, res):
    """The fragment program of texture_packer.js:103-121 evaluated per output texel in f32."""
    px = img["pixels"]
    h, w = px.shape[0], px.shape[1]
    src = _srgb_to_linear(px) if img.get("corrected") else px.astype(np.float32) / np.float32(255.0)
    if img.get("corrected"):
        src[..., 3] = px[..., 3].astype(np.float32) / np.float32(255.0)  # alpha of SRGB8_ALPHA8 stays linear
    # texImage2D(image): row 0 of the texture = top of the image (no UNPACK_FLIP_Y), t axis up
    fc = (np.arange(res, dtype=np.float32) + np.float32(0.5)) / np.float32(res)  # gl_FragCoord.xy / dims
    u = fc
    v = np.float32(1.0) - fc                                                     # uv.y = 1.0 - uv.y
    x = u * np.float32(w) - np.float32(0.5)
    y = v * np.float32(h) - np.float32(0.5)
    fx, fy = np.floor(x), np.floor(y)
    ax, ay = (x - fx).astype(np.float32), (y - fy).astype(np.float32)
    i0 = np.mod(fx.astype(np.int64), w)
    i1 = np.mod(fx.astype(np.int64) + 1, w)             # WRAP_S = REPEAT (:91)
    j0 = np.clip(fy.astype(np.int64), 0, h - 1)
    j1 = np.clip(fy.astype(np.int64) + 1, 0, h - 1)     # WRAP_T = CLAMP_TO_EDGE (:92)
    t00 = src[j0][:, i0]
    t10 = src[j0][:, i1]
    t01 = src[j1][:, i0]
    t11 = src[j1][:, i1]
    AX = ax[None, :, None]
    AY = ay[:, None, None]
    one = np.float32(1.0)
    c = ((one - AX) * (one - AY)) * t00 + (AX * (one - AY)) * t10 + ((one - AX) * AY) * t01 + (AX * AY) * t11
    sw = img.get("swizzle") or [0, 1, 2, 3]
    c = c[..., [int(s) for s in sw]]
    rgb = c[..., :3] * c[..., 3:4]                       # fragColor = vec4(c.rgb * c.a, 1.0)
    o = np.empty((res, res, 4), np.uint8)
    o[..., :3] = np.floor(np.clip(rgb, 0.0, 1.0) * np.float32(255.0) + np.float32(0.5)).astype(np.uint8)
    o[..., 3] = 255
    return o  # row index = gl_FragCoord.y: readPixels returns rows bottom-up and they are uploaded as-is (:178-184)
