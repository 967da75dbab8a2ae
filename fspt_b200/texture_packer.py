"""texture_packer.js without WebGL: dedup + resample every material map into one res x res RGBA8 layer
(the resampling blit itself is native: csrc/atlas_packer.cpp, fspt_pack_layer).

Reference: TexturePacker (texture_packer.js:5-63) and the blit program of WebGLTextureWriter (:103-121,
:152-176).  Images are dicts {"src": str, "pixels": (h,w,4) uint8 with row 0 = image top, "swizzle": [4]}.
"""
import numpy as np


class TexturePacker:
    def __init__(self, atlas_res, pack_layer=None):
        self.res = atlas_res          # texture_packer.js:7
        self._pack_layer = pack_layer  # the blit; default = the native fspt_pack_layer
        self.imageSet = []
        self.imageKeys = {}
        self.maxRes = 1

    def addTexture(self, image, corrected=False):
        key = image["src"]
        if self.imageKeys.get(key):   # `if (this.imageKeys[key])`: index 0 is never deduplicated (:14)
            return self.imageKeys[key]
        self.maxRes = max(self.maxRes, image["pixels"].shape[0])
        # the reference stores the Image OBJECT and sets `image.corrected` on it (:18-19): every layer that refers to
        # the same object is blitted with the object's final `corrected` / `swizzle` (read in getPixels), so the
        # record is shared and mutated here, not copied
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
                blit = self._pack_layer
                if blit is None:
                    from . import capi  # the blit itself is native (csrc/atlas_packer.cpp)
                    blit = capi.pack_layer
                out[i] = blit(img["pixels"], res, bool(img.get("corrected")), img.get("swizzle"))
        return out


def _js_num(x):
    if float(x) == int(x):
        return str(int(x))
    return repr(float(x))
