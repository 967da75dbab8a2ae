// atlas_packer.cpp -- texture_packer.js without WebGL: resamples one material map into a res x res
// RGBA8 atlas layer exactly as the reference's blit program does (texture_packer.js:103-121,159-184):
//   uv = gl_FragCoord.xy / dims;  uv.y = 1 - uv.y;  c = texture(tex, uv)   [LINEAR, S = REPEAT, T = CLAMP_TO_EDGE,
//   SRGB8_ALPHA8 upload for base-colour maps => sRGB -> linear per texel BEFORE filtering]
//   c = swizzle(c);  fragColor = vec4(c.rgb * c.a, 1)  -> RGBA8 canvas -> readPixels (bottom-up rows, uploaded as-is).
// Host code (no GPU needed), multi-threaded over rows; all arithmetic is single IEEE f32 operations.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <thread>
#include <vector>

#include "../../include/fspt_b200.h"

namespace {

struct Lut {
  float lin[256];   // c / 255
  float srgb[256];  // sRGB EOTF of c / 255 (GL ES 3.0 section 3.8.16)
  Lut() {
    for (int i = 0; i < 256; ++i) {
      const float c = (float)i / 255.0f;
      lin[i] = c;
      srgb[i] = c <= 0.04045f ? c / 12.92f : (float)pow(((double)c + 0.055) / 1.055, 2.4);
    }
  }
};

inline uint8_t quant8(float v) {  // RGBA8 colour-buffer write: clamp, round to nearest
  if (!(v > 0.0f)) return 0;
  if (v >= 1.0f) return 255;
  return (uint8_t)(int)floorf(v * 255.0f + 0.5f);
}

}  // namespace

extern "C" int fspt_pack_layer(const uint8_t* rgba8, int32_t w, int32_t h, int32_t res, int32_t corrected,
                               const int32_t* swizzle, uint8_t* out, int32_t n_threads) {
  if (!rgba8 || !out || w <= 0 || h <= 0 || res <= 0) return FSPT_E_INVALID;
  int sw[4] = {0, 1, 2, 3};
  if (swizzle)
    for (int k = 0; k < 4; ++k) {
      if (swizzle[k] < 0 || swizzle[k] > 3) return FSPT_E_INVALID;
      sw[k] = swizzle[k];
    }
  static const Lut lut;
  const float* rgb_lut = corrected ? lut.srgb : lut.lin;
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  n_threads = std::max(1, std::min(n_threads, (int)res));
  // horizontal taps are the same for every row: precompute them once
  std::vector<int> i0((size_t)res), i1((size_t)res);
  std::vector<float> ax((size_t)res);
  for (int x = 0; x < res; ++x) {
    const float u = ((float)x + 0.5f) / (float)res;       // gl_FragCoord.x / dims.x
    const float fx = u * (float)w - 0.5f;
    const float fl = floorf(fx);
    ax[x] = fx - fl;
    long long ix = (long long)fl;
    int m0 = (int)(ix % w), m1 = (int)((ix + 1) % w);
    if (m0 < 0) m0 += w;
    if (m1 < 0) m1 += w;
    i0[x] = m0; i1[x] = m1;                               // WRAP_S = REPEAT (texture_packer.js:91)
  }
  std::atomic<int> next(0);
  auto work = [&]() {
    for (;;) {
      const int y = next.fetch_add(1);
      if (y >= res) break;
      float v = ((float)y + 0.5f) / (float)res;           // gl_FragCoord.y / dims.y
      v = 1.0f - v;                                       // uv.y = 1.0 - uv.y
      const float fy = v * (float)h - 0.5f;
      const float fl = floorf(fy);
      const float ay = fy - fl;
      long long iy = (long long)fl;
      const int j0 = (int)std::max<long long>(0, std::min<long long>(h - 1, iy));      // WRAP_T = CLAMP_TO_EDGE
      const int j1 = (int)std::max<long long>(0, std::min<long long>(h - 1, iy + 1));
      const uint8_t* r0 = rgba8 + (size_t)j0 * w * 4;     // texture row j = image row j (no UNPACK_FLIP_Y)
      const uint8_t* r1 = rgba8 + (size_t)j1 * w * 4;
      uint8_t* o = out + (size_t)y * res * 4;             // readPixels row y = gl_FragCoord.y
      for (int x = 0; x < res; ++x) {
        const uint8_t* t00 = r0 + (size_t)i0[x] * 4; const uint8_t* t10 = r0 + (size_t)i1[x] * 4;
        const uint8_t* t01 = r1 + (size_t)i0[x] * 4; const uint8_t* t11 = r1 + (size_t)i1[x] * 4;
        const float a = ax[x];
        const float w00 = (1.0f - a) * (1.0f - ay), w10 = a * (1.0f - ay), w01 = (1.0f - a) * ay, w11 = a * ay;
        float c[4];
        for (int k = 0; k < 4; ++k) {
          const float* l = k < 3 ? rgb_lut : lut.lin;     // alpha of SRGB8_ALPHA8 stays linear
          c[k] = w00 * l[t00[k]] + w10 * l[t10[k]] + w01 * l[t01[k]] + w11 * l[t11[k]];
        }
        const float s0 = c[sw[0]], s1 = c[sw[1]], s2 = c[sw[2]], s3 = c[sw[3]];
        o[4 * x + 0] = quant8(s0 * s3);                   // fragColor = vec4(c.rgb * c.a, 1.0)
        o[4 * x + 1] = quant8(s1 * s3);
        o[4 * x + 2] = quant8(s2 * s3);
        o[4 * x + 3] = 255;
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < n_threads; ++t) th.emplace_back(work);
  work();
  for (auto& t : th) t.join();
  return FSPT_OK;
}
