/*
 * oracle/oracle_texunit.h -- TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * The GL ES 3.0 texture unit as the oracle models it -- "platform" behaviour, like oracle_math.h: nothing in the
 * reference's own sources defines it, the WebGL2 implementation does.  Shared by the restatement
 * (fspt_oracle.cpp) and by the CPU execution of the reference's own shader sources (oracle/glsl_cpu), so that the
 * two differ only in who wrote the shader logic.
 *
 *   - texture() with LINEAR filtering (GL ES 3.0 section 3.8.10): tau = (1-a)(1-b) t00 + a(1-b) t10 + (1-a)b t01 +
 *     ab t11 in binary32, texels taken at floor(u*size - 0.5) and +1, REPEAT = mathematical modulo, CLAMP_TO_EDGE =
 *     clamp of the texel index; no mip levels (main.js:170-180, 551-555 set LINEAR / LINEAR);
 *   - unorm8 -> f32 is c / 255.0f (GL ES 3.0 section 2.1.6.1);
 *   - the array layer is floor(layer + 0.5) clamped to the layer range (GL ES 3.0 section 3.8.10.2);
 *   - float -> int conversions of coordinates treat NaN / |f| > 1e9 as 0 so that CPU and GPU agree on inputs GLSL
 *     leaves undefined.
 */
#ifndef FSPT_ORACLE_TEXUNIT_H
#define FSPT_ORACLE_TEXUNIT_H

#include <stddef.h>
#include <stdint.h>

#include "oracle_math.h"

namespace om {

static inline long long tu_coord_to_int(float f) {
  if (!(f >= -1.0e9f && f <= 1.0e9f)) f = 0.0f;
  return (long long)f;
}
static inline int tu_wrap_repeat(long long i, int size) {
  long long m = i % size;
  if (m < 0) m += size;
  return (int)m;
}
static inline int tu_wrap_clamp(long long i, int size) { return (int)(i < 0 ? 0 : (i >= size ? size - 1 : i)); }
static inline v4 tu_texel8(const uint8_t* p) {
  v4 r = {(float)p[0] / 255.0f, (float)p[1] / 255.0f, (float)p[2] / 255.0f, (float)p[3] / 255.0f};
  return r;
}
static inline v4 tu_bilerp(v4 t00, v4 t10, v4 t01, v4 t11, float a, float b) {
  float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
  v4 r;
  r.x = w00 * t00.x + w10 * t10.x + w01 * t01.x + w11 * t11.x;
  r.y = w00 * t00.y + w10 * t10.y + w01 * t01.y + w11 * t11.y;
  r.z = w00 * t00.z + w10 * t10.z + w01 * t01.z + w11 * t11.z;
  r.w = w00 * t00.w + w10 * t10.w + w01 * t01.w + w11 * t11.w;
  return r;
}
/* texture(sampler2DArray, vec3(u, v, layer)): RGBA8, REPEAT / REPEAT, LINEAR (main.js:551-555) */
static inline v4 tu_texture_array(const uint8_t* atlas, int R, int layers, float u, float v, float layerf) {
  long long Lq = tu_coord_to_int(floorf(layerf + 0.5f));
  int L = (int)(Lq < 0 ? 0 : (Lq >= layers ? layers - 1 : Lq));
  float x = u * (float)R - 0.5f, y = v * (float)R - 0.5f;
  float fx = floorf(x), fy = floorf(y);
  float a = x - fx, b = y - fy;
  long long ix = tu_coord_to_int(fx), iy = tu_coord_to_int(fy);
  int i0 = tu_wrap_repeat(ix, R), i1 = tu_wrap_repeat(ix + 1, R);
  int j0 = tu_wrap_repeat(iy, R), j1 = tu_wrap_repeat(iy + 1, R);
  const uint8_t* base = atlas + (size_t)L * R * R * 4;
  return tu_bilerp(tu_texel8(base + ((size_t)j0 * R + i0) * 4), tu_texel8(base + ((size_t)j0 * R + i1) * 4),
                   tu_texel8(base + ((size_t)j1 * R + i0) * 4), tu_texel8(base + ((size_t)j1 * R + i1) * 4), a, b);
}
/* texture(sampler2D, vec2(u, v)) on an RGBA8 or SRGB8_ALPHA8 texture with S REPEAT, T CLAMP_TO_EDGE, LINEAR -- the state
 * of both 2-D textures the reference samples with filtering: the environment map (main.js:170-180, RGBA8) and the image
 * being blitted into the atlas (texture_packer.js:88-95,159-165; SRGB8_ALPHA8 when `corrected`).  sRGB texels are
 * decoded before filtering (GL ES 3.0 section 3.8.16), alpha is linear. */
static inline v4 tu_texel8_srgb(const uint8_t* p, int srgb) {
  v4 r = tu_texel8(p);
  if (srgb) {
    float* c = &r.x;
    for (int k = 0; k < 3; ++k) c[k] = c[k] <= 0.04045f ? c[k] / 12.92f : (float)pow(((double)c[k] + 0.055) / 1.055, 2.4);
  }
  return r;
}
static inline v4 tu_texture_2d(const uint8_t* tex, int W, int H, int srgb, float u, float v) {
  float x = u * (float)W - 0.5f, y = v * (float)H - 0.5f;
  float fx = floorf(x), fy = floorf(y);
  float a = x - fx, b = y - fy;
  long long ix = tu_coord_to_int(fx), iy = tu_coord_to_int(fy);
  int i0 = tu_wrap_repeat(ix, W), i1 = tu_wrap_repeat(ix + 1, W);
  int j0 = tu_wrap_clamp(iy, H), j1 = tu_wrap_clamp(iy + 1, H);
  return tu_bilerp(tu_texel8_srgb(tex + ((size_t)j0 * W + i0) * 4, srgb), tu_texel8_srgb(tex + ((size_t)j0 * W + i1) * 4, srgb),
                   tu_texel8_srgb(tex + ((size_t)j1 * W + i0) * 4, srgb), tu_texel8_srgb(tex + ((size_t)j1 * W + i1) * 4, srgb), a, b);
}
static inline v4 tu_texture_env(const uint8_t* env, int W, int H, float u, float v) { return tu_texture_2d(env, W, H, 0, u, v); }
/* RGBA8 colour-buffer write: clamp to [0, 1], round to nearest (GL ES 3.0 section 2.1.6.1) */
static inline uint8_t tu_quant8(float v) {
  if (!(v > 0.0f)) return 0;
  if (v >= 1.0f) return 255;
  return (uint8_t)(int)floorf(v * 255.0f + 0.5f);
}

}  // namespace om
#endif
