/*
 * oracle/glsl_cpu/ref_driver.cpp -- TEST INFRASTRUCTURE.  The reference's own fragment shaders, executed on the CPU.
 *
 * Compiles /root/reference/shader/{camera,tracer,bvh_test,draw}.fs and the atlas blit shader embedded in
 * /root/reference/texture_packer.js -- adapted lexically by glsl2cpp.py into
 * oracle/_ref/*.gen.inc at build time, never copied into the repository -- behind the GLSL subset of glsl_body.inc,
 * one namespace per shader, and runs their main() once per fragment the way main.js's draw calls do
 * (main.js:728-807: camera pass -> tracer pass with ping-ponged accumulation targets -> draw pass).  The entry points
 * mirror oracle/fspt_oracle.cpp's (same scene struct, same buffers), so tests can put the two side by side:
 * whatever differs is a difference between the oracle's restatement and the reference's shader text.
 *
 * Built only where /root/reference exists (oracle/Makefile target `ref`); output oracle/_ref/libfspt_ref.so.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#include "../oracle_math.h"
#include "../oracle_texunit.h"

extern "C" {
/* == OScene of fspt_oracle.cpp: the GL resources main.js hands to the tracer program, un-padded */
struct RefScene {
  const float* bvh;
  const float* tris;
  const float* mats;
  const float* norms;
  const float* uvs;
  const uint8_t* atlas;
  const uint8_t* env;
  const uint16_t* bins;
  int32_t n_nodes, n_tris, atlas_res, atlas_layers, env_w, env_h, n_bins, leaf_size;
};
}

namespace {

template <class Fn>
void parallel_rows(int rows, int nthreads, Fn fn) {
  if (nthreads <= 0) nthreads = (int)std::thread::hardware_concurrency();
  if (nthreads < 1) nthreads = 1;
  if (nthreads > rows) nthreads = rows > 0 ? rows : 1;
  std::atomic<int> next(0);
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; ++t)
    th.emplace_back([&]() {
      bool first = true;
      for (;;) {
        int r = next.fetch_add(1);
        if (r >= rows) break;
        fn(r, first);
        first = false;
      }
    });
  for (auto& x : th) x.join();
}

/* padBuffer, main.js:143-154: the texture that holds `n_floats` floats of `channels`-channel texels, `per_element`
 * texels per record */
void pad_dims(long long n_floats, int per_element, int channels, int& width, int& height) {
  const double num_pixels = (double)n_floats / channels;
  const double root = sqrt(num_pixels);
  width = (int)(ceil(root / per_element) * per_element);
  height = width > 0 ? (int)ceil(num_pixels / width) : 0;
}

uint8_t quant8(float v) { return om::tu_quant8(v); } /* RGBA8 colour-buffer write */

}  // namespace

#define DATA_SAMPLER(S, PTR, N_FLOATS, PER_ELEMENT, CHANNELS)          \
  do {                                                                 \
    (S).f32 = (PTR);                                                   \
    (S).ch = (CHANNELS);                                               \
    (S).valid = (N_FLOATS);                                            \
    pad_dims((N_FLOATS), (PER_ELEMENT), (CHANNELS), (S).w, (S).h);     \
  } while (0)
#define FRAME_SAMPLER(S, PTR, W, H)                                    \
  do {                                                                 \
    (S).f32 = (PTR);                                                   \
    (S).ch = 4;                                                        \
    (S).w = (W);                                                       \
    (S).h = (H);                                                       \
    (S).valid = (long long)(W) * (H) * 4;                              \
  } while (0)

/* ------------------------------------------------------------------------------------------------------------------ */
namespace camera_fs {
#include "glsl_body.inc"
#include "camera.gen.inc"
}  // namespace camera_fs

extern "C" void ref_camera(int W, int H, const float* P, const float* I, float fovScale, const float* lens, float randBase,
                           float* pos4, float* dir4, int nthreads) {
  using namespace camera_fs;
  parallel_rows(H, nthreads, [&](int y, bool first) {
    if (first) { /* gl.uniform* of drawCamera, main.js:728-756 */
      camera_fs::randBase = randBase;
      camera_fs::fovScale = fovScale;
      resolution = vec2((float)W, (float)H);
      lensFeatures = vec2(lens[0], lens[1]);
      camera_fs::P = vec3(P[0], P[1], P[2]);
      camera_fs::I = vec3(I[0], I[1], I[2]);
    }
    for (int x = 0; x < W; ++x) {
      gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.5f, 1.0f);
      /* the `uv` varying: corner.xy of the oversized triangle interpolated at the fragment centre (camera.vs) */
      uv = vec2((gl_FragCoord.x / (float)W) * 2.0f - 1.0f, (gl_FragCoord.y / (float)H) * 2.0f - 1.0f);
      shader_main();
      const size_t k = ((size_t)y * W + x) * 4;
      pos4[k] = fragColor[0].x; pos4[k + 1] = fragColor[0].y; pos4[k + 2] = fragColor[0].z; pos4[k + 3] = fragColor[0].w;
      dir4[k] = fragColor[1].x; dir4[k + 1] = fragColor[1].y; dir4[k + 2] = fragColor[1].z; dir4[k + 3] = fragColor[1].w;
    }
  });
}

/* ------------------------------------------------------------------------------------------------------------------ */
namespace tracer_fs {
#include "glsl_body.inc"
#include "tracer.gen.inc"
}  // namespace tracer_fs

/* one drawTracer() pass, main.js:758-807: fb_out = tracer.fs main() per fragment, fbTex = fb_prev (zeros at tick 0) */
extern "C" void ref_trace(const RefScene* s, const float* pos4, const float* dir4, int W, int H, uint32_t tick,
                          float randBase, float envTheta, const float* fb_prev, float* fb_out, int nthreads) {
  using namespace tracer_fs;
  std::vector<float> zeros;
  if (!fb_prev) { zeros.assign((size_t)W * H * 4, 0.0f); fb_prev = zeros.data(); }
  parallel_rows(H, nthreads, [&](int y, bool first) {
    if (first) {
      tracer_fs::tick = tick;
      tracer_fs::randBase = randBase;
      tracer_fs::envTheta = envTheta;
      numLights = 0.0f;
      glsl_env_bins = s->n_bins;
      glsl_leaf_size = s->leaf_size;
      radianceBins.v.resize(s->n_bins);
      for (int i = 0; i < s->n_bins; ++i)
        radianceBins.v[i] = uvec4{s->bins[4 * i], s->bins[4 * i + 1], s->bins[4 * i + 2], s->bins[4 * i + 3]};
      DATA_SAMPLER(bvhTex, s->bvh, 9LL * s->n_nodes, 3, 3);  /* main.js:408-437 */
      DATA_SAMPLER(matTex, s->mats, 12LL * s->n_tris, 4, 3);
      DATA_SAMPLER(triTex, s->tris, 9LL * s->n_tris, 3, 3);
      DATA_SAMPLER(normTex, s->norms, 27LL * s->n_tris, 9, 3);
      DATA_SAMPLER(uvTex, s->uvs, 6LL * s->n_tris, 3, 2);
      envTex.u8 = s->env; envTex.w = s->env_w; envTex.h = s->env_h; envTex.ch = 4;
      texArray.u8 = s->atlas; texArray.res = s->atlas_res; texArray.layers = s->atlas_layers;
      FRAME_SAMPLER(fbTex, fb_prev, W, H);
      FRAME_SAMPLER(cameraPosTex, pos4, W, H);
      FRAME_SAMPLER(cameraDirTex, dir4, W, H);
    }
    for (int x = 0; x < W; ++x) {
      gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.5f, 1.0f);
      shader_main();
      const size_t k = ((size_t)y * W + x) * 4;
      fb_out[k] = fragColor.x; fb_out[k + 1] = fragColor.y; fb_out[k + 2] = fragColor.z; fb_out[k + 3] = fragColor.w;
    }
  });
}

/* ------------------------------------------------------------------------------------------------------------------ */
namespace bvh_test_fs {
#include "glsl_body.inc"
#include "bvh_test.gen.inc"
}  // namespace bvh_test_fs

/* bvh_test.fs's intersectScene (:173-221) called per ray: (result.index, result.t, count) -- what the shader's main()
 * folds into one heat-map colour (:224-232).  `heat` (optional, n x 4) receives that colour for tick 0 as well. */
extern "C" void ref_bvh_test(const RefScene* s, const float* pos4, const float* dir4, int n, int32_t* index, float* t,
                             int32_t* count, float* heat, int nthreads) {
  using namespace bvh_test_fs;
  const int chunk = 1024;
  const int rows = (n + chunk - 1) / chunk;
  std::vector<float> zeros((size_t)n * 4, 0.0f);
  parallel_rows(rows, nthreads, [&](int r, bool first) {
    if (first) {
      bvh_test_fs::tick = 0;
      bvh_test_fs::randBase = 0.0f;
      glsl_leaf_size = s->leaf_size;
      DATA_SAMPLER(bvhTex, s->bvh, 9LL * s->n_nodes, 3, 3);
      DATA_SAMPLER(triTex, s->tris, 9LL * s->n_tris, 3, 3);
      FRAME_SAMPLER(fbTex, zeros.data(), n, 1);  /* the rays as an n x 1 frame */
      FRAME_SAMPLER(cameraPosTex, pos4, n, 1);
      FRAME_SAMPLER(cameraDirTex, dir4, n, 1);
    }
    const int lo = r * chunk, hi = n < lo + chunk ? n : lo + chunk;
    for (int i = lo; i < hi; ++i) {
      Ray ray;
      ray.origin = vec3(pos4[4 * (size_t)i], pos4[4 * (size_t)i + 1], pos4[4 * (size_t)i + 2]);
      ray.dir = vec3(dir4[4 * (size_t)i], dir4[4 * (size_t)i + 1], dir4[4 * (size_t)i + 2]);
      int c = 0;
      const Hit h = intersectScene(ray, c);
      index[i] = h.index;
      t[i] = h.t;
      count[i] = c;
      if (heat) {
        gl_FragCoord = vec4((float)i + 0.5f, 0.5f, 0.5f, 1.0f);
        shader_main();
        heat[4 * (size_t)i] = fragColor.x; heat[4 * (size_t)i + 1] = fragColor.y;
        heat[4 * (size_t)i + 2] = fragColor.z; heat[4 * (size_t)i + 3] = fragColor.w;
      }
    }
  });
}

/* ------------------------------------------------------------------------------------------------------------------ */
namespace draw_fs {
#include "glsl_body.inc"
#include "draw.gen.inc"
}  // namespace draw_fs
#undef INV_PI
#undef INV_SQRT_OF_2PI

/* draw.fs main() per fragment (main.js:809-826); rgba8 = the canvas the colour lands in, out4 (optional) the f32
 * value before the fixed-function conversion */
extern "C" void ref_draw(const float* fb, int W, int H, float exposure, float saturation, int denoise, float maxSigma,
                         float scale, uint8_t* rgba8, float* out4, int nthreads) {
  using namespace draw_fs;
  parallel_rows(H, nthreads, [&](int y, bool first) {
    if (first) {
      FRAME_SAMPLER(fbTex, fb, W, H);
      draw_fs::exposure = exposure;
      draw_fs::saturation = saturation;
      draw_fs::scale = scale;
      draw_fs::maxSigma = maxSigma;
      draw_fs::denoise = denoise != 0;
    }
    for (int x = 0; x < W; ++x) {
      gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.5f, 1.0f);
      shader_main();
      const size_t k = ((size_t)y * W + x) * 4;
      rgba8[k] = quant8(fragColor.x); rgba8[k + 1] = quant8(fragColor.y); rgba8[k + 2] = quant8(fragColor.z);
      rgba8[k + 3] = quant8(fragColor.w);
      if (out4) { out4[k] = fragColor.x; out4[k + 1] = fragColor.y; out4[k + 2] = fragColor.z; out4[k + 3] = fragColor.w; }
    }
  });
}

/* ------------------------------------------------------------------------------------------------------------------ */
namespace blit_fs { /* the fragment shader inside texture_packer.js (:103-121, template string fsStr) */
#include "glsl_body.inc"
#include "blit.gen.inc"
}  // namespace blit_fs

/* One image layer of TexturePacker.getPixels(): setAndDrawTexture(img) (texture_packer.js:159-176) + readPixels (:178-184).
 * rgba8: w*h*4, row 0 = image top (texImage2D without FLIP_Y: texture row 0).  out: res*res*4, row y = gl_FragCoord.y. */
extern "C" void ref_pack_layer(const uint8_t* rgba8, int w, int h, int res, int corrected, const int32_t* swz, uint8_t* out) {
  using namespace blit_fs;
  tex.u8 = rgba8; tex.w = w; tex.h = h; tex.ch = 4; tex.srgb = corrected;
  dims = vec2((float)res, (float)res);
  swizzle = swz ? uvec4{(uint)swz[0], (uint)swz[1], (uint)swz[2], (uint)swz[3]} : uvec4{0u, 1u, 2u, 3u}; /* :173 */
  for (int y = 0; y < res; ++y)
    for (int x = 0; x < res; ++x) {
      gl_FragCoord = vec4((float)x + 0.5f, (float)y + 0.5f, 0.5f, 1.0f);
      shader_main();
      uint8_t* o = out + ((size_t)y * res + x) * 4;
      o[0] = quant8(fragColor.x); o[1] = quant8(fragColor.y); o[2] = quant8(fragColor.z); o[3] = quant8(fragColor.w);
    }
}

extern "C" int ref_abi_version() { return 2; }
