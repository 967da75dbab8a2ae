// fspt_api.cu -- context, scene upload (layout repack), wavefront render loop and the extern "C" ABI of
// libfspt_b200.so (include/fspt_b200.h).  Host side of what main.js does between initBVH() and tick():
// texImage uploads (main.js:408-437,548-560,170-180), drawCamera/drawTracer/drawQuad (main.js:741-824).
// Map: Ctx (state of one context) -- stage_atlas (the atlas part of an upload: runs inline or on the context's atlas
// thread) -- launch_trace / render_wave (one wave of the wavefront) -- NCCL loaded at run time, broadcast_phase2 (the
// deferred atlas part of fspt_scene_broadcast) -- extern "C": create / destroy, scene_upload_impl (pre-passes and record
// builders live in scene_pack.h, the worker pool in host_pool.h), render, resolve, accumulation access, debug entry
// points, tiles and collectives, parameters and statistics.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is resolved with dlopen at fspt_comm_init, there is no link-time dependency
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__SSE2__)
#include <emmintrin.h>
#endif
#if defined(__SSSE3__)
#include <tmmintrin.h>
#endif
#include <algorithm>
#include <array>
#include <atomic>
#include <map>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/fspt_b200.h"
#include "device_common.cuh"
#include "host_pool.h"
#include "scene_pack.h"
#include "shade.cuh"
#include "traverse.cuh"

namespace {

thread_local std::string g_create_error;

struct Ctx {
  int device = 0, sm_count = 0;
  int width = 0, height = 0, n_pixels = 0;
  int rx0 = 0, ry0 = 0, rw = 0, rh = 0;  // the pixel rectangle this context renders (fspt_set_tile); default = frame
  cudaStream_t stream = nullptr;
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  std::vector<cudaEvent_t> ev_trace;  // pairs around traversal (tag 0) / shading (tag 1) launches of the last render
  std::vector<int> ev_tag;
  size_t ev_trace_used = 0;
  bool render_timed = false;
  std::string error;

  // scene
  bool has_scene = false, has_dielectric = false;
  DeviceScene sc{};
  void *d_nodes = nullptr, *d_tris = nullptr, *d_shade = nullptr, *d_bins = nullptr, *d_layer_info = nullptr;
  cudaArray_t atlas_arr = nullptr, env_arr = nullptr, mat_arr = nullptr;
  int mat_R = 0, mat_L = 0;
  bool mat_surface = false;           // mat_arr was created with cudaArraySurfaceLoadStore (GPU-side interleave)
  cudaSurfaceObject_t mat_surf = 0;   // write view of mat_arr for k_interleave_atlas
  void* d_raw = nullptr;              // distinct varying atlas layers as uploaded (RGBA8), source of the GPU interleave
  void* d_mat_src = nullptr;          // MatSrc per textured material
  size_t cap_raw = 0, cap_mat_src = 0;
  void* d_mat_info = nullptr;
  size_t cap_mat_info = 0;
  cudaTextureObject_t nodes_tex = 0;  // the node array again as a linear texture (second L1 data pipe)
  int atlas_R = 0, atlas_L = 0, env_W = 0, env_H = 0;       // dims of the resident arrays (reused across uploads)
  size_t cap_nodes = 0, cap_tris = 0, cap_shade = 0, cap_bins = 0, cap_layer_info = 0;
  uint8_t* h_stage = nullptr;                               // pinned staging for the atlas (one slot per layer)
  cudaStream_t copy_stream = nullptr;                       // atlas DMA, overlaps the primary traversal
  cudaEvent_t ev_atlas = nullptr;
  // fspt_scene_upload_async: the atlas part of the upload (constant-layer scan, host interleave into pinned memory,
  // band-wise DMA) runs on this thread after the call has returned; joined before the first shading launch and by
  // everything that touches scene state
  std::thread atlas_thread;
  // page-locked sources are DMA'd in place: they stay borrowed until the copies have completed, not just been enqueued
  bool env_in_place = false, atlas_in_place = false;
  cudaEvent_t ev_env = nullptr;  // behind the environment copies on the context's stream (recorded when env_in_place)
  std::atomic<uint64_t> atlas_launches{0};  // kernels launched by the atlas part (folded into stats.kernel_launches)
  int atlas_rc = 0;            // outcome of the atlas part (written by the atlas thread, read after joining it)
  std::string atlas_error;
  HostPool* pool = nullptr;    // host workers of fspt_scene_upload (created at the first upload)
  uint8_t* h_ring = nullptr;   // pinned ring the geometry records are staged through, chunk by chunk
  size_t ring_bytes = 0;
  std::vector<cudaEvent_t> ev_ring;  // per ring slot: the copies that read it have completed
  uint8_t* h_geo = nullptr;                                 // pinned block of the small tables (bins, layer / material tables) + env
  size_t geo_stage_bytes = 0;
  size_t stage_bytes = 0;
  size_t scene_bytes = 0;

  // frame state
  float4* d_fb = nullptr;          // accumulation target (RGBA32F)
  float4* d_last_color = nullptr;
  float4* d_sample_color = nullptr;
  float4* d_cam_pos = nullptr;     // debug_primary only
  float4* d_cam_dir = nullptr;
  uchar4* d_rgba8 = nullptr;
  uint32_t next_tick = 0;
  uint64_t accum_samples = 0;
  int accum_mode = 0;
  int sanitize = 1;
  int max_refractions = 64;
  int anyhit = 1;

  // wavefront state
  int wave_samples = 0;       // samples in flight per wave (whole frame)
  int wave_cap = 64;
  size_t wave_paths = 0;
  PathState ps{};             // = ps2[0] (debug entry points)
  PathState ps2[2]{};         // dense path-record arrays, ping-pong: shade reads one, compacts survivors into the other
  float4* d_shadow = nullptr; // shadow rays of the current bounce, dense: origin | record position, direction | -
  int* d_counts = nullptr;    // set k at [2k, 2k+1] = (#continuation, #shadow), [32] fetch cursor (own cache line)
  int* d_count_out = nullptr; // per-slot visit count (debug)
  unsigned char* d_hit_flag = nullptr;  // hit / miss per record position
  unsigned long long* d_stats = nullptr;  // rays, nodes, leaves, capped
  float* d_rb = nullptr;      // rand bases of one render call: [0,cap) camera, [cap,2cap) tracer
  // pinned staging, a ring of RB_SLOTS blocks of 2*rb_cap floats: a slot is rewritten only after the copy that last
  // read it has completed (its event), so fspt_render never waits for the GPU to drain
  static constexpr int RB_SLOTS = 8;
  float* h_rb = nullptr;
  cudaEvent_t ev_rb[RB_SLOTS] = {};
  int rb_slot = 0;
  int rb_cap = 0;
  // refraction bounces beyond NUM_BOUNCES (tracer.fs:488): path counts polled two iterations behind the enqueue
  static constexpr int POLL_SLOTS = 8;
  int* h_poll = nullptr;      // pinned
  cudaEvent_t ev_poll[POLL_SLOTS] = {};
  // collectives (fspt_comm_init): NCCL communicator over NVLink / NVSwitch
  ncclComm_t comm = nullptr;
  int comm_rank = 0, comm_world = 1;
  cudaEvent_t ev_red0 = nullptr, ev_red1 = nullptr;  // around the last fspt_reduce_accum
  bool reduce_timed = false;
  void* d_lin = nullptr;      // linear staging of the environment + atlas arrays for fspt_scene_broadcast
  size_t cap_lin = 0;
  void* d_hdr = nullptr;      // broadcast headers
  void* d_lin_env = nullptr;  // linear staging of the environment array (phase 1 of the broadcast)
  size_t cap_lin_env = 0;
  int bcast_pending_root = -1;  // >= 0: phase 2 of fspt_scene_broadcast (the atlas) has not been issued yet
  cudaEvent_t ev_bcast = nullptr;  // end of phase 1 on the context's stream
  size_t bytes_nodes = 0, bytes_tris = 0, bytes_shade = 0, bytes_bins = 0, bytes_layer_info = 0,
         bytes_mat_info = 0;
  int trace_blocks = 0, trace_blocks_cnt = 0, trace_blocks_cam = 0, shade_blocks = 0;
  int trace_blocks_nt = 0, trace_blocks_cnt_nt = 0, trace_blocks_cam_nt = 0;  // NODE_TEX = false instantiations

  fspt_stats stats{};
};

void comm_free(Ctx* c);
int broadcast_phase2(Ctx* c);

int fail(Ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->error = buf; else g_create_error = buf;
  return code;
}

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) return fail(c, FSPT_E_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

template <class T>
void dfree(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

void free_scene(Ctx* c) {
  if (c->sc.atlas) cudaDestroyTextureObject(c->sc.atlas);
  if (c->sc.env) cudaDestroyTextureObject(c->sc.env);
  if (c->nodes_tex) cudaDestroyTextureObject(c->nodes_tex);
  c->nodes_tex = 0;
  if (c->sc.mat_tex) cudaDestroyTextureObject(c->sc.mat_tex);
  if (c->mat_surf) cudaDestroySurfaceObject(c->mat_surf);
  c->mat_surf = 0; c->mat_surface = false;
  dfree(c->d_raw); dfree(c->d_mat_src); c->cap_raw = c->cap_mat_src = 0;
  if (c->atlas_arr) cudaFreeArray(c->atlas_arr);
  if (c->env_arr) cudaFreeArray(c->env_arr);
  if (c->mat_arr) cudaFreeArray(c->mat_arr);
  c->atlas_arr = c->env_arr = c->mat_arr = nullptr;
  c->mat_R = c->mat_L = 0;
  dfree(c->d_mat_info); c->cap_mat_info = 0;
  c->atlas_R = c->atlas_L = c->env_W = c->env_H = 0;
  dfree(c->d_nodes); dfree(c->d_tris); dfree(c->d_shade); dfree(c->d_bins); dfree(c->d_layer_info);
  c->cap_nodes = c->cap_tris = c->cap_shade = c->cap_bins = c->cap_layer_info = 0;
  if (c->h_stage) cudaFreeHost(c->h_stage);
  c->h_stage = nullptr; c->stage_bytes = 0;
  if (c->h_geo) cudaFreeHost(c->h_geo);
  c->h_geo = nullptr; c->geo_stage_bytes = 0;
  c->sc = DeviceScene{};
  c->has_scene = false;
}

// device buffer that is only re-allocated when it has to grow
int ensure(Ctx* c, void*& p, size_t& cap, size_t bytes) {
  if (p && cap >= bytes) return FSPT_OK;
  if (p) cudaFree(p);
  p = nullptr; cap = 0;
  CK(cudaMalloc(&p, bytes));
  cap = bytes;
  return FSPT_OK;
}

int alloc_wave(Ctx* c) {
  // samples in flight: traversal launches amortise their ramp/tail over tens of millions of rays (measured at
  // 1280x720: 4 -> 16 -> 32 -> 64 samples per wave = +14 % -> +3 % -> +2.7 %); 64 M paths x (2 x 80 B records + lists)
  // = ~14 GB of 180 GB
  const size_t target_paths = (size_t)64 << 20;
  int S = (int)std::max<size_t>(1, target_paths / (size_t)c->n_pixels);
  S = std::min(S, 64);
  c->wave_cap = 64;
  if (const char* e = getenv("FSPT_WAVE_SAMPLES")) c->wave_cap = S = std::max(1, std::min(64, atoi(e)));  // tuning knob
  c->wave_samples = S;
  c->wave_paths = (size_t)S * c->n_pixels;
  const size_t W = c->wave_paths;
  for (int k = 0; k < 2; ++k) {
    CK(cudaMalloc(&c->ps2[k].rec, W * 16 * FSPT_PATH_WORDS));
    CK(cudaMalloc(&c->ps2[k].sh, W));
  }
  c->ps = c->ps2[0];
  CK(cudaMalloc(&c->d_shadow, W * 32));
  CK(cudaMalloc(&c->d_counts, 64 * sizeof(int)));  // [0..3] the two count pairs; [32] the fetch cursor, on its own 128-byte
                                                   // line: every warp's atomicAdd hits it, nothing else should
  CK(cudaMemset(c->d_counts, 0, 64 * sizeof(int)));
  CK(cudaMalloc(&c->d_count_out, W * sizeof(int)));
  CK(cudaMalloc(&c->d_hit_flag, W));
  CK(cudaMalloc(&c->d_stats, 8 * sizeof(unsigned long long)));  // [0..3] live, [4..7] snapshot at render start
  CK(cudaMemset(c->d_stats, 0, 8 * sizeof(unsigned long long)));
  CK(cudaMalloc(&c->d_sample_color, W * 16));
  c->rb_cap = 64;
  CK(cudaMalloc(&c->d_rb, 2 * (size_t)c->rb_cap * sizeof(float)));
  CK(cudaMallocHost(&c->h_rb, Ctx::RB_SLOTS * 2 * (size_t)c->rb_cap * sizeof(float)));
  for (auto& e : c->ev_rb) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  CK(cudaMallocHost(&c->h_poll, Ctx::POLL_SLOTS * sizeof(int)));
  for (auto& e : c->ev_poll) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  return FSPT_OK;
}

// ---- traversal launch helpers ---------------------------------------------------------------------------
void record_trace_begin(Ctx* c, int tag = 0) {
  if (c->ev_trace_used + 2 > c->ev_trace.size()) {
    for (int i = 0; i < 2; ++i) { cudaEvent_t e; cudaEventCreate(&e); c->ev_trace.push_back(e); }
  }
  if (c->ev_tag.size() < c->ev_trace.size() / 2) c->ev_tag.resize(c->ev_trace.size() / 2);
  c->ev_tag[c->ev_trace_used / 2] = tag;
  cudaEventRecord(c->ev_trace[c->ev_trace_used], c->stream);
}
void record_trace_end(Ctx* c) {
  cudaEventRecord(c->ev_trace[c->ev_trace_used + 1], c->stream);
  c->ev_trace_used += 2;
}

// One texel of a material layer = the RGBA8 texels of its four maps (device_common.cuh "MatTexel"), each taken from an
// uploaded raw layer or, for a colour layer, from its constant.
struct MatSrc { int raw[4]; unsigned cst[4]; };
// RGB24: the raw layers travelled without their alpha channel (3 bytes per texel) -- tracer.fs:453-456 reads .rgb / .rg of
// the four maps, never .a, so alpha cannot influence an image and is not worth a quarter of the PCIe time.
template <bool RGB24>
__global__ void __launch_bounds__(256) k_interleave_atlas(cudaSurfaceObject_t surf, const uint8_t* __restrict__ raw,
                                                          const MatSrc* __restrict__ src, int R, size_t layer_texels) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y * blockDim.y + threadIdx.y;
  if (x >= R || y >= R) return;
  const MatSrc m = src[blockIdx.z];
  const size_t i = (size_t)y * R + x;
  auto fetch = [&](int k) -> unsigned {
    if (m.raw[k] < 0) return m.cst[k];
    if (RGB24) {
      const uint8_t* p = raw + ((size_t)m.raw[k] * layer_texels + i) * 3;
      return (unsigned)p[0] | ((unsigned)p[1] << 8) | ((unsigned)p[2] << 16);
    }
    return reinterpret_cast<const uint32_t*>(raw)[(size_t)m.raw[k] * layer_texels + i];
  };
  uint4 v;
  v.x = fetch(0); v.y = fetch(1); v.z = fetch(2); v.w = fetch(3);
  surf2DLayeredwrite(v, surf, x * 16, y, blockIdx.z);
}

__global__ void k_set_counts(int* counts, int n_cont, int n_shadow) {
  counts[0] = n_cont; counts[1] = n_shadow; counts[2] = 0; counts[3] = 0; counts[32] = 0;
}

// true when `p` points into page-locked host memory (cudaHostRegister / cudaMallocHost, by whoever): DMA can read it in place
bool host_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

// Page-locked sources (fspt_host_register) are read by the copy engines where they lie: the caller gets them back when
// those copies have completed.  (Staged sources were copied out before the upload returned / the atlas thread ended.)
int release_borrowed(Ctx* c) {
  if (c->env_in_place) { CK(cudaEventSynchronize(c->ev_env)); c->env_in_place = false; }
  if (c->atlas_in_place) { CK(cudaEventSynchronize(c->ev_atlas)); c->atlas_in_place = false; }
  return FSPT_OK;
}

// Waits for the atlas part of the last fspt_scene_upload_async (no-op otherwise) and reports its outcome.
int atlas_join(Ctx* c) {
  if (c->atlas_thread.joinable()) c->atlas_thread.join();
  const int rc = c->atlas_rc;
  c->atlas_rc = FSPT_OK;
  if (rc) { c->has_scene = false; c->error = c->atlas_error; }
  return rc;
}

// ---- the atlas part of a scene upload (main.js:548-560) ------------------------------------------------------
// Constant-colour layers are detected (exact: every texel compared), then the atlas is re-interleaved per material into
// 16-byte texels (device_common.cuh "MatTexel") in pinned memory by the host workers and DMA'd band by band as the bands
// land; a scene with so many layer combinations that this would not fit falls back to the plain RGBA8 layered array.
// Everything travels on the copy stream and ends with ev_atlas: only k_shade reads what this function uploads, so a
// render's camera + primary traversal launch overlaps it.  Runs inline (fspt_scene_upload) or on the context's atlas
// thread (fspt_scene_upload_async); every input is held by value or lives in the context's pinned blocks.
struct AtlasJob {
  const uint8_t* atlas;                     // caller's layers, RGBA8, L x R x R
  int L, R;
  std::vector<std::array<int, 4>> mats;     // distinct layer quadruples (diffuse, emission, metallic-roughness, normal)
  uint32_t* layer_info;                     // pinned: per layer {constant?, first texel}
  int32_t* mat_info;                        // pinned: per material 2 x int4
  size_t n_mat_info;                        // ints
  MatSrc* mat_src;                          // pinned: per textured material (GPU-side interleave)
  std::vector<uint8_t> varied;               // per layer: 1 = not a constant colour (scanned next to the geometry staging);
                                             // empty: the atlas part scans the layers itself
  bool src_pinned;                           // the caller's atlas is page-locked (fspt_host_register): no staging copy
  int workers;
  bool timing;
};

// constant-colour layers: constant <=> every texel equals its successor (exact: every texel compared); work item =
// (layer, band), abandoned once the layer is known to vary
constexpr int SCAN_BANDS = 8;
inline void scan_layer_band(const uint8_t* atlas, size_t layer_texels, int item, std::atomic<int>* varied) {
  const int l = item / SCAN_BANDS, band = item % SCAN_BANDS;
  if (varied[l].load(std::memory_order_relaxed)) return;
  const uint32_t* px = reinterpret_cast<const uint32_t*>(atlas) + (size_t)l * layer_texels;
  const size_t i0 = layer_texels * band / SCAN_BANDS, i1 = std::min(layer_texels - 1, layer_texels * (band + 1) / SCAN_BANDS);
  for (size_t i = i0; i < i1;) {
    const size_t n = std::min<size_t>(i1 - i, 16384);
    if (memcmp(px + i, px + i + 1, n * 4) != 0) { varied[l].store(1, std::memory_order_relaxed); return; }
    i += n;
  }
}

int stage_atlas(Ctx* c, const AtlasJob& J) {
  auto afail = [&](int code, const char* what, cudaError_t e) {
    char buf[256];
    snprintf(buf, sizeof buf, "atlas upload: %s failed: %s", what, cudaGetErrorString(e));
    c->atlas_error = buf;
    return code;
  };
#define ACK(call)                                                         \
  do {                                                                    \
    cudaError_t e_ = (call);                                              \
    if (e_ != cudaSuccess) return afail(FSPT_E_CUDA, #call, e_);          \
  } while (0)
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!J.timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[fspt upload]   atlas: %-24s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  const int L = J.L, R = J.R;
  const size_t layer_texels = (size_t)R * R, layer_bytes = layer_texels * 4;
  const std::vector<std::array<int, 4>>& mats = J.mats;
  uint32_t* layer_info = J.layer_info;
  int32_t* mat_info = J.mat_info;
  HostPool& pool = *c->pool;
  {
    std::vector<std::atomic<int>> varied((size_t)L);
    if (J.varied.empty()) {
      for (auto& v : varied) v.store(0);
      pool.run(L * SCAN_BANDS, J.workers, [&](int item) { scan_layer_band(J.atlas, layer_texels, item, varied.data()); });
      lap("constant-layer scan");
    }
    for (int l = 0; l < L; ++l) {
      layer_info[2 * l] = (J.varied.empty() ? varied[l].load() : J.varied[l]) ? 0u : 1u;
      memcpy(&layer_info[2 * l + 1], J.atlas + (size_t)l * layer_bytes, 4);
    }
  }
  memset(mat_info, 0, J.n_mat_info * 4);
  int n_tex_mats = 0;
  for (size_t m = 0; m < mats.size(); ++m) {
    bool all_const = true;
    for (int k = 0; k < 4; ++k) all_const = all_const && layer_info[2 * mats[m][k]];
    mat_info[8 * m] = all_const ? -1 : n_tex_mats++;
    for (int k = 0; k < 4; ++k) mat_info[8 * m + 1 + k] = (int32_t)layer_info[2 * mats[m][k] + 1];
  }
  // The interleaved atlas costs res^2 * 16 bytes per TEXTURED MATERIAL (distinct layer quadruple) and the same again in
  // pinned staging; quadruples can outnumber layers, so it is bounded against the plain atlas (4 x its bytes, at least
  // 256 MB) and against the free device memory, and any allocation failure falls back to the plain RGBA8 array.
  bool use_mat_tex = !getenv("FSPT_PLAIN_ATLAS");
  {
    const size_t inter = (size_t)n_tex_mats * layer_texels * 16, plain = (size_t)L * layer_bytes;
    // the array of the previous upload is reused when it has the same shape: then nothing is allocated and the driver is
    // not asked for the free memory (cudaMemGetInfo is a resource-manager call: usually 0.1 ms, but 20-100 ms every few
    // dozen calls on the virtualised hosts measured, which showed up as spikes in the end-to-end step time)
    const bool reuse = c->mat_arr && c->mat_R == R && c->mat_L == std::max(1, n_tex_mats);
    size_t free_b = (size_t)48 << 30, total_b = 0;
    if (!reuse && cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { (void)cudaGetLastError(); free_b = (size_t)48 << 30; }
    const size_t resident = c->mat_arr ? (size_t)c->mat_R * c->mat_R * 16 * (size_t)c->mat_L : 0;  // freed before the new one
    if (inter > std::max<size_t>(4 * plain, (size_t)256 << 20) || inter > ((size_t)48 << 30) ||
        (!reuse && inter > (free_b + resident) / 2))
      use_mat_tex = false;
    if (getenv("FSPT_FORCE_MAT_TEX")) use_mat_tex = true;  // test knob: exercise the allocation-failure fallback
  }
  lap("tables + sizing");
  cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeArray;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  std::atomic<int> cuda_err(0);
  std::mutex mu;  // serialises the enqueues of the staging items
  bool plain_atlas = !use_mat_tex;
  if (use_mat_tex) {
    if (c->sc.atlas) { cudaDestroyTextureObject(c->sc.atlas); c->sc.atlas = 0; }
    if (c->atlas_arr) { cudaFreeArray(c->atlas_arr); c->atlas_arr = nullptr; c->atlas_R = c->atlas_L = 0; }
    const int ML = std::max(1, n_tex_mats);
    // Where the interleave runs.  On the host (default): the 16-byte texels are built in pinned memory and DMA'd.  On the
    // GPU: only the DISTINCT varying layers cross PCIe as raw RGBA8 and k_interleave_atlas builds the texels through a
    // surface -- a material with one image map and three colours then costs 1 layer of staging and DMA instead of 4.
    // Measured break-even (bench scene, 7 varying layers behind 2 materials: GPU path 20 % slower): chosen when the raw
    // layers are at most half of the interleaved bytes.  FSPT_ATLAS_INTERLEAVE=gpu|cpu overrides.
    std::vector<int> raw_of((size_t)L, -1), raw_layers;
    MatSrc* mat_src = J.mat_src;
    for (size_t m = 0; m < mats.size(); ++m) {
      const int tl = mat_info[8 * m];
      if (tl < 0) continue;
      for (int k = 0; k < 4; ++k) {
        const int l = mats[m][k];
        if (!layer_info[2 * l] && raw_of[l] < 0) { raw_of[l] = (int)raw_layers.size(); raw_layers.push_back(l); }
        mat_src[tl].raw[k] = layer_info[2 * l] ? -1 : raw_of[l];
        mat_src[tl].cst[k] = layer_info[2 * l + 1];
      }
    }
    // A page-locked source (the host registered its atlas, fspt_host_register) needs no staging at all on the GPU path:
    // the distinct varying layers are DMA'd from where they lie, the host does nothing but enqueue the copies.
    bool gpu_interleave = n_tex_mats > 0 && (J.src_pinned || raw_layers.size() * 2 <= (size_t)n_tex_mats * 4);
    if (const char* e = getenv("FSPT_ATLAS_INTERLEAVE")) gpu_interleave = n_tex_mats > 0 && !strcmp(e, "gpu");
    const bool direct = gpu_interleave && J.src_pinned && !getenv("FSPT_ATLAS_NO_DIRECT");
    c->atlas_in_place = direct;
    // wire format of the staged GPU path: RGB, 3 bytes per texel (the shaders never read alpha, tracer.fs:453-456)
    const bool rgb24 = gpu_interleave && !direct && !getenv("FSPT_ATLAS_RGBA_WIRE");
    const size_t wire_layer_bytes = rgb24 ? layer_texels * 3 : layer_bytes;
    if (!c->mat_arr || c->mat_R != R || c->mat_L != ML || c->mat_surface != gpu_interleave) {
      if (c->sc.mat_tex) cudaDestroyTextureObject(c->sc.mat_tex);
      c->sc.mat_tex = 0;
      if (c->mat_surf) cudaDestroySurfaceObject(c->mat_surf);
      c->mat_surf = 0;
      if (c->mat_arr) cudaFreeArray(c->mat_arr);
      c->mat_arr = nullptr;
      cudaChannelFormatDesc fmt4 = cudaCreateChannelDesc<uint4>();
      c->mat_R = c->mat_L = 0;
      const size_t max_ml = getenv("FSPT_FORCE_MAT_TEX") ? (size_t)atoi(getenv("FSPT_FORCE_MAT_TEX")) : (size_t)1 << 30;
      if ((size_t)ML > max_ml ||  // (test knob: pretend the device cannot hold more than that many material layers)
          cudaMalloc3DArray(&c->mat_arr, &fmt4, make_cudaExtent(R, R, ML),
                            cudaArrayLayered | (gpu_interleave ? cudaArraySurfaceLoadStore : 0)) != cudaSuccess) {
        (void)cudaGetLastError();
        c->mat_arr = nullptr;
        plain_atlas = true;
      } else {
        rd.res.array.array = c->mat_arr;
        ACK(cudaCreateTextureObject(&c->sc.mat_tex, &rd, &td, nullptr));
        if (gpu_interleave) ACK(cudaCreateSurfaceObject(&c->mat_surf, &rd));
        c->mat_R = R; c->mat_L = ML; c->mat_surface = gpu_interleave;
      }
    }
    const size_t need = gpu_interleave ? wire_layer_bytes * raw_layers.size() : layer_texels * 16 * (size_t)ML;
    if (!plain_atlas && !direct && c->stage_bytes < need) {
      if (c->h_stage) cudaFreeHost(c->h_stage);
      c->h_stage = nullptr; c->stage_bytes = 0;
      if (cudaMallocHost(&c->h_stage, need) != cudaSuccess) {
        (void)cudaGetLastError();
        c->h_stage = nullptr;
        plain_atlas = true;
      } else {
        c->stage_bytes = need;
      }
    }
    lap("arrays + pinned block");
    if (plain_atlas) {
      // could not hold the interleaved atlas: every material becomes "plain" (the tables are not consulted by k_shade<false>)
    } else if (gpu_interleave) {
      if (!(c->d_raw && c->cap_raw >= need)) {
        if (c->d_raw) cudaFree(c->d_raw);
        c->d_raw = nullptr; c->cap_raw = 0;
        ACK(cudaMalloc(&c->d_raw, need));
        c->cap_raw = need;
      }
      if (!(c->d_mat_src && c->cap_mat_src >= sizeof(MatSrc) * (size_t)ML)) {
        if (c->d_mat_src) cudaFree(c->d_mat_src);
        c->d_mat_src = nullptr; c->cap_mat_src = 0;
        ACK(cudaMalloc(&c->d_mat_src, sizeof(MatSrc) * (size_t)ML));
        c->cap_mat_src = sizeof(MatSrc) * (size_t)ML;
      }
      if (direct) {
        for (size_t ri = 0; ri < raw_layers.size(); ++ri)
          ACK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(c->d_raw) + ri * layer_bytes, J.atlas + (size_t)raw_layers[ri] * layer_bytes,
                              layer_bytes, cudaMemcpyHostToDevice, c->copy_stream));
      }
      // work item = (raw layer, band of rows): copy into the pinned block, DMA the band
      const int bands = std::max(1, std::min(R, 16));
      pool.run(direct ? 0 : (int)raw_layers.size() * bands, J.workers, [&](int item) {
        const int ri = item / bands, band = item % bands;
        const size_t y0 = (size_t)R * band / bands, y1 = (size_t)R * (band + 1) / bands;
        const size_t bpt = rgb24 ? 3 : 4;
        const size_t off = (size_t)ri * wire_layer_bytes + y0 * R * bpt, bytes = (y1 - y0) * R * bpt;
        const uint8_t* src = J.atlas + (size_t)raw_layers[ri] * layer_bytes + y0 * R * 4;
        if (!rgb24) {
          memcpy(c->h_stage + off, src, bytes);
        } else {
          uint8_t* dst = c->h_stage + off;
          const size_t n = (y1 - y0) * R;
          size_t i = 0;
#if defined(__SSSE3__)
          // 4 texels (16 bytes) -> 12 bytes per shuffle; the 16-byte store's last 4 bytes are overwritten by the next one
          const __m128i sh = _mm_setr_epi8(0, 1, 2, 4, 5, 6, 8, 9, 10, 12, 13, 14, -1, -1, -1, -1);
          for (; i + 8 <= n; i += 4)
            _mm_storeu_si128(reinterpret_cast<__m128i*>(dst + 3 * i),
                             _mm_shuffle_epi8(_mm_loadu_si128(reinterpret_cast<const __m128i*>(src + 4 * i)), sh));
#endif
          for (; i < n; ++i) { dst[3 * i] = src[4 * i]; dst[3 * i + 1] = src[4 * i + 1]; dst[3 * i + 2] = src[4 * i + 2]; }
        }
        std::lock_guard<std::mutex> g(mu);
        cudaError_t e = cudaMemcpyAsync(reinterpret_cast<uint8_t*>(c->d_raw) + off, c->h_stage + off, bytes, cudaMemcpyHostToDevice, c->copy_stream);
        if (e != cudaSuccess) cuda_err.store((int)e);
      });
      ACK(cudaMemcpyAsync(c->d_mat_src, mat_src, sizeof(MatSrc) * (size_t)n_tex_mats, cudaMemcpyHostToDevice, c->copy_stream));
      const dim3 blk(32, 8), grd((R + 31) / 32, (R + 7) / 8, n_tex_mats);
      if (rgb24)
        k_interleave_atlas<true><<<grd, blk, 0, c->copy_stream>>>(c->mat_surf, reinterpret_cast<const uint8_t*>(c->d_raw),
                                                                  reinterpret_cast<const MatSrc*>(c->d_mat_src), R, layer_texels);
      else
        k_interleave_atlas<false><<<grd, blk, 0, c->copy_stream>>>(c->mat_surf, reinterpret_cast<const uint8_t*>(c->d_raw),
                                                                   reinterpret_cast<const MatSrc*>(c->d_mat_src), R, layer_texels);
      c->atlas_launches.fetch_add(1);
      ACK(cudaGetLastError());
    } else {
      // work item = (textured material, band of rows): interleave the four source layers, DMA the band
      const int bands = std::max(1, std::min(R, 16));
      std::vector<int> tex_mat_ids;
      for (size_t m = 0; m < mats.size(); ++m) if (mat_info[8 * m] >= 0) tex_mat_ids.push_back((int)m);
      pool.run((int)tex_mat_ids.size() * bands, J.workers, [&](int item) {
        const int m = tex_mat_ids[item / bands], band = item % bands;
        const int tl = mat_info[8 * m];
        const int y0 = (int)((long long)R * band / bands), y1 = (int)((long long)R * (band + 1) / bands);
        const uint32_t* src[4];
        for (int k = 0; k < 4; ++k) src[k] = reinterpret_cast<const uint32_t*>(J.atlas) + (size_t)mats[m][k] * layer_texels;
        uint32_t* dst = reinterpret_cast<uint32_t*>(c->h_stage) + ((size_t)tl * layer_texels + (size_t)y0 * R) * 4;
        const size_t i0 = (size_t)y0 * R, n = (size_t)(y1 - y0) * R;
        size_t i = 0;
#if defined(__SSE2__)
        // 4x4 transpose of 32-bit texels, written with non-temporal stores: the staging block is only read by the DMA
        // engine, so it should neither be fetched for ownership nor displace the source layers from the CPU caches
        for (; i + 4 <= n; i += 4) {
          const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[0] + i0 + i));
          const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[1] + i0 + i));
          const __m128i c2 = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[2] + i0 + i));
          const __m128i d = _mm_loadu_si128(reinterpret_cast<const __m128i*>(src[3] + i0 + i));
          const __m128i ab_lo = _mm_unpacklo_epi32(a, b), ab_hi = _mm_unpackhi_epi32(a, b);
          const __m128i cd_lo = _mm_unpacklo_epi32(c2, d), cd_hi = _mm_unpackhi_epi32(c2, d);
          __m128i* o = reinterpret_cast<__m128i*>(dst + 4 * i);  // 16-byte aligned: pinned block + multiples of 16
          _mm_stream_si128(o + 0, _mm_unpacklo_epi64(ab_lo, cd_lo));
          _mm_stream_si128(o + 1, _mm_unpackhi_epi64(ab_lo, cd_lo));
          _mm_stream_si128(o + 2, _mm_unpacklo_epi64(ab_hi, cd_hi));
          _mm_stream_si128(o + 3, _mm_unpackhi_epi64(ab_hi, cd_hi));
        }
        _mm_sfence();
#endif
        for (; i < n; ++i) {
          dst[4 * i + 0] = src[0][i0 + i]; dst[4 * i + 1] = src[1][i0 + i];
          dst[4 * i + 2] = src[2][i0 + i]; dst[4 * i + 3] = src[3][i0 + i];
        }
        cudaMemcpy3DParms cp = {};
        cp.srcPtr = make_cudaPitchedPtr(dst, (size_t)R * 16, R, y1 - y0);
        cp.dstArray = c->mat_arr;
        cp.dstPos = make_cudaPos(0, y0, tl);
        cp.extent = make_cudaExtent(R, y1 - y0, 1);
        cp.kind = cudaMemcpyHostToDevice;
        std::lock_guard<std::mutex> g(mu);
        cudaError_t e = cudaMemcpy3DAsync(&cp, c->copy_stream);
        if (e != cudaSuccess) cuda_err.store((int)e);
      });
    }
  }
  if (plain_atlas) {
    if (c->sc.mat_tex) { cudaDestroyTextureObject(c->sc.mat_tex); c->sc.mat_tex = 0; }
    if (c->mat_surf) { cudaDestroySurfaceObject(c->mat_surf); c->mat_surf = 0; }
    c->mat_surface = false;
    if (c->mat_arr) { cudaFreeArray(c->mat_arr); c->mat_arr = nullptr; c->mat_R = c->mat_L = 0; }
    if (!c->atlas_arr || c->atlas_R != R || c->atlas_L != L) {
      if (c->sc.atlas) cudaDestroyTextureObject(c->sc.atlas);
      c->sc.atlas = 0;
      if (c->atlas_arr) cudaFreeArray(c->atlas_arr);
      c->atlas_arr = nullptr;
      ACK(cudaMalloc3DArray(&c->atlas_arr, &fmt, make_cudaExtent(R, R, L), cudaArrayLayered));
      rd.res.array.array = c->atlas_arr;
      ACK(cudaCreateTextureObject(&c->sc.atlas, &rd, &td, nullptr));
      c->atlas_R = R; c->atlas_L = L;
    }
    if (c->stage_bytes < layer_bytes * L) {
      if (c->h_stage) cudaFreeHost(c->h_stage);
      c->h_stage = nullptr; c->stage_bytes = 0;
      ACK(cudaMallocHost(&c->h_stage, layer_bytes * L));
      c->stage_bytes = layer_bytes * L;
    }
    pool.run(L, J.workers, [&](int l) {
      uint8_t* dst = c->h_stage + (size_t)l * layer_bytes;
      memcpy(dst, J.atlas + (size_t)l * layer_bytes, layer_bytes);
      cudaMemcpy3DParms cp = {};
      cp.srcPtr = make_cudaPitchedPtr(dst, (size_t)R * 4, R, R);
      cp.dstArray = c->atlas_arr;
      cp.dstPos = make_cudaPos(0, 0, l);
      cp.extent = make_cudaExtent(R, R, 1);
      cp.kind = cudaMemcpyHostToDevice;
      std::lock_guard<std::mutex> g(mu);
      cudaError_t e = cudaMemcpy3DAsync(&cp, c->copy_stream);
      if (e != cudaSuccess) cuda_err.store((int)e);
    });
  }
  if (cuda_err.load()) return afail(FSPT_E_CUDA, "a staging copy", (cudaError_t)cuda_err.load());
  lap("stage + enqueue");
  // the two small tables k_shade reads next to the texels: same stream, so the same event covers them
  auto grow = [&](void*& p, size_t& cap, size_t bytes) -> cudaError_t {
    if (p && cap >= bytes) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  };
  ACK(grow(c->d_layer_info, c->cap_layer_info, (size_t)L * 8));
  ACK(grow(c->d_mat_info, c->cap_mat_info, J.n_mat_info * 4));
  ACK(cudaMemcpyAsync(c->d_layer_info, layer_info, (size_t)L * 8, cudaMemcpyHostToDevice, c->copy_stream));
  ACK(cudaMemcpyAsync(c->d_mat_info, mat_info, J.n_mat_info * 4, cudaMemcpyHostToDevice, c->copy_stream));
  c->sc.layer_info = reinterpret_cast<const uint2*>(c->d_layer_info);
  c->sc.mat_info = reinterpret_cast<const int4*>(c->d_mat_info);
  c->bytes_layer_info = (size_t)L * 8; c->bytes_mat_info = J.n_mat_info * 4;
  ACK(cudaEventRecord(c->ev_atlas, c->copy_stream));
  return FSPT_OK;
#undef ACK
}

// A device array of 16-byte words as a linear texture (tex1Dfetch): the second L1 data path next to the LSU.  Arrays
// beyond the 2^27-texel limit get no texture (0) and are read with plain loads.
int linear_tex(Ctx* c, cudaTextureObject_t* tex, void* ptr, size_t bytes, bool allowed = true) {
  if (*tex) cudaDestroyTextureObject(*tex);
  *tex = 0;
  if (!allowed || bytes / 16 > ((size_t)1 << 27) || bytes == 0) return FSPT_OK;
  cudaResourceDesc r = {};
  r.resType = cudaResourceTypeLinear;
  r.res.linear.devPtr = ptr;
  r.res.linear.desc = cudaCreateChannelDesc<float4>();
  r.res.linear.sizeInBytes = bytes;
  cudaTextureDesc t = {};
  t.readMode = cudaReadModeElementType;
  CK(cudaCreateTextureObject(tex, &r, &t, nullptr));
  return FSPT_OK;
}

// Traverses the continuation ray of every record of array `which` (d_counts[2*which] of them) + d_counts[2*which+1]
// shadow rays (d_shadow); results go into the records, plus one hit/miss byte per position (for k_shade).
int launch_trace(Ctx* c, int which, bool hit_flags, bool write_count, const FrameParams* cam = nullptr,
                 const float* rb_cam = nullptr, int n_samples = 1) {
  TraceArgs A;
  if (cam) A.f = *cam; else memset(&A.f, 0, sizeof A.f);
  A.rb_cam = rb_cam; A.n_samples = n_samples;
  A.div_s = make_fastdiv((unsigned)std::max(1, n_samples), c->wave_paths);
  A.anyhit = c->anyhit;
  A.nodes = c->sc.nodes; A.tris = c->sc.tris; A.root_ref = c->sc.root_ref; A.nodes_tex = c->nodes_tex;
  A.ps = c->ps2[which];
  A.shadow_rays = c->d_shadow;
  A.counts = c->d_counts + 2 * which;
  A.next = c->d_counts + 32;
  A.stats = c->d_stats;
  A.count_out = write_count ? c->d_count_out : nullptr;
  A.hit_flag = hit_flags ? c->d_hit_flag : nullptr;
  record_trace_begin(c, cam ? 2 : 0);  // tag 2: the camera-fused primary launch of a wave
  if (A.nodes_tex) {
    if (cam) k_trace<false, true, true><<<c->trace_blocks_cam, TRACE_THREADS, 0, c->stream>>>(A);
    else if (write_count) k_trace<true, false, true><<<c->trace_blocks_cnt, TRACE_THREADS, 0, c->stream>>>(A);
    else k_trace<false, false, true><<<c->trace_blocks, TRACE_THREADS, 0, c->stream>>>(A);
  } else {  // node array beyond the linear-texture limit
    if (cam) k_trace<false, true, false><<<c->trace_blocks_cam_nt, TRACE_THREADS, 0, c->stream>>>(A);
    else if (write_count) k_trace<true, false, false><<<c->trace_blocks_cnt_nt, TRACE_THREADS, 0, c->stream>>>(A);
    else k_trace<false, false, false><<<c->trace_blocks_nt, TRACE_THREADS, 0, c->stream>>>(A);
  }
  record_trace_end(c);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  return FSPT_OK;
}

int set_counts(Ctx* c, int n_cont, int n_shadow) {  // also zeroes #hit, #miss and the fetch cursor
  k_set_counts<<<1, 1, 0, c->stream>>>(c->d_counts, n_cont, n_shadow);
  CK(cudaGetLastError());
  return FSPT_OK;
}

FrameParams make_frame(const Ctx* c, const fspt_frame_params* f, bool whole_frame = false) {
  FrameParams p;
  for (int k = 0; k < 3; ++k) { p.eye[k] = f->eye[k]; p.dir[k] = f->dir[k]; }
  p.fov_scale = f->fov_scale; p.lens0 = f->lens_features[0]; p.lens1 = f->lens_features[1];
  p.env_theta = f->env_theta;
  {
    const v3 I = mk3(p.dir[0], p.dir[1], p.dir[2]);
    const v3 bx = normalize(cross(I, mk3(0.0f, 1.0f, 0.0f)));  // camera.fs:39
    const v3 by = normalize(cross(bx, I));                     // camera.fs:40
    p.basis_x[0] = bx.x; p.basis_x[1] = bx.y; p.basis_x[2] = bx.z;
    p.basis_y[0] = by.x; p.basis_y[1] = by.y; p.basis_y[2] = by.z;
  }
  p.width = c->width; p.height = c->height;
  p.rx0 = whole_frame ? 0 : c->rx0; p.ry0 = whole_frame ? 0 : c->ry0;
  p.rw = whole_frame ? c->width : c->rw; p.rh = whole_frame ? c->height : c->rh;
  p.tiled = (p.rw % 8 == 0 && p.rh % 4 == 0) ? 1 : 0;
  // slot -> pixel: one division by the tiles per row (dividend: a tile index) or by the row width (dividend: a pixel index)
  p.div_row = p.tiled ? make_fastdiv((unsigned)(p.rw >> 3), ((unsigned long long)p.rw * p.rh) >> 5)
                      : make_fastdiv((unsigned)p.rw, (unsigned long long)p.rw * p.rh);
  return p;
}

// Rand bases of one call -> device: [0,cap) camera, [cap,2cap) tracer.  The pinned staging is a ring; a slot is reused
// only after the DMA that read it has finished, so there is no stream synchronisation on the render path.
int stage_rand_bases(Ctx* c, const float* cam, const float* trace, int n) {
  if (n > c->rb_cap) {  // grow (rare): everything in flight must finish first
    CK(cudaStreamSynchronize(c->stream));
    dfree(c->d_rb);
    if (c->h_rb) cudaFreeHost(c->h_rb);
    c->h_rb = nullptr;
    c->rb_cap = std::max(n, 2 * c->rb_cap);
    CK(cudaMalloc(&c->d_rb, 2 * (size_t)c->rb_cap * sizeof(float)));
    CK(cudaMallocHost(&c->h_rb, Ctx::RB_SLOTS * 2 * (size_t)c->rb_cap * sizeof(float)));
  }
  if (n <= 0) return FSPT_OK;
  const int slot = c->rb_slot;
  c->rb_slot = (slot + 1) % Ctx::RB_SLOTS;
  CK(cudaEventSynchronize(c->ev_rb[slot]));  // returns at once unless RB_SLOTS calls are still queued
  float* h = c->h_rb + (size_t)slot * 2 * c->rb_cap;
  memcpy(h, cam, (size_t)n * sizeof(float));
  if (trace) memcpy(h + c->rb_cap, trace, (size_t)n * sizeof(float));
  CK(cudaMemcpyAsync(c->d_rb, h, 2 * (size_t)c->rb_cap * sizeof(float), cudaMemcpyHostToDevice, c->stream));
  CK(cudaEventRecord(c->ev_rb[slot], c->stream));
  return FSPT_OK;
}

// One wave.  Path records are dense arrays that ping-pong: set k = (array ps2[k], counts d_counts[2k .. 2k+1] =
// #records, #shadow rays); the shadow list holds record positions; fetch cursor at d_counts[32].
int render_wave(Ctx* c, const FrameParams& fp, uint32_t first_tick, int S, const float* rb_cam, const float* rb_trace) {
  const int P = fp.rw * fp.rh;
  const int n_paths = S * P;
  int rc = set_counts(c, n_paths, 0);
  if (rc) return rc;
  rc = launch_trace(c, 0, true, false, &fp, rb_cam, S);  // camera.fs + primary rays (tracer.fs:440), fused
  if (rc) return rc;
  ShadeArgs A;
  A.f = fp;
  A.rb_trace = rb_trace;
  A.hit_flag = c->d_hit_flag;
  A.shadow_rays_out = c->d_shadow;
  A.sample_color = c->d_sample_color;
  A.capped = c->d_stats + 3;
  A.n_samples = S;
  A.div_s = make_fastdiv((unsigned)S, c->wave_paths);
  A.max_refractions = c->max_refractions;
  A.anyhit = c->anyhit;
  const int hard_cap = FSPT_NUM_BOUNCES + 1 + (c->has_dielectric ? c->max_refractions + 2 : 0);
  // atlas of the last upload: an asynchronous upload's staging thread is joined HERE, with the primary traversal already
  // running on the GPU (the event it records has to exist before the stream can be made to wait on it); then the DMA
  // itself is waited for on the device (no-op once it has completed)
  // multi-GPU: the atlas travels to the other ranks now, behind the traversal.  (Phase 2 first: it joins the atlas thread
  // itself and, should the root's atlas part have failed, tells the other ranks instead of leaving them in the collective.)
  if ((rc = broadcast_phase2(c))) return rc;
  if ((rc = atlas_join(c))) return rc;
  A.sc = c->sc;  // (the atlas part sets the texture objects and table pointers)
  CK(cudaStreamWaitEvent(c->stream, c->ev_atlas, 0));
  int cur = 0;
  for (int b = 0; b < hard_cap; ++b) {
    const int nxt = cur ^ 1;
    CK(cudaMemsetAsync(c->d_counts + 2 * nxt, 0, 2 * sizeof(int), c->stream));
    A.first = (b == 0);
    A.ps = c->ps2[cur];
    A.ps_out = c->ps2[nxt];
    A.counts_in = c->d_counts + 2 * cur;
    A.counts_out = c->d_counts + 2 * nxt;
    record_trace_begin(c, 1);
    if (c->sc.mat_tex) k_shade<true><<<c->shade_blocks, SHADE_THREADS, 1024, c->stream>>>(A);
    else k_shade<false><<<c->shade_blocks, SHADE_THREADS, 1024, c->stream>>>(A);
    record_trace_end(c);
    c->stats.kernel_launches++;
    CK(cudaGetLastError());
    cur = nxt;
    if (b >= FSPT_NUM_BOUNCES) {
      if (!c->has_dielectric) break;  // every surviving path had i == NUM_BOUNCES: nothing was appended
      // Refractions do not count as bounces (`i--`, tracer.fs:488), so paths may survive.  The number of live paths is
      // copied out after every extra iteration, but the host reads the value of TWO iterations ago: the GPU always has
      // work queued and never waits for the host; at most two empty iterations (kernels that find zero items) are wasted.
      const int k = b % Ctx::POLL_SLOTS;
      CK(cudaMemcpyAsync(c->h_poll + k, c->d_counts + 2 * cur, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CK(cudaEventRecord(c->ev_poll[k], c->stream));
      if (b >= FSPT_NUM_BOUNCES + 2) {
        const int k2 = (b - 2) % Ctx::POLL_SLOTS;
        CK(cudaEventSynchronize(c->ev_poll[k2]));
        if (c->h_poll[k2] == 0) break;
      }
    }
    CK(cudaMemsetAsync(c->d_counts + 32, 0, sizeof(int), c->stream));  // fetch cursor
    rc = launch_trace(c, cur, true, false);  // tracer.fs:501,507
    if (rc) return rc;
  }
  k_accumulate<<<(P + 255) / 256, 256, 0, c->stream>>>(c->d_sample_color, c->d_fb, c->d_last_color, fp.rx0, fp.ry0, fp.rw,
                                                       fp.rh, fp.width, S, first_tick, c->accum_mode, c->sanitize);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  return FSPT_OK;
}

// ---- NCCL, resolved at run time ----------------------------------------------------------------------------------
// The library has no link-time dependency on NCCL: a single-GPU host never loads it.  dlopen("libnccl.so.2") returns
// the copy already mapped into the process when there is one (e.g. the one a PyTorch host brought along), else the
// system library.
struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Reduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};
NcclApi* nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    const char* names[] = {getenv("FSPT_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) { api.error = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return; }
    auto sym = [&](const char* name) {
      void* p = dlsym(api.handle, name);
      if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + name;
      return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Reduce = reinterpret_cast<decltype(api.Reduce)>(sym("ncclReduce"));
    api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return &api;
}
#define NK(call)                                                                                               \
  do {                                                                                                         \
    ncclResult_t r_ = (call);                                                                                  \
    if (r_ != ncclSuccess) return fail(c, FSPT_E_NCCL, "%s failed: %s", #call, nccl_api()->GetErrorString(r_)); \
  } while (0)

void comm_free(Ctx* c) {
  if (c->comm) nccl_api()->CommDestroy(c->comm);
  c->comm = nullptr;
  c->comm_rank = 0; c->comm_world = 1;
}

// what a receiving rank has to know before the buffers of fspt_scene_broadcast arrive
struct SceneHeader {
  uint64_t bytes_nodes, bytes_tris, bytes_shade, bytes_bins, bytes_layer_info, bytes_mat_info, scene_bytes;
  int32_t root_ref, n_tris, n_interior, atlas_res, atlas_layers, env_w, env_h, n_bins;
  int32_t mat_R, mat_L, use_mat_tex, has_dielectric, magic;
};

// what a receiving rank has to know before the atlas part of fspt_scene_broadcast arrives (phase 2)
struct SceneHeader2 {
  uint64_t bytes_layer_info, bytes_mat_info;
  int32_t mat_R, mat_L, use_mat_tex, magic;
};

// The scene resident on `root` -> every other rank, device to device over NVLink: the records fspt_scene_upload built
// (Node64, Tri48, ShadeRec, bins, layer / material tables) and linear copies of the environment and atlas arrays.  One
// host stages and uploads a scene once instead of every rank repeating the same 200 MB of host work.
//
// Two phases, like the upload itself.  Phase 1 (here, on the context's stream): what the traversal kernel needs -- and
// the environment.  Phase 2 (the atlas and its tables, on the copy stream) is DEFERRED to the point where the root's
// atlas is needed anyway: the first shading launch of the next fspt_render, behind its primary traversal launch (or the
// next fspt_synchronize / fspt_reduce_accum / upload / broadcast, whichever comes first -- every rank reaches one of them,
// in the same order relative to the other collectives).  An asynchronous upload on the root therefore overlaps its atlas
// transfer with the primary traversal on EVERY rank, and the root's host is not held in the broadcast call.
int broadcast_phase2(Ctx* c) {
  if (c->bcast_pending_root < 0) return FSPT_OK;
  const int root = c->bcast_pending_root;
  c->bcast_pending_root = -1;
  const bool is_root = c->comm_rank == root;
  NcclApi* N = nccl_api();
  SceneHeader2 h;
  memset(&h, 0, sizeof h);
  static_assert(sizeof(SceneHeader2) <= 128, "header buffer");
  uint8_t* d_hdr2 = reinterpret_cast<uint8_t*>(c->d_hdr) + 256;
  // the collectives of one communicator run one after the other: this one starts behind phase 1 (its last collective
  // recorded the event), not behind the traversal launch that may have been enqueued since
  CK(cudaStreamWaitEvent(c->copy_stream, c->ev_bcast, 0));
  if (is_root) {
    int rcj = atlas_join(c);  // the root's atlas part ends here at the latest; its outcome travels in the header
    if (rcj) h.magic = -1;    // the ranks must not wait for an atlas that will not come
    else {
      h.bytes_layer_info = c->bytes_layer_info; h.bytes_mat_info = c->bytes_mat_info;
      h.use_mat_tex = c->sc.mat_tex ? 1 : 0;
      h.mat_R = h.use_mat_tex ? c->mat_R : c->atlas_R;
      h.mat_L = h.use_mat_tex ? c->mat_L : c->atlas_L;
      h.magic = 0x46535032;
    }
    CK(cudaMemcpyAsync(d_hdr2, &h, sizeof h, cudaMemcpyHostToDevice, c->copy_stream));
    NK(N->Broadcast(d_hdr2, d_hdr2, sizeof h, ncclUint8, root, c->comm, c->copy_stream));
    if (rcj) return rcj;
  } else {
    NK(N->Broadcast(d_hdr2, d_hdr2, sizeof h, ncclUint8, root, c->comm, c->copy_stream));
    CK(cudaMemcpyAsync(&h, d_hdr2, sizeof h, cudaMemcpyDeviceToHost, c->copy_stream));
    CK(cudaStreamSynchronize(c->copy_stream));
    if (h.magic != 0x46535032) { c->has_scene = false; return fail(c, FSPT_E_NCCL, "fspt_scene_broadcast: rank %d has no atlas to send (its upload failed)", root); }
  }
  const size_t texel = h.use_mat_tex ? 16 : 4;
  const size_t atlas_bytes = (size_t)h.mat_R * h.mat_R * texel * (size_t)h.mat_L;
  int rc;
  if ((rc = ensure(c, c->d_lin, c->cap_lin, atlas_bytes))) return rc;
  uint8_t* lin = reinterpret_cast<uint8_t*>(c->d_lin);
  cudaMemcpy3DParms cp = {};
  cp.extent = make_cudaExtent(h.mat_R, h.mat_R, h.mat_L);
  cp.kind = cudaMemcpyDeviceToDevice;
  if (is_root) {
    // the atlas DMA of the upload runs on this same stream (in order)
    cp.srcArray = h.use_mat_tex ? c->mat_arr : c->atlas_arr;
    cp.dstPtr = make_cudaPitchedPtr(lin, (size_t)h.mat_R * texel, h.mat_R, h.mat_R);
    CK(cudaMemcpy3DAsync(&cp, c->copy_stream));
  } else {
    // storage on the receiving side: same reuse rules as fspt_scene_upload
    if ((rc = ensure(c, c->d_layer_info, c->cap_layer_info, h.bytes_layer_info))) return rc;
    if ((rc = ensure(c, c->d_mat_info, c->cap_mat_info, h.bytes_mat_info))) return rc;
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    if (h.use_mat_tex) {
      if (c->sc.atlas) { cudaDestroyTextureObject(c->sc.atlas); c->sc.atlas = 0; }
      if (c->atlas_arr) { cudaFreeArray(c->atlas_arr); c->atlas_arr = nullptr; c->atlas_R = c->atlas_L = 0; }
      if (!c->mat_arr || c->mat_R != h.mat_R || c->mat_L != h.mat_L) {
        if (c->sc.mat_tex) cudaDestroyTextureObject(c->sc.mat_tex);
        c->sc.mat_tex = 0;
        if (c->mat_surf) cudaDestroySurfaceObject(c->mat_surf);
        c->mat_surf = 0; c->mat_surface = false;
        if (c->mat_arr) cudaFreeArray(c->mat_arr);
        c->mat_arr = nullptr;
        cudaChannelFormatDesc fmt4 = cudaCreateChannelDesc<uint4>();
        CK(cudaMalloc3DArray(&c->mat_arr, &fmt4, make_cudaExtent(h.mat_R, h.mat_R, h.mat_L), cudaArrayLayered));
        rd.res.array.array = c->mat_arr;
        CK(cudaCreateTextureObject(&c->sc.mat_tex, &rd, &td, nullptr));
        c->mat_R = h.mat_R; c->mat_L = h.mat_L;
      }
    } else {
      if (c->sc.mat_tex) { cudaDestroyTextureObject(c->sc.mat_tex); c->sc.mat_tex = 0; }
      if (c->mat_surf) { cudaDestroySurfaceObject(c->mat_surf); c->mat_surf = 0; }
      c->mat_surface = false;
      if (c->mat_arr) { cudaFreeArray(c->mat_arr); c->mat_arr = nullptr; c->mat_R = c->mat_L = 0; }
      if (!c->atlas_arr || c->atlas_R != h.mat_R || c->atlas_L != h.mat_L) {
        if (c->sc.atlas) cudaDestroyTextureObject(c->sc.atlas);
        c->sc.atlas = 0;
        if (c->atlas_arr) cudaFreeArray(c->atlas_arr);
        c->atlas_arr = nullptr;
        CK(cudaMalloc3DArray(&c->atlas_arr, &fmt, make_cudaExtent(h.mat_R, h.mat_R, h.mat_L), cudaArrayLayered));
        rd.res.array.array = c->atlas_arr;
        CK(cudaCreateTextureObject(&c->sc.atlas, &rd, &td, nullptr));
        c->atlas_R = h.mat_R; c->atlas_L = h.mat_L;
      }
    }
  }
  NK(N->GroupStart());
  NK(N->Broadcast(c->d_layer_info, c->d_layer_info, h.bytes_layer_info, ncclUint8, root, c->comm, c->copy_stream));
  NK(N->Broadcast(c->d_mat_info, c->d_mat_info, h.bytes_mat_info, ncclUint8, root, c->comm, c->copy_stream));
  NK(N->Broadcast(lin, lin, atlas_bytes, ncclUint8, root, c->comm, c->copy_stream));
  NK(N->GroupEnd());
  c->stats.kernel_launches += 4;
  if (!is_root) {
    cp.srcPtr = make_cudaPitchedPtr(lin, (size_t)h.mat_R * texel, h.mat_R, h.mat_R);
    cp.dstArray = h.use_mat_tex ? c->mat_arr : c->atlas_arr;
    CK(cudaMemcpy3DAsync(&cp, c->copy_stream));
    c->bytes_layer_info = h.bytes_layer_info; c->bytes_mat_info = h.bytes_mat_info;
    c->sc.layer_info = reinterpret_cast<const uint2*>(c->d_layer_info);
    c->sc.mat_info = reinterpret_cast<const int4*>(c->d_mat_info);
  }
  // what the first shading launch waits for; later collectives on the context's stream come behind it as well
  CK(cudaEventRecord(c->ev_atlas, c->copy_stream));
  CK(cudaStreamWaitEvent(c->stream, c->ev_atlas, 0));
  return FSPT_OK;
}

int pull_stats(Ctx* c) {
  unsigned long long h[8];
  CK(cudaMemcpy(h, c->d_stats, sizeof h, cudaMemcpyDeviceToHost));
  c->stats.rays = h[0]; c->stats.node_visits = h[1]; c->stats.leaf_visits = h[2]; c->stats.capped_paths = h[3];
  c->stats.last_rays = h[0] - h[4]; c->stats.last_node_visits = h[1] - h[5]; c->stats.last_leaf_visits = h[2] - h[6];
  return FSPT_OK;
}

}  // namespace

extern "C" {

int fspt_abi_version(void) { return FSPT_ABI_VERSION; }

const char* fspt_last_error(const fspt_ctx* ctx) {
  if (!ctx) return g_create_error.c_str();
  return reinterpret_cast<const Ctx*>(ctx)->error.c_str();
}

int fspt_create(fspt_ctx** out, int32_t width, int32_t height, int32_t device) {
  Ctx* c = nullptr;
  if (!out || width <= 0 || height <= 0 || (int64_t)width * height > (1 << 27))
    return fail(c, FSPT_E_INVALID, "fspt_create: bad arguments (%d x %d)", width, height);
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev <= 0)
    return fail(c, FSPT_E_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
  if (device < 0 || device >= n_dev) return fail(c, FSPT_E_INVALID, "device %d out of range (%d present)", device, n_dev);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return fail(c, FSPT_E_CUDA, "device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
  Ctx* ctx = new Ctx();
  c = ctx;
  c->device = device; c->sm_count = prop.multiProcessorCount;
  c->width = width; c->height = height; c->n_pixels = width * height;
  c->rx0 = c->ry0 = 0; c->rw = width; c->rh = height;
  cudaError_t err;
#define CKC(call) if ((err = (call)) != cudaSuccess) { g_create_error = std::string(#call) + ": " + cudaGetErrorString(err); delete ctx; return FSPT_E_CUDA; }
  CKC(cudaSetDevice(device));
  CKC(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  CKC(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  CKC(cudaEventCreateWithFlags(&c->ev_atlas, cudaEventDisableTiming));
  CKC(cudaEventCreate(&c->ev_begin));
  CKC(cudaEventCreate(&c->ev_end));
  CKC(cudaMalloc(&c->d_fb, (size_t)c->n_pixels * 16));
  CKC(cudaMemset(c->d_fb, 0, (size_t)c->n_pixels * 16));
  CKC(cudaMalloc(&c->d_last_color, (size_t)c->n_pixels * 16));
  CKC(cudaMalloc(&c->d_cam_pos, (size_t)c->n_pixels * 16));
  CKC(cudaMalloc(&c->d_cam_dir, (size_t)c->n_pixels * 16));
  CKC(cudaMalloc(&c->d_rgba8, (size_t)c->n_pixels * 4));
#undef CKC
  int rc = alloc_wave(c);
  if (rc) { g_create_error = c->error; fspt_destroy(reinterpret_cast<fspt_ctx*>(c)); return rc; }
  // persistent grids: resident CTAs per SM x SM count.  The traversal kernel keeps its short stacks and a few per-ray
  // values in shared memory; everything else of the unified array should stay L1 (BVH nodes and leaf blocks live
  // there), so the carveout is set to what the resident CTAs need and no more.
  int max_smem_sm = 0;
  cudaDeviceGetAttribute(&max_smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
  auto persistent = [&](auto kernel) {
    int per_sm = 0;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TRACE_THREADS, 0);
    per_sm = std::max(1, per_sm);
    cudaFuncAttributes fa{};
    cudaFuncGetAttributes(&fa, kernel);
    if (!getenv("FSPT_NO_CARVEOUT") && max_smem_sm > 0) {
      const size_t need = (size_t)per_sm * (fa.sharedSizeBytes + 1024);  // + the per-CTA reservation
      int pct = (int)std::min<size_t>(100, (need * 100 + (size_t)max_smem_sm - 1) / (size_t)max_smem_sm);
      if (const char* e = getenv("FSPT_CARVEOUT_PCT")) pct = std::max(0, std::min(100, atoi(e)));  // measurement knob
      cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
    }
    return per_sm * c->sm_count;
  };
  c->trace_blocks = persistent(k_trace<false, false, true>);
  c->trace_blocks_cnt = persistent(k_trace<true, false, true>);
  c->trace_blocks_cam = persistent(k_trace<false, true, true>);
  c->trace_blocks_nt = persistent(k_trace<false, false, false>);
  c->trace_blocks_cnt_nt = persistent(k_trace<true, false, false>);
  c->trace_blocks_cam_nt = persistent(k_trace<false, true, false>);
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_shade<true>, SHADE_THREADS, 1024);
  c->shade_blocks = std::max(1, per_sm) * c->sm_count;
  *out = reinterpret_cast<fspt_ctx*>(c);
  return FSPT_OK;
}

void fspt_destroy(fspt_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->atlas_thread.joinable()) c->atlas_thread.join();
  c->bcast_pending_root = -1;  // (no collective in a destructor)
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  delete c->pool;
  c->pool = nullptr;
  if (c->h_ring) cudaFreeHost(c->h_ring);
  for (auto e : c->ev_ring) cudaEventDestroy(e);
  free_scene(c);
  dfree(c->d_fb); dfree(c->d_last_color); dfree(c->d_sample_color); dfree(c->d_cam_pos); dfree(c->d_cam_dir); dfree(c->d_rgba8);
  dfree(c->ps2[0].rec); dfree(c->ps2[1].rec); dfree(c->ps2[0].sh); dfree(c->ps2[1].sh);
  dfree(c->d_shadow);
  dfree(c->d_counts); dfree(c->d_count_out); dfree(c->d_hit_flag); dfree(c->d_stats); dfree(c->d_rb);
  if (c->h_rb) cudaFreeHost(c->h_rb);
  if (c->h_poll) cudaFreeHost(c->h_poll);
  for (auto e : c->ev_rb) if (e) cudaEventDestroy(e);
  for (auto e : c->ev_poll) if (e) cudaEventDestroy(e);
  comm_free(c);
  if (c->ev_red0) cudaEventDestroy(c->ev_red0);
  if (c->ev_red1) cudaEventDestroy(c->ev_red1);
  dfree(c->d_lin); dfree(c->d_hdr); dfree(c->d_lin_env);
  if (c->ev_bcast) cudaEventDestroy(c->ev_bcast);
  if (c->ev_env) cudaEventDestroy(c->ev_env);
  for (auto e : c->ev_trace) cudaEventDestroy(e);
  if (c->ev_begin) cudaEventDestroy(c->ev_begin);
  if (c->ev_end) cudaEventDestroy(c->ev_end);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->ev_atlas) cudaEventDestroy(c->ev_atlas);
  if (c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

// fspt_scene_upload / fspt_scene_upload_async.
// Order of work (round 2, second half: the geometry used to be staged next to the atlas by a quarter of the host
// threads and DMA'd only when all of it was ready, which put ~3 ms between the call and the first traversal launch):
//   1. pre-passes over nodes and triangles (parallel): interior-record numbering, material ids;
//   2. geometry records (Node64 / Tri48 / ShadeRec), bins and environment: built chunk by chunk by ALL host workers into
//      a ring of pinned slots, every chunk DMA'd the moment it is complete -- the traversal kernel's inputs are on
//      their way before anything else is touched, and a 10 M-triangle scene (2.6 GB of records) streams through
//      256 MB of pinned memory instead of a pageable copy;
//   3. the atlas part (stage_atlas): inline, or -- async_atlas -- on the context's atlas thread after this function has
//      returned; every other buffer of the scene has been consumed by then.
static int scene_upload_impl(fspt_ctx* ctx, const fspt_scene_desc* s, bool async_atlas) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  const bool timing = getenv("FSPT_TIMING") != nullptr;
  auto t_prev = std::chrono::steady_clock::now();
  auto lap = [&](const char* what) {
    if (!timing) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[fspt upload] %-28s %7.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
    t_prev = now;
  };
  if (!c) return FSPT_E_INVALID;
  if (!s || !s->bvh || !s->triangles || !s->materials || !s->normals || !s->uvs || !s->atlas || !s->env || !s->radiance_bins)
    return fail(c, FSPT_E_INVALID, "scene_upload: NULL buffer");
  if (s->n_nodes <= 0 || s->n_triangles <= 0 || s->atlas_res <= 0 || s->atlas_layers <= 0 || s->env_width <= 0 ||
      s->env_height <= 0 || s->env_bins <= 0)
    return fail(c, FSPT_E_INVALID, "scene_upload: non-positive size (an environment with >= 1 bin is mandatory, main.js:303-308)");
  if (s->leaf_size != 4) return fail(c, FSPT_E_INVALID, "scene_upload: LEAF_SIZE must be 4 (main.js:45), got %d", s->leaf_size);
  CK(cudaSetDevice(c->device));
  if (int rcb = broadcast_phase2(c)) return rcb;  // (a broadcast nobody rendered from: finish it before its source goes away)
  (void)atlas_join(c);  // the atlas thread of the previous upload still reads the pinned blocks and the arrays
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaStreamSynchronize(c->copy_stream));
  c->has_scene = false;  // device buffers, arrays and texture objects of the previous scene are reused when they fit
  lap("sync");
  const int N = s->n_nodes, T = s->n_triangles;
  const int L = s->atlas_layers;
  // host threads for staging: all cores of a single-process host; one process per GPU shares them
  // (dist.share_host_threads sets FSPT_UPLOAD_THREADS = cores / processes on this node)
  int hw = (int)std::max(4u, std::min(32u, std::thread::hardware_concurrency()));
  if (const char* e = getenv("FSPT_UPLOAD_THREADS")) hw = std::max(4, std::min(64, atoi(e)));
  if (!c->pool) { const int device = c->device; c->pool = new HostPool([device]() { cudaSetDevice(device); }); }
  HostPool& pool = *c->pool;
  // ---- pre-passes (scene_pack.h): interior-record numbering, material ids, depth / tree check
  ScenePrepass pre;
  {
    const int rcp = scene_prepass(pool, hw, s, pre);
    if (rcp) return fail(c, rcp, "%s", pre.error.c_str());
  }
  const std::vector<int32_t>& ref = pre.ref;
  std::vector<std::array<int, 4>>& mats = pre.mats;
  const bool dielectric = pre.dielectric;
  const size_t NI = pre.NI();
  lap("node + material pre-pass");

  // ---- pinned block for the small tables and the environment (kept between uploads)
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  const size_t env_bytes = (size_t)s->env_width * s->env_height * 4;
  const size_t o_bins = 0, o_layer = o_bins + al((size_t)s->env_bins * 16), o_mat = o_layer + al((size_t)L * 8),
               o_matsrc = o_mat + al(std::max<size_t>(32, mats.size() * 32)),
               o_env = o_matsrc + al(std::max<size_t>(32, mats.size() * sizeof(MatSrc))), small_bytes = o_env + al(env_bytes);
  if (c->geo_stage_bytes < small_bytes) {
    if (c->h_geo) cudaFreeHost(c->h_geo);
    c->h_geo = nullptr; c->geo_stage_bytes = 0;
    CK(cudaMallocHost(&c->h_geo, small_bytes));
    c->geo_stage_bytes = small_bytes;
  }
  uint8_t* hg = c->h_geo;
  float* bins = reinterpret_cast<float*>(hg + o_bins);
  const size_t n_mat_info = std::max<size_t>(8, mats.size() * 8);
  // ---- device buffers, environment array
  int rc_;
  const size_t nodes_bytes = std::max<size_t>(NI, 1) * 64, tris_bytes = (size_t)(T + 3) * 48, shade_bytes = (size_t)T * 192,
               bins_bytes = (size_t)s->env_bins * 16;
  if ((rc_ = ensure(c, c->d_nodes, c->cap_nodes, nodes_bytes))) return rc_;
  if ((rc_ = ensure(c, c->d_tris, c->cap_tris, tris_bytes))) return rc_;
  if ((rc_ = ensure(c, c->d_shade, c->cap_shade, shade_bytes))) return rc_;
  if ((rc_ = ensure(c, c->d_bins, c->cap_bins, bins_bytes))) return rc_;
  if (!c->env_arr || c->env_W != s->env_width || c->env_H != s->env_height) {  // 2D array, RGBA8 RGBE (main.js:170-180)
    cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeArray;
    cudaTextureDesc td = {};
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    if (c->sc.env) cudaDestroyTextureObject(c->sc.env);
    c->sc.env = 0;
    if (c->env_arr) cudaFreeArray(c->env_arr);
    c->env_arr = nullptr;
    CK(cudaMallocArray(&c->env_arr, &fmt, s->env_width, s->env_height));
    rd.res.array.array = c->env_arr;
    CK(cudaCreateTextureObject(&c->sc.env, &rd, &td, nullptr));
    c->env_W = s->env_width; c->env_H = s->env_height;
  }
  // ---- work items of the staging region: bands of the environment (+ the bins), node chunks, triangle chunks -- each
  // DMA'd the moment it is complete -- and, behind them, the constant-layer scan of the atlas, which the workers that
  // run out of geometry pick up.  Chunks are sized so that a small scene still gives every worker a few items; each
  // geometry chunk owns one slot of the pinned ring.
  const int workers = hw;
  const int tri_chunk = (int)std::max<size_t>(1024, std::min<size_t>(16384, ((size_t)T + 3) / (4 * (size_t)workers) + 1));
  const int node_chunk = (int)std::max<size_t>(4096, std::min<size_t>(65536, NI / (4 * (size_t)workers) + 1));
  const int n_node_items = NI ? (int)((NI + node_chunk - 1) / node_chunk) : 1;
  const int n_tri_items = (T + 3 + tri_chunk - 1) / tri_chunk;
  const int n_env_items = std::max(1, std::min(s->env_height, (int)(env_bytes >> 20)));  // ~1 MB of rows each
  const int n_geo_items = n_node_items + n_tri_items;
  // (a page-locked atlas under an asynchronous upload: the scan is all the host work the atlas part has left, and it
  // should not hold up this function's return -- it moves to the atlas thread)
  const bool atlas_pinned = host_is_pinned(s->atlas);
  const bool scan_here = !(async_atlas && atlas_pinned);
  const int n_scan_items = scan_here ? L * SCAN_BANDS : 0;
  const int n_items = n_env_items + n_geo_items + n_scan_items;
  const size_t layer_texels = (size_t)s->atlas_res * s->atlas_res;
  std::vector<std::atomic<int>> varied((size_t)L);
  for (auto& v : varied) v.store(0);
  const size_t slot_bytes = al(std::max<size_t>((size_t)node_chunk * 64, (size_t)tri_chunk * (48 + 192)));
  int n_slots = (int)std::max<size_t>(1, std::min<size_t>((size_t)n_geo_items, std::max<size_t>(2 * (size_t)workers, ((size_t)256 << 20) / slot_bytes)));
  if (const char* e = getenv("FSPT_RING_SLOTS")) n_slots = std::max(1, std::min(n_slots, atoi(e)));  // test knob: force slot reuse
  const bool ring_wraps = n_geo_items > n_slots;  // otherwise every chunk has a slot of its own and no event is needed
  if (c->ring_bytes < slot_bytes * (size_t)n_slots) {
    if (c->h_ring) cudaFreeHost(c->h_ring);
    c->h_ring = nullptr; c->ring_bytes = 0;
    CK(cudaMallocHost(&c->h_ring, slot_bytes * (size_t)n_slots));
    c->ring_bytes = slot_bytes * (size_t)n_slots;
  }
  while ((int)c->ev_ring.size() < n_slots) {
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    c->ev_ring.push_back(e);
  }
  lap("buffers + pinned ring");
  struct GeoStatus { int code = FSPT_OK; char msg[192] = {0}; } geo;
  std::mutex geo_mu;
  auto geo_fail = [&](int code, const char* fmt, int a0, int a1, int a2) {
    std::lock_guard<std::mutex> g(geo_mu);
    if (geo.code == FSPT_OK) { geo.code = code; snprintf(geo.msg, sizeof geo.msg, fmt, a0, a1, a2); }
  };
  std::atomic<int> dma_err(0);
  // (No lock around the enqueues: the runtime is thread-safe, and a slot's event only has to follow the slot's own copies,
  // which the recording thread enqueued itself.  A mutex held across the two or three calls of an item serialised ~60
  // items x ~12 us on the bench scene -- most of the region's 1.4 ms.)
  const bool src_env_pinned = host_is_pinned(s->env);
  std::vector<std::atomic<int>> slot_done((size_t)n_slots);  // last item whose copies have been enqueued from the slot
  for (auto& v : slot_done) v.store(-1);
  pool.run(n_items, workers, [&](int item) {
    if (item < n_env_items) {  // a band of environment rows (+ the bins): their own pinned block
      if (item == 0)
        for (size_t i = 0; i < (size_t)s->env_bins * 4; ++i) bins[i] = (float)s->radiance_bins[i];  // vec4(radianceBins[idx]), tracer.fs:424
      const size_t pitch = (size_t)s->env_width * 4;
      const int y0 = (int)((long long)s->env_height * item / n_env_items), y1 = (int)((long long)s->env_height * (item + 1) / n_env_items);
      const uint8_t* src = s->env + y0 * pitch;   // page-locked by the caller (fspt_host_register): DMA'd where it lies
      if (!src_env_pinned) { memcpy(hg + o_env + y0 * pitch, src, (size_t)(y1 - y0) * pitch); src = hg + o_env + y0 * pitch; }
      cudaError_t e = item == 0 ? cudaMemcpyAsync(c->d_bins, bins, bins_bytes, cudaMemcpyHostToDevice, c->stream) : cudaSuccess;
      if (e == cudaSuccess)
        e = cudaMemcpy2DToArrayAsync(c->env_arr, 0, y0, src, pitch, pitch, y1 - y0, cudaMemcpyHostToDevice, c->stream);
      if (e != cudaSuccess) dma_err.store((int)e);
      return;
    }
    if (item >= n_env_items + n_geo_items) {  // atlas: constant-layer scan
      scan_layer_band(s->atlas, layer_texels, item - n_env_items - n_geo_items, varied.data());
      return;
    }
    // ring slot of this chunk: free once the copies of the chunk that used it last have completed
    const int ring_item = item - n_env_items, slot = ring_item % n_slots;
    if (ring_item >= n_slots) {
      while (slot_done[slot].load(std::memory_order_acquire) != ring_item - n_slots) std::this_thread::yield();
      cudaEventSynchronize(c->ev_ring[slot]);
    }
    uint8_t* hs = c->h_ring + (size_t)slot * slot_bytes;
    cudaError_t e = cudaSuccess;
    if (ring_item < n_node_items) {
      const size_t k0 = (size_t)ring_item * node_chunk, k1 = std::min(NI, k0 + node_chunk);
      float* nodes = reinterpret_cast<float*>(hs);
      int bl = 0, br = 0;
      const int bad = pack_node_chunk(s, pre, k0, k1, nodes, &bl, &br);
      if (bad >= 0) geo_fail(FSPT_E_INVALID, "node %d: child index out of range (%d, %d)", bad, bl, br);
      const size_t bytes = NI ? (k1 - k0) * 64 : 64;
      e = cudaMemcpyAsync(reinterpret_cast<uint8_t*>(c->d_nodes) + k0 * 64, nodes, bytes, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess && ring_wraps) e = cudaEventRecord(c->ev_ring[slot], c->stream);
    } else {
      const int t0 = (ring_item - n_node_items) * tri_chunk, t1 = std::min(T + 3, t0 + tri_chunk), ts = std::min(T, t1);
      float* tris = reinterpret_cast<float*>(hs);
      float* shade = reinterpret_cast<float*>(hs + (size_t)tri_chunk * 48);
      pack_tri_chunk(s, pre, t0, t1, tris, shade);
      e = cudaMemcpyAsync(reinterpret_cast<uint8_t*>(c->d_tris) + (size_t)t0 * 48, tris, (size_t)(t1 - t0) * 48, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess && ts > t0)
        e = cudaMemcpyAsync(reinterpret_cast<uint8_t*>(c->d_shade) + (size_t)t0 * 192, shade, (size_t)(ts - t0) * 192, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess && ring_wraps) e = cudaEventRecord(c->ev_ring[slot], c->stream);
    }
    if (e != cudaSuccess) dma_err.store((int)e);
    slot_done[slot].store(ring_item, std::memory_order_release);
  });
  if (dma_err.load()) return fail(c, FSPT_E_CUDA, "geometry upload failed: %s", cudaGetErrorString((cudaError_t)dma_err.load()));
  c->env_in_place = src_env_pinned;
  if (src_env_pinned) {
    if (!c->ev_env) CK(cudaEventCreateWithFlags(&c->ev_env, cudaEventDisableTiming));
    CK(cudaEventRecord(c->ev_env, c->stream));
  }
  lap("geometry + env + layer scan");
  // env: test knob for the LSU-only instantiation of k_trace
  if ((rc_ = linear_tex(c, &c->nodes_tex, c->d_nodes, nodes_bytes, !getenv("FSPT_NO_NODE_TEX")))) return rc_;
  c->sc.nodes = reinterpret_cast<const float4*>(c->d_nodes);
  c->sc.tris = reinterpret_cast<const float4*>(c->d_tris);
  c->sc.shade = reinterpret_cast<const float4*>(c->d_shade);
  c->sc.bins = reinterpret_cast<const float4*>(c->d_bins);
  c->bytes_nodes = nodes_bytes; c->bytes_tris = tris_bytes; c->bytes_shade = shade_bytes; c->bytes_bins = bins_bytes;
  c->sc.root_ref = ref[0];
  c->sc.n_tris = T; c->sc.n_interior = (int)NI;
  c->sc.atlas_res = s->atlas_res; c->sc.atlas_layers = s->atlas_layers; c->sc.env_w = s->env_width; c->sc.env_h = s->env_height;
  c->sc.n_bins = s->env_bins;
  set_env_constants(c->sc);
  c->has_dielectric = dielectric;
  c->scene_bytes = (size_t)N * 36 + (size_t)T * (36 + 48 + 108 + 24) + (size_t)s->atlas_res * s->atlas_res * 4 * s->atlas_layers +
                   env_bytes + (size_t)s->env_bins * 8;
  // ---- the atlas part: here, or on the context's atlas thread (the caller's atlas stays borrowed until it is joined).
  // Measured alternatives (e2e step of the bench scene, asynchronous upload, 29.4 ms as built): the layer scan inside the
  // atlas part instead of the staging region 29.8 (the upload returns 0.2 ms earlier but the atlas lands 0.7 ms later, and
  // the first shading launch waits for that); the whole atlas part started before the geometry staging on workers of
  // its own 32.5 (32 threads on 16 cores: the geometry, which the first traversal launch waits for, takes 3 ms).
  AtlasJob job;
  job.atlas = s->atlas; job.L = L; job.R = s->atlas_res;
  job.mats = std::move(mats);
  job.layer_info = reinterpret_cast<uint32_t*>(hg + o_layer);
  job.mat_info = reinterpret_cast<int32_t*>(hg + o_mat);
  job.n_mat_info = n_mat_info;
  job.mat_src = reinterpret_cast<MatSrc*>(hg + o_matsrc);
  if (scan_here) {
    job.varied.resize((size_t)L);
    for (int l = 0; l < L; ++l) job.varied[l] = (uint8_t)varied[l].load();
  }
  job.src_pinned = atlas_pinned;
  job.workers = hw;
  job.timing = timing;
  c->atlas_rc = FSPT_OK;
  c->atlas_in_place = false;
  struct StagingGuard {  // an error return below must not leave the atlas thread reading the caller's buffer
    Ctx* c; bool ok = false;
    ~StagingGuard() { if (!ok && c->atlas_thread.joinable()) c->atlas_thread.join(); }
  } staging_guard{c};
  if (async_atlas) {
    const int device = c->device;
    c->atlas_thread = std::thread([c, device, job = std::move(job)]() {
      cudaSetDevice(device);
      c->atlas_rc = stage_atlas(c, job);
    });
    lap("atlas part handed to its thread");
  } else {
    if ((rc_ = stage_atlas(c, job))) { c->error = c->atlas_error; return rc_; }
    lap("atlas part");
  }
  if (geo.code != FSPT_OK) return fail(c, geo.code, "%s", geo.msg);
  // no synchronisation: everything the DMA engines still read lives in the context's pinned blocks, which the next
  // upload (and destroy) only touch after synchronising the streams; work enqueued by fspt_render waits in order
  if (timing && !async_atlas) { CK(cudaStreamSynchronize(c->stream)); CK(cudaStreamSynchronize(c->copy_stream)); lap("sync (timing only)"); }
  c->has_scene = true;
  staging_guard.ok = true;
  if (!async_atlas) {  // a synchronous upload has consumed every buffer when it returns, page-locked ones included
    int rcw = release_borrowed(c);
    if (rcw) return rcw;
  }
  return FSPT_OK;
}

int fspt_scene_upload(fspt_ctx* ctx, const fspt_scene_desc* s) { return scene_upload_impl(ctx, s, false); }
int fspt_scene_upload_async(fspt_ctx* ctx, const fspt_scene_desc* s) { return scene_upload_impl(ctx, s, true); }
int fspt_debug_pack_scene(const fspt_scene_desc* s, float* node64_out, float* tri48_out, float* shaderec_out,
                          int32_t* mat_id_out, int32_t* info_out, int32_t n_threads) {
  if (!s || !s->bvh || !s->triangles || !s->materials || !s->normals || !s->uvs || !info_out || s->n_nodes <= 0 ||
      s->n_triangles <= 0 || s->atlas_layers <= 0)
    return FSPT_E_INVALID;
  const int hw = std::max(1, std::min(64, n_threads > 0 ? n_threads : (int)std::thread::hardware_concurrency()));
  HostPool pool;
  ScenePrepass pre;
  const int rc = scene_prepass(pool, hw, s, pre);
  if (rc) { g_create_error = pre.error; return rc; }
  const size_t NI = pre.NI();
  info_out[0] = (int32_t)NI; info_out[1] = (int32_t)pre.mats.size(); info_out[2] = pre.ref[0];
  info_out[3] = pre.dielectric ? 1 : 0; info_out[4] = pre.max_depth;
  if (!node64_out) return FSPT_OK;  // first call: sizes only
  if (!tri48_out || !shaderec_out) return FSPT_E_INVALID;
  const int T = s->n_triangles;
  const size_t node_chunk = 4096;
  const int tri_chunk = 1024;
  const int n_node_items = NI ? (int)((NI + node_chunk - 1) / node_chunk) : 1, n_tri_items = (T + 3 + tri_chunk - 1) / tri_chunk;
  std::atomic<int> bad(-1);
  pool.run(n_node_items + n_tri_items, hw, [&](int item) {
    if (item < n_node_items) {
      const size_t k0 = (size_t)item * node_chunk, k1 = std::min(NI, k0 + node_chunk);
      int bl = 0, br = 0;
      const int b = pack_node_chunk(s, pre, k0, k1, node64_out + k0 * 16, &bl, &br);
      if (b >= 0) { int exp = -1; bad.compare_exchange_strong(exp, b); }
      return;
    }
    const int t0 = (item - n_node_items) * tri_chunk, t1 = std::min(T + 3, t0 + tri_chunk);
    pack_tri_chunk(s, pre, t0, t1, tri48_out + (size_t)t0 * 12, shaderec_out + (size_t)t0 * 48);
  });
  if (bad.load() >= 0) { g_create_error = "node " + std::to_string(bad.load()) + ": child index out of range"; return FSPT_E_INVALID; }
  if (mat_id_out) memcpy(mat_id_out, pre.mat_id.data(), (size_t)T * 4);
  return FSPT_OK;
}

int fspt_host_register(void* p, uint64_t bytes) {
  if (!p || !bytes) return FSPT_E_INVALID;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { (void)cudaGetLastError(); return FSPT_OK; }
  if (e != cudaSuccess) { (void)cudaGetLastError(); g_create_error = std::string("cudaHostRegister: ") + cudaGetErrorString(e); return FSPT_E_CUDA; }
  return FSPT_OK;
}
int fspt_host_unregister(void* p) {
  if (!p) return FSPT_E_INVALID;
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) { (void)cudaGetLastError(); g_create_error = std::string("cudaHostUnregister: ") + cudaGetErrorString(e); return FSPT_E_CUDA; }
  return FSPT_OK;
}

int fspt_scene_upload_wait(fspt_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  if (int rc = atlas_join(c)) return rc;
  return release_borrowed(c);
}

int fspt_clear(fspt_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaMemsetAsync(c->d_fb, 0, (size_t)c->n_pixels * 16, c->stream));
  c->next_tick = 0;
  c->accum_samples = 0;
  c->stats.samples = 0;
  return FSPT_OK;
}

int fspt_render(fspt_ctx* ctx, const fspt_frame_params* frame, uint32_t first_tick, int32_t n_samples,
                const float* rand_base_camera, const float* rand_base_tracer) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  if (!frame || n_samples < 0 || (n_samples > 0 && (!rand_base_camera || !rand_base_tracer)))
    return fail(c, FSPT_E_INVALID, "fspt_render: bad arguments");
  if (!c->has_scene) return fail(c, FSPT_E_STATE, "fspt_render before fspt_scene_upload");
  CK(cudaSetDevice(c->device));
  const FrameParams fp = make_frame(c, frame);
  c->ev_trace_used = 0;
  CK(cudaMemcpyAsync(c->d_stats + 4, c->d_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c->stream));
  // all rand bases of this call go up in one copy: [0,cap) camera, [cap,2cap) tracer
  {
    int rc0 = stage_rand_bases(c, rand_base_camera, rand_base_tracer, n_samples);
    if (rc0) return rc0;
  }
  CK(cudaEventRecord(c->ev_begin, c->stream));
  // samples in flight per wave: the path-state arrays hold wave_paths records; a tile of the frame keeps more samples
  // in flight than the whole frame would (3840x2160 whole: 8, one of 8 tiles: 64)
  const int wave_S = (int)std::max<size_t>(1, std::min<size_t>((size_t)c->wave_cap, c->wave_paths / ((size_t)fp.rw * fp.rh)));
  for (int done = 0; done < n_samples;) {
    const int S = std::min(wave_S, n_samples - done);
    int rc = render_wave(c, fp, first_tick + (uint32_t)done, S, c->d_rb + done, c->d_rb + c->rb_cap + done);
    if (rc) return rc;
    done += S;
  }
  CK(cudaEventRecord(c->ev_end, c->stream));
  c->render_timed = true;
  c->next_tick = first_tick + (uint32_t)n_samples;
  c->accum_samples += (uint64_t)n_samples;
  c->stats.samples += (uint64_t)n_samples * (uint64_t)fp.rw * (uint64_t)fp.rh;
  return FSPT_OK;
}

int fspt_synchronize(fspt_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  if (int rc = broadcast_phase2(c)) return rc;  // (joins the atlas thread on the root, see render_wave)
  if (int rc = atlas_join(c)) return rc;
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaStreamSynchronize(c->copy_stream));
  return FSPT_OK;
}

int fspt_resolve(fspt_ctx* ctx, const fspt_post_params* post, uint8_t* rgba8_out) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  if (!post || !rgba8_out) return fail(c, FSPT_E_INVALID, "fspt_resolve: NULL argument");
  CK(cudaSetDevice(c->device));
  dim3 blk(32, 8), grd((c->width + 31) / 32, (c->height + 7) / 8);
  const int use_div = c->accum_mode == 1;  // sum mode: every pixel is divided by its own sample count (alpha channel)
  k_post<<<grd, blk, 0, c->stream>>>(c->d_fb, c->d_rgba8, c->width, c->height, post->exposure, post->saturation,
                                     post->denoise ? 1 : 0, post->max_sigma, post->scale, use_div);
  c->stats.kernel_launches++;
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(rgba8_out, c->d_rgba8, (size_t)c->n_pixels * 4, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FSPT_OK;
}

int fspt_read_accum(fspt_ctx* ctx, float* out) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !out) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(out, c->d_fb, (size_t)c->n_pixels * 16, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FSPT_OK;
}

int fspt_write_accum(fspt_ctx* ctx, const float* in, uint32_t next_tick) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !in) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(c->d_fb, in, (size_t)c->n_pixels * 16, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->next_tick = next_tick;
  c->accum_samples = next_tick;
  return FSPT_OK;
}

int fspt_set_accum_mode(fspt_ctx* ctx, int32_t mode) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || (mode != 0 && mode != 1)) return FSPT_E_INVALID;
  c->accum_mode = mode;
  return FSPT_OK;
}

int fspt_accum_device_ptr(fspt_ctx* ctx, void** dptr, uint64_t* n_floats, uint64_t* n_samples) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  if (dptr) *dptr = c->d_fb;
  if (n_floats) *n_floats = (uint64_t)c->n_pixels * 4;
  if (n_samples) *n_samples = c->accum_samples;
  return FSPT_OK;
}

int fspt_set_accum_samples(fspt_ctx* ctx, uint64_t n) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  c->accum_samples = n;
  return FSPT_OK;
}

int fspt_debug_primary(fspt_ctx* ctx, const fspt_frame_params* frame, float rand_base_camera, int32_t* index_out,
                       float* t_out, int32_t* count_out, float* pos4_out, float* dir4_out) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !frame) return FSPT_E_INVALID;
  if (!c->has_scene) return fail(c, FSPT_E_STATE, "fspt_debug_primary before fspt_scene_upload");
  CK(cudaSetDevice(c->device));
  const FrameParams fp = make_frame(c, frame, true);  // always the whole frame
  const int P = c->n_pixels;
  {
    int rc0 = stage_rand_bases(c, &rand_base_camera, nullptr, 1);
    if (rc0) return rc0;
  }
  k_camera<<<(P + 255) / 256, 256, 0, c->stream>>>(fp, c->d_rb, P, make_fastdiv(1u, (unsigned long long)P), c->ps, c->d_cam_pos, c->d_cam_dir);
  c->stats.kernel_launches++;
  int rc = set_counts(c, P, 0);
  if (rc) return rc;
  c->ev_trace_used = 0;
  rc = launch_trace(c, 0, false, true);
  if (rc) return rc;
  CK(cudaStreamSynchronize(c->stream));
  // un-swizzle path slots -> pixels on the host
  std::vector<float4> ro(P), rdv(P);
  std::vector<int> cnt(P);
  CK(cudaMemcpy2D(ro.data(), 16, c->ps.rec + 0, 16 * FSPT_PATH_WORDS, 16, P, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy2D(rdv.data(), 16, c->ps.rec + 1, 16 * FSPT_PATH_WORDS, 16, P, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(cnt.data(), c->d_count_out, (size_t)P * 4, cudaMemcpyDeviceToHost));
  for (int j = 0; j < P; ++j) {
    int x, y;
    if (fp.tiled) {
      const int tiles_x = fp.width >> 3, tile = j >> 5, l = j & 31;
      x = ((tile % tiles_x) << 3) + (l & 7);
      y = ((tile / tiles_x) << 2) + (l >> 3);
    } else { x = j % fp.width; y = j / fp.width; }
    const size_t px = (size_t)y * fp.width + x;
    if (t_out) t_out[px] = ro[j].w;
    if (index_out) memcpy(&index_out[px], &rdv[j].w, 4);
    if (count_out) count_out[px] = cnt[j];
  }
  if (pos4_out) CK(cudaMemcpy(pos4_out, c->d_cam_pos, (size_t)P * 16, cudaMemcpyDeviceToHost));
  if (dir4_out) CK(cudaMemcpy(dir4_out, c->d_cam_dir, (size_t)P * 16, cudaMemcpyDeviceToHost));
  return pull_stats(c);
}

int fspt_debug_trace(fspt_ctx* ctx, const float* pos4, const float* dir4, int32_t n_rays, int32_t* index_out,
                     float* t_out, int32_t* count_out) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !pos4 || !dir4 || n_rays < 0) return FSPT_E_INVALID;
  if (!c->has_scene) return fail(c, FSPT_E_STATE, "fspt_debug_trace before fspt_scene_upload");
  CK(cudaSetDevice(c->device));
  std::vector<float4> ro, rdv;
  std::vector<int> cnt;
  for (int64_t done = 0; done < n_rays;) {
    const int n = (int)std::min<int64_t>((int64_t)c->wave_paths, n_rays - done);
    CK(cudaMemcpy2DAsync(c->ps.rec + 0, 16 * FSPT_PATH_WORDS, pos4 + 4 * done, 16, 16, n, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpy2DAsync(c->ps.rec + 1, 16 * FSPT_PATH_WORDS, dir4 + 4 * done, 16, 16, n, cudaMemcpyHostToDevice, c->stream));
    int rc = set_counts(c, n, 0);
    if (rc) return rc;
    c->ev_trace_used = 0;
    rc = launch_trace(c, 0, false, true);
    if (rc) return rc;
    ro.resize(n); rdv.resize(n); cnt.resize(n);
    CK(cudaMemcpy2DAsync(ro.data(), 16, c->ps.rec + 0, 16 * FSPT_PATH_WORDS, 16, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpy2DAsync(rdv.data(), 16, c->ps.rec + 1, 16 * FSPT_PATH_WORDS, 16, n, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(cnt.data(), c->d_count_out, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < n; ++i) {
      if (t_out) t_out[done + i] = ro[i].w;
      if (index_out) memcpy(&index_out[done + i], &rdv[i].w, 4);
      if (count_out) count_out[done + i] = cnt[i];
    }
    done += n;
  }
  return pull_stats(c);
}

int fspt_debug_last_color(fspt_ctx* ctx, float* out) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !out) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaMemcpyAsync(out, c->d_last_color, (size_t)c->n_pixels * 16, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return FSPT_OK;
}

int fspt_debug_math(fspt_ctx* ctx, int32_t fn, const float* x, const float* y, float* out, int32_t n) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !x || !out || n < 0) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  float* buf = nullptr;  // one allocation (x | y | out), released on every path
  const size_t stride = (size_t)n + 1;
  CK(cudaMalloc(&buf, 3 * stride * 4));
  struct Free { float* p; ~Free() { cudaFree(p); } } guard{buf};
  float *dx = buf, *dy = buf + stride, *dout = buf + 2 * stride;
  CK(cudaMemcpy(dx, x, (size_t)n * 4, cudaMemcpyHostToDevice));
  if (y) CK(cudaMemcpy(dy, y, (size_t)n * 4, cudaMemcpyHostToDevice)); else CK(cudaMemset(dy, 0, (size_t)n * 4));
  if (n) k_debug_math<<<(n + 255) / 256, 256, 0, c->stream>>>(fn, dx, dy, dout, n);
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaMemcpy(out, dout, (size_t)n * 4, cudaMemcpyDeviceToHost));
  return FSPT_OK;
}

__global__ void __launch_bounds__(256) k_read_bw(const uint4* __restrict__ buf, size_t n16, int iters, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n16; i += 4 * stride) {  // four independent loads in flight per thread
      const uint4 a = __ldcg(buf + i), b = __ldcg(buf + i + stride), c2 = __ldcg(buf + i + 2 * stride), d = __ldcg(buf + i + 3 * stride);
      acc ^= a.x ^ a.y ^ a.z ^ a.w ^ b.x ^ b.y ^ b.z ^ b.w ^ c2.x ^ c2.y ^ c2.z ^ c2.w ^ d.x ^ d.y ^ d.z ^ d.w;
    }
    for (; i < n16; i += stride) { const uint4 a = __ldcg(buf + i); acc ^= a.x ^ a.y ^ a.z ^ a.w; }
  }
  if (acc == 0x9e3779b9u) *sink = acc;  // keeps the loads alive
}

int fspt_debug_read_bandwidth(fspt_ctx* ctx, uint64_t bytes, int32_t iters, double* gb_per_s_out) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !gb_per_s_out || bytes < 4096 || iters < 1) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  uint4* buf = nullptr;
  unsigned* sink = nullptr;
  const size_t n16 = (size_t)bytes / 16;
  CK(cudaMalloc(&buf, n16 * 16));
  CK(cudaMalloc(&sink, 4));
  CK(cudaMemsetAsync(buf, 1, n16 * 16, c->stream));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  const int blocks = c->sm_count * 8;
  k_read_bw<<<blocks, 256, 0, c->stream>>>(buf, n16, 1, sink);  // warm-up: brings the buffer into L2 when it fits
  CK(cudaEventRecord(e0, c->stream));
  k_read_bw<<<blocks, 256, 0, c->stream>>>(buf, n16, iters, sink);
  CK(cudaEventRecord(e1, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  float ms = 0.0f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(buf); cudaFree(sink);
  *gb_per_s_out = (double)n16 * 16.0 * iters / (ms * 1e-3) / 1e9;
  return FSPT_OK;
}

int fspt_set_tile(fspt_ctx* ctx, int32_t x0, int32_t y0, int32_t w, int32_t h) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  if (w <= 0 || h <= 0 || x0 < 0 || y0 < 0 || (int64_t)x0 + w > c->width || (int64_t)y0 + h > c->height)
    return fail(c, FSPT_E_INVALID, "fspt_set_tile: rectangle %d,%d %dx%d is not inside the %dx%d frame", x0, y0, w, h, c->width, c->height);
  c->rx0 = x0; c->ry0 = y0; c->rw = w; c->rh = h;
  return FSPT_OK;
}

int fspt_comm_unique_id(uint8_t* id_out) {
  Ctx* c = nullptr;
  if (!id_out) return FSPT_E_INVALID;
  NcclApi* N = nccl_api();
  if (!N->error.empty()) return fail(c, FSPT_E_NCCL, "%s", N->error.c_str());
  ncclUniqueId id;
  NK(N->GetUniqueId(&id));
  static_assert(sizeof(id) == FSPT_COMM_ID_BYTES, "ncclUniqueId size");
  memcpy(id_out, &id, sizeof id);
  return FSPT_OK;
}

int fspt_comm_init(fspt_ctx* ctx, const uint8_t* id_bytes, int32_t rank, int32_t world) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !id_bytes || world < 1 || rank < 0 || rank >= world) return FSPT_E_INVALID;
  NcclApi* N = nccl_api();
  if (!N->error.empty()) return fail(c, FSPT_E_NCCL, "%s", N->error.c_str());
  CK(cudaSetDevice(c->device));
  comm_free(c);
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof id);
  NK(N->CommInitRank(&c->comm, world, id, rank));
  c->comm_rank = rank; c->comm_world = world;
  if (!c->d_hdr) CK(cudaMalloc(&c->d_hdr, 512));  // [0, 256) geometry header, [256, 512) atlas header
  return FSPT_OK;
}

int fspt_comm_destroy(fspt_ctx* ctx) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  comm_free(c);
  return FSPT_OK;
}

// sum of the accumulation targets of every rank -> root's target, in place, on the context's stream (ordered after the
// render kernels already enqueued, no host synchronisation).  In sum mode the alpha channel carries every pixel's sample
// count, so tiles, sample sets and unequal shards all resolve correctly on the root.
int fspt_reduce_accum(fspt_ctx* ctx, int32_t root) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  if (!c->comm) return fail(c, FSPT_E_STATE, "fspt_reduce_accum before fspt_comm_init");
  if (root < 0 || root >= c->comm_world) return fail(c, FSPT_E_INVALID, "fspt_reduce_accum: root %d of %d ranks", root, c->comm_world);
  if (c->accum_mode != 1) return fail(c, FSPT_E_STATE, "fspt_reduce_accum needs the sum accumulation mode (fspt_set_accum_mode(ctx, 1))");
  CK(cudaSetDevice(c->device));
  if (int rc = broadcast_phase2(c)) return rc;  // (a rank that rendered nothing since the broadcast)
  NcclApi* N = nccl_api();
  if (!c->ev_red0) { CK(cudaEventCreate(&c->ev_red0)); CK(cudaEventCreate(&c->ev_red1)); }
  CK(cudaEventRecord(c->ev_red0, c->stream));
  NK(N->Reduce(c->d_fb, c->d_fb, (size_t)c->n_pixels * 4, ncclFloat32, ncclSum, root, c->comm, c->stream));
  CK(cudaEventRecord(c->ev_red1, c->stream));
  c->reduce_timed = true;
  c->stats.kernel_launches++;
  return FSPT_OK;
}

int fspt_scene_broadcast(fspt_ctx* ctx, int32_t root) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  if (!c->comm) return fail(c, FSPT_E_STATE, "fspt_scene_broadcast before fspt_comm_init");
  if (root < 0 || root >= c->comm_world) return fail(c, FSPT_E_INVALID, "fspt_scene_broadcast: root %d of %d ranks", root, c->comm_world);
  const bool is_root = c->comm_rank == root;
  if (is_root && !c->has_scene) return fail(c, FSPT_E_STATE, "fspt_scene_broadcast: the root has no scene (fspt_scene_upload first)");
  CK(cudaSetDevice(c->device));
  int rc;
  if ((rc = broadcast_phase2(c))) return rc;  // (a previous broadcast nobody rendered from)
  if (!is_root && (rc = atlas_join(c))) return rc;
  NcclApi* N = nccl_api();
  SceneHeader h;
  memset(&h, 0, sizeof h);
  static_assert(sizeof(SceneHeader) <= 256, "header buffer");
  if (!c->ev_bcast) CK(cudaEventCreateWithFlags(&c->ev_bcast, cudaEventDisableTiming));
  if (is_root) {
    h.bytes_nodes = c->bytes_nodes; h.bytes_tris = c->bytes_tris; h.bytes_shade = c->bytes_shade;
    h.bytes_bins = c->bytes_bins;
    h.scene_bytes = c->scene_bytes;
    h.root_ref = c->sc.root_ref; h.n_tris = c->sc.n_tris; h.n_interior = c->sc.n_interior;
    h.atlas_res = c->sc.atlas_res; h.atlas_layers = c->sc.atlas_layers; h.env_w = c->sc.env_w; h.env_h = c->sc.env_h;
    h.n_bins = c->sc.n_bins;
    h.has_dielectric = c->has_dielectric ? 1 : 0;
    h.magic = 0x46535054;
    CK(cudaMemcpyAsync(c->d_hdr, &h, sizeof h, cudaMemcpyHostToDevice, c->stream));
  } else {
    CK(cudaStreamSynchronize(c->stream));  // nothing of the previous scene may still be in flight
    CK(cudaStreamSynchronize(c->copy_stream));
  }
  NK(N->Broadcast(c->d_hdr, c->d_hdr, sizeof h, ncclUint8, root, c->comm, c->stream));
  if (!is_root) {
    CK(cudaMemcpyAsync(&h, c->d_hdr, sizeof h, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h.magic != 0x46535054) return fail(c, FSPT_E_NCCL, "fspt_scene_broadcast: bad header from rank %d", root);
  }
  const size_t env_bytes = (size_t)h.env_w * h.env_h * 4;
  if ((rc = ensure(c, c->d_lin_env, c->cap_lin_env, env_bytes))) return rc;
  uint8_t* lin = reinterpret_cast<uint8_t*>(c->d_lin_env);
  if (is_root) {
    CK(cudaMemcpy2DFromArrayAsync(lin, (size_t)h.env_w * 4, c->env_arr, 0, 0, (size_t)h.env_w * 4, h.env_h,
                                  cudaMemcpyDeviceToDevice, c->stream));
  } else {
    // storage on the receiving side: same reuse rules as fspt_scene_upload
    c->has_scene = false;
    if ((rc = ensure(c, c->d_nodes, c->cap_nodes, h.bytes_nodes))) return rc;
    if ((rc = ensure(c, c->d_tris, c->cap_tris, h.bytes_tris))) return rc;
    if ((rc = ensure(c, c->d_shade, c->cap_shade, h.bytes_shade))) return rc;
    if ((rc = ensure(c, c->d_bins, c->cap_bins, h.bytes_bins))) return rc;
    if (!c->env_arr || c->env_W != h.env_w || c->env_H != h.env_h) {
      cudaResourceDesc rd = {};
      rd.resType = cudaResourceTypeArray;
      cudaTextureDesc td = {};
      td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
      td.filterMode = cudaFilterModePoint;
      td.readMode = cudaReadModeElementType;
      cudaChannelFormatDesc fmt = cudaCreateChannelDesc<uchar4>();
      if (c->sc.env) cudaDestroyTextureObject(c->sc.env);
      c->sc.env = 0;
      if (c->env_arr) cudaFreeArray(c->env_arr);
      c->env_arr = nullptr;
      CK(cudaMallocArray(&c->env_arr, &fmt, h.env_w, h.env_h));
      rd.res.array.array = c->env_arr;
      CK(cudaCreateTextureObject(&c->sc.env, &rd, &td, nullptr));
      c->env_W = h.env_w; c->env_H = h.env_h;
    }
  }
  NK(N->GroupStart());
  NK(N->Broadcast(c->d_nodes, c->d_nodes, h.bytes_nodes, ncclUint8, root, c->comm, c->stream));
  NK(N->Broadcast(c->d_tris, c->d_tris, h.bytes_tris, ncclUint8, root, c->comm, c->stream));
  NK(N->Broadcast(c->d_shade, c->d_shade, h.bytes_shade, ncclUint8, root, c->comm, c->stream));
  NK(N->Broadcast(c->d_bins, c->d_bins, h.bytes_bins, ncclUint8, root, c->comm, c->stream));
  NK(N->Broadcast(lin, lin, env_bytes, ncclUint8, root, c->comm, c->stream));
  NK(N->GroupEnd());
  c->stats.kernel_launches += 6;
  if (!is_root) {
    CK(cudaMemcpy2DToArrayAsync(c->env_arr, 0, 0, lin, (size_t)h.env_w * 4, (size_t)h.env_w * 4, h.env_h,
                                cudaMemcpyDeviceToDevice, c->stream));
    if ((rc = linear_tex(c, &c->nodes_tex, c->d_nodes, h.bytes_nodes))) return rc;
    c->bytes_nodes = h.bytes_nodes; c->bytes_tris = h.bytes_tris; c->bytes_shade = h.bytes_shade;
    c->bytes_bins = h.bytes_bins;
    c->sc.nodes = reinterpret_cast<const float4*>(c->d_nodes);
    c->sc.tris = reinterpret_cast<const float4*>(c->d_tris);
    c->sc.shade = reinterpret_cast<const float4*>(c->d_shade);
    c->sc.bins = reinterpret_cast<const float4*>(c->d_bins);
    c->sc.root_ref = h.root_ref; c->sc.n_tris = h.n_tris; c->sc.n_interior = h.n_interior;
    c->sc.atlas_res = h.atlas_res; c->sc.atlas_layers = h.atlas_layers; c->sc.env_w = h.env_w; c->sc.env_h = h.env_h;
    c->sc.n_bins = h.n_bins;
    set_env_constants(c->sc);
    c->has_dielectric = h.has_dielectric != 0;
    c->scene_bytes = h.scene_bytes;
    c->has_scene = true;
  }
  CK(cudaEventRecord(c->ev_bcast, c->stream));
  c->bcast_pending_root = root;  // phase 2: the atlas, behind the next primary traversal launch
  return FSPT_OK;
}

int fspt_set_param(fspt_ctx* ctx, int32_t key, int32_t value) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c) return FSPT_E_INVALID;
  switch (key) {
    case FSPT_PARAM_ANYHIT: c->anyhit = value ? 1 : 0; return FSPT_OK;
    case FSPT_PARAM_MAX_REFRACTIONS:
      if (value < 0 || value > 30000) return fail(c, FSPT_E_INVALID, "max_refractions out of range");
      c->max_refractions = value; return FSPT_OK;
    case FSPT_PARAM_SANITIZE_NAN: c->sanitize = value ? 1 : 0; return FSPT_OK;
  }
  return fail(c, FSPT_E_INVALID, "fspt_set_param: unknown key %d", key);
}

int fspt_get_stats(fspt_ctx* ctx, fspt_stats* out) {
  Ctx* c = reinterpret_cast<Ctx*>(ctx);
  if (!c || !out) return FSPT_E_INVALID;
  CK(cudaSetDevice(c->device));
  CK(cudaStreamSynchronize(c->stream));
  int rc = pull_stats(c);
  if (rc) return rc;
  float ms = 0.0f;
  if (c->render_timed && cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end) == cudaSuccess) c->stats.render_ms = ms;
  double tr = 0.0, sh = 0.0, pr = 0.0;
  for (size_t i = 0; i + 1 < c->ev_trace_used; i += 2) {
    float t = 0.0f;
    if (cudaEventElapsedTime(&t, c->ev_trace[i], c->ev_trace[i + 1]) != cudaSuccess) continue;
    const int tag = c->ev_tag[i / 2];
    (tag == 1 ? sh : tr) += t;
    if (tag == 2) pr += t;
  }
  c->stats.shade_ms = sh;
  c->stats.primary_trace_ms = pr;
  c->stats.reduce_ms = 0.0;
  if (c->reduce_timed && cudaEventElapsedTime(&ms, c->ev_red0, c->ev_red1) == cudaSuccess) c->stats.reduce_ms = ms;
  (void)cudaGetLastError();  // event queries must not leave a sticky status for the next launch check
  c->stats.trace_ms = tr;
  *out = c->stats;
  out->kernel_launches += c->atlas_launches.load();
  return FSPT_OK;
}

}  // extern "C"
