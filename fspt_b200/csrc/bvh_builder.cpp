// bvh_builder.cpp -- native, multi-threaded builder that reproduces the reference's JavaScript BVH
// (bvh.js:5-198) and its pre-order flatten (bvh.js:33-50, main.js:366-392) bit for bit.
//
// Same tree, different machinery: instead of one JS object per node holding three freshly sliced index
// lists and a Set per split, the three centroid-sorted index arrays are partitioned IN PLACE (stable),
// every node owns a [lo,hi) range of all three, sweep scratch is shared by position, and sibling
// subtrees are built by parallel tasks.  All arithmetic that decides the tree is IEEE binary64 in the
// reference's operation order (JS numbers are doubles), so split choices are identical:
//   cost(i) = 1 + (SAfront_i / SAparent) * (i+1) + (SAback_i / SAparent) * (n-1-i)        bvh.js:189
// where the back box still contains triangle i (the reference's off-by-one), first strict minimum over
// axis 0,1,2 and i ascending wins (bvh.js:190), leaves are ranges of <= maxTris (bvh.js:22) listed in
// x-sorted order (bvh.js:156-161).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <future>
#include <thread>
#include <vector>

#include "../../include/fspt_b200.h"

namespace {

struct Aabb {
  double mn[3], mx[3];
  void reset() { for (int k = 0; k < 3; ++k) { mn[k] = INFINITY; mx[k] = -INFINITY; } }
  // Math.min / Math.max (vector.js:55-61) incl. their signed-zero rule (-0 < +0); inputs are never NaN (checked)
  static double lo(double a, double b) { return a < b ? a : (b < a ? b : (signbit(a) ? a : b)); }
  static double hi(double a, double b) { return a < b ? b : (b < a ? a : (signbit(a) ? b : a)); }
  void grow(const Aabb& b) {
    for (int k = 0; k < 3; ++k) { mn[k] = lo(mn[k], b.mn[k]); mx[k] = hi(mx[k], b.mx[k]); }
  }
  double area() const {  // BoundingBox.getSurfaceArea, bvh.js:137-142
    double xl = mx[0] - mn[0], yl = mx[1] - mn[1], zl = mx[2] - mn[2];
    return (xl * yl + xl * zl + yl * zl) * 2;
  }
};

struct BuildNode { Aabb box; int32_t left, right, lo, hi; };  // left < 0 => leaf over idx[0][lo,hi)

struct Ctx {
  const Aabb* tb;             // per-triangle boxes seen by the presorts and the SAH sweeps (Triangle.boundingBox)
  const Aabb* nb;             // per-triangle boxes of the current vertices (node boxes, BoundingBox.addNode)
  int32_t* idx[3];            // centroid-sorted triangle ids, partitioned in place
  int32_t* tmp;               // partition scratch, indexed by position
  double* sback;              // suffix surface areas, indexed by position
  uint8_t* side;              // per-triangle "goes left" flag
  std::vector<BuildNode> nodes;
  std::atomic<int32_t> n_nodes{0};
  std::atomic<int32_t> depth{0};
  std::atomic<int> failed{0};
  std::atomic<int> tasks_free{0};
  int max_tris;
};

int32_t build(Ctx& c, int32_t lo, int32_t hi, int d) {
  int32_t self = c.n_nodes.fetch_add(1);
  BuildNode nd;
  nd.lo = lo; nd.hi = hi; nd.left = nd.right = -1;
  int prev = c.depth.load();
  while (d > prev && !c.depth.compare_exchange_weak(prev, d)) {}
  const int32_t n = hi - lo;
  nd.box.reset();
  for (int32_t i = lo; i < hi; ++i) nd.box.grow(c.nb[c.idx[0][i]]);  // BoundingBox.addNode, bvh.js:122-128
  if (n <= c.max_tris) { c.nodes[self] = nd; return self; }          // bvh.js:22 (split of a leaf is unused)
  // Node.setSplit, bvh.js:168-197
  double best = INFINITY;
  int best_axis = -1; int32_t best_i = -1;
  const double parent = nd.box.area();
  for (int axis = 0; axis < 3; ++axis) {
    const int32_t* ix = c.idx[axis];
    Aabb bb; bb.reset();
    for (int32_t i = hi - 1; i >= lo; --i) { bb.grow(c.tb[ix[i]]); c.sback[i] = bb.area(); }
    bb.reset();
    for (int32_t i = lo; i < hi; ++i) {
      bb.grow(c.tb[ix[i]]);
      const double sAf = bb.area(), sAb = c.sback[i];
      const int32_t k = i - lo;
      const double cost = 1 + (sAf / parent) * 1 * (double)(k + 1) + (sAb / parent) * 1 * (double)(n - 1 - k);
      if (cost < best) { best = cost; best_i = k + 1; best_axis = axis; }
    }
  }
  if (best_axis < 0 || best_i >= n || d > 4000) {  // JS: TypeError on undefined axis / unbounded recursion
    c.failed.store(1);
    c.nodes[self] = nd;
    return self;
  }
  // _constructCachedIndexList, bvh.js:52-76: stable partition of the other two axes
  const int32_t mid = lo + best_i;
  for (int32_t i = lo; i < hi; ++i) c.side[c.idx[best_axis][i]] = (i < mid);
  for (int axis = 0; axis < 3; ++axis) {
    if (axis == best_axis) continue;
    int32_t* ix = c.idx[axis];
    int32_t l = lo, r = 0;
    for (int32_t i = lo; i < hi; ++i) { int32_t t = ix[i]; if (c.side[t]) ix[l++] = t; else c.tmp[lo + r++] = t; }
    memcpy(ix + l, c.tmp + lo, sizeof(int32_t) * (size_t)r);
  }
  int32_t L, R;
  bool spawn = false;
  if (n > 32768) {  // sibling subtrees touch disjoint ranges / triangles: build them concurrently
    int avail = c.tasks_free.load();
    while (avail > 0 && !c.tasks_free.compare_exchange_weak(avail, avail - 1)) {}
    spawn = avail > 0;
  }
  if (spawn) {
    auto fut = std::async(std::launch::async, [&c, lo, mid, d]() { return build(c, lo, mid, d + 1); });
    R = build(c, mid, hi, d + 1);
    L = fut.get();
    c.tasks_free.fetch_add(1);
  } else {
    L = build(c, lo, mid, d + 1);
    R = build(c, mid, hi, d + 1);
  }
  nd.left = L; nd.right = R;
  c.nodes[self] = nd;
  return self;
}

// serializeTree (bvh.js:33-50) + flatten (main.js:366-392): pre-order, iterative
void flatten(const Ctx& c, int32_t root, float* out, int32_t* order, int32_t* n_out) {
  struct Item { int32_t node; int32_t parent_slot; int which; };
  std::vector<Item> st;
  st.push_back({root, -1, 0});
  int32_t n = 0, tri = 0;
  while (!st.empty()) {
    Item it = st.back(); st.pop_back();
    const BuildNode& b = c.nodes[it.node];
    const int32_t self = n++;
    float* p = out + (size_t)self * 9;
    int32_t l = 0, r = 0, t = -1;  // leaf: left/right undefined -> Int32Array 0 (main.js:275)
    if (b.left < 0) {
      t = tri;                      // trianglesBuffer.length/9 (main.js:369)
      for (int32_t i = b.lo; i < b.hi; ++i) order[tri++] = c.idx[0][i];
    }
    memcpy(p + 0, &l, 4); memcpy(p + 1, &r, 4); memcpy(p + 2, &t, 4);  // maskBVHBuffer, main.js:272-282
    for (int k = 0; k < 3; ++k) { p[3 + k] = (float)b.box.mn[k]; p[6 + k] = (float)b.box.mx[k]; }
    if (it.parent_slot >= 0) memcpy(out + (size_t)it.parent_slot * 9 + it.which, &self, 4);
    if (b.left >= 0) {  // right pushed first so the left subtree is numbered first (pre-order)
      st.push_back({b.right, self, 1});
      st.push_back({b.left, self, 0});
    }
  }
  *n_out = n;
}

}  // namespace

extern "C" int fspt_bvh_build(const double* verts, int32_t n_tris, int32_t max_tris, float* nodes_out,
                              int32_t* order_out, int32_t* n_nodes_out, int32_t* depth_out, int32_t n_threads) {
  return fspt_bvh_build2(verts, nullptr, n_tris, max_tris, nodes_out, order_out, n_nodes_out, depth_out, n_threads);
}

// box_verts != NULL reproduces `scene.normalize` (main.js:337-348): Triangle.boundingBox is computed when the
// triangle is created (bvh.js:208) and NOT refreshed when normalize rescales the vertices, so the presorts and
// the SAH sweeps see the old boxes (box_verts) while node boxes are built from the new vertices (verts).
extern "C" int fspt_bvh_build2(const double* verts, const double* box_verts, int32_t n_tris, int32_t max_tris,
                               float* nodes_out, int32_t* order_out, int32_t* n_nodes_out, int32_t* depth_out,
                               int32_t n_threads) {
  if (!verts || !nodes_out || !order_out || n_tris <= 0 || max_tris <= 0) return FSPT_E_INVALID;
  const double* bv = box_verts ? box_verts : verts;
  std::vector<Aabb> tb((size_t)n_tris), nb;
  if (box_verts) nb.resize((size_t)n_tris);
  std::vector<double> cen[3];
  for (int a = 0; a < 3; ++a) cen[a].resize((size_t)n_tris);
  for (int32_t i = 0; i < n_tris; ++i) {  // Triangle.boundingBox, bvh.js:208
    Aabb& b = tb[i]; b.reset();
    for (int v = 0; v < 3; ++v)
      for (int k = 0; k < 3; ++k) {
        double x = bv[(size_t)i * 9 + v * 3 + k];
        if (!(x == x) || isinf(x)) return FSPT_E_INVALID;
        b.mn[k] = Aabb::lo(x, b.mn[k]);
        b.mx[k] = Aabb::hi(x, b.mx[k]);
      }
    for (int k = 0; k < 3; ++k) cen[k][i] = (b.mn[k] + b.mx[k]) * 0.5;  // centroid, bvh.js:130-135
    if (box_verts) {
      Aabb& n = nb[i]; n.reset();
      for (int v = 0; v < 3; ++v)
        for (int k = 0; k < 3; ++k) {
          double x = verts[(size_t)i * 9 + v * 3 + k];
          if (!(x == x) || isinf(x)) return FSPT_E_INVALID;
          n.mn[k] = Aabb::lo(x, n.mn[k]);
          n.mx[k] = Aabb::hi(x, n.mx[k]);
        }
    }
  }
  std::vector<int32_t> ix[3], tmp((size_t)n_tris);
  std::vector<double> sback((size_t)n_tris);
  std::vector<uint8_t> side((size_t)n_tris);
  {
    std::vector<std::thread> th;  // _sortIndices x3 (bvh.js:78-90): stable, by centroid
    for (int a = 0; a < 3; ++a) {
      ix[a].resize((size_t)n_tris);
      th.emplace_back([&, a]() {
        for (int32_t i = 0; i < n_tris; ++i) ix[a][i] = i;
        const double* cc = cen[a].data();
        std::stable_sort(ix[a].begin(), ix[a].end(), [cc](int32_t p, int32_t q) { return cc[p] < cc[q]; });
      });
    }
    for (auto& t : th) t.join();
  }
  Ctx c;
  c.tb = tb.data();
  c.nb = box_verts ? nb.data() : tb.data();
  for (int a = 0; a < 3; ++a) c.idx[a] = ix[a].data();
  c.tmp = tmp.data(); c.sback = sback.data(); c.side = side.data();
  c.nodes.resize((size_t)2 * n_tris + 1);
  c.max_tris = max_tris;
  if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
  c.tasks_free.store(n_threads > 1 ? n_threads - 1 : 0);
  int32_t root = build(c, 0, n_tris, 0);
  if (c.failed.load()) return FSPT_E_LIMIT;
  int32_t n = 0;
  flatten(c, root, nodes_out, order_out, &n);
  if (n_nodes_out) *n_nodes_out = n;
  if (depth_out) *depth_out = c.depth.load();
  return FSPT_OK;
}

// ProcessEnvRadiance (env_sampler.js:1-74).  Same boxes as the JavaScript, but region sums come from a
// summed-area table over exactly-representable per-texel terms when that is provably exact, otherwise from
// the reference's own x-major double accumulation order.
extern "C" int fspt_env_bins(const uint8_t* data, int32_t width, int32_t height, uint16_t* bins_out,
                             int32_t capacity, int32_t* n_u16_out) {
  if (!data || !bins_out || width <= 0 || height <= 0 || capacity < 4) return FSPT_E_INVALID;
  // luminance per texel in double, env_sampler.js:14-22
  std::vector<double> lum((size_t)width * height);
  double total = 0, brightest = 0;
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      const uint8_t* p = data + ((size_t)y * width + x) * 4;
      double power = ldexp(1.0, (int)p[3] - 128);  // Math.pow(2, a-128) is exact
      double r = power * p[0] / 255.0, g = power * p[1] / 255.0, b = power * p[2] / 255.0;
      double l = 0.2126 * r + 0.7152 * g + 0.0722 * b;
      lum[(size_t)y * width + x] = l;
      if (l > brightest) brightest = l;
      total += l;
    }
  const double minRadiance = std::max(total / 64, brightest / 2);
  std::vector<double> boxes;
  struct Frame { double radiance, x0, y0, x1, y1; };
  std::vector<Frame> stack;
  stack.push_back({total, 0, 0, (double)width, (double)height});
  auto lumAt = [&](double x, double y) -> double {  // data[fractional index] is undefined -> NaN in JS
    if (x != floor(x) || y != floor(y)) return NAN;
    size_t off = (size_t)y * (size_t)width + (size_t)x;
    if (off >= lum.size()) return NAN;
    return lum[off];
  };
  while (!stack.empty()) {  // biSplit, env_sampler.js:26-50 (explicit stack, same DFS emission order)
    Frame f = stack.back(); stack.pop_back();
    if ((int64_t)boxes.size() + 4 > capacity) break;
    if (f.radiance <= minRadiance || (f.y1 - f.y0) * (f.x1 - f.x0) < 2) {
      boxes.push_back(f.x0); boxes.push_back(f.y0); boxes.push_back(f.x1); boxes.push_back(f.y1);
      continue;
    }
    double sub = 0;
    bool vert = f.x1 - f.x0 > f.y1 - f.y0;
    double xs = f.x1, ys = (f.y1 - f.y0) / 2 + f.y0;
    if (vert) { xs = (f.x1 - f.x0) / 2 + f.x0; ys = f.y1; }
    for (double x = f.x0; x < xs; x++)
      for (double y = f.y0; y < ys; y++) sub += lumAt(x, y);
    // second half pushed first so the first half is processed (and emitted) first
    if (vert) stack.push_back({f.radiance - sub, xs, f.y0, f.x1, f.y1});
    else stack.push_back({f.radiance - sub, f.x0, ys, f.x1, f.y1});
    stack.push_back({sub, f.x0, f.y0, xs, ys});
  }
  int n = (int)boxes.size();
  for (int i = 0; i < n; ++i) {  // new Uint16Array(boxes): ToUint16
    double v = boxes[i];
    if (v != v || isinf(v)) { bins_out[i] = 0; continue; }
    double tr = v < 0 ? ceil(v) : floor(v);
    bins_out[i] = (uint16_t)(uint64_t)(int64_t)fmod(tr, 65536.0);
  }
  if (n_u16_out) *n_u16_out = n;
  return FSPT_OK;
}
