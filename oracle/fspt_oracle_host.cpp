/*
 * oracle/fspt_oracle_host.cpp -- TEST INFRASTRUCTURE.  Literal CPU restatements of the
 * reference's host-side scene compilers that feed the hot path:
 *   bvh.js (exact-sweep SAH builder + pre-order serialisation), main.js:360-392 (flatten),
 *   env_sampler.js (HDRi radiance bins).
 * JavaScript numbers are IEEE binary64, so everything here is double until the
 * Float32Array upload points (main.js:412-437).  Deliberately naive (one node object
 * per tree node, index lists copied per split, like the JS) -- the product has its own
 * fast builder (fspt_b200/csrc/bvh_builder.cpp) which tests compare against this one.
 *
 * PARITY PIN: tests/test_reference_js_pin.py runs the reference's own bvh.js / env_sampler.js in a real ECMAScript
 * engine (oracle/reference_js.py: Qt's QJSEngine through ctypes) and demands the same nodes, triangle order, depth and
 * bins, bit for bit, from this file and from the product's native builder.  oracle_pack_layer (the WebGL blit of
 * texture_packer.js) is compared with the blit shader's own text compiled for the CPU (tests/test_reference_pin.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <memory>
#include <vector>

#include "oracle_texunit.h"

namespace {

struct Box { /* bvh.js:93-143 BoundingBox */
  double mn[3], mx[3];
  Box() { for (int i = 0; i < 3; ++i) { mn[i] = INFINITY; mx[i] = -INFINITY; } }
  void addVertex(const double* v) { /* Vec3.min/Vec3.max = Math.min/Math.max, vector.js:55-61 */
    for (int i = 0; i < 3; ++i) { mn[i] = jsmin(v[i], mn[i]); mx[i] = jsmax(v[i], mx[i]); }
  }
  void addBox(const Box& b) {
    for (int i = 0; i < 3; ++i) { mn[i] = jsmin(mn[i], b.mn[i]); mx[i] = jsmax(mx[i], b.mx[i]); }
  }
  double surfaceArea() const { /* bvh.js:137-142 */
    double xl = mx[0] - mn[0], yl = mx[1] - mn[1], zl = mx[2] - mn[2];
    return (xl * yl + xl * zl + yl * zl) * 2;
  }
  double centroid(int axis) const { return (mn[axis] + mx[axis]) * 0.5; } /* bvh.js:130-135 */
  /* Math.min / Math.max: NaN-propagating, and -0 < +0 (ECMA-262 21.3.2.24-25) */
  static double jsmin(double a, double b) {
    if (a != a || b != b) return NAN;
    if (a < b) return a;
    if (b < a) return b;
    return signbit(a) ? a : b;
  }
  static double jsmax(double a, double b) {
    if (a != a || b != b) return NAN;
    if (a < b) return b;
    if (b < a) return a;
    return signbit(a) ? b : a;
  }
};

struct Tri { const double* v; Box box; };  /* v: current verts; box: Triangle.boundingBox (may be stale) */

struct BNode { /* bvh.js:145-198 Node */
  std::vector<int> idx[3];
  Box box;
  bool leaf = false;
  int splitIndex = -1, splitAxis = -1;
  std::unique_ptr<BNode> left, right;
};

struct Builder {
  std::vector<Tri> tris;
  int maxTris = 4;
  int depth = 0;
  bool failed = false;

  void setSplit(BNode& n) { /* bvh.js:168-197 */
    double bestCost = INFINITY;
    double parentSurfaceArea = n.box.surfaceArea();
    for (int axis = 0; axis < 3; ++axis) {
      Box bbFront, bbBack;
      const std::vector<int>& ic = n.idx[axis];
      size_t len = ic.size();
      std::vector<double> surfacesFront(len), surfacesBack(len);
      for (size_t i = 0; i < len; ++i) {
        bbFront.addBox(tris[ic[i]].box);
        bbBack.addBox(tris[ic[len - 1 - i]].box);
        surfacesFront[i] = bbFront.surfaceArea();
        surfacesBack[i] = bbBack.surfaceArea();
      }
      for (size_t i = 0; i < len; ++i) {
        double sAf = surfacesFront[i];
        double sAb = surfacesBack[len - 1 - i];
        double cost = 1 + (sAf / parentSurfaceArea) * 1 * (double)(i + 1) +
                      (sAb / parentSurfaceArea) * 1 * (double)(len - 1 - i);
        if (cost < bestCost) {
          bestCost = cost;
          n.splitIndex = (int)i + 1;
          n.splitAxis = axis;
        }
      }
    }
  }

  std::unique_ptr<BNode> buildTree(std::vector<int> (&indices)[3], int d) { /* bvh.js:19-31 */
    depth = std::max(d, depth);
    std::unique_ptr<BNode> root(new BNode());
    for (int a = 0; a < 3; ++a) root->idx[a] = indices[a];
    for (int i : root->idx[0]) /* BoundingBox.addNode, bvh.js:122-128 */
      for (int k = 0; k < 3; ++k) root->box.addVertex(tris[i].v + 3 * k);
    setSplit(*root);
    int ax = root->splitAxis < 0 ? 0 : root->splitAxis; /* `root.splitAxis || 0` */
    if ((int)root->idx[ax].size() <= maxTris) {
      root->leaf = true;
      return root;
    }
    if (root->splitAxis < 0 || d > 4096) { failed = true; root->leaf = true; return root; } /* JS would throw */
    /* _constructCachedIndexList, bvh.js:52-76 */
    int sa = root->splitAxis, si = root->splitIndex;
    std::vector<int> L[3], R[3];
    L[sa].assign(root->idx[sa].begin(), root->idx[sa].begin() + si);
    R[sa].assign(root->idx[sa].begin() + si, root->idx[sa].end());
    if (R[sa].empty()) { failed = true; root->leaf = true; return root; } /* JS: infinite recursion */
    std::vector<char> setLeft(tris.size(), 0);
    for (int i : L[sa]) setLeft[i] = 1;
    for (int axis = 0; axis < 3; ++axis) {
      if (axis == sa) continue;
      for (int id : root->idx[axis]) (setLeft[id] ? L[axis] : R[axis]).push_back(id);
    }
    root->left = buildTree(L, d + 1);
    root->right = buildTree(R, d + 1);
    for (int a = 0; a < 3; ++a) { root->idx[a].clear(); root->idx[a].shrink_to_fit(); } /* clearTempBuffers */
    return root;
  }
};

struct Flat { std::vector<float>* nodes; std::vector<int>* order; };

/* serializeTree (bvh.js:33-50) fused with the flatten loop (main.js:366-392) */
int flatten(const BNode* n, std::vector<float>& nodes, std::vector<int32_t>& order) {
  int self = (int)(nodes.size() / 9);
  nodes.resize(nodes.size() + 9);
  int32_t left = 0, right = 0, triIndex = -1; /* leaf: e.left/e.right undefined -> Int32Array 0 */
  if (n->leaf) {
    triIndex = (int32_t)order.size(); /* trianglesBuffer.length / 3 / 3, main.js:369 */
    for (int i : n->idx[0]) order.push_back(i); /* getTriangles() walks indices[0], bvh.js:156-161 */
  } else {
    left = flatten(n->left.get(), nodes, order);
    right = flatten(n->right.get(), nodes, order);
  }
  float* p = nodes.data() + (size_t)self * 9;
  memcpy(p + 0, &left, 4); /* maskBVHBuffer, main.js:272-282 */
  memcpy(p + 1, &right, 4);
  memcpy(p + 2, &triIndex, 4);
  for (int k = 0; k < 3; ++k) { p[3 + k] = (float)n->box.mn[k]; p[6 + k] = (float)n->box.mx[k]; }
  return self;
}

}  // namespace

extern "C" {

/* new BVH(tris, maxTris) + serializeTree + flatten.
 * verts: n_tris*9 doubles (post-transform world space, as in Triangle.verts).
 * nodes_out: capacity 2*n_tris*9 floats; order_out: n_tris ints (source triangle of each triTex slot).
 * Returns node count, or -1 where the JS would crash / recurse forever. */
int oracle_bvh_build2(const double* verts, const double* box_verts, int n_tris, int max_tris, float* nodes_out,
                      int32_t* order_out, int32_t* depth_out);
int oracle_bvh_build(const double* verts, int n_tris, int max_tris, float* nodes_out, int32_t* order_out,
                     int32_t* depth_out) {
  return oracle_bvh_build2(verts, nullptr, n_tris, max_tris, nodes_out, order_out, depth_out);
}
/* box_verts: the vertices Triangle.boundingBox was computed from (bvh.js:208); main.js:337-348 (`normalize`)
 * rescales Triangle.verts afterwards and leaves the boxes stale.  NULL = same as verts. */
int oracle_bvh_build2(const double* verts, const double* box_verts, int n_tris, int max_tris, float* nodes_out,
                      int32_t* order_out, int32_t* depth_out) {
  const double* bv = box_verts ? box_verts : verts;
  Builder b;
  b.maxTris = max_tris;
  b.tris.resize(n_tris);
  for (int i = 0; i < n_tris; ++i) {
    b.tris[i].v = verts + (size_t)i * 9;
    for (int k = 0; k < 3; ++k) b.tris[i].box.addVertex(bv + (size_t)i * 9 + 3 * k); /* bvh.js:208 */
  }
  std::vector<int> idx[3];
  for (int a = 0; a < 3; ++a) {
    idx[a].resize(n_tris);
    for (int i = 0; i < n_tris; ++i) idx[a][i] = i;
    /* _sortIndices, bvh.js:78-90: Array.prototype.sort is stable (ES2019) */
    std::stable_sort(idx[a].begin(), idx[a].end(), [&](int i1, int i2) {
      return b.tris[i1].box.centroid(a) < b.tris[i2].box.centroid(a);
    });
  }
  std::unique_ptr<BNode> root = b.buildTree(idx, 0);
  if (b.failed) return -1;
  std::vector<float> nodes;
  std::vector<int32_t> order;
  flatten(root.get(), nodes, order);
  memcpy(nodes_out, nodes.data(), nodes.size() * 4);
  memcpy(order_out, order.data(), order.size() * 4);
  if (depth_out) *depth_out = b.depth;
  return (int)(nodes.size() / 9);
}

/* ProcessEnvRadiance, env_sampler.js:1-74.  data = RGBA8 as getImageData returns it, row 0 = top.
 * boxes_out capacity: 4*w*h u16.  Returns the number of u16 written (4 per bin). */
int oracle_env_bins(const uint8_t* data, int width, int height, uint16_t* boxes_out, int capacity) {
  auto getRadiance = [&](double x, double y) -> double { /* env_sampler.js:14-22 */
    if (x != floor(x) || y != floor(y) || x < 0 || y < 0) return NAN; /* data[fractional] === undefined */
    size_t off = ((size_t)y * ((size_t)width * 4)) + ((size_t)x * 4);
    if (off + 3 >= (size_t)width * height * 4) return NAN;
    double c0 = data[off], c1 = data[off + 1], c2 = data[off + 2], c3 = data[off + 3];
    double power = pow(2.0, c3 - 128);
    double n0 = power * c0 / 255.0, n1 = power * c1 / 255.0, n2 = power * c2 / 255.0;
    return 0.2126 * n0 + 0.7152 * n1 + 0.0722 * n2;
  };
  double totalRadiance = 0, brightestTexel = 0;
  for (int y = 0; y < height; ++y)
    for (int x = 0; x < width; ++x) {
      double rad = getRadiance(x, y);
      brightestTexel = (rad != rad || brightestTexel != brightestTexel) ? NAN : (rad > brightestTexel ? rad : brightestTexel);
      totalRadiance += rad;
    }
  double minRadiance = std::max(totalRadiance / 64, brightestTexel / 2);
  std::vector<double> boxes;
  struct Rec {
    decltype(getRadiance)& gr; double minRadiance; std::vector<double>& boxes; size_t cap;
    void biSplit(double radiance, double x0, double y0, double x1, double y1) { /* env_sampler.js:26-50 */
      if (boxes.size() >= cap) return;
      if (radiance <= minRadiance || (y1 - y0) * (x1 - x0) < 2) {
        boxes.push_back(x0); boxes.push_back(y0); boxes.push_back(x1); boxes.push_back(y1);
        return;
      }
      double subRadiance = 0;
      bool vertSplit = x1 - x0 > y1 - y0;
      double xs = x1;
      double ys = (y1 - y0) / 2 + y0;
      if (vertSplit) { xs = (x1 - x0) / 2 + x0; ys = y1; }
      for (double x = x0; x < xs; x++)
        for (double y = y0; y < ys; y++) subRadiance += gr(x, y);
      biSplit(subRadiance, x0, y0, xs, ys);
      if (vertSplit) biSplit(radiance - subRadiance, xs, y0, x1, y1);
      else biSplit(radiance - subRadiance, x0, ys, x1, y1);
    }
  } rec{getRadiance, minRadiance, boxes, (size_t)capacity};
  rec.biSplit(totalRadiance, 0, 0, width, height);
  int n = (int)std::min(boxes.size(), (size_t)capacity);
  for (int i = 0; i < n; ++i) { /* new Uint16Array(boxes): ToUint16 */
    double v = boxes[i];
    if (v != v || isinf(v)) { boxes_out[i] = 0; continue; }
    double tr = v < 0 ? ceil(v) : floor(v);
    boxes_out[i] = (uint16_t)(uint64_t)(int64_t)fmod(tr, 65536.0);
  }
  return n;
}

/* One image layer of TexturePacker.getPixels(): the fragment program of WebGLTextureWriter evaluated per fragment
 * (texture_packer.js:103-121) on a texture uploaded by setAndDrawTexture (:159-176) and read back by getPixels
 * (:178-184).  rgba8: w*h*4, row 0 = image top.  out: res*res*4, row y = gl_FragCoord.y. */
void oracle_pack_layer(const uint8_t* rgba8, int w, int h, int res, int corrected, const int32_t* swizzle, uint8_t* out) {
  for (int y = 0; y < res; ++y)
    for (int x = 0; x < res; ++x) {
      float u = ((float)x + 0.5f) / (float)res, v = ((float)y + 0.5f) / (float)res; /* gl_FragCoord.xy / dims */
      v = 1.0f - v;                                                                /* uv.y = 1.0 - uv.y */
      const om::v4 t = om::tu_texture_2d(rgba8, w, h, corrected, u, v);            /* texture(tex, uv), oracle_texunit.h */
      const float c[4] = {t.x, t.y, t.z, t.w};
      float sc[4];
      for (int k = 0; k < 4; ++k) sc[k] = c[swizzle ? swizzle[k] : k];             /* c[k] = copy[swizzle[k]] */
      float rgb[3] = {sc[0] * sc[3], sc[1] * sc[3], sc[2] * sc[3]};                /* vec4(c.rgb * c.a, 1.0) */
      uint8_t* o = out + ((size_t)y * res + x) * 4;
      for (int k = 0; k < 3; ++k) o[k] = om::tu_quant8(rgb[k]);
      o[3] = 255;
    }
}

}  // extern "C"
