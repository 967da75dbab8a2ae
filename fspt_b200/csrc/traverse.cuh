// traverse.cuh -- closest-hit BVH traversal + Moller-Trumbore, persistent warps with dynamic ray fetch.
//
// Replaces intersectScene / rayBoxIntersect / processLeaf / rayTriangleIntersect of the reference
// (tracer.fs:300-326,355-404; counter variant bvh_test.fs:173-221).  Per ray the sequence of visited nodes,
// the pruning comparisons and the triangle-test order are exactly the reference's, so (index, t, count)
// are bit-identical:
//   - near child first, `leftHit > rightHit` sends the ray right (ties go left), far child deferred
//     (tracer.fs:382-392); deferred nodes are NOT re-tested against result.t when popped (tracer.fs:401);
//   - a leaf tests exactly LEAF_SIZE = 4 consecutive triangles starting at its first one, over-reading
//     into the next leaf (tracer.fs:355-364); strict `<` keeps the first hit on ties (tracer.fs:359);
//   - slab test: (b - o) * (1/d), IEEE division, min/max = minNum/maxNum, `tMax >= tMin && tMax > 0`
//     (tracer.fs:317-326); no FMA contraction anywhere (file is compiled with --fmad=false).
// The two box tests the reference also performs on leaf visits (children 0,0 -> root box, tracer.fs:377-378)
// have no observable effect and are skipped.
//
// Execution model: one ray per lane, a warp keeps running until fewer than REFILL lanes still have a ray,
// then the idle lanes are refilled from a global queue with one atomicAdd per warp (__ballot_sync +
// popc prefix = the compaction), so short rays do not wait for the longest ray of their warp.  Inside the
// loop the warp votes each iteration whether to run an interior step or a leaf step (majority of lanes).
// Results go back into the path record plus one hit/miss byte per record position for the shading kernel.
#pragma once
#include "camera.cuh"
#include "device_common.cuh"

struct TraceArgs {
  const float4* nodes;
  const float4* tris;
  cudaTextureObject_t nodes_tex;  // the node / triangle arrays again as linear textures (second L1 data pipe)
  cudaTextureObject_t tris_tex;
  int root_ref;
  PathState ps;           // rays in, hits out (words 0..2 of the path record)
  const int* list_shadow; // record positions of the paths that cast a shadow ray (continuation rays: every record)
  const int* counts;      // counts[0] = #continuation, counts[1] = #shadow
  int* next;              // work-fetch cursor (zeroed before launch)
  unsigned long long* stats;  // [0] rays, [1] node visits, [2] leaf visits
  int* count_out;         // per-slot visit count (debug / bvh_test mode) or NULL
  unsigned char* hit_flag;  // per record position: 1 = hit (read coalesced by k_shade) or NULL
  FrameParams f;          // CAMERA mode: primary rays are generated in the fetch instead of being read
  const float* rb_cam;
  int n_samples;
  int anyhit;             // 1: hit-or-miss rays (shadow rays, marked last-bounce rays) stop at their first intersection
};

#ifndef TRACE_THREADS
#define TRACE_THREADS 128
#endif
#ifndef TRACE_NODE_TEX
#define TRACE_NODE_TEX 15  /* bit k: word k of the node record comes through the texture pipe */
#endif
#ifndef TRACE_TRI_TEX
#define TRACE_TRI_TEX 0    /* bit k: word k of a triangle record comes through the texture pipe */
#endif
#ifndef TRACE_DUP_LOADS
#define TRACE_DUP_LOADS 0
#endif
#ifndef TRACE_REFILL
#define TRACE_REFILL 12
#endif
#ifndef TRACE_INT_WEIGHT
#define TRACE_INT_WEIGHT 1
#define TRACE_LEAF_WEIGHT 1
#endif

__device__ __forceinline__ float slab(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz,
                                      float ox, float oy, float oz, float ix, float iy, float iz) {
  const float t1x = (bminx - ox) * ix, t2x = (bmaxx - ox) * ix;
  const float t1y = (bminy - oy) * iy, t2y = (bmaxy - oy) * iy;
  const float t1z = (bminz - oz) * iz, t2z = (bmaxz - oz) * iz;
  const float tMax = fminf(fminf(fmaxf(t1x, t2x), fmaxf(t1y, t2y)), fmaxf(t1z, t2z));
  const float tMin = fmaxf(fmaxf(fminf(t1x, t2x), fminf(t1y, t2y)), fminf(t1z, t2z));
  return (tMax >= tMin && tMax > 0.0f) ? tMin : FSPT_MAX_T;
}

// Both child boxes at once with Blackwell's packed f32x2 pipe: FADD2 / FMUL2 perform two independent IEEE
// round-to-nearest f32 operations per instruction (lane 0 = left child, lane 1 = right child), so the 24
// subtract/multiply operations of the two slab tests issue as 12 instructions.  Bitwise identical to slab().
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ void slab_pair(const float4 w0, const float4 w1, const float4 w2, f32x2 ox2, f32x2 oy2, f32x2 oz2,
                                          f32x2 ix2, f32x2 iy2, f32x2 iz2, float& lh, float& rh) {
  // record words: w0 = (lmin.x rmin.x lmin.y rmin.y)  w1 = (lmin.z rmin.z lmax.x rmax.x)  w2 = (lmax.y rmax.y lmax.z rmax.z)
  const f32x2 t1x = mul2(sub2(pack2(w0.x, w0.y), ox2), ix2), t1y = mul2(sub2(pack2(w0.z, w0.w), oy2), iy2);
  const f32x2 t1z = mul2(sub2(pack2(w1.x, w1.y), oz2), iz2), t2x = mul2(sub2(pack2(w1.z, w1.w), ox2), ix2);
  const f32x2 t2y = mul2(sub2(pack2(w2.x, w2.y), oy2), iy2), t2z = mul2(sub2(pack2(w2.z, w2.w), oz2), iz2);
  float a1x, b1x, a1y, b1y, a1z, b1z, a2x, b2x, a2y, b2y, a2z, b2z;
  unpack2(t1x, a1x, b1x); unpack2(t1y, a1y, b1y); unpack2(t1z, a1z, b1z);
  unpack2(t2x, a2x, b2x); unpack2(t2y, a2y, b2y); unpack2(t2z, a2z, b2z);
  const float lMax = fminf(fminf(fmaxf(a1x, a2x), fmaxf(a1y, a2y)), fmaxf(a1z, a2z));
  const float lMin = fmaxf(fmaxf(fminf(a1x, a2x), fminf(a1y, a2y)), fminf(a1z, a2z));
  const float rMax = fminf(fminf(fmaxf(b1x, b2x), fmaxf(b1y, b2y)), fmaxf(b1z, b2z));
  const float rMin = fmaxf(fmaxf(fminf(b1x, b2x), fminf(b1y, b2y)), fminf(b1z, b2z));
  lh = (lMax >= lMin && lMax > 0.0f) ? lMin : FSPT_MAX_T;
  rh = (rMax >= rMin && rMax > 0.0f) ? rMin : FSPT_MAX_T;
}

// rayTriangleIntersect with e1/e2 precomputed; predicates are the reference's conditions, un-negated, so
// NaN behaves identically (tracer.fs:305,309,312,314)
__device__ __forceinline__ float tri_test(const float4 q0, const float4 q1, const float4 q2, float ox, float oy,
                                          float oz, float dx, float dy, float dz) {
  const float e1x = q0.w, e1y = q1.x, e1z = q1.y, e2x = q1.z, e2y = q1.w, e2z = q2.x;
  const float px = dy * e2z - e2y * dz, py = dz * e2x - e2z * dx, pz = dx * e2y - e2x * dy;  // cross(dir,e2)
  const float det = e1x * px + e1y * py + e1z * pz;
  if (fabsf(det) < FSPT_EPSILON) return FSPT_MAX_T;
  const float invDet = 1.0f / det;
  const float tx = ox - q0.x, ty = oy - q0.y, tz = oz - q0.z;
  const float u = (tx * px + ty * py + tz * pz) * invDet;
  if (u < 0.0f || u > 1.0f) return FSPT_MAX_T;
  const float qx = ty * e1z - e1y * tz, qy = tz * e1x - e1z * tx, qz = tx * e1y - e1x * ty;  // cross(t,e1)
  const float v = (dx * qx + dy * qy + dz * qz) * invDet;
  if (v < 0.0f || u + v > 1.0f) return FSPT_MAX_T;
  const float dist = (e2x * qx + e2y * qy + e2z * qz) * invDet;
  return dist > FSPT_EPSILON ? dist : FSPT_MAX_T;
}

// CAMERA = the primary launch of a render wave: the ray of slot `my` is generated on the fly (camera.fs) and the whole
// ray + hit record is written at retirement, so the camera pass, its 32 B/path of writes and this launch's
// record reads disappear.
// NODE_TEX = false: the node array is too large for a linear texture (2^27 texels), everything goes through the LSU.
template <bool WRITE_COUNT, bool CAMERA, bool NODE_TEX>
__global__ void __launch_bounds__(TRACE_THREADS) k_trace(const TraceArgs A) {
  const unsigned lane = threadIdx.x & 31u;
  const unsigned FULL = 0xffffffffu;
  const int n_cont = A.counts[0];
  const int total = n_cont + A.counts[1];

  int stack[FSPT_STACK];
  int cur = FSPT_SENTINEL, sp = 0;
  int slot = -1, kind = 0, cnt = 0, item = 0;
  float ox = 0, oy = 0, oz = 0, dx = 0, dy = 0, dz = 0;
  f32x2 ox2 = 0, oy2 = 0, oz2 = 0, ix2 = 0, iy2 = 0, iz2 = 0;  // (v, v) pairs for the packed slab test
  float tbest = FSPT_MAX_T;
  int ibest = -1;
  unsigned long long n_rays = 0, n_nodes = 0, n_leaves = 0;
  bool drained = false;
  bool boolean_ray = false;  // only hit-or-miss is consumed (tracer.fs:502, :509 at the last bounce)

  for (;;) {
    const bool need = (cur == FSPT_SENTINEL);
    const bool retire = need && slot >= 0;
    if (retire) {  // retire the finished ray
      if (CAMERA) {
        st_path(A.ps.ro(slot), make_float4(ox, oy, oz, tbest));
        st_path(A.ps.rd(slot), make_float4(dx, dy, dz, __int_as_float(ibest)));
      } else if (kind == 0) {
        st_path_w(A.ps.ro(slot), tbest);
        st_path_w(A.ps.rd(slot), __int_as_float(ibest));
      } else {
        st_path_w(A.ps.sd(slot), __int_as_float((ibest == -1) ? 2 : 3));
      }
      if (WRITE_COUNT) A.count_out[slot] = cnt;
      if (A.hit_flag && kind == 0) A.hit_flag[item] = (ibest != -1);
      n_nodes += (unsigned long long)cnt;
    }
    if (retire) slot = -1;
    if (!drained) {
      const unsigned m = __ballot_sync(FULL, need);
      if (m) {
        const int leader = __ffs(m) - 1;
        const int want = __popc(m);
        int base = 0;
        if ((int)lane == leader) base = atomicAdd(A.next, want);
        base = __shfl_sync(FULL, base, leader);
        if (base + want >= total) drained = true;
        if (need) {
          const int my = base + __popc(m & ((1u << lane) - 1u));
          if (my < total) {
            item = my;
            if (CAMERA) {
              kind = 0;
              slot = my;
              v3 o, d;
              int px, py;
              camera_ray(A.f, A.rb_cam, A.n_samples, my, o, d, px, py);
              ox = o.x; oy = o.y; oz = o.z;
              dx = d.x; dy = d.y; dz = d.z;
            } else {
              kind = my >= n_cont;
              slot = kind ? ld_list(A.list_shadow + (my - n_cont)) : my;
              const float4 o4 = ld_path(A.ps.ro(slot));
              const float4 d4 = ld_path(kind ? A.ps.sd(slot) : A.ps.rd(slot));
              ox = o4.x; oy = o4.y; oz = o4.z;
              dx = d4.x; dy = d4.y; dz = d4.z;
              // shadow rays always; continuation rays when k_shade marked the path's last bounce (index word = -2)
              boolean_ray = A.anyhit && (kind == 1 || __float_as_int(d4.w) == -2);
            }
            const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;  // `vec3 inverse = 1.0 / ray.dir`, tracer.fs:318
            ox2 = pack2(ox, ox); oy2 = pack2(oy, oy); oz2 = pack2(oz, oz);
            ix2 = pack2(ix, ix); iy2 = pack2(iy, iy); iz2 = pack2(iz, iz);
            tbest = FSPT_MAX_T; ibest = -1; cnt = 0;
            stack[0] = FSPT_SENTINEL; sp = 1;
            cur = A.root_ref;
            n_rays++;
          }
        }
      }
    }
    if (__ballot_sync(FULL, cur != FSPT_SENTINEL) == 0u) break;

    const int refill = drained ? 1 : TRACE_REFILL;
    for (;;) {
      // Warp-level phase scheduling: every iteration runs ONE step for the larger group of lanes -- those
      // standing on an interior node or those standing on a leaf -- and the other group waits.  (A classic
      // while-while loop lets a few long interior walks hold the whole warp: measured 6.8 of 32 lanes active.)
      const bool is_int = cur >= 0;
      const bool is_leaf = !is_int && cur != FSPT_SENTINEL;
      const int ni = __popc(__ballot_sync(FULL, is_int)), nl = __popc(__ballot_sync(FULL, is_leaf));
      if (ni + nl < refill) break;
      if (ni * TRACE_INT_WEIGHT >= nl * TRACE_LEAF_WEIGHT) {
        // ---- interior node: both child boxes from one 64-byte record --------------------------------
        if (is_int) {
          cnt++;
          // The record's four words are split between the two L1 data pipes (texture fetch / LSU load): the
          // kernel is bound by L1 wavefronts (ncu: l1tex data-pipe ~60 % busy), not by issue slots.
          const float4* np = A.nodes + 4 * (size_t)cur;
          const float4 a = ((TRACE_NODE_TEX & 1) && NODE_TEX) ? tex1Dfetch<float4>(A.nodes_tex, 4 * cur) : __ldg(np);
          const float4 b = ((TRACE_NODE_TEX & 2) && NODE_TEX) ? tex1Dfetch<float4>(A.nodes_tex, 4 * cur + 1) : __ldg(np + 1);
          const float4 c = ((TRACE_NODE_TEX & 4) && NODE_TEX) ? tex1Dfetch<float4>(A.nodes_tex, 4 * cur + 2) : __ldg(np + 2);
          const float4 df = ((TRACE_NODE_TEX & 8) && NODE_TEX) ? tex1Dfetch<float4>(A.nodes_tex, 4 * cur + 3) : __ldg(np + 3);
          int4 d = make_int4(__float_as_int(df.x), __float_as_int(df.y), 0, 0);
#if TRACE_DUP_LOADS  /* sensitivity experiment: issue the record's loads a second time through the LSU pipe */
          {
            float4 e0, e1, e2, e3;
            asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e0.x), "=f"(e0.y), "=f"(e0.z), "=f"(e0.w) : "l"(np));
            asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e1.x), "=f"(e1.y), "=f"(e1.z), "=f"(e1.w) : "l"(np + 1));
            asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e2.x), "=f"(e2.y), "=f"(e2.z), "=f"(e2.w) : "l"(np + 2));
            asm volatile("ld.global.ca.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(e3.x), "=f"(e3.y), "=f"(e3.z), "=f"(e3.w) : "l"(np + 3));
            if (e0.x + e1.x + e2.x + e3.x == 12345.678f) d.x = 0;  // keep the loads alive
          }
#endif
          float lh, rh;
          slab_pair(a, b, c, ox2, oy2, oz2, ix2, iy2, iz2, lh, rh);
          const bool tl = lh < tbest, tr = rh < tbest;
          const bool right_first = lh > rh;                   // tracer.fs:384 (ties go left)
          if (tl && tr) {
            stack[sp++] = right_first ? d.x : d.y;            // deferred child, tracer.fs:391
            cur = right_first ? d.y : d.x;
          } else if (tl || tr) {
            cur = tl ? d.x : d.y;
          } else {
            cur = stack[--sp];
          }
        }
      } else {
        // ---- leaf: 4 consecutive triangles ----------------------------------------------------------
        if (is_leaf) {
          cnt++;
          n_leaves++;
          const int first = ~cur;
          const float4* tp = A.tris + 3 * (size_t)first;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 q0 = (TRACE_TRI_TEX & 1) ? tex1Dfetch<float4>(A.tris_tex, 3 * (first + k)) : __ldg(tp + 3 * k);
            const float4 q1 = (TRACE_TRI_TEX & 2) ? tex1Dfetch<float4>(A.tris_tex, 3 * (first + k) + 1) : __ldg(tp + 3 * k + 1);
            const float4 q2 = (TRACE_TRI_TEX & 4) ? tex1Dfetch<float4>(A.tris_tex, 3 * (first + k) + 2) : __ldg(tp + 3 * k + 2);
            const float res = tri_test(q0, q1, q2, ox, oy, oz, dx, dy, dz);
            if (res < tbest) { ibest = first + k; tbest = res; }
          }
          cur = stack[--sp];
          if (!CAMERA && boolean_ray && ibest != -1) cur = FSPT_SENTINEL;  // any hit settles it
        }
      }
    }
  }
  // per-warp statistics -> 3 atomics per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    n_rays += __shfl_xor_sync(FULL, n_rays, o);
    n_nodes += __shfl_xor_sync(FULL, n_nodes, o);
    n_leaves += __shfl_xor_sync(FULL, n_leaves, o);
  }
  if (lane == 0) {
    atomicAdd(A.stats + 0, n_rays);
    atomicAdd(A.stats + 1, n_nodes);
    atomicAdd(A.stats + 2, n_leaves);
  }
}
