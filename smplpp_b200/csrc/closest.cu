// Closest point on the posed mesh for every marker point of every frame, and the re-seated attachment (face index +
// triangle vertex weights) the MoSh body stage derives from it.
//
// Reference: node/node.cpp:949-1001.  After the QP step every IkTask's point p = calcActualPos() + tangents_ * phi is
// projected onto the CURRENT mesh with igl::point_mesh_squared_distance (libigl v2.4.0, cmake/libigl.cmake:9 - not under
// /root/reference; an AABB-tree accelerated exact point-triangle search), then faceIdx_ = closest face and
// vertexWeights_ = calcTriangleVertexWeights(closest point, face vertices) (GeometryUtils.h:42-52).
//
// Here: one CTA per frame.  The frame's 6890 vertices are staged in shared memory (82 680 B, two CTAs per SM); thread t
// owns (point t % n, triangle slice t / n) and scans its slice with the exact region-based point-triangle test
// (Ericson, Real-Time Collision Detection 5.1.5 - the same seven Voronoi regions igl's point_simplex_squared_distance
// resolves by clamped barycentrics).  Lanes of a warp mostly share the triangle, so the nine vertex loads per test are
// shared-memory broadcasts.  The per-point minimum over slices breaks ties towards the lower face index, which makes
// the result independent of the launch geometry.  Exhaustive over the n x F pairs of a frame (41 x 13 776 = 565 k), but
// the exact test only runs where the bound |p - a| - r_t (r_t = reach of triangle t from its first vertex) can still beat
// the running minimum, which starts at the distance to the nearest vertex; a bounding volume hierarchy is not worth its
// per-frame rebuild at this mesh size.
#include <cuda_fp16.h>

#include <cfloat>

#include "common.cuh"

using namespace sb;

namespace
{
struct CpParams
{
  int V, F, n, B;
  const int32_t * faces;   // (F, 3) 0-based
  const float * verts;     // (B, V, 3)
  const float * points;    // (B, n, 3)
  int32_t * face_out;      // (B, n) 0-based closest face
  float * closest_out;     // (B, n, 3) or null
  float * sqdist_out;      // (B, n) or null
  float * weights_out;     // (B, n, 3) or null: calcTriangleVertexWeights(closest, face)
};

struct V3
{
  float x, y, z;
};
__device__ __forceinline__ V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
__device__ __forceinline__ float dot(V3 a, V3 b) { return fmaf(a.x, b.x, fmaf(a.y, b.y, a.z * b.z)); }
__device__ __forceinline__ V3 madd(V3 a, V3 d, float s) { return {fmaf(d.x, s, a.x), fmaf(d.y, s, a.y), fmaf(d.z, s, a.z)}; }
__device__ __forceinline__ V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// closest point of triangle (a, b, c) to p
__device__ __forceinline__ V3 closest_on_triangle(V3 p, V3 a, V3 b, V3 c)
{
  const V3 ab = sub(b, a), ac = sub(c, a), ap = sub(p, a);
  const float d1 = dot(ab, ap), d2 = dot(ac, ap);
  if(d1 <= 0.f && d2 <= 0.f) return a;
  const V3 bp = sub(p, b);
  const float d3 = dot(ab, bp), d4 = dot(ac, bp);
  if(d3 >= 0.f && d4 <= d3) return b;
  const float vc = d1 * d4 - d3 * d2;
  if(vc <= 0.f && d1 >= 0.f && d3 <= 0.f) return madd(a, ab, d1 / (d1 - d3));
  const V3 cp = sub(p, c);
  const float d5 = dot(ab, cp), d6 = dot(ac, cp);
  if(d6 >= 0.f && d5 <= d6) return c;
  const float vb = d5 * d2 - d1 * d6;
  if(vb <= 0.f && d2 >= 0.f && d6 <= 0.f) return madd(a, ac, d2 / (d2 - d6));
  const float va = d3 * d6 - d5 * d4;
  if(va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) return madd(b, sub(c, b), (d4 - d3) / ((d4 - d3) + (d5 - d6)));
  const float denom = 1.f / (va + vb + vc);
  return madd(madd(a, ab, vb * denom), ac, vc * denom);
}

constexpr int CP_THREADS = 512;

__global__ void __launch_bounds__(CP_THREADS, 2) closest_point_kernel(const CpParams p)
{
  extern __shared__ __align__(16) float cp_sm[];
  float * s_v = cp_sm;                                        // V * 3
  float * s_best = s_v + 3 * p.V;                             // slices * n   squared distance
  int * s_face = reinterpret_cast<int *>(s_best + CP_THREADS); // slices * n   face
  __half * s_r = reinterpret_cast<__half *>(s_face + CP_THREADS); // F            per-triangle reach, rounded up
  const int f = blockIdx.x;
  const int tid = threadIdx.x;
  {
    // the frame is 82 680 B = 8 mod 16: copy as float2 (the buffer is at least 8-byte aligned when V is even), else scalar
    const float * gv = p.verts + static_cast<size_t>(f) * p.V * 3;
    if(((reinterpret_cast<uintptr_t>(gv) & 7) == 0) && ((3 * p.V) % 2 == 0))
    {
      const float2 * g2 = reinterpret_cast<const float2 *>(gv);
      float2 * s2 = reinterpret_cast<float2 *>(s_v);
      for(int i = tid; i < 3 * p.V / 2; i += CP_THREADS) s2[i] = __ldg(g2 + i);
    }
    else
      for(int i = tid; i < 3 * p.V; i += CP_THREADS) s_v[i] = __ldg(gv + i);
  }
  const int slices = CP_THREADS / p.n;             // >= 1 (n <= CP_THREADS is checked by the host)
  const int m = tid % p.n, sl = tid / p.n;
  const bool on = sl < slices;
  V3 pt = {0.f, 0.f, 0.f};
  if(on)
  {
    const float * gp = p.points + (static_cast<size_t>(f) * p.n + m) * 3;
    pt = {__ldg(gp), __ldg(gp + 1), __ldg(gp + 2)};
  }
  __syncthreads();
  // ---- pass 1: r_t = the longer of the two edges leaving triangle t's first vertex a.  Every point x of the triangle
  // has |x - a| <= r_t, hence dist(p, t) >= |p - a| - r_t: the exact test of pass 3 only runs where this bound can still
  // beat the running minimum.  (Kept per triangle, as fp16 rounded UP: a frame-wide maximum is useless on meshes with a
  // few long triangles - the synthetic hull mesh has 0.6 cm .. 75 cm edges.) ----
  for(int t = tid; t < p.F; t += CP_THREADS)
  {
    const int i0 = __ldg(p.faces + 3 * t), i1 = __ldg(p.faces + 3 * t + 1), i2 = __ldg(p.faces + 3 * t + 2);
    const V3 a = {s_v[3 * i0], s_v[3 * i0 + 1], s_v[3 * i0 + 2]};
    const V3 ab = sub({s_v[3 * i1], s_v[3 * i1 + 1], s_v[3 * i1 + 2]}, a), ac = sub({s_v[3 * i2], s_v[3 * i2 + 1], s_v[3 * i2 + 2]}, a);
    s_r[t] = __float2half_ru(sqrtf(fmaxf(dot(ab, ab), dot(ac, ac))) * 1.0001f);
  }
  // ---- pass 2: the nearest VERTEX of each point bounds its distance to the mesh from above ----
  float best = FLT_MAX;
  if(on)
  {
    const int per = (p.V + slices - 1) / slices;
    const int v0 = sl * per, v1 = min(p.V, v0 + per);
    for(int v = v0; v < v1; v++)
    {
      const V3 d = sub({s_v[3 * v], s_v[3 * v + 1], s_v[3 * v + 2]}, pt);
      best = fminf(best, dot(d, d));
    }
    s_best[sl * p.n + m] = best;
  }
  __syncthreads();
  if(on)
    for(int s = 0; s < slices; s++) best = fminf(best, s_best[s * p.n + m]);
  __syncthreads();
  // ---- pass 3: exact point-triangle test on the triangles that survive the bound ----
  best = best * 1.0001f + 1e-12f; // the vertex distance is an upper bound: keep it strictly above the true minimum
  int best_face = 0x7fffffff;
  if(on)
  {
    float sb = sqrtf(best) * 1.0001f;
    const int per = (p.F + slices - 1) / slices;
    const int t0 = sl * per, t1 = min(p.F, t0 + per);
    auto exact = [&](int t, V3 a) {
      const int i1 = __ldg(p.faces + 3 * t + 1), i2 = __ldg(p.faces + 3 * t + 2);
      const V3 b = {s_v[3 * i1], s_v[3 * i1 + 1], s_v[3 * i1 + 2]};
      const V3 c = {s_v[3 * i2], s_v[3 * i2 + 1], s_v[3 * i2 + 2]};
      const V3 d = sub(closest_on_triangle(pt, a, b, c), pt);
      const float d2 = dot(d, d);
      if(d2 < best) // ascending t: ties keep the lower face index
      {
        best = d2, best_face = t;
        sb = sqrtf(best) * 1.0001f;
      }
    };
    // the bound test is a dependent chain of ~16 instructions (index, vertex, reach): four triangles per trip give
    // the scheduler independent work; the rare survivors take the exact test in ascending order
    constexpr int U = 4;
    int t = t0;
    for(; t + U <= t1; t += U)
    {
      V3 a[U];
      float over[U];
#pragma unroll
      for(int u = 0; u < U; u++)
      {
        const int i0 = __ldg(p.faces + 3 * (t + u));
        a[u] = {s_v[3 * i0], s_v[3 * i0 + 1], s_v[3 * i0 + 2]};
        const V3 ap = sub(pt, a[u]);
        const float reach = sb + __half2float(s_r[t + u]);
        over[u] = dot(ap, ap) - reach * reach;
      }
#pragma unroll
      for(int u = 0; u < U; u++)
      {
        if(over[u] > 0.f) continue; // (sb only shrinks: a test that passed with the older, larger reach is merely conservative)
        exact(t + u, a[u]);
      }
    }
    for(; t < t1; t++)
    {
      const int i0 = __ldg(p.faces + 3 * t);
      const V3 a = {s_v[3 * i0], s_v[3 * i0 + 1], s_v[3 * i0 + 2]};
      const V3 ap = sub(pt, a);
      const float reach = sb + __half2float(s_r[t]);
      if(dot(ap, ap) > reach * reach) continue;
      exact(t, a);
    }
    s_best[sl * p.n + m] = best;
    s_face[sl * p.n + m] = best_face;
  }
  __syncthreads();
  if(tid < p.n)
  {
    // a slice that found no triangle below the vertex bound reports face 0x7fffffff and loses every comparison
    for(int s = 1; s < slices; s++)
    {
      const float d2 = s_best[s * p.n + tid];
      const int fc = s_face[s * p.n + tid];
      if(fc != 0x7fffffff && (best_face == 0x7fffffff || d2 < best || (d2 == best && fc < best_face))) best = d2, best_face = fc;
    }
    const size_t o = static_cast<size_t>(f) * p.n + tid;
    p.face_out[o] = best_face;
    if(p.sqdist_out) p.sqdist_out[o] = best;
    if((p.closest_out || p.weights_out) && best_face != 0x7fffffff) // NaN points never beat the sentinel
    {
      const int i0 = p.faces[3 * best_face], i1 = p.faces[3 * best_face + 1], i2 = p.faces[3 * best_face + 2];
      const V3 a = {s_v[3 * i0], s_v[3 * i0 + 1], s_v[3 * i0 + 2]};
      const V3 b = {s_v[3 * i1], s_v[3 * i1 + 1], s_v[3 * i1 + 2]};
      const V3 c = {s_v[3 * i2], s_v[3 * i2 + 1], s_v[3 * i2 + 2]};
      const V3 q = closest_on_triangle(pt, a, b, c);
      if(p.closest_out) p.closest_out[3 * o] = q.x, p.closest_out[3 * o + 1] = q.y, p.closest_out[3 * o + 2] = q.z;
      if(p.weights_out)
      {
        // calcTriangleVertexWeights (GeometryUtils.h:42-52): w_i ~ |(v_{i+1} - q) x (v_{i+2} - q)|, normalised to sum 1
        const V3 r0 = cross(sub(b, q), sub(c, q)), r1 = cross(sub(c, q), sub(a, q)), r2 = cross(sub(a, q), sub(b, q));
        const float w0 = sqrtf(dot(r0, r0)), w1 = sqrtf(dot(r1, r1)), w2 = sqrtf(dot(r2, r2));
        const float s = w0 + w1 + w2;
        p.weights_out[3 * o] = w0 / s, p.weights_out[3 * o + 1] = w1 / s, p.weights_out[3 * o + 2] = w2 / s;
      }
    }
  }
}
} // namespace

extern "C" int smplpp_closest_points(const smplpp_model_t * model, void * stream, int64_t batch, int64_t n_points,
                                     const float * vertices, const float * points, int32_t * face_idx, float * closest,
                                     float * sq_dist, float * vertex_weights)
{
  if(!model || batch < 1 || !vertices || !points || !face_idx)
    return fail(SMPLPP_ERR_INVALID, "IkTask", "Failed to project points onto the mesh!");
  if(n_points < 1 || n_points > CP_THREADS)
    return fail(SMPLPP_ERR_INVALID, "IkTask", "Failed to project points onto the mesh! (1 .. 512 points per frame)");
  const ModelDev & d = model->d;
  if(d.F < 1) return fail(SMPLPP_ERR_INVALID, "IkTask", "Failed to project points onto the mesh! (the model has no faces)");
  CpParams p;
  p.V = d.V;
  p.F = d.F;
  p.n = static_cast<int>(n_points);
  p.B = static_cast<int>(batch);
  p.faces = d.faces;
  p.verts = vertices;
  p.points = points;
  p.face_out = face_idx;
  p.closest_out = closest;
  p.sqdist_out = sq_dist;
  p.weights_out = vertex_weights;
  const size_t smem = (static_cast<size_t>(3) * d.V + 2 * CP_THREADS) * sizeof(float) + (static_cast<size_t>(d.F) + 8) * sizeof(__half);
  // per launch: the attribute is per device, and one process may drive several
  SB_CUDA(cudaFuncSetAttribute(closest_point_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 115000));
  if(smem > 115000) return fail(SMPLPP_ERR_INVALID, "IkTask", "Failed to project points onto the mesh! (mesh too large)");
  closest_point_kernel<<<static_cast<unsigned>(batch), CP_THREADS, smem, as_stream(stream)>>>(p);
  SB_LAUNCHED();
  return SMPLPP_OK;
}
