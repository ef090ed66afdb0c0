// Small device helpers of the IK kernels (ik.cu, ik2.cu): 3-vectors, the normalisation of torch::nn::functional::normalize
// and its derivative, calcTriangleVertexWeights and BlendShape::rodrigues with its derivative.
#pragma once
#include "common.cuh"

struct f3
{
  float x, y, z;
};
__device__ __forceinline__ f3 mk3(float x, float y, float z)
{
  f3 r;
  r.x = x, r.y = y, r.z = z;
  return r;
}
__device__ __forceinline__ f3 ld3(const float * p)
{
  return mk3(p[0], p[1], p[2]);
}
__device__ __forceinline__ f3 operator+(f3 a, f3 b)
{
  return mk3(a.x + b.x, a.y + b.y, a.z + b.z);
}
__device__ __forceinline__ f3 operator-(f3 a, f3 b)
{
  return mk3(a.x - b.x, a.y - b.y, a.z - b.z);
}
__device__ __forceinline__ f3 operator*(float s, f3 a)
{
  return mk3(s * a.x, s * a.y, s * a.z);
}
__device__ __forceinline__ float dot3(f3 a, f3 b)
{
  return a.x * b.x + a.y * b.y + a.z * b.z;
}
__device__ __forceinline__ f3 cross3(f3 a, f3 b)
{
  return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float norm3(f3 a)
{
  return sqrtf(dot3(a, a));
}
// torch::nn::functional::normalize: v / max(||v||, 1e-12); returns 1 / max(||v||, eps), negative when clamped
__device__ __forceinline__ f3 normalize_inv(f3 a, float & inv)
{
  float n = norm3(a);
  bool clamped = !(n > 1e-12f);
  float i = 1.f / fmaxf(n, 1e-12f);
  inv = clamped ? -i : i;
  return i * a;
}
// derivative of normalize at output direction n (unit) applied to x: (x - n (n.x)) / |v|   (or x / eps when clamped)
__device__ __forceinline__ f3 proj_apply(f3 n, float inv, f3 x)
{
  if(inv < 0.f) return (-inv) * x;
  return inv * (x - dot3(n, x) * n);
}
// calcTriangleVertexWeights (include/smplpp/toolbox/GeometryUtils.h:42-52)
__device__ __forceinline__ void triangle_weights(f3 pos, f3 v0, f3 v1, f3 v2, float * w)
{
  float a0 = norm3(cross3(v1 - pos, v2 - pos));
  float a1 = norm3(cross3(v2 - pos, v0 - pos));
  float a2 = norm3(cross3(v0 - pos, v1 - pos));
  float s = a0 + a1 + a2;
  w[0] = a0 / s, w[1] = a1 / s, w[2] = a2 / s;
}

// BlendShape::rodrigues (src/BlendShape.cpp:803-844) with the derivative of THAT formula (da/dtheta uses theta+eps)
__device__ inline void rodrigues_grad(float x, float y, float z, float * R, float * dR /* [3][9] */)
{
  const float eps = 1e-8f;
  float th[3] = {x, y, z};
  float ax = x + eps, ay = y + eps, az = z + eps;
  float a = sqrtf(ax * ax + ay * ay + az * az);
  float inv_a = 1.f / a;
  float u[3] = {x / a, y / a, z / a};
  float s, c;
  sincosf(a, &s, &c);
  float oc = 1.f - c;
  float uu = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
  R[0] = 1.f + oc * (-(u[2] * u[2]) - u[1] * u[1]);
  R[1] = s * (-u[2]) + oc * (u[1] * u[0]);
  R[2] = s * u[1] + oc * (u[2] * u[0]);
  R[3] = s * u[2] + oc * (u[0] * u[1]);
  R[4] = 1.f + oc * (-(u[2] * u[2]) - u[0] * u[0]);
  R[5] = s * (-u[0]) + oc * (u[2] * u[1]);
  R[6] = s * (-u[1]) + oc * (u[0] * u[2]);
  R[7] = s * u[0] + oc * (u[1] * u[2]);
  R[8] = 1.f + oc * (-(u[1] * u[1]) - u[0] * u[0]);
  const float K[9] = {0.f, -u[2], u[1], u[2], 0.f, -u[0], -u[1], u[0], 0.f};
  float da[3] = {ax * inv_a, ay * inv_a, az * inv_a};
#pragma unroll
  for(int q = 0; q < 3; q++)
  {
    float du[3];
#pragma unroll
    for(int i = 0; i < 3; i++) du[i] = ((i == q) ? inv_a : 0.f) - u[i] * da[q] * inv_a;
    (void)th;
    const float dK[9] = {0.f, -du[2], du[1], du[2], 0.f, -du[0], -du[1], du[0], 0.f};
    float udu = u[0] * du[0] + u[1] * du[1] + u[2] * du[2];
#pragma unroll
    for(int i = 0; i < 3; i++)
#pragma unroll
      for(int k = 0; k < 3; k++)
      {
        float K2 = u[i] * u[k] - ((i == k) ? uu : 0.f);
        float dK2 = du[i] * u[k] + u[i] * du[k] - ((i == k) ? 2.f * udu : 0.f);
        dR[q * 9 + i * 3 + k] = c * da[q] * K[i * 3 + k] + s * dK[i * 3 + k] + s * da[q] * K2 + oc * dK2;
      }
  }
}

