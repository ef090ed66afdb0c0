// Sweep-grid occupancy of the posed mesh (node/node.cpp:1023-1073): the 2.5 cm voxel grid around the mesh
// of one frame (toolbox/GridUtils.hpp:28-63: index = floor / ceil of position / GRID_SCALE, position = GRID_SCALE * index)
// and, per grid point, the generalized winding number of the triangle mesh (igl::winding_number, un-vendored: the sum of
// the signed solid angles of all faces over 4 pi, Van Oosterom & Strackee); cells with a winding number above 0.5 are
// inside the body (node.cpp:1054-1059).
//
// One thread per grid point, the faces of the frame gathered once into (F, 9) floats and streamed through shared memory
// 256 at a time (every thread of the block reads the same face: broadcast loads); the sum runs in double.
#include "common.cuh"

using namespace sb;

namespace
{
constexpr double kGridScale = 0.025; // smplpp::GRID_SCALE
constexpr int TPB = 256;

__global__ void bounds_kernel(int V, const float * __restrict__ verts, float * __restrict__ out) // out: min xyz | max xyz
{
  __shared__ float s_lo[3][TPB / 32], s_hi[3][TPB / 32];
  float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
  for(int v = threadIdx.x; v < V; v += TPB)
#pragma unroll
    for(int a = 0; a < 3; a++)
    {
      const float x = verts[3 * v + a];
      lo[a] = fminf(lo[a], x), hi[a] = fmaxf(hi[a], x);
    }
#pragma unroll
  for(int a = 0; a < 3; a++)
  {
    for(int o = 16; o > 0; o >>= 1)
    {
      lo[a] = fminf(lo[a], __shfl_xor_sync(0xffffffffu, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xffffffffu, hi[a], o));
    }
    if((threadIdx.x & 31) == 0) s_lo[a][threadIdx.x >> 5] = lo[a], s_hi[a][threadIdx.x >> 5] = hi[a];
  }
  __syncthreads();
  if(threadIdx.x < 3)
  {
    float l = INFINITY, h = -INFINITY;
    for(int w = 0; w < TPB / 32; w++) l = fminf(l, s_lo[threadIdx.x][w]), h = fmaxf(h, s_hi[threadIdx.x][w]);
    out[threadIdx.x] = l, out[3 + threadIdx.x] = h;
  }
}

__global__ void gather_faces_kernel(int F, const int32_t * __restrict__ faces, const float * __restrict__ verts,
                                    float * __restrict__ tri)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= 9 * F) return;
  const int f = i / 9, e = i - 9 * f;
  tri[i] = verts[3 * faces[3 * f + e / 3] + e % 3];
}

// grid point g -> index (ix, iy, iz) in the reference's loop order (x outermost, z innermost, node.cpp:1038-1048)
__global__ void __launch_bounds__(TPB) winding_kernel(int F, const float * __restrict__ tri, int3 idx_min, int3 num, long long total,
                                                      float * __restrict__ winding, uint8_t * __restrict__ occupied)
{
  __shared__ float s_tri[TPB * 9];
  const long long g = static_cast<long long>(blockIdx.x) * TPB + threadIdx.x;
  const long long gc = g < total ? g : total - 1;
  const int iz = static_cast<int>(gc % num.z), iy = static_cast<int>((gc / num.z) % num.y), ix = static_cast<int>(gc / (static_cast<long long>(num.z) * num.y));
  // Eigen: GRID_SCALE (double) * int cast to float -> float product of float(0.025) and the index
  const float gs = static_cast<float>(kGridScale);
  const float px = gs * static_cast<float>(idx_min.x + ix), py = gs * static_cast<float>(idx_min.y + iy),
              pz = gs * static_cast<float>(idx_min.z + iz);
  double sum = 0.0;
  for(int f0 = 0; f0 < F; f0 += TPB)
  {
    const int nf = min(TPB, F - f0);
    __syncthreads();
    for(int i = threadIdx.x; i < 9 * nf; i += TPB) s_tri[i] = tri[static_cast<size_t>(9) * f0 + i];
    __syncthreads();
    float part = 0.f;
    for(int f = 0; f < nf; f++)
    {
      const float * t = s_tri + 9 * f;
      const float ax = t[0] - px, ay = t[1] - py, az = t[2] - pz;
      const float bx = t[3] - px, by = t[4] - py, bz = t[5] - pz;
      const float cx = t[6] - px, cy = t[7] - py, cz = t[8] - pz;
      const float la = sqrtf(ax * ax + ay * ay + az * az), lb = sqrtf(bx * bx + by * by + bz * bz),
                  lc = sqrtf(cx * cx + cy * cy + cz * cz);
      const float det = ax * (by * cz - bz * cy) - ay * (bx * cz - bz * cx) + az * (bx * cy - by * cx);
      const float den = la * lb * lc + (ax * bx + ay * by + az * bz) * lc + (bx * cx + by * cy + bz * cz) * la
                        + (cx * ax + cy * ay + cz * az) * lb;
      part += atan2f(det, den); // half the signed solid angle
    }
    sum += static_cast<double>(part);
  }
  if(g < total)
  {
    const float w = static_cast<float>(sum * (2.0 / (4.0 * 3.14159265358979323846)));
    if(winding) winding[g] = w;
    if(occupied) occupied[g] = w > 0.5f ? 1 : 0;
  }
}
} // namespace

extern "C" int smplpp_sweep_grid_bounds(const smplpp_model_t * model, void * stream, const float * vertices_dev,
                                        int32_t * grid_idx_min, int32_t * grid_num)
{
  if(!model || !vertices_dev || !grid_idx_min || !grid_num) return fail(SMPLPP_ERR_INVALID, "SMPL", "invalid sweep grid arguments!");
  cudaStream_t st = as_stream(stream);
  float * d_box = nullptr;
  SB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&d_box), 6 * sizeof(float), st));
  bounds_kernel<<<1, TPB, 0, st>>>(model->d.V, vertices_dev, d_box);
  SB_LAUNCHED();
  float box[6];
  SB_CUDA(cudaMemcpyAsync(box, d_box, sizeof(box), cudaMemcpyDeviceToHost, st));
  SB_CUDA(cudaFreeAsync(d_box, st));
  SB_CUDA(cudaStreamSynchronize(st));
  for(int a = 0; a < 3; a++)
  {
    if(!std::isfinite(box[a]) || !std::isfinite(box[3 + a])) return fail(SMPLPP_ERR_INVALID, "SMPL", "mesh is not finite!");
    // getGridIdxFloor / getGridIdxCeil<float> (GridUtils.hpp:49-63): float position / double GRID_SCALE evaluated by Eigen
    // in float (the scalar is cast to the matrix scalar type)
    const float q_lo = box[a] / static_cast<float>(kGridScale), q_hi = box[3 + a] / static_cast<float>(kGridScale);
    const int lo = static_cast<int>(std::floor(q_lo)), hi = static_cast<int>(std::ceil(q_hi));
    grid_idx_min[a] = lo;
    grid_num[a] = hi - lo + 1;
  }
  return SMPLPP_OK;
}

extern "C" int smplpp_sweep_grid_winding(const smplpp_model_t * model, void * stream, const float * vertices_dev,
                                         const int32_t * grid_idx_min, const int32_t * grid_num, float * winding_dev,
                                         uint8_t * occupied_dev)
{
  if(!model || !vertices_dev || !grid_idx_min || !grid_num || (!winding_dev && !occupied_dev))
    return fail(SMPLPP_ERR_INVALID, "SMPL", "invalid sweep grid arguments!");
  if(grid_num[0] < 1 || grid_num[1] < 1 || grid_num[2] < 1) return fail(SMPLPP_ERR_INVALID, "SMPL", "empty sweep grid!");
  const ModelDev & d = model->d;
  if(d.F < 1) return fail(SMPLPP_ERR_INVALID, "SMPL", "the model has no faces!");
  const long long total = static_cast<long long>(grid_num[0]) * grid_num[1] * grid_num[2];
  cudaStream_t st = as_stream(stream);
  float * tri = nullptr;
  SB_CUDA(cudaMallocAsync(reinterpret_cast<void **>(&tri), static_cast<size_t>(9) * d.F * sizeof(float), st));
  gather_faces_kernel<<<(9 * d.F + 255) / 256, 256, 0, st>>>(d.F, d.faces, vertices_dev, tri);
  SB_LAUNCHED();
  winding_kernel<<<static_cast<unsigned>((total + TPB - 1) / TPB), TPB, 0, st>>>(
      d.F, tri, make_int3(grid_idx_min[0], grid_idx_min[1], grid_idx_min[2]), make_int3(grid_num[0], grid_num[1], grid_num[2]), total,
      winding_dev, occupied_dev);
  SB_LAUNCHED();
  SB_CUDA(cudaFreeAsync(tri, st));
  return SMPLPP_OK;
}
