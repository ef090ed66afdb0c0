// Micro-benchmarks that sized the epilogue of the tcgen05 blend+skinning kernel (run on the B200 box):
//   (1) tcgen05.ld throughput per SM with 4 / 8 warps (x8 / x16 / x32 shapes)
//   (2) shared-memory wavefront cost of ld.shared.v4 with uniform / quarter-uniform / distinct addresses
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu && ./ubench
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../smplpp_b200/csrc/tc_ptx.cuh"
using namespace sb;

template<int X>
__device__ __forceinline__ void tld(uint32_t taddr, uint32_t & acc)
{
  if constexpr(X == 8)
  {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for(int i = 0; i < 8; i++) acc ^= r[i];
  }
  else if constexpr(X == 16)
  {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for(int i = 0; i < 16; i++) acc ^= r[i];
  }
  else
  {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for(int i = 0; i < 32; i++) acc ^= r[i];
  }
}

// two loads in flight before the wait (what a pipelined epilogue does)
template<int X>
__global__ void tmem_ld_bench(int iters, long long * cycles, uint32_t * sink)
{
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if(warp == 0) ptx::tmem_alloc<512>(&slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t base = slot + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for(int i = 0; i < iters; i++)
  {
#pragma unroll
    for(int c = 0; c < 512; c += X) tld<X>(base + c, acc);
  }
  __syncthreads();
  const long long t1 = clock64();
  if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  ptx::tc_fence_before();
  __syncthreads();
  if(warp == 0) ptx::tmem_dealloc<512>(slot);
}

// mode 0: all lanes same address; 1: each quarter-warp (8 lanes) its own address; 2: each lane distinct consecutive;
// 3: lanes pick among 4 addresses at random-ish (lane % 4)
__global__ void lds_bench(int mode, int iters, long long * cycles, float * sink)
{
  extern __shared__ float4 sm[];
  for(int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_float4(i, 1, 2, 3);
  __syncthreads();
  const int lane = threadIdx.x & 31;
  int idx = mode == 0 ? 0 : mode == 1 ? (lane >> 3) * 3 : mode == 2 ? lane : (lane & 3) * 3;
  const uint32_t a0 = ptx::smem_u32(sm) + idx * 16;
  float4 acc = make_float4(0, 0, 0, 0);
  __syncthreads();
  const long long t0 = clock64();
  for(int i = 0; i < iters; i++)
  {
#pragma unroll
    for(int u = 0; u < 16; u++)
    {
      const float4 v = ptx::lds128(a0 + ((i * 16 + u) & 127) * 512);
      acc.x += v.x, acc.y += v.y, acc.z += v.z, acc.w += v.w;
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  if(threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
}

int main()
{
  long long * cyc;
  uint32_t * sink;
  cudaMalloc(&cyc, 1024 * sizeof(long long));
  cudaMalloc(&sink, 148 * 1024 * sizeof(uint32_t));
  long long h[148];
  const int iters = 200;
  for(int warps : {4, 8})
  {
    for(int x : {8, 16, 32})
    {
      if(x == 8) tmem_ld_bench<8><<<148, warps * 32>>>(iters, cyc, sink);
      if(x == 16) tmem_ld_bench<16><<<148, warps * 32>>>(iters, cyc, sink);
      if(x == 32) tmem_ld_bench<32><<<148, warps * 32>>>(iters, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double bytes = double(iters) * warps * 32 * 512 * 4;
      printf("tcgen05.ld 32x32b.x%-2d  %d warps: %lld cycles, %.1f B/cycle/SM  (%s)\n", x, warps, h[0], bytes / h[0], cudaGetErrorString(e));
    }
  }
  cudaFuncSetAttribute(lds_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 8192);
  for(int warps : {4, 8, 16})
    for(int mode = 0; mode < 4; mode++)
    {
      lds_bench<<<148, warps * 32, 65536 + 8192>>>(mode, 500, cyc, reinterpret_cast<float *>(sink));
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      const double n = 500.0 * 16 * warps;
      printf("lds128 mode %d  %2d warps: %.2f cycles per warp-instruction (%s)\n", mode, warps, h[0] / n, cudaGetErrorString(e));
    }
  return 0;
}
