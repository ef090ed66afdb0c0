// Issue rates of the fp64 paths ik_solve_mma_kernel leans on, per SM sub-partition: mma.m8n8k4.f64 (DMMA), cvt.f64.f32
// (F2F), fma.f64 (DFMA), rsqrt.  One CTA per SM, W warps, every warp runs ITER rounds of U independent instructions.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fp64 ubench_fp64.cu && ./ubench_fp64
#include <cstdio>
#include <cuda_runtime.h>

template<int MODE, int U>
__global__ void k(int iters, double * out, long long * cyc)
{
  double a[U], c0[U], c1[U];
  float fsrc[U];
#pragma unroll
  for(int i = 0; i < U; i++) a[i] = 1.0 + threadIdx.x * 1e-3 + i, c0[i] = 0.0, c1[i] = 0.0, fsrc[i] = 1.f + i + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  for(int it = 0; it < iters; it++)
  {
#pragma unroll
    for(int i = 0; i < U; i++)
    {
      if(MODE == 0)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0[i]), "+d"(c1[i]) : "d"(a[i]), "d"(a[i]));
      else if(MODE == 1)
      {
        asm volatile("cvt.f64.f32 %0, %1;" : "=d"(c0[i]) : "f"(fsrc[i]));
        fsrc[i] += 1.f;
      }
      else if(MODE == 2)
        asm volatile("fma.rn.f64 %0, %1, %1, %0;" : "+d"(c0[i]) : "d"(a[i]));
      else
        c0[i] += rsqrt(a[i] + c0[i]);
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for(int i = 0; i < U; i++) s += c0[i] + c1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if(threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template<int MODE, int U>
void run(const char * name, int warps)
{
  double * out;
  long long * cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(double));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const int iters = 2000;
  k<MODE, U><<<148, warps * 32>>>(iters, out, cyc);
  k<MODE, U><<<148, warps * 32>>>(iters, out, cyc);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double per_smsp = static_cast<double>(iters) * U * warps / 4.0; // warp instructions per sub-partition
  printf("%-10s %2d warps/SM: %.2f cycles per warp instruction and sub-partition (%s)\n", name, warps, h[0] / per_smsp,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
  cudaFree(cyc);
}

int main()
{
  for(int w : {4, 8, 16})
  {
    run<0, 8>("DMMA", w);
    run<1, 8>("F2F", w);
    run<2, 8>("DFMA", w);
    run<3, 4>("rsqrt", w);
  }
  return 0;
}
