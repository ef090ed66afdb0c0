// Micro-benchmark: latency / throughput of tcgen05.mma kind::f16 (M = 128) as a function of N, for a chain of MMAs that
// accumulate into the SAME TMEM columns (dependent) versus round-robin over several accumulators (independent), with
// the A operand in shared memory (SS) or in TMEM (TS).  Sized the GEMM 2 sub-batches of skin_tc.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o ubench_mma ubench_mma.cu && ./ubench_mma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../smplpp_b200/csrc/tc_ptx.cuh"
using namespace sb;

// mode bit0: 1 = A from TMEM; naccum: accumulators walked round-robin (1 = fully dependent chain)
__global__ void mma_bench(int n, int ts, int naccum, int count, long long * cycles)
{
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  for(int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if(threadIdx.x == 0)
  {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if(threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if(threadIdx.x == 0)
  {
    const uint32_t idesc = ptx::make_idesc_f16(128, n);
    const uint32_t sa = ptx::smem_u32(smem);
    const uint64_t adesc = ptx::make_smem_desc<64>(sa);
    const uint64_t bdesc = ptx::make_smem_desc<64>(sa + 16384);
    const long long t0 = clock64();
    for(int i = 0; i < count; i++)
    {
      const uint32_t d = tmem + (naccum == 1 ? 0 : (naccum == 2 ? (i & 1) : (i % 3))) * n;
      if(ts)
        ptx::umma_f16_ts(d, tmem + 480, bdesc, idesc, 1u);
      else
        ptx::umma_f16_ss(d, adesc, bdesc, idesc, 1u);
    }
    ptx::tc_commit(&bar);
    const long long t1 = clock64();
    ptx::mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if(blockIdx.x == 0)
    {
      cycles[0] = t1 - t0;
      cycles[1] = t2 - t0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if(threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}


// variant 1: `if(elect_one())` around a straight-line block of 8 MMAs (what skin_tc.cu does)
// variant 2: no C++ branch at all: every lane runs the asm, the instruction is predicated on an elected lane inside the asm
__global__ void mma_bench2(int n, int variant, int count, long long * cycles)
{
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  for(int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if(threadIdx.x == 0)
  {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if(threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if(threadIdx.x < 32)
  {
    const uint32_t idesc = ptx::make_idesc_f16(128, n);
    const uint32_t sa = ptx::smem_u32(smem);
    const uint64_t adesc = ptx::make_smem_desc<64>(sa);
    const uint64_t bdesc = ptx::make_smem_desc<64>(sa + 16384);
    long long t0 = 0, t1 = 0;
    if(variant == 1)
    {
      if(ptx::elect_one())
      {
        t0 = clock64();
        for(int i = 0; i < count; i += 8)
        {
#pragma unroll
          for(int u = 0; u < 8; u++) ptx::umma_f16_ss(tmem + (u % 3) * n, adesc, bdesc, idesc, 1u);
        }
        ptx::tc_commit(&bar);
        t1 = clock64();
        ptx::mbar_wait(&bar, 0);
        cycles[0] = t1 - t0;
        cycles[1] = clock64() - t0;
      }
    }
    else
    {
      t0 = clock64();
      for(int i = 0; i < count; i += 8)
      {
#pragma unroll
        for(int u = 0; u < 8; u++)
        {
          const uint32_t d = tmem + (u % 3) * n;
          asm volatile(
              "{\n\t"
              ".reg .pred p, q;\n\t"
              "elect.sync _|q, 0xffffffff;\n\t"
              "setp.ne.b32 p, %4, 0;\n\t"
              "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
              "}\n" ::"r"(d),
              "l"(adesc), "l"(bdesc), "r"(idesc), "r"(1u)
              : "memory");
        }
      }
      if(ptx::elect_one())
      {
        ptx::tc_commit(&bar);
        t1 = clock64();
        ptx::mbar_wait(&bar, 0);
        cycles[0] = t1 - t0;
        cycles[1] = clock64() - t0;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if(threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}


__global__ void mma_bench3(int n, int ts, int count, long long * cycles)
{
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar[2];
  for(int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0;
  if(threadIdx.x == 0)
  {
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
  }
  if(threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if(threadIdx.x < 32 && ptx::elect_one())
  {
    const uint32_t idesc = ptx::make_idesc_f16(128, n);
    const uint32_t sa = ptx::smem_u32(smem);
    const long long t0 = clock64();
    int ph = 0;
    for(int i = 0; i < count; i += 6)
    {
      const int b = (i / 6) & 1;
#pragma unroll
      for(int prod = 0; prod < 3; prod++)
#pragma unroll
        for(int ks = 0; ks < 2; ks++)
        {
          const uint64_t bdesc = ptx::make_smem_desc<64>(sa + 16384 + (prod == 2 ? 16384 : 0) + ks * 32);
          if(ts)
            ptx::umma_f16_ts(tmem + b * n, tmem + 480 + (prod == 1 ? 16 : 0) + ks * 8, bdesc, idesc, (prod | ks) ? 1u : 0u);
          else
            ptx::umma_f16_ss(tmem + b * n, ptx::make_smem_desc<64>(sa + (prod == 1 ? 8192 : 0) + ks * 32), bdesc, idesc, (prod | ks) ? 1u : 0u);
        }
      ptx::tc_commit(&bar[b]);
    }
    const long long t1 = clock64();
    ptx::tc_commit(&bar[0]);
    cycles[0] = t1 - t0;
    (void)ph;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if(threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}


// chain of 6 dependent TS MMAs per buffer while `ldwarps` other warps hammer TMEM with tcgen05.ld (epilogue traffic)
__global__ void mma_bench4(int n, int ldwarps, int count, long long * cycles, uint32_t * sink)
{
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar[2];
  __shared__ volatile int done;
  for(int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(smem)[i] = 0x3c003c00u;
  if(threadIdx.x == 0)
  {
    ptx::mbar_init(&bar[0], 1);
    ptx::mbar_init(&bar[1], 1);
    ptx::fence_barrier_init();
    done = 0;
  }
  if(threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  if(warp == 0)
  {
    if(ptx::elect_one())
    {
      const uint32_t idesc = ptx::make_idesc_f16(128, n);
      const uint32_t sa = ptx::smem_u32(smem);
      const long long t0 = clock64();
      for(int i = 0; i < count; i += 6)
      {
        const int b = (i / 6) & 1;
#pragma unroll
        for(int prod = 0; prod < 3; prod++)
#pragma unroll
          for(int ks = 0; ks < 2; ks++)
          {
            const uint64_t bdesc = ptx::make_smem_desc<64>(sa + 16384 + (prod == 2 ? 16384 : 0) + ks * 32);
            ptx::umma_f16_ts(tmem + 288 + b * n, tmem + 480 + (prod == 1 ? 16 : 0) + ks * 8, bdesc, idesc, (prod | ks) ? 1u : 0u);
          }
        ptx::tc_commit(&bar[b]);
      }
      const long long t1 = clock64();
      cycles[0] = t1 - t0;
      done = 1;
    }
  }
  else if(warp <= ldwarps)
  {
    const uint32_t base = tmem + (static_cast<uint32_t>((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    while(!done)
    {
#pragma unroll
      for(int c = 0; c < 288; c += 16)
      {
        float v[16];
        ptx::tmem_ld_x16(base + c, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for(int i = 0; i < 16; i++) acc ^= __float_as_uint(v[i]);
      }
    }
    sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if(threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

int main()
{
  long long * cyc;
  cudaMalloc(&cyc, 16);
  cudaFuncSetAttribute(mma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int count = 96;
  for(int ts = 0; ts < 2; ts++)
    for(int naccum : {1, 2, 3})
      for(int n : {16, 48, 96, 128, 240})
      {
        if(naccum * n > 480) continue;
        long long h[2];
        for(int rep = 0; rep < 3; rep++) mma_bench<<<148, 128, 64 * 1024>>>(n, ts, naccum, count, cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
        printf("%s naccum %d N %3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (%s)\n", ts ? "TS" : "SS", naccum, n,
               double(h[0]) / count, double(h[1]) / count, cudaGetErrorString(e));
      }
  cudaFuncSetAttribute(mma_bench2, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for(int variant : {1, 2})
    for(int n : {16, 48, 128, 160})
    {
      long long h[2];
      for(int rep = 0; rep < 3; rep++) mma_bench2<<<148, 128, 64 * 1024>>>(n, variant, 96, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      printf("variant %d (unrolled x8, 3 accumulators) N %3d: issue %6.1f cyc/mma, complete %6.1f cyc/mma (%s)\n", variant, n,
             double(h[0]) / 96, double(h[1]) / 96, cudaGetErrorString(e));
    }
  cudaFuncSetAttribute(mma_bench3, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for(int ts : {0, 1})
    for(int n : {48, 96, 240})
    {
      long long h[2];
      for(int rep = 0; rep < 3; rep++) mma_bench3<<<148, 128, 64 * 1024>>>(n, ts, 96, cyc);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      printf("chain of 6 per buffer (%s) N %3d: issue %6.1f cyc/mma (%s)\n", ts ? "TS" : "SS", n, double(h[0]) / 96, cudaGetErrorString(e));
    }
  cudaFuncSetAttribute(mma_bench4, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  uint32_t * sink;
  cudaMalloc(&sink, 148 * 1024 * 4);
  for(int ldwarps : {0, 4, 8, 16})
    for(int n : {48, 96})
    {
      long long h[2];
      for(int rep = 0; rep < 3; rep++) mma_bench4<<<148, 32 * 17, 64 * 1024>>>(n, ldwarps, 192, cyc, sink);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(h, cyc, 16, cudaMemcpyDeviceToHost);
      printf("TS chain of 6, %2d warps of concurrent tcgen05.ld, N %3d: %6.1f cyc/mma (%s)\n", ldwarps, n, double(h[0]) / 192, cudaGetErrorString(e));
    }
  return 0;
}
