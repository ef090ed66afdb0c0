// Micro-benchmark: what one SM can ingest from L2 with cp.async.bulk (1-D TMA copies) as a function of the copy size
// and the bytes in flight, with 1 CTA and with one CTA per SM running at once.  Sized the stage ring of skin_tc3.cu.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_tma ubench_tma.cu && ./ubench_tma
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "../../smplpp_b200/csrc/tc_ptx.cuh"
using namespace sb;

// ring of `slots` slots of `chunk` bytes; one thread issues, waits slot by slot; total bytes per CTA = iters * chunk
__device__ __forceinline__ bool mbar_test_wait(uint64_t * bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred P;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(ptx::smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

template<bool kSpin>
__global__ void tma_ingest(const uint8_t * src, size_t src_bytes, int chunk, int slots, int iters, long long * cycles)
{
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[64];
  if(threadIdx.x == 0)
  {
    for(int s = 0; s < slots; s++) ptx::mbar_init(&bar[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    // every CTA walks its own window of the (L2-resident) source so that CTAs do not share lines
    size_t off = (static_cast<size_t>(blockIdx.x) * 1315423911u) % (src_bytes / 2);
    off &= ~size_t(1023);
    const long long t0 = clock64();
    int issued = 0;
    for(; issued < slots && issued < iters; issued++)
    {
      ptx::mbar_expect_tx(&bar[issued], chunk);
      ptx::bulk_load_1d(smem + issued * chunk, src + (off + static_cast<size_t>(issued) * chunk) % (src_bytes - chunk), chunk, &bar[issued]);
    }
    for(int i = 0; i < iters; i++)
    {
      const int s = i % slots;
      if constexpr(kSpin)
      {
        while(!mbar_test_wait(&bar[s], (i / slots) & 1)) {}
      }
      else
        ptx::mbar_wait(&bar[s], (i / slots) & 1);
      if(issued < iters)
      {
        ptx::mbar_expect_tx(&bar[s], chunk);
        ptx::bulk_load_1d(smem + s * chunk, src + (off + static_cast<size_t>(issued) * chunk) % (src_bytes - chunk), chunk, &bar[s]);
        issued++;
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

__device__ __forceinline__ void expect_tx_relaxed(uint64_t * bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.relaxed.cta.shared::cta.b64 _, [%0], %1;" ::"r"(ptx::smem_u32(bar)), "r"(bytes) : "memory");
}
// ring mode with per-iteration timestamps of CTA 0: when the wait returned, when expect_tx returned, when the copy was issued
template<bool kRelaxed>
__global__ void tma_trace(const uint8_t * src, int chunk, int slots, int iters, long long * ts)
{
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[64];
  if(threadIdx.x == 0)
  {
    for(int s = 0; s < slots; s++) ptx::mbar_init(&bar[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    const long long t0 = clock64();
    int issued = 0;
    for(; issued < slots; issued++)
    {
      ptx::mbar_expect_tx(&bar[issued], chunk);
      ptx::bulk_load_1d(smem + issued * chunk, src + static_cast<size_t>(issued) * chunk, chunk, &bar[issued]);
    }
    ts[0] = clock64() - t0;
    for(int i = 0; i < iters; i++)
    {
      const int s = i % slots;
      ptx::mbar_wait(&bar[s], (i / slots) & 1);
      const long long a = clock64();
      if constexpr(kRelaxed)
        expect_tx_relaxed(&bar[s], chunk);
      else
        ptx::mbar_expect_tx(&bar[s], chunk);
      const long long b = clock64();
      ptx::bulk_load_1d(smem + s * chunk, src + static_cast<size_t>(issued) * chunk, chunk, &bar[s]);
      issued++;
      const long long c = clock64();
      ts[1 + 3 * i] = a - t0, ts[2 + 3 * i] = b - t0, ts[3 + 3 * i] = c - t0;
    }
  }
}

// cost of a wait on an ALREADY COMPLETE barrier: back to back, and right after issuing a bulk copy to another slot
__global__ void wait_cost(const uint8_t * src, long long * ts)
{
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[16];
  if(threadIdx.x == 0)
  {
    for(int s = 0; s < 16; s++) ptx::mbar_init(&bar[s], 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    for(int s = 0; s < 8; s++) ptx::mbar_arrive(&bar[s]); // phase 0 of bars 0..7 complete
    long long t0 = clock64();
    while(clock64() - t0 < 3000) {}
    t0 = clock64();
    for(int s = 0; s < 8; s++) ptx::mbar_wait(&bar[s], 0);
    ts[0] = clock64() - t0; // 8 try_waits on complete barriers
    t0 = clock64();
    for(int s = 0; s < 8; s++) while(!mbar_test_wait(&bar[s], 0)) {}
    ts[1] = clock64() - t0; // 8 test_waits
    // copy to slot s (barrier 8 + s), then wait on complete barrier s
    for(int s = 0; s < 8; s++)
    {
      ptx::mbar_expect_tx(&bar[8 + s], 2048);
      ptx::bulk_load_1d(smem + s * 2048, src + s * 2048, 2048, &bar[8 + s]);
      const long long a = clock64();
      ptx::mbar_wait(&bar[s], 0);
      ts[2 + s] = clock64() - a;
    }
    for(int s = 0; s < 8; s++) ptx::mbar_wait(&bar[8 + s], 0);
    for(int s = 0; s < 8; s++)
    {
      ptx::mbar_expect_tx(&bar[8 + s], 2048);
      ptx::bulk_load_1d(smem + s * 2048, src + s * 2048, 2048, &bar[8 + s]);
      const long long a = clock64();
      while(!mbar_test_wait(&bar[s], 0)) {}
      ts[10 + s] = clock64() - a;
    }
    for(int s = 0; s < 8; s++) ptx::mbar_wait(&bar[8 + s], 1);
    // latency of one copy measured with a test_wait spin
    ptx::mbar_expect_tx(&bar[8], 2048);
    const long long a = clock64();
    ptx::bulk_load_1d(smem, src, 2048, &bar[8]);
    while(!mbar_test_wait(&bar[8], 0)) {}
    ts[18] = clock64() - a;
    // 8 copies issued back to back on 8 barriers, then completion time of each observed with test_wait spins
    for(int s = 0; s < 8; s++)
    {
      ptx::mbar_expect_tx(&bar[s], 2048); // phase 1 of bars 0..7
    }
    const long long b = clock64();
    for(int s = 0; s < 8; s++) ptx::bulk_load_1d(smem + s * 2048, src + 65536 + s * 2048, 2048, &bar[s]);
    for(int s = 0; s < 8; s++)
    {
      while(!mbar_test_wait(&bar[s], 1)) {}
      ts[19 + s] = clock64() - b;
    }
  }
}

// mode A: `n` copies of `chunk` bytes issued back to back by one thread onto ONE barrier, one wait per round
__global__ void tma_round(const uint8_t * src, size_t src_bytes, int chunk, int n, int rounds, long long * cycles)
{
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  if(threadIdx.x == 0)
  {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if(threadIdx.x == 0)
  {
    size_t off = (static_cast<size_t>(blockIdx.x) * 1315423911u) % (src_bytes / 2);
    off &= ~size_t(1023);
    const long long t0 = clock64();
    for(int r = 0; r < rounds; r++)
    {
      ptx::mbar_expect_tx(&bar, chunk * n);
      for(int i = 0; i < n; i++)
        ptx::bulk_load_1d(smem + i * chunk, src + off + (static_cast<size_t>(r) * n + i) * chunk, chunk, &bar);
      ptx::mbar_wait(&bar, r & 1);
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
}

int main()
{
  const size_t src_bytes = 32u << 20; // 32 MB: L2-resident after the first pass
  uint8_t * src;
  cudaMalloc(&src, src_bytes);
  cudaMemset(src, 1, src_bytes);
  long long * cyc;
  cudaMalloc(&cyc, 256 * sizeof(long long));
  cudaFuncSetAttribute(tma_ingest<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(tma_ingest<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(tma_round, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("spin | chunk KB | in flight KB | grid | B/cycle/SM (avg over CTAs) | chip B/cycle\n");
  for(int spin : {0, 1})
  for(int grid : {1, sms})
    for(int chunk : {2048, 4096, 8192, 16384, 49152})
      for(int inflight_kb : {48, 96, 192})
      {
        const int slots = inflight_kb * 1024 / chunk;
        if(slots < 1 || slots > 64) continue;
        const int iters = (4 << 20) / chunk; // 4 MB per CTA
        for(int rep = 0; rep < 2; rep++)
        {
          if(spin)
            tma_ingest<true><<<grid, 32, slots * chunk + 1024>>>(src, src_bytes, chunk, slots, iters, cyc);
          else
            tma_ingest<false><<<grid, 32, slots * chunk + 1024>>>(src, src_bytes, chunk, slots, iters, cyc);
        }
        if(cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for(long long c : h) avg += static_cast<double>(c);
        avg /= grid;
        const double bpc = static_cast<double>(iters) * chunk / avg;
        printf("%d | %8.0f | %12d | %4d | %10.1f | %10.0f\n", spin, chunk / 1024.0, slots * chunk / 1024, grid, bpc, bpc * grid);
      }
  {
    long long * ts;
    cudaMalloc(&ts, 4096 * sizeof(long long));
    cudaFuncSetAttribute(tma_trace<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(tma_trace<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    for(int chunk : {2048, 16384, -16384})
    {
      const bool relaxed = chunk < 0;
      if(relaxed) chunk = -chunk;
      const int slots = 8, iters = 24;
      for(int rep = 0; rep < 2; rep++)
      {
        if(relaxed)
          tma_trace<true><<<1, 32, slots * chunk + 1024>>>(src, chunk, slots, iters, ts);
        else
          tma_trace<false><<<1, 32, slots * chunk + 1024>>>(src, chunk, slots, iters, ts);
      }
      printf("relaxed expect_tx: %d\n", relaxed ? 1 : 0);
      cudaDeviceSynchronize();
      std::vector<long long> h(1 + 3 * iters);
      cudaMemcpy(h.data(), ts, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
      printf("trace chunk %d, %d slots: initial burst issued by +%lld; per iteration (wait done, expect_tx done, copy issued):\n", chunk, slots, h[0]);
      for(int i = 0; i < iters; i++) printf("  %2d: %6lld %6lld %6lld\n", i, h[1 + 3 * i], h[2 + 3 * i], h[3 + 3 * i]);
    }
  }
  {
    long long * ts;
    cudaMalloc(&ts, 64 * sizeof(long long));
    for(int rep = 0; rep < 2; rep++) wait_cost<<<1, 32, 32 * 1024>>>(src, ts);
    cudaDeviceSynchronize();
    long long h[27];
    cudaMemcpy(h, ts, sizeof(h), cudaMemcpyDeviceToHost);
    printf("wait cost: 8 try_wait on complete barriers %lld cycles | 8 test_wait %lld\n  try_wait right after a bulk copy issue:", h[0], h[1]);
    for(int i = 0; i < 8; i++) printf(" %lld", h[2 + i]);
    printf("\n  test_wait right after a bulk copy issue:");
    for(int i = 0; i < 8; i++) printf(" %lld", h[10 + i]);
    printf("\n  latency of one 2 KB copy (test_wait spin): %lld\n  8 copies back to back, completion seen at:", h[18]);
    for(int i = 0; i < 8; i++) printf(" %lld", h[19 + i]);
    printf("\n");
  }
  printf("mode A: n copies on one barrier | chunk KB | n | grid | cycles per round | B/cycle/SM\n");
  for(int grid : {1, sms})
    for(int chunk : {2048, 8192, 49152})
      for(int n : {1, 4, 16})
      {
        if(chunk * n > 192 * 1024) continue;
        const int rounds = 64;
        for(int rep = 0; rep < 2; rep++) tma_round<<<grid, 32, chunk * n + 1024>>>(src, src_bytes, chunk, n, rounds, cyc);
        if(cudaDeviceSynchronize() != cudaSuccess) { printf("error %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        std::vector<long long> h(grid);
        cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
        double avg = 0;
        for(long long c : h) avg += static_cast<double>(c);
        avg /= grid * rounds;
        printf("%8.0f | %3d | %4d | %8.0f | %8.1f\n", chunk / 1024.0, n, grid, avg, chunk * n / avg);
      }
  return 0;
}
