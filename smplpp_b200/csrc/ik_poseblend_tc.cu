// C1b: the pose-blend (and shape-blend) columns of the IK Jacobian on tcgen05.
//
// Reference: the node gets these entries from autograd through BlendShape (src/BlendShape.cpp:670-683, 803-928) one
// residual row at a time (node/node.cpp:790-877).  Analytically, for residual row r of task m,
//
//   J[r, theta_kc] += sum_e Q_m[r, 9 (k - 1) + e] . d vec(R_k)[e] / d theta_kc,     Q_m = CA_m . P_m
//
// where P_m (3 np x 224) stacks the x / y / z basis rows of the np vertices the task depends on (corners + 1-rings) and
// CA_m (ROWS x 3 np) = d(residual rows) / d(rest vertex) is per frame.  P_m is the SAME for every frame of a task set:
// ik_jacobian_kernel (one CTA per frame) re-read those 1.3 MB of basis rows from L2 for every frame (18.8 GB per 16384
// frames, the phase ran at the L2 -> SM bandwidth: 2.5 of the kernel's 5.2 ms).  Here 32 frames x 4 row slots form one
// M = 128 tile, P_m^T is a per-task-set fp16 hi | lo stage image (K-major, SWIZZLE_64B, K-blocks of 32) fetched with one
// cp.async.bulk per K-block and CTA, CA_m of the 32 frames is converted to the same format by builder warps, and
// Q_m (128 x 224, fp32) accumulates in TMEM (two buffers: the epilogue of task m runs under the MMAs of task m + 1).
// Split precision as in skin_tc3.cu / vposer_tc.cu: hi.hi + lo.hi + hi.lo, basis pre-scaled by a power of two.
//
// warp 0: TMA producer of the P_m ring | warp 1: TMEM allocator + MMA issuer | warps 2-3: idle |
// warps 4-11: builders of the CA_m operand (two groups, alternating K-blocks) |
// warps 12-19: epilogue (two groups: basis columns 0..107 | 108..223), thread = TMEM lane = (frame, row slot):
// contraction with d vec(R_k) / d theta_k from shared memory, result added to the J rows that ik_jacobian_kernel left
// holding the kinematic-chain part (red.global.add: one addend per element, so the sum is order-independent).
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "ik_poseblend.cuh"
#include "skin_common.cuh"
#include "tasks.cuh"
#include "tc_ptx.cuh"

using namespace sb;

namespace pbtc
{
constexpr int FPB = 32;                  // frames per CTA
constexpr int ROWB = 64;                 // 32 fp16 = one SWIZZLE_64B span
constexpr int NCOL = kBlendK;            // 224 basis columns = UMMA N
constexpr int B_PART = NCOL * ROWB;      // 14336
constexpr int A_PART = 128 * ROWB;       // 8192
constexpr int SLOT = 2 * B_PART + 2 * A_PART; // 45056: [P hi | P lo | CA hi | CA lo] of one K-block
constexpr int SLOTS = 3;
constexpr int DRJ = 28;                  // per joint: 27 entries d vec(R_k)[e] / d theta_kc at [9 c + e], padded to 7 float4
constexpr int DR = DRJ * (kJoints - 1);  // 644 floats per frame (4 mod 32: the 8 frames of a warp read 16 bytes each from
                                         // different banks)
constexpr int OFF_DR = SLOTS * SLOT;
constexpr int OFF_BAR = OFF_DR + FPB * DR * 4;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 256;
constexpr int CTRL_WARPS = 4, BUILD_WARPS = 8, EPI_WARPS = 8;
constexpr int THREADS = 32 * (CTRL_WARPS + BUILD_WARPS + EPI_WARPS);
constexpr int SPLIT_COL = 108;           // epilogue group 0: joints 1..12, group 1: joints 13..23 and the shape columns
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(B_PART % 1024 == 0 && A_PART % 1024 == 0, "swizzle atoms stay aligned");

__host__ __device__ constexpr uint32_t swz64(uint32_t o)
{
  return o ^ (((o >> 7) & 3u) << 4);
}

struct Params
{
  int B, n, rows, use_ring;
  int beta_col;                // first beta column of J (75 + phi columns), < 0: no beta columns
  int ca_stride;               // floats per frame of ca
  const int * slot_off;        // (n + 1) first K-block of every task in the image / in ca (x 128 floats)
  const int * pair_off;        // (n + 1)
  const uint8_t * img;         // [K-block][hi | lo][224][64 B]
  const float * ca;            // (B, ca_stride): per task [4 row slots][32 * K-blocks]
  const float * dr;            // (B, 23, 28)
  float out_scale;             // 2^-basis_exp
  float * J;                   // (B, 4 n, ld)
  int ld;
};

__device__ __forceinline__ void red_add_v2(float * addr, float a, float b)
{
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float * addr, float a, float b, float c, float d)
{
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template<int G>
__device__ __forceinline__ void epilogue_task(const Params & p, uint32_t taddr, uint32_t dr, float * jrow, bool live)
{
  // G = 0: columns [0, 108) = joints 1..12; G = 1: columns [108, 224) = joints 13..23, then 10 shape columns.
  // dr: SHARED-space address of this frame's derivative table (explicit ld.shared: through a generic pointer the compiler
  // emitted LD.E + R2UR pairs, 27 % of the kernel's stall samples; seven 16-byte loads per joint: with one 4-byte load per
  // entry the epilogue ran at the issue rate of the shared-memory pipe)
  constexpr int C0 = G == 0 ? 0 : 96, C1 = G == 0 ? 112 : 224; // chunks of 16 columns that cover the range
  constexpr int K0 = G == 0 ? 1 : 13, NK = G == 0 ? 12 : 11;
  float acc[NK][3];
#pragma unroll
  for(int k = 0; k < NK; k++) acc[k][0] = acc[k][1] = acc[k][2] = 0.f;
  float shp[kShapeDim];
  float jd[DRJ]; // derivative entries of the joint whose columns are being read
#pragma unroll
  for(int ch = C0; ch < C1; ch += 16)
  {
    float v[16];
    ptx::tmem_ld_x16(taddr + ch, v);
    ptx::tmem_ld_wait();
#pragma unroll
    for(int i = 0; i < 16; i++)
    {
      const int d = ch + i;
      if(d < kPoseDim)
      {
        const int k = d / 9 + 1, e = d % 9;
        if(k >= K0 && k < K0 + NK)
        {
          if(e == 0)
          {
#pragma unroll
            for(int q = 0; q < DRJ / 4; q++)
            {
              const float4 t4 = ptx::lds128(dr + 4 * (DRJ * (k - 1) + 4 * q));
              jd[4 * q] = t4.x, jd[4 * q + 1] = t4.y, jd[4 * q + 2] = t4.z, jd[4 * q + 3] = t4.w;
            }
          }
#pragma unroll
          for(int c = 0; c < 3; c++) acc[k - K0][c] = fmaf(v[i], jd[9 * c + e], acc[k - K0][c]);
        }
      }
      else if(G == 1 && d < kPoseDim + kShapeDim)
        shp[d - kPoseDim] = v[i];
    }
  }
  if(!live) return;
  // J row + 3 + 3 K0 .. : 36 (G = 0, columns 6..41) or 33 (G = 1, columns 42..74) consecutive floats of a 16-byte aligned
  // row (ld % 4 == 0): two leading floats, then 16-byte vector reductions (one L2 transaction per 4 values instead of 4)
  float o[3 * NK + 3];
#pragma unroll
  for(int k = 0; k < NK; k++)
#pragma unroll
    for(int c = 0; c < 3; c++) o[3 * k + c] = acc[k][c] * p.out_scale;
  float * dst = jrow + 3 + 3 * K0; // column 6 or 42: 8 bytes past a 16-byte boundary
  red_add_v2(dst, o[0], o[1]);
  constexpr int NV4 = (3 * NK - 2) / 4; // G = 0: 8 (columns 8..39), G = 1: 7 (columns 44..71)
#pragma unroll
  for(int i = 0; i < NV4; i++) red_add_v4(dst + 2 + 4 * i, o[2 + 4 * i], o[3 + 4 * i], o[4 + 4 * i], o[5 + 4 * i]);
  if(G == 0)
    red_add_v2(dst + 34, o[34], o[35]);
  else
  {
    red_add_v2(dst + 30, o[30], o[31]);
    atomicAdd(dst + 32, o[32]);
  }
  if(G == 1 && p.beta_col >= 0)
  {
#pragma unroll
    for(int i = 0; i < kShapeDim; i++) atomicAdd(jrow + p.beta_col + i, shp[i] * p.out_scale);
  }
}
} // namespace pbtc

__global__ void __launch_bounds__(pbtc::THREADS, 1) ik_poseblend_tc_kernel(const pbtc::Params p)
{
  using namespace pbtc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float * s_dr = reinterpret_cast<float *>(smem + OFF_DR);
  uint64_t * bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  uint64_t * full_b = bars;               // [SLOTS] TMA -> MMA
  uint64_t * full_a = full_b + SLOTS;     // [SLOTS] builders -> MMA
  uint64_t * empty = full_a + SLOTS;      // [SLOTS] MMA -> TMA + builders
  uint64_t * acc_full = empty + SLOTS;    // [2] MMA -> epilogue
  uint64_t * acc_empty = acc_full + 2;    // [2] epilogue -> MMA
  uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f0 = static_cast<int>(blockIdx.x) * FPB;

  if(warp == 0 && lane == 0)
  {
    for(int s = 0; s < SLOTS; s++)
    {
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&full_a[s], BUILD_WARPS / 2);
      ptx::mbar_init(&empty[s], 1);
    }
    for(int b = 0; b < 2; b++)
    {
      ptx::mbar_init(&acc_full[b], 1);
      ptx::mbar_init(&acc_empty[b], EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if(warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  // d vec(R_k) / d theta_k of this CTA's frames
  for(int i = threadIdx.x; i < FPB * DR; i += THREADS)
  {
    const int fl = i / DR;
    s_dr[i] = f0 + fl < p.B ? __ldg(p.dr + static_cast<size_t>(f0) * DR + i) : 0.f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // K-blocks of task m that carry data: all of them with the 1-rings, the first one (3 corners = 9 values) without
  auto task_kb = [&](int m) { return p.use_ring ? __ldg(p.slot_off + m + 1) - __ldg(p.slot_off + m) : 1; };

  // registers: 96 per thread at launch (640 threads); the control warps and the builders hand some back, the epilogue
  // warps (36 accumulators + 28 derivative entries + 16 TMEM words live) take 128
  // (setmaxnreg inside the role branches: ptxas only honours it there)
  if(warp < CTRL_WARPS)
  {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
  if(warp == 0)
  {
    if(ptx::elect_one())
    {
      int it = 0;
      for(int m = 0; m < p.n; m++)
      {
        const int s0 = __ldg(p.slot_off + m), nkb = task_kb(m);
        for(int kb = 0; kb < nkb; kb++, it++)
        {
          const int s = it % SLOTS;
          ptx::mbar_wait(&empty[s], ((it / SLOTS) & 1) ^ 1);
          ptx::mbar_expect_tx(&full_b[s], 2 * B_PART);
          ptx::bulk_load_1d(smem + s * SLOT, p.img + static_cast<size_t>(s0 + kb) * (2 * B_PART), 2 * B_PART, &full_b[s]);
        }
      }
    }
  }
  else if(warp == 1)
  {
    if(ptx::elect_one())
    {
      constexpr uint32_t DHI = ptx::smem_desc_hi<ROWB>();
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, NCOL);
      const uint32_t smem16 = ptx::smem_u32(smem) >> 4;
      int it = 0;
      for(int m = 0; m < p.n; m++)
      {
        const int b = m & 1;
        if(m >= 2)
        {
          ptx::mbar_wait(&acc_empty[b], ((m >> 1) - 1) & 1);
          ptx::tc_fence_after();
        }
        const int nkb = task_kb(m);
        const int np = p.use_ring ? __ldg(p.pair_off + m + 1) - __ldg(p.pair_off + m) : 3;
        const int ksteps = (3 * np + 15) >> 4;
        for(int kb = 0; kb < nkb; kb++, it++)
        {
          const int s = it % SLOTS;
          ptx::mbar_wait(&full_b[s], (it / SLOTS) & 1);
          ptx::mbar_wait(&full_a[s], (it / SLOTS) & 1);
          ptx::tc_fence_after();
          const uint32_t st16 = smem16 + s * (SLOT >> 4);
          const int nks = min(2, ksteps - 2 * kb);
#pragma unroll
          for(int prod = 0; prod < 3; prod++)
          {
            const int pa = prod == 1 ? 1 : 0, pb = prod == 2 ? 1 : 0; // hi.hi, lo.hi, hi.lo
            for(int ks = 0; ks < nks; ks++)
              ptx::umma_f16_ss_lo(tmem_base + b * 256, st16 + ((2 * B_PART + pa * A_PART + ks * 32) >> 4),
                                  st16 + ((pb * B_PART + ks * 32) >> 4), DHI, idesc, (kb | prod | ks) != 0 ? 1u : 0u);
          }
          ptx::tc_commit(&empty[s]);
        }
        ptx::tc_commit(&acc_full[b]);
      }
    }
  }
  } // control warps
  else if(warp < CTRL_WARPS + BUILD_WARPS)
  {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 88;");
    // ---- CA_m K-blocks: thread = row (frame fl, row slot r); group gb builds every other K-block ----
    const int bw = warp - CTRL_WARPS, gb = bw >> 2;
    const int L = (bw & 3) * 32 + lane, fl = L >> 2, r = L & 3;
    const bool live = f0 + fl < p.B && r < p.rows;
    const float * ca_f = p.ca + static_cast<size_t>(f0 + fl) * p.ca_stride;
    // K-block `it` of the whole sequence -> (task, K-block in task): this group's K-blocks are it = gb, gb + 2, ...; the
    // loads of the next one are in flight while the current one is converted and stored
    int m_nx = 0, kb_nx = 0, it_nx = 0;
    auto advance = [&](int steps) { // move the cursor `steps` K-blocks on; m_nx = p.n at the end
      while(steps > 0 && m_nx < p.n)
      {
        const int nkb = task_kb(m_nx);
        if(kb_nx + steps < nkb)
        {
          kb_nx += steps, it_nx += steps, steps = 0;
        }
        else
        {
          const int adv = nkb - kb_nx;
          steps -= adv, it_nx += adv, kb_nx = 0, m_nx++;
        }
      }
    };
    float4 nxt[8];
    auto prefetch = [&]() {
      if(m_nx >= p.n) return;
      const int s0 = __ldg(p.slot_off + m_nx), nkb_all = __ldg(p.slot_off + m_nx + 1) - s0;
      const int np = p.use_ring ? __ldg(p.pair_off + m_nx + 1) - __ldg(p.pair_off + m_nx) : 3;
      const int kvalid = 3 * np;
      const float * row = ca_f + static_cast<size_t>(s0) * 128 + r * (32 * nkb_all);
#pragma unroll
      for(int c = 0; c < 8; c++)
      {
        const int k = 32 * kb_nx + 4 * c;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if(live && k < kvalid) v = __ldg(reinterpret_cast<const float4 *>(row + k));
        if(k + 1 >= kvalid) v.y = 0.f;
        if(k + 2 >= kvalid) v.z = 0.f;
        if(k + 3 >= kvalid) v.w = 0.f;
        nxt[c] = v;
      }
    };
    advance(gb);
    prefetch();
    while(m_nx < p.n)
    {
      const int it = it_nx, s = it % SLOTS;
      float x[32];
#pragma unroll
      for(int c = 0; c < 8; c++) x[4 * c] = nxt[c].x, x[4 * c + 1] = nxt[c].y, x[4 * c + 2] = nxt[c].z, x[4 * c + 3] = nxt[c].w;
      advance(2);
      prefetch();
      uint4 hi[4], lo[4];
#pragma unroll
      for(int c = 0; c < 4; c++)
      {
        uint32_t h[4], l[4];
#pragma unroll
        for(int e = 0; e < 4; e++)
        {
          const float a = fminf(fmaxf(x[8 * c + 2 * e], -6.0e4f), 6.0e4f), b = fminf(fmaxf(x[8 * c + 2 * e + 1], -6.0e4f), 6.0e4f);
          const float ah = __half2float(__float2half_rn(a)), bh = __half2float(__float2half_rn(b));
          h[e] = skin::pack_half2(ah, bh);
          l[e] = skin::pack_half2(a - ah, b - bh);
        }
        hi[c] = make_uint4(h[0], h[1], h[2], h[3]);
        lo[c] = make_uint4(l[0], l[1], l[2], l[3]);
      }
      ptx::mbar_wait(&empty[s], ((it / SLOTS) & 1) ^ 1);
      uint8_t * dst = smem + s * SLOT + 2 * B_PART;
#pragma unroll
      for(int c = 0; c < 4; c++)
      {
        const uint32_t o = static_cast<uint32_t>(L * ROWB + c * 16);
        *reinterpret_cast<uint4 *>(dst + swz64(o)) = hi[c];
        *reinterpret_cast<uint4 *>(dst + swz64(A_PART + o)) = lo[c];
      }
      ptx::fence_proxy_async(); // generic-proxy stores -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if(lane == 0) ptx::mbar_arrive(&full_a[s]);
    }
  }
  else
  {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 128;");
    // ---- epilogue: thread = TMEM lane L = 4 fl + r ----
    const int ew = warp - CTRL_WARPS - BUILD_WARPS, ge = ew >> 2, q = warp & 3;
    const int L = q * 32 + lane, fl = L >> 2, r = L & 3;
    const bool live = f0 + fl < p.B && r < p.rows;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t dr = ptx::smem_u32(s_dr + fl * DR);
    float * jf = p.J + (static_cast<size_t>(f0 + fl) * 4 * p.n + r) * p.ld;
    for(int m = 0; m < p.n; m++)
    {
      const int b = m & 1;
      ptx::mbar_wait(&acc_full[b], (m >> 1) & 1);
      ptx::tc_fence_after();
      float * jrow = jf + static_cast<size_t>(4 * m) * p.ld;
      // the accumulators are read before the buffer is handed back; the additions to J follow
      if(ge == 0)
        epilogue_task<0>(p, lane_taddr + b * 256, dr, jrow, live);
      else
        epilogue_task<1>(p, lane_taddr + b * 256, dr, jrow, live);
      ptx::tc_fence_before();
      __syncwarp();
      if(lane == 0) ptx::mbar_arrive(&acc_empty[b]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if(warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}


// ------------------------------------------------------------------------------------------------------------
// Rest shape of the task vertices (BlendShape + the first half of JointRegression on the ~480 rows the tasks touch,
// src/BlendShape.cpp:670-683, src/JointRegression.cpp:551-565) on tcgen05: rest[f, 3 u + a] = T + sum_c basis[3 u + a, c] coef[f, c].
// The FFMA kernel of the forward path did this in 0.33 ms per 16384 frames (43 % of the FMA peak); here 128 frames are one
// M = 128 tile whose coefficient rows (K = 224, fp16 hi | lo) are converted ONCE per CTA and stay in shared memory, the
// basis rows of 224 coordinates at a time are per-task-set stage images (the array is K-major already) streamed through a
// three-slot ring, and the epilogue (thread = frame) adds the template in fp32 and transposes 32 x 32 blocks through shared
// memory so that the (B, 3 nU) result goes out in 128-byte row pieces.
// ------------------------------------------------------------------------------------------------------------
namespace rstc
{
using pbtc::A_PART;
using pbtc::B_PART;
using pbtc::NCOL;
using pbtc::ROWB;
using pbtc::swz64;
constexpr int FPB = 128;                  // frames per CTA
constexpr int NKB = kBlendK / 32;         // 7 K-blocks
constexpr int A_BYTES = NKB * 2 * A_PART; // 114688: coefficient rows, all K-blocks, hi | lo
constexpr int SLOTS = 3;
constexpr int SLOT = 2 * B_PART;          // 28672: one (coordinate tile, K-block) of the basis, hi | lo
constexpr int OFF_RING = A_BYTES;
constexpr int OFF_TMPL = OFF_RING + SLOTS * SLOT;
constexpr int MAX_TILES = 8;              // 8 x 224 = 1792 coordinates = 597 task vertices
constexpr int OFF_TR = OFF_TMPL + MAX_TILES * NCOL * 4; // 4 warps x [32 frames][33]: transpose tiles of the epilogue
constexpr int OFF_BAR = OFF_TR + 4 * 32 * 33 * 4;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 128;
constexpr int THREADS = 256;              // warp 0 TMA, warp 1 MMA, warps 2-3 idle, warps 4-7 builders then epilogue
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params
{
  int B, ncoord, ntiles;   // coordinates = 3 * vertices used, tiles of 224
  const uint8_t * img;     // [tile][K-block][hi | lo][224][64 B]
  const float * tmpl;      // (ntiles * 224) template coordinate, 0 beyond ncoord
  const float * coef;      // (B, 224): pose feature | beta | 1 | 0
  float out_scale;
  float * rest;            // (B, ncoord)
};
} // namespace rstc

__global__ void __launch_bounds__(rstc::THREADS, 1) ik_restshape_tc_kernel(const rstc::Params p)
{
  using namespace rstc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float * s_tmpl = reinterpret_cast<float *>(smem + OFF_TMPL);
  uint64_t * bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  uint64_t * full_b = bars;            // [SLOTS]
  uint64_t * empty = full_b + SLOTS;   // [SLOTS]
  uint64_t * acc_full = empty + SLOTS; // [2]
  uint64_t * acc_empty = acc_full + 2; // [2]
  uint64_t * a_ready = acc_empty + 2;  // builders -> MMA
  uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(a_ready + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f0 = static_cast<int>(blockIdx.x) * FPB;

  if(warp == 0 && lane == 0)
  {
    for(int s = 0; s < SLOTS; s++)
    {
      ptx::mbar_init(&full_b[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for(int b = 0; b < 2; b++)
    {
      ptx::mbar_init(&acc_full[b], 1);
      ptx::mbar_init(&acc_empty[b], 4);
    }
    ptx::mbar_init(a_ready, 4);
    ptx::fence_barrier_init();
  }
  if(warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  for(int i = threadIdx.x; i < p.ntiles * NCOL; i += THREADS) s_tmpl[i] = __ldg(p.tmpl + i);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int total = p.ntiles * NKB;

  if(warp == 0)
  {
    if(ptx::elect_one())
      for(int it = 0; it < total; it++)
      {
        const int s = it % SLOTS;
        ptx::mbar_wait(&empty[s], ((it / SLOTS) & 1) ^ 1);
        ptx::mbar_expect_tx(&full_b[s], SLOT);
        ptx::bulk_load_1d(smem + OFF_RING + s * SLOT, p.img + static_cast<size_t>(it) * SLOT, SLOT, &full_b[s]);
      }
  }
  else if(warp == 1)
  {
    if(ptx::elect_one())
    {
      constexpr uint32_t DHI = ptx::smem_desc_hi<ROWB>();
      constexpr uint32_t idesc = ptx::make_idesc_f16(128, NCOL);
      const uint32_t smem16 = ptx::smem_u32(smem) >> 4;
      ptx::mbar_wait(a_ready, 0);
      ptx::tc_fence_after();
      int it = 0;
      for(int nt = 0; nt < p.ntiles; nt++)
      {
        const int b = nt & 1;
        if(nt >= 2)
        {
          ptx::mbar_wait(&acc_empty[b], ((nt >> 1) - 1) & 1);
          ptx::tc_fence_after();
        }
        for(int kb = 0; kb < NKB; kb++, it++)
        {
          const int s = it % SLOTS;
          ptx::mbar_wait(&full_b[s], (it / SLOTS) & 1);
          ptx::tc_fence_after();
          const uint32_t a16 = smem16 + ((kb * 2 * A_PART) >> 4), b16 = smem16 + ((OFF_RING + s * SLOT) >> 4);
#pragma unroll
          for(int prod = 0; prod < 3; prod++)
          {
            const int pa = prod == 1 ? 1 : 0, pb = prod == 2 ? 1 : 0; // hi.hi, lo.hi, hi.lo
#pragma unroll
            for(int ks = 0; ks < 2; ks++)
              ptx::umma_f16_ss_lo(tmem_base + b * 256, a16 + ((pa * A_PART + ks * 32) >> 4), b16 + ((pb * B_PART + ks * 32) >> 4),
                                  DHI, idesc, (kb | prod | ks) != 0 ? 1u : 0u);
          }
          ptx::tc_commit(&empty[s]);
        }
        ptx::tc_commit(&acc_full[b]);
      }
    }
  }
  else if(warp >= 4)
  {
    // ---- coefficient rows of this CTA's 128 frames -> fp16 hi | lo K-block images (thread = frame) ----
    const int L = (warp - 4) * 32 + lane;
    const bool live = f0 + L < p.B;
    const float4 * src = reinterpret_cast<const float4 *>(p.coef + static_cast<size_t>(live ? f0 + L : 0) * kBlendK);
#pragma unroll 1
    for(int kb = 0; kb < NKB; kb++)
    {
      float x[32];
#pragma unroll
      for(int c = 0; c < 8; c++)
      {
        const float4 v = live ? __ldg(src + 8 * kb + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        x[4 * c] = v.x, x[4 * c + 1] = v.y, x[4 * c + 2] = v.z, x[4 * c + 3] = v.w;
      }
#pragma unroll
      for(int i = 0; i < 32; i++)
        if(32 * kb + i >= kBlendKUsed - 1) x[i] = 0.f; // the template column (coefficient 1) is added in fp32
      uint8_t * dst = smem + kb * 2 * A_PART;
#pragma unroll
      for(int c = 0; c < 4; c++)
      {
        uint32_t h[4], l[4];
#pragma unroll
        for(int e = 0; e < 4; e++)
        {
          const float a = x[8 * c + 2 * e], b = x[8 * c + 2 * e + 1];
          const float ah = __half2float(__float2half_rn(a)), bh = __half2float(__float2half_rn(b));
          h[e] = skin::pack_half2(ah, bh);
          l[e] = skin::pack_half2(a - ah, b - bh);
        }
        const uint32_t o = static_cast<uint32_t>(L * ROWB + c * 16);
        *reinterpret_cast<uint4 *>(dst + swz64(o)) = make_uint4(h[0], h[1], h[2], h[3]);
        *reinterpret_cast<uint4 *>(dst + swz64(A_PART + o)) = make_uint4(l[0], l[1], l[2], l[3]);
      }
    }
    ptx::fence_proxy_async();
    __syncwarp();
    if(lane == 0) ptx::mbar_arrive(a_ready);
    // ---- epilogue: thread = TMEM lane = frame; a 32 x 32 (frame, coordinate) block is transposed through shared memory
    //      so that every store instruction writes 32 consecutive coordinates of ONE frame (128 contiguous bytes) ----
    const int q = warp & 3;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    float * tile = reinterpret_cast<float *>(smem + OFF_TR) + q * 32 * 33;
    const int fq = f0 + 32 * q; // first frame of this warp
    for(int nt = 0; nt < p.ntiles; nt++)
    {
      const int b = nt & 1;
      ptx::mbar_wait(&acc_full[b], (nt >> 1) & 1);
      ptx::tc_fence_after();
#pragma unroll 1
      for(int ch = 0; ch < NCOL; ch += 32)
      {
        float v[32];
        ptx::tmem_ld_x16(lane_taddr + b * 256 + ch, v);
        ptx::tmem_ld_x16(lane_taddr + b * 256 + ch + 16, v + 16);
        ptx::tmem_ld_wait();
        const int col0 = nt * NCOL + ch;
#pragma unroll
        for(int i = 0; i < 32; i++) tile[lane * 33 + i] = fmaf(v[i], p.out_scale, s_tmpl[col0 + i]);
        __syncwarp();
        if(col0 + lane < p.ncoord)
        {
          const int nfr = min(32, p.B - fq);
          float * dst = p.rest + static_cast<size_t>(fq) * p.ncoord + col0 + lane;
          for(int r = 0; r < nfr; r++) dst[static_cast<size_t>(r) * p.ncoord] = tile[r * 33 + lane];
        }
        __syncwarp();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if(lane == 0) ptx::mbar_arrive(&acc_empty[b]);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if(warp == 1) ptx::tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
namespace sb
{
std::atomic<int> g_poseblend_variant{0};

// stage images of P_m^T for every task from the compact basis rows (3 nUpad, 224) of the task set
int poseblend_tc_prepare(const std::vector<float> & basis, const std::vector<int32_t> & pair_off,
                         const std::vector<int32_t> & pair_vert, PoseBlendTc & out, std::vector<void *> & allocations)
{
  using namespace pbtc;
  out.ready = false;
  int dev = 0, major = 0;
  SB_CUDA(cudaGetDevice(&dev));
  SB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if(major != 10) return SMPLPP_OK; // tcgen05 needs sm_100
  const int n = static_cast<int>(pair_off.size()) - 1;
  std::vector<int32_t> slot_off(n + 1, 0);
  for(int m = 0; m < n; m++)
  {
    const int np = pair_off[m + 1] - pair_off[m];
    if(3 * np > 64) return SMPLPP_OK; // more than two K-blocks: the FFMA phase of ik_jacobian_kernel stays in charge
    slot_off[m + 1] = slot_off[m] + (3 * np + 31) / 32;
  }
  float mx = 0.f;
  for(int m = 0; m < n; m++)
    for(int q = pair_off[m]; q < pair_off[m + 1]; q++)
      for(int a = 0; a < 3; a++)
        for(int c = 0; c < kBlendKUsed - 1; c++) // the template column takes no part
          mx = std::fmax(mx, std::fabs(basis[(static_cast<size_t>(3) * pair_vert[q] + a) * kBlendK + c]));
  int e = 0;
  if(mx > 0.f)
  {
    while(mx * std::ldexp(1.f, e + 1) <= 1024.f && e < 30) e++;
    while(mx * std::ldexp(1.f, e) > 1024.f && e > -30) e--;
  }
  const float scale = std::ldexp(1.f, e);
  const size_t total = static_cast<size_t>(slot_off[n]) * 2 * B_PART;
  std::vector<uint8_t> img(total, 0);
  for(int m = 0; m < n; m++)
  {
    const int np = pair_off[m + 1] - pair_off[m];
    for(int k = 0; k < 3 * np; k++)
    {
      const int u = pair_vert[pair_off[m] + k / 3], a = k % 3;
      uint8_t * blk = img.data() + static_cast<size_t>(slot_off[m] + k / 32) * (2 * B_PART);
      for(int c = 0; c < kBlendKUsed - 1; c++)
      {
        const float x = basis[(static_cast<size_t>(3) * u + a) * kBlendK + c] * scale;
        const __half hi = __float2half_rn(x);
        const __half lo = __float2half_rn(x - __half2float(hi));
        const uint32_t o = static_cast<uint32_t>(c * ROWB + (k % 32) * 2);
        *reinterpret_cast<__half *>(blk + swz64(o)) = hi;
        *reinterpret_cast<__half *>(blk + swz64(static_cast<uint32_t>(B_PART) + o)) = lo;
      }
    }
  }
  void * d_img = nullptr;
  void * d_off = nullptr;
  SB_CUDA(cudaMalloc(&d_img, std::max<size_t>(total, 16)));
  allocations.push_back(d_img);
  SB_CUDA(cudaMemcpy(d_img, img.data(), total, cudaMemcpyHostToDevice));
  SB_CUDA(cudaMalloc(&d_off, slot_off.size() * sizeof(int32_t)));
  allocations.push_back(d_off);
  SB_CUDA(cudaMemcpy(d_off, slot_off.data(), slot_off.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  out.img = static_cast<const uint8_t *>(d_img);
  out.slot_off = static_cast<const int32_t *>(d_off);
  out.slots = slot_off[n];
  out.basis_exp = e;
  // rest-shape operand: the K-major basis rows of the nU task vertices, 224 coordinates per tile (the scale e above bounds
  // every used row: each vertex of the task set belongs to a pair)
  {
    const int nU = static_cast<int>(basis.size() / (static_cast<size_t>(3) * kBlendK)); // nUpad; rows beyond nU are zero
    int used = 0;
    for(int32_t u : pair_vert) used = std::max(used, u + 1);
    const int ncoord = 3 * used, ntiles = (ncoord + NCOL - 1) / NCOL;
    if(ntiles <= rstc::MAX_TILES && used <= nU)
    {
      std::vector<uint8_t> rimg(static_cast<size_t>(ntiles) * rstc::NKB * rstc::SLOT, 0);
      std::vector<float> tmpl(static_cast<size_t>(ntiles) * NCOL, 0.f);
      for(int r = 0; r < ncoord; r++)
      {
        const int nt = r / NCOL, row = r % NCOL;
        tmpl[r] = basis[static_cast<size_t>(r) * kBlendK + kBlendKUsed - 1];
        for(int c = 0; c < kBlendKUsed - 1; c++)
        {
          const float x = basis[static_cast<size_t>(r) * kBlendK + c] * scale;
          const __half hi = __float2half_rn(x);
          const __half lo = __float2half_rn(x - __half2float(hi));
          uint8_t * blk = rimg.data() + (static_cast<size_t>(nt) * rstc::NKB + c / 32) * rstc::SLOT;
          const uint32_t o = static_cast<uint32_t>(row * ROWB + (c % 32) * 2);
          *reinterpret_cast<__half *>(blk + swz64(o)) = hi;
          *reinterpret_cast<__half *>(blk + swz64(static_cast<uint32_t>(B_PART) + o)) = lo;
        }
      }
      void * d_rimg = nullptr;
      void * d_tmpl = nullptr;
      SB_CUDA(cudaMalloc(&d_rimg, rimg.size()));
      allocations.push_back(d_rimg);
      SB_CUDA(cudaMemcpy(d_rimg, rimg.data(), rimg.size(), cudaMemcpyHostToDevice));
      SB_CUDA(cudaMalloc(&d_tmpl, tmpl.size() * sizeof(float)));
      allocations.push_back(d_tmpl);
      SB_CUDA(cudaMemcpy(d_tmpl, tmpl.data(), tmpl.size() * sizeof(float), cudaMemcpyHostToDevice));
      out.rest_img = static_cast<const uint8_t *>(d_rimg);
      out.rest_tmpl = static_cast<const float *>(d_tmpl);
      out.rest_ready = true;
    }
  }
  out.ready = true;
  return SMPLPP_OK;
}

// rest (B, 3 * n_vertices) of the first n_vertices task vertices
int launch_restshape_tc(const PoseBlendTc & pb, cudaStream_t st, int B, int n_vertices, const float * coef, float * rest)
{
  using namespace rstc;
  static bool attr_done[64] = {};
  if(first_call_on_device(attr_done))
    SB_CUDA(cudaFuncSetAttribute(ik_restshape_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  Params p{};
  p.B = B, p.ncoord = 3 * n_vertices, p.ntiles = (p.ncoord + NCOL - 1) / NCOL;
  p.img = pb.rest_img, p.tmpl = pb.rest_tmpl, p.coef = coef;
  p.out_scale = std::ldexp(1.f, -pb.basis_exp);
  p.rest = rest;
  ik_restshape_tc_kernel<<<(B + FPB - 1) / FPB, THREADS, SMEM_BYTES, st>>>(p);
  SB_LAUNCHED();
  return SMPLPP_OK;
}

int launch_poseblend_tc(const PoseBlendTc & pb, const TasksDev & t, cudaStream_t st, int B, int rows, int use_ring, int beta_col,
                        const float * ca, const float * dr, float * J, int ld)
{
  using namespace pbtc;
  static bool attr_done[64] = {};
  if(first_call_on_device(attr_done))
    SB_CUDA(cudaFuncSetAttribute(ik_poseblend_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  Params p{};
  p.B = B, p.n = t.n, p.rows = rows, p.use_ring = use_ring, p.beta_col = beta_col;
  p.ca_stride = pb.slots * 128;
  p.slot_off = pb.slot_off, p.pair_off = t.pair_off, p.img = pb.img, p.ca = ca, p.dr = dr;
  p.out_scale = std::ldexp(1.f, -pb.basis_exp);
  p.J = J, p.ld = ld;
  ik_poseblend_tc_kernel<<<(B + FPB - 1) / FPB, THREADS, SMEM_BYTES, st>>>(p);
  SB_LAUNCHED();
  return SMPLPP_OK;
}
} // namespace sb
