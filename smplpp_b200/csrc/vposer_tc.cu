// VPoser decoder Jacobian d(axis-angle)/d(latent) (63 x 32 per frame) on tcgen05.
//
// Reference: the node obtains these entries row by row from autograd through VPoserDecoderImpl::forward
// (src/VPoser.cpp:143-167, node/node.cpp:761-772, 823-873).  Forward mode, 32 tangents per frame:
//
//   X   = D1 W0                         (512 x 32)    D1 = diag(LeakyReLU'(h1)) in {1, 0.01}
//   C   = W3 X                          (512 x 32)    16.8 MFLOP per frame: the bulk
//   T3  = W5 (D2 C)                     (126 x 32)     4.1 MFLOP per frame
//   J   = d aa / d y6 . T3              (63 x 32)     6 FMA per entry
//
// The FFMA kernel (vposer.cu) spends ~13 ms per 16384 frames on it.  Here four frames form one M = 128 tile
// (TMEM lane n = 32 frame + tangent) and both products run TRANSPOSED so that the result of the first is already laid
// out as the TMEM-resident A operand of the second:
//
//   GEMM 1   C^T[n, i]  = sum_k X^T[n, k] W3[i, k]      A = X^T K-block generated in shared memory by the worker warps
//                                                        (fp16 hi | lo, SWIZZLE_64B), B = W3 stage image (cp.async.bulk),
//                                                        N = 256 hidden units per half, fp32 accumulators in TMEM [0,256)
//   convert  A2[n, i]   = fp16 hi | lo of d2[i] C^T[n, i]  in place: 256 fp32 columns -> 128 hi + 128 lo columns
//   GEMM 2   T3^T[n, o] += sum_i A2[n, i] W5[o, i]       A in TMEM, B = W5 K-block image, accumulators in TMEM [384,512)
//   epilogue J[f, 3j + r, t] = sum_c daa[f][j][r][c] T3^T[n, 6j + c]   thread = (frame, tangent): coalesced rows of J
//
// Split precision as in skin_tc3.cu: every operand is fp16 hi + lo with power-of-two pre-scales chosen at create time
// from the weights (the A2 scale from the bound max_i sum_k |W3[i,k]| . max |W0|, so it cannot overflow), three
// products hi.hi + lo.hi + hi.lo accumulated in fp32.
// warp 0: TMA producer of the W3 ring | warp 1: TMEM allocator + MMA issuer | warp 2: TMA producer of the W5 ring |
// warp 3: idle | warps 4-19: workers (operand generation, conversion, epilogue), 112 registers via setmaxnreg.
#include <cuda_fp16.h>

#include <cmath>
#include <vector>

#include "skin_common.cuh"
#include "tc_ptx.cuh"
#include "vposer.cuh"

using namespace sb;

namespace vtc
{
constexpr int H = SMPLPP_VPOSER_HIDDEN, L = SMPLPP_LATENT_DIM, NJ = SMPLPP_VPOSER_JOINTS, OUT = 6 * NJ;
constexpr int FPP = 4;                      // frames per pass: 4 x 32 tangents = 128 TMEM lanes
constexpr int ROWB = 64;                    // K-block of 32 fp16 = one SWIZZLE_64B span
constexpr int NKB = H / 32;                 // 16 K-blocks
constexpr int NH = 256;                     // hidden units per half (UMMA N of GEMM 1)
constexpr int A_PART = 128 * ROWB;          // 8192: X^T K-block, one part
constexpr int B_PART = NH * ROWB;           // 16384: W3 K-block of one half, one part
constexpr int STAGE = 2 * A_PART + 2 * B_PART; // 49152
constexpr int STAGES = 2;
constexpr int W5_PART = 128 * ROWB;         // 8192: W5 K-block (126 outputs padded to 128), one part
constexpr int W5_SLOT = 2 * W5_PART;        // 16384
constexpr int W5_SLOTS = 2;
constexpr int W0T_LD = H + 4;               // padded row of the fp32 W0^T copy in shared memory
constexpr int AUX_FLOATS = 2 * H + NJ * 18 + 6; // per frame: d1 | d2 | daa | pad  (1408)
constexpr int OFF_W5 = STAGES * STAGE;
constexpr int OFF_W0T = OFF_W5 + W5_SLOTS * W5_SLOT;
constexpr int OFF_AUX = OFF_W0T + L * W0T_LD * 4;
constexpr int OFF_BAR = OFF_AUX + FPP * AUX_FLOATS * 4;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 256;
constexpr int CTRL_WARPS = 4, WORK_WARPS = 16, THREADS = 32 * (CTRL_WARPS + WORK_WARPS);
constexpr int COL_T3 = 384;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
static_assert(AUX_FLOATS % 4 == 0 && OFF_W0T % 16 == 0 && OFF_AUX % 16 == 0, "alignment");

__host__ __device__ constexpr uint32_t swz64(uint32_t o)
{
  return o ^ (((o >> 7) & 3u) << 4);
}

struct Params
{
  int B, npass;
  const uint8_t * img_w3;   // [half][K-block][part][256][64 B]
  const uint8_t * img_w5;   // [K-block][part][128][64 B]
  const float * w0t;        // (32, 512) fp32, scaled by 2^ex
  const float * aux;        // (B, AUX_FLOATS) from the forward kernel
  float conv_scale;         // 2^(ea - ex - e3)
  float out_scale;          // 2^-(ea + e5)
  float * jac;              // (B, 63, 32)
};
} // namespace vtc

// fp32 (rows, 512) weight -> K-block images [block][part][rows_per_block][32 fp16], SWIZZLE_64B, scaled by `scale`;
// row r of the weight goes to block (r / rows_per_block) * 16 + kb, row r % rows_per_block; rows >= rows_valid are zero
__global__ void weight_image_kernel(const float * __restrict__ w, int rows_valid, int rows_total, int rows_per_block, float scale,
                                    uint8_t * __restrict__ img)
{
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if(i >= static_cast<long long>(rows_total) * vtc::H) return;
  const int k = static_cast<int>(i % vtc::H), r = static_cast<int>(i / vtc::H);
  const float x = r < rows_valid ? w[static_cast<size_t>(r) * vtc::H + k] * scale : 0.f;
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  const size_t part_bytes = static_cast<size_t>(rows_per_block) * vtc::ROWB;
  uint8_t * blk = img + (static_cast<size_t>(r / rows_per_block) * vtc::NKB + k / 32) * (2 * part_bytes);
  const uint32_t o = static_cast<uint32_t>((r % rows_per_block) * vtc::ROWB + (k % 32) * 2);
  *reinterpret_cast<__half *>(blk + vtc::swz64(o)) = hi;
  *reinterpret_cast<__half *>(blk + vtc::swz64(static_cast<uint32_t>(part_bytes) + o)) = lo;
}

__global__ void __launch_bounds__(vtc::THREADS, 1) vposer_jac_tc_kernel(const vtc::Params p)
{
  using namespace vtc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float * s_w0t = reinterpret_cast<float *>(smem + OFF_W0T);
  float * s_aux = reinterpret_cast<float *>(smem + OFF_AUX);
  uint64_t * bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  uint64_t * b_full = bars;                  // [STAGES]   TMA -> MMA (W3 K-block)
  uint64_t * a_full = b_full + STAGES;       // [STAGES]   workers -> MMA (X^T K-block)
  uint64_t * empty = a_full + STAGES;        // [STAGES]   MMA -> TMA + workers
  uint64_t * w5_full = empty + STAGES;       // [W5_SLOTS]
  uint64_t * w5_empty = w5_full + W5_SLOTS;  // [W5_SLOTS]
  uint64_t * c_full = w5_empty + W5_SLOTS;   // MMA -> workers: GEMM 1 of a half complete
  uint64_t * a2_ready = c_full + 1;          // workers -> MMA: converted A operand stored
  uint64_t * g2_done = a2_ready + 1;         // MMA -> MMA / workers: GEMM 2 of a half complete (TMEM [0,256) reusable)
  uint64_t * t3_free = g2_done + 1;          // workers -> MMA: T3 accumulators read
  uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(t3_free + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // passes of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
  const int my_passes = (p.npass - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if(warp == 0 && lane == 0)
  {
    for(int s = 0; s < STAGES; s++)
    {
      ptx::mbar_init(&b_full[s], 1);
      ptx::mbar_init(&a_full[s], WORK_WARPS);
      ptx::mbar_init(&empty[s], 1);
    }
    for(int s = 0; s < W5_SLOTS; s++)
    {
      ptx::mbar_init(&w5_full[s], 1);
      ptx::mbar_init(&w5_empty[s], 1);
    }
    ptx::mbar_init(c_full, 1);
    ptx::mbar_init(a2_ready, WORK_WARPS);
    ptx::mbar_init(g2_done, 1);
    ptx::mbar_init(t3_free, WORK_WARPS);
    ptx::fence_barrier_init();
  }
  if(warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  // W0^T (pre-scaled fp32) stays in shared memory for the whole kernel
  for(int i = threadIdx.x; i < L * H; i += THREADS) s_w0t[(i / H) * W0T_LD + (i % H)] = __ldg(p.w0t + i);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if(warp < CTRL_WARPS)
  {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
    if(warp == 0)
    {
      // ---- W3 ring: per pass 2 halves x 16 K-blocks ----
      if(ptx::elect_one())
      {
        const int total = my_passes * 2 * NKB;
        for(int n = 0; n < total; n++)
        {
          const int s = n % STAGES, hk = n % (2 * NKB);
          ptx::mbar_wait(&empty[s], ((n / STAGES) & 1) ^ 1);
          ptx::mbar_expect_tx(&b_full[s], 2 * B_PART);
          ptx::bulk_load_1d(smem + s * STAGE + 2 * A_PART, p.img_w3 + static_cast<size_t>(hk) * (2 * B_PART), 2 * B_PART, &b_full[s]);
        }
      }
    }
    else if(warp == 2)
    {
      // ---- W5 ring: per pass 16 K-blocks ----
      if(ptx::elect_one())
      {
        const int total = my_passes * NKB;
        for(int n = 0; n < total; n++)
        {
          const int s = n % W5_SLOTS;
          ptx::mbar_wait(&w5_empty[s], ((n / W5_SLOTS) & 1) ^ 1);
          ptx::mbar_expect_tx(&w5_full[s], W5_SLOT);
          ptx::bulk_load_1d(smem + OFF_W5 + s * W5_SLOT, p.img_w5 + static_cast<size_t>(n % NKB) * W5_SLOT, W5_SLOT, &w5_full[s]);
        }
      }
    }
    else if(warp == 1)
    {
      if(ptx::elect_one())
      {
        constexpr uint32_t DHI = ptx::smem_desc_hi<ROWB>();
        constexpr uint32_t idesc1 = ptx::make_idesc_f16(128, NH);
        constexpr uint32_t idesc2 = ptx::make_idesc_f16(128, 128);
        const uint32_t smem16 = ptx::smem_u32(smem) >> 4;
        int n1 = 0, n5 = 0, nhalf = 0;
        for(int ps = 0; ps < my_passes; ps++)
        {
          for(int h = 0; h < 2; h++, nhalf++)
          {
            // GEMM 1 of this half overwrites TMEM [0,256): the previous half's GEMM 2 (which reads it as A) must be done
            if(nhalf > 0)
            {
              ptx::mbar_wait(g2_done, (nhalf - 1) & 1);
              ptx::tc_fence_after();
            }
            for(int kb = 0; kb < NKB; kb++, n1++)
            {
              const int s = n1 % STAGES;
              ptx::mbar_wait(&b_full[s], (n1 / STAGES) & 1);
              ptx::mbar_wait(&a_full[s], (n1 / STAGES) & 1);
              ptx::tc_fence_after();
              const uint32_t st16 = smem16 + s * (STAGE >> 4);
#pragma unroll
              for(int prod = 0; prod < 3; prod++)
              {
                const int pa = prod == 1 ? 1 : 0, pb = prod == 2 ? 1 : 0; // hi.hi, lo.hi, hi.lo
#pragma unroll
                for(int ks = 0; ks < 2; ks++)
                  ptx::umma_f16_ss_lo(tmem_base, st16 + ((pa * A_PART + ks * 32) >> 4), st16 + ((2 * A_PART + pb * B_PART + ks * 32) >> 4),
                                      DHI, idesc1, (kb | prod | ks) != 0 ? 1u : 0u);
              }
              ptx::tc_commit(&empty[s]);
            }
            ptx::tc_commit(c_full);
            // GEMM 2 over the 256 hidden units of this half
            ptx::mbar_wait(a2_ready, nhalf & 1);
            if(h == 0 && ps > 0) ptx::mbar_wait(t3_free, (ps - 1) & 1);
            ptx::tc_fence_after();
            for(int kb2 = 0; kb2 < NKB / 2; kb2++, n5++)
            {
              const int s = n5 % W5_SLOTS;
              ptx::mbar_wait(&w5_full[s], (n5 / W5_SLOTS) & 1);
              ptx::tc_fence_after();
              const uint32_t sw16 = smem16 + ((OFF_W5 + s * W5_SLOT) >> 4);
#pragma unroll
              for(int prod = 0; prod < 3; prod++)
              {
                const int pa = prod == 1 ? 1 : 0, pb = prod == 2 ? 1 : 0;
#pragma unroll
                for(int ks = 0; ks < 2; ks++)
                  ptx::umma_f16_ts_lo(tmem_base + COL_T3, tmem_base + pa * 128 + kb2 * 16 + ks * 8,
                                      sw16 + ((pb * W5_PART + ks * 32) >> 4), DHI, idesc2, (h | kb2 | prod | ks) != 0 ? 1u : 0u);
              }
              ptx::tc_commit(&w5_empty[s]);
            }
            ptx::tc_commit(g2_done);
          }
        }
      }
    }
  }
  else
  {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int ww = warp - CTRL_WARPS;      // 0..15
    const int q = warp & 3;                // TMEM lane quadrant = frame of the pass
    const int g = ww >> 2;                 // column group / joint group
    const int wt = threadIdx.x - CTRL_WARPS * 32; // 0..511
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    int n1 = 0, nhalf = 0;
    for(int ps = 0; ps < my_passes; ps++)
    {
      const int f0 = (static_cast<int>(blockIdx.x) + ps * static_cast<int>(gridDim.x)) * FPP;
      // every worker is done with the previous pass's aux data (named barrier 1: the 512 worker threads)
      asm volatile("bar.sync 1, 512;" ::: "memory");
      for(int i = wt; i < FPP * AUX_FLOATS / 4; i += WORK_WARPS * 32)
      {
        const int f = i / (AUX_FLOATS / 4);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if(f0 + f < p.B) v = __ldg(reinterpret_cast<const float4 *>(p.aux + static_cast<size_t>(f0) * AUX_FLOATS) + i);
        reinterpret_cast<float4 *>(s_aux)[i] = v;
      }
      asm volatile("bar.sync 1, 512;" ::: "memory");
      for(int h = 0; h < 2; h++, nhalf++)
      {
        // ---- X^T K-blocks: thread -> row n = wt / 4 (frame n / 32, tangent n % 32), 16-byte chunk c = wt % 4 ----
        {
          const int n = wt >> 2, c = wt & 3;
          const float * d1 = s_aux + (n >> 5) * AUX_FLOATS;
          const float * w0r = s_w0t + (n & 31) * W0T_LD;
          const uint32_t o = static_cast<uint32_t>(n * ROWB + c * 16);
          for(int kb = 0; kb < NKB; kb++, n1++)
          {
            const int s = n1 % STAGES;
            const int k0 = kb * 32 + c * 8;
            const float4 da = *reinterpret_cast<const float4 *>(d1 + k0), db = *reinterpret_cast<const float4 *>(d1 + k0 + 4);
            const float4 wa = *reinterpret_cast<const float4 *>(w0r + k0), wb = *reinterpret_cast<const float4 *>(w0r + k0 + 4);
            const float x[8] = {da.x * wa.x, da.y * wa.y, da.z * wa.z, da.w * wa.w, db.x * wb.x, db.y * wb.y, db.z * wb.z, db.w * wb.w};
            uint32_t hi[4], lo[4];
#pragma unroll
            for(int e = 0; e < 4; e++)
            {
              const float a = x[2 * e], b = x[2 * e + 1];
              const float ah = __half2float(__float2half_rn(a)), bh = __half2float(__float2half_rn(b));
              hi[e] = skin::pack_half2(ah, bh);
              lo[e] = skin::pack_half2(a - ah, b - bh);
            }
            ptx::mbar_wait(&empty[s], ((n1 / STAGES) & 1) ^ 1);
            uint8_t * dst = smem + s * STAGE;
            *reinterpret_cast<uint4 *>(dst + swz64(o)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4 *>(dst + swz64(A_PART + o)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            ptx::fence_proxy_async(); // generic-proxy stores -> visible to the tensor core's async-proxy reads
            __syncwarp();
            if(lane == 0) ptx::mbar_arrive(&a_full[s]);
          }
        }
        // ---- convert: C^T (fp32, 64 columns per warp) -> d2-scaled fp16 hi | lo, in place ----
        ptx::mbar_wait(c_full, nhalf & 1);
        ptx::tc_fence_after();
        {
          float v[64];
#pragma unroll
          for(int i = 0; i < 4; i++) ptx::tmem_ld_x16(lane_taddr + g * 64 + i * 16, v + i * 16);
          ptx::tmem_ld_wait();
          // the four warps of this quadrant overwrite each other's source columns: all loads first (named barrier 2 + q)
          asm volatile("bar.sync %0, 128;" ::"r"(2 + q) : "memory");
          const float * d2 = s_aux + q * AUX_FLOATS + H + h * NH + g * 64;
#pragma unroll
          for(int i = 0; i < 2; i++) // 32 columns at a time: 64 source + 32 packed registers stay under the 112 budget
          {
            uint32_t hi[16], lo[16];
#pragma unroll
            for(int e = 0; e < 16; e++)
            {
              const int c = 32 * i + 2 * e;
              const float a = v[c] * d2[c] * p.conv_scale, b = v[c + 1] * d2[c + 1] * p.conv_scale;
              const float ah = __half2float(__float2half_rn(a)), bh = __half2float(__float2half_rn(b));
              hi[e] = skin::pack_half2(ah, bh);
              lo[e] = skin::pack_half2(a - ah, b - bh);
            }
            ptx::tmem_st_x16(lane_taddr + g * 32 + i * 16, hi);
            ptx::tmem_st_x16(lane_taddr + 128 + g * 32 + i * 16, lo);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if(lane == 0) ptx::mbar_arrive(a2_ready);
        }
      }
      // ---- epilogue: J rows of this warp's joints; thread = (frame q, tangent lane) ----
      ptx::mbar_wait(g2_done, (nhalf - 1) & 1);
      ptx::tc_fence_after();
      {
        const int j0 = g == 0 ? 0 : 1 + 5 * g, j1 = 6 + 5 * g; // joints 0-5 | 6-10 | 11-15 | 16-20
        const float * daa = s_aux + q * AUX_FLOATS + 2 * H;
        const bool live = f0 + q < p.B;
        float * jrow = p.jac + (static_cast<size_t>(f0 + q) * 63) * L + lane;
        for(int j = j0; j < j1; j++)
        {
          float t3[8];
          ptx::tmem_ld_x8(lane_taddr + COL_T3 + 6 * j, t3);
          ptx::tmem_ld_wait();
          if(live)
          {
#pragma unroll
            for(int r = 0; r < 3; r++)
            {
              float acc = 0.f;
#pragma unroll
              for(int c = 0; c < 6; c++) acc = fmaf(daa[j * 18 + r * 6 + c], t3[c], acc);
              jrow[static_cast<size_t>(3 * j + r) * L] = acc * p.out_scale;
            }
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if(lane == 0) ptx::mbar_arrive(t3_free);
      }
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
namespace
{
// largest power of two that keeps max_abs * 2^e <= limit
int pow2_scale(float max_abs, float limit)
{
  int e = 0;
  if(!(max_abs > 0.f)) return 0;
  while(max_abs * ldexpf(1.f, e + 1) <= limit && e < 30) e++;
  while(max_abs * ldexpf(1.f, e) > limit && e > -30) e--;
  return e;
}
} // namespace

// builds the fp16 stage images of W3 / W5, the scaled W0^T and the scales; host weights are row-major (out, in)
int vposer_tc_prepare(smplpp_vposer & v, const float * w0, const float * w3, const float * w5)
{
  using namespace vtc;
  v.tc_ready = false;
  int dev = 0, major = 0;
  SB_CUDA(cudaGetDevice(&dev));
  SB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  if(major != 10) return SMPLPP_OK; // tcgen05 needs sm_100: the FFMA kernel stays in charge elsewhere
  float m0 = 0.f, m3 = 0.f, m5 = 0.f, rowsum3 = 0.f;
  for(int i = 0; i < H * L; i++) m0 = std::fmax(m0, std::fabs(w0[i]));
  for(int i = 0; i < H; i++)
  {
    float rs = 0.f;
    for(int k = 0; k < H; k++)
    {
      const float a = std::fabs(w3[static_cast<size_t>(i) * H + k]);
      m3 = std::fmax(m3, a);
      rs += a;
    }
    rowsum3 = std::fmax(rowsum3, rs);
  }
  for(int i = 0; i < OUT * H; i++) m5 = std::fmax(m5, std::fabs(w5[i]));
  const int ex = pow2_scale(m0, 1024.f), e3 = pow2_scale(m3, 1024.f), e5 = pow2_scale(m5, 1024.f);
  const int ea = pow2_scale(rowsum3 * m0, 30000.f); // |d2 C| <= rowsum(|W3|) max|W0|: the converted operand cannot overflow
  v.tc_conv_scale = ldexpf(1.f, ea - ex - e3);
  v.tc_out_scale = ldexpf(1.f, -(ea + e5));
  std::vector<float> w0t(static_cast<size_t>(L) * H);
  for(int k = 0; k < H; k++)
    for(int t = 0; t < L; t++) w0t[static_cast<size_t>(t) * H + k] = w0[static_cast<size_t>(k) * L + t] * ldexpf(1.f, ex);
  SB_CUDA(cudaMalloc(&v.tc_w0t, w0t.size() * sizeof(float)));
  SB_CUDA(cudaMemcpy(v.tc_w0t, w0t.data(), w0t.size() * sizeof(float), cudaMemcpyHostToDevice));
  float *d3 = nullptr, *d5 = nullptr;
  SB_CUDA(cudaMalloc(&d3, static_cast<size_t>(H) * H * sizeof(float)));
  SB_CUDA(cudaMalloc(&d5, static_cast<size_t>(OUT) * H * sizeof(float)));
  SB_CUDA(cudaMemcpy(d3, w3, static_cast<size_t>(H) * H * sizeof(float), cudaMemcpyHostToDevice));
  SB_CUDA(cudaMemcpy(d5, w5, static_cast<size_t>(OUT) * H * sizeof(float), cudaMemcpyHostToDevice));
  SB_CUDA(cudaMalloc(&v.tc_img_w3, static_cast<size_t>(2) * NKB * 2 * B_PART));
  SB_CUDA(cudaMalloc(&v.tc_img_w5, static_cast<size_t>(NKB) * W5_SLOT));
  weight_image_kernel<<<(H * H + 255) / 256, 256>>>(d3, H, H, NH, ldexpf(1.f, e3), static_cast<uint8_t *>(v.tc_img_w3));
  SB_LAUNCHED();
  weight_image_kernel<<<(128 * H + 255) / 256, 256>>>(d5, OUT, 128, 128, ldexpf(1.f, e5), static_cast<uint8_t *>(v.tc_img_w5));
  SB_LAUNCHED();
  SB_CUDA(cudaDeviceSynchronize());
  cudaFree(d3);
  cudaFree(d5);
  SB_CUDA(cudaFuncSetAttribute(vposer_jac_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  SB_CUDA(cudaDeviceGetAttribute(&v.tc_sms, cudaDevAttrMultiProcessorCount, dev));
  v.tc_ready = v.tc_sms > 0;
  return SMPLPP_OK;
}

void vposer_tc_release(smplpp_vposer & v)
{
  cudaFree(v.tc_w0t);
  cudaFree(v.tc_img_w3);
  cudaFree(v.tc_img_w5);
  v.tc_w0t = nullptr;
  v.tc_img_w3 = v.tc_img_w5 = nullptr;
  v.tc_ready = false;
}

size_t vposer_tc_aux_floats()
{
  return vtc::AUX_FLOATS;
}

// aux (B, AUX_FLOATS) holds d1 | d2 | daa of every frame (written by the forward kernel); jac (B, 63, 32)
int launch_vposer_jac_tc(const smplpp_vposer & v, cudaStream_t st, int B, const float * aux, float * jac)
{
  using namespace vtc;
  if(!v.tc_ready) return fail(SMPLPP_ERR_INVALID, "VPoser", "tensor-core Jacobian is not available on this device");
  Params p;
  p.B = B;
  p.npass = (B + FPP - 1) / FPP;
  p.img_w3 = static_cast<const uint8_t *>(v.tc_img_w3);
  p.img_w5 = static_cast<const uint8_t *>(v.tc_img_w5);
  p.w0t = v.tc_w0t;
  p.aux = aux;
  p.conv_scale = v.tc_conv_scale;
  p.out_scale = v.tc_out_scale;
  p.jac = jac;
  const int grid = p.npass < v.tc_sms ? p.npass : v.tc_sms;
  vposer_jac_tc_kernel<<<grid, THREADS, SMEM_BYTES, st>>>(p);
  SB_LAUNCHED();
  return SMPLPP_OK;
}
} // namespace sb
