// K2'' (tcgen05): fused blend-shape contraction + linear blend skinning with BOTH products on the tensor cores.
//
// Reference semantics: BlendShape::poseBlend / shapeBlend (src/BlendShape.cpp:764, 670-683), the rest shape
// T + S + P (src/JointRegression.cpp:551-565) and LinearBlendSkinning::skinning (src/LinearBlendSkinning.cpp:445-553).
//
//   rest[v,f]   = T[v] + sum_k basis[v,k] coef[f,k]                 GEMM 1: (128 vertices x 3 planes) x 96 frames, K = 224
//   M[v,f]      = sum_j W[v,j] G'[f,j]  (3x4 per vertex and frame)  GEMM 2: 128 vertices x (8 frames x 12), K = 24 -> 32
//   vert[v,f]   = M[v,f] [rest; 1] / sum_j W[v,j] + trans[f]        epilogue: 15 FMA per vertex and frame
//
// Moving the skinning matrices to the tensor cores removes what bounded the first tcgen05 kernel (blend_tc.cu): its
// epilogue fetched 4 joints x 48 B of transforms from shared memory per vertex and frame (91 % of the cycles,
// l1tex-bound, tensor pipe 15 % busy).  Here the epilogue reads 12 + 3 TMEM words per vertex and frame instead.
//
// Precision: both GEMMs are fp32 in the reference.  Every operand is split into fp16 hi + lo (power-of-two
// pre-scaled so that lo stays a normal fp16) and three products are accumulated in fp32 TMEM,
// hi.hi + lo.hi + hi.lo (~2^-22 relative, the accuracy of 3xTF32 at twice its rate).  The ~1 m template stays out of
// GEMM 1 and is added in fp32.
//
// CTA = 128 vertices x 96 frames.  TMEM (512 columns): [0,288) rest accumulators (x | y | z planes x 96 frames),
// [288,480) two 96-column buffers of skinning matrices (8 frames x 12), [480,512) the CTA's W tile as the A operand
// of GEMM 2 (fp16 hi | lo, written once with tcgen05.st by the threads that own the vertices).
// warp 0: TMA producer | warp 1: TMEM allocator + MMA issuer | warps 2-17: epilogue (four per TMEM lane quadrant,
// one per frame pair of the 8-frame sub-batch; every warp visits every sub-batch, which keeps all sixteen busy).
// Shared memory: 3 x 60 KB GEMM 1 stages + 2 x 12 KB transform sub-batches loaded up front; once GEMM 1 is complete
// its stages are reused for the output staging (12 KB) and for the tile's other 10 transform sub-batches, all
// requested at once (a 4-deep ring was bound by the TMA round trip).
// Sizing (measured, scripts/ubench/ubench_mma.cu): one thread issues a tcgen05.mma every max(N / 2, ~40) cycles when
// consecutive MMAs hit different accumulators and every ~67 cycles when they chain on one accumulator, whatever N is.
// The kernel is bound by its MMA COUNT (126 per tile for GEMM 1, 6 per sub-batch for GEMM 2), so GEMM 2 wants the
// largest N the TMEM left over by the rest accumulators allows: 96 frames x 3 + 2 x 96 + 32 = 512 columns.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <vector>

#include "forward.cuh"
#include "skin_common.cuh"
#include "tc_ptx.cuh"

using namespace sb;

namespace tc2
{
constexpr int MV = 128;                          // vertices per tile (UMMA M)
constexpr int NF = 96;                           // frames per tile (UMMA N of GEMM 1)
constexpr int ROWB = 64;                         // bytes of K per shared-memory row (one SWIZZLE_64B span = 32 fp16)
constexpr int KP = kBlendK;                      // 224
constexpr int KUSED = kPoseDim + kShapeDim;      // 217: the template column is excluded
constexpr int NKB = KP * 2 / ROWB;               // 7 K-blocks of 32
constexpr int STAGES = 3;
constexpr int A_PART = 3 * MV * ROWB;            // 24576: one part (hi or lo) of the basis tile, 3 planes
constexpr int B_PART = NF * ROWB;                // 8192
constexpr int STAGE = 2 * A_PART + 2 * B_PART;   // 65536
constexpr int SUBF = 8;                          // frames per skinning sub-batch
constexpr int SUBN = SUBF * kXformFloats;        // 96 = UMMA N of GEMM 2
constexpr int KJ = 32;                           // joints padded to two K = 16 steps
constexpr int G_PART = SUBN * ROWB;              // 3072
constexpr int G_STAGE = 2 * G_PART;              // 6144
constexpr int GS = 2;                            // transform sub-batches with a slot of their own (loaded up front)
constexpr int NSUB = NF / SUBF;                  // 12 sub-batches per tile
constexpr int EPI_WARPS = 16;                     // four per TMEM lane quadrant: {matrix buffer 0, 1} x {frame pair 0, 1}
constexpr int EPI_FR = 2;                         // frames of a sub-batch handled by one epilogue warp
constexpr int THREADS = 32 * (2 + EPI_WARPS);
constexpr int STG_FLOATS = EPI_FR * 32 * 3;       // per warp: 2 frames x 32 vertices x 3
constexpr int OFF_G = STAGES * STAGE;
constexpr int OFF_STG = 0;                       // output staging aliases GEMM 1 stage 0 (dead once p_full completed)
constexpr int OFF_BAR = OFF_G + GS * G_STAGE;
constexpr int STG_BYTES = EPI_WARPS * STG_FLOATS * 4;
// sub-batches >= GS land in the GEMM 1 stages as well (behind the output staging) once GEMM 1 has consumed them
constexpr int OFF_G_LATE = OFF_STG + STG_BYTES;
static_assert(STG_BYTES % 1024 == 0 && OFF_G_LATE + (NSUB - GS) * G_STAGE <= STAGES * STAGE, "late transform slots must fit");
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 512;
constexpr int TMEM_COLS = 512;
constexpr int COL_M = 3 * NF;                    // 288
constexpr int COL_W = COL_M + 2 * SUBN;          // 480
// power-of-two operand scales (exact; undone in the epilogue); W_EXP / G_EXP live in skin_common.cuh
constexpr int COEF_EXP = 6, W_EXP = skin::W_EXP, G_EXP = skin::G_EXP;
static_assert(COL_W + 2 * (KJ / 2) == TMEM_COLS, "TMEM column map");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params
{
  int V, B, Bpad, ntiles;
  float scale_p, scale_m;       // 2^-(basis_exp + COEF_EXP), 2^-(W_EXP + G_EXP)
  const float * basis;          // (3 Vpad, 224): column 217 = template
  const float * weights;        // (V, 24) dense
  const float * wsum;           // (Vpad)
  const float * theta;          // (B, 25, 3): row 0 = root translation
  float * out;                  // (B, V, 3)
  long long * dbg;              // optional per-CTA phase timestamps (SMPLPP_TC2_DBG)
};
} // namespace tc2


// ------------------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------------------
// basis (3 Vpad, 224) fp32 -> [part][tile][plane][128][224] fp16, scaled by 2^e
__global__ void split_basis_f16_kernel(const float * __restrict__ basis, int V, int ntiles, float scale, __half * __restrict__ dst)
{
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_part = static_cast<long long>(ntiles) * 3 * tc2::MV * tc2::KP;
  if(i >= per_part) return;
  const int k = static_cast<int>(i % tc2::KP);
  const long long row = i / tc2::KP;
  const int r = static_cast<int>(row % tc2::MV);
  const int plane = static_cast<int>((row / tc2::MV) % 3);
  const int tile = static_cast<int>(row / (3 * tc2::MV));
  const int v = tile * tc2::MV + r;
  const float x = (v < V && k < tc2::KUSED) ? basis[(static_cast<size_t>(3) * v + plane) * kBlendK + k] * scale : 0.f;
  const __half hi = __float2half_rn(x);
  dst[i] = hi;
  dst[per_part + i] = __float2half_rn(x - __half2float(hi));
}

// per-call operands: coef (B,224) fp32 -> coef16 [part][Bpad][224]; xforms (B,24,12) fp32 -> xf16 [part][Bpad*12][32]
// (row = frame * 12 + element of the 3x4, column = joint: K-major B operand of GEMM 2).  Padding is zero-filled.
__global__ void split_frame_operands_kernel(const float * __restrict__ coef, const float * __restrict__ xforms, int B, int Bpad,
                                            __half * __restrict__ coef16, __half * __restrict__ xf16)
{
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n_coef = static_cast<long long>(Bpad) * tc2::KP;
  const long long n_xf = static_cast<long long>(Bpad) * kXformFloats * tc2::KJ;
  if(i < n_coef)
  {
    const int k = static_cast<int>(i % tc2::KP);
    const long long f = i / tc2::KP;
    const float x = (f < B && k < tc2::KUSED) ? coef[i] * static_cast<float>(1 << tc2::COEF_EXP) : 0.f;
    const __half hi = __float2half_rn(x);
    coef16[i] = hi;
    coef16[n_coef + i] = __float2half_rn(x - __half2float(hi));
  }
  else if(i < n_coef + n_xf)
  {
    const long long o = i - n_coef;
    const int j = static_cast<int>(o % tc2::KJ);
    const long long row = o / tc2::KJ;
    const int e = static_cast<int>(row % kXformFloats);
    const long long f = row / kXformFloats;
    const float x = (f < B && j < kJoints) ? xforms[(f * kJoints + j) * kXformFloats + e] * static_cast<float>(1 << tc2::G_EXP) : 0.f;
    const __half hi = __float2half_rn(x);
    xf16[o] = hi;
    xf16[n_xf + o] = __float2half_rn(x - __half2float(hi));
  }
}

// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(tc2::THREADS, 1)
    blend_skin_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                          const __grid_constant__ CUtensorMap tmG, const tc2::Params p)
{
  using namespace tc2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float * stg = reinterpret_cast<float *>(smem + OFF_STG);
  uint64_t * bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  uint64_t * full = bars;              // [STAGES]  TMA -> MMA (GEMM 1 stages)
  uint64_t * empty = full + STAGES;    // [STAGES]  MMA -> TMA
  uint64_t * g_full = empty + STAGES;  // [NSUB]    TMA -> MMA (transform sub-batches, one-shot)
  uint64_t * m_full = g_full + NSUB;   // [2]       MMA -> epilogue (skinning matrices of a sub-batch)
  uint64_t * m_empty = m_full + 2;     // [2]       epilogue -> MMA
  uint64_t * p_full = m_empty + 2;     //           MMA -> epilogue (rest accumulators complete)
  uint64_t * w_ready = p_full + 1;     //           epilogue -> MMA (W tile stored in TMEM)
  uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(w_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.y;              // frame tiles fastest: CTAs in flight share the (3x larger) basis tile
  const int f0 = blockIdx.x * NF;
  const int nsub = (min(NF, p.B - f0) + SUBF - 1) / SUBF;

  if(warp == 0 && lane == 0)
  {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    ptx::prefetch_tensormap(&tmG);
    for(int s = 0; s < STAGES; s++)
    {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for(int s = 0; s < NSUB; s++) ptx::mbar_init(&g_full[s], 1);
    for(int i = 0; i < 2; i++)
    {
      ptx::mbar_init(&m_full[i], 1);
      ptx::mbar_init(&m_empty[i], EPI_WARPS);
    }
    ptx::mbar_init(p_full, 1);
    ptx::mbar_init(w_ready, 4);
    ptx::fence_barrier_init();
  }
  if(warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if(warp == 0)
  {
    if(ptx::elect_one())
    {
      auto load_g = [&](int sb) {
        ptx::mbar_expect_tx(&g_full[sb], G_STAGE);
        uint8_t * dst = smem + (sb < GS ? OFF_G + sb * G_STAGE : OFF_G_LATE + (sb - GS) * G_STAGE);
        const int row = (f0 + sb * SUBF) * kXformFloats;
#pragma unroll
        for(int part = 0; part < 2; part++)
          ptx::tma_load_2d(dst + part * G_PART, &tmG, 0, part * p.Bpad * kXformFloats + row, &g_full[sb]);
      };
      for(int sb = 0; sb < min(GS, nsub); sb++) load_g(sb);
      for(int kb = 0; kb < NKB; kb++)
      {
        const int s = kb % STAGES;
        ptx::mbar_wait(&empty[s], ((kb / STAGES) & 1) ^ 1);
        ptx::mbar_expect_tx(&full[s], STAGE);
        uint8_t * dst = smem + s * STAGE;
#pragma unroll
        for(int part = 0; part < 2; part++)
        {
          const int row = (part * p.ntiles + tile) * (3 * MV);
          ptx::tma_load_2d(dst + part * A_PART, &tmA, kb * 32, row, &full[s]);
          ptx::tma_load_2d(dst + part * A_PART + (3 * MV / 2) * ROWB, &tmA, kb * 32, row + 3 * MV / 2, &full[s]);
        }
#pragma unroll
        for(int part = 0; part < 2; part++)
          ptx::tma_load_2d(dst + 2 * A_PART + part * B_PART, &tmB, kb * 32, part * p.Bpad + f0, &full[s]);
      }
      // GEMM 1 has consumed its stages (last use of stage s = its ceil((NKB - s) / STAGES)-th): the remaining transform
      // sub-batches of the tile land there, all in flight at once
      for(int s = 0; s < STAGES; s++) ptx::mbar_wait(&empty[s], (((NKB - s + STAGES - 1) / STAGES) - 1) & 1);
      for(int sb = GS; sb < nsub; sb++) load_g(sb);
    }
  }
  else if(warp == 1)
  {
    if(ptx::elect_one())
    {
      // ---- GEMM 1: rest accumulators, three products per K step ----
      constexpr uint32_t idesc1 = ptx::make_idesc_f16(MV, NF);
      long long * dbg = p.dbg ? p.dbg + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 256 : nullptr;
      if(dbg) dbg[0] = clock64();
      for(int kb = 0; kb < NKB; kb++)
      {
        const int s = kb % STAGES;
        ptx::mbar_wait(&full[s], (kb / STAGES) & 1);
        if(dbg) dbg[1 + kb] = clock64();
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + s * STAGE);
#pragma unroll
        for(int prod = 0; prod < 3; prod++)
        {
          const int pa = prod == 1 ? 1 : 0, pb = prod == 2 ? 1 : 0; // hi.hi, lo.hi, hi.lo
#pragma unroll
          for(int ks = 0; ks < 2; ks++)
          {
            const uint64_t bdesc = ptx::make_smem_desc<ROWB>(sa + 2 * A_PART + pb * B_PART + ks * 32);
#pragma unroll
            for(int c = 0; c < 3; c++)
            {
              const uint64_t adesc = ptx::make_smem_desc<ROWB>(sa + pa * A_PART + c * MV * ROWB + ks * 32);
              ptx::umma_f16_ss(tmem_base + c * NF, adesc, bdesc, idesc1, (kb | prod | ks) != 0 ? 1u : 0u);
            }
          }
        }
        ptx::tc_commit(&empty[s]);
      }
      ptx::tc_commit(p_full);
      if(dbg) dbg[8] = clock64();
      // ---- GEMM 2: skinning matrices of 4 frames at a time, A = W tile in TMEM ----
      constexpr uint32_t idesc2 = ptx::make_idesc_f16(MV, SUBN);
      ptx::mbar_wait(w_ready, 0);
      ptx::tc_fence_after();
      if(dbg) dbg[9] = clock64();
      for(int sb = 0; sb < nsub; sb++)
      {
        const int b = sb & 1;
        ptx::mbar_wait(&g_full[sb], 0);
        if(dbg && sb < 32) dbg[192 + sb] = clock64();
        ptx::mbar_wait(&m_empty[b], ((sb >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        if(dbg && sb < 32) dbg[64 + sb] = clock64();
        const uint32_t sg = ptx::smem_u32(smem + (sb < GS ? OFF_G + sb * G_STAGE : OFF_G_LATE + (sb - GS) * G_STAGE));
#pragma unroll
        for(int prod = 0; prod < 3; prod++)
        {
          const int pa = prod == 1 ? 1 : 0, pb = prod == 2 ? 1 : 0;
#pragma unroll
          for(int ks = 0; ks < 2; ks++)
          {
            const uint64_t bdesc = ptx::make_smem_desc<ROWB>(sg + pb * G_PART + ks * 32);
            ptx::umma_f16_ts(tmem_base + COL_M + b * SUBN, tmem_base + COL_W + pa * (KJ / 2) + ks * 8, bdesc, idesc2,
                             (prod | ks) != 0 ? 1u : 0u);
          }
        }
        ptx::tc_commit(&m_full[b]);
        if(dbg && sb < 40) dbg[10 + sb] = clock64();
      }
    }
  }
  else
  {
    const int ew = warp - 2;
    const int q = warp & 3;        // TMEM lane quadrant this warp may access (hardware rule: warp id % 4)
    const int fp = ew >> 2;        // which frame pair of the 8-frame sub-batch (every warp visits every sub-batch)
    const int wv0 = tile * MV + q * 32;
    const int v = wv0 + lane;
    const int vc = min(v, p.V - 1);
    const int nvalid = max(0, min(32, p.V - wv0));
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if(ew < 4)
    {
      skin::store_w_row_tmem(p.weights + static_cast<size_t>(vc) * kJoints, lane_taddr + COL_W);
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if(lane == 0) ptx::mbar_arrive(w_ready);
    }
    float T[3];
#pragma unroll
    for(int k = 0; k < 3; k++) T[k] = __ldg(p.basis + (static_cast<size_t>(3) * vc + k) * kBlendK + KUSED);
    const float sm = p.scale_m / p.wsum[vc]; // homogeneous divide (LinearBlendSkinning.cpp:545-550) folded into the scale
    const float sp = p.scale_p;
    const uint32_t my_stg = ptx::smem_u32(stg + ew * STG_FLOATS);

    ptx::mbar_wait(p_full, 0);
    ptx::tc_fence_after();
    long long * dbg = (p.dbg && ew == 0 && lane == 0) ? p.dbg + (static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 256 : nullptr;
    if(dbg) dbg[50] = clock64();
    for(int sb = 0; sb < nsub; sb++)
    {
      const int h = sb & 1; // matrix buffer of this sub-batch
      const int fl = sb * SUBF + fp * EPI_FR; // first of this warp's frames within the tile
      // root translation (theta row 0, SMPL.cpp:726-727): uniform loads, in flight during the barrier wait
      float tr[EPI_FR][3];
#pragma unroll
      for(int t = 0; t < EPI_FR; t++)
      {
        const float * trp = p.theta + static_cast<size_t>(min(f0 + fl + t, p.B - 1)) * ((kJoints + 1) * 3);
        tr[t][0] = __ldg(trp), tr[t][1] = __ldg(trp + 1), tr[t][2] = __ldg(trp + 2);
      }
      ptx::mbar_wait(&m_full[h], (sb >> 1) & 1);
      ptx::tc_fence_after();
      if(dbg && sb < 32) dbg[96 + sb] = clock64();
      float M[EPI_FR * kXformFloats], X[EPI_FR], Y[EPI_FR], Z[EPI_FR];
      const uint32_t mcol = lane_taddr + COL_M + h * SUBN + fp * (EPI_FR * kXformFloats);
      ptx::tmem_ld_x16(mcol, M);
      ptx::tmem_ld_x8p(mcol + 16, M + 16);
      ptx::tmem_ld_x2(lane_taddr + 0 * NF + fl, X);
      ptx::tmem_ld_x2(lane_taddr + 1 * NF + fl, Y);
      ptx::tmem_ld_x2(lane_taddr + 2 * NF + fl, Z);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if(lane == 0) ptx::mbar_arrive(&m_empty[h]); // the MMA warp may overwrite this matrix buffer
      if(p.dbg && lane == 0 && (sb == 8 || sb == 9))
        p.dbg[(static_cast<size_t>(blockIdx.y) * gridDim.x + blockIdx.x) * 256 + 224 + ew] = clock64();
      if(dbg && sb < 32) dbg[128 + sb] = clock64();
#pragma unroll
      for(int t = 0; t < EPI_FR; t++)
      {
        const float rx = fmaf(X[t], sp, T[0]), ry = fmaf(Y[t], sp, T[1]), rz = fmaf(Z[t], sp, T[2]);
        const float * m = M + kXformFloats * t;
        const float ox = fmaf(m[0], rx, fmaf(m[1], ry, fmaf(m[2], rz, m[3])));
        const float oy = fmaf(m[4], rx, fmaf(m[5], ry, fmaf(m[6], rz, m[7])));
        const float oz = fmaf(m[8], rx, fmaf(m[9], ry, fmaf(m[10], rz, m[11])));
        const uint32_t sa = my_stg + (t * 96 + lane * 3) * 4;
        ptx::sts32(sa, fmaf(ox, sm, tr[t][0]));
        ptx::sts32(sa + 4, fmaf(oy, sm, tr[t][1]));
        ptx::sts32(sa + 8, fmaf(oz, sm, tr[t][2]));
      }
      __syncwarp();
      // each frame's 32 vertices are 384 contiguous bytes: 8-byte coalesced streaming stores
#pragma unroll
      for(int i = 0; i < EPI_FR * 48 / 32; i++)
      {
        const int idx = 32 * i + lane; // float2 index over EPI_FR frames x 48
        const int t = idx / 48;
        const int w2 = idx - 48 * t;
        const int f = f0 + fl + t;
        const float2 val = ptx::lds64(my_stg + idx * 8);
        if(f < p.B && 2 * w2 < 3 * nvalid)
          __stcs(reinterpret_cast<float2 *>(p.out + (static_cast<size_t>(f) * p.V + wv0) * 3) + w2, val);
      }
      __syncwarp();
      if(dbg && sb < 32) dbg[160 + sb] = clock64();
    }
    if(dbg) dbg[51] = clock64();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if(warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
namespace
{
using EncodeTiledFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                   const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void * p = nullptr;
    cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
    if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess
       && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  });
  return fn;
}

// 2-D K-major fp16 tensor map: rows x cols elements, box = 32 elements (64 B) x box_rows, SWIZZLE_64B
bool encode_f16(CUtensorMap * out, void * base, uint64_t rows, uint64_t cols, uint32_t box_rows)
{
  EncodeTiledFn fn = encode_fn();
  if(!fn) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {tc2::ROWB / 2, box_rows};
  cuuint32_t estr[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
         == CUDA_SUCCESS;
}
} // namespace

namespace sb
{
bool tc_encode_f16(void * out_map, void * base, uint64_t rows, uint64_t cols, uint32_t box_rows)
{
  return encode_f16(static_cast<CUtensorMap *>(out_map), base, rows, cols, box_rows);
}

size_t tc2_frame_operand_bytes(int64_t batch)
{
  const size_t bpad = align_up(static_cast<size_t>(batch), tc2::NF);
  return align_up(2 * bpad * tc2::KP * sizeof(__half)) + align_up(2 * bpad * kXformFloats * tc2::KJ * sizeof(__half));
}

// builds the fp16 split basis and its tensor map; called once from smplpp_model_create after tc_prepare_model
int tc2_prepare_model(ModelDev & d, float basis_max_abs)
{
  d.tc2_ready = false;
  if(!tc_blend_available() || (d.V & 1) || d.V < tc2::MV || !encode_fn()) return SMPLPP_OK;
  const int ntiles = (d.V + tc2::MV - 1) / tc2::MV;
  // largest power of two that keeps |basis| 2^e <= 1024: lo = x - hi stays a normal fp16 for all but tiny entries
  int e = 0;
  if(basis_max_abs > 0.f)
  {
    while(basis_max_abs * ldexpf(1.f, e + 1) <= 1024.f && e < 24) e++;
    while(basis_max_abs * ldexpf(1.f, e) > 1024.f && e > -24) e--;
  }
  d.tc2_basis_exp = e;
  const long long per_part = static_cast<long long>(ntiles) * 3 * tc2::MV * tc2::KP;
  SB_CUDA(cudaMalloc(&d.basis_f16, static_cast<size_t>(2) * per_part * sizeof(__half)));
  split_basis_f16_kernel<<<static_cast<unsigned>((per_part + 255) / 256), 256>>>(d.basis, d.V, ntiles, ldexpf(1.f, e),
                                                                                static_cast<__half *>(d.basis_f16));
  SB_LAUNCHED();
  if(!encode_f16(reinterpret_cast<CUtensorMap *>(d.tmapA16), d.basis_f16, static_cast<uint64_t>(2) * ntiles * 3 * tc2::MV,
                 tc2::KP, 3 * tc2::MV / 2))
    return fail(SMPLPP_ERR_CUDA, "CUDA", "cuTensorMapEncodeTiled failed for the fp16 blend basis");
  SB_CUDA(cudaDeviceSynchronize());
  SB_CUDA(cudaFuncSetAttribute(blend_skin_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, tc2::SMEM_BYTES));
  d.tc2_tiles = ntiles;
  d.tc2_ready = true;
  return SMPLPP_OK;
}

void tc2_release_model(ModelDev & d)
{
  if(d.basis_f16) cudaFree(d.basis_f16);
  d.basis_f16 = nullptr;
  d.tc2_ready = false;
}

// coef (B,224) and xforms (B,24,12) fp32 from K1; scratch: tc2_frame_operand_bytes(B)
int launch_blend_skin_tc2(const ModelDev & d, cudaStream_t st, int B, const float * coef, const float * xforms, void * scratch,
                          const float * theta, float * out)
{
  if(!d.tc2_ready) return fail(SMPLPP_ERR_INVALID, "SMPL", "tcgen05 skinning variant is not available for this model");
  if(reinterpret_cast<uintptr_t>(out) & 7) return fail(SMPLPP_ERR_INVALID, "SMPL", "tcgen05 variants need 8-byte aligned vertices");
  const int Bpad = static_cast<int>(align_up(static_cast<size_t>(B), tc2::NF));
  __half * coef16 = static_cast<__half *>(scratch);
  __half * xf16 = reinterpret_cast<__half *>(static_cast<char *>(scratch) + align_up(static_cast<size_t>(2) * Bpad * tc2::KP * sizeof(__half)));
  const long long n = static_cast<long long>(Bpad) * (tc2::KP + kXformFloats * tc2::KJ);
  split_frame_operands_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(coef, xforms, B, Bpad, coef16, xf16);
  SB_LAUNCHED();
  alignas(64) CUtensorMap tmB, tmG;
  if(!encode_f16(&tmB, coef16, static_cast<uint64_t>(2) * Bpad, tc2::KP, tc2::NF)
     || !encode_f16(&tmG, xf16, static_cast<uint64_t>(2) * Bpad * kXformFloats, tc2::KJ, tc2::SUBN))
    return fail(SMPLPP_ERR_CUDA, "CUDA", "cuTensorMapEncodeTiled failed for the frame operands");
  tc2::Params p;
  p.V = d.V;
  p.B = B;
  p.Bpad = Bpad;
  p.ntiles = d.tc2_tiles;
  p.scale_p = ldexpf(1.f, -(d.tc2_basis_exp + tc2::COEF_EXP));
  p.scale_m = ldexpf(1.f, -(tc2::W_EXP + tc2::G_EXP));
  p.basis = d.basis;
  p.weights = d.weights_dense;
  p.wsum = d.lbs_wsum;
  p.theta = theta;
  p.out = out;
  const dim3 grid(Bpad / tc2::NF, d.tc2_tiles);
  static const bool dbg_on = getenv("SMPLPP_TC2_DBG") != nullptr;
  p.dbg = nullptr;
  if(dbg_on) SB_CUDA(cudaMalloc(&p.dbg, static_cast<size_t>(grid.x) * grid.y * 256 * sizeof(long long)));
  const CUtensorMap & tmA = *reinterpret_cast<const CUtensorMap *>(d.tmapA16);
  blend_skin_tc2_kernel<<<grid, tc2::THREADS, tc2::SMEM_BYTES, st>>>(tmA, tmB, tmG, p);
  SB_LAUNCHED();
  if(dbg_on)
  {
    SB_CUDA(cudaStreamSynchronize(st));
    std::vector<long long> h(static_cast<size_t>(grid.x) * grid.y * 256);
    SB_CUDA(cudaMemcpy(h.data(), p.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    for(size_t cta : {size_t(200), h.size() / 256 / 2})
    {
      const long long * t = h.data() + cta * 256;
      fprintf(stderr, "[tc2 dbg] cta %zu: first stage +%lld | k-blocks", cta, t[1] - t[0]);
      for(int kb = 1; kb < 7; kb++) fprintf(stderr, " %lld", t[1 + kb] - t[kb]);
      fprintf(stderr, " | gemm1 issue done +%lld | w_ready +%lld | p_full seen by epilogue +%lld | sub-batches", t[8] - t[0],
              t[9] - t[0], t[50] - t[0]);
      for(int sb = 0; sb < 32; sb += 4) fprintf(stderr, " %lld", t[10 + sb] - t[0]);
      fprintf(stderr, " | last %lld | epilogue end +%lld\n", t[10 + 31] - t[0], t[51] - t[0]);
      for(int sb = 8; sb < 16; sb++)
        fprintf(stderr, "   sb %2d: mma waits done +%lld, issued +%lld%s", sb, t[64 + sb] - t[0], t[10 + sb] - t[0], (sb & 1) ? "\n" : " ||");
      for(int sb = 8; sb < 14; sb++) fprintf(stderr, "   sb %2d: g_full done +%lld, m_empty done +%lld\n", sb, t[192 + sb] - t[0], t[64 + sb] - t[0]);
      fprintf(stderr, "   m_empty arrivals of sb 8/9 per epilogue warp:");
      for(int w = 0; w < 16; w++) fprintf(stderr, " %lld", t[224 + w] - t[0]);
      fprintf(stderr, "\n");
      for(int sb = 8; sb < 16; sb += 2)
        fprintf(stderr, "   epi(buf0) sb %2d: m_full seen +%lld, ld done+arrive +%lld, loop end +%lld\n", sb, t[96 + sb] - t[0],
                t[128 + sb] - t[0], t[160 + sb] - t[0]);
    }
  }
  return SMPLPP_OK;
}
} // namespace sb
