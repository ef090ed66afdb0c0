// K2''' (tcgen05, persistent + pipelined): fused blend-shape contraction + linear blend skinning, both products on the
// tensor cores, with the contraction of work item i+1 running UNDER the skinning GEMM + epilogue of work item i.
//
// Reference semantics: BlendShape::poseBlend / shapeBlend (src/BlendShape.cpp:764, 670-683), the rest shape
// T + S + P (src/JointRegression.cpp:551-565) and LinearBlendSkinning::skinning (src/LinearBlendSkinning.cpp:445-553).
//
//   rest[v,f]   = T[v] + sum_k basis[v,k] coef[f,k]                 GEMM 1: (128 vertices x 3 planes) x 96 frames, K = 224
//   M[v,f]      = sum_j W[v,j] G'[f,j]  (3x4 per vertex and frame)  GEMM 2: 128 vertices x (8 frames x 12), K = 24 -> 32
//   vert[v,f]   = M[v,f] [rest; 1] / sum_j W[v,j] + trans[f]        epilogue: 12 FMA per vertex and frame
//
// What changed against K2'' (skin_tc.cu), whose CTA ran wait(3.2 k) -> GEMM 1 (8.6 k) -> GEMM 2 + epilogue (11 k cycles)
// back to back with the tensor pipe idle for two thirds of the time:
//   * one persistent CTA per SM walks a contiguous range of (vertex tile, frame block) work items;
//   * when GEMM 1 of an item completes, the sixteen epilogue warps DRAIN its 288 accumulator columns into registers
//     (72 per thread, pre-scaled by 1 / sum_j W), which frees the TMEM columns: one thread then issues the 126 MMAs of
//     the NEXT item's GEMM 1 while another issues the 12 sub-batches of this item's GEMM 2;
//   * two TMA producer threads, one per ring (2 x 60 KB GEMM 1 stages, 6 x 12 KB transform sub-batches), so neither
//     ring can starve the other;
//   * the frames of a 96-frame block are permuted in the fp16 coefficient operand so that the 24 frames an epilogue
//     warp owns are 24 adjacent accumulator columns (6 tcgen05.ld per drain instead of 36).
//   * every operand tile is stored in global memory as the exact (SWIZZLE_64B) shared-memory image of its pipeline
//     stage, so a stage is two contiguous cp.async.bulk copies (48 KB basis + 12 KB coefficients) and a transform
//     sub-batch is one (12 KB): no tensor maps, a handful of large copies per item (DESIGN.md 4.1 has the measurements).
// Precision is that of K2'': fp16 hi + lo split of every operand, hi.hi + lo.hi + hi.lo in fp32 TMEM.
//
// TMEM (512 columns): [0,288) rest accumulators (x | y | z planes x 96 frames), [288,480) two 96-column buffers of
// skinning matrices (8 frames x 12), [480,512) the W tile (A operand of GEMM 2, fp16 hi | lo).
// warp 0: TMA producer of the GEMM 1 stages | warp 1: TMEM allocator + issuer of GEMM 2 | warp 2: TMA producer of the transform
// sub-batches | warp 3: issuer of GEMM 1 (the first warpgroup hands its registers back with setmaxnreg) | warps 4-19: epilogue,
// 112 registers each (four per TMEM lane quadrant, one per frame pair of an 8-frame sub-batch).
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#include "forward.cuh"
#include "skin_common.cuh"
#include "tc3_layout.cuh"
#include "tc_ptx.cuh"

using namespace sb;

namespace tc3
{
struct Params
{
  int V, B, Bpad, ntiles, nfb, nitems;
  float scale_p, scale_m;       // 2^-(basis_exp + COEF_EXP), 2^-(W_EXP + G_EXP)
  const uint8_t * img_a;        // [tile][K-block][part][plane][128][64 B]   stage image of the fp16 basis
  const uint8_t * img_b;        // [frame block][K-block][part][96][64 B]   stage image of the fp16 coefficients
  const uint8_t * img_g;        // [frame block][sub-batch][part][96][64 B] stage image of the fp16 transforms
  const float * basis;          // (3 Vpad, 224): column 217 = template
  const float * weights;        // (V, 24) dense
  const float * wsum;           // (Vpad)
  const float * theta;          // (B, 25, 3): row 0 = root translation
  float * out;                  // (B, V, 3)
  long long * dbg;              // optional per-CTA timestamps (SMPLPP_TC3_DBG)
  int dbg_mode;                 // DBG build only (SMPLPP_TC3_DBG = bit mask): 1 no matrix tcgen05.ld, 2 no stores, 4 no GEMM 2 MMAs, 8 no GEMM 1 MMAs, 16 no drain loads, 32 no stage ring, 64 no transform ring, 128 no epilogue arithmetic, 256 no matrix-buffer hand-off, 512 no drain hand-off (results are wrong: timing decomposition)
};

} // namespace tc3

// basis (3 Vpad, 224) fp32 -> stage images [tile][K-block][part hi | lo][plane][128 rows][32 fp16], scaled by 2^e
__global__ void basis_image_f16_kernel(const float * __restrict__ basis, int V, int ntiles, float scale, uint8_t * __restrict__ img)
{
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long n = static_cast<long long>(ntiles) * 3 * tc3::MV * tc3::KP;
  if(i >= n) return;
  const int k = static_cast<int>(i % tc3::KP);
  const long long row = i / tc3::KP;
  const int r = static_cast<int>(row % tc3::MV);
  const int plane = static_cast<int>((row / tc3::MV) % 3);
  const int tile = static_cast<int>(row / (3 * tc3::MV));
  const int v = tile * tc3::MV + r;
  const float x = (v < V && k < tc3::KUSED) ? basis[(static_cast<size_t>(3) * v + plane) * kBlendK + k] * scale : 0.f;
  const __half hi = __float2half_rn(x);
  const __half lo = __float2half_rn(x - __half2float(hi));
  uint8_t * blk = img + (static_cast<size_t>(tile) * tc3::NKB + k / 32) * (2 * tc3::A_PART);
  const uint32_t o = static_cast<uint32_t>(plane * tc3::MV * tc3::ROWB + r * tc3::ROWB + (k % 32) * 2);
  *reinterpret_cast<__half *>(blk + tc3::swz64(o)) = hi;
  *reinterpret_cast<__half *>(blk + tc3::swz64(tc3::A_PART + o)) = lo;
}

// per-call operands: coef (B,224) fp32 -> img_b (rows of every 96-frame block permuted by coef_row, x 2^6);
// xforms (B,24,12) fp32 -> img_g (row = frame in sub-batch * 12 + element of the 3x4, column = joint, x 2^4).
// Padding (frames >= B, K >= 217, joints >= 24) is zero-filled.  One thread per 16-byte chunk (8 fp16 of one row:
// the swizzle moves whole chunks), hi and lo written as one 16-byte store each.
__global__ void frame_images3_kernel(const float * __restrict__ coef, const float * __restrict__ xforms, int B, int Bpad,
                                     uint8_t * __restrict__ img_b, uint8_t * __restrict__ img_g)
{
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  constexpr int CK = tc3::KP / 8;   // 28 chunks per coefficient row
  constexpr int CJ = tc3::KJ / 8;   // 4 chunks per transform row
  const long long n_coef = static_cast<long long>(Bpad) * CK;
  const long long n_xf = static_cast<long long>(Bpad) * kXformFloats * CJ;
  float x[8];
  if(i < n_coef)
  {
    const int ck = static_cast<int>(i % CK);
    const long long f = i / CK;
    if(f < B)
    {
      const float4 a = __ldg(reinterpret_cast<const float4 *>(coef + f * tc3::KP + ck * 8));
      const float4 b = __ldg(reinterpret_cast<const float4 *>(coef + f * tc3::KP + ck * 8 + 4));
      x[0] = a.x, x[1] = a.y, x[2] = a.z, x[3] = a.w, x[4] = b.x, x[5] = b.y, x[6] = b.z, x[7] = b.w;
    }
#pragma unroll
    for(int e = 0; e < 8; e++) x[e] = (f < B && ck * 8 + e < tc3::KUSED) ? x[e] * static_cast<float>(1 << tc3::COEF_EXP) : 0.f;
    const long long fb = f / tc3::NF;
    const int r = tc3::coef_row(static_cast<int>(f - fb * tc3::NF));
    uint8_t * blk = img_b + (static_cast<size_t>(fb) * tc3::NKB + ck / 4) * (2 * tc3::B_PART);
    const uint32_t o = static_cast<uint32_t>(r * tc3::ROWB + (ck % 4) * 16);
    tc3::split8_store(x, blk + tc3::swz64(o), blk + tc3::swz64(tc3::B_PART + o));
  }
  else if(i < n_coef + n_xf)
  {
    const long long q = i - n_coef;
    const int cj = static_cast<int>(q % CJ);
    const long long row = q / CJ;
    const int e = static_cast<int>(row % kXformFloats);
    const long long f = row / kXformFloats;
#pragma unroll
    for(int jj = 0; jj < 8; jj++)
    {
      const int j = cj * 8 + jj;
      x[jj] = (f < B && j < kJoints) ? __ldg(xforms + (f * kJoints + j) * kXformFloats + e) * static_cast<float>(1 << tc3::G_EXP) : 0.f;
    }
    const long long fb = f / tc3::NF;
    const int nf = static_cast<int>(f - fb * tc3::NF);
    uint8_t * blk = img_g + (static_cast<size_t>(fb) * tc3::NSUB + nf / tc3::SUBF) * tc3::G_STAGE;
    const uint32_t o = static_cast<uint32_t>(((nf % tc3::SUBF) * kXformFloats + e) * tc3::ROWB + cj * 16);
    tc3::split8_store(x, blk + tc3::swz64(o), blk + tc3::swz64(tc3::G_PART + o));
  }
}

template<int STAGES, int GSLOTS, int EPI, bool DBG = false>
__global__ void __launch_bounds__(tc3::THREADS, 1)
    blend_skin_tc3_kernel(const tc3::Params p)
{
  using namespace tc3;
  using L = Layout<STAGES, GSLOTS, EPI>;
  constexpr int OFF_G = L::OFF_G, OFF_STG = L::OFF_STG, OFF_BAR = L::OFF_BAR;
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t * bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  // the same base as a 32-bit shared-window address, aligned in that space (the window base is itself 1 KB aligned): the
  // hot loops address stages and barriers as smem32 + constant instead of converting a generic pointer at every use
  const uint32_t smem32 = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar32 = smem32 + OFF_BAR;
  uint64_t * full = bars;                // [STAGES]  TMA -> MMA (GEMM 1 stages)
  uint64_t * empty = full + STAGES;      // [STAGES]  MMA -> TMA
  uint64_t * g_full = empty + STAGES;    // [GSLOTS]  TMA -> MMA (transform sub-batches)
  uint64_t * g_empty = g_full + GSLOTS;  // [GSLOTS]  MMA -> TMA
  uint64_t * m_full = g_empty + GSLOTS;  // [2]       MMA -> epilogue (skinning matrices of a sub-batch)
  uint64_t * m_empty = m_full + 2;       // [2]       epilogue -> MMA
  uint64_t * p_full = m_empty + 2;       //           MMA -> epilogue (rest accumulators of an item complete)
  uint64_t * rest_free = p_full + 1;     //           epilogue -> MMA (accumulators drained into registers)
  uint64_t * w_ready = rest_free + 1;    //           epilogue -> MMA (W tile stored in TMEM)
  uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(w_ready + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous range of work items, vertex tile major / frame block minor: a CTA changes its vertex tile (W tile,
  // template, 1 / sum w) at most a couple of times
  const int it0 = static_cast<int>(static_cast<long long>(blockIdx.x) * p.nitems / gridDim.x);
  const int it1 = static_cast<int>(static_cast<long long>(blockIdx.x + 1) * p.nitems / gridDim.x);
  const int nit = it1 - it0;

  if(warp == 0 && lane == 0)
  {
    for(int s = 0; s < STAGES; s++)
    {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    for(int s = 0; s < GSLOTS; s++)
    {
      ptx::mbar_init(&g_full[s], 1);
      ptx::mbar_init(&g_empty[s], 1);
    }
    for(int i = 0; i < 2; i++)
    {
      ptx::mbar_init(&m_full[i], 1);
      ptx::mbar_init(&m_empty[i], EPI_WARPS);
    }
    ptx::mbar_init(p_full, 1);
    ptx::mbar_init(rest_free, EPI_WARPS);
    ptx::mbar_init(w_ready, 4);
    ptx::fence_barrier_init();
  }
  if(warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  long long * dbg = (DBG && p.dbg) ? p.dbg + static_cast<size_t>(blockIdx.x) * 256 : nullptr;
  // DBG build: cycles the MMA thread / one epilogue warp spend blocked in each kind of wait (SMPLPP_TC3_DBG)
  long long w_stage = 0, w_gfull = 0, w_mempty = 0, w_rest = 0, w_pfull = 0, w_mfull = 0;
#define TC3_OFF(bit) (DBG && (p.dbg_mode & (bit)))
#define TC3_TIMED(acc, stmt) do { if constexpr(DBG) { const long long t_ = clock64(); stmt; acc += clock64() - t_; } else { stmt; } } while(0)

  if(warp < CTRL_WARPS)
  {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
  if(warp == 0)
  {
    // ---- producer of the GEMM 1 stage ring ----
    // (Measured, scripts/ubench/ubench_tma.cu: [expect_tx, copy] groups issued back to back by ONE thread onto different
    // barriers retire ~410 cycles apart whatever their size, so a group is as large as a stage, and each ring has its own
    // producer thread with plain blocking waits.  A single polling producer over both rings performed the same.)
    if(ptx::elect_one())
    {
      const int total_kb = TC3_OFF(32) ? 0 : nit * NKB;
      // (An L2 prefetch of all seven K-blocks of the first item at kernel start was measured: no gain.)
      // The reload of a stage is on the critical path of the GEMM 1 issuer (a quarter of its samples are stage waits):
      // source addresses (two integer divisions) are ready BEFORE the wait, barriers and stages are 32-bit shared addresses.
      int kb = 0, tile = it0 / p.nfb, fb = it0 % p.nfb;
      for(int kbn = 0; kbn < total_kb; kbn++)
      {
        const uint32_t s = static_cast<uint32_t>(kbn % STAGES);
        const uint8_t * srca = p.img_a + (static_cast<size_t>(tile) * NKB + kb) * (2 * A_PART);
        const uint8_t * srcb = p.img_b + (static_cast<size_t>(fb) * NKB + kb) * (2 * B_PART);
        const uint32_t dst = smem32 + s * STAGE, fbar = bar32 + 8 * s;
        ptx::mbar_wait_a(bar32 + 8 * (STAGES + s), ((kbn / STAGES) & 1) ^ 1);
        ptx::mbar_expect_tx_a(fbar, STAGE);
        ptx::bulk_load_1d_a(dst, srca, 2 * A_PART, fbar);
        ptx::bulk_load_1d_a(dst + 2 * A_PART, srcb, 2 * B_PART, fbar);
        if(dbg && kbn < 14) dbg[34 + kbn] = clock64();
        if(++kb == NKB)
        {
          kb = 0;
          if(++fb == p.nfb) fb = 0, tile++;
        }
      }
    }
  }
  else if(warp == 2)
  {
    // ---- producer of the transform sub-batch ring ----
    if(ptx::elect_one())
    {
      const int total_g = TC3_OFF(64) ? 0 : nit * NSUB;
      int sb = 0, fb = it0 % p.nfb;
      for(int gn = 0; gn < total_g; gn++)
      {
        const uint32_t s = static_cast<uint32_t>(gn % GSLOTS);
        const uint8_t * src = p.img_g + (static_cast<size_t>(fb) * NSUB + sb) * G_STAGE;
        const uint32_t fbar = bar32 + 8 * (2 * STAGES + s);
        ptx::mbar_wait_a(bar32 + 8 * (2 * STAGES + GSLOTS + s), ((gn / GSLOTS) & 1) ^ 1);
        ptx::mbar_expect_tx_a(fbar, G_STAGE);
        ptx::bulk_load_1d_a(smem32 + OFF_G + s * G_STAGE, src, G_STAGE, fbar);
        if(++sb == NSUB)
        {
          sb = 0;
          if(++fb == p.nfb) fb = 0;
        }
      }
    }
  }
  else if(warp == 1 || warp == 3)
  {
    if(ptx::elect_one())
    {
      constexpr uint32_t idesc1 = ptx::make_idesc_f16(MV, NF);
      constexpr uint32_t idesc2 = ptx::make_idesc_f16(MV, SUBN);
      // Every instruction between two tcgen05.mma is on the critical path of its issuing thread.  ncu's source page of the
      // single-issuer version (r02h) had that thread at 1800 instructions per item for 198 MMAs and blocked on a full
      // tensor queue in only a quarter of its samples: the tensor pipe starved behind the issue loop.  Hence (a) TWO
      // issuing threads - warp 3 the 126 GEMM 1 MMAs of the next item, warp 1 the 72 GEMM 2 MMAs of this one; the tensor
      // queue interleaves them (0.1725 -> 0.1607 ms per launch) - and (b) lean streams: the sub-batches and half K-blocks
      // are unrolled, so that ring slots, parities and operand offsets are compile-time constants; only the stage of a
      // K-block (ring position over all items) and the item parity are run-time values, and barriers are 32-bit shared
      // addresses.  Descriptor = constant high word | (shared address >> 4).
      static_assert(NSUB % GSLOTS == 0 && ((NSUB / GSLOTS) & 1) == 0, "transform ring: slot and parity of a sub-batch do not depend on the item");
      static_assert((NSUB & 3) == 0, "matrix buffers: parity of a sub-batch does not depend on the item");
      constexpr uint32_t DHI = ptx::smem_desc_hi<ROWB>();
      const uint32_t smem16 = smem32 >> 4;
      const uint32_t a_full = bar32; // one base register, every other barrier at a compile-time offset
      const uint32_t a_empty = a_full + 8 * STAGES, a_gfull = a_empty + 8 * STAGES, a_gempty = a_gfull + 8 * GSLOTS,
                     a_mfull = a_gempty + 8 * GSLOTS, a_mempty = a_mfull + 16, a_pfull = a_mempty + 16, a_rest = a_pfull + 8,
                     a_wready = a_rest + 8;
      // MMA i (0..8) of a half K-block: triple i / 3 (operand parts and K step), plane i % 3.
      //   half 0: (hi.hi, ks 0), (hi.hi, ks 1), (lo.hi, ks 0)      half 1: (lo.hi, ks 1), (hi.lo, ks 0), (hi.lo, ks 1)
      auto g1_mma = [&](uint32_t st16, int hpar, int i, uint32_t acc) {
        const int t = i / 3, c = i % 3;
        const int a_part = hpar == 0 ? (t == 2 ? 1 : 0) : (t == 0 ? 1 : 0);
        const int b_part = hpar == 0 ? 0 : (t == 0 ? 0 : 1);
        const int ks = hpar == 0 ? (t == 1 ? 1 : 0) : (t == 1 ? 0 : 1);
        ptx::umma_f16_ss_lo(tmem_base + c * NF, st16 + ((a_part * A_PART + c * MV * ROWB + ks * 32) >> 4),
                            st16 + ((2 * A_PART + b_part * B_PART + ks * 32) >> 4), DHI, idesc1, acc);
      };
      // GEMM 2 MMA j (0..5) of a sub-batch: hi.hi, lo.hi, hi.lo with two K = 16 steps each
      auto g2_mma = [&](uint32_t dm, uint32_t aw, uint32_t sg16, int j) {
        const int prod = j >> 1, ks = j & 1;
        const int pa = prod == 1 ? 1 : 0, pb = prod == 2 ? 1 : 0;
        ptx::umma_f16_ts_lo(dm, aw + pa * (KJ / 2) + ks * 8, sg16 + ((pb * G_PART + ks * 32) >> 4), DHI, idesc2, j != 0 ? 1u : 0u);
      };
      // half K-block h (compile-time after unrolling: K-block h / 2, operand half h % 2) of the GEMM 1 whose first K-block
      // has ring position kbn0 (run-time).  The even half waits for the TMA data, the odd half hands the stage back.
      auto g1_half = [&](int h, int kbn0) {
        const int kbn = kbn0 + (h >> 1);
        const uint32_t s = static_cast<uint32_t>(kbn % STAGES);
        const uint32_t st16 = smem16 + s * (STAGE >> 4);
        if((h & 1) == 0)
        {
          if(!TC3_OFF(32)) TC3_TIMED(w_stage, ptx::mbar_wait_a(a_full + 8 * s, (kbn / STAGES) & 1));
          ptx::tc_fence_after();
          if(dbg && kbn < 14) dbg[48 + kbn] = clock64();
#pragma unroll
          for(int i = 0; i < 9; i++)
            if(!DBG || !(p.dbg_mode & 8)) g1_mma(st16, 0, i, (h == 0 && i < 3) ? 0u : 1u);
        }
        else
        {
#pragma unroll
          for(int i = 0; i < 9; i++)
            if(!DBG || !(p.dbg_mode & 8)) g1_mma(st16, 1, i, 1u);
          if(!TC3_OFF(32)) ptx::tc_commit_a(a_empty + 8 * s);
        }
      };
      if(dbg) dbg[0] = clock64();
      if(nit > 0 && warp == 3)
      {
#pragma unroll
        for(int h = 0; h < 2 * NKB; h++) g1_half(h, 0);
        if(!TC3_OFF(512)) ptx::tc_commit_a(a_pfull);
      }
      if(dbg) dbg[1] = clock64();
      if(warp == 3)
      {
        // ---- GEMM 1 issuer: the 126 MMAs of item k + 1 as soon as item k's accumulators have been drained into the epilogue
        //      warps' registers; the stage ring and the tensor queue pace it ----
        for(int k = 0; k + 1 < nit; k++)
        {
          if(!TC3_OFF(512)) TC3_TIMED(w_rest, ptx::mbar_wait_a(a_rest, k & 1));
          ptx::tc_fence_after();
#pragma unroll
          for(int h = 0; h < 2 * NKB; h++) g1_half(h, (k + 1) * NKB);
          if(!TC3_OFF(512)) ptx::tc_commit_a(a_pfull);
        }
        if(dbg) dbg[64] = w_stage, dbg[67] = w_rest, dbg[69] = clock64();
      }
      // ---- GEMM 2 issuer (warp 1): the twelve sub-batches of every item, two matrix buffers ----
      int wcount = 0, prev_tile = -1;
      for(int k = 0; k < (warp == 3 ? 0 : nit); k++)
      {
        const int tile = (it0 + k) / p.nfb;
        if(tile != prev_tile)
        {
          ptx::mbar_wait_a(a_wready, wcount & 1);
          ptx::tc_fence_after();
          wcount++;
          prev_tile = tile;
        }
        const uint32_t aw = tmem_base + COL_W;
#pragma unroll
        for(int sb = 0; sb < NSUB; sb++)
        {
          const int b = sb & 1, gs = sb % GSLOTS;
          if(!TC3_OFF(64)) TC3_TIMED(w_gfull, ptx::mbar_wait_a(a_gfull + 8 * gs, (sb / GSLOTS) & 1));
          if(!TC3_OFF(256)) TC3_TIMED(w_mempty, ptx::mbar_wait_a(a_mempty + 8 * b, ((sb >> 1) & 1) ^ 1));
          ptx::tc_fence_after();
          const uint32_t sg16 = smem16 + ((OFF_G + gs * G_STAGE) >> 4);
          const uint32_t dm = tmem_base + COL_M + b * SUBN;
#pragma unroll
          for(int j = 0; j < 6; j++)
            if(!DBG || !(p.dbg_mode & 4)) g2_mma(dm, aw, sg16, j);
          if(!TC3_OFF(256)) ptx::tc_commit_a(a_mfull + 8 * b);
          if(!TC3_OFF(64)) ptx::tc_commit_a(a_gempty + 8 * gs);
        }
        if(dbg && k < 30) dbg[2 + k] = clock64();
      }
      if(dbg && warp == 1) dbg[65] = w_gfull, dbg[66] = w_mempty, dbg[68] = clock64();
    }
  }
  }
  else
  {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 112;");
    const int ew = warp - CTRL_WARPS;
    const int q = warp & 3;        // TMEM lane quadrant this warp may access (hardware rule: warp id % 4)
    const int fp = ew >> 2;        // which frame pair of every 8-frame sub-batch
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    const uint32_t my_stg = ptx::smem_u32(smem + OFF_STG) + ew * STG_FLOATS * 4;
    const float sp = p.scale_p;
    // barriers as 32-bit shared addresses, frame stride and validity bounds hoisted: the r02h source page showed ~120
    // instructions per sub-batch and warp for 24 FMA + 6 stores (cvta sequences, 64-bit store addresses rebuilt from the
    // kernel parameters for every frame)
    const uint32_t a_mfull = bar32 + 16 * (STAGES + GSLOTS); // one base register, compile-time offsets
    const uint32_t a_mempty = a_mfull + 16, a_pfull = a_mempty + 16, a_rest = a_pfull + 8, a_wready = a_rest + 8;
    const unsigned fstride = static_cast<unsigned>(p.V) * 3u; // floats per frame of the output
    // (At 112 registers ptxas rebuilds part of this state in the first sub-batches instead of keeping it; forcing it into
    // registers with opaque asm operands gave 72 instead of 75-120 instructions per sub-batch and was 2 % SLOWER.)
    int prev_tile = -1;
    int wv0 = 0, nvalid = 0;
    float T0 = 0.f, T1 = 0.f, T2 = 0.f, sm = 0.f;
    if(dbg && ew == 0 && lane == 0) dbg[32] = clock64();
    for(int k = 0; k < nit; k++)
    {
      const int item = it0 + k;
      const int tile = item / p.nfb, f0 = (item - tile * p.nfb) * NF;
      if(tile != prev_tile)
      {
        prev_tile = tile;
        wv0 = tile * MV + q * 32;
        const int vc = min(wv0 + lane, p.V - 1);
        nvalid = max(0, min(32, p.V - wv0));
        if(ew < 4)
        {
          // every MMA that read the previous W tile has completed: this warp has seen m_full of the item's last sub-batch
          skin::store_w_row_tmem(p.weights + static_cast<size_t>(vc) * kJoints, lane_taddr + COL_W);
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          __syncwarp();
          if(lane == 0) ptx::mbar_arrive_a(a_wready);
        }
        sm = p.scale_m / p.wsum[vc]; // homogeneous divide (LinearBlendSkinning.cpp:545-550) folded into the scale
        const float * tp = p.basis + static_cast<size_t>(3) * vc * kBlendK + KUSED;
        T0 = __ldg(tp) * sm, T1 = __ldg(tp + kBlendK) * sm, T2 = __ldg(tp + 2 * kBlendK) * sm;
      }
      // root translations (theta row 0, SMPL.cpp:726-727) of this warp's 24 frames: lane l holds frame pair l / 2, frame
      // l % 2; fetched once per item (in flight during the p_full wait) and broadcast by shuffles in the sub-batch loop
      float trx = 0.f, try_ = 0.f, trz = 0.f;
      if(lane < FR_WARP)
      {
        const int f = min(f0 + (lane >> 1) * SUBF + fp * EPI_FR + (lane & 1), p.B - 1);
        const float * tq = p.theta + static_cast<size_t>(f) * ((kJoints + 1) * 3);
        trx = __ldg(tq), try_ = __ldg(tq + 1), trz = __ldg(tq + 2);
      }
      // ---- drain the rest accumulators of this item: 24 frames x (x, y, z), pre-multiplied by scale / sum w ----
      float R[3][FR_WARP];
      if(!TC3_OFF(512)) TC3_TIMED(w_pfull, ptx::mbar_wait_a(a_pfull, k & 1));
      ptx::tc_fence_after();
#pragma unroll
      for(int c = 0; c < 3; c++)
      {
        if(DBG && (p.dbg_mode & 16))
        {
#pragma unroll
          for(int i = 0; i < FR_WARP; i++) R[c][i] = 1.f;
          continue;
        }
        ptx::tmem_ld_x16(lane_taddr + c * NF + fp * FR_WARP, R[c]);
        ptx::tmem_ld_x8p(lane_taddr + c * NF + fp * FR_WARP + 16, R[c] + 16);
      }
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if(lane == 0 && !TC3_OFF(512)) ptx::mbar_arrive_a(a_rest);
      {
        const float s = sp * sm;
#pragma unroll
        for(int i = 0; i < FR_WARP; i++)
        {
          R[0][i] = fmaf(R[0][i], s, T0);
          R[1][i] = fmaf(R[1][i], s, T1);
          R[2][i] = fmaf(R[2][i], s, T2);
        }
      }
      float * outp = p.out + (static_cast<size_t>(f0 + fp * EPI_FR) * p.V + wv0) * 3;
      float * outl = outp + lane * 3;                                  // this lane's vertex in the warp's first frame
      const int nfv = lane < nvalid ? p.B - f0 - fp * EPI_FR : 0;      // frame slots (sb * SUBF + t) below nfv are stored
#pragma unroll
      for(int sb = 0; sb < NSUB; sb++)
      {
        const int h = sb & 1; // matrix buffer of this sub-batch (NSUB is even: the same for every item)
        float tr[EPI_FR][3];
#pragma unroll
        for(int t = 0; t < EPI_FR; t++)
        {
          tr[t][0] = __shfl_sync(0xffffffffu, trx, sb * EPI_FR + t);
          tr[t][1] = __shfl_sync(0xffffffffu, try_, sb * EPI_FR + t);
          tr[t][2] = __shfl_sync(0xffffffffu, trz, sb * EPI_FR + t);
        }
        if(!TC3_OFF(256)) TC3_TIMED(w_mfull, ptx::mbar_wait_a(a_mfull + 8 * h, (sb >> 1) & 1)); // 6 ring rounds per item: the parity does not depend on the item
        ptx::tc_fence_after();
        float M[EPI_FR * kXformFloats];
        const uint32_t mcol = lane_taddr + COL_M + h * SUBN + fp * (EPI_FR * kXformFloats);
        if(DBG && (p.dbg_mode & 1))
        {
#pragma unroll
          for(int i = 0; i < EPI_FR * kXformFloats; i++) M[i] = 1.f + i;
        }
        else
        {
          ptx::tmem_ld_x16(mcol, M);
          ptx::tmem_ld_x8p(mcol + 16, M + 16);
        }
        ptx::tmem_ld_wait();
        ptx::tc_fence_before();
        __syncwarp();
        if(lane == 0 && !TC3_OFF(256)) ptx::mbar_arrive_a(a_mempty + 8 * h); // the MMA warp may overwrite this matrix buffer
        if(TC3_OFF(128)) continue;
#pragma unroll
        for(int t = 0; t < EPI_FR; t++)
        {
          const float rx = R[0][sb * EPI_FR + t], ry = R[1][sb * EPI_FR + t], rz = R[2][sb * EPI_FR + t];
          const float * m = M + kXformFloats * t;
          const float ox = fmaf(m[0], rx, fmaf(m[1], ry, fmaf(m[2], rz, fmaf(m[3], sm, tr[t][0]))));
          const float oy = fmaf(m[4], rx, fmaf(m[5], ry, fmaf(m[6], rz, fmaf(m[7], sm, tr[t][1]))));
          const float oz = fmaf(m[8], rx, fmaf(m[9], ry, fmaf(m[10], rz, fmaf(m[11], sm, tr[t][2]))));
          if constexpr(EPI == 0)
          {
            const uint32_t sa = my_stg + (t * 96 + lane * 3) * 4;
            ptx::sts32(sa, ox);
            ptx::sts32(sa + 4, oy);
            ptx::sts32(sa + 8, oz);
          }
          else
          {
            // no staging: three 4-byte streaming stores per vertex and frame (a warp's 32 vertices are 384 contiguous
            // bytes; L2 merges the sectors).  Measured +2.4 % against the staged float2 stores: the kernel is bound by
            // shared-memory bandwidth (UMMA operand reads + TMA writes), which the staging round trip competes for.
            float * o = outl + static_cast<size_t>((sb * SUBF + t) * fstride);
            if(sb * SUBF + t < nfv && (!DBG || !(p.dbg_mode & 2) || ox == 1.2345e30f))
            {
              __stcs(o, ox);
              __stcs(o + 1, oy);
              __stcs(o + 2, oz);
            }
          }
        }
        if constexpr(EPI != 0) continue;
        __syncwarp();
        // each frame's 32 vertices are 384 contiguous bytes: 8-byte coalesced streaming stores
#pragma unroll
        for(int i = 0; i < EPI_FR * 48 / 32; i++)
        {
          const int idx = 32 * i + lane; // float2 index over EPI_FR frames x 48
          const int t = idx / 48;
          const int w2 = idx - 48 * t;
          const int f = f0 + sb * SUBF + fp * EPI_FR + t;
          const float2 val = ptx::lds64(my_stg + idx * 8);
          if(f < p.B && 2 * w2 < 3 * nvalid)
            __stcs(reinterpret_cast<float2 *>(outp + (static_cast<size_t>(sb * SUBF + t) * p.V) * 3) + w2, val);
        }
        __syncwarp();
      }
    }
    if(dbg && (ew == 0 || ew == 15) && lane == 0)
    {
      if(ew == 0) dbg[33] = clock64();
      dbg[70 + (ew ? 2 : 0)] = w_pfull, dbg[71 + (ew ? 2 : 0)] = w_mfull;
    }
  }
#undef TC3_TIMED
#undef TC3_OFF
  ptx::tc_fence_before();
  __syncthreads();
  if(warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
namespace sb
{
size_t tc3_frame_operand_bytes(int64_t batch)
{
  const size_t bpad = align_up(static_cast<size_t>(batch), tc3::NF);
  return align_up(2 * bpad * tc3::KP * sizeof(__half)) + align_up(2 * bpad * kXformFloats * tc3::KJ * sizeof(__half));
}

// builds the stage images of the fp16 split basis (scaled like the tc2 copy); called once from smplpp_model_create
int tc3_prepare_model(ModelDev & d)
{
  d.tc3_ready = false;
  if(!d.tc2_ready) return SMPLPP_OK;
  const long long n = static_cast<long long>(d.tc2_tiles) * 3 * tc3::MV * tc3::KP;
  SB_CUDA(cudaMalloc(&d.basis_img16, static_cast<size_t>(2) * n * sizeof(__half)));
  basis_image_f16_kernel<<<static_cast<unsigned>((n + 255) / 256), 256>>>(d.basis, d.V, d.tc2_tiles, ldexpf(1.f, d.tc2_basis_exp),
                                                                         static_cast<uint8_t *>(d.basis_img16));
  SB_LAUNCHED();
  SB_CUDA(cudaDeviceSynchronize());
  SB_CUDA(cudaFuncSetAttribute(blend_skin_tc3_kernel<2, 6, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::Layout<2, 6, 1>::SMEM_BYTES));
  SB_CUDA(cudaFuncSetAttribute(blend_skin_tc3_kernel<2, 6, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::Layout<2, 6, 1>::SMEM_BYTES));
  SB_CUDA(cudaFuncSetAttribute(blend_skin_tc3_kernel<3, 3, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::Layout<3, 3, 1>::SMEM_BYTES));
  SB_CUDA(cudaFuncSetAttribute(blend_skin_tc3_kernel<2, 6, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc3::Layout<2, 6, 0>::SMEM_BYTES));
  int dev = 0;
  SB_CUDA(cudaGetDevice(&dev));
  SB_CUDA(cudaDeviceGetAttribute(&d.sm_count, cudaDevAttrMultiProcessorCount, dev));
  d.tc3_ready = d.sm_count > 0;
  return SMPLPP_OK;
}

void tc3_release_model(ModelDev & d)
{
  if(d.basis_img16) cudaFree(d.basis_img16);
  d.basis_img16 = nullptr;
  d.tc3_ready = false;
}

// coef (B,224) and xforms (B,24,12) fp32 from K1; scratch: tc3_frame_operand_bytes(B)
int launch_blend_skin_tc3(const ModelDev & d, cudaStream_t st, int B, const float * coef, const float * xforms, void * scratch,
                          const float * theta, float * out, bool images_ready)
{
  if(!d.tc3_ready) return fail(SMPLPP_ERR_INVALID, "SMPL", "pipelined tcgen05 skinning variant is not available for this model");
  if(reinterpret_cast<uintptr_t>(out) & 7) return fail(SMPLPP_ERR_INVALID, "SMPL", "tcgen05 variants need 8-byte aligned vertices");
  if(reinterpret_cast<uintptr_t>(scratch) & 127) return fail(SMPLPP_ERR_INVALID, "SMPL", "tcgen05 variants need a 128-byte aligned workspace");
  const int Bpad = static_cast<int>(align_up(static_cast<size_t>(B), tc3::NF));
  uint8_t * img_b = static_cast<uint8_t *>(scratch);
  uint8_t * img_g = img_b + tc3::img_g_offset(Bpad);
  if(!images_ready)
  {
    const long long n = static_cast<long long>(Bpad) * (tc3::KP + kXformFloats * tc3::KJ) / 8; // 16-byte chunks
    frame_images3_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(coef, xforms, B, Bpad, img_b, img_g);
    SB_LAUNCHED();
  }
  tc3::Params p;
  p.V = d.V;
  p.B = B;
  p.Bpad = Bpad;
  p.ntiles = d.tc2_tiles;
  p.nfb = Bpad / tc3::NF;
  p.nitems = p.ntiles * p.nfb;
  p.scale_p = ldexpf(1.f, -(d.tc2_basis_exp + tc3::COEF_EXP));
  p.scale_m = ldexpf(1.f, -(tc3::W_EXP + tc3::G_EXP));
  p.img_a = static_cast<const uint8_t *>(d.basis_img16);
  p.img_b = img_b;
  p.img_g = img_g;
  p.basis = d.basis;
  p.weights = d.weights_dense;
  p.wsum = d.lbs_wsum;
  p.theta = theta;
  p.out = out;
  int grid = p.nitems < d.sm_count ? p.nitems : d.sm_count;
  static const int grid_env = getenv("SMPLPP_TC3_GRID") ? atoi(getenv("SMPLPP_TC3_GRID")) : 0;
  static const int ring_env = getenv("SMPLPP_TC3_RING") ? atoi(getenv("SMPLPP_TC3_RING")) : 0;
  if(grid_env > 0 && grid_env < grid) grid = grid_env;
  static const bool dbg_on = getenv("SMPLPP_TC3_DBG") != nullptr;
  p.dbg = nullptr;
  p.dbg_mode = dbg_on ? atoi(getenv("SMPLPP_TC3_DBG")) >> 1 : 0; // SMPLPP_TC3_DBG = 1 + 2 * mode mask
  if(dbg_on)
  {
    SB_CUDA(cudaMalloc(&p.dbg, static_cast<size_t>(grid) * 256 * sizeof(long long)));
    SB_CUDA(cudaMemsetAsync(p.dbg, 0, static_cast<size_t>(grid) * 256 * sizeof(long long), st));
  }
  if(ring_env == 1) // SMPLPP_TC3_RING: alternatives kept for measurement (3 stages + 3 transform slots; staged stores)
    blend_skin_tc3_kernel<3, 3, 1><<<grid, tc3::THREADS, tc3::Layout<3, 3, 1>::SMEM_BYTES, st>>>(p);
  else if(ring_env == 2)
    blend_skin_tc3_kernel<2, 6, 0><<<grid, tc3::THREADS, tc3::Layout<2, 6, 0>::SMEM_BYTES, st>>>(p);
  else if(dbg_on)
    blend_skin_tc3_kernel<2, 6, 1, true><<<grid, tc3::THREADS, tc3::Layout<2, 6, 1>::SMEM_BYTES, st>>>(p);
  else
    blend_skin_tc3_kernel<2, 6, 1><<<grid, tc3::THREADS, tc3::Layout<2, 6, 1>::SMEM_BYTES, st>>>(p);
  SB_LAUNCHED();
  if(dbg_on)
  {
    SB_CUDA(cudaStreamSynchronize(st));
    std::vector<long long> h(static_cast<size_t>(grid) * 256);
    SB_CUDA(cudaMemcpy(h.data(), p.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
    cudaFree(p.dbg);
    for(int cta : {0, grid / 2})
    {
      const long long * t = h.data() + static_cast<size_t>(cta) * 256;
      fprintf(stderr, "[tc3 dbg] cta %d: first GEMM 1 issued +%lld | items", cta, t[1] - t[0]);
      for(int k = 0; k < 30 && t[2 + k]; k++) fprintf(stderr, " %lld", t[2 + k] - (k ? t[1 + k] : t[1]));
      fprintf(stderr, " | epilogue %lld .. %lld\n", t[32] - t[0], t[33] - t[0]);
      fprintf(stderr, "   stage loads issued (K-block 0..13):");
      for(int i = 0; i < 14; i++) fprintf(stderr, " %lld", t[34 + i] - t[0]);
      fprintf(stderr, "\n   stage seen full by the MMA thread:  ");
      for(int i = 0; i < 14; i++) fprintf(stderr, " %lld", t[48 + i] - t[0]);
      fprintf(stderr, "\n   GEMM 1 issuer: total %lld cycles, blocked on stage %lld, drain %lld | GEMM 2 issuer: total %lld, blocked on transforms %lld, matrix buffer %lld | epilogue warp 0 / 15 blocked on rest accumulators %lld / %lld, on matrices %lld / %lld\n",
              t[69] - t[0], t[64], t[67], t[68] - t[0], t[65], t[66], t[70], t[72], t[71], t[73]);
    }
  }
  return SMPLPP_OK;
}
} // namespace sb
