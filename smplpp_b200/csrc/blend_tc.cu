// K2' (tcgen05): fused pose/shape blend contraction + linear blend skinning on the 5th-generation tensor cores.
//
// Reference semantics: BlendShape::poseBlend / shapeBlend (src/BlendShape.cpp:764, 670-683), the rest shape
// T + S + P (src/JointRegression.cpp:551-565) and LinearBlendSkinning::skinning (src/LinearBlendSkinning.cpp:445-553).
//
// The contraction is fp32 in the reference.  A single TF32 / BF16 pass misses the 1e-5 m bound, so every operand is
// split into hi + lo (both exactly representable in the tensor-core input format) and three products are
// accumulated in fp32 TMEM:  D = A_hi B_hi + A_lo B_hi + A_hi B_lo.  The ~1 m template term stays OUT of the
// tensor-core sum (added in fp32 in the epilogue) so that accumulator rounding acts on cm-scale offsets only.
//
// Mapping (one CTA = 128 vertices x 128 frames, K = 224 in 64-byte K-blocks, SWIZZLE_64B, 2-stage TMA ring):
//   UMMA M = 128 vertices (TMEM lanes), N = 128 frames (TMEM columns); three accumulators = x / y / z planes of the
//   basis, so the thread that owns TMEM lane v sees all three coordinates of vertex v for every frame of the tile and
//   keeps that vertex's <= 4 skinning weights in registers.  The frames' 24 x (3x4) transforms stream through a
//   double-buffered shared-memory window (32 frames per bulk copy).  Stores go through a per-warp staging buffer so
//   that each warp writes 384 contiguous bytes per frame with 8-byte vector stores.
//   warp 0: TMA producer | warp 1: TMEM allocator + MMA issuer | warps 2-9: epilogue (2 warps per TMEM lane quadrant)
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>

#include "forward.cuh"
#include "tc_ptx.cuh"

using namespace sb;

namespace tc
{
constexpr int MV = 128;                            // vertices per tile (UMMA M)
constexpr int NF = 128;                            // frames per tile (UMMA N)
constexpr int ROWB = 64;                           // bytes of K per shared-memory row (one SWIZZLE_64B span)
constexpr int STAGES = 2;
constexpr int A_PART = 3 * MV * ROWB;              // one part (hi or lo) of the basis tile: 3 planes x 128 rows
constexpr int B_PART = NF * ROWB;                  // one part of the coefficient tile
constexpr int STAGE = 2 * A_PART + 2 * B_PART;     // 65536
constexpr int XF_FR = 32;                          // frames per transform window
constexpr int XF_FLOATS = XF_FR * kJoints * kXformFloats;
constexpr int XF_BYTES = XF_FLOATS * 4;            // 36864
constexpr int EPI_WARPS = 8;
constexpr int THREADS = 32 * (2 + EPI_WARPS);
constexpr int STG_FR = 4;                          // frames staged per warp between two warp syncs
constexpr int STG_FLOATS = STG_FR * 32 * 3;        // per warp: 4 frames x 32 vertices x 3
constexpr int OFF_XF = STAGES * STAGE;
constexpr int OFF_STG = OFF_XF + 2 * XF_BYTES;
constexpr int OFF_BAR = OFF_STG + EPI_WARPS * STG_FLOATS * 4;
constexpr int SMEM_BYTES = 1024 + OFF_BAR + 256;
constexpr int TMEM_COLS = 512;
constexpr int KP = kBlendK;                        // padded K per part (224)
constexpr int KUSED = kPoseDim + kShapeDim;        // 217: the template column is excluded
static_assert(3 * NF <= TMEM_COLS, "three accumulators must fit TMEM");
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");

struct Params
{
  int V, B, Bpad, ntiles, nkb, ke, ell_stride, kmax, frames_fastest;
  const float * basis;
  const uint8_t * lbs_joint;
  const float * lbs_weight;
  const float * lbs_wsum;
  const float * xforms;
  const float * theta;
  float * out;
};
} // namespace tc

// ------------------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float round_tf32(float x)
{
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r & 0xFFFFE000u);
}

// basis (3 Vpad, 224) fp32 -> [part][tile][plane][128][224] in the tensor-core input format
template<bool kTf32>
__global__ void split_basis_kernel(const float * __restrict__ basis, int V, int ntiles, void * __restrict__ dst)
{
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long per_part = static_cast<long long>(ntiles) * 3 * tc::MV * tc::KP;
  if(i >= per_part) return;
  const int k = static_cast<int>(i % tc::KP);
  const long long row = i / tc::KP;
  const int r = static_cast<int>(row % tc::MV);
  const int plane = static_cast<int>((row / tc::MV) % 3);
  const int tile = static_cast<int>(row / (3 * tc::MV));
  const int v = tile * tc::MV + r;
  const float x = (v < V && k < tc::KUSED) ? basis[(static_cast<size_t>(3) * v + plane) * kBlendK + k] : 0.f;
  if constexpr(kTf32)
  {
    float * d = static_cast<float *>(dst);
    const float hi = round_tf32(x);
    d[i] = hi;
    d[per_part + i] = round_tf32(x - hi);
  }
  else
  {
    __nv_bfloat16 * d = static_cast<__nv_bfloat16 *>(dst);
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    d[i] = hi;
    d[per_part + i] = __float2bfloat16_rn(x - __bfloat162float(hi));
  }
}

// coefficient rows (B, 224) fp32 (as written by K1) -> [part][Bpad][224]
template<bool kTf32>
__global__ void split_coef_kernel(const float * __restrict__ coef, int B, int Bpad, void * __restrict__ dst)
{
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if(i >= static_cast<long long>(B) * tc::KP) return;
  const int k = static_cast<int>(i % tc::KP);
  const float x = k < tc::KUSED ? coef[i] : 0.f;
  const long long per_part = static_cast<long long>(Bpad) * tc::KP;
  if constexpr(kTf32)
  {
    float * d = static_cast<float *>(dst);
    const float hi = round_tf32(x);
    d[i] = hi;
    d[per_part + i] = round_tf32(x - hi);
  }
  else
  {
    __nv_bfloat16 * d = static_cast<__nv_bfloat16 *>(dst);
    const __nv_bfloat16 hi = __float2bfloat16_rn(x);
    d[i] = hi;
    d[per_part + i] = __float2bfloat16_rn(x - __bfloat162float(hi));
  }
}

// ------------------------------------------------------------------------------------------------------------
// the kernel
// ------------------------------------------------------------------------------------------------------------
// xfr: shared-memory byte address of the frame's 24 x (3x4) transforms; joff: byte offsets of the vertex's joints
__device__ __forceinline__ void skin_vertex(uint32_t xfr, const int (&joff)[4], const float (&jw)[4], float rx, float ry,
                                            float rz, float & ox, float & oy, float & oz)
{
  ox = oy = oz = 0.f;
#pragma unroll
  for(int k = 0; k < 4; k++)
  {
    const uint32_t g = xfr + joff[k];
    const float4 r0 = ptx::lds128(g), r1 = ptx::lds128(g + 16), r2 = ptx::lds128(g + 32);
    const float w = jw[k];
    ox = fmaf(w, fmaf(r0.x, rx, fmaf(r0.y, ry, fmaf(r0.z, rz, r0.w))), ox);
    oy = fmaf(w, fmaf(r1.x, rx, fmaf(r1.y, ry, fmaf(r1.z, rz, r1.w))), oy);
    oz = fmaf(w, fmaf(r2.x, rx, fmaf(r2.y, ry, fmaf(r2.z, rz, r2.w))), oz);
  }
}

template<bool kTf32>
__global__ void __launch_bounds__(tc::THREADS, 1)
    blend_skin_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                         const tc::Params p)
{
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t * smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float * xf = reinterpret_cast<float *>(smem + OFF_XF);
  float * stg = reinterpret_cast<float *>(smem + OFF_STG);
  uint64_t * bars = reinterpret_cast<uint64_t *>(smem + OFF_BAR);
  uint64_t * full = bars;                  // [STAGES]  TMA -> MMA
  uint64_t * empty = bars + STAGES;        // [STAGES]  MMA -> TMA
  uint64_t * tmem_full = bars + 2 * STAGES; // MMA -> epilogue
  uint64_t * xf_full = tmem_full + 1;      // [2] TMA -> epilogue
  uint64_t * xf_empty = xf_full + 2;       // [2] epilogue -> TMA
  uint32_t * tmem_slot = reinterpret_cast<uint32_t *>(xf_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // kFramesFastest: consecutive CTAs share the (3x larger) basis tile instead of the coefficient tile
  const int tile = p.frames_fastest ? blockIdx.y : blockIdx.x;
  const int f0 = (p.frames_fastest ? blockIdx.x : blockIdx.y) * NF;

  if(warp == 0 && lane == 0)
  {
    ptx::prefetch_tensormap(&tmA);
    ptx::prefetch_tensormap(&tmB);
    for(int s = 0; s < STAGES; s++)
    {
      ptx::mbar_init(&full[s], 1);
      ptx::mbar_init(&empty[s], 1);
    }
    ptx::mbar_init(tmem_full, 1);
    for(int i = 0; i < 2; i++)
    {
      ptx::mbar_init(&xf_full[i], 1);
      ptx::mbar_init(&xf_empty[i], EPI_WARPS);
    }
    ptx::fence_barrier_init();
  }
  if(warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int nsb = min(NF / XF_FR, (p.B - f0 + XF_FR - 1) / XF_FR); // transform windows holding live frames

  if(warp == 0)
  {
    if(ptx::elect_one())
    {
      auto load_xf = [&](int sb) {
        const int buf = sb & 1;
        ptx::mbar_wait(&xf_empty[buf], ((sb >> 1) & 1) ^ 1);
        ptx::mbar_expect_tx(&xf_full[buf], XF_BYTES);
        ptx::bulk_load_1d(xf + buf * XF_FLOATS, p.xforms + static_cast<size_t>(f0 + sb * XF_FR) * (kJoints * kXformFloats),
                          XF_BYTES, &xf_full[buf]);
      };
      for(int sb = 0; sb < min(2, nsb); sb++) load_xf(sb);
      for(int kb = 0; kb < p.nkb; kb++)
      {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        ptx::mbar_wait(&empty[s], ph ^ 1);
        ptx::mbar_expect_tx(&full[s], STAGE);
        uint8_t * dst = smem + s * STAGE;
#pragma unroll
        for(int part = 0; part < 2; part++)
        {
          const int row = (part * p.ntiles + tile) * (3 * MV);
          ptx::tma_load_2d(dst + part * A_PART, &tmA, kb * p.ke, row, &full[s]);
          ptx::tma_load_2d(dst + part * A_PART + (3 * MV / 2) * ROWB, &tmA, kb * p.ke, row + 3 * MV / 2, &full[s]);
        }
#pragma unroll
        for(int part = 0; part < 2; part++)
          ptx::tma_load_2d(dst + 2 * A_PART + part * B_PART, &tmB, kb * p.ke, part * p.Bpad + f0, &full[s]);
      }
      for(int sb = 2; sb < nsb; sb++) load_xf(sb);
    }
  }
  else if(warp == 1)
  {
    if(ptx::elect_one())
    {
      constexpr uint32_t idesc = ptx::make_idesc(kTf32, MV, NF);
      for(int kb = 0; kb < p.nkb; kb++)
      {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        ptx::mbar_wait(&full[s], ph);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + s * STAGE);
#pragma unroll
        for(int prod = 0; prod < 3; prod++)
        {
          // hi.hi, lo.hi, hi.lo
          const int pa = prod == 1 ? 1 : 0, pb = prod == 2 ? 1 : 0;
#pragma unroll
          for(int ks = 0; ks < 2; ks++)
          {
            const uint64_t bdesc = ptx::make_smem_desc<ROWB>(sa + 2 * A_PART + pb * B_PART + ks * 32);
#pragma unroll
            for(int c = 0; c < 3; c++)
            {
              const uint64_t adesc = ptx::make_smem_desc<ROWB>(sa + pa * A_PART + c * MV * ROWB + ks * 32);
              ptx::umma<kTf32>(tmem_base + c * NF, adesc, bdesc, idesc, (kb | prod | ks) != 0 ? 1u : 0u);
            }
          }
        }
        ptx::tc_commit(&empty[s]); // frees the stage once these MMAs have read it
      }
      ptx::tc_commit(tmem_full);
    }
  }
  else
  {
    const int ew = warp - 2;
    const int q = warp & 3; // TMEM lane quadrant this warp may read (hardware rule: warp id % 4)
    const int h = ew >> 2;  // which half of every window's 8-frame groups
    const int wv0 = tile * MV + q * 32;
    const int v = wv0 + lane;
    const int vc = min(v, p.V - 1);
    const int nvalid = max(0, min(32, p.V - wv0));
    float T[3];
#pragma unroll
    for(int k = 0; k < 3; k++) T[k] = __ldg(p.basis + (static_cast<size_t>(3) * vc + k) * kBlendK + KUSED);
    int joff[4];
    float jw[4];
#pragma unroll
    for(int k = 0; k < 4; k++)
    {
      const bool on = k < p.kmax;
      joff[k] = on ? p.lbs_joint[static_cast<size_t>(k) * p.ell_stride + vc] * (kXformFloats * 4) : 0;
      jw[k] = on ? p.lbs_weight[static_cast<size_t>(k) * p.ell_stride + vc] : 0.f;
    }
    const float iw = 1.f / p.lbs_wsum[vc];
    const uint32_t my_stg = ptx::smem_u32(stg + ew * STG_FLOATS);
    const uint32_t xf_addr = ptx::smem_u32(xf);
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    ptx::mbar_wait(tmem_full, 0);
    ptx::tc_fence_after();
    for(int sb = 0; sb < nsb; sb++)
    {
      // root translation (theta row 0, SMPL.cpp:726-727) of this window's frames: lane l holds frame sb * 32 + l;
      // issued before the wait so that the global-load latency hides behind it, broadcast per frame by shuffles
      const float * trp = p.theta + static_cast<size_t>(min(f0 + sb * XF_FR + lane, p.B - 1)) * ((kJoints + 1) * 3);
      const float wtx = __ldg(trp), wty = __ldg(trp + 1), wtz = __ldg(trp + 2);
      ptx::mbar_wait(&xf_full[sb & 1], (sb >> 1) & 1);
      const uint32_t xfb = xf_addr + (sb & 1) * XF_BYTES;
#pragma unroll 1
      for(int gg = 0; gg < 2; gg++)
      {
        const int g = h + 2 * gg;
        const int fl = sb * XF_FR + g * 8; // first frame of the group within the tile
        float X[8], Y[8], Z[8];
        ptx::tmem_ld_x8(lane_taddr + 0 * NF + fl, X);
        ptx::tmem_ld_x8(lane_taddr + 1 * NF + fl, Y);
        ptx::tmem_ld_x8(lane_taddr + 2 * NF + fl, Z);
        ptx::tmem_ld_wait();
#pragma unroll
        for(int hf = 0; hf < 8 / STG_FR; hf++)
        {
#pragma unroll
          for(int t = 0; t < STG_FR; t++)
          {
            const int fi = STG_FR * hf + t;
            const float trx = __shfl_sync(0xffffffffu, wtx, g * 8 + fi);
            const float try_ = __shfl_sync(0xffffffffu, wty, g * 8 + fi);
            const float trz = __shfl_sync(0xffffffffu, wtz, g * 8 + fi);
            float ox, oy, oz;
            skin_vertex(xfb + (g * 8 + fi) * (kJoints * kXformFloats * 4), joff, jw, X[fi] + T[0], Y[fi] + T[1],
                        Z[fi] + T[2], ox, oy, oz);
            const uint32_t sa = my_stg + (t * 96 + lane * 3) * 4;
            ptx::sts32(sa, fmaf(ox, iw, trx));
            ptx::sts32(sa + 4, fmaf(oy, iw, try_));
            ptx::sts32(sa + 8, fmaf(oz, iw, trz));
          }
          __syncwarp();
#pragma unroll
          for(int i = 0; i < STG_FR * 48 / 32; i++)
          {
            const int idx = 32 * i + lane; // float2 index over STG_FR frames x 48
            const int t = idx / 48;
            const int w = idx - 48 * t;
            const int f = f0 + fl + STG_FR * hf + t;
            const float2 val = ptx::lds64(my_stg + idx * 8);
            if(f < p.B && 2 * w < 3 * nvalid)
              __stcs(reinterpret_cast<float2 *>(p.out + (static_cast<size_t>(f) * p.V + wv0) * 3) + w, val);
          }
          __syncwarp();
        }
      }
      __syncwarp();
      if(lane == 0) ptx::mbar_arrive(&xf_empty[sb & 1]);
    }
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

// 2-D K-major tensor map: rows x 224 elements, box = (64 bytes of K) x box_rows, SWIZZLE_64B
bool encode_kmajor(CUtensorMap * out, bool tf32, void * base, uint64_t rows, uint32_t box_rows)
{
  EncodeTiledFn fn = encode_fn();
  if(!fn) return false;
  const uint32_t esize = tf32 ? 4 : 2;
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(tc::KP), rows};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(tc::KP) * esize};
  cuuint32_t box[2] = {tc::ROWB / esize, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, tf32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}
} // namespace

namespace sb
{
std::atomic<int> g_tc_grid_order{1};

bool tc_blend_available()
{
  static int ok = -1;
  if(ok < 0)
  {
    int dev = 0, major = 0;
    ok = 0;
    if(cudaGetDevice(&dev) == cudaSuccess
       && cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) == cudaSuccess && major == 10
       && encode_fn() != nullptr)
      ok = 1;
    cudaGetLastError();
  }
  return ok == 1;
}

bool tc_model_ready(const ModelDev & d)
{
  return d.tc_ready;
}

size_t tc_coef_split_bytes(int64_t batch)
{
  const size_t bpad = align_up(static_cast<size_t>(batch), tc::NF);
  return align_up(2 * bpad * tc::KP * sizeof(float));
}

// builds the split basis (both input formats) and its tensor maps; called once from smplpp_model_create
int tc_prepare_model(ModelDev & d)
{
  d.tc_ready = false;
  if(!tc_blend_available() || (d.V & 1) || d.V < tc::MV || d.kmax > 4) return SMPLPP_OK;
  const int ntiles = (d.V + tc::MV - 1) / tc::MV;
  d.tc_tiles = ntiles;
  const long long per_part = static_cast<long long>(ntiles) * 3 * tc::MV * tc::KP;
  for(int kind = 0; kind < 2; kind++)
  {
    const bool tf32 = kind == 1;
    const size_t bytes = static_cast<size_t>(2) * per_part * (tf32 ? 4 : 2);
    SB_CUDA(cudaMalloc(&d.basis_split[kind], bytes));
    const unsigned grid = static_cast<unsigned>((per_part + 255) / 256);
    if(tf32)
      split_basis_kernel<true><<<grid, 256>>>(d.basis, d.V, ntiles, d.basis_split[kind]);
    else
      split_basis_kernel<false><<<grid, 256>>>(d.basis, d.V, ntiles, d.basis_split[kind]);
    SB_LAUNCHED();
    if(!encode_kmajor(reinterpret_cast<CUtensorMap *>(d.tmapA[kind]), tf32, d.basis_split[kind],
                      static_cast<uint64_t>(2) * ntiles * 3 * tc::MV, 3 * tc::MV / 2))
      return fail(SMPLPP_ERR_CUDA, "CUDA", "cuTensorMapEncodeTiled failed for the blend basis");
  }
  SB_CUDA(cudaDeviceSynchronize());
  SB_CUDA(cudaFuncSetAttribute(blend_skin_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
  SB_CUDA(cudaFuncSetAttribute(blend_skin_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
  d.tc_ready = true;
  return SMPLPP_OK;
}

void tc_release_model(ModelDev & d)
{
  for(int kind = 0; kind < 2; kind++)
  {
    if(d.basis_split[kind]) cudaFree(d.basis_split[kind]);
    d.basis_split[kind] = nullptr;
  }
  d.tc_ready = false;
}

// coef (B,224) fp32 from K1; coef_split: tc_coef_split_bytes(B) of scratch; xforms must be readable up to
// align_up(B, 128) frames (the forward workspace guarantees it)
int launch_blend_skin_tc(const ModelDev & d, cudaStream_t st, int B, const float * coef, void * coef_split,
                         const float * xforms, const float * theta, float * out, bool tf32)
{
  if(!d.tc_ready) return fail(SMPLPP_ERR_INVALID, "SMPL", "tcgen05 blend variant is not available for this model");
  if((reinterpret_cast<uintptr_t>(out) & 7) || (reinterpret_cast<uintptr_t>(xforms) & 15))
    return fail(SMPLPP_ERR_INVALID, "SMPL", "tcgen05 blend variant needs 8-byte aligned vertices");
  const int Bpad = static_cast<int>(align_up(static_cast<size_t>(B), tc::NF));
  const long long n = static_cast<long long>(B) * tc::KP;
  if(tf32)
    split_coef_kernel<true><<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(coef, B, Bpad, coef_split);
  else
    split_coef_kernel<false><<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(coef, B, Bpad, coef_split);
  SB_LAUNCHED();
  alignas(64) CUtensorMap tmB;
  if(!encode_kmajor(&tmB, tf32, coef_split, static_cast<uint64_t>(2) * Bpad, tc::NF))
    return fail(SMPLPP_ERR_CUDA, "CUDA", "cuTensorMapEncodeTiled failed for the blend coefficients");
  tc::Params p;
  p.V = d.V;
  p.B = B;
  p.Bpad = Bpad;
  p.ntiles = d.tc_tiles;
  const int esize = tf32 ? 4 : 2;
  p.ke = tc::ROWB / esize;
  p.nkb = tc::KP / p.ke;
  p.ell_stride = d.Vpad;
  p.kmax = d.kmax;
  p.basis = d.basis;
  p.lbs_joint = d.lbs_joint;
  p.lbs_weight = d.lbs_weight;
  p.lbs_wsum = d.lbs_wsum;
  p.xforms = xforms;
  p.theta = theta;
  p.out = out;
  p.frames_fastest = g_tc_grid_order;
  dim3 grid(d.tc_tiles, Bpad / tc::NF);
  if(p.frames_fastest) grid = dim3(Bpad / tc::NF, d.tc_tiles);
  const CUtensorMap & tmA = *reinterpret_cast<const CUtensorMap *>(d.tmapA[tf32 ? 1 : 0]);
  if(tf32)
    blend_skin_tc_kernel<true><<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(tmA, tmB, p);
  else
    blend_skin_tc_kernel<false><<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(tmA, tmB, p);
  SB_LAUNCHED();
  return SMPLPP_OK;
}
} // namespace sb
