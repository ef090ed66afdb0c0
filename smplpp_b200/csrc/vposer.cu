// VPoser v2 decoder on sm_100a (reference src/VPoser.cpp):
//   MLP 32 -> 512 -> 512 -> 126 with LeakyReLU(0.01), Dropout = identity in eval   (VPoser.cpp:143-161)
//   6D -> rotation by Gram-Schmidt, columns b1,b2,b3                                (VPoser.cpp:129-141)
//   rotation -> axis-angle with the reference's branch rules                         (VPoser.cpp:25-120)
// and the 63x32 Jacobian d(axis-angle)/d(latent) that the reference obtains from autograd, computed here in
// forward mode (32 tangents): T2 = D2 W3 D1 W0, T3 = W5 T2, J = dAA/dy6 . T3.
#include "common.cuh"
#include "vposer.cuh"

using namespace sb;

// ------------------------------------------------------------------------------------------------------------
// forward-mode dual numbers (N tangents)
// ------------------------------------------------------------------------------------------------------------
template<int N>
struct Dual
{
  float v;
  float d[N];
};

template<int N>
__device__ __forceinline__ Dual<N> dconst(float c)
{
  Dual<N> r;
  r.v = c;
#pragma unroll
  for(int i = 0; i < N; i++) r.d[i] = 0.f;
  return r;
}
template<int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N> & a, const Dual<N> & b)
{
  Dual<N> r;
  r.v = a.v + b.v;
#pragma unroll
  for(int i = 0; i < N; i++) r.d[i] = a.d[i] + b.d[i];
  return r;
}
template<int N>
__device__ __forceinline__ Dual<N> operator-(const Dual<N> & a, const Dual<N> & b)
{
  Dual<N> r;
  r.v = a.v - b.v;
#pragma unroll
  for(int i = 0; i < N; i++) r.d[i] = a.d[i] - b.d[i];
  return r;
}
template<int N>
__device__ __forceinline__ Dual<N> operator*(const Dual<N> & a, const Dual<N> & b)
{
  Dual<N> r;
  r.v = a.v * b.v;
#pragma unroll
  for(int i = 0; i < N; i++) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
  return r;
}
template<int N>
__device__ __forceinline__ Dual<N> operator*(float s, const Dual<N> & a)
{
  Dual<N> r;
  r.v = s * a.v;
#pragma unroll
  for(int i = 0; i < N; i++) r.d[i] = s * a.d[i];
  return r;
}
template<int N>
__device__ __forceinline__ Dual<N> operator+(const Dual<N> & a, float s)
{
  Dual<N> r = a;
  r.v += s;
  return r;
}
template<int N>
__device__ __forceinline__ Dual<N> operator/(const Dual<N> & a, const Dual<N> & b)
{
  Dual<N> r;
  float inv = 1.f / b.v;
  r.v = a.v / b.v;
#pragma unroll
  for(int i = 0; i < N; i++) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
  return r;
}
template<int N>
__device__ __forceinline__ Dual<N> dsqrt(const Dual<N> & a)
{
  Dual<N> r;
  r.v = sqrtf(a.v);
  float g = 0.5f / r.v;
#pragma unroll
  for(int i = 0; i < N; i++) r.d[i] = g * a.d[i];
  return r;
}
template<int N>
__device__ __forceinline__ Dual<N> dacos(const Dual<N> & a)
{
  Dual<N> r;
  r.v = acosf(a.v);
  float g = -1.f / sqrtf(1.f - a.v * a.v);
#pragma unroll
  for(int i = 0; i < N; i++) r.d[i] = g * a.d[i];
  return r;
}
template<int N>
__device__ __forceinline__ Dual<N> dsin(const Dual<N> & a)
{
  Dual<N> r;
  r.v = sinf(a.v);
  float g = cosf(a.v);
#pragma unroll
  for(int i = 0; i < N; i++) r.d[i] = g * a.d[i];
  return r;
}

// torch::nn::functional::normalize (eps 1e-12) of a 3-vector
template<int N>
__device__ __forceinline__ void dnormalize(Dual<N> * x)
{
  Dual<N> n = dsqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  if(n.v > 1e-12f)
  {
#pragma unroll
    for(int k = 0; k < 3; k++) x[k] = x[k] / n;
  }
  else
  {
#pragma unroll
    for(int k = 0; k < 3; k++) x[k] = 1e12f * x[k];
  }
}

// convertRotMatToAxisAngle (VPoser.cpp:25-120) for one matrix R (row-major duals) -> aa[3]
template<int N>
__device__ void rotmat_to_axis_angle_dual(const Dual<N> * R, Dual<N> * aa)
{
  const float eps = 1.1920928955078125e-07f;       // FLT_EPSILON
  const float eps_sqrt = 3.4526698300124393e-04f;  // sqrt(eps)
  const float eps_sqrt2 = 1.8581361171917516e-02f; // eps^(1/4)
  const float half_one_minus_eps = 0.49999994039535522f;
  const float pi_minus = 3.1414926535897931f; // M_PI - 1e-4

  Dual<N> trace = R[0] + R[4] + R[8];
  Dual<N> theta = dacos(half_one_minus_eps * (trace + (-1.f)));
  Dual<N> w[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  if(1.f + trace.v < eps_sqrt2)
  {
    // near pi (:54-101)
    Dual<N> one_minus_tr = dconst<N>(1.f) - trace;
    Dual<N> three_minus_tr = dconst<N>(3.f) - trace;
    float sign[3] = {1.f, 1.f, 1.f};
    Dual<N> t[3];
#pragma unroll
    for(int k = 0; k < 3; k++)
    {
      Dual<N> s = (2.f * R[4 * k] + one_minus_tr) / three_minus_tr;
      t[k] = dsqrt(s + eps) * theta;
    }
    if(theta.v > pi_minus)
    {
      if(t[0].v > 0.f)
      {
        if(R[1].v + R[3].v < 0.f) sign[1] = -1.f;
        if(R[2].v + R[6].v < 0.f) sign[2] = -1.f;
      }
      else if(t[1].v > 0.f)
      {
        if(R[5].v + R[7].v < 0.f) sign[2] = -1.f;
      }
    }
    else
    {
#pragma unroll
      for(int k = 0; k < 3; k++)
        if(!(w[k].v >= 0.f)) sign[k] = -1.f;
    }
#pragma unroll
    for(int k = 0; k < 3; k++) aa[k] = sign[k] * t[k];
  }
  else if(fabsf(3.f - trace.v) < eps_sqrt)
  {
    // near zero (:105-111): 0.5 w (1 + th^2/6 + 7 th^4/360)
    Dual<N> t2 = theta * theta;
    Dual<N> series = ((1.f / 6.f) * t2 + 1.f) + (1.f / 360.f) * (7.f * (t2 * t2));
#pragma unroll
    for(int k = 0; k < 3; k++) aa[k] = (0.5f * w[k]) * series;
  }
  else
  {
    // generic (:112-116): w theta / (2 sin theta)
    Dual<N> f = theta / (2.f * dsin(theta));
#pragma unroll
    for(int k = 0; k < 3; k++) aa[k] = w[k] * f;
  }
}

// ContinousRotReprDecoder (VPoser.cpp:129-141) + axis-angle for one joint; y6 -> aa (3) and d aa / d y6 (3x6)
__device__ void decode_joint(const float * y6, float * aa_out, float * daa /* [3][6] or null */)
{
  Dual<6> y[6];
#pragma unroll
  for(int i = 0; i < 6; i++)
  {
    y[i] = dconst<6>(y6[i]);
    y[i].d[i] = 1.f;
  }
  // input.view(-1,3,2): a = (y0,y2,y4), b = (y1,y3,y5)
  Dual<6> b1[3] = {y[0], y[2], y[4]};
  Dual<6> c2[3] = {y[1], y[3], y[5]};
  dnormalize(b1);
  Dual<6> dot = b1[0] * c2[0] + b1[1] * c2[1] + b1[2] * c2[2];
  Dual<6> b2[3] = {c2[0] - dot * b1[0], c2[1] - dot * b1[1], c2[2] - dot * b1[2]};
  dnormalize(b2);
  Dual<6> b3[3] = {b1[1] * b2[2] - b1[2] * b2[1], b1[2] * b2[0] - b1[0] * b2[2], b1[0] * b2[1] - b1[1] * b2[0]};
  Dual<6> R[9] = {b1[0], b2[0], b3[0], b1[1], b2[1], b3[1], b1[2], b2[2], b3[2]};
  Dual<6> aa[3];
  rotmat_to_axis_angle_dual<6>(R, aa);
#pragma unroll
  for(int r = 0; r < 3; r++)
  {
    aa_out[r] = aa[r].v;
    if(daa)
    {
#pragma unroll
      for(int c = 0; c < 6; c++) daa[r * 6 + c] = aa[r].d[c];
    }
  }
}

__global__ void rotmat_to_axis_angle_kernel(long long n, const float * __restrict__ rot, float * __restrict__ out)
{
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if(i >= n) return;
  Dual<1> R[9], aa[3];
#pragma unroll
  for(int e = 0; e < 9; e++) R[e] = dconst<1>(rot[9 * i + e]);
  rotmat_to_axis_angle_dual<1>(R, aa);
#pragma unroll
  for(int k = 0; k < 3; k++) out[3 * i + k] = aa[k].v;
}

// ------------------------------------------------------------------------------------------------------------
// decoder kernel: persistent CTAs, FB = 2 frames per pass, 512 threads (thread i <-> hidden unit i)
// ------------------------------------------------------------------------------------------------------------
namespace vp
{
constexpr int H = SMPLPP_VPOSER_HIDDEN, L = SMPLPP_LATENT_DIM, NJ = SMPLPP_VPOSER_JOINTS, OUT = 6 * NJ;
constexpr int THREADS = 512;
constexpr int FB_JAC = 2; // frames per pass of the FFMA Jacobian variant (bounded by the T2 tile in shared memory)
constexpr int FB_FWD = 8; // forward-only variant: W3 (1 MB from L2) is re-read once per pass, so more frames per pass
                          // (ncu r01h: 1.8 ms per 16384 frames at 2 frames per pass, FMA pipe 10 % active)
template<int FB>
struct Smem
{
  float w0[H][L];          // decoder_net.0.weight (out, in) = [k][t] of the tangent recursion
  float t2[FB][H][L];      // T2 (jacobian only), later reused for T3
  float z[FB][L];
  float h1[FB][H], d1[FB][H], h2[FB][H], d2[FB][H];
  float y[FB][OUT + 2];
  float daa[FB][NJ][18];
};
// forward-only variant: the activations are stored [unit][frame], so that the FB frames of one input unit are one or two
// 16-byte broadcast loads per weight (frame-major rows cost one 4-byte load per FMA: the kernel ran at the issue rate of the
// shared-memory pipe); no LeakyReLU' copies (they go to the aux buffer only): 113 KB, two CTAs per SM at 64 registers.
// 0.71 -> 0.58 ms per 16384 frames; 16 frames per pass with W0 read through L1 measured the same (0.574 ms)
template<int FB>
struct SmemFwd
{
  float w0[H][L];
  float z[FB][L];
  alignas(16) float h1t[H][FB];
  alignas(16) float h2t[H][FB];
  float y[FB][OUT + 2];
  float daa[FB][NJ][18];
};
} // namespace vp

template<bool kJac, int FB>
__global__ void __launch_bounds__(vp::THREADS, kJac ? 1 : 2)
    vposer_decode_kernel(const float * __restrict__ w0, const float * __restrict__ b0, const float * __restrict__ w3t,
                         const float * __restrict__ b3, const float * __restrict__ w5t, const float * __restrict__ b5,
                         int B, const float * __restrict__ latent, long long latent_stride, float * __restrict__ aa_out,
                         long long aa_stride, float * __restrict__ jac_out, float * __restrict__ aux_out, int aux_ld)
{
  using namespace vp;
  using S = typename std::conditional<kJac, Smem<FB>, SmemFwd<FB>>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  S & s = *reinterpret_cast<S *>(smem_raw);
  const int tid = threadIdx.x;

  for(int i = tid; i < H * L / 4; i += THREADS)
    reinterpret_cast<float4 *>(&s.w0[0][0])[i] = __ldg(reinterpret_cast<const float4 *>(w0) + i);
  const float bias0 = b0[tid], bias3 = b3[tid];

  for(int f0 = blockIdx.x * FB; f0 < B; f0 += gridDim.x * FB)
  {
    __syncthreads();
    if(tid < FB * L)
    {
      int f = tid / L, t = tid % L;
      s.z[f][t] = (f0 + f < B) ? latent[static_cast<size_t>(f0 + f) * latent_stride + t] : 0.f;
    }
    __syncthreads();
    // layer 0
    {
      float acc[FB];
#pragma unroll
      for(int f = 0; f < FB; f++) acc[f] = bias0;
#pragma unroll
      for(int t = 0; t < L; t++)
      {
        float w = s.w0[tid][t];
#pragma unroll
        for(int f = 0; f < FB; f++) acc[f] = fmaf(w, s.z[f][t], acc[f]);
      }
#pragma unroll
      for(int f = 0; f < FB; f++)
      {
        bool pos = acc[f] > 0.f;
        if constexpr(kJac)
        {
          s.h1[f][tid] = pos ? acc[f] : 0.01f * acc[f];
          s.d1[f][tid] = pos ? 1.f : 0.01f;
        }
        else
          s.h1t[tid][f] = pos ? acc[f] : 0.01f * acc[f];
        if(aux_out && f0 + f < B) aux_out[static_cast<size_t>(f0 + f) * aux_ld + tid] = pos ? 1.f : 0.01f;
      }
    }
    __syncthreads();
    // layer 3 (w3t is (in, out): coalesced over tid)
    {
      float acc[FB];
#pragma unroll
      for(int f = 0; f < FB; f++) acc[f] = bias3;
#pragma unroll 8
      for(int k = 0; k < H; k++)
      {
        float w = __ldg(w3t + static_cast<size_t>(k) * H + tid);
        if constexpr(kJac)
        {
#pragma unroll
          for(int f = 0; f < FB; f++) acc[f] = fmaf(w, s.h1[f][k], acc[f]);
        }
        else
        {
          static_assert(kJac || FB % 4 == 0, "forward variant: frames in groups of four");
#pragma unroll
          for(int q = 0; q < FB / 4; q++)
          {
            const float4 h = reinterpret_cast<const float4 *>(s.h1t[k])[q];
            acc[4 * q] = fmaf(w, h.x, acc[4 * q]), acc[4 * q + 1] = fmaf(w, h.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(w, h.z, acc[4 * q + 2]), acc[4 * q + 3] = fmaf(w, h.w, acc[4 * q + 3]);
          }
        }
      }
#pragma unroll
      for(int f = 0; f < FB; f++)
      {
        bool pos = acc[f] > 0.f;
        if constexpr(kJac)
        {
          s.h2[f][tid] = pos ? acc[f] : 0.01f * acc[f];
          s.d2[f][tid] = pos ? 1.f : 0.01f;
        }
        else
          s.h2t[tid][f] = pos ? acc[f] : 0.01f * acc[f];
        if(aux_out && f0 + f < B) aux_out[static_cast<size_t>(f0 + f) * aux_ld + H + tid] = pos ? 1.f : 0.01f;
      }
    }
    __syncthreads();
    // layer 5: threads (frame group, o): every thread covers FB * 128 / THREADS frames so that a weight is loaded once
    {
      constexpr int FPT = FB * 128 > THREADS ? FB * 128 / THREADS : 1; // frames per thread
      const int o = tid & 127, fg = tid >> 7;
      if(o < OUT && fg * FPT < FB)
      {
        float acc[FPT];
#pragma unroll
        for(int i = 0; i < FPT; i++) acc[i] = b5[o];
#pragma unroll 8
        for(int k = 0; k < H; k++)
        {
          const float w = __ldg(w5t + static_cast<size_t>(k) * 128 + o);
          if constexpr(kJac)
          {
#pragma unroll
            for(int i = 0; i < FPT; i++) acc[i] = fmaf(w, s.h2[fg * FPT + i][k], acc[i]);
          }
          else
          {
#pragma unroll
            for(int i = 0; i < FPT; i++) acc[i] = fmaf(w, s.h2t[k][fg * FPT + i], acc[i]);
          }
        }
#pragma unroll
        for(int i = 0; i < FPT; i++) s.y[fg * FPT + i][o] = acc[i];
      }
    }
    __syncthreads();
    // 6D -> R -> axis-angle (+ d aa / d y6)
    if(tid < FB * NJ)
    {
      int f = tid / NJ, j = tid % NJ;
      if(f0 + f < B)
      {
        float aa[3];
        decode_joint(&s.y[f][6 * j], aa, (kJac || aux_out) ? s.daa[f][j] : nullptr);
        float * dst = aa_out + static_cast<size_t>(f0 + f) * aa_stride + 3 * j;
        dst[0] = aa[0], dst[1] = aa[1], dst[2] = aa[2];
        if(aux_out)
        {
          float * da = aux_out + static_cast<size_t>(f0 + f) * aux_ld + 2 * H + 18 * j;
#pragma unroll
          for(int e = 0; e < 18; e++) da[e] = s.daa[f][j][e];
        }
      }
    }
    if(kJac)
    {
      Smem<FB> & sj = *reinterpret_cast<Smem<FB> *>(smem_raw);
      // T2[f][i][t] = d2[f][i] * sum_k W3[i][k] d1[f][k] W0[k][t]   (thread i = tid)
      float acc[FB][L];
#pragma unroll
      for(int f = 0; f < FB; f++)
#pragma unroll
        for(int t = 0; t < L; t++) acc[f][t] = 0.f;
#pragma unroll 2
      for(int k = 0; k < H; k++)
      {
        float w = __ldg(w3t + static_cast<size_t>(k) * H + tid);
        float wf[FB];
#pragma unroll
        for(int f = 0; f < FB; f++) wf[f] = w * sj.d1[f][k];
        const float4 * w0k = reinterpret_cast<const float4 *>(sj.w0[k]);
#pragma unroll
        for(int q = 0; q < L / 4; q++)
        {
          float4 v = w0k[q];
#pragma unroll
          for(int f = 0; f < FB; f++)
          {
            acc[f][4 * q + 0] = fmaf(wf[f], v.x, acc[f][4 * q + 0]);
            acc[f][4 * q + 1] = fmaf(wf[f], v.y, acc[f][4 * q + 1]);
            acc[f][4 * q + 2] = fmaf(wf[f], v.z, acc[f][4 * q + 2]);
            acc[f][4 * q + 3] = fmaf(wf[f], v.w, acc[f][4 * q + 3]);
          }
        }
      }
#pragma unroll
      for(int f = 0; f < FB; f++)
      {
        float d = sj.d2[f][tid];
#pragma unroll
        for(int q = 0; q < L / 4; q++)
          reinterpret_cast<float4 *>(sj.t2[f][tid])[q] =
              make_float4(d * acc[f][4 * q], d * acc[f][4 * q + 1], d * acc[f][4 * q + 2], d * acc[f][4 * q + 3]);
      }
      __syncthreads();
      // T3[f][o][t] = sum_i W5[o][i] T2[f][i][t]: thread -> (o = tid & 127, t-quarter = tid >> 7), both frames
      const int o = tid & 127, tq = tid >> 7;
      float a3[FB][8];
#pragma unroll
      for(int f = 0; f < FB; f++)
#pragma unroll
        for(int t = 0; t < 8; t++) a3[f][t] = 0.f;
      if(o < OUT)
      {
#pragma unroll 4
        for(int i = 0; i < H; i++)
        {
          float w = __ldg(w5t + static_cast<size_t>(i) * 128 + o);
#pragma unroll
          for(int f = 0; f < FB; f++)
          {
            const float4 * tp = reinterpret_cast<const float4 *>(&sj.t2[f][i][8 * tq]);
            float4 u0 = tp[0], u1 = tp[1];
            a3[f][0] = fmaf(w, u0.x, a3[f][0]);
            a3[f][1] = fmaf(w, u0.y, a3[f][1]);
            a3[f][2] = fmaf(w, u0.z, a3[f][2]);
            a3[f][3] = fmaf(w, u0.w, a3[f][3]);
            a3[f][4] = fmaf(w, u1.x, a3[f][4]);
            a3[f][5] = fmaf(w, u1.y, a3[f][5]);
            a3[f][6] = fmaf(w, u1.z, a3[f][6]);
            a3[f][7] = fmaf(w, u1.w, a3[f][7]);
          }
        }
      }
      __syncthreads(); // everyone is done reading T2
      if(o < OUT)
      {
#pragma unroll
        for(int f = 0; f < FB; f++)
#pragma unroll
          for(int t = 0; t < 8; t++) sj.t2[f][o][8 * tq + t] = a3[f][t]; // T3 stored over T2 rows 0..125
      }
      __syncthreads();
      // J[f][3j+r][t] = sum_c daa[f][j][r][c] T3[f][6j+c][t]
      for(int idx = tid; idx < FB * 63 * L; idx += THREADS)
      {
        int f = idx / (63 * L), rem = idx % (63 * L), row = rem / L, t = rem % L;
        if(f0 + f >= B) continue;
        int j = row / 3, r = row % 3;
        float acc2 = 0.f;
#pragma unroll
        for(int c = 0; c < 6; c++) acc2 = fmaf(sj.daa[f][j][r * 6 + c], sj.t2[f][6 * j + c][t], acc2);
        jac_out[(static_cast<size_t>(f0 + f) * 63 + row) * L + t] = acc2;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------------
namespace sb
{
std::atomic<int> g_vposer_jac_variant{0};

int launch_vposer_decode(const smplpp_vposer * vposer, cudaStream_t st, int B, const float * latent,
                         long long latent_stride, float * aa, long long aa_stride, float * jac, float * aux_ws)
{
  // per launch: the attribute is per device, and one process may drive several
  SB_CUDA(cudaFuncSetAttribute(vposer_decode_kernel<true, vp::FB_JAC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(sizeof(vp::Smem<vp::FB_JAC>))));
  SB_CUDA(cudaFuncSetAttribute(vposer_decode_kernel<false, vp::FB_FWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               static_cast<int>(sizeof(vp::SmemFwd<vp::FB_FWD>))));
  int sms = 148;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const bool jac_ffma = jac && !(vposer->tc_ready && g_vposer_jac_variant == 0);
  const int fb = jac_ffma ? vp::FB_JAC : vp::FB_FWD;
  int passes = (B + fb - 1) / fb;
  if(jac && vposer->tc_ready && g_vposer_jac_variant == 0)
  {
    // forward pass (FFMA, 0.7 MFLOP per frame) leaves LeakyReLU'(h1), LeakyReLU'(h2) and d aa / d y6 of every frame in a
    // scratch buffer; the 21 MFLOP per frame of the Jacobian chain run on the tensor cores (vposer_tc.cu)
    const int aux_ld = static_cast<int>(vposer_tc_aux_floats());
    float * aux = aux_ws;
    if(!aux)
    {
      const size_t need = static_cast<size_t>(B) * aux_ld;
      if(vposer->tc_aux_floats < need)
      {
        SB_CUDA(cudaStreamSynchronize(st)); // earlier launches may still read the old buffer
        if(vposer->tc_aux) cudaFree(vposer->tc_aux);
        vposer->tc_aux = nullptr;
        vposer->tc_aux_floats = 0;
        SB_CUDA(cudaMalloc(reinterpret_cast<void **>(&vposer->tc_aux), need * sizeof(float)));
        vposer->tc_aux_floats = need;
      }
      aux = vposer->tc_aux;
    }
    int grid = passes < 2 * sms ? passes : 2 * sms;
    vposer_decode_kernel<false, vp::FB_FWD><<<grid, vp::THREADS, sizeof(vp::SmemFwd<vp::FB_FWD>), st>>>(
        vposer->w0, vposer->b0, vposer->w3t, vposer->b3, vposer->w5t, vposer->b5, B, latent, latent_stride, aa,
        aa_stride, nullptr, aux, aux_ld);
    SB_LAUNCHED();
    return launch_vposer_jac_tc(*vposer, st, B, aux, jac);
  }
  if(jac)
  {
    int grid = passes < sms ? passes : sms;
    vposer_decode_kernel<true, vp::FB_JAC><<<grid, vp::THREADS, sizeof(vp::Smem<vp::FB_JAC>), st>>>(
        vposer->w0, vposer->b0, vposer->w3t, vposer->b3, vposer->w5t, vposer->b5, B, latent, latent_stride, aa,
        aa_stride, jac, nullptr, 0);
  }
  else
  {
    int grid = passes < 2 * sms ? passes : 2 * sms;
    vposer_decode_kernel<false, vp::FB_FWD><<<grid, vp::THREADS, sizeof(vp::SmemFwd<vp::FB_FWD>), st>>>(
        vposer->w0, vposer->b0, vposer->w3t, vposer->b3, vposer->w5t, vposer->b5, B, latent, latent_stride, aa,
        aa_stride, nullptr, nullptr, 0);
  }
  SB_LAUNCHED();
  return SMPLPP_OK;
}
} // namespace sb

extern "C" int smplpp_vposer_create(const smplpp_vposer_desc * desc, smplpp_vposer_t ** out)
{
  using namespace vp;
  if(!desc || !out || !desc->w0 || !desc->b0 || !desc->w3 || !desc->b3 || !desc->w5 || !desc->b5)
    return fail(SMPLPP_ERR_INVALID, "VPoser", "invalid dimension of decoder_net parameters!");
  if(smplpp_device_count() < 1) return fail(SMPLPP_ERR_CUDA, "CUDA", "no CUDA device (there is no CPU fallback)");
  auto v = new smplpp_vposer();
  std::vector<float> w3t(static_cast<size_t>(H) * H), w5t(static_cast<size_t>(H) * 128, 0.f);
  for(int o = 0; o < H; o++)
    for(int k = 0; k < H; k++) w3t[static_cast<size_t>(k) * H + o] = desc->w3[static_cast<size_t>(o) * H + k];
  for(int o = 0; o < OUT; o++)
    for(int k = 0; k < H; k++) w5t[static_cast<size_t>(k) * 128 + o] = desc->w5[static_cast<size_t>(o) * H + k];
  auto up = [&](float ** dst, const float * src, size_t n) -> int {
    SB_CUDA(cudaMalloc(reinterpret_cast<void **>(dst), n * sizeof(float)));
    SB_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return SMPLPP_OK;
  };
  int rc = up(&v->w0, desc->w0, static_cast<size_t>(H) * L);
  if(rc == SMPLPP_OK) rc = up(&v->b0, desc->b0, H);
  if(rc == SMPLPP_OK) rc = up(&v->w3t, w3t.data(), w3t.size());
  if(rc == SMPLPP_OK) rc = up(&v->b3, desc->b3, H);
  if(rc == SMPLPP_OK) rc = up(&v->w5t, w5t.data(), w5t.size());
  if(rc == SMPLPP_OK) rc = up(&v->b5, desc->b5, OUT);
  if(rc == SMPLPP_OK) rc = vposer_tc_prepare(*v, desc->w0, desc->w3, desc->w5);
  if(rc != SMPLPP_OK)
  {
    smplpp_vposer_destroy(v);
    return rc;
  }
  *out = v;
  return SMPLPP_OK;
}

extern "C" void smplpp_vposer_destroy(smplpp_vposer_t * v)
{
  if(!v) return;
  cudaFree(v->w0);
  cudaFree(v->b0);
  cudaFree(v->w3t);
  cudaFree(v->b3);
  cudaFree(v->w5t);
  cudaFree(v->b5);
  vposer_tc_release(*v);
  if(v->tc_aux) cudaFree(v->tc_aux);
  delete v;
}

extern "C" int smplpp_vposer_decode(const smplpp_vposer_t * vposer, void * stream, int64_t batch, const float * latent,
                                    float * axis_angle, float * jacobian)
{
  if(!vposer || batch < 1 || !latent || !axis_angle)
    return fail(SMPLPP_ERR_INVALID, "VPoser", "invalid latent tensor!");
  return launch_vposer_decode(vposer, as_stream(stream), static_cast<int>(batch), latent, SMPLPP_LATENT_DIM, axis_angle,
                              63, jacobian);
}

extern "C" int smplpp_rotmat_to_axis_angle(void * stream, int64_t n, const float * rotmat, float * axis_angle)
{
  if(n < 1 || !rotmat || !axis_angle) return fail(SMPLPP_ERR_INVALID, "VPoser", "invalid rotation tensor!");
  rotmat_to_axis_angle_kernel<<<static_cast<unsigned>((n + 127) / 128), 128, 0, as_stream(stream)>>>(n, rotmat,
                                                                                                      axis_angle);
  SB_LAUNCHED();
  return SMPLPP_OK;
}
