// C2': per-frame normal equations on the fp64 tensor cores + blocked Cholesky (node/node.cpp:884-968 of the reference:
// A = J'J in double, damping, VPoser prior, LLT solve, update; and the partial elimination of the shared-beta stage).
//
// One CTA (4 warps) per frame.  The augmented Gram matrix [J | e]' [J | e] = [A b; b' |e|^2] (N = D + 1 <= 8 NB) is cut
// into 8x8 tiles of the lower triangle; a warp owns every fourth tile (column-major order) and keeps their accumulators
// in REGISTERS as mma.m8n8k4.f64 C fragments for the whole kernel: four live rows of J are one k-step, so a chunk of four
// rows costs every lane NB loads + NB conversions and the warp NT/4 DMMAs (the scalar kernel: 8 loads, 8 conversions and
// 16 DFMAs per thread and row, 3.0 G warp instructions per 16384 frames, 63 % issue-bound).  The rows arrive through an
// eight-stage cp.async ring.  Cholesky is right-looking over 8-wide panels: the owners of column p store their tiles to
// shared memory, the owner of the diagonal tile factors it with warp shuffles, one thread per row solves the panel against
// it, and every warp applies the rank-8 update to its own register tiles with two DMMAs per tile.  Because the right-hand
// side is row D of the matrix, the forward substitution falls out of the elimination (L[D][k] = (L^-1 b)[k]) and only the
// back substitution remains; with fewer pivots than unknowns (shared-beta stage) the trailing tiles are left holding the
// Schur complement S | r exactly as ik_solve_kernel leaves them.
#include <utility>

#include "ik_solve.cuh"

namespace sb
{
std::atomic<int> g_solve_variant{0};
}

using namespace sb;

namespace
{
constexpr int NW = 4;
constexpr int THREADS = NW * 32;
constexpr int NSTAGE = 8; // 7 chunks of 1.4 KB in flight per CTA: the J rows come from HBM, two chunks in flight left the loop latency-bound

constexpr int col_start(int NB, int j) // index of tile (j, j) in the column-major list of lower-triangular tiles
{
  return j * NB - j * (j - 1) / 2;
}
constexpr int tile_col(int NB, int t)
{
  int j = 0;
  while(col_start(NB, j + 1) <= t) j++;
  return j;
}
constexpr int tile_row(int NB, int t)
{
  const int j = tile_col(NB, t);
  return j + (t - col_start(NB, j));
}

template<int NB>
struct Cfg
{
  static constexpr int N8 = 8 * NB;
  static constexpr int NT = NB * (NB + 1) / 2;
  static constexpr int TPW = (NT + NW - 1) / NW;
  // floats per staged row: 8 or 24 mod 32, so that the fragment loads (lane (g, t) reads row t, column 8 blk + g) hit 32
  // different banks
  static constexpr int LDS = (NB % 2 == 0) ? N8 + 8 : N8;
  static constexpr int off_lt = 0;                               // NT tiles of 64 doubles (swizzled rows)
  static constexpr int off_p = off_lt + NT * 512;                // 2 x NB panel tiles (columns >= npl zeroed)
  static constexpr int off_invd = off_p + 2 * NB * 512;          // 1 / L[i][i]
  static constexpr int off_x = off_invd + N8 * 8;                // solution
  static constexpr int off_misc = off_x + N8 * 8;                // |e|^2 (double), ok flag (int)
  static constexpr int off_stage = off_misc + 16;                // NSTAGE x 4 rows x LDS floats
  static constexpr int off_e = off_stage + NSTAGE * 4 * LDS * 4; // 4 n floats
};

// element (r, c) of an 8x8 tile: rows 2, 3, 6, 7 swap their column halves, so that the fragment loads of a warp
// (lane (g, t) reads [g][t] and [g][t + 4], 8 bytes each) are conflict-free
__device__ __forceinline__ int swz(int r, int c)
{
  return r * 8 + (c ^ ((r & 2) << 1));
}

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

__device__ __forceinline__ double dneg(double a) // sign flip on the integer pipe
{
  return __hiloint2double(__double2hiint(a) ^ 0x80000000, __double2loint(a));
}

// The kernel body exists once per warp of the CTA (static register indices): one out-of-line barrier, so that the four
// warps arrive at the SAME instruction (compute-sanitizer's synccheck reports warps that meet at a named barrier from
// different code addresses as divergent).
__device__ __noinline__ void cta_sync()
{
  asm volatile("bar.sync 1, %0;\n" ::"n"(THREADS) : "memory");
}

__device__ __forceinline__ void cp_async16(void * smem, const void * gmem)
{
  const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit()
{
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template<int N>
__device__ __forceinline__ void cp_async_wait()
{
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

template<typename F, int... K>
__device__ __forceinline__ void for_each_k(F && f, std::integer_sequence<int, K...>)
{
  (f(std::integral_constant<int, K>{}), ...);
}

// Cholesky of the leading npl pivots of one 8x8 tile, executed by a whole warp (lane & 7 = row, the four lane groups
// compute the same thing); rows below the last pivot keep the Schur complement.  invd receives 1 / L[r][r].
__device__ __forceinline__ void diag_factor(double * tile, int npl, double * invd, int * ok, int lane)
{
  const int r = lane & 7;
  double a[8];
#pragma unroll
  for(int c = 0; c < 8; c++) a[c] = tile[swz(r, c)];
  double myinv = 0.0;
  bool bad = false;
#pragma unroll
  for(int k = 0; k < 8; k++)
  {
    if(k < npl)
    {
      const double dk = __shfl_sync(0xffffffffu, a[k], k);
      if(!(dk > 0.0)) bad = true;
      const double inv = rsqrt(dk);
      const double l = a[k] * inv;
      a[k] = l;
      if(r == k) myinv = inv;
#pragma unroll
      for(int c = k + 1; c < 8; c++)
      {
        const double lc = __shfl_sync(0xffffffffu, l, c);
        if(r >= c) a[c] = fma(-l, lc, a[c]);
      }
    }
  }
  __syncwarp(); // lanes 8..31 have read the rows that lanes 0..7 overwrite
  if(lane < 8)
  {
#pragma unroll
    for(int c = 0; c < 8; c++)
      if(c <= r) tile[swz(r, c)] = a[c];
    invd[r] = myinv;
  }
  if(bad && lane == 0) *ok = 0;
  __syncwarp();
}

// one row of the panel below the diagonal tile: x L_pp' = t for the npl pivot columns, the remaining columns keep
// t - x L_pp[c][:npl]' (Schur complement); result to the tile itself and, with the non-pivot columns zeroed, to the panel
template<bool FULL>
__device__ __forceinline__ void trsm_row(double * row_l, double * row_p, int rr, const double * lpp, const double * invd, int npl)
{
  const int s = (rr & 2) << 1;
  double tv[8], xv[8];
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    const double2 v2 = *reinterpret_cast<const double2 *>(row_l + ((2 * q) ^ s));
    tv[2 * q] = v2.x, tv[2 * q + 1] = v2.y;
  }
#pragma unroll
  for(int c = 0; c < 8; c++)
  {
    double acc = tv[c];
#pragma unroll
    for(int k = 0; k < c; k++)
      if(FULL || k < npl) acc = fma(-xv[k], lpp[swz(c, k)], acc);
    xv[c] = (FULL || c < npl) ? acc * invd[c] : acc;
  }
#pragma unroll
  for(int q = 0; q < 4; q++)
  {
    *reinterpret_cast<double2 *>(row_l + ((2 * q) ^ s)) = make_double2(xv[2 * q], xv[2 * q + 1]);
    double2 pv = make_double2(xv[2 * q], xv[2 * q + 1]);
    if(!FULL)
    {
      if(2 * q >= npl) pv.x = 0.0;
      if(2 * q + 1 >= npl) pv.y = 0.0;
    }
    *reinterpret_cast<double2 *>(row_p + ((2 * q) ^ s)) = pv;
  }
}

template<int NB, int W>
__device__ __forceinline__ void solve_body(const IkSolveParams & p, unsigned char * smem)
{
  using C = Cfg<NB>;
  constexpr int N8 = C::N8, NT = C::NT, TPW = C::TPW, LDS = C::LDS;
  double * Lt = reinterpret_cast<double *>(smem + C::off_lt);
  double * Pn = reinterpret_cast<double *>(smem + C::off_p);
  double * invd = reinterpret_cast<double *>(smem + C::off_invd);
  double * x = reinterpret_cast<double *>(smem + C::off_x);
  double * s_esq = reinterpret_cast<double *>(smem + C::off_misc);
  int * s_ok = reinterpret_cast<int *>(smem + C::off_misc + 8);
  float * stage = reinterpret_cast<float *>(smem + C::off_stage);
  float * s_e = reinterpret_cast<float *>(smem + C::off_e);
  const auto ks = std::make_integer_sequence<int, TPW>{};

  const int tid = threadIdx.x, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int f = blockIdx.x;
  const int D = p.D, rpt = p.rows_per_task, nlive = p.n * rpt, nchunks = (nlive + 3) >> 2;
  const int ld = p.ld, ld4 = ld >> 2;
  const int gD = D & 7; // the right-hand side is row D = 8 (NB - 1) + gD
  const int npiv = p.schur ? D - p.beta_cols : D;
  const int npanels = (npiv + 7) >> 3;
  const float * J = p.J + static_cast<size_t>(f) * 4 * p.n * ld;
  const float * e = p.e + static_cast<size_t>(f) * 4 * p.n;
  const int nvalid = p.frame_info[2 * f], bad = p.frame_info[2 * f + 1];
  const bool too_few = p.skip_if_too_few && nvalid < p.n / 2; // node.cpp:785

  for(int i = tid; i < NSTAGE * 4 * LDS; i += THREADS) stage[i] = 0.f;
  for(int i = tid; i < 4 * p.n; i += THREADS) s_e[i] = e[i];
  if(tid == 0) *s_ok = 1;
  cta_sync();

  // ---- [J | e]' [J | e]: chunks of four LIVE rows (a task contributes rpt = 3 or 4 of its 4 row slots) ----
  const int q_ld = tid / ld4, c4_ld = tid - q_ld * ld4; // this thread's 16 bytes of a chunk
  const bool loader = tid < 4 * ld4;
  const bool fixer = loader && c4_ld == min(D >> 2, ld4 - 1); // also writes e into column D and clears (D, ld)
  auto live_row = [&](int l) { // global row of live row l
    const int task = l / rpt;
    return 4 * task + (l - task * rpt);
  };
  auto issue = [&](int c, int buf) {
    if(c < nchunks && loader)
    {
      const int l = 4 * c + q_ld;
      float * dst = stage + (buf * 4 + q_ld) * LDS + 4 * c4_ld;
      if(l < nlive)
        cp_async16(dst, J + static_cast<size_t>(live_row(l)) * ld + 4 * c4_ld);
      else
        *reinterpret_cast<float4 *>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    cp_async_commit();
  };
  double acc[TPW][2];
#pragma unroll
  for(int k = 0; k < TPW; k++) acc[k][0] = acc[k][1] = 0.0;
#pragma unroll
  for(int c = 0; c < NSTAGE - 1; c++) issue(c, c);
  int buf = 0;
  for(int c = 0; c < nchunks; c++)
  {
    cp_async_wait<NSTAGE - 2>();
    if(fixer)
    {
      const int l = 4 * c + q_ld;
      float * row = stage + (buf * 4 + q_ld) * LDS;
      row[D] = l < nlive ? s_e[live_row(l)] : 0.f;
      for(int cc = D + 1; cc < ld; cc++) row[cc] = 0.f;
    }
    cta_sync();
    issue(c + NSTAGE - 1, buf == 0 ? NSTAGE - 1 : buf - 1);
    const float * st = stage + (buf * 4 + t) * LDS + g;
    double v[NB];
#pragma unroll
    for(int b = 0; b < NB; b++) v[b] = static_cast<double>(st[8 * b]);
    for_each_k(
        [&](auto kc) {
          constexpr int K = decltype(kc)::value, T = NW * K + W;
          if constexpr(T < NT)
          {
            constexpr int I = tile_row(NB, T), Jc = tile_col(NB, T);
            dmma(acc[K], v[I], v[Jc]);
          }
        },
        ks);
    buf = buf == NSTAGE - 1 ? 0 : buf + 1;
  }
  // ---- |e|^2 = element (D, D) ----
  for_each_k(
      [&](auto kc) {
        constexpr int K = decltype(kc)::value, T = NW * K + W;
        if constexpr(T == NT - 1)
          if(g == gD)
          {
            if(2 * t == gD) *s_esq = acc[K][0];
            if(2 * t + 1 == gD) *s_esq = acc[K][1];
          }
      },
      ks);
  cta_sync();
  const double esq = *s_esq;
  // ---- damping (node.cpp:887-893) and the VPoser prior (:895-904) ----
  auto prior_w = [&](int i) -> double { return i < 6 ? 0.0 : (i >= p.theta_dim - 6 ? p.hand_reg : p.latent_reg); };
  for_each_k(
      [&](auto kc) {
        constexpr int K = decltype(kc)::value, T = NW * K + W;
        if constexpr(T < NT)
        {
          constexpr int I = tile_row(NB, T), Jc = tile_col(NB, T);
          if constexpr(I == Jc)
          {
#pragma unroll
            for(int h = 0; h < 2; h++)
              if(2 * t + h == g)
              {
                const int i = 8 * I + g;
                if(i < D)
                {
                  double reg = i < p.theta_dim ? p.reg_theta : (i < p.theta_dim + p.phi_cols ? p.reg_phi : p.reg_beta);
                  double add = reg + esq;
                  if(p.schur && i >= npiv) add = 0.0; // the beta block is damped once, globally, in the apply step
                  if(p.vposer && i < p.theta_dim) add += prior_w(i);
                  acc[K][h] += add;
                }
              }
          }
          if constexpr(I == NB - 1)
          {
            if(p.vposer && g == gD)
            {
#pragma unroll
              for(int h = 0; h < 2; h++)
              {
                const int i = 8 * Jc + 2 * t + h;
                if(i < p.theta_dim)
                  acc[K][h] += prior_w(i) * static_cast<double>(p.theta_state[static_cast<size_t>(f) * p.theta_dim + i]);
              }
            }
          }
        }
      },
      ks);
  // ---- optional outputs in the reference layout ----
  if(p.a_out || p.b_out)
  {
    auto ref_col = [&](int c) {
      if(c < p.theta_dim + p.phi_cols) return c;
      return p.theta_dim + 2 * p.n + (c - p.theta_dim - p.phi_cols);
    };
    double * A = p.a_out ? p.a_out + static_cast<size_t>(f) * p.dim_ref * p.dim_ref : nullptr;
    double * bo = p.b_out ? p.b_out + static_cast<size_t>(f) * p.dim_ref : nullptr;
    if(A)
      for(int i = tid; i < p.dim_ref * p.dim_ref; i += THREADS) A[i] = 0.0;
    if(bo)
      for(int i = tid; i < p.dim_ref; i += THREADS) bo[i] = 0.0;
    cta_sync();
    for_each_k(
        [&](auto kc) {
          constexpr int K = decltype(kc)::value, T = NW * K + W;
          if constexpr(T < NT)
          {
            constexpr int I = tile_row(NB, T), Jc = tile_col(NB, T);
#pragma unroll
            for(int h = 0; h < 2; h++)
            {
              const int r = 8 * I + g, c = 8 * Jc + 2 * t + h;
              if(A && r < D && c <= r)
              {
                A[ref_col(r) * p.dim_ref + ref_col(c)] = acc[K][h];
                A[ref_col(c) * p.dim_ref + ref_col(r)] = acc[K][h];
              }
              if(bo && r == D && c < D) bo[ref_col(c)] = acc[K][h];
            }
          }
        },
        ks);
    if(A && !p.phi_cols)
      for(int i = tid; i < 2 * p.n; i += THREADS)
        A[(p.theta_dim + i) * p.dim_ref + p.theta_dim + i] = static_cast<double>(p.reg_phi) + esq;
  }

  // ---- right-looking Cholesky of the leading npiv pivots over 8-wide panels ----
  int pbuf = 0;
  for(int pp = 0; pp < npanels; pp++)
  {
    const int npl = min(8, npiv - 8 * pp);
    const int tpp = pp * NB - pp * (pp - 1) / 2; // tile (pp, pp)
    for_each_k(
        [&](auto kc) {
          constexpr int K = decltype(kc)::value, T = NW * K + W;
          if constexpr(T < NT)
            if(constexpr int Jc = tile_col(NB, T); Jc == pp)
              *reinterpret_cast<double2 *>(Lt + T * 64 + swz(g, 2 * t)) = make_double2(acc[K][0], acc[K][1]);
        },
        ks);
    if((tpp & (NW - 1)) == W)
    {
      __syncwarp();
      diag_factor(Lt + tpp * 64, npl, invd + 8 * pp, s_ok, lane);
    }
    cta_sync();
    {
      const int R = 8 * (pp + 1) + tid;
      if(R < N8)
      {
        const int i = R >> 3, rr = R & 7;
        double * row_l = Lt + (tpp + i - pp) * 64 + rr * 8;
        double * row_p = Pn + (pbuf * NB + i) * 64 + rr * 8;
        if(npl == 8)
          trsm_row<true>(row_l, row_p, rr, Lt + tpp * 64, invd + 8 * pp, 8);
        else
          trsm_row<false>(row_l, row_p, rr, Lt + tpp * 64, invd + 8 * pp, npl);
      }
    }
    cta_sync();
    {
      const double * pb = Pn + pbuf * NB * 64;
      for_each_k(
          [&](auto kc) {
            constexpr int K = decltype(kc)::value, T = NW * K + W;
            if constexpr(T < NT)
            {
              constexpr int I = tile_row(NB, T), Jc = tile_col(NB, T);
              if(Jc > pp)
              {
                const double a0 = pb[I * 64 + swz(g, t)], a1 = pb[I * 64 + swz(g, t + 4)];
                const double b0 = pb[Jc * 64 + swz(g, t)], b1 = pb[Jc * 64 + swz(g, t + 4)];
                dmma(acc[K], dneg(a0), b0);
                dmma(acc[K], dneg(a1), b1);
              }
            }
          },
          ks);
    }
    pbuf ^= 1;
  }
  // tiles right of the last panel (shared-beta stage): the Schur complement
  for_each_k(
      [&](auto kc) {
        constexpr int K = decltype(kc)::value, T = NW * K + W;
        if constexpr(T < NT)
          if(tile_col(NB, T) >= npanels)
            *reinterpret_cast<double2 *>(Lt + T * 64 + swz(g, 2 * t)) = make_double2(acc[K][0], acc[K][1]);
      },
      ks);
  cta_sync();
  auto lget = [&](int r, int c) -> double { // r >= c
    const int i = r >> 3, j = c >> 3;
    return Lt[(j * NB - j * (j - 1) / 2 + i - j) * 64 + swz(r & 7, c & 7)];
  };
  const int ok = *s_ok;

  if(p.schur)
  {
    // ---- shared-beta stage: S | r | |e|^2 and the factor rows for the apply step (layout of ik_solve_kernel) ----
    double * out = p.schur_out + static_cast<size_t>(f) * 111;
    const bool good = ok && !bad && !too_few;
    for(int i = tid; i < 111; i += THREADS)
    {
      double v = 0.0;
      if(good)
      {
        if(i < 100)
        {
          const int r = i / 10, c = i % 10;
          v = r >= c ? lget(npiv + r, npiv + c) : lget(npiv + c, npiv + r);
        }
        else if(i < 110)
          v = lget(D, npiv + (i - 100));
        else
          v = esq;
      }
      out[i] = v;
    }
    const int nff = npiv * (npiv + 1) / 2;
    const int Pf = nff + (p.beta_cols + 1) * npiv;
    double * fw = p.factor_ws + static_cast<size_t>(f) * Pf;
    for(int r = tid >> 5; r < npiv; r += NW)
      for(int c = lane; c <= r; c += 32) fw[r * (r + 1) / 2 + c] = lget(r, c);
    for(int i = tid; i < (p.beta_cols + 1) * npiv; i += THREADS)
    {
      const int r = i / npiv, c = i - r * npiv;
      fw[nff + i] = lget(npiv + r, c);
    }
    if(tid == 0) p.status[f] = too_few ? 1 : ((bad || !ok) ? 2 : 0);
    return;
  }

  // ---- back substitution L' x = -y, y = row D of the factor (warp 0; eight unknowns per step, every lane solves the
  //      diagonal tile redundantly) ----
  if constexpr(W == 0)
  {
    for(int k = lane; k < 8 * npanels; k += 32) x[k] = k < D ? -lget(D, k) : 0.0;
    __syncwarp();
    for(int pb = npanels - 1; pb >= 0; pb--)
    {
      const int nloc = min(8, D - 8 * pb);
      const double * lpp = Lt + (pb * NB - pb * (pb - 1) / 2) * 64;
      double xs[8];
#pragma unroll
      for(int c = 7; c >= 0; c--)
      {
        double v = 0.0;
        if(c < nloc)
        {
          v = x[8 * pb + c];
#pragma unroll
          for(int k = c + 1; k < 8; k++)
            if(k < nloc) v = fma(-lpp[swz(k, c)], xs[k], v);
          v *= invd[8 * pb + c];
        }
        xs[c] = v;
      }
      __syncwarp();
      if(lane == 0)
      {
#pragma unroll
        for(int c = 0; c < 8; c++) x[8 * pb + c] = xs[c];
      }
      for(int i = lane; i < 8 * pb; i += 32)
      {
        const int j = i >> 3;
        const double * tl = Lt + (j * NB - j * (j - 1) / 2 + pb - j) * 64;
        double v = x[i];
#pragma unroll
        for(int c = 0; c < 8; c++)
          if(c < nloc) v = fma(-tl[swz(c, i & 7)], xs[c], v);
        x[i] = v;
      }
      __syncwarp();
    }
  }
  cta_sync();
  int status = ok ? 0 : 2;
  if(bad) status = 2;
  if(status == 0 && too_few) status = 1;
  // ---- outputs + update (node.cpp:946-968) ----
  if(p.delta_out)
  {
    double * dout = p.delta_out + static_cast<size_t>(f) * p.dim_ref;
    for(int i = tid; i < p.dim_ref; i += THREADS) dout[i] = 0.0;
    cta_sync();
    for(int i = tid; i < D; i += THREADS)
    {
      const int c = i < p.theta_dim + p.phi_cols ? i : p.theta_dim + 2 * p.n + (i - p.theta_dim - p.phi_cols);
      dout[c] = status == 2 ? 0.0 : x[i];
    }
  }
  if(p.update_state && status == 0)
  {
    for(int i = tid; i < p.theta_dim; i += THREADS)
      p.theta_state[static_cast<size_t>(f) * p.theta_dim + i] += static_cast<float>(x[i]);
    if(p.beta_cols && p.beta)
      for(int i = tid; i < p.beta_cols; i += THREADS)
        p.beta[static_cast<size_t>(f) * p.beta_stride + i] += static_cast<float>(x[p.theta_dim + p.phi_cols + i]);
  }
  if(tid == 0) p.status[f] = status;
}

template<int NB>
__global__ void __launch_bounds__(THREADS, 4) ik_solve_mma_kernel(const IkSolveParams p)
{
  extern __shared__ __align__(16) unsigned char smem_mma[];
  switch(threadIdx.x >> 5)
  {
  case 0:
    solve_body<NB, 0>(p, smem_mma);
    break;
  case 1:
    solve_body<NB, 1>(p, smem_mma);
    break;
  case 2:
    solve_body<NB, 2>(p, smem_mma);
    break;
  default:
    solve_body<NB, 3>(p, smem_mma);
    break;
  }
}

template<int NB>
int launch_nb(const IkSolveParams & p, cudaStream_t st)
{
  const size_t smem = Cfg<NB>::off_e + static_cast<size_t>(4) * p.n * sizeof(float) + 16;
  if(smem > 227 * 1024) return fail(SMPLPP_ERR_INVALID, "IkTask", "IK problem too large for one CTA per frame");
  SB_CUDA(cudaFuncSetAttribute(ik_solve_mma_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  ik_solve_mma_kernel<NB><<<p.B, THREADS, smem, st>>>(p);
  SB_LAUNCHED();
  return SMPLPP_OK;
}
} // namespace

int sb::launch_ik_solve_mma(const IkSolveParams & p, cudaStream_t st, bool * handled)
{
  *handled = false;
  if(g_solve_variant == 1) return SMPLPP_OK;
  const bool qp = p.enable_qp && (p.phi_cols > 0 || p.beta_cols > 0) && p.a_ws != nullptr;
  if((!p.schur && qp) || p.phi_cols > 0 || p.ld > 88 || (p.ld & 3)) return SMPLPP_OK;
  const int nb = (p.D + 1 + 7) / 8;
  int rc;
  switch(nb)
  {
  case 6:
    rc = launch_nb<6>(p, st);
    break;
  case 7:
    rc = launch_nb<7>(p, st);
    break;
  case 10:
    rc = launch_nb<10>(p, st);
    break;
  case 11:
    rc = launch_nb<11>(p, st);
    break;
  default:
    return SMPLPP_OK;
  }
  *handled = rc == SMPLPP_OK;
  return rc;
}
