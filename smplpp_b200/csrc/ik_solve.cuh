// Parameters of the IK normal-equation / solve kernels (ik.cu: ik_solve_kernel, ik_solve_mma.cu: ik_solve_mma_kernel).
#pragma once
#include "common.cuh"

namespace sb
{
struct IkSolveParams
{
  int B, n, rows_per_task; // rows actually populated per task (3 or 4)
  int theta_dim, phi_cols, beta_cols, D; // D = theta_dim + phi_cols + beta_cols (compact)
  int ld;                  // row stride of J
  int rows_off;            // byte offset of the row staging area [2][4][ld] floats in the dynamic shared memory
  int vposer, enable_qp, skip_if_too_few, update_state;
  int schur;               // shared-beta stage: eliminate only the first D - beta_cols pivots
  float reg_theta, reg_phi, reg_beta, phi_limit, beta_limit, latent_reg, hand_reg;
  const float * J;   // (B, 4n, ld)
  const float * e;   // (B, 4n)
  const int * frame_info;
  float * theta_state; // (B, theta_dim) in/out
  float * beta;        // (B, 10) in/out when beta_cols && !schur
  long long beta_stride;
  int * status;        // (B)
  // optional outputs in the reference layout dim_ref = theta_dim + 2n + (beta ? 10 : 0)
  int dim_ref;
  double * a_out;
  double * b_out;
  double * delta_out;
  double * a_ws;       // (B, D(D+1)/2) preserved A for the active-set QP (null when no bound can bind)
  // shared-beta stage
  double * schur_out;  // (B, 111): S (10x10 row-major) | r (10) | ||e||^2
  double * factor_ws;  // (B, P) packed factor rows kept for the apply step
};

// J'J on the fp64 tensor cores + blocked Cholesky (ik_solve_mma.cu).  Returns SMPLPP_OK and sets *handled when the problem
// shape is one the kernel covers (D + 1 <= 88 unknowns incl. the right-hand side row, no bound can bind); otherwise
// *handled = false and nothing was launched (the caller falls back to ik_solve_kernel).
int launch_ik_solve_mma(const IkSolveParams & p, cudaStream_t st, bool * handled);
extern std::atomic<int> g_solve_variant; // 0: auto (tensor-core kernel where it applies), 1: ik_solve_kernel always
} // namespace sb
