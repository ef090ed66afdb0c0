// Pose-blend columns of the IK Jacobian on tcgen05 (ik_poseblend_tc.cu): per-task-set operand images and the launcher.
#pragma once
#include <vector>

#include "common.cuh"

namespace sb
{
struct TasksDev;
struct PoseBlendTc
{
  bool ready = false;
  const uint8_t * img = nullptr;      // [K-block][hi | lo][224 basis columns][32 fp16]: P_m^T stage images, SWIZZLE_64B
  const int32_t * slot_off = nullptr; // (n + 1) first K-block of every task
  int slots = 0;                      // K-blocks of all tasks; the per-frame CA buffer holds slots * 128 floats
  int basis_exp = 0;                  // images are scaled by 2^basis_exp
  // rest shape of the task vertices (ik_restshape_tc_kernel): [tile of 224 coordinates][K-block][hi | lo][224][32 fp16]
  bool rest_ready = false;
  const uint8_t * rest_img = nullptr;
  const float * rest_tmpl = nullptr;  // template coordinate of every row (added in fp32)
};
// builds the images from the task set's compact basis rows; leaves out.ready = false where the kernel does not apply
// (not sm_100, a task with more than 21 vertices)
int poseblend_tc_prepare(const std::vector<float> & basis, const std::vector<int32_t> & pair_off,
                         const std::vector<int32_t> & pair_vert, PoseBlendTc & out, std::vector<void *> & allocations);
// J[f, 4 m + r, 3 + 3 k + c] += sum_e (CA_m P_m)[r, 9 (k - 1) + e] dr[f, 27 (k - 1) + 9 c + e]  (k = 1..23), and the shape
// columns 207..216 of CA_m P_m added to J[.., beta_col + i] when beta_col >= 0
int launch_poseblend_tc(const PoseBlendTc & pb, const TasksDev & t, cudaStream_t st, int B, int rows, int use_ring, int beta_col,
                        const float * ca, const float * dr, float * J, int ld);
// rest[f][3 u + a] = T + basis . coef for the first n_vertices task vertices (corners first)
int launch_restshape_tc(const PoseBlendTc & pb, cudaStream_t st, int B, int n_vertices, const float * coef, float * rest);
extern std::atomic<int> g_poseblend_variant; // 0: auto (tensor cores where prepared), 1: the FFMA phase inside ik_jacobian_kernel, 2: tensor-core columns, FFMA rest shape
} // namespace sb
