// Fused IK step (ik2.cu): internal launcher shared with ik.cu.
#pragma once
#include "common.cuh"
#include "tasks.cuh"
#include "vposer.cuh"

namespace sb
{
// dimensions of one IK problem in the tile-aligned layout of the fused kernel:
// unknowns [theta (padded to 4) | phi 2n (padded to 4) | beta 10 (padded to 12)], padding unknowns are identity rows
struct Ik2Dims
{
  int theta_dim, thp, phi_cols, php, beta_cols, Dp, ldf, ld;
};
Ik2Dims ik2_dims(int n, bool vposer, bool phi, bool beta);
size_t ik2_qp_ws_doubles(const Ik2Dims & d); // doubles per frame of the active-set QP's copy of A
size_t ik2_rec_bytes(int64_t batch, int n);  // per-frame attachment records

struct Ik2Call
{
  const smplpp_model * model = nullptr;
  const smplpp_vposer * vposer = nullptr;
  const smplpp_tasks * tasks = nullptr;
  const smplpp_ik_options * opt = nullptr;
  cudaStream_t st = nullptr;
  int B = 0;
  bool schur = false;
  const float * theta75 = nullptr; // (B, 75) assembled theta (theta_state itself without VPoser)
  float * theta_state = nullptr;
  float * beta = nullptr;
  long long beta_stride = 0;
  float * vertex_weights = nullptr;
  const float * target_pos = nullptr;
  const float * target_normal = nullptr;
  const float * pos_task_weight = nullptr;
  const float * vjac = nullptr;          // (B, 63, 32) decoder Jacobian
  const TaskRec * frame_recs = nullptr;  // (B, n) per-frame attachments; null: the task set's records
  const TaskSkin * frame_skins = nullptr; // (B, n) skinning rows of frame_recs (models with <= 4 influences), else null
  int32_t * status = nullptr;
  float * e_out = nullptr;
  float * j_out = nullptr;
  double * a_out = nullptr;
  double * b_out = nullptr;
  double * delta_out = nullptr;
  float * dphi_out = nullptr;
  double * a_ws = nullptr;
  double * schur_out = nullptr;
  double * factor_ws = nullptr;
};
int launch_ik_fused(const Ik2Call & c);
// one record from the host copies of the model topology; -1 when the 1-rings exceed the record limits
int build_task_rec_host(const smplpp_model * model, int64_t face, TaskRec & rec);
void build_task_skin_host(const smplpp_model * model, const TaskRec & rec, TaskSkin & skin); // model kmax <= 4
// (total) records from per-frame face indices on the device (total = frames * tasks)
// skins (nullable): also the skinning rows of every record (d.kmax <= 4)
int launch_task_topo(const ModelDev & d, cudaStream_t st, long long total, const int32_t * face_idx, TaskRec * out, TaskSkin * skins);
size_t ik2_skin_bytes(int64_t batch, int n);
extern std::atomic<int> g_ik_variant; // 0: auto (two kernels when the frames share the attachments, fused kernel otherwise), 1: two kernels, 2: fused
} // namespace sb
