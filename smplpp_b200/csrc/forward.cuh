// Internal launch helpers shared by forward.cu, blend_tc.cu and ik.cu.
#pragma once
#include "common.cuh"

namespace sb
{
struct ChainTopo
{
  int parent[kJoints];
  int depth[kJoints];
  int max_depth;
};

ChainTopo make_topo(const ModelDev & d);

// K1: rodrigues + pose features + joints + kinematic chain (one warp per frame)
// tc3_scratch (nullable, tc3_frame_operand_bytes(B) bytes): also write the stage images of K2'''
int launch_pose_chain(const ModelDev & d, cudaStream_t st, int B, const float * beta, long long beta_stride,
                      const float * theta, float * coef, float * xforms, float * joints, float * xforms44,
                      void * tc3_scratch = nullptr);
// K2 (FFMA): fused blend contraction (+ skinning when skin == true, else writes the rest shape)
int launch_blend_skin_ffma(const ModelDev & d, cudaStream_t st, int B, const float * coef, const float * xforms,
                           const float * theta, float * out, bool skin);
// K2' (tcgen05 split precision, blend_tc.cu): same contract as the skin == true FFMA kernel; tf32: 3xTF32, else 3xBF16
bool tc_blend_available();
int tc_prepare_model(ModelDev & d);
void tc_release_model(ModelDev & d);
size_t tc_coef_split_bytes(int64_t batch);
int launch_blend_skin_tc(const ModelDev & d, cudaStream_t st, int B, const float * coef, void * coef_split,
                         const float * xforms, const float * theta, float * out, bool tf32);
// K2'' (skin_tc.cu): fp16 split-precision blend AND skinning matrices on tcgen05 (any number of influences per vertex)
int tc2_prepare_model(ModelDev & d, float basis_max_abs);
void tc2_release_model(ModelDev & d);
size_t tc2_frame_operand_bytes(int64_t batch);
int launch_blend_skin_tc2(const ModelDev & d, cudaStream_t st, int B, const float * coef, const float * xforms, void * scratch,
                          const float * theta, float * out);
// K2''' (skin_tc3.cu): the same two GEMMs in a persistent CTA per SM; GEMM 1 of item i+1 runs under GEMM 2 + epilogue of item i
int tc3_prepare_model(ModelDev & d);
void tc3_release_model(ModelDev & d);
size_t tc3_frame_operand_bytes(int64_t batch);
// images_ready: the scratch area already holds the stage images (written by K1); else they are built from coef / xforms
int launch_blend_skin_tc3(const ModelDev & d, cudaStream_t st, int B, const float * coef, const float * xforms, void * scratch,
                          const float * theta, float * out, bool images_ready = false);
bool tc_encode_f16(void * out_map, void * base, uint64_t rows, uint64_t cols, uint32_t box_rows);
// K3: standalone skinning; affine: xforms are (B,24,3,4) else (B,24,4,4)
int launch_lbs(const ModelDev & d, cudaStream_t st, int B, const float * rest, const float * xforms, bool affine,
               const float * root, int root_stride, float * out);

// K3' (lbs_tma.cu): per-warp TMA pipelines; needs even V, 16-byte aligned tensors, affine (3x4) transforms
bool lbs_tma_usable(const ModelDev & d, const float * rest, const float * out, const float * xforms);
int launch_lbs_tma(const ModelDev & d, cudaStream_t st, int B, const float * rest, const float * xforms, const float * root,
                   int root_stride, float * out);
// K3'' (lbs_tc.cu): skinning matrices on tcgen05; needs the tc2 model data, even V >= 128, 16-byte aligned tensors
bool lbs_tc_usable(const ModelDev & d, const float * rest, const float * out, const float * xforms);
int launch_lbs_tc(const ModelDev & d, cudaStream_t st, int B, const float * rest, const float * xforms, int xf_floats,
                  const float * root, int root_stride, float * out);
extern std::atomic<int> g_lbs_variant; // 0: FFMA TMA pipeline when usable, 1: FFMA register-pipelined kernel, 2: tcgen05 (default)
// frees the streams, events and buffers of smplpp_forward_host (host_pipe.cu)
void release_host_pipe(smplpp_model * m);

extern std::atomic<int> g_forward_variant;
extern std::atomic<int> g_tc_grid_order; // 1: frame tiles fastest in the tcgen05 kernel's grid (CTAs in flight share the basis tile)
} // namespace sb
