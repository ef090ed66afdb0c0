// VPoser decoder handle + internal launcher (shared with ik.cu).
#pragma once
#include "common.cuh"

struct smplpp_vposer
{
  float * w0 = nullptr;  // (512, 32)  decoder_net.0.weight, (out, in)
  float * b0 = nullptr;  // (512)
  float * w3t = nullptr; // (512 in, 512 out) transposed decoder_net.3.weight
  float * b3 = nullptr;  // (512)
  float * w5t = nullptr; // (512 in, 128 out-padded) transposed decoder_net.5.weight
  float * b5 = nullptr;  // (126)
  // tensor-core Jacobian (vposer_tc.cu): fp16 hi | lo stage images of W3 / W5, scaled fp32 W0^T, power-of-two scales
  void * tc_img_w3 = nullptr; // [half][K-block][part][256][32 fp16]
  void * tc_img_w5 = nullptr; // [K-block][part][128][32 fp16]
  float * tc_w0t = nullptr;   // (32, 512)
  float tc_conv_scale = 1.f, tc_out_scale = 1.f;
  int tc_sms = 0;
  bool tc_ready = false;
  mutable float * tc_aux = nullptr; // grow-only scratch of smplpp_vposer_decode (the IK step passes its own workspace)
  mutable size_t tc_aux_floats = 0;
};

namespace sb
{
// latent (B, 32) with row stride latent_stride -> aa (B, 63) with row stride aa_stride; jac (B, 63, 32) nullable
// aux: B * vposer_tc_aux_floats() floats of caller-owned scratch for the tensor-core Jacobian (nullable: the handle's own
// grow-only buffer is used, which restricts Jacobian calls on one handle to one stream at a time)
int launch_vposer_decode(const smplpp_vposer * vposer, cudaStream_t st, int B, const float * latent,
                         long long latent_stride, float * aa, long long aa_stride, float * jac, float * aux = nullptr);
// tensor-core Jacobian path (vposer_tc.cu)
int vposer_tc_prepare(smplpp_vposer & v, const float * w0, const float * w3, const float * w5);
void vposer_tc_release(smplpp_vposer & v);
size_t vposer_tc_aux_floats();
int launch_vposer_jac_tc(const smplpp_vposer & v, cudaStream_t st, int B, const float * aux, float * jac);
extern std::atomic<int> g_vposer_jac_variant; // 0: tensor cores when available, 1: FFMA kernel
} // namespace sb
