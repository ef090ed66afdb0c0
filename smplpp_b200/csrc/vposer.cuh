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
};

namespace sb
{
// latent (B, 32) with row stride latent_stride -> aa (B, 63) with row stride aa_stride; jac (B, 63, 32) nullable
int launch_vposer_decode(const smplpp_vposer * vposer, cudaStream_t st, int B, const float * latent,
                         long long latent_stride, float * aa, long long aa_stride, float * jac);
} // namespace sb
