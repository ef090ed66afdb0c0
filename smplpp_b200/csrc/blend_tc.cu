// K2 (tcgen05): placeholder until the 3xTF32 tensor-core variant lands; the FFMA kernel is the default path.
#include "forward.cuh"

namespace sb
{
bool tc_blend_available()
{
  return false;
}

int launch_blend_skin_tc(const ModelDev &, cudaStream_t, int, const float *, const float *, const float *, float *)
{
  return fail(SMPLPP_ERR_INVALID, "SMPL", "tcgen05 blend variant is not built");
}
} // namespace sb
