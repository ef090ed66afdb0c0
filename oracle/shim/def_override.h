// Test infrastructure (oracle build only): force-included in front of every reference TU together with
// -DDEF_H so that include/smplpp/definition/def.h (whose BATCH_SIZE / VERTEX_NUM are `static`, i.e. one
// private copy per translation unit, def.h:8-9) is replaced by ONE settable pair shared by all TUs.  The
// five constants are the values of def.h:10-14.  With BATCH_SIZE left at 1 the build behaves exactly like
// the stock reference; the harness raises it to run the same unmodified module sources batched.
#pragma once
#include <cstdint>

namespace smplpp
{
inline int64_t BATCH_SIZE = 1;
inline int64_t VERTEX_NUM = 6890;
constexpr int64_t JOINT_NUM = 24;
constexpr int64_t SHAPE_BASIS_DIM = 10;
constexpr int64_t POSE_BASIS_DIM = 207;
constexpr int64_t FACE_INDEX_NUM = 13776;
constexpr int64_t LATENT_DIM = 32;
} // namespace smplpp
