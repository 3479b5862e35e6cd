// fused evaluation + triangular contraction, kernel class 4 (see fused_impl.cuh); one translation unit per class so
// that the variants compile in parallel
#include "fused_impl.cuh"

namespace gpd {
GPD_DEFINE_FUSED_TU(4)
}  // namespace gpd
