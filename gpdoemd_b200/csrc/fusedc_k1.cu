// concurrent evaluation + triangular contraction, kernel class 1 (see fusedc_impl.cuh); one translation unit per class
// so that the variants compile in parallel
#include "fusedc_impl.cuh"

namespace gpd {
GPD_DEFINE_FUSEDC_TU(1)
}  // namespace gpd
