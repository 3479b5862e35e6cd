// K1 instantiations for kernel class 0 (see eval_impl.cuh)
#include "eval_impl.cuh"
namespace gpd {
GPD_DEFINE_EVAL_TU(0)
}
