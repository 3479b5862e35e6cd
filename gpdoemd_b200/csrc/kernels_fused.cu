// Dispatch of the fused evaluation + contraction kernel over the kernel classes (fused_k0.cu .. fused_k5.cu).
#include "common.cuh"

namespace gpd {

#define GPD_DECL_F(K) int launch_fused_k##K(cudaStream_t s, int grid, int nx, int ncols, const FusedArgs& args);
GPD_DECL_F(0) GPD_DECL_F(1) GPD_DECL_F(2) GPD_DECL_F(3) GPD_DECL_F(4) GPD_DECL_F(5)
#undef GPD_DECL_F

int launch_fused(cudaStream_t s, int kernel, int grid, int nx, int ncols, const FusedArgs& args) {
    if (nx < 1 || nx > kMaxXDims || ncols < 1 || ncols > 11) return -1;     // no fused variant: the caller falls back
    switch (kernel) {
#define GPD_F_CASE(K) case K: return launch_fused_k##K(s, grid, nx, ncols, args);
        GPD_F_CASE(0) GPD_F_CASE(1) GPD_F_CASE(2) GPD_F_CASE(3) GPD_F_CASE(4) GPD_F_CASE(5)
#undef GPD_F_CASE
        default: return -1;
    }
}

}  // namespace gpd
