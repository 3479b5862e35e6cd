// Dispatch of the fused evaluation + contraction kernel over the kernel classes (fused_k0.cu .. fused_k5.cu).
#include "common.cuh"

namespace gpd {

#define GPD_DECL_F(K) int launch_fused_k##K(cudaStream_t s, int grid, int nx, int ncols, const FusedArgs& args);
GPD_DECL_F(0) GPD_DECL_F(1) GPD_DECL_F(2) GPD_DECL_F(3) GPD_DECL_F(4) GPD_DECL_F(5)
#undef GPD_DECL_F
#define GPD_DECL_FC(K) int launch_fusedc_k##K(cudaStream_t s, int grid, int nx, int ncols, const FusedArgs& args);
GPD_DECL_FC(0) GPD_DECL_FC(1) GPD_DECL_FC(2) GPD_DECL_FC(3) GPD_DECL_FC(4) GPD_DECL_FC(5)
#undef GPD_DECL_FC

// option "fused" = 2: producer warpgroup inside the persistent contraction CTA (fusedc_impl.cuh)
int launch_fusedc(cudaStream_t s, int kernel, int grid, int nx, int ncols, const FusedArgs& args) {
    if (nx < 1 || nx > kMaxXDims || ncols < 1 || ncols > 11) return -1;
    switch (kernel) {
#define GPD_FC_CASE(K) case K: return launch_fusedc_k##K(s, grid, nx, ncols, args);
        GPD_FC_CASE(0) GPD_FC_CASE(1) GPD_FC_CASE(2) GPD_FC_CASE(3) GPD_FC_CASE(4) GPD_FC_CASE(5)
#undef GPD_FC_CASE
        default: return -1;
    }
}

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
