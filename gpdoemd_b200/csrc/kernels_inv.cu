// One-time per GP: X = L^-1 of the cached Cholesky factor on the FP64 tensor pipe (setup of K2, kernels_tri.cu).
//
// Reference: the factor is GPy's `posterior.woodbury_chol` (gp_model.py:199 reaches it through predict_noiseless);
// the reference solves against it with LAPACK dtrtrs per call.  Here it is inverted once so that the per-candidate
// solve becomes the triangular GEMM of tri_kernel.
//
// Algorithm: the 128 x 128 diagonal blocks are inverted by substitution (diag_inv_kernel, kernels_tri.cu); the
// off-diagonal part follows by recursive doubling over tile pairs,
//      [ L11  0  ]^-1   [  X11            0  ]
//      [ L21 L22 ]    = [ -X22 L21 X11   X22 ] ,
// level ts = 1, 2, 4, ... tiles: every pair of neighbouring ts-tile diagonal blocks is merged with two batched GEMMs
//      T   = L21 * X11      (X11 lower triangular: k tiles c .. end of the left half)
//      X21 = -X22 * T       (X22 lower triangular: k tiles start of the right half .. r)
// log2(nblk) levels, 2 launches each, n^3/3 flop in total on mma.sync m8n8k4 (DMMA).  T has no storage of its own: tile
// (r, c) of T is kept at the mirrored, strictly upper tile position (c, r) of X, which nothing else reads (the packing
// kernels read the lower tiles only).  The round-1 kernel (offdiag_kernel: block forward substitution on FP64 FMAs,
// (nblk-1) x 4 CTAs) took 70 ms at n = 4000; it stays behind option "setup_variant" = 0 for A/B tests.
#include <cuda_runtime.h>

#include "common.cuh"
#include "tri_core.cuh"

namespace gpd {

constexpr int kInvLdA = 20;     // As[128 rows][16 k + 4]: the 16 lanes of a half warp (4 rows x 4 k) hit 16 distinct banks
constexpr int kInvLdB = 132;    // Bs[16 k][128 cols + 4]: (4 k x 4 cols) likewise

struct InvArgs {
    const double* L;   // n x n row-major lower factor (rows >= n read as zero off the diagonal)
    int n;
    double* X;         // n_pad x n_pad, diagonal tiles hold the inverted diagonal blocks
    int n_pad, nblk, ts;
};

// MODE 1: T(r, c) = sum_kt L(r, kt) X(kt, c) -> stored at tile (c, r);  MODE 2: X(r, c) = -sum_kt X(r, kt) T(kt, c)
template <int MODE>
__global__ void __launch_bounds__(512, 1) inv_gemm_kernel(const InvArgs a) {
    const int ts = a.ts;
    const int c = blockIdx.x;
    // heavy tiles first: MODE 1 sums k tiles c .. half-1 (long for small c), MODE 2 half .. r (long for large r)
    const int r = MODE == 1 ? (int)blockIdx.y : a.nblk - 1 - (int)blockIdx.y;
    if ((c / ts) & 1) return;                  // c must lie in the left half of its pair
    if (r / ts != c / ts + 1) return;          // r in the right half of the same pair
    const int half = (c / ts + 1) * ts;        // first tile of the right half
    const int kt_lo = MODE == 1 ? c : half;
    const int kt_hi = MODE == 1 ? half - 1 : r;

    __shared__ __align__(16) double As[kBlk * kInvLdA];
    __shared__ __align__(16) double Bs[kChunkK * kInvLdB];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 3, wn = warp & 7;       // 2 row groups x 8 column groups: 64 x 16 warp tiles
    const int lq = lane & 3, lr = lane >> 2;
    double acc[8][2][2];
#pragma unroll
    for (int mb = 0; mb < 8; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;

    const int ar = tid >> 2, ak = (tid & 3) * 4;   // A stage: row ar, k ak .. ak+3
    const int bk = tid >> 5, bc = (tid & 31) * 4;  // B stage: k bk, columns bc .. bc+3
    double ra[4], rb[4];
    auto fetch = [&](int step) {
        const int kt = kt_lo + (step >> 3), k0 = (step & 7) * kChunkK;
        if (MODE == 1) {
            const int gi = r * kBlk + ar;
            const double* __restrict__ src = a.L + (size_t)gi * a.n + kt * kBlk + k0 + ak;   // columns < n (see header)
#pragma unroll
            for (int q = 0; q < 4; ++q) ra[q] = gi < a.n ? src[q] : 0.0;
        } else {
            const double2* __restrict__ src =
                reinterpret_cast<const double2*>(a.X + (size_t)(r * kBlk + ar) * a.n_pad + kt * kBlk + k0 + ak);
            const double2 u = src[0], v = src[1];
            ra[0] = u.x; ra[1] = u.y; ra[2] = v.x; ra[3] = v.y;
        }
        const double* bsrc = MODE == 1 ? a.X + (size_t)(kt * kBlk + k0 + bk) * a.n_pad + c * kBlk + bc
                                       : a.X + (size_t)(c * kBlk + k0 + bk) * a.n_pad + kt * kBlk + bc;
        const double2 u = reinterpret_cast<const double2*>(bsrc)[0], v = reinterpret_cast<const double2*>(bsrc)[1];
        rb[0] = u.x; rb[1] = u.y; rb[2] = v.x; rb[3] = v.y;
    };

    const int nsteps = (kt_hi - kt_lo + 1) * kChunksPerBlk;
    fetch(0);
#pragma unroll 1
    for (int step = 0; step < nsteps; ++step) {
        __syncthreads();                           // everyone is done with the previous stage
        *reinterpret_cast<double2*>(As + ar * kInvLdA + ak) = make_double2(ra[0], ra[1]);
        *reinterpret_cast<double2*>(As + ar * kInvLdA + ak + 2) = make_double2(ra[2], ra[3]);
        *reinterpret_cast<double2*>(Bs + bk * kInvLdB + bc) = make_double2(rb[0], rb[1]);
        *reinterpret_cast<double2*>(Bs + bk * kInvLdB + bc + 2) = make_double2(rb[2], rb[3]);
        __syncthreads();
        if (step + 1 < nsteps) fetch(step + 1);    // global loads of the next stage fly during the DMMAs
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double av[8], bv[2];
#pragma unroll
            for (int mb = 0; mb < 8; ++mb) av[mb] = As[(wm * 64 + mb * 8 + lr) * kInvLdA + ks * 4 + lq];
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) bv[nb] = Bs[(ks * 4 + lq) * kInvLdB + wn * 16 + nb * 8 + lr];
#pragma unroll
            for (int mb = 0; mb < 8; ++mb)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], av[mb], bv[nb]);
        }
    }
    double* out = MODE == 1 ? a.X + (size_t)(c * kBlk) * a.n_pad + r * kBlk : a.X + (size_t)(r * kBlk) * a.n_pad + c * kBlk;
    const double sgn = MODE == 1 ? 1.0 : -1.0;
#pragma unroll
    for (int mb = 0; mb < 8; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb)
            *reinterpret_cast<double2*>(out + (size_t)(wm * 64 + mb * 8 + lr) * a.n_pad + wn * 16 + nb * 8 + 2 * lq) =
                make_double2(sgn * acc[mb][nb][0], sgn * acc[mb][nb][1]);
}

// off-diagonal tiles of X = L^-1 given its inverted diagonal blocks; 2 * ceil(log2(nblk)) launches
int launch_invert_offdiag_dmma(cudaStream_t s, const double* L_dev, int n, int n_pad, double* X, uint64_t* launches) {
    const int nblk = n_pad / kBlk;
    InvArgs a;
    a.L = L_dev;
    a.n = n;
    a.X = X;
    a.n_pad = n_pad;
    a.nblk = nblk;
    for (int ts = 1; ts < nblk; ts *= 2) {
        a.ts = ts;
        inv_gemm_kernel<1><<<dim3(nblk, nblk), 512, 0, s>>>(a);
        GPD_CUDA(cudaGetLastError());
        inv_gemm_kernel<2><<<dim3(nblk, nblk), 512, 0, s>>>(a);
        GPD_CUDA(cudaGetLastError());
        *launches += 2;
    }
    return 0;
}

}  // namespace gpd
