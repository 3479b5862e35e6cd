// GP-state build on the device: covariance matrix, Cholesky factor and alpha of one exact GP.
//
// Reference: GPdoemd/models/gp_model.py:96-143 builds `GPRegression(Zr, Yr, kernx * kernp)` per (output, binary
// combination) and gp_load_hyp (:145-147) re-runs the inference whenever hyper-parameters are loaded; GPy's
// ExactGaussianInference then forms  Ky = K + (noise + 1e-8) I,  L = jitchol(Ky),  alpha = Ky^-1 y  (SURVEY.md
// Appendix A.3).  The hot path only needs L^-1 (packed) and alpha on the device, so building them here avoids the
// host LAPACK factorisation and the n^2 upload per GP (128 MB at n = 4000) on every (re)training.
//
// This is setup work, O(n^3 / 3) once per GP: plain FP64 FMA kernels, blocked right-looking Cholesky with
// 128-wide block columns:   for J:  factor A_JJ  ->  A_IJ <- A_IJ L_JJ^-T  ->  A_II' -= A_IJ A_I'J^T.
// The panel step is a true substitution (not a product with an explicit inverse) so the factor is backward stable
// like LAPACK's dpotrf; alpha comes from two blocked substitutions (dpotrs).
#include <cuda_runtime.h>

#include "common.cuh"
#include "radial.cuh"

namespace gpd {

// ---------------------------------------------------------------------------------------------
// Ky = var_x kappa(r_x) * var_p kappa(r_p) + diag_add I, lower triangle, row-major n x n
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) kmat_kernel(const KmatArgs a, const double* __restrict__ Z, double* __restrict__ A) {
    // 32 x 32 tile per CTA, tiles strictly above the diagonal exit
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj > ti) return;
    __shared__ double zi[32][kMaxXDims + kMaxP], zj[32][kMaxXDims + kMaxP];
    const int nd = a.nx + a.P;
    for (int k = threadIdx.x; k < 32 * nd; k += blockDim.x) {
        const int r = k / nd, d = k % nd;
        const int col = d < a.nx ? a.x_active[d] : a.D + (d - a.nx);
        const double ls = d < a.nx ? a.ls_x[d] : a.ls_p[d - a.nx];
        const int gi = ti * 32 + r, gj = tj * 32 + r;
        // GPy scales each argument by the lengthscale first (Stationary._scaled_dist, ARD branch)
        zi[r][d] = gi < a.n ? Z[(size_t)gi * a.dz + col] / ls : 0.0;
        zj[r][d] = gj < a.n ? Z[(size_t)gj * a.dz + col] / ls : 0.0;
    }
    __syncthreads();
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int rr = ty; rr < 32; rr += 8) {
        const int i = ti * 32 + rr, j = tj * 32 + tx;
        if (i >= a.n || j > i) continue;
        double r2x = 0.0, r2p = 0.0;
        for (int d = 0; d < a.nx; ++d) {
            const double df = zi[rr][d] - zj[tx][d];
            r2x = fma(df, df, r2x);
        }
        for (int d = 0; d < a.P; ++d) {
            const double df = zi[rr][a.nx + d] - zj[tx][a.nx + d];
            r2p = fma(df, df, r2p);
        }
        double kx, kp, d1, d2;
        radial_profile(a.kernel, sqrt(r2x), a.power_x, kx, d1, d2);
        radial_profile(a.kernel, sqrt(r2p), a.power_p, kp, d1, d2);
        double v = (a.var_x * kx) * (a.var_p * kp);
        if (i == j) v += a.diag_add;
        A[(size_t)i * a.n + j] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// factor the diagonal block A_JJ (bs x bs, bs <= 128) in shared memory; info = 1 + first bad pivot
// ---------------------------------------------------------------------------------------------
constexpr int kCholLd = kBlk + 1;

__global__ void __launch_bounds__(256) chol_diag_kernel(double* __restrict__ A, int n, int J, int* __restrict__ info) {
    extern __shared__ double S[];   // [bs][kCholLd]
    const int r0 = J * kBlk, bs = min(kBlk, n - r0);
    const int tid = threadIdx.x;
    for (int k = tid; k < bs * bs; k += blockDim.x) {
        const int i = k / bs, j = k % bs;
        S[i * kCholLd + j] = (j <= i) ? A[(size_t)(r0 + i) * n + r0 + j] : 0.0;
    }
    __syncthreads();
    for (int k = 0; k < bs; ++k) {
        if (tid == 0) {
            double d = S[k * kCholLd + k];
            if (!(d > 0.0)) {               // not positive definite (also catches NaN): GPy jitchol retries with jitter
                atomicMin(info, r0 + k + 1);
                d = 1.0;
            }
            S[k * kCholLd + k] = sqrt(d);
        }
        __syncthreads();
        const double piv = S[k * kCholLd + k];
        for (int i = k + 1 + tid; i < bs; i += blockDim.x) S[i * kCholLd + k] /= piv;
        __syncthreads();
        // trailing update of the lower triangle: rows i > k, columns k < j <= i
        const int m = bs - k - 1;
        for (int e = tid; e < m * m; e += blockDim.x) {
            const int i = k + 1 + e / m, j = k + 1 + e % m;
            if (j <= i) S[i * kCholLd + j] = fma(-S[i * kCholLd + k], S[j * kCholLd + k], S[i * kCholLd + j]);
        }
        __syncthreads();
    }
    for (int k = tid; k < bs * bs; k += blockDim.x) {
        const int i = k / bs, j = k % bs;
        if (j <= i) A[(size_t)(r0 + i) * n + r0 + j] = S[i * kCholLd + j];
    }
}

// ---------------------------------------------------------------------------------------------
// panel: rows below the diagonal block,  X L_JJ^T = A_IJ  by forward substitution along the columns
// ---------------------------------------------------------------------------------------------
constexpr int kPanelRows = 64;
constexpr int kPanelLd = kPanelRows + 1;

__global__ void __launch_bounds__(kPanelRows) chol_panel_kernel(double* __restrict__ A, int n, int J) {
    extern __shared__ double sm[];
    double* Ls = sm;                          // [128][128] L_JJ (lower)
    double* Xs = sm + kBlk * kBlk;            // [128][kPanelLd] this CTA's rows, column-major
    const int c0 = J * kBlk;
    const int row0 = (J + 1) * kBlk + blockIdx.x * kPanelRows;
    const int nrows = min(kPanelRows, n - row0);
    const int tid = threadIdx.x;
    for (int k = tid; k < kBlk * kBlk; k += blockDim.x) {
        const int c = k / kBlk, m = k % kBlk;
        Ls[k] = (m <= c) ? A[(size_t)(c0 + c) * n + c0 + m] : 0.0;
    }
    for (int k = tid; k < kPanelRows * kBlk; k += blockDim.x) {
        const int r = k / kBlk, c = k % kBlk;
        Xs[c * kPanelLd + r] = (r < nrows) ? A[(size_t)(row0 + r) * n + c0 + c] : 0.0;
    }
    __syncthreads();
    if (tid < nrows) {
        for (int c = 0; c < kBlk; ++c) {
            const double* __restrict__ lrow = Ls + c * kBlk;
            double s0 = Xs[c * kPanelLd + tid], s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int m = 0;
            for (; m + 4 <= c; m += 4) {
                s0 = fma(-Xs[(m + 0) * kPanelLd + tid], lrow[m + 0], s0);
                s1 = fma(-Xs[(m + 1) * kPanelLd + tid], lrow[m + 1], s1);
                s2 = fma(-Xs[(m + 2) * kPanelLd + tid], lrow[m + 2], s2);
                s3 = fma(-Xs[(m + 3) * kPanelLd + tid], lrow[m + 3], s3);
            }
            for (; m < c; ++m) s0 = fma(-Xs[m * kPanelLd + tid], lrow[m], s0);
            Xs[c * kPanelLd + tid] = ((s0 + s1) + (s2 + s3)) / lrow[c];
        }
    }
    __syncthreads();
    for (int k = tid; k < kPanelRows * kBlk; k += blockDim.x) {
        const int r = k / kBlk, c = k % kBlk;
        if (r < nrows) A[(size_t)(row0 + r) * n + c0 + c] = Xs[c * kPanelLd + r];
    }
}

// ---------------------------------------------------------------------------------------------
// trailing update: for 64 x 64 tiles (ti >= tj) of the rows/columns below block J:
//   A[ti][tj] -= A[ti][J] * A[tj][J]^T        (k extent 128)
// ---------------------------------------------------------------------------------------------
constexpr int kUpd = 64;

__global__ void __launch_bounds__(256) chol_update_kernel(double* __restrict__ A, int n, int J) {
    const int ti = blockIdx.y, tj = blockIdx.x;
    if (tj > ti) return;
    __shared__ double As[16][kUpd + 4], Bs[16][kUpd + 4];
    const int base = (J + 1) * kBlk;
    const int i0 = base + ti * kUpd, j0 = base + tj * kUpd, c0 = J * kBlk;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
    const int lr = tid >> 2, lk = (tid & 3) * 4;       // loader: row lr (0..63), k offset lk..lk+3
    for (int k0 = 0; k0 < kBlk; k0 += 16) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int gi = i0 + lr, gj = j0 + lr;
            As[lk + q][lr] = gi < n ? A[(size_t)gi * n + c0 + k0 + lk + q] : 0.0;
            Bs[lk + q][lr] = gj < n ? A[(size_t)gj * n + c0 + k0 + lk + q] : 0.0;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; ++kk) {
            double av[4], bv[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) av[a] = As[kk][ty * 4 + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) bv[b] = Bs[kk][tx * 4 + b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const int gi = i0 + ty * 4 + a;
        if (gi >= n) continue;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int gj = j0 + tx * 4 + b;
            if (gj <= gi) A[(size_t)gi * n + gj] -= acc[a][b];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// alpha = L^-T (L^-1 y): two blocked substitutions in ONE CTA (n^2 flops, reads L once per direction)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) chol_solve_kernel(const double* __restrict__ L, int n, const double* __restrict__ y,
                                                          double* __restrict__ x) {
    extern __shared__ double sm[];
    double* Ls = sm;                 // [128][kCholLd] current diagonal block
    double* xb = sm + kBlk * kCholLd;  // [128] solution of the current block
    const int tid = threadIdx.x, nt = blockDim.x;
    const int nblk = (n + kBlk - 1) / kBlk;
    for (int i = tid; i < n; i += nt) x[i] = y[i];
    __syncthreads();
    // ---- forward: L t = y ----
    for (int J = 0; J < nblk; ++J) {
        const int r0 = J * kBlk, bs = min(kBlk, n - r0);
        for (int k = tid; k < bs * bs; k += nt) {
            const int i = k / bs, j = k % bs;
            Ls[i * kCholLd + j] = (j <= i) ? L[(size_t)(r0 + i) * n + r0 + j] : 0.0;
        }
        if (tid < bs) xb[tid] = x[r0 + tid];
        __syncthreads();
        for (int c = 0; c < bs; ++c) {
            if (tid == c) xb[c] /= Ls[c * kCholLd + c];
            __syncthreads();
            if (tid > c && tid < bs) xb[tid] = fma(-Ls[tid * kCholLd + c], xb[c], xb[tid]);
            __syncthreads();
        }
        if (tid < bs) x[r0 + tid] = xb[tid];
        // rows below: x_i -= L[i][r0 .. r0+bs) . xb
        for (int i = r0 + bs + tid; i < n; i += nt) {
            const double* __restrict__ lrow = L + (size_t)i * n + r0;
            double s0 = 0.0, s1 = 0.0;
            int c = 0;
            for (; c + 2 <= bs; c += 2) {
                s0 = fma(lrow[c], xb[c], s0);
                s1 = fma(lrow[c + 1], xb[c + 1], s1);
            }
            if (c < bs) s0 = fma(lrow[c], xb[c], s0);
            x[i] -= s0 + s1;
        }
        __syncthreads();
    }
    // ---- backward: L^T alpha = t ----
    for (int J = nblk - 1; J >= 0; --J) {
        const int r0 = J * kBlk, bs = min(kBlk, n - r0);
        for (int k = tid; k < bs * bs; k += nt) {
            const int i = k / bs, j = k % bs;
            Ls[i * kCholLd + j] = (j <= i) ? L[(size_t)(r0 + i) * n + r0 + j] : 0.0;
        }
        if (tid < bs) xb[tid] = x[r0 + tid];
        __syncthreads();
        for (int c = bs - 1; c >= 0; --c) {
            if (tid == c) xb[c] /= Ls[c * kCholLd + c];
            __syncthreads();
            if (tid < c) xb[tid] = fma(-Ls[c * kCholLd + tid], xb[c], xb[tid]);   // (L^T)[tid][c] = L[c][tid]
            __syncthreads();
        }
        if (tid < bs) x[r0 + tid] = xb[tid];
        // columns to the left: x_i -= sum_c L[r0 + c][i] xb[c]   (coalesced over i)
        for (int i = tid; i < r0; i += nt) {
            double s0 = 0.0, s1 = 0.0;
            int c = 0;
            for (; c + 2 <= bs; c += 2) {
                s0 = fma(L[(size_t)(r0 + c) * n + i], xb[c], s0);
                s1 = fma(L[(size_t)(r0 + c + 1) * n + i], xb[c + 1], s1);
            }
            if (c < bs) s0 = fma(L[(size_t)(r0 + c) * n + i], xb[c], s0);
            x[i] -= s0 + s1;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
int launch_kmat(cudaStream_t s, const KmatArgs& a, const double* Z_dev, double* A_dev) {
    const int t = (a.n + 31) / 32;
    kmat_kernel<<<dim3(t, t), 256, 0, s>>>(a, Z_dev, A_dev);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// in-place lower Cholesky of the n x n row-major matrix A (lower triangle read and written); *info_dev = n + 1 on
// entry (by the caller), 1 + index of the first non-positive pivot on failure
int launch_cholesky(cudaStream_t s, double* A_dev, int n, int* info_dev, uint64_t* launches) {
    static bool attr_done_dev[64] = {};          // function attributes are per device
    int attr_dev = 0;
    GPD_CUDA(cudaGetDevice(&attr_dev));
    bool& attr_done = attr_done_dev[attr_dev & 63];
    const int diag_smem = kBlk * kCholLd * (int)sizeof(double);
    const int panel_smem = (kBlk * kBlk + kBlk * kPanelLd) * (int)sizeof(double);
    if (!attr_done) {
        GPD_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
        GPD_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, panel_smem));
        attr_done = true;
    }
    const int nblk = (n + kBlk - 1) / kBlk;
    for (int J = 0; J < nblk; ++J) {
        chol_diag_kernel<<<1, 256, diag_smem, s>>>(A_dev, n, J, info_dev);
        GPD_CUDA(cudaGetLastError());
        *launches += 1;
        const int below = n - (J + 1) * kBlk;
        if (below <= 0) break;
        chol_panel_kernel<<<(below + kPanelRows - 1) / kPanelRows, kPanelRows, panel_smem, s>>>(A_dev, n, J);
        GPD_CUDA(cudaGetLastError());
        const int t = (below + kUpd - 1) / kUpd;
        chol_update_kernel<<<dim3(t, t), 256, 0, s>>>(A_dev, n, J);
        GPD_CUDA(cudaGetLastError());
        *launches += 2;
    }
    return 0;
}

int launch_chol_solve(cudaStream_t s, const double* L_dev, int n, const double* y_dev, double* x_dev) {
    static bool attr_done_dev[64] = {};          // function attributes are per device
    int attr_dev = 0;
    GPD_CUDA(cudaGetDevice(&attr_dev));
    bool& attr_done = attr_done_dev[attr_dev & 63];
    const int smem = (kBlk * kCholLd + kBlk) * (int)sizeof(double);
    if (!attr_done) {
        GPD_CUDA(cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        attr_done = true;
    }
    chol_solve_kernel<<<1, 1024, smem, s>>>(L_dev, n, y_dev, x_dev);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gpd
