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
#include "tri_core.cuh"

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

// =============================================================================================
// Round-2 setup kernels (option "setup_variant" = 1, the default).  Same algorithms and the same substitutions as
// above (blocked right-looking Cholesky, dpotrs-style solves), re-organised so that a thread owns a ROW and keeps a
// 16-column slice of it in registers: no integer divisions in inner loops, two barriers per column in the diagonal
// block and none at all in the panel; alpha's two substitutions run over many CTAs instead of one.
// =============================================================================================
constexpr int kSlice = 16;                  // columns of a row held in registers

// A_JJ = L_JJ L_JJ^T in shared memory, one thread per row.  Left-looking over 16-column slices:
//   p[j] = A[i][c0+j] - sum_{m<c0} L[i][m] L[c0+j][m]      (registers, L[c0+j][m] is a broadcast load)
// then the 16 columns of the slice are finished right-looking inside the slice.
__global__ void __launch_bounds__(kBlk) chol_diag_rows_kernel(double* __restrict__ A, int n, int J, int* __restrict__ info) {
    extern __shared__ double S[];   // [128][kCholLd]
    const int r0 = J * kBlk, bs = min(kBlk, n - r0);
    const int i = threadIdx.x;
    // rows/columns >= bs are identity padding, so the loops below have compile-time bounds
#pragma unroll 16                                                  // 16 independent row loads in flight
    for (int r = 0; r < kBlk; ++r) {
        double v = (r == i) ? 1.0 : 0.0;
        if (r < bs && i < bs) v = (i <= r) ? A[(size_t)(r0 + r) * n + r0 + i] : 0.0;   // coalesced rows of the lower triangle
        S[r * kCholLd + i] = v;
    }
    __syncthreads();
    double* __restrict__ rowi = S + i * kCholLd;
    for (int c0 = 0; c0 < kBlk; c0 += kSlice) {
        double p[kSlice];
#pragma unroll
        for (int j = 0; j < kSlice; ++j) p[j] = rowi[c0 + j];
        if (i >= c0) {
            for (int m = 0; m < c0; ++m) {
                const double li = rowi[m];
#pragma unroll
                for (int j = 0; j < kSlice; ++j) p[j] = fma(-li, S[(c0 + j) * kCholLd + m], p[j]);
            }
        }
#pragma unroll
        for (int k = 0; k < kSlice; ++k) {
            const int g = c0 + k;
            if (i == g) {
                double d = p[k];
                if (!(d > 0.0)) {           // not positive definite (also catches NaN): GPy jitchol retries with jitter
                    atomicMin(info, r0 + g + 1);
                    d = 1.0;
                }
                rowi[g] = sqrt(d);
            }
            __syncthreads();
            if (i > g) {
                p[k] = p[k] / S[g * kCholLd + g];
                rowi[g] = p[k];
            }
            __syncthreads();
            if (i > g) {
#pragma unroll
                for (int j = k + 1; j < kSlice; ++j) p[j] = fma(-p[k], S[(c0 + j) * kCholLd + g], p[j]);
            }
        }
    }
    __syncthreads();
    for (int r = 0; r < bs; ++r)
        if (i <= r) A[(size_t)(r0 + r) * n + r0 + i] = S[r * kCholLd + i];
}

// panel rows below the diagonal block:  X L_JJ^T = A_IJ, one thread per row, 64 rows per CTA, no barriers in the
// substitution.  The CTA's rows are staged column-major in shared memory (conflict-free) and overwritten in place by
// the solution; the current 16 columns of a row live in registers.  All 128 threads load and store, 64 substitute.
constexpr int kPanelRows2 = 64;
constexpr int kPanelLd2 = kPanelRows2 + 1;

__global__ void __launch_bounds__(kBlk) chol_panel_rows_kernel(double* __restrict__ A, int n, int J) {
    extern __shared__ double sm[];
    double* Ls = sm;                          // [128][kCholLd] L_JJ (lower), row c = coefficients of column c of X
    double* Xs = sm + kBlk * kCholLd;         // [128][kPanelLd2] this CTA's rows, column-major
    const int c0g = J * kBlk;
    const int row0 = (J + 1) * kBlk + blockIdx.x * kPanelRows2;
    const int nrows = min(kPanelRows2, n - row0);
    const int t = threadIdx.x;
#pragma unroll 16
    for (int r = 0; r < kBlk; ++r) Ls[r * kCholLd + t] = (t <= r) ? A[(size_t)(c0g + r) * n + c0g + t] : 0.0;
#pragma unroll 16
    for (int r = 0; r < kPanelRows2; ++r) Xs[t * kPanelLd2 + r] = (r < nrows) ? A[(size_t)(row0 + r) * n + c0g + t] : 0.0;
    __syncthreads();
    if (t < nrows) {
        for (int c0 = 0; c0 < kBlk; c0 += kSlice) {
            double p[kSlice];
#pragma unroll
            for (int j = 0; j < kSlice; ++j) p[j] = Xs[(c0 + j) * kPanelLd2 + t];
            for (int m = 0; m < c0; ++m) {
                const double xm = Xs[m * kPanelLd2 + t];
#pragma unroll
                for (int j = 0; j < kSlice; ++j) p[j] = fma(-xm, Ls[(c0 + j) * kCholLd + m], p[j]);
            }
#pragma unroll
            for (int k = 0; k < kSlice; ++k) {
                const int g = c0 + k;
                p[k] = p[k] / Ls[g * kCholLd + g];
                Xs[g * kPanelLd2 + t] = p[k];
#pragma unroll
                for (int j = k + 1; j < kSlice; ++j) p[j] = fma(-p[k], Ls[(c0 + j) * kCholLd + g], p[j]);
            }
        }
    }
    __syncthreads();
    for (int r = 0; r < nrows; ++r) A[(size_t)(row0 + r) * n + c0g + t] = Xs[t * kPanelLd2 + r];   // coalesced row stores
}

// ---- alpha = L^-T (L^-1 y) as blocked substitution over many CTAs: per diagonal block one small solve kernel and
// one update kernel for the rest of the vector (dpotrs order of operations inside a block) ----
__global__ void __launch_bounds__(kBlk) solve_diag_kernel(const double* __restrict__ L, int n, int J, int transposed,
                                                          double* __restrict__ x) {
    // 128 unknowns as 4 warp blocks of 32: the owning warp solves its 32 x 32 triangle column by column with shuffles
    // (no CTA barrier), then the remaining warps subtract that block's contribution: 2 CTA barriers per warp block
    // instead of 2 per column.  Rows / columns >= bs are identity padding.
    extern __shared__ double S[];   // [128][kCholLd]
    __shared__ double xb[kBlk];
    const int r0 = J * kBlk, bs = min(kBlk, n - r0);
    const int i = threadIdx.x, warp = i >> 5, lane = i & 31;
#pragma unroll 16
    for (int r = 0; r < kBlk; ++r) {
        double v = (r == i) ? 1.0 : 0.0;
        if (r < bs && i < bs) v = (i <= r) ? L[(size_t)(r0 + r) * n + r0 + i] : 0.0;
        S[r * kCholLd + i] = v;
    }
    xb[i] = i < bs ? x[r0 + i] : 0.0;
    __syncthreads();
    if (!transposed) {
        for (int b = 0; b < 4; ++b) {
            if (warp == b) {
                double xi = xb[i];
                const double dii = S[i * kCholLd + i];
#pragma unroll 8
                for (int c = 0; c < 32; ++c) {
                    if (lane == c) xi = xi / dii;
                    const double xc = __shfl_sync(0xffffffffu, xi, c);
                    if (lane > c) xi = fma(-S[i * kCholLd + b * 32 + c], xc, xi);
                }
                xb[i] = xi;
            }
            __syncthreads();
            if (warp > b) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
                for (int c = 0; c < 32; c += 2) {
                    s0 = fma(S[i * kCholLd + b * 32 + c], xb[b * 32 + c], s0);
                    s1 = fma(S[i * kCholLd + b * 32 + c + 1], xb[b * 32 + c + 1], s1);
                }
                xb[i] -= s0 + s1;
            }
            __syncthreads();
        }
    } else {
        for (int b = 3; b >= 0; --b) {
            if (warp == b) {
                double xi = xb[i];
                const double dii = S[i * kCholLd + i];
#pragma unroll 8
                for (int c = 31; c >= 0; --c) {
                    if (lane == c) xi = xi / dii;
                    const double xc = __shfl_sync(0xffffffffu, xi, c);
                    if (lane < c) xi = fma(-S[(b * 32 + c) * kCholLd + i], xc, xi);   // (L^T)[i][c] = L[c][i]
                }
                xb[i] = xi;
            }
            __syncthreads();
            if (warp < b) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll 8
                for (int c = 0; c < 32; c += 2) {
                    s0 = fma(S[(b * 32 + c) * kCholLd + i], xb[b * 32 + c], s0);
                    s1 = fma(S[(b * 32 + c + 1) * kCholLd + i], xb[b * 32 + c + 1], s1);
                }
                xb[i] -= s0 + s1;
            }
            __syncthreads();
        }
    }
    if (i < bs) x[r0 + i] = xb[i];
}

// forward: x_i -= L[i][r0 .. r0+bs) . x_J for the rows below block J; one warp per row, lanes across the columns
__global__ void __launch_bounds__(256) solve_update_fwd_kernel(const double* __restrict__ L, int n, int J, double* __restrict__ x) {
    __shared__ double xb[kBlk];
    const int r0 = J * kBlk, bs = min(kBlk, n - r0);
    if (threadIdx.x < kBlk) xb[threadIdx.x] = threadIdx.x < bs ? x[r0 + threadIdx.x] : 0.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = r0 + bs + blockIdx.x * 8 + warp;
    if (i >= n) return;
    const double* __restrict__ lrow = L + (size_t)i * n + r0;
    double s = 0.0;
    for (int c = lane; c < bs; c += 32) s = fma(lrow[c], xb[c], s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) x[i] -= s;
}

// backward: x_i -= sum_c L[r0 + c][i] x_J[c] for the entries left of block J; one thread per i (coalesced rows of L)
__global__ void __launch_bounds__(256) solve_update_bwd_kernel(const double* __restrict__ L, int n, int J, double* __restrict__ x) {
    __shared__ double xb[kBlk];
    const int r0 = J * kBlk, bs = min(kBlk, n - r0);
    if (threadIdx.x < kBlk) xb[threadIdx.x] = threadIdx.x < bs ? x[r0 + threadIdx.x] : 0.0;
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= r0) return;
    double s0 = 0.0, s1 = 0.0;
    int c = 0;
    for (; c + 2 <= bs; c += 2) {
        s0 = fma(L[(size_t)(r0 + c) * n + i], xb[c], s0);
        s1 = fma(L[(size_t)(r0 + c + 1) * n + i], xb[c + 1], s1);
    }
    if (c < bs) s0 = fma(L[(size_t)(r0 + c) * n + i], xb[c], s0);
    x[i] -= s0 + s1;
}

// trailing update on the tensor pipe: 128 x 128 tiles (ti >= tj) of the rows / columns below block J,
//   A[ti][tj] -= A[ti][J] * A[tj][J]^T      (k extent 128, mma.sync m8n8k4 f64; 16 warps as 2 x 8, 64 x 16 warp tiles)
constexpr int kUpdLdA = 20, kUpdLdB = 132;   // bank-conflict-free fragment reads, see kernels_inv.cu

__global__ void __launch_bounds__(512, 1) chol_update_dmma_kernel(double* __restrict__ A, int n, int J) {
    const int tj = blockIdx.x, ti = blockIdx.y;
    if (tj > ti) return;
    __shared__ __align__(16) double As[kBlk * kUpdLdA];
    __shared__ __align__(16) double Bs[kChunkK * kUpdLdB];
    const int base = (J + 1) * kBlk, c0 = J * kBlk;
    const int i0 = base + ti * kBlk, j0 = base + tj * kBlk;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wm = warp >> 3, wn = warp & 7, lq = lane & 3, lr = lane >> 2;
    double acc[8][2][2];
#pragma unroll
    for (int mb = 0; mb < 8; ++mb)
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
    const int ar = tid >> 2, ak = (tid & 3) * 4;    // both operands: row ar of the tile, k ak .. ak+3 of the 16-k stage
    double ra[4], rb[4];
    auto fetch = [&](int step) {
        const int gi = i0 + ar, gj = j0 + ar;
        const double* __restrict__ pa = A + (size_t)gi * n + c0 + step * kChunkK + ak;
        const double* __restrict__ pb = A + (size_t)gj * n + c0 + step * kChunkK + ak;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            ra[q] = gi < n ? pa[q] : 0.0;
            rb[q] = gj < n ? pb[q] : 0.0;
        }
    };
    fetch(0);
#pragma unroll 1
    for (int step = 0; step < kChunksPerBlk; ++step) {
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            As[ar * kUpdLdA + ak + q] = ra[q];
            Bs[(ak + q) * kUpdLdB + ar] = rb[q];               // transposed: Bs[k][column of the output tile]
        }
        __syncthreads();
        if (step + 1 < kChunksPerBlk) fetch(step + 1);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            double av[8], bv[2];
#pragma unroll
            for (int mb = 0; mb < 8; ++mb) av[mb] = As[(wm * 64 + mb * 8 + lr) * kUpdLdA + ks * 4 + lq];
#pragma unroll
            for (int nb = 0; nb < 2; ++nb) bv[nb] = Bs[(ks * 4 + lq) * kUpdLdB + wn * 16 + nb * 8 + lr];
#pragma unroll
            for (int mb = 0; mb < 8; ++mb)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], av[mb], bv[nb]);
        }
    }
#pragma unroll
    for (int mb = 0; mb < 8; ++mb) {
        const int gi = i0 + wm * 64 + mb * 8 + lr;
        if (gi >= n) continue;
#pragma unroll
        for (int nb = 0; nb < 2; ++nb) {
            const int gj = j0 + wn * 16 + nb * 8 + 2 * lq;
            if (gj <= gi) A[(size_t)gi * n + gj] -= acc[mb][nb][0];
            if (gj + 1 <= gi) A[(size_t)gi * n + gj + 1] -= acc[mb][nb][1];
        }
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
int launch_cholesky(cudaStream_t s, double* A_dev, int n, int* info_dev, uint64_t* launches, int setup_variant) {
    static bool attr_done_dev[64] = {};          // function attributes are per device
    int attr_dev = 0;
    GPD_CUDA(cudaGetDevice(&attr_dev));
    bool& attr_done = attr_done_dev[attr_dev & 63];
    const int diag_smem = kBlk * kCholLd * (int)sizeof(double);
    const int panel_smem = (kBlk * kBlk + kBlk * kPanelLd) * (int)sizeof(double);
    const int panel_smem2 = (kBlk * kCholLd + kBlk * kPanelLd2) * (int)sizeof(double);
    if (!attr_done) {
        GPD_CUDA(cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
        GPD_CUDA(cudaFuncSetAttribute(chol_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, panel_smem));
        GPD_CUDA(cudaFuncSetAttribute(chol_diag_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
        GPD_CUDA(cudaFuncSetAttribute(chol_panel_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, panel_smem2));
        attr_done = true;
    }
    const int nblk = (n + kBlk - 1) / kBlk;
    for (int J = 0; J < nblk; ++J) {
        if (setup_variant >= 1) chol_diag_rows_kernel<<<1, kBlk, diag_smem, s>>>(A_dev, n, J, info_dev);
        else chol_diag_kernel<<<1, 256, diag_smem, s>>>(A_dev, n, J, info_dev);
        GPD_CUDA(cudaGetLastError());
        *launches += 1;
        const int below = n - (J + 1) * kBlk;
        if (below <= 0) break;
        if (setup_variant >= 1)
            chol_panel_rows_kernel<<<(below + kPanelRows2 - 1) / kPanelRows2, kBlk, panel_smem2, s>>>(A_dev, n, J);
        else
            chol_panel_kernel<<<(below + kPanelRows - 1) / kPanelRows, kPanelRows, panel_smem, s>>>(A_dev, n, J);
        GPD_CUDA(cudaGetLastError());
        if (setup_variant >= 1) {
            const int t2 = (below + kBlk - 1) / kBlk;
            chol_update_dmma_kernel<<<dim3(t2, t2), 512, 0, s>>>(A_dev, n, J);
        } else {
            const int t = (below + kUpd - 1) / kUpd;
            chol_update_kernel<<<dim3(t, t), 256, 0, s>>>(A_dev, n, J);
        }
        GPD_CUDA(cudaGetLastError());
        *launches += 2;
    }
    return 0;
}

int launch_chol_solve(cudaStream_t s, const double* L_dev, int n, const double* y_dev, double* x_dev, uint64_t* launches,
                      int setup_variant) {
    static bool attr_done_dev[64] = {};          // function attributes are per device
    int attr_dev = 0;
    GPD_CUDA(cudaGetDevice(&attr_dev));
    bool& attr_done = attr_done_dev[attr_dev & 63];
    const int smem = (kBlk * kCholLd + kBlk) * (int)sizeof(double);
    const int diag_smem = kBlk * kCholLd * (int)sizeof(double);
    if (!attr_done) {
        GPD_CUDA(cudaFuncSetAttribute(chol_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        GPD_CUDA(cudaFuncSetAttribute(solve_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, diag_smem));
        attr_done = true;
    }
    if (setup_variant >= 1) {
        const int nblk = (n + kBlk - 1) / kBlk;
        GPD_CUDA(cudaMemcpyAsync(x_dev, y_dev, sizeof(double) * n, cudaMemcpyDeviceToDevice, s));
        for (int J = 0; J < nblk; ++J) {                       // L t = y
            solve_diag_kernel<<<1, kBlk, diag_smem, s>>>(L_dev, n, J, 0, x_dev);
            GPD_CUDA(cudaGetLastError());
            *launches += 1;
            const int below = n - (J + 1) * kBlk;
            if (below > 0) {
                solve_update_fwd_kernel<<<(below + 7) / 8, 256, 0, s>>>(L_dev, n, J, x_dev);
                GPD_CUDA(cudaGetLastError());
                *launches += 1;
            }
        }
        for (int J = nblk - 1; J >= 0; --J) {                  // L^T alpha = t
            solve_diag_kernel<<<1, kBlk, diag_smem, s>>>(L_dev, n, J, 1, x_dev);
            GPD_CUDA(cudaGetLastError());
            *launches += 1;
            if (J > 0) {
                solve_update_bwd_kernel<<<(J * kBlk + 255) / 256, 256, 0, s>>>(L_dev, n, J, x_dev);
                GPD_CUDA(cudaGetLastError());
                *launches += 1;
            }
        }
        return 0;
    }
    *launches += 1;
    chol_solve_kernel<<<1, 1024, smem, s>>>(L_dev, n, y_dev, x_dev);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gpd
