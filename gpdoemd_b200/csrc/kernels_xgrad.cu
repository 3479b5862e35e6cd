// Design gradients: the first-order Taylor marginal (mu, S) of one model and its derivative with respect to the
// continuous design variables x* (the caller side of demos/case_study.py:279-293 refines the
// grid argmax; the reference itself has no such gradient).
//
//   k*_i(x) = kappa(r_i) G0_i,  r_i = |(x - X_i) / l|,   d kappa_i / d x_d = kappa'(r_i) u_d / (r_i l_d),  u = (x - X_i) / l
//   mu_e     = sum_i kappa_i C_0i                 d mu_e / dx_d   = sum_i dkappa_i/dx_d C_0i
//   J_ea     = sum_i kappa_i C_(1+a)i             d J_ea / dx_d   = sum_i dkappa_i/dx_d C_(1+a)i
//   s2_e     = k** - |v|^2, v = L^-1 k*           d s2_e / dx_d   = -2 v . w_d,  w_d = L^-1 (dk*/dx_d)
//   S        = diag(s2) + J Sigma J^T             dS/dx_d = diag(ds2/dx_d) + dJ Sigma J^T + J Sigma dJ^T
//
// The 1 + nx right-hand sides (k*, dk*/dx_1..nx) go through the SAME tri_kernel / gram_kernel as the variance
// derivatives with respect to p (row 0 of the Gram matrix holds |v|^2 and v . w_d).  Refinement batches are small
// (tens to thousands of designs), so the evaluation below is a plain kernel: one thread per candidate, run-time loops,
// the library's radial profiles (radial.cuh) instead of the table exponential of the hot evaluation kernel.
#include <math.h>

#include "common.cuh"
#include "radial.cuh"

namespace gpd {


// grid (tiles, nGP), 128 threads: thread t = candidate t of the tile.  acc_out: [gp][chunk_n][(1+nx)*(1+P)]
__global__ void __launch_bounds__(kBlk) xgrad_eval_kernel(const GpView* __restrict__ gps, const double* __restrict__ X,
                                                          int ldx, int chunk_n, double* __restrict__ panels,
                                                          double* __restrict__ acc_out) {
    const GpView& g = gps[blockIdx.y];
    const int tile = blockIdx.x, t = threadIdx.x;
    const int cand = tile * kBlk + t;
    const bool valid = cand < chunk_n;
    const int nx = g.nx, P = g.P, n_pad = g.n_pad, nk = 1 + nx, ncol = 1 + P;
    const double cs = (g.kernel == 0) ? kRbfCoordScale : 1.0;      // Xs and xscale carry kRbfCoordScale for the RBF kernel
    double xs[kMaxXDims];
    for (int d = 0; d < nx; ++d)
        xs[d] = valid ? (X[(size_t)cand * ldx + g.x_active[d]] - g.xmin[d]) * g.xscale[d] / cs : 0.0;   // = trans_x(x)/l
    double acc[(1 + kMaxXDims) * (1 + kMaxP)];
    for (int k = 0; k < nk * ncol; ++k) acc[k] = 0.0;
    double* pan = panels + g.panel_off_w + (size_t)tile * nk * n_pad * kBlk;
    for (int i = 0; i < n_pad; ++i) {
        double u[kMaxXDims], r2 = 0.0;
        for (int d = 0; d < nx; ++d) {
            u[d] = xs[d] - g.Xs[(size_t)i * nx + d] / cs;
            r2 = fma(u[d], u[d], r2);
        }
        const double r = sqrt(r2);
        double k, dk, d2k;
        radial_profile(g.kernel, r, g.power_x, k, dk, d2k);
        const double w = (r != 0.0) ? dk / r : 0.0;                  // GPy gradients_X: 1/r := 0 at r = 0
        const double g0 = g.G[i];                                    // var_x * kp_i (0 on padding rows)
        const int pos = t ^ ((i & 3) << 2);
        pan[((size_t)0 * n_pad + i) * kBlk + pos] = valid ? k * g0 : 0.0;
        double dkx[kMaxXDims];
        for (int d = 0; d < nx; ++d) {
            dkx[d] = w * u[d] / g.ls[d];                              // d kappa / d x_d (transformed x in [0, 1])
            pan[((size_t)(1 + d) * n_pad + i) * kBlk + pos] = valid ? dkx[d] * g0 : 0.0;
        }
        for (int c = 0; c < ncol; ++c) {
            const double cc = g.C[(size_t)c * n_pad + i];
            acc[c] = fma(k, cc, acc[c]);
            for (int d = 0; d < nx; ++d) acc[(1 + d) * ncol + c] = fma(dkx[d], cc, acc[(1 + d) * ncol + c]);
        }
    }
    if (!valid) return;
    double* o = acc_out + ((size_t)blockIdx.y * chunk_n + cand) * (size_t)(nk * ncol);
    for (int k = 0; k < nk * ncol; ++k) o[k] = acc[k];
}

__global__ void __launch_bounds__(128) xgrad_finalize_kernel(const XgradArgs a) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.chunk_n) return;
    const int E = a.E, P = a.P, D = a.D, nx = a.nx, ncol = 1 + P, nk = 1 + nx;
    int r = 0;
    for (int v = 0; v < a.nb; ++v)
        if (a.X[(size_t)n * a.ldx + a.bcols[v]] - 0.5 > 0.0) r += a.weights[v];
    const ModelXform* xf = a.xf;
    // J[e][p], dJ[e][p][d] in original units; kept in local memory (E <= 8, P <= 16, nx <= 8)
    double J[kMaxE][kMaxP], s2[kMaxE], ds2[kMaxE][kMaxXDims];
    for (int e = 0; e < E; ++e) {
        const int gi = e * a.R + r;
        const GpView& g = a.gps[gi];
        const double ystd = xf->ystd[e];
        if (g.n == 0) {                      // no GP for this binary combination: the reference leaves zeros
            a.mu[(size_t)n * E + e] = xf->ymean[e];
            s2[e] = 0.0;
            for (int p = 0; p < P; ++p) J[e][p] = 0.0;
            for (int d = 0; d < nx; ++d) ds2[e][d] = 0.0;
            for (int d = 0; d < D; ++d) a.dmu[((size_t)n * E + e) * D + d] = 0.0;
            continue;
        }
        const double* c = a.acc + ((size_t)gi * a.chunk_n + n) * (size_t)(nk * ncol);
        const double* gr = a.gram + ((size_t)gi * a.chunk_n + n) * a.ngram;
        a.mu[(size_t)n * E + e] = xf->ymean[e] + c[0] * ystd;
        s2[e] = fmax(g.kss - gr[0], xf->var_floor) * (ystd * ystd);
        for (int p = 0; p < P; ++p) J[e][p] = c[1 + p] * ystd / xf->pdiff[p];
        for (int d = 0; d < D; ++d) a.dmu[((size_t)n * E + e) * D + d] = 0.0;
        for (int d = 0; d < nx; ++d) {
            const double inv_dx = 1.0 / g.xdiff[d];                   // d/dx_original = (1 / (xmax - xmin)) d/dx_transformed
            a.dmu[((size_t)n * E + e) * D + a.x_active[d]] = c[(1 + d) * ncol] * ystd * inv_dx;
            ds2[e][d] = -2.0 * gr[1 + d] * (ystd * ystd) * inv_dx;    // gram row 0: v.v, v.w_1, ..., v.w_nx
        }
    }
    // S and dS/dx
    for (int e1 = 0; e1 < E; ++e1)
        for (int e2 = 0; e2 < E; ++e2) {
            double v = 0.0;
            for (int p = 0; p < P; ++p) {
                double tq = 0.0;
                for (int q = 0; q < P; ++q) tq = fma(xf->Sigma[p * P + q], J[e2][q], tq);
                v = fma(J[e1][p], tq, v);
            }
            bool clipped = false;
            if (e1 == e2) {
                const double raw = s2[e1] + v;
                clipped = !(raw > 1e-15);
                v = fmax(1e-15, raw);                                   // taylor_first_order.py:41
            }
            a.S[((size_t)n * E + e1) * E + e2] = v;
            for (int d = 0; d < D; ++d) a.dS[(((size_t)n * E + e1) * E + e2) * D + d] = 0.0;
            if (clipped) continue;                                      // the clip is flat
            for (int d = 0; d < nx; ++d) {
                const GpView& g1 = a.gps[e1 * a.R + r];
                const GpView& g2 = a.gps[e2 * a.R + r];
                double dv = 0.0;
                if (g1.n && g2.n) {
                    const double* c1 = a.acc + ((size_t)(e1 * a.R + r) * a.chunk_n + n) * (size_t)(nk * ncol) + (1 + d) * ncol;
                    const double* c2 = a.acc + ((size_t)(e2 * a.R + r) * a.chunk_n + n) * (size_t)(nk * ncol) + (1 + d) * ncol;
                    const double i1 = xf->ystd[e1] / g1.xdiff[d], i2 = xf->ystd[e2] / g2.xdiff[d];
                    for (int p = 0; p < P; ++p) {
                        double t1 = 0.0, t2 = 0.0;                      // (Sigma J_e2)_p and (Sigma dJ_e2)_p
                        for (int q = 0; q < P; ++q) {
                            t1 = fma(xf->Sigma[p * P + q], J[e2][q], t1);
                            t2 = fma(xf->Sigma[p * P + q], c2[1 + q] * i2 / xf->pdiff[q], t2);
                        }
                        dv = fma(c1[1 + p] * i1 / xf->pdiff[p], t1, dv);
                        dv = fma(J[e1][p], t2, dv);
                    }
                }
                if (e1 == e2) dv += ds2[e1][d];
                a.dS[(((size_t)n * E + e1) * E + e2) * D + a.x_active[d]] = dv;
            }
        }
}

int launch_xgrad_eval(cudaStream_t s, const GpView* gps_dev, int ngp, int tiles, const double* X_dev, int ldx, int chunk_n,
                      double* panels, double* acc) {
    xgrad_eval_kernel<<<dim3(tiles, ngp), kBlk, 0, s>>>(gps_dev, X_dev, ldx, chunk_n, panels, acc);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

int launch_xgrad_finalize(cudaStream_t s, const XgradArgs& a) {
    xgrad_finalize_kernel<<<(a.chunk_n + 127) / 128, 128, 0, s>>>(a);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gpd
