// K0 (parameter-part state) and K1 (fused x-kernel row + mean/derivative contraction + RHS panels).
//
// Reference arithmetic: GPy Stationary.K / gradients_X / gradients_XX as called from
// GPdoemd/models/gp_model.py:199,217,232-233,266-276 with the radial profiles of
// GPdoemd/kernels/stationary.py:33-99.
//
// Because the test point's parameter part p* = trans_p(pmean) is the same for every candidate
// (surrogate_model.py:175, :105), kp(p*,P_i) and its p-derivatives are candidate independent.  K0
// folds them, alpha and the x-variance into column-major coefficient tables; K1 then needs a single
// transcendental per (candidate, training point).
#include <stdlib.h>
#include "common.cuh"
#include "radial.cuh"

namespace gpd {

// ---------------------------------------------------------------------------------------------
// K0: per training row, the candidate-independent p-part and its derivatives
// ---------------------------------------------------------------------------------------------
__global__ void pstate_kernel(int kernel, int n, int n_pad, int P, int dim_z, const double* __restrict__ Zp,
                              const double* __restrict__ alpha, const double* __restrict__ pstar,
                              const double* __restrict__ ls_p, double var_x, double var_p, double power_p,
                              double* __restrict__ C, double* __restrict__ G, double* __restrict__ H) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pad) return;
    const int npair = P * (P + 1) / 2;
    if (i >= n) {
        for (int c = 0; c < 1 + P + npair; ++c) C[(size_t)c * n_pad + i] = 0.0;
        for (int c = 0; c < 1 + P; ++c) G[(size_t)c * n_pad + i] = 0.0;
        for (int c = 0; c < npair; ++c) H[(size_t)c * n_pad + i] = 0.0;
        return;
    }
    double d[kMaxP], l2[kMaxP];
    double r2 = 0.0;
    for (int a = 0; a < P; ++a) {
        double ls = ls_p[a];
        d[a] = pstar[a] - Zp[(size_t)i * dim_z + a];
        l2[a] = ls * ls;
        // difference first, then the scaling: when p* comes within 1e-5 of a training parameter the 1/r terms of
        // the second derivative (Exponential: kappa has a cusp at r = 0) need r to full relative accuracy; scaling
        // each argument first (GPy) or GPy's expanded-form distance lose u / r there (tests/test_d2mu_truth.py)
        double s = d[a] / ls;
        r2 += s * s;
    }
    double r = sqrt(r2);
    double k, dk, d2k;
    radial_profile(kernel, r, power_p, k, dk, d2k);
    const double invd = (r != 0.0) ? 1.0 / r : 0.0;
    const double invd2 = invd * invd;
    const double kp = var_p * k;
    double tmp1 = (var_p * dk) * invd;      // dK_dr * invdist
    const double w1 = tmp1;
    if (invd2 == 0.0) tmp1 -= var_p;        // GPy gradients_XX: (dK/dr)/r := -variance at r = 0
    const double tmp2 = (var_p * d2k) * invd2;
    const double al = alpha[i];
    C[i] = al * (var_x * kp);
    G[i] = var_x * kp;
    for (int a = 0; a < P; ++a) {
        double dkp = w1 * d[a] / l2[a];
        G[(size_t)(1 + a) * n_pad + i] = var_x * dkp;
        C[(size_t)(1 + a) * n_pad + i] = al * (var_x * dkp);
    }
    int pr = 0;
    for (int a = 0; a < P; ++a)
        for (int b = a; b < P; ++b, ++pr) {
            double v = ((tmp2 - tmp1 * invd2) * (d[a] * d[b]) / l2[a] + (a == b ? tmp1 : 0.0)) / l2[b];
            H[(size_t)pr * n_pad + i] = var_x * v;
            C[(size_t)(1 + P + pr) * n_pad + i] = al * (var_x * v);
        }
}

int launch_pstate(cudaStream_t s, int kernel, int n, int n_pad, int P, int dim_z, const double* Zp,
                  const double* alpha, const double* pstar, const double* ls_p, double var_x, double var_p,
                  double power_p, double* C, double* G, double* H) {
    int threads = 128, blocks = (n_pad + threads - 1) / threads;
    pstate_kernel<<<blocks, threads, 0, s>>>(kernel, n, n_pad, P, dim_z, Zp, alpha, pstar, ls_p, var_x, var_p,
                                             power_p, C, G, H);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// K1 dispatch: the kernel template lives in eval_impl.cuh and is instantiated per kernel class in
// eval_k0.cu .. eval_k5.cu (parallel compilation).
// ---------------------------------------------------------------------------------------------
#define GPD_DECL_K(K)                                                                                          \
    int launch_eval_k##K(cudaStream_t s, int mode, int nitems, int ncols, int nx, const GpView* gps,            \
                         const Item* items, const double* X, int ldx, int chunk_n, int nrhs, double* panels,   \
                         const double* beta, const RawOut* raws, int ldraw);
GPD_DECL_K(0) GPD_DECL_K(1) GPD_DECL_K(2) GPD_DECL_K(3) GPD_DECL_K(4) GPD_DECL_K(5)
#undef GPD_DECL_K

static int dispatch_kernel(cudaStream_t s, int mode, int kernel, int nx, int nitems, int ncols, const GpView* gps,
                           const Item* items, const double* X, int ldx, int chunk_n, int nrhs, double* panels,
                           const double* beta, const RawOut* raws, int ldraw) {
    switch (kernel) {
#define GPD_K_CASE(K) \
    case K: return launch_eval_k##K(s, mode, nitems, ncols, nx, gps, items, X, ldx, chunk_n, nrhs, panels, beta, raws, ldraw);
        GPD_K_CASE(0) GPD_K_CASE(1) GPD_K_CASE(2) GPD_K_CASE(3) GPD_K_CASE(4) GPD_K_CASE(5)
#undef GPD_K_CASE
        default:
            set_error("unknown kernel id");
            return 3;
    }
}

int launch_eval(cudaStream_t s, const GpView* gps_dev, const Item* items_dev, int nitems, const double* X_dev,
                int ldx, int chunk_n, int nrhs, int ncols, int kernel, int nx, double* scratch,
                const RawOut* raw_dev, int ldraw) {
    if (nitems == 0) return 0;
    return dispatch_kernel(s, 0, kernel, nx, nitems, ncols, gps_dev, items_dev, X_dev, ldx, chunk_n, nrhs, scratch,
                           nullptr, raw_dev, ldraw);
}

int launch_beta_contract(cudaStream_t s, const GpView* gps_dev, const Item* items_dev, int nitems,
                         const double* X_dev, int ldx, int chunk_n, int kernel, int nx, int P,
                         const double* beta_panels, const RawOut* raw_dev) {
    if (nitems == 0) return 0;
    return dispatch_kernel(s, 1, kernel, nx, nitems, P * (P + 1) / 2, gps_dev, items_dev, X_dev, ldx, chunk_n, 1,
                           nullptr, beta_panels, raw_dev, 0);
}

}  // namespace gpd
