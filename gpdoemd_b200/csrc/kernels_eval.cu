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
#include "common.cuh"

namespace gpd {

// ---------------------------------------------------------------------------------------------
// radial profiles  kappa(r) (variance excluded), first and second derivative
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void radial_profile(int kernel, double r, double power, double& k, double& dk,
                                               double& d2k) {
    switch (kernel) {
        case 0: {  // RBF: GPy rbf.py, dK2_drdr = (r^2-1) k
            k = exp(-0.5 * r * r);
            dk = -r * k;
            d2k = (r * r - 1.0) * k;
        } break;
        case 1: {  // Exponential, stationary.py:46-47
            k = exp(-r);
            dk = -k;
            d2k = k;
        } break;
        case 2: {  // Matern32, stationary.py:57-59
            const double s3 = 1.7320508075688772;
            double e = exp(-s3 * r);
            k = (1.0 + s3 * r) * e;
            dk = -3.0 * r * e;
            d2k = 3.0 * (s3 * r - 1.0) * e;
        } break;
        case 3: {  // Matern52, stationary.py:69-71
            const double s5 = 2.23606797749979;
            double e = exp(-s5 * r);
            k = (1.0 + s5 * r + 5.0 / 3.0 * r * r) * e;
            dk = (10.0 / 3.0 * r - 5.0 * r - 5.0 * s5 / 3.0 * r * r) * e;
            double ar = s5 * r;
            d2k = 5.0 / 3.0 * (ar * ar - ar - 1.0) * e;
        } break;
        case 4: {  // Cosine, stationary.py:81-82
            k = cos(r);
            dk = -sin(r);
            d2k = -k;
        } break;
        default: {  // RatQuad, stationary.py:92-99
            double r2 = r * r;
            double lp = log1p(0.5 * r2);
            k = exp(-power * lp);
            double a = (power + 1.0) * lp;
            dk = -power * r * exp(-a);
            double dr1 = -power * exp(-a);
            double dr2 = power * (power + 1.0) * r2 * exp(-(a + lp));
            d2k = dr1 + dr2;
        } break;
    }
}

template <int KERN>
__device__ __forceinline__ double kappa_of_r2(double r2, double power) {
    if (KERN == 0) return exp(-0.5 * r2);
    double r = sqrt(r2);
    if (KERN == 1) return exp(-r);
    if (KERN == 2) {
        const double s3 = 1.7320508075688772;
        return (1.0 + s3 * r) * exp(-s3 * r);
    }
    if (KERN == 3) {
        const double s5 = 2.23606797749979;
        return (1.0 + s5 * r + 5.0 / 3.0 * (r * r)) * exp(-s5 * r);
    }
    if (KERN == 4) return cos(r);
    return exp(-power * log1p(0.5 * r2));
}

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
        double s = pstar[a] / ls - Zp[(size_t)i * dim_z + a] / ls;   // GPy scales each argument first
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
// K1: one thread per candidate of a 128-wide tile, training rows staged through shared memory.
//   MODE 0: acc[c] += kx * C[c][i]            and write RHS panels  kx * G[a][i]  (swizzled, layout.h)
//   MODE 1: acc[c] += beta[i][t] * kx * H[c][i]   (second-order variance term, beta read from a panel)
// ---------------------------------------------------------------------------------------------
constexpr int kEvalRows = 32;   // training rows per shared-memory stage

template <int KERN, int NX, int NCB, int MODE>
__global__ void __launch_bounds__(kBlk) eval_kernel(const GpView* __restrict__ gps, const Item* __restrict__ items,
                                                    const double* __restrict__ X, int ldx, int chunk_n, int nrhs,
                                                    int ncols, double* __restrict__ panels,
                                                    const double* __restrict__ beta_panels,
                                                    const RawOut* __restrict__ raws, int ldraw) {
    __shared__ double sX[kEvalRows * NX];
    __shared__ double sC[NCB * kEvalRows];
    __shared__ double sG[(MODE == 0 ? (1 + kMaxP) : 1) * kEvalRows];

    const Item it = items[blockIdx.x];
    const GpView& g = gps[it.gp];
    const int count = g.count ? *g.count : chunk_n;
    const int slot0 = it.tile * kBlk;
    if (slot0 >= count) return;
    const int t = threadIdx.x;
    const int slot = slot0 + t;
    const bool valid = slot < count;
    const int cand = valid ? (g.idx ? g.idx[slot] : slot) : 0;
    const int n_pad = g.n_pad;
    const double power = g.power_x;

    double xs[NX];
#pragma unroll
    for (int d = 0; d < NX; ++d) {
        double x = valid ? X[(size_t)cand * ldx + g.x_active[d]] : 0.0;
        xs[d] = ((x - g.xmin[d]) / g.xdiff[d]) / g.ls[d];   // trans_x, then GPy's X / lengthscale
    }
    double acc[NCB];
#pragma unroll
    for (int c = 0; c < NCB; ++c) acc[c] = 0.0;

    const double* __restrict__ gX = g.Xs;
    const double* __restrict__ gC = (MODE == 0) ? g.C : g.H;
    const double* __restrict__ gG = g.G;
    double* __restrict__ pan = nullptr;
    const double* __restrict__ bpan = nullptr;
    if (MODE == 0) pan = panels + g.panel_off * nrhs + (size_t)it.tile * nrhs * n_pad * kBlk;
    else bpan = beta_panels + g.panel_off + (size_t)it.tile * n_pad * kBlk;   // beta panels hold one RHS kind

    for (int i0 = 0; i0 < n_pad; i0 += kEvalRows) {
        __syncthreads();
        for (int k = t; k < kEvalRows * NX; k += kBlk) sX[k] = gX[(size_t)i0 * NX + k];
        for (int k = t; k < ncols * kEvalRows; k += kBlk) {
            int c = k / kEvalRows, i = k % kEvalRows;
            sC[k] = gC[(size_t)c * n_pad + i0 + i];
        }
        if (MODE == 0) {
            for (int k = t; k < nrhs * kEvalRows; k += kBlk) {
                int a = k / kEvalRows, i = k % kEvalRows;
                sG[k] = gG[(size_t)a * n_pad + i0 + i];
            }
        }
        __syncthreads();
#pragma unroll 2
        for (int i = 0; i < kEvalRows; ++i) {
            double r2 = 0.0;
#pragma unroll
            for (int d = 0; d < NX; ++d) {
                double df = xs[d] - sX[i * NX + d];
                r2 = fma(df, df, r2);
            }
            double kap = kappa_of_r2<KERN>(r2, power);
            const int row = i0 + i;
            if (MODE == 1) kap *= bpan[(size_t)row * kBlk + (t ^ ((row & 3) << 2))];
#pragma unroll
            for (int c = 0; c < NCB; ++c)
                if (c < ncols) acc[c] = fma(kap, sC[c * kEvalRows + i], acc[c]);
            if (MODE == 0) {
                for (int a = 0; a < nrhs; ++a)
                    pan[((size_t)a * n_pad + row) * kBlk + (t ^ ((row & 3) << 2))] = kap * sG[a * kEvalRows + i];
            }
        }
    }
    if (!valid) return;
    const RawOut& raw = raws[g.model];
    const size_t orow = (size_t)g.out_e * chunk_n + cand;
    if (MODE == 0) {
#pragma unroll
        for (int c = 0; c < NCB; ++c)
            if (c < ncols) raw.cmu[orow * ldraw + c] = acc[c];
        raw.kss[orow] = g.kss;
    } else {
        const int np = g.P * (g.P + 1) / 2;
#pragma unroll
        for (int c = 0; c < NCB; ++c)
            if (c < ncols) raw.bh[orow * np + c] = acc[c];
    }
}

template <int KERN, int NX, int MODE>
static int dispatch_ncb(cudaStream_t s, int nitems, int ncols, const GpView* gps, const Item* items,
                        const double* X, int ldx, int chunk_n, int nrhs, double* panels, const double* beta,
                        const RawOut* raws, int ldraw) {
    if (ncols <= 8)
        eval_kernel<KERN, NX, 8, MODE><<<nitems, kBlk, 0, s>>>(gps, items, X, ldx, chunk_n, nrhs, ncols, panels,
                                                                beta, raws, ldraw);
    else if (ncols <= 24)
        eval_kernel<KERN, NX, 24, MODE><<<nitems, kBlk, 0, s>>>(gps, items, X, ldx, chunk_n, nrhs, ncols, panels,
                                                                 beta, raws, ldraw);
    else if (ncols <= 72)
        eval_kernel<KERN, NX, 72, MODE><<<nitems, kBlk, 0, s>>>(gps, items, X, ldx, chunk_n, nrhs, ncols, panels,
                                                                 beta, raws, ldraw);
    else {
        set_error("evaluation kernel supports at most 72 coefficient columns (P <= 10 for second order)");
        return 3;
    }
    GPD_CUDA(cudaGetLastError());
    return 0;
}

template <int KERN, int MODE>
static int dispatch_nx(cudaStream_t s, int nx, int nitems, int ncols, const GpView* gps, const Item* items,
                       const double* X, int ldx, int chunk_n, int nrhs, double* panels, const double* beta,
                       const RawOut* raws, int ldraw) {
#define GPD_NX_CASE(NXV)                                                                                   \
    case NXV:                                                                                              \
        return dispatch_ncb<KERN, NXV, MODE>(s, nitems, ncols, gps, items, X, ldx, chunk_n, nrhs, panels,  \
                                             beta, raws, ldraw);
    switch (nx) {
        GPD_NX_CASE(1)
        GPD_NX_CASE(2)
        GPD_NX_CASE(3)
        GPD_NX_CASE(4)
        GPD_NX_CASE(5)
        GPD_NX_CASE(6)
        GPD_NX_CASE(7)
        GPD_NX_CASE(8)
        default:
            set_error("x-kernel acts on 1..8 non-binary design dimensions");
            return 3;
    }
#undef GPD_NX_CASE
}

template <int MODE>
static int dispatch_kernel(cudaStream_t s, int kernel, int nx, int nitems, int ncols, const GpView* gps,
                           const Item* items, const double* X, int ldx, int chunk_n, int nrhs, double* panels,
                           const double* beta, const RawOut* raws, int ldraw) {
#define GPD_K_CASE(KV)                                                                                     \
    case KV:                                                                                               \
        return dispatch_nx<KV, MODE>(s, nx, nitems, ncols, gps, items, X, ldx, chunk_n, nrhs, panels, beta, \
                                     raws, ldraw);
    switch (kernel) {
        GPD_K_CASE(0)
        GPD_K_CASE(1)
        GPD_K_CASE(2)
        GPD_K_CASE(3)
        GPD_K_CASE(4)
        GPD_K_CASE(5)
        default:
            set_error("unknown kernel id");
            return 3;
    }
#undef GPD_K_CASE
}

int launch_eval(cudaStream_t s, const GpView* gps_dev, const Item* items_dev, int nitems, const double* X_dev,
                int ldx, int chunk_n, int nrhs, int ncols, int kernel, int nx, double* scratch,
                const RawOut* raw_dev, int ldraw) {
    if (nitems == 0) return 0;
    return dispatch_kernel<0>(s, kernel, nx, nitems, ncols, gps_dev, items_dev, X_dev, ldx, chunk_n, nrhs, scratch,
                              nullptr, raw_dev, ldraw);
}

int launch_beta_contract(cudaStream_t s, const GpView* gps_dev, const Item* items_dev, int nitems,
                         const double* X_dev, int ldx, int chunk_n, int kernel, int nx, int P,
                         const double* beta_panels, const RawOut* raw_dev) {
    if (nitems == 0) return 0;
    return dispatch_kernel<1>(s, kernel, nx, nitems, P * (P + 1) / 2, gps_dev, items_dev, X_dev, ldx, chunk_n,
                              1, nullptr, beta_panels, raw_dev, 0);
}

}  // namespace gpd
