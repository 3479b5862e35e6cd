// Internal declarations shared by the translation units of libgpdoemd_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "layout.h"

namespace gpd {

constexpr int kMaxXDims = 8;     // active (non-binary) design dimensions handled by the evaluation kernel
constexpr int kMaxP = 16;        // model parameters
constexpr int kMaxE = 8;         // outputs (register-resident E x E criteria)
constexpr int kMaxBinary = 6;
constexpr int kMaxKinds = 5;      // criteria scored in one fused pass
// RBF GPs keep their x-coordinates (training rows Xs and the candidate scale xscale) multiplied by sqrt(128 / ln 2): the
// squared distance is then the argument of the exponential table in table units (eval_impl.cuh: exp_neg_tab_units)
constexpr double kRbfCoordScale = 13.589148804608305;

// Device-side view of one GP (array in global memory, indexed by work items).
struct GpView {
    int kernel, n, n_pad, nblk, nx, P;
    double kss;                  // var_x * var_p = k(z*, z*)
    double power_x;
    const double* Xs;            // [n_pad][nx]  training x / lengthscale   (pad rows 0)
    const double* XsE;           // RBF only: [n_pad][nx + 1] = {-2 Xs, |Xs|^2}, the expanded-form record of eval_kernel
    const double* C;             // [1+P+P(P+1)/2][n_pad] alpha*var_x*{kp, dkp/dp_a, d2kp/dp_a dp_b (a<=b)}
    const double* G;             // [1+P][n_pad]          var_x*{kp, dkp/dp_a}          (pad rows 0)
    const double* H;             // [P(P+1)/2][n_pad]     var_x*d2kp/dp_a dp_b
    const double* LinvF;         // packed forward factor  (layout.h)
    const double* LinvB;         // packed backward factor
    double ls[kMaxXDims];        // x lengthscales of the active dims
    double xmin[kMaxXDims], xdiff[kMaxXDims];
    double xscale[kMaxXDims];    // 1 / (xdiff * lengthscale) [* sqrt(1/2) for RBF]: candidate -> kernel coordinate
    int x_active[kMaxXDims];
    // routing of this GP inside the current chunk
    const int* idx;              // candidate positions (within chunk) handled by this GP, or null = identity
    const int* count;            // device count of those positions, or null = chunk size
    int out_e;                   // output row e of the owning model
    int model;                   // model slot inside the current call
    long long panel_off;         // offset of this GP's region in a ONE-kind panel buffer (first order, beta), in doubles
    long long panel_off_w;       // offset in a buffer holding this GP's own 1 + P kinds per tile (variance derivatives):
                                 // rival models may differ in P, so this is NOT panel_off * kinds
};

// Work item of the evaluation / triangular kernels: GP + candidate tile (128 columns).
// I0 / I1: row blocks [I0, I1) of the triangular contraction this unit covers; I1 == 0 means all of them (the value an
// `Item{gp, tile}` initialiser leaves).  The units of one (GP, tile) that is cut into parts publish their ||v||^2
// partials in distinct slots (part * row groups + row group) and, scheduled on neighbouring CTAs, share the tile's
// right-hand-side panel through L2 instead of re-reading it from HBM once per row block.
struct Item { int gp; int tile; short I0, I1; short part, nparts; };

// Fused evaluation + contraction kernel (fused_impl.cuh): candidates per work item and CTAs per SM.  128 / 1 = one
// 512-thread CTA per SM (3 x 32 k pipeline stages); 64 / 2 = two 256-thread CTAs per SM on half tiles (4 x 16 k), which
// was measured slower at every size (profiles/r02_summary.md) and is kept as a compile-time variant only.
#ifndef GPD_FUSED_TILE
#define GPD_FUSED_TILE 128
#endif
constexpr int kFusedTile = GPD_FUSED_TILE;
constexpr int kFusedCtasPerSm = kBlk / kFusedTile;

// Work item of the fused evaluation + contraction kernel: row blocks [I0, I1) of (GP, tile); the parts of one (GP, tile)
// write their ||v||^2 partials to distinct slots
typedef Item FItem;

// Per-model raw outputs for a chunk (device pointers), transformed units.
struct RawOut {
    double* cmu;    // [E][Nc][ldraw]  kx . C columns
    double* kss;    // [E][Nc]         k** of the GP that handled the row (0 = unrouted)
    double* gram;   // [E][Nc][ng]     ng = 1 (v'v) or (1+P)(2+P)/2 upper triangle of [v W]'[v W]
    double* bh;     // [E][Nc][P(P+1)/2]   beta . (kx * d2kp)
};

struct ModelXform {              // device copy of the per-model unit conversions
    int E, P;
    double ymean[kMaxE], ystd[kMaxE];
    double pdiff[kMaxP];
    double Sigma[kMaxP * kMaxP]; // row-major, stride P
    double var_floor;            // lower clip of the predictive variance in transformed units: 1e-15 for sparse GPs
                                 // (GPy's base Posterior._raw_predict clips), -inf for exact GPs (PosteriorExact does not)
};

struct PostModel {               // per model slot of a fused scoring call (device table)
    const ModelXform* xf;
    int P, ldraw, npart;
};

struct KmatArgs {                // covariance-matrix build of one GP (kernels_build.cu)
    int kernel, n, dz, D, P, nx;
    int x_active[kMaxXDims];
    double ls_x[kMaxXDims], ls_p[kMaxP];
    double var_x, var_p, power_x, power_p;
    double diag_add;             // Gaussian noise variance + GPy's 1e-8 (+ jitchol jitter)
};

struct FusedArgs {                // fused_impl.cuh
    const GpView* gps;
    const FItem* items;
    int nitems;
    const double* X;
    int ldx, chunk_n;
    double* priv;              // [gridDim.x][priv_stride] private right-hand-side panels
    long long priv_stride;     // doubles per CTA (>= n_pad_max * 128)
    const RawOut* raws;
    int ldraw, ncols, ng;      // ng = row length of raw.gram (= 2 * max parts)
    int trim;
};

struct XgradArgs {               // finalize step of the design-gradient path (kernels_xgrad.cu)
    int E, R, D, P, nx, nb, ngram;
    int x_active[kMaxXDims];
    int bcols[kMaxBinary], weights[kMaxBinary];
    const ModelXform* xf;
    const GpView* gps;            // E * R views, index e * R + r; missing GPs have n == 0
    const double* X;
    int ldx, chunk_n;
    const double* acc;            // [gp][chunk_n][(1+nx)*(1+P)]
    const double* gram;           // [gp][chunk_n][ngram]  upper triangle of [v W]^T [v W]
    double *mu, *S, *dmu, *dS;    // (N,E) (N,E,E) (N,E,D) (N,E,E,D), original units
};

void set_error(const std::string& msg);
#define GPD_CUDA(call)                                                                        \
    do {                                                                                      \
        cudaError_t err__ = (call);                                                           \
        if (err__ != cudaSuccess) {                                                           \
            ::gpd::set_error(std::string(#call) + ": " + cudaGetErrorString(err__) + " (" +   \
                             __FILE__ + ":" + std::to_string(__LINE__) + ")");                \
            return 1;                                                                         \
        }                                                                                     \
    } while (0)
#define GPD_CHECK(cond, msg)                                                                  \
    do {                                                                                      \
        if (!(cond)) {                                                                        \
            ::gpd::set_error(std::string(msg) + " [" #cond "] (" + __FILE__ + ":" +          \
                             std::to_string(__LINE__) + ")");                                 \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)
#define GPD_TRY(expr)                                                                         \
    do {                                                                                      \
        int rc__ = (expr);                                                                    \
        if (rc__) return rc__;                                                                \
    } while (0)

// ---- kernel launchers (0 = ok) ----------------------------------------------------------------
// kernels_eval.cu
int launch_pstate(cudaStream_t s, int kernel, int n, int n_pad, int P, int dim_z, const double* Zp,
                  const double* alpha, const double* pstar, const double* ls_p, double var_x, double var_p,
                  double power_p, double* C, double* G, double* H);
int launch_eval(cudaStream_t s, const GpView* gps_dev, const Item* items_dev, int nitems, const double* X_dev,
                int ldx, int chunk_n, int nrhs, int ncols, int kernel, int nx, double* panels,
                const RawOut* raw_dev, int ldraw);
int launch_beta_contract(cudaStream_t s, const GpView* gps_dev, const Item* items_dev, int nitems,
                         const double* X_dev, int ldx, int chunk_n, int kernel, int nx, int P,
                         const double* beta_panels, const RawOut* raw_dev);
// kernels_tri.cu
int launch_invert_and_pack(cudaStream_t s, const double* L_dev, int n, int n_pad, double* dense_tmp,
                           double* packF, double* packB, uint64_t* launches, int factor_kind, int setup_variant);
// kernels_inv.cu
int launch_invert_offdiag_dmma(cudaStream_t s, const double* L_dev, int n, int n_pad, double* X, uint64_t* launches);
int launch_tri(cudaStream_t s, const GpView* gps_dev, const Item* items_dev, int nitems, int nwork_rhs,
               int in_stride, int out_stride, int backward, const double* panels_in, double* panels_out,
               const RawOut* raw_dev, int ng, int chunk_n, int num_sms, int variant, int cps, int trim);
int launch_gram(cudaStream_t s, const GpView* gps_dev, const Item* items_dev, int nitems, int nrhs,
                const double* panels, const RawOut* raw_dev, int chunk_n);
int launch_dmma_probe(cudaStream_t s, int num_sms, int iters, double* sink_dev);
// kernels_fused.cu: K1 + K2 in one persistent kernel (first-order path); returns -1 when no variant covers the shape
int launch_fused(cudaStream_t s, int kernel, int grid, int nx, int ncols, const FusedArgs& args);
int launch_fusedc(cudaStream_t s, int kernel, int grid, int nx, int ncols, const FusedArgs& args);   // two private panels per CTA
// kernels_xgrad.cu
int launch_xgrad_eval(cudaStream_t s, const GpView* gps_dev, int ngp, int tiles, const double* X_dev, int ldx, int chunk_n,
                      double* panels, double* acc);
int launch_xgrad_finalize(cudaStream_t s, const XgradArgs& a);
// kernels_build.cu
int launch_kmat(cudaStream_t s, const KmatArgs& a, const double* Z_dev, double* A_dev);
int launch_cholesky(cudaStream_t s, double* A_dev, int n, int* info_dev, uint64_t* launches, int setup_variant);
int launch_chol_solve(cudaStream_t s, const double* L_dev, int n, const double* y_dev, double* x_dev, uint64_t* launches,
                      int setup_variant);
// kernels_crit.cu
int launch_route(cudaStream_t s, const double* X_dev, int ldx, int chunk_n, int nb, const int* bcols,
                 const int* weights, int ncombo, int* idx, int* counts, int idx_stride);
int launch_finalize(cudaStream_t s, const ModelXform* xf_dev, const RawOut* raw_dev, int model_slot, int E, int P,
                    int chunk_n, int ldraw, int flags, int nrhs, int npart, double* mu, int ldmu, double* s2,
                    double* dmu, double* d2mu, double* ds2, double* d2s2);
int tri_row_groups(int variant);   // WM of the selected tri_kernel variant = number of ||v||^2 partials
int launch_assemble(cudaStream_t s, int order, int N, int E, int P, double* mu, int ldmu, const double* s2,
                    const double* dmu, const double* d2mu, const double* d2s2, const double* Sigma_dev, double* S,
                    int ldS, double* tmpA, double* tmp_s2);
int launch_criterion(cudaStream_t s, int kind, const double* mu, const double* S, int N, int M, int E,
                     const double* noise_dev, const double* pps_dev, double* dc, double* work);
size_t criterion_work_doubles(int N, int M, int E);
int launch_gauss_loglik(cudaStream_t s, const double* Y, const double* Mu, const double* S, long long n, int N, int E,
                        double* logpdf, double* maha);
int launch_argmax(cudaStream_t s, const double* dc, long long N, long long idx_base, double* best_val_dev,
                  long long* best_idx_dev, int first, void* partial_dev);
size_t argmax_partial_bytes();
size_t best_pair_bytes();
int launch_singular_fetch(cudaStream_t s, int* host_out);
int launch_best_pack(cudaStream_t s, const double* bv, const long long* bi, int nk, int empty, void* out);
int launch_best_reduce(cudaStream_t s, const void* all, int world, int nk, double* bv, long long* bi);
size_t crit_partial_bytes(int N, int nk);
int launch_post1(cudaStream_t s, const PostModel* pms_dev, const RawOut* raws_dev, int M, int E, int N, int chunk_n,
                 const double* noise_dev, double* work, const int* kinds, int nk);
int launch_crit_prep(cudaStream_t s, const double* mu, const double* S, int N, int M, int E, const double* noise_dev,
                     double* work, const int* kinds, int nk);
int launch_crit_multi(cudaStream_t s, const int* kinds, int nk, double* const* dc, int N, int M, int E,
                      const double* noise_dev, const double* pps_dev, double* work, long long idx_base, void* partial_dev,
                      int first, double* best_val_dev, long long* best_idx_dev);
int launch_fill_uniform(cudaStream_t s, double* X, long long N, int D, uint64_t seed, long long first_row,
                        const double* lo_dev, const double* hi_dev);
int launch_flush(cudaStream_t s, double* buf, size_t n);

}  // namespace gpd
