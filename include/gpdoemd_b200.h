/*
 * gpdoemd_b200 -- C ABI of the B200-native GPdoemd design-evaluation hot path.
 *
 * The reference (cog-imperial/GPdoemd) is pure Python and has no FFI; this header is the
 * boundary a maintainer would bind with ctypes.  Every entry point names
 * the reference interface it replaces (file:line relative to the reference repository).
 *
 * Conventions
 *   - all pointers are HOST pointers to row-major (C-order) float64 unless the name ends in
 *     `_dev` (device pointer owned by the caller) ; the caller owns every buffer it passes;
 *   - the library owns its device copies until the matching `*_free`;
 *   - every function returns 0 on success, non-zero on failure; the message is available
 *     through gpd_last_error(); nothing is computed on the CPU: if no CUDA device / kernel image
 *     is usable the call fails, it never falls back;
 *   - one context per device, not re-entrant per context.
 *
 * Units: `gpd_gp_desc` holds the GP in the reference's *transformed* units (Z in [0,1],
 * Y standardised: surrogate_model.py:70-156); the model-level calls take and return
 * ORIGINAL units exactly like the public SurrogateModel wrappers (surrogate_model.py:170-215).
 */
#ifndef GPDOEMD_B200_H
#define GPDOEMD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gpd_ctx gpd_ctx;
typedef struct gpd_gp gpd_gp;
typedef struct gpd_model gpd_model;
typedef struct gpd_comm gpd_comm;
typedef struct gpd_resident gpd_resident;   /* marginals of all rival models kept in device memory for the criteria */   /* one NCCL communicator bound to a context (multi-GPU scoring) */

/* GPdoemd/kernels/stationary.py:33-99 -- the six ARD stationary kernels */
enum gpd_kernel {
    GPD_KERN_RBF = 0,         /* stationary.py:33 */
    GPD_KERN_EXPONENTIAL = 1, /* stationary.py:41 */
    GPD_KERN_MATERN32 = 2,    /* stationary.py:52 */
    GPD_KERN_MATERN52 = 3,    /* stationary.py:64 */
    GPD_KERN_COSINE = 4,      /* stationary.py:76 */
    GPD_KERN_RATQUAD = 5      /* stationary.py:87 */
};

/* GPdoemd/design_criteria.py:49,66,95,126,147 */
enum gpd_criterion_kind { GPD_HR = 0, GPD_BH = 1, GPD_BF = 2, GPD_AW = 3, GPD_JR = 4 };

/*
 * One exact GP = one GPy `GPRegression(Zr, Yr, kernx*kernp)` of gp_model.py:123, handed over
 * after training.  The fields are the attributes the reference itself reads at
 * gp_model.py:231-233,262-275: gp._predictive_variable, gp.posterior.woodbury_vector,
 * gp.posterior.woodbury_chol, gp.kern.kernx/kernp.{variance,lengthscale,active_dims}.
 */
typedef struct gpd_gp_desc {
    int32_t kernel;        /* enum gpd_kernel; both product parts use this class (gp_model.py:60 quirk) */
    int32_t n;             /* training rows */
    int32_t dim_x;         /* D: design columns of Z (binary ones included) */
    int32_t dim_p;         /* P: model-parameter columns of Z (columns D..D+P-1) */
    int32_t n_x_active;    /* number of non-binary design columns the x-kernel acts on */
    const int32_t* x_active; /* their column indices (gp_model.py:117 non_binary_variables) */
    double var_x;          /* kernx.variance */
    const double* ls_x;    /* kernx.lengthscale, n_x_active entries */
    double power_x;        /* kernx.power (RatQuad only) */
    double var_p;          /* kernp.variance */
    const double* ls_p;    /* kernp.lengthscale, P entries */
    double power_p;        /* kernp.power (RatQuad only) */
    const double* Z;       /* (n, D+P) gp._predictive_variable, transformed units */
    const double* alpha;   /* (n)      gp.posterior.woodbury_vector */
    const double* L;       /* (n, n)   lower triangle used; meaning given by factor_kind */
    int32_t factor_kind;   /* GPD_FACTOR_CHOL: L = gp.posterior.woodbury_chol = chol(K + noise), the library forms
                            *   L^-1 (exact GP, gp_model.py:123);
                            * GPD_FACTOR_ROOT: L = T, lower triangular with gp.posterior.woodbury_inv = T^T T, used
                            *   as is: variance = k** - |T k*|^2 (sparse GP, sparse_gp_model.py:35-74: n = number
                            *   of inducing inputs, Z = gp._predictive_variable = the inducing inputs) */
} gpd_gp_desc;
enum gpd_factor_kind { GPD_FACTOR_CHOL = 0, GPD_FACTOR_ROOT = 1 };

/* return codes: 0 ok; 1 CUDA error; 2 invalid argument; 3 unsupported size; 4 covariance not positive definite
 * (gpd_gp_build); 5 NCCL error; GPD_ERR_SINGULAR: an E x E covariance handed to a criterion was exactly singular
 * (np.linalg.inv in design_criteria.py:83,115,137,173 raises numpy.linalg.LinAlgError there) */
enum { GPD_ERR_SINGULAR = 6 };

/* ---- context ----------------------------------------------------------------------------- */
int  gpd_create(int device, gpd_ctx** out);
void gpd_destroy(gpd_ctx* ctx);
const char* gpd_last_error(void);
/* bytes of per-context scratch the fused paths may use for the K(x*,X) / L^-1 K panels (default 6 GiB) */
int  gpd_set_scratch_bytes(gpd_ctx* ctx, uint64_t bytes);
/* tuning knobs: "tri_variant" 0 = 16 warps as 4x4 (32x32 warp tiles), 1 = 8 warps as 2x4 (64x32), 2 (default) = 16 warps
 * as 2x8 (64x16: two row groups halve the idle time of the low row groups in the diagonal tiles);
 * "tri_cps" 0 = automatic, 1 = 4 pipeline stages of 16 k, 2 = 3 stages of 32 k;
 * "tail_overlap" 1 (default) = first-order path: the candidate tiles that would form the last, partial wave of the
 *   persistent triangular kernel are evaluated and contracted first, concurrently with the evaluation of the rest;
 * "tri_trim" 1 (default) = the all-padding 8-row blocks (rows n..n_pad) of the last row block are skipped;
 * "fused" 0 (default) / 1 / 2 = first-order path through ONE persistent kernel that evaluates the kernel rows into a private
 *   L2-resident panel and contracts them in place: O(SMs * n) scratch instead of O(N * n), measured 1-8 % slower than the
 *   separate kernels; 2 = the same with a producer warpgroup evaluating the next item concurrently
 *   (measured slower still);
 * "scratch_redzones" 0 (default) / 1 = debug red zones behind every scratch sub-buffer, see gpd_debug_check_scratch;
 * "head_periods" 0 (default) .. 16 / "head_grid" 0 (all SMs) or a CTA count: a longer tail-overlap head on a restricted
 *   grid (measured neutral on C2);
 * "tri_parts" 0 (default) = whole work items, 1..8 = every (GP, tile) item of the first-order contraction cut into that
 *   many row-block parts, 9 = by factor size (measured slower at n = 4000, kept for experiments);
 * "setup_variant" 1 (default) = per-GP setup (gpd_gp_upload / gpd_gp_build) with the factor inverse and the Cholesky
 *   trailing update on the FP64 tensor pipe, 0 = the round-1 FP64 FMA kernels (A/B tests) */
int  gpd_set_option(gpd_ctx* ctx, const char* name, int64_t value);
/* number of CUDA kernels launched by this context so far (bench.py `gpu_launches`) */
uint64_t gpd_launch_count(gpd_ctx* ctx);
/* device-time accounting per kernel family with CUDA events on the library's stream:
 * enable, run, then read the accumulated milliseconds / launch counts.  family index:
 * 0 eval (kernel+derivative rows), 1 tri (DMMA triangular contraction), 2 gram, 3 marginal,
 * 4 criteria, 5 argmax, 6 setup (p-state, routing, packing), 7 tri_head (the small triangular launch over the
 * head tiles that runs concurrently with the evaluation of the remaining tiles, option "tail_overlap"),
 * 8 comm (pack + ncclAllGather + fold of the per-rank (max, index) pairs in gpd_score_sharded) */
enum { GPD_NFAMILIES = 9 };
int  gpd_profile_enable(gpd_ctx* ctx, int on);
int  gpd_profile_read(gpd_ctx* ctx, double ms_out[GPD_NFAMILIES], uint64_t launches_out[GPD_NFAMILIES],
                      double flops_out[GPD_NFAMILIES]);
/* CUDA-event stopwatch on the library's stream (bench.py timed region) */
int  gpd_timer_start(gpd_ctx* ctx);
int  gpd_timer_stop(gpd_ctx* ctx, double* ms_out);
int  gpd_sync(gpd_ctx* ctx);
/* FP64 DMMA (mma.sync m8n8k4) register-resident peak probe: roofline denominator measured live */
int  gpd_probe_fp64_peak(gpd_ctx* ctx, double* tflops_out);
/* write a buffer larger than L2 (bench.py L2 flush between timed iterations) */
int  gpd_flush_l2(gpd_ctx* ctx);

/* ---- GP state hand-off (replaces nothing: training stays on GPy, gp_model.py:96-181) ------- */
int  gpd_gp_upload(gpd_ctx* ctx, const gpd_gp_desc* desc, gpd_gp** out);
/* the exact-GP inference itself on the device (gp_model.py:96-147: GPRegression(Zr, Yr, kernx*kernp) in
 * gp_surrogate, re-run by gp_load_hyp's `gp[:] = hyp`): Ky = K(Z,Z) + (noise_var + 1e-8) I, L = jitchol(Ky)
 * (blocked FP64 Cholesky; on a non-positive pivot up to 5 retries with jitter mean(diag) * 1e-6 * 10^k like GPy),
 * alpha = Ky^-1 Y.  desc->alpha and desc->L are ignored; factor_kind must be GPD_FACTOR_CHOL.  Y has n entries
 * (standardised targets of this output).  alpha_out (n), L_out (n*n, row-major; only the lower triangle is
 * meaningful) and jitter_out may be NULL: they expose gp.posterior.woodbury_vector / woodbury_chol to the host. */
int  gpd_gp_build(gpd_ctx* ctx, const gpd_gp_desc* desc, const double* Y, double noise_var, gpd_gp** out,
                  double* alpha_out, double* L_out, double* jitter_out);
void gpd_gp_free(gpd_gp* gp);

/* ---- model = GPModel after gp_surrogate/gp_load_hyp (gp_model.py:96-143) ------------------- */
/* gps: E * R handles, index e*R + r, r = binary-combination index of utils.py:27-48; NULL where the
 * reference has gps[e][r] = None (gp_model.py:111-113).  binary: n_binary column indices.
 * xmin..ystd are the transforms of surrogate_model.py:77-85,139-143. */
int  gpd_model_create(gpd_ctx* ctx, int32_t E, int32_t dim_x, int32_t dim_p,
                      int32_t n_binary, const int32_t* binary,
                      gpd_gp* const* gps,
                      const double* xmin, const double* xmax,
                      const double* pmin, const double* pmax,
                      const double* ymean, const double* ystd,
                      gpd_model** out);
void gpd_model_free(gpd_model* model);
/* model.pmean / model.Sigma (model.py:100-136), ORIGINAL parameter units; Sigma may be NULL */
int  gpd_model_set_pmean_sigma(gpd_model* model, const double* pmean, const double* Sigma);
/* the transformed parameter point p* = trans_p(pmean) directly (second argument of GPModel._predict,
 * gp_model.py:184); pmean is updated to the matching original-unit value */
int  gpd_model_set_pstar(gpd_model* model, const double* pstar);

/* ---- the hooks, all E outputs at once, ORIGINAL units ------------------------------------- */
/* SurrogateModel.predict -> GPModel._predict   (surrogate_model.py:170-178, gp_model.py:184-200) */
int  gpd_predict(gpd_model* model, const double* X, int64_t N, double* M_out /*N,E*/, double* S_out /*N,E*/);
/* d_mu_d_p   (surrogate_model.py:185-189, gp_model.py:206-218)   out (N,E,P)   */
int  gpd_dmu_dp(gpd_model* model, const double* X, int64_t N, double* out);
/* d2_mu_d_p2 (surrogate_model.py:193-198, gp_model.py:220-235)   out (N,E,P,P) */
int  gpd_d2mu_dp2(gpd_model* model, const double* X, int64_t N, double* out);
/* d_s2_d_p   (surrogate_model.py:202-206, gp_model.py:237-249)   out (N,E,P)   */
int  gpd_ds2_dp(gpd_model* model, const double* X, int64_t N, double* out);
/* d2_s2_d_p2 (surrogate_model.py:210-215, gp_model.py:251-280)   out (N,E,P,P) */
int  gpd_d2s2_dp2(gpd_model* model, const double* X, int64_t N, double* out);

/* every hook in one pass over X (any output may be NULL): what a host wrapper that caches per X
 * calls once instead of E separate per-output calls (taylor_first_order.py:32-33 loops over e).
 * transformed != 0: X and all outputs are in the reference's TRANSFORMED units, i.e. exactly the
 * `_predict/_d_mu_d_p/_d2_mu_d_p2/_d_s2_d_p/_d2_s2_d_p2` hooks of gp_model.py:184-280 */
int  gpd_hooks(gpd_model* model, const double* X, int64_t N, int32_t transformed, double* mu /*N,E*/,
               double* s2 /*N,E*/, double* dmu /*N,E,P*/, double* d2mu /*N,E,P,P*/, double* ds2 /*N,E,P*/,
               double* d2s2 /*N,E,P,P*/);

/* ---- marginals ----------------------------------------------------------------------------- */
/* taylor_first_order / taylor_second_order on a device model
 * (marginal/taylor_first_order.py:27-43, taylor_second_order.py:27-60); order = 1 or 2 */
int  gpd_marginal(gpd_model* model, const double* X, int64_t N, int order,
                  double* mu_out /*N,E*/, double* S_out /*N,E,E*/);
/* the assembly step alone, for duck-typed models whose hooks ran elsewhere
 * (taylor_first_order.py:36-42, taylor_second_order.py:39-59): inputs already in original units;
 * ddmu/dds2 (N,E,P,P) are NULL for order 1; mu_io (N,E) is updated in place for order 2 */
int  gpd_marginal_assemble(gpd_ctx* ctx, int order, int64_t N, int32_t E, int32_t P,
                           double* mu_io, const double* s2, const double* dmu,
                           const double* ddmu, const double* dds2, const double* Sigma,
                           double* S_out);

/* ---- design criteria (design_criteria.py:49-194) ------------------------------------------- */
/* mu (N,M,E), S (N,M,E,E), noise (E,E), pps (M) already normalised as by _reshape (:202-238);
 * inputs are never modified (the reference's BH/BF add noise into the caller's s2, :81,:114) */
int  gpd_criterion(gpd_ctx* ctx, int kind, const double* mu, const double* S, int64_t N,
                   int32_t M, int32_t E, const double* noise, const double* pps, double* dc_out);
/* ---- discrimination criteria on the observed set (discrimination_criteria.py:32-124) -----------
 * gaussian_posterior (:32-52), gaussian_posterior_update (:57-66), chi2 (:72-103) and akaike (:109-130) all reduce
 * to the Gaussian log-density and the squared Mahalanobis distance of every observation j under every model i:
 *   logpdf_out[j,i] = log N(Y[j]; M[j,i], S[j,i])   (scipy.stats.multivariate_normal.logpdf, :40,:118)
 *   maha_out[j,i]   = (M[j,i]-Y[j])^T S[j,i]^-1 (M[j,i]-Y[j])                                   (:93-98)
 * Y (n,E), M (n,N,E), S (n,N,E,E); the sums over j and the weights are a few host flops (n, N are tiny). */
int  gpd_gauss_loglik(gpd_ctx* ctx, const double* Y, const double* M, const double* S, int64_t n, int32_t N,
                      int32_t E, double* logpdf_out /*n,N*/, double* maha_out /*n,N*/);
/* np.argmax with numpy semantics: first index of the maximum, NaN wins (demos/case_study.py:291) */
int  gpd_argmax(gpd_ctx* ctx, const double* dc, int64_t N, double* best_val, int64_t* best_idx);

/* ---- fused scoring: predict + marginal + criterion + argmax, nothing N-sized leaves the GPU --
 * the loop of demos/case_study.py:284-293 for one criterion.  dc_out may be NULL.
 * idx_offset is added to best_idx (global index of this shard's first candidate). */
int  gpd_score(gpd_ctx* ctx, gpd_model* const* models, int32_t M, const double* X, int64_t N,
               int order, int kind, const double* noise /*E,E*/, const double* pps /*M*/,
               double* dc_out, double* best_val, int64_t* best_idx, int64_t idx_offset);
/* same with the candidates (and optionally dc) already resident in device memory */
int  gpd_score_dev(gpd_ctx* ctx, gpd_model* const* models, int32_t M, const double* X_dev, int64_t N,
                   int order, int kind, const double* noise, const double* pps,
                   double* dc_out_dev, double* best_val, int64_t* best_idx, int64_t idx_offset);
/* several criteria in ONE pass over the candidates (the demo scores all five on the same mu, S:
 * demos/case_study.py:289-293).  X / dc_out are host or device pointers according to the two flags;
 * dc_out (may be NULL) is [nkinds][N]; best_val / best_idx have nkinds entries. */
int  gpd_score_multi(gpd_ctx* ctx, gpd_model* const* models, int32_t M, const double* X, int32_t x_on_device,
                     int64_t N, int order, const int32_t* kinds, int32_t nkinds, const double* noise,
                     const double* pps, double* dc_out, int32_t dc_on_device, double* best_val,
                     int64_t* best_idx, int64_t idx_offset);
/* ---- design gradients (refining the grid argmax of demos/case_study.py:279-293) --------------------
 * taylor_first_order (marginal/taylor_first_order.py:27-43) together with its derivative with respect to the design:
 * mu (N,E), S (N,E,E), dmu_dx (N,E,D), dS_dx (N,E,E,D), ORIGINAL units; columns of binary design variables hold 0.
 * The reference has no such gradient (it only evaluates a grid); parity = central finite differences of the marginal. */
int  gpd_marginal_dx(gpd_model* model, const double* X, int64_t N, double* mu_out, double* S_out,
                     double* dmu_dx_out, double* dS_dx_out);

/* ---- the drop-in call pattern without the host round trip (demos/case_study.py:284-293) -------------------------
 * gpd_marginals = taylor_{first,second}_order for ALL rival models in one pass: mu_out (N,M,E) and S_out (N,M,E,E) are
 * the arrays the reference's loop `mu[:,i], s2[:,i] = taylor(M_i, X)` fills (either may be NULL).  With keep != NULL the
 * device copies stay alive in *keep, and gpd_criterion_resident scores them with nothing but dc (N) crossing PCIe.
 * The host wrapper returns mu / S as read-only arrays so that the resident copy cannot go stale. */
int  gpd_marginals(gpd_ctx* ctx, gpd_model* const* models, int32_t M, const double* X, int64_t N, int order,
                   double* mu_out, double* S_out, gpd_resident** keep);
int  gpd_criterion_resident(gpd_ctx* ctx, int kind, const gpd_resident* marginals, const double* noise /*E,E*/,
                            const double* pps /*M*/, double* dc_out /*N*/);
void gpd_resident_free(gpd_resident* marginals);

/* ---- multi-GPU (BASELINE configs[4]): candidates sharded in contiguous blocks, GP state replicated,
 * ONE collective per scoring call = the all-gather of the per-rank (best value, best global index) pairs.  The
 * reference is single-process (demos/case_study.py:279-293 takes np.argmax over the whole grid); these entry points
 * are what a maintainer binds to spread that grid over the GPUs of a box -- one context (and one host thread or one
 * process) per GPU.  NCCL is resolved at run time (dlopen of libnccl.so.2; an instance already loaded by the process,
 * e.g. PyTorch's, is reused); without it these calls fail, there is no fallback. */
enum { GPD_NCCL_ID_BYTES = 128 };          /* sizeof(ncclUniqueId) */
int  gpd_nccl_version(int32_t* version_out);
/* ncclGetUniqueId: call on ONE rank, hand the 128 bytes to the others (any channel: torch.distributed, MPI, a file) */
int  gpd_nccl_unique_id(char id_out[GPD_NCCL_ID_BYTES]);
/* ncclCommInitRank on the context's device; collective over the `world` ranks (threads or processes) */
int  gpd_comm_init(gpd_ctx* ctx, const char id[GPD_NCCL_ID_BYTES], int32_t rank, int32_t world, gpd_comm** out);
/* wrap an ncclComm_t the caller already owns (not destroyed by gpd_comm_destroy) */
int  gpd_comm_adopt(gpd_ctx* ctx, void* nccl_comm /* ncclComm_t */, int32_t rank, int32_t world, gpd_comm** out);
void gpd_comm_destroy(gpd_comm* comm);
/* gpd_score_multi on this rank's block [idx_offset, idx_offset + N_local) (N_local may be 0), then the all-gather and
 * the np.argmax fold (NaN wins, ties -> lowest global index) on the library stream; best_val / best_idx (nkinds
 * entries) are the GLOBAL winners, identical on every rank.  dc_out (may be NULL) receives this rank's block only. */
int  gpd_score_sharded(gpd_ctx* ctx, gpd_comm* comm, gpd_model* const* models, int32_t M, const double* X,
                       int32_t x_on_device, int64_t N_local, int order, const int32_t* kinds, int32_t nkinds,
                       const double* noise, const double* pps, double* dc_out, int32_t dc_on_device,
                       double* best_val, int64_t* best_idx, int64_t idx_offset);
/* device buffer helpers so a host language without a CUDA binding can stage candidates once */
int  gpd_dev_alloc(gpd_ctx* ctx, uint64_t bytes, void** out_dev);
int  gpd_dev_free(gpd_ctx* ctx, void* dev);
int  gpd_dev_upload(gpd_ctx* ctx, void* dst_dev, const void* src_host, uint64_t bytes);
int  gpd_dev_download(gpd_ctx* ctx, void* dst_host, const void* src_dev, uint64_t bytes);
/* page-locked host staging buffers (the end-to-end path copies candidates from pinned memory) */
int  gpd_host_alloc_pinned(gpd_ctx* ctx, uint64_t bytes, void** out_host);
int  gpd_host_free_pinned(gpd_ctx* ctx, void* host);
/* counter-based uniform candidates U[lo,hi]^D generated on the device: row i of the
 * global stream starts at counter (first_row + i); shards reproduce the same rows */
int  gpd_dev_fill_uniform(gpd_ctx* ctx, double* X_dev, int64_t N, int32_t D, uint64_t seed,
                          int64_t first_row, const double* lo, const double* hi);

/* ---- debug: a memcheck of our own for the scratch region (compute-sanitizer is not available on the target pool) ----
 * with gpd_set_option(ctx, "scratch_redzones", 1) every sub-buffer of the per-call scratch region is followed by a
 * 256-byte red zone; this call reports how many zones the current layout has and how many of their bytes were
 * overwritten (must be 0).  tests/test_scratch_redzones.py runs the ragged-size matrix with it. */
int  gpd_debug_check_scratch(gpd_ctx* ctx, int64_t* zones_out, int64_t* bad_bytes_out);

/* ---- layout self-description (used by the CPU tests to check the DMMA operand packing) ---- */
int32_t gpd_layout_a_index(int32_t ms, int32_t ks, int32_t lane);
int32_t gpd_layout_a_row(int32_t ms, int32_t lane);
int32_t gpd_layout_a_col(int32_t ks, int32_t lane);
int32_t gpd_layout_b_index(int32_t k, int32_t t);
int64_t gpd_layout_tile_offset(int32_t nblk, int32_t I, int32_t kb, int32_t backward);
int32_t gpd_binary_combo_weight(int32_t n_binary, int32_t v);

#ifdef __cplusplus
}
#endif
#endif /* GPDOEMD_B200_H */
