// K1: fused x-kernel row + mean/derivative contraction + right-hand-side panels.
//
// Reference arithmetic: GPy Stationary.K for kx(x*, X) (gp_model.py:199,232,266) followed by the products
// with alpha / the p-part that GPy's predictive_gradients / gradients_XX form (gp_model.py:217,233,275);
// one transcendental per (candidate, training point).
//
// Work item = (GP, tile of 128 candidates).  A CTA of 128 threads is split into CPT row groups; every
// thread owns CPT candidates (register blocking: the per-row constants are read from shared memory once
// per CPT kernel evaluations) and every row group takes every CPT-th training row.  Training rows are
// staged through shared memory as records [x (NXB) | C columns (NCB)], zero padded, so the inner loops
// have compile-time bounds and no predicates.
//   MODE 0: acc[c] += kx * C[c][i]; panel[a][i][t] = kx * G[a][i]        (swizzled for the DMMA kernel)
//   MODE 1: acc[c] += beta[i][t] * kx * H[c][i]                           (second variance derivative)
#pragma once
#include "common.cuh"
#include "exp2_table.h"
#include "log_table.h"

namespace gpd {

__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src)
                 : "memory");
}

// exp(-s) for s >= 0, ~1e-16 relative: -s = (m/256) ln2 + u, |u| <= ln2/512;  e^-s = 2^(m>>8) * 2^((m&255)/256) * e^u
// with a degree-4 Taylor polynomial for e^u (truncation < 4e-17) and a 256-entry table.  9 FP64 instructions
// (the library exp() needs ~17 FP64 + ~10 integer ones); the underflow guard is an integer compare on the high
// word of s, which keeps the FP64 pipe free (s >= 700 or +inf -> 0, NaN propagates).
__device__ __forceinline__ double exp_neg_fast(double s, const double* __restrict__ tab) {
    const double kMagic = 6755399441055744.0;          // 1.5 * 2^52: the FMA rounds -s*256/ln2 to an integer
    const double z = fma(s, -369.3299304675746, kMagic);
    const int m = __double2loint(z);
    const double zr = z - kMagic;
    double u = fma(zr, -0.00270760617331689, -s);      // ln2/256 high part (32 significant bits: m * hi is exact)
    u = fma(zr, -7.453964567463233e-13, u);            // low part
    double p = fma(u, 1.0 / 24.0, 1.0 / 6.0);
    p = fma(p, u, 0.5);
    p = fma(p, u, 1.0);
    p = fma(p, u, 1.0);
    const double r = tab[m & (kExpTabSize - 1)] * p;
    const double v = __hiloint2double(__double2hiint(r) + ((m >> kExpTabBits) << 20), __double2loint(r));
    const unsigned hs = (unsigned)__double2hiint(s);
    return (hs - 0x4085E000u <= 0x7FF00000u - 0x4085E000u) ? 0.0 : v;
}

// The RBF profile exp(-r^2/2) is evaluated on coordinates pre-scaled by sqrt(128 / ln 2) (training rows at upload,
// candidates when they are loaded, kRbfCoordScale in common.cuh): the squared distance then IS the table argument
// t = (r^2 / 2) * 256 / ln 2, and the range reduction needs no multiply at all:
//   m = -round(t) (magic add), d = t - round(t) in [-1/2, 1/2] (exact), e^-s = 2^(m>>8) * 2^((m&255)/256) * e^(-d ln2/256)
// with the degree-4 polynomial in d (coefficients (-ln2/256)^k / k!, truncation < 4e-17): 3 DADD + 4 DFMA + 1 DMUL.
__device__ __forceinline__ double exp_neg_tab_units(double t, const double* __restrict__ tab) {
    const double kMagic = 6755399441055744.0;          // 1.5 * 2^52
    const double z = kMagic - t;
    const int m = __double2loint(z);                   // -round(t)
    const double zr = z - kMagic;
    const double d = t + zr;
    double p = fma(d, 2.239395190875157e-12, -3.308302680541371e-09);
    p = fma(p, d, 3.665565596910106e-06);
    p = fma(p, d, -0.0027076061740622863);
    p = fma(p, d, 1.0);
    const double r = tab[m & (kExpTabSize - 1)] * p;
    const double v = __hiloint2double(__double2hiint(r) + ((m >> kExpTabBits) << 20), __double2loint(r));
    const unsigned ht = (unsigned)__double2hiint(t);   // t >= 258531 (s >= 700) or +inf -> 0, NaN propagates
    return (ht - 0x410F8F17u <= 0x7FF00000u - 0x410F8F17u) ? 0.0 : v;
}

// sqrt(x) for x in [kTinyR2, 1e300], branch-free: hardware reciprocal-square-root estimate (MUFU.RSQ64H, ~2^-20) and two
// coupled Newton / Goldschmidt steps (error 2^-20 -> 2^-39 -> below one ulp; result within ~1 ulp, not correctly rounded).
// The library sqrt() carries a slow-path CALL behind a branch per call, which cut the unrolled pair loop of the Matern /
// Exponential kernels into one basic block per pair; the squared distance is started from kTinyR2 instead of 0 so that the
// estimate never sees 0 (r = 1e-150 at an exact coincidence changes nothing below 1e-150).
constexpr double kTinyR2 = 1e-300;
__device__ __forceinline__ double sqrt_fast_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = x * y, h = 0.5 * y;
    double e = fma(-g, h, 0.5);
    g = fma(g, e, g);
    h = fma(h, e, h);
    e = fma(-g, h, 0.5);
    return fma(g, e, g);
}

// shared-memory tables of one evaluation CTA: 2^(j/256) (kernels 0..3 and 5), then 1 / c_j and ln c_j (kernel 5 only)
constexpr int kKernTabSize = kExpTabSize + 2 * kLogTabSize;
template <int KERN>
__device__ __forceinline__ void load_kernel_tables(double* __restrict__ sTab, int tid, int nthreads) {
    if (KERN <= 3 || KERN == 5)
        for (int k = tid; k < kExpTabSize; k += nthreads) sTab[k] = kExp2Tab[k];
    if (KERN == 5)
        for (int k = tid; k < kLogTabSize; k += nthreads) {
            sTab[kExpTabSize + k] = kLogInvTab[k];
            sTab[kExpTabSize + kLogTabSize + k] = kLogLnTab[k];
        }
}

// ln(u) for u in [1, 2^1000], branch-free, absolute error ~ (1 + k) 1e-16: u = 2^k m with m in [1, 2); the top 7 mantissa
// bits pick c_j, r = m / c_j - 1 (|r| <= 2^-8) and log1p(r) is a degree-6 polynomial (truncation 2e-18).  The library
// log1p / exp pair of the RatQuad profile carried slow-path CALLs behind branches (one basic block per pair, like sqrt).
__device__ __forceinline__ double log_ge1_tab(double u, const double* __restrict__ tlog) {
    const int hi = __double2hiint(u);
    const int k = (hi >> 20) - 1023;
    const int j = (hi >> 13) & (kLogTabSize - 1);
    const double m = __hiloint2double((hi & 0x000FFFFF) | 0x3FF00000, __double2loint(u));
    const double r = fma(m, tlog[j], -1.0);
    double p = fma(r, -1.0 / 6.0, 0.2);
    p = fma(p, r, -0.25);
    p = fma(p, r, 1.0 / 3.0);
    p = fma(p, r, -0.5);
    p = fma(p, r, 1.0);
    return fma((double)k, 0.6931471805599453, fma(p, r, tlog[kLogTabSize + j]));
}

template <int KERN>
__device__ __forceinline__ double kappa_of_r2(double r2, double power, const double* __restrict__ tab) {
    if (KERN == 0) return exp_neg_tab_units(r2, tab);            // RBF: r2 is already (r^2 / 2) * 256 / ln 2
    double r = sqrt_fast_pos(r2);
    if (KERN == 1) return exp_neg_fast(r, tab);                  // Exponential
    if (KERN == 2) {                                             // Matern32
        const double s3 = 1.7320508075688772;
        const double a = s3 * r;
        return (1.0 + a) * exp_neg_fast(a, tab);
    }
    if (KERN == 3) {                                             // Matern52
        const double s5 = 2.23606797749979;
        const double a = s5 * r;
        return fma(r2, 5.0 / 3.0, 1.0 + a) * exp_neg_fast(a, tab);
    }
    if (KERN == 4) return cos(r);                                // Cosine
    // RatQuad (1 + r^2 / 2)^-power = exp(-power ln u), power > 0; (r2 - r2) carries a NaN design through
    return exp_neg_fast(fma(power, log_ge1_tab(fma(0.5, r2, 1.0), tab + kExpTabSize), r2 - r2), tab);
}


// Candidate (column of the 128-wide tile) of slot j of thread s in a row group of TPG threads: adjacent pairs
// (2s, 2s+1) so that one 16-byte store / load per pair serves the swizzled panels (the swizzle flips bits 2..3 only).
template <int CPT, int TPG>
__device__ __forceinline__ int eval_col(int s, int j) {
    return (CPT == 1) ? s : 2 * s + (j & 1) + (j >> 1) * 2 * TPG;
}

#ifndef GPD_EVAL_UNROLL
#define GPD_EVAL_UNROLL 4
#endif
// training rows per unrolled loop body of eval_kernel: 4 for the RBF kernel's small column buckets (C2: 0.843 -> 0.808 ms against
// 2; more independent FP64 chains per warp at 25 % occupancy), 2 where the per-row coefficient registers dominate
template <int KERN, int NCB>
struct EvalUnroll { static constexpr int value = (KERN <= 3 && NCB <= 5) ? GPD_EVAL_UNROLL : 2; };   // larger buckets spill at 4

// DB: the record / G staging buffers are double buffered and filled with cp.async (single-RHS small buckets)
template <int NXB, int NCB, int CPT, int RG, int MODE, bool DB = false>
struct EvalSmem {
#ifndef GPD_EVAL_ROWS
#define GPD_EVAL_ROWS 128
#endif
    static constexpr int kRows = (NCB > 28) ? 32 : ((NCB <= 11) ? GPD_EVAL_ROWS : 64);   // training rows per shared-memory stage
    static constexpr int RS = (NXB + NCB + 1) & ~1;              // record stride, even (16-byte vector loads)
    static constexpr int kRec = kRows * RS;
    static constexpr int kG = (MODE == 0) ? kRows * (DB ? 1 : kMaxP + 1) : 0;
    static constexpr int kRed = (RG > 1) ? kBlk * NCB : 0;
    static constexpr int kStage = kRec + kG;                     // one staging buffer
    static constexpr int kMain = kStage * (DB ? 2 : 1);
    static constexpr int kDoubles = (kMain > kRed ? kMain : kRed) + kKernTabSize;
};

template <int KERN, int NXB, int NCB, int CPT, int RG, int MODE, bool ONE_RHS>
#ifndef GPD_EVAL_TPSM
#define GPD_EVAL_TPSM 512         // resident threads per SM the small-bucket kernels are compiled for (640 / 768: measured, see profiles)
#endif
#ifndef GPD_EVAL_CPT_MID
#define GPD_EVAL_CPT_MID 4        // candidates per thread of the 11-column bucket (first order with P = 6..10)
#define GPD_EVAL_MINB_MID 1       // resident CTAs per SM the compiler must allow for it
#endif
__global__ void __launch_bounds__(RG * (kBlk / CPT), (NCB <= 8) ? (CPT == 4 ? GPD_EVAL_TPSM : 1024) / (RG * (kBlk / CPT)) : (NCB <= 11 ? GPD_EVAL_MINB_MID : 1)) eval_kernel(const GpView* __restrict__ gps, const Item* __restrict__ items,
                                                            const double* __restrict__ X, int ldx, int chunk_n, int nrhs,
                                                            int ncols, double* __restrict__ panels,
                                                            const double* __restrict__ beta_panels,
                                                            const RawOut* __restrict__ raws, int ldraw) {
    // single right-hand side + small column bucket (the first-order scoring path): staging is double buffered
    constexpr bool DB = ONE_RHS && MODE == 0 && NCB <= 11;
    // RBF: the squared distance is formed in EXPANDED form, t = |x|^2 + |X_i|^2 - 2 x.X_i (4 instead of 6 FP64
    // instructions per pair; the absolute error u (|x|^2 + |X_i|^2) of t is a relative 1e-14 of exp(-t ln2/256), far
    // inside the 1e-9 gate -- it is also the form GPy itself uses).  The record then carries -2 X_i and |X_i|^2
    // (GpView::XsE, nx + 1 doubles per row) instead of X_i.
    constexpr bool kExpanded = (KERN == 0);
    constexpr int NXR = NXB + (kExpanded ? 1 : 0);               // x slots of a record
    using SM = EvalSmem<NXR, NCB, CPT, RG, MODE, DB>;
    constexpr int RS = SM::RS;
    constexpr int kEvalRows = SM::kRows;
    constexpr int TPG = kBlk / CPT;                              // threads per row group (each covers the 128 candidates)
    constexpr int kEvalThreads = RG * TPG;
    constexpr int kEvalUnroll = EvalUnroll<KERN, NCB>::value;
    __shared__ __align__(16) double smem[SM::kDoubles];
    double* sRec = smem;
    double* sG = smem + SM::kRec;                                // DB: buffer b lives at smem + b * SM::kStage
    double* sRed = smem;                                         // reused after the main loop
    double* sTab = smem + (SM::kDoubles - kKernTabSize);

    const Item it = items[blockIdx.x];
    const GpView& g = gps[it.gp];
    const int count = g.count ? *g.count : chunk_n;
    const int slot0 = it.tile * kBlk;
    if (slot0 >= count) return;
    const int tid = threadIdx.x;
    const int rg = tid / TPG, s = tid % TPG;
    const int n_pad = g.n_pad, nx = g.nx;
    const double power = g.power_x;
    load_kernel_tables<KERN>(sTab, tid, kEvalThreads);           // kernels evaluated with exp_neg_fast / log_ge1_tab

    double xs[CPT][NXB], cx[CPT];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int slot = slot0 + eval_col<CPT, TPG>(s, j);
        const bool valid = slot < count;
        const int cand = valid ? (g.idx ? g.idx[slot] : slot) : 0;
        cx[j] = 0.0;
#pragma unroll
        for (int d = 0; d < NXB; ++d) {
            double v = 0.0;
            if (d < nx && valid) {
                const double x = X[(size_t)cand * ldx + g.x_active[d]];
                v = (x - g.xmin[d]) * g.xscale[d];              // trans_x, then GPy's X / lengthscale, as one multiply
            }
            xs[j][d] = v;
            cx[j] = fma(v, v, cx[j]);                            // |x|^2 of the expanded form
        }
    }
    double acc[CPT][NCB];
#pragma unroll
    for (int j = 0; j < CPT; ++j)
#pragma unroll
        for (int c = 0; c < NCB; ++c) acc[j][c] = 0.0;

    const double* __restrict__ gX = kExpanded ? g.XsE : g.Xs;
    const int nxs = kExpanded ? nx + 1 : nx;                     // doubles per training row in gX
    const double* __restrict__ gC = (MODE == 0) ? g.C : g.H;
    const double* __restrict__ gG = g.G;
    double* __restrict__ pan = nullptr;
    const double* __restrict__ bpan = nullptr;
    if (MODE == 0) pan = panels + (nrhs == 1 ? g.panel_off : g.panel_off_w) + (size_t)it.tile * nrhs * n_pad * kBlk;
    else bpan = beta_panels + g.panel_off + (size_t)it.tile * n_pad * kBlk;   // beta panels hold one RHS kind

    // DB staging: asynchronous 8-byte copies of the next kEvalRows records while the current ones are consumed.
    // Padding slots (d >= nx, c >= ncols) are zeroed once and never written again.
    auto stage_async = [&](int i0, int buf) {
        double* dRec = smem + buf * SM::kStage;
        double* dG = dRec + SM::kRec;
        for (int k = tid; k < kEvalRows * NXR; k += kEvalThreads) {
            const int row = k / NXR, d = k % NXR;
            // expanded form: source column nx (|X_i|^2) goes to the last x slot, the padding slots in between stay 0
            if (d < nx) cp_async8(dRec + row * RS + d, gX + (size_t)(i0 + row) * nxs + d);
            else if (kExpanded && d == NXB) cp_async8(dRec + row * RS + d, gX + (size_t)(i0 + row) * nxs + nx);
        }
        for (int k = tid; k < kEvalRows * NCB; k += kEvalThreads) {
            const int c = k / kEvalRows, row = k % kEvalRows;
            if (c < ncols) cp_async8(dRec + row * RS + NXR + c, gC + (size_t)c * n_pad + i0 + row);
        }
        for (int k = tid; k < kEvalRows; k += kEvalThreads) cp_async8(dG + k, gG + i0 + k);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (DB) {
        for (int k = tid; k < SM::kMain; k += kEvalThreads) smem[k] = 0.0;
        __syncthreads();
        stage_async(0, 0);
    }
    for (int i0 = 0; i0 < n_pad; i0 += kEvalRows) {
        if (DB) {
            const int buf = (i0 / kEvalRows) & 1;
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            __syncthreads();                                     // stage landed; everyone is done with the other buffer
            if (i0 + kEvalRows < n_pad) stage_async(i0 + kEvalRows, buf ^ 1);
            sRec = smem + buf * SM::kStage;
            sG = sRec + SM::kRec;
        } else {
        __syncthreads();
        for (int k = tid; k < kEvalRows * NXR; k += kEvalThreads) {
            const int row = k / NXR, d = k % NXR;
            double v = 0.0;
            if (d < nx) v = gX[(size_t)(i0 + row) * nxs + d];
            else if (kExpanded && d == NXB) v = gX[(size_t)(i0 + row) * nxs + nx];
            sRec[row * RS + d] = v;
        }
        for (int k = tid; k < kEvalRows * NCB; k += kEvalThreads) {
            const int c = k / kEvalRows, row = k % kEvalRows;
            sRec[row * RS + NXR + c] = (c < ncols) ? gC[(size_t)c * n_pad + i0 + row] : 0.0;
        }
        if (MODE == 0) {
            for (int k = tid; k < kEvalRows * nrhs; k += kEvalThreads) {
                const int a = k / kEvalRows, row = k % kEvalRows;
                sG[a * kEvalRows + row] = gG[(size_t)a * n_pad + i0 + row];
            }
        }
        __syncthreads();
        }
#pragma unroll kEvalUnroll
        for (int ii = rg; ii < kEvalRows; ii += RG) {
            const double* __restrict__ rec = sRec + ii * RS;
            double xr[NXR], cr[NCB];
            if (NXR % 2 == 0 && !kExpanded) {
#pragma unroll
                for (int d = 0; d < NXB; d += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(rec + d);
                    xr[d] = v.x;
                    xr[d + 1] = v.y;
                }
#pragma unroll
                for (int c = 0; c < NCB; ++c) cr[c] = rec[NXB + c];
            } else {
                // odd NXB / expanded form: the record [x | C columns] is read as RS/2 16-byte vectors
                double rv[RS];
#pragma unroll
                for (int d = 0; d < RS; d += 2) {
                    const double2 v = *reinterpret_cast<const double2*>(rec + d);
                    rv[d] = v.x;
                    rv[d + 1] = v.y;
                }
#pragma unroll
                for (int d = 0; d < NXR; ++d) xr[d] = rv[d];
#pragma unroll
                for (int c = 0; c < NCB; ++c) cr[c] = rv[NXR + c];
            }
            const int row = i0 + ii;
            const int sw = (row & 3) << 2;
            double kap[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                double r2;
                if (kExpanded) {
                    r2 = cx[j] + xr[NXB];                        // |x|^2 + |X_i|^2
#pragma unroll
                    for (int d = 0; d < NXB; ++d) r2 = fma(xs[j][d], xr[d], r2);   // xr = -2 X_i
                } else {
                    r2 = (KERN == 5) ? 0.0 : kTinyR2;            // RatQuad uses r^2 itself, the others take sqrt_fast_pos
#pragma unroll
                    for (int d = 0; d < NXB; ++d) {
                        const double df = xs[j][d] - xr[d];
                        r2 = fma(df, df, r2);
                    }
                }
                kap[j] = kappa_of_r2<KERN>(r2, power, sTab);
            }
            if (MODE == 1) {
                const double* __restrict__ brow = bpan + (size_t)row * kBlk;
                if (CPT == 1) {
                    kap[0] *= brow[s ^ sw];
                } else {
#pragma unroll
                    for (int jp = 0; jp < CPT / 2; ++jp) {
                        const double2 b2 = *reinterpret_cast<const double2*>(brow + (eval_col<CPT, TPG>(s, 2 * jp) ^ sw));
                        kap[2 * jp] *= b2.x;
                        kap[2 * jp + 1] *= b2.y;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j)
#pragma unroll
                for (int c = 0; c < NCB; ++c) acc[j][c] = fma(kap[j], cr[c], acc[j][c]);
            if (MODE == 0) {
                const int na = ONE_RHS ? 1 : nrhs;
                for (int a = 0; a < na; ++a) {
                    const double ga = sG[a * kEvalRows + ii];
                    double* __restrict__ prow = pan + ((size_t)a * n_pad + row) * kBlk;
                    if (CPT == 1) {
                        prow[s ^ sw] = kap[0] * ga;
                    } else {
#pragma unroll
                        for (int jp = 0; jp < CPT / 2; ++jp)
                            *reinterpret_cast<double2*>(prow + (eval_col<CPT, TPG>(s, 2 * jp) ^ sw)) =
                                make_double2(kap[2 * jp] * ga, kap[2 * jp + 1] * ga);
                    }
                }
            }
        }
    }
    // ---- combine the row groups (group 0 accumulates), then write the raw contractions ----
    if (RG > 1) {
#pragma unroll 1
        for (int gsrc = 1; gsrc < RG; ++gsrc) {
            __syncthreads();
            if (rg == gsrc) {
#pragma unroll
                for (int j = 0; j < CPT; ++j)
#pragma unroll
                    for (int c = 0; c < NCB; ++c) sRed[eval_col<CPT, TPG>(s, j) * NCB + c] = acc[j][c];
            }
            __syncthreads();
            if (rg == 0) {
#pragma unroll
                for (int j = 0; j < CPT; ++j)
#pragma unroll
                    for (int c = 0; c < NCB; ++c) acc[j][c] += sRed[eval_col<CPT, TPG>(s, j) * NCB + c];
            }
        }
    }
    if (rg != 0) return;
    const RawOut& raw = raws[g.model];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int slot = slot0 + eval_col<CPT, TPG>(s, j);
        if (slot >= count) continue;
        const int cand = g.idx ? g.idx[slot] : slot;
        const size_t orow = (size_t)g.out_e * chunk_n + cand;
        if (MODE == 0) {
#pragma unroll
            for (int c = 0; c < NCB; ++c)
                if (c < ncols) raw.cmu[orow * ldraw + c] = acc[j][c];
            raw.kss[orow] = g.kss;
        } else {
            const int np = g.P * (g.P + 1) / 2;
#pragma unroll
            for (int c = 0; c < NCB; ++c)
                if (c < ncols) raw.bh[orow * np + c] = acc[j][c];
        }
    }
}

// ---- bucket dispatch: NXB in {2,3,4,8}; MODE 0 column buckets {2,5,8,11,17,28,72}; MODE 1 {3,10,28,55} ----
template <int KERN, int NXB, int NCB, int CPT, int RG, int MODE>
static int launch_one(cudaStream_t s, int nitems, int ncols, const GpView* gps, const Item* items, const double* X,
                      int ldx, int chunk_n, int nrhs, double* panels, const double* beta, const RawOut* raws,
                      int ldraw) {
    constexpr int threads = RG * (kBlk / CPT);
    if (MODE == 0 && nrhs == 1 && NCB <= 17)
        eval_kernel<KERN, NXB, NCB, CPT, RG, MODE, true><<<nitems, threads, 0, s>>>(gps, items, X, ldx, chunk_n, nrhs, ncols,
                                                                                    panels, beta, raws, ldraw);
    else
        eval_kernel<KERN, NXB, NCB, CPT, RG, MODE, false><<<nitems, threads, 0, s>>>(gps, items, X, ldx, chunk_n, nrhs, ncols,
                                                                                     panels, beta, raws, ldraw);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

template <int KERN, int NXB>
static int dispatch_cols(cudaStream_t s, int mode, int nitems, int ncols, const GpView* gps, const Item* items,
                         const double* X, int ldx, int chunk_n, int nrhs, double* panels, const double* beta,
                         const RawOut* raws, int ldraw) {
#ifndef GPD_EVAL_CPT_SMALL
#define GPD_EVAL_CPT_SMALL 4
#define GPD_EVAL_RG_SMALL 4
#endif
#define GPD_GO(NCB, CPT, MODE) \
    return launch_one<KERN, NXB, NCB, CPT, CPT, MODE>(s, nitems, ncols, gps, items, X, ldx, chunk_n, nrhs, panels, beta, raws, ldraw)
#define GPD_GO_SMALL(NCB, MODE) \
    return launch_one<KERN, NXB, NCB, GPD_EVAL_CPT_SMALL, GPD_EVAL_RG_SMALL, MODE>(s, nitems, ncols, gps, items, X, ldx, chunk_n, nrhs, panels, beta, raws, ldraw)
    if (mode == 0) {
        if (ncols <= 2) GPD_GO_SMALL(2, 0);
        if (ncols <= 5) GPD_GO_SMALL(5, 0);
        if (ncols <= 8) GPD_GO_SMALL(8, 0);
        if (ncols <= 11) GPD_GO(11, GPD_EVAL_CPT_MID, 0);
        if (ncols <= 17) GPD_GO(17, 2, 0);
        if (ncols <= 28) GPD_GO(28, 2, 0);
        if (ncols <= 72) GPD_GO(72, 1, 0);
        set_error("evaluation kernel supports at most 72 coefficient columns (P <= 10 for second order)");
        return 3;
    }
    if (ncols <= 3) GPD_GO(3, 4, 1);
    if (ncols <= 10) GPD_GO(10, 4, 1);
    if (ncols <= 28) GPD_GO(28, 2, 1);
    if (ncols <= 55) GPD_GO(55, 1, 1);
    set_error("second variance derivative supports P <= 10 model parameters");
    return 3;
#undef GPD_GO
#undef GPD_GO_SMALL
}

template <int KERN>
static int dispatch_nx(cudaStream_t s, int mode, int nitems, int ncols, int nx, const GpView* gps, const Item* items,
                       const double* X, int ldx, int chunk_n, int nrhs, double* panels, const double* beta,
                       const RawOut* raws, int ldraw) {
    if (nx <= 2) return dispatch_cols<KERN, 2>(s, mode, nitems, ncols, gps, items, X, ldx, chunk_n, nrhs, panels, beta, raws, ldraw);
    if (nx == 3) return dispatch_cols<KERN, 3>(s, mode, nitems, ncols, gps, items, X, ldx, chunk_n, nrhs, panels, beta, raws, ldraw);
    if (nx <= 4) return dispatch_cols<KERN, 4>(s, mode, nitems, ncols, gps, items, X, ldx, chunk_n, nrhs, panels, beta, raws, ldraw);
    if (nx <= 8) return dispatch_cols<KERN, 8>(s, mode, nitems, ncols, gps, items, X, ldx, chunk_n, nrhs, panels, beta, raws, ldraw);
    set_error("x-kernel acts on 1..8 non-binary design dimensions");
    return 3;
}

#define GPD_DEFINE_EVAL_TU(K)                                                                                  \
    int launch_eval_k##K(cudaStream_t s, int mode, int nitems, int ncols, int nx, const GpView* gps,            \
                         const Item* items, const double* X, int ldx, int chunk_n, int nrhs, double* panels,   \
                         const double* beta, const RawOut* raws, int ldraw) {                                  \
        return dispatch_nx<K>(s, mode, nitems, ncols, nx, gps, items, X, ldx, chunk_n, nrhs, panels, beta, raws, ldraw); \
    }

}  // namespace gpd
