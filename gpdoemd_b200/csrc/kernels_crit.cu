// K3/K4: routing, unit conversion, Taylor marginal assembly, design criteria, argmax.
//
// Reference arithmetic:
//   routing      GPdoemd/utils.py:27-48 (binary_dimensions)
//   units        GPdoemd/models/surrogate_model.py:170-215, GPdoemd/transform.py:27-73
//   marginals    GPdoemd/marginal/taylor_first_order.py:27-43, taylor_second_order.py:27-60
//   criteria     GPdoemd/design_criteria.py:49-194 (HR, BH, BF, AW, JR)
//   argmax       np.argmax in demos/case_study.py:291 (first index of the maximum, NaN wins)
#include <math.h>

#include "common.cuh"

namespace gpd {

// ---------------------------------------------------------------------------------------------
// routing: candidate -> binary-combination index, compacted position lists per combination
// ---------------------------------------------------------------------------------------------
__global__ void route_kernel(const double* __restrict__ X, int ldx, int chunk_n, int nb, const int* __restrict__ bcols,
                             const int* __restrict__ weights, int* __restrict__ idx, int* __restrict__ counts,
                             int idx_stride) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= chunk_n) return;
    int r = 0;
    for (int v = 0; v < nb; ++v) {
        // binary columns are box-scaled with min 0 / max 1 (surrogate_model.py:80-82): z = x
        double z = X[(size_t)i * ldx + bcols[v]] - 0.5;
        if (z > 0.0) r += weights[v];
    }
    int pos = atomicAdd(&counts[r], 1);
    idx[(size_t)r * idx_stride + pos] = i;
}

int launch_route(cudaStream_t s, const double* X_dev, int ldx, int chunk_n, int nb, const int* bcols,
                 const int* weights, int ncombo, int* idx, int* counts, int idx_stride) {
    GPD_CUDA(cudaMemsetAsync(counts, 0, sizeof(int) * ncombo, s));
    route_kernel<<<(chunk_n + 255) / 256, 256, 0, s>>>(X_dev, ldx, chunk_n, nb, bcols, weights, idx, counts, idx_stride);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// finalize: raw contractions (transformed units) -> hook outputs (original units)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int gram_index(int NR, int a, int b) { return a * NR - a * (a - 1) / 2 + (b - a); }

__global__ void finalize_kernel(const ModelXform* __restrict__ xf, const RawOut* __restrict__ raws, int model_slot,
                                int chunk_n, int ldraw, int flags, int nrhs, int npart, double* __restrict__ mu, int ldmu,
                                double* __restrict__ s2, double* __restrict__ dmu, double* __restrict__ d2mu,
                                double* __restrict__ ds2, double* __restrict__ d2s2) {
    const int E = xf->E, P = xf->P;
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)chunk_n * E) return;
    const int n = (int)(tid / E), e = (int)(tid % E);
    const RawOut& raw = raws[model_slot];
    const size_t row = (size_t)e * chunk_n + n;
    const double* c = raw.cmu + row * ldraw;
    const double ystd = xf->ystd[e];
    // first order: npart row-group partials of ||v||^2 from tri_kernel, added in a fixed order;
    // otherwise the upper triangle of [v W]'[v W] from gram_kernel
    const int ng = npart > 0 ? npart : nrhs * (nrhs + 1) / 2;
    const double* gr = raw.gram + row * ng;
    double vv = gr[0];
    for (int k = 1; k < npart; ++k) vv += gr[k];
    mu[(size_t)n * ldmu + e] = xf->ymean[e] + c[0] * ystd;                // transform.py:69-73 back=True
    s2[(size_t)n * E + e] = fmax(raw.kss[row] - vv, xf->var_floor) * (ystd * ystd);   // k** - v'v  (GPy _raw_predict)
    if (flags & 1) {                                                     // d_mu_d_p  surrogate_model.py:185-189
        double* o = dmu + ((size_t)n * E + e) * P;
        for (int a = 0; a < P; ++a) o[a] = c[1 + a] * ystd / xf->pdiff[a];
    }
    if (flags & 2) {                                                     // d2_mu_d_p2  :193-198
        double* o = d2mu + ((size_t)n * E + e) * P * P;
        int pr = 0;
        for (int a = 0; a < P; ++a)
            for (int b = a; b < P; ++b, ++pr) {
                double v = c[1 + P + pr] * ystd / (xf->pdiff[a] * xf->pdiff[b]);
                o[a * P + b] = v;
                o[b * P + a] = v;
            }
    }
    if (flags & 4) {                                                     // d_s2_d_p  :202-206 ; -2 W'v
        double* o = ds2 + ((size_t)n * E + e) * P;
        for (int a = 0; a < P; ++a) o[a] = (-2.0 * gr[gram_index(nrhs, 0, 1 + a)]) * (ystd * ystd) / xf->pdiff[a];
    }
    if (flags & 8) {                                                     // d2_s2_d_p2  :210-215
        double* o = d2s2 + ((size_t)n * E + e) * P * P;
        const double* bh = raw.bh + row * (P * (P + 1) / 2);
        int pr = 0;
        for (int a = 0; a < P; ++a)
            for (int b = a; b < P; ++b, ++pr) {
                double v = -2.0 * (bh[pr] + gr[gram_index(nrhs, 1 + a, 1 + b)]);
                v = v * (ystd * ystd) / (xf->pdiff[a] * xf->pdiff[b]);
                o[a * P + b] = v;
                o[b * P + a] = v;
            }
    }
}

int launch_finalize(cudaStream_t s, const ModelXform* xf_dev, const RawOut* raw_dev, int model_slot, int E, int P,
                    int chunk_n, int ldraw, int flags, int nrhs, int npart, double* mu, int ldmu, double* s2,
                    double* dmu, double* d2mu, double* ds2, double* d2s2) {
    (void)P;
    long long total = (long long)chunk_n * E;
    finalize_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(xf_dev, raw_dev, model_slot, chunk_n, ldraw,
                                                                    flags, nrhs, npart, mu, ldmu, s2, dmu, d2mu, ds2, d2s2);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// Taylor marginal assembly.  mu/S may be strided so several models interleave into (N, M, E[,E]).
// ---------------------------------------------------------------------------------------------
// second order, step 1 (taylor_second_order.py:40-45): A_e = H_mu,e Sigma ; mu += tr(A)/2 ; s2 += sum(H_s2*Sigma)/2
__global__ void assemble2_prep_kernel(int N, int E, int P, double* __restrict__ mu, int ldmu,
                                      const double* __restrict__ s2_in, double* __restrict__ s2_out,
                                      const double* __restrict__ d2mu, const double* __restrict__ d2s2,
                                      const double* __restrict__ Sigma, double* __restrict__ A) {
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)N * E) return;
    const int n = (int)(tid / E), e = (int)(tid % E);
    const double* Hm = d2mu + (size_t)tid * P * P;
    const double* Hs = d2s2 + (size_t)tid * P * P;
    double* Ae = A + (size_t)tid * P * P;
    double tr = 0.0, corr = 0.0;
    for (int a = 0; a < P; ++a)
        for (int b = 0; b < P; ++b) {
            double v = 0.0;
            for (int k = 0; k < P; ++k) v = fma(Hm[a * P + k], Sigma[k * P + b], v);
            Ae[a * P + b] = v;
            if (a == b) tr += v;
            corr += Hs[a * P + b] * Sigma[a * P + b];
        }
    mu[(size_t)n * ldmu + e] += 0.5 * tr;
    s2_out[tid] = s2_in[tid] + 0.5 * corr;
}

// S[n,e1,e2] = [e1==e2] s2 + J_e1 Sigma J_e2^T (+ 1/2 sum(A_e1 * A_e2^T)), diagonal clipped at 1e-15
__global__ void assemble_kernel(int order, int N, int E, int P, const double* __restrict__ s2,
                                const double* __restrict__ dmu, const double* __restrict__ Sigma,
                                const double* __restrict__ A, double* __restrict__ S, int ldS) {
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)N * E * E) return;
    const int n = (int)(tid / (E * E));
    int e1 = (int)((tid / E) % E), e2 = (int)(tid % E);
    const int r1 = e1, r2 = e2;
    if (order == 2 && e1 > e2) { int t = e1; e1 = e2; e2 = t; }   // taylor_second_order.py:56-57 mirrors the upper triangle
    const double* J1 = dmu + ((size_t)n * E + e1) * P;
    const double* J2 = dmu + ((size_t)n * E + e2) * P;
    double v = 0.0;
    for (int a = 0; a < P; ++a) {
        double w = 0.0;
        for (int b = 0; b < P; ++b) w = fma(Sigma[a * P + b], J2[b], w);
        v = fma(J1[a], w, v);
    }
    if (e1 == e2) v = s2[(size_t)n * E + e1] + v;
    if (order == 2) {
        const double* A1 = A + ((size_t)n * E + e1) * P * P;
        const double* A2 = A + ((size_t)n * E + e2) * P * P;
        double q = 0.0;
        for (int a = 0; a < P; ++a)
            for (int b = 0; b < P; ++b) q = fma(A1[a * P + b], A2[b * P + a], q);
        v += 0.5 * q;
    }
    if (e1 == e2) v = fmax(1e-15, v);
    S[(size_t)n * ldS + r1 * E + r2] = v;
}

int launch_assemble(cudaStream_t s, int order, int N, int E, int P, double* mu, int ldmu, const double* s2,
                    const double* dmu, const double* d2mu, const double* d2s2, const double* Sigma_dev, double* S,
                    int ldS, double* tmpA, double* tmp_s2) {
    const double* s2_use = s2;
    if (order == 2) {
        long long t1 = (long long)N * E;
        assemble2_prep_kernel<<<(unsigned)((t1 + 127) / 128), 128, 0, s>>>(N, E, P, mu, ldmu, s2, tmp_s2, d2mu, d2s2,
                                                                           Sigma_dev, tmpA);
        GPD_CUDA(cudaGetLastError());
        s2_use = tmp_s2;
    }
    long long total = (long long)N * E * E;
    assemble_kernel<<<(unsigned)((total + 127) / 128), 128, 0, s>>>(order, N, E, P, s2_use, dmu, Sigma_dev, tmpA, S,
                                                                    ldS);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// register-resident E x E inverse + determinant (Gauss-Jordan, partial pivoting like LAPACK getrf)
// ---------------------------------------------------------------------------------------------
// An exactly zero pivot = a singular matrix: np.linalg.inv (design_criteria.py:83,115,137,173) raises LinAlgError there.
// The kernels cannot raise, so they set this per-device flag; the host entry points read and reset it at their final
// synchronisation and return GPD_ERR_SINGULAR.  (All writers store 1: the race is benign.)
__device__ int g_singular_flag = 0;

template <int E>
__device__ __forceinline__ double inv_inplace(double (&A)[E][E]) {
    int piv[E];
    double det = 1.0;
#pragma unroll
    for (int c = 0; c < E; ++c) {
        int p = c;
        double best = fabs(A[c][c]);
#pragma unroll
        for (int r = c + 1; r < E; ++r) {
            double v = fabs(A[r][c]);
            if (v > best) { best = v; p = r; }
        }
        piv[c] = p;
#pragma unroll
        for (int r = c + 1; r < E; ++r)
            if (p == r) {
#pragma unroll
                for (int k = 0; k < E; ++k) { double t = A[c][k]; A[c][k] = A[r][k]; A[r][k] = t; }
                det = -det;
            }
        const double d = A[c][c];
        if (d == 0.0) g_singular_flag = 1;
        det *= d;
        const double id = 1.0 / d;
        A[c][c] = 1.0;
#pragma unroll
        for (int k = 0; k < E; ++k) A[c][k] *= id;
#pragma unroll
        for (int r = 0; r < E; ++r)
            if (r != c) {
                const double f = A[r][c];
                A[r][c] = 0.0;
#pragma unroll
                for (int k = 0; k < E; ++k) A[r][k] = fma(-f, A[c][k], A[r][k]);
            }
    }
#pragma unroll
    for (int c = E - 1; c >= 0; --c)
#pragma unroll
        for (int r = c + 1; r < E; ++r)
            if (piv[c] == r) {
#pragma unroll
                for (int k = 0; k < E; ++k) { double t = A[k][c]; A[k][c] = A[k][r]; A[k][r] = t; }
            }
    return det;
}

// work layout (structure of arrays, coalesced over candidates):
//   Sn [M][E][E][N]  s2 + noise      iS [M][E][E][N]  inverse      dS [M][N]  determinant     mu [M][E][N]
struct CritWork {
    double *Sn, *iS, *dS, *mu;
};

template <int E>
__global__ void __launch_bounds__(128) crit_prep_kernel(const double* __restrict__ mu, const double* __restrict__ S, int N,
                                                        int M, const double* __restrict__ noise, CritWork w, int need_inv) {
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)N * M) return;
    const int m = (int)(tid / N), n = (int)(tid % N);          // consecutive threads = consecutive candidates
    const double* Sp = S + ((size_t)n * M + m) * E * E;
    double A[E][E];
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b) {
            A[a][b] = Sp[a * E + b] + noise[a * E + b];
            w.Sn[((size_t)(m * E + a) * E + b) * N + n] = A[a][b];
        }
#pragma unroll
    for (int a = 0; a < E; ++a) w.mu[((size_t)m * E + a) * N + n] = mu[((size_t)n * M + m) * E + a];
    if (!need_inv) return;
    const double det = inv_inplace<E>(A);
    w.dS[(size_t)m * N + n] = det;
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b) w.iS[((size_t)(m * E + a) * E + b) * N + n] = A[a][b];
}

#define SOA(ptr, m, a, b) ptr[((size_t)((m) * E + (a)) * E + (b)) * N + n]
#define MUS(m, a) w.mu[((size_t)(m) * E + (a)) * N + n]

template <int E>
__device__ __forceinline__ double crit_eval(int kind, int n, int N, int M, const double* __restrict__ noise,
                                            const double* __restrict__ pps, const CritWork& w) {
    double out = 0.0;
    if (kind == 0) {                                   // HR  design_criteria.py:49-63
        for (int i = 0; i < M - 1; ++i)
            for (int j = i + 1; j < M; ++j) {
                double s = 0.0;
#pragma unroll
                for (int a = 0; a < E; ++a) { double d = MUS(i, a) - MUS(j, a); s += d * d; }
                out += s;
            }
    } else if (kind == 1) {                            // BH  :66-92
        for (int i = 0; i < M - 1; ++i)
            for (int j = i + 1; j < M; ++j) {
                double t1 = 0.0, t2 = 0.0;
                double r[E];
#pragma unroll
                for (int a = 0; a < E; ++a) r[a] = MUS(i, a) - MUS(j, a);
#pragma unroll
                for (int a = 0; a < E; ++a) {
                    double p1 = 0.0, p2 = 0.0, q = 0.0;
#pragma unroll
                    for (int b = 0; b < E; ++b) {
                        const double isi = SOA(w.iS, i, b, a), isj = SOA(w.iS, j, b, a);
                        p1 = fma(SOA(w.Sn, i, a, b), isj, p1);
                        p2 = fma(SOA(w.Sn, j, a, b), isi, p2);
                        q = fma(SOA(w.iS, i, a, b) + SOA(w.iS, j, a, b), r[b], q);
                    }
                    t1 += (p1 + p2) - 2.0;
                    t2 = fma(r[a], q, t2);
                }
                out += pps[i] * pps[j] * (t1 + t2);
            }
        out *= 0.5;
    } else if (kind == 2) {                            // BF  :95-123
        for (int i = 0; i < M - 1; ++i)
            for (int j = i + 1; j < M; ++j) {
                double A[E][E], r[E];
#pragma unroll
                for (int a = 0; a < E; ++a) {
                    r[a] = MUS(i, a) - MUS(j, a);
#pragma unroll
                    for (int b = 0; b < E; ++b) A[a][b] = SOA(w.Sn, i, a, b) + SOA(w.Sn, j, a, b);
                }
                inv_inplace<E>(A);
                double t1 = 0.0, t2 = 0.0;
#pragma unroll
                for (int a = 0; a < E; ++a) {
                    double q = 0.0;
#pragma unroll
                    for (int b = 0; b < E; ++b) {
                        t1 = fma(noise[a * E + b], A[b][a], t1);
                        q = fma(A[a][b], r[b], q);
                    }
                    t2 = fma(r[a], q, t2);
                }
                out += t1 + t2;
            }
    } else if (kind == 3) {                            // AW  :126-144
        for (int i = 0; i < M; ++i) {
            double wsum = 0.0;
            for (int j = 0; j < M; ++j) {
                double r[E], t1 = 0.0;
#pragma unroll
                for (int a = 0; a < E; ++a) r[a] = MUS(i, a) - MUS(j, a);
#pragma unroll
                for (int a = 0; a < E; ++a) {
                    double q = 0.0;
#pragma unroll
                    for (int b = 0; b < E; ++b) q = fma(SOA(w.iS, i, a, b), r[b], q);
                    t1 = fma(r[a], q, t1);
                }
                wsum += exp(-0.5 * t1);
            }
            out += (1.0 / wsum) * pps[i];
        }
    } else {                                           // JR  :147-194
        const double log4pi = 2.5310242469692907, log2pi = 1.8378770664093453;
        double T1 = 0.0, T2 = 0.0;
        const double scale = exp2(0.5 * E);
        for (int i = 0; i < M; ++i) {
            const double dSi = w.dS[(size_t)i * N + n];
            T1 += pps[i] * ((0.5 * E) * log4pi + 0.5 * log(dSi));
            T2 += pps[i] * pps[i] / (scale * sqrt(dSi));
        }
        for (int i = 0; i < M; ++i) {
            double iSmi[E], miiSmi = 0.0;
#pragma unroll
            for (int a = 0; a < E; ++a) {
                double q = 0.0;
#pragma unroll
                for (int b = 0; b < E; ++b) q = fma(SOA(w.iS, i, a, b), MUS(i, b), q);
                iSmi[a] = q;
                miiSmi = fma(MUS(i, a), q, miiSmi);
            }
            const double ldSi = log(w.dS[(size_t)i * N + n]);
            for (int j = i + 1; j < M; ++j) {
                double A[E][E], mij[E], mjiSmj = 0.0;
#pragma unroll
                for (int a = 0; a < E; ++a) {
                    double q = 0.0;
#pragma unroll
                    for (int b = 0; b < E; ++b) {
                        const double isj = SOA(w.iS, j, a, b);
                        q = fma(isj, MUS(j, b), q);
                        A[a][b] = SOA(w.iS, i, a, b) + isj;
                    }
                    mjiSmj = fma(MUS(j, a), q, mjiSmj);
                    mij[a] = iSmi[a] + q;
                }
                const double detA = inv_inplace<E>(A);
                double quad = 0.0;
#pragma unroll
                for (int a = 0; a < E; ++a) {
                    double q = 0.0;
#pragma unroll
                    for (int b = 0; b < E; ++b) q = fma(A[a][b], mij[b], q);
                    quad = fma(mij[a], q, quad);
                }
                const double ldSj = log(w.dS[(size_t)j * N + n]);
                const double phi = miiSmi + mjiSmj - quad + ldSi + ldSj + log(detA);
                T2 += 2.0 * pps[i] * pps[j] * exp(-0.5 * phi);
            }
        }
        out = ((0.5 * E) * log2pi - log(T2)) - T1;
    }
    return out;
}

template <int E>
__global__ void __launch_bounds__(128) crit_kernel(int kind, int N, int M, const double* __restrict__ noise,
                                                   const double* __restrict__ pps, CritWork w, double* __restrict__ dc) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    dc[n] = crit_eval<E>(kind, n, N, M, noise, pps, w);
}
#undef SOA
#undef MUS

size_t criterion_work_doubles(int N, int M, int E) {
    return (size_t)N * M * (2 * (size_t)E * E + 1 + E);
}

template <int E>
static int crit_dispatch(cudaStream_t s, int kind, const double* mu, const double* S, int N, int M,
                         const double* noise_dev, const double* pps_dev, double* dc, double* work) {
    CritWork w;
    w.Sn = work;
    w.iS = w.Sn + (size_t)N * M * E * E;
    w.dS = w.iS + (size_t)N * M * E * E;
    w.mu = w.dS + (size_t)N * M;
    const int need_inv = (kind == 1 || kind == 3 || kind == 4);
    long long total = (long long)N * M;
    crit_prep_kernel<E><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(mu, S, N, M, noise_dev, w, need_inv);
    GPD_CUDA(cudaGetLastError());
    crit_kernel<E><<<(N + 127) / 128, 128, 0, s>>>(kind, N, M, noise_dev, pps_dev, w, dc);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

int launch_criterion(cudaStream_t s, int kind, const double* mu, const double* S, int N, int M, int E,
                     const double* noise_dev, const double* pps_dev, double* dc, double* work) {
    switch (E) {
        case 1: return crit_dispatch<1>(s, kind, mu, S, N, M, noise_dev, pps_dev, dc, work);
        case 2: return crit_dispatch<2>(s, kind, mu, S, N, M, noise_dev, pps_dev, dc, work);
        case 3: return crit_dispatch<3>(s, kind, mu, S, N, M, noise_dev, pps_dev, dc, work);
        case 4: return crit_dispatch<4>(s, kind, mu, S, N, M, noise_dev, pps_dev, dc, work);
        case 5: return crit_dispatch<5>(s, kind, mu, S, N, M, noise_dev, pps_dev, dc, work);
        case 6: return crit_dispatch<6>(s, kind, mu, S, N, M, noise_dev, pps_dev, dc, work);
        case 7: return crit_dispatch<7>(s, kind, mu, S, N, M, noise_dev, pps_dev, dc, work);
        case 8: return crit_dispatch<8>(s, kind, mu, S, N, M, noise_dev, pps_dev, dc, work);
        default:
            set_error("criteria support 1 <= E <= 8 outputs");
            return 3;
    }
}

// ---------------------------------------------------------------------------------------------
// Gaussian log-density of observations under every rival model's marginal prediction
// (GPdoemd/discrimination_criteria.py:32-124: gaussian_posterior, gaussian_posterior_update, chi2, akaike all
// reduce to  log N(y_j; m_ji, S_ji)  and the squared Mahalanobis distance per (observation j, model i))
// ---------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(128) gauss_loglik_kernel(const double* __restrict__ Y, const double* __restrict__ Mu,
                                                           const double* __restrict__ S, long long n, int N,
                                                           double* __restrict__ logpdf, double* __restrict__ maha) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n * N) return;
    const long long j = tid / N;
    double A[E][E], d[E];
#pragma unroll
    for (int a = 0; a < E; ++a) {
        d[a] = Y[j * E + a] - Mu[tid * E + a];
#pragma unroll
        for (int b = 0; b < E; ++b) A[a][b] = S[(tid * E + a) * E + b];
    }
    const double det = inv_inplace<E>(A);
    double q = 0.0;
#pragma unroll
    for (int a = 0; a < E; ++a) {
        double t = 0.0;
#pragma unroll
        for (int b = 0; b < E; ++b) t = fma(A[a][b], d[b], t);
        q = fma(d[a], t, q);
    }
    maha[tid] = q;
    logpdf[tid] = -0.5 * (E * 1.8378770664093453 + log(det) + q);      // log(2 pi)
}

int launch_gauss_loglik(cudaStream_t s, const double* Y, const double* Mu, const double* S, long long n, int N, int E,
                        double* logpdf, double* maha) {
    const long long total = n * N;
    const unsigned blocks = (unsigned)((total + 127) / 128);
    switch (E) {
#define GPD_LL_CASE(EV) \
    case EV: gauss_loglik_kernel<EV><<<blocks, 128, 0, s>>>(Y, Mu, S, n, N, logpdf, maha); break;
        GPD_LL_CASE(1) GPD_LL_CASE(2) GPD_LL_CASE(3) GPD_LL_CASE(4) GPD_LL_CASE(5) GPD_LL_CASE(6) GPD_LL_CASE(7) GPD_LL_CASE(8)
#undef GPD_LL_CASE
        default:
            set_error("discrimination criteria support 1 <= E <= 8 outputs");
            return 3;
    }
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// argmax with numpy semantics, two stages, running (value, index) pair kept on the device
// ---------------------------------------------------------------------------------------------
struct Best { double v; long long i; };

__device__ __forceinline__ bool better(double a, long long ia, double b, long long ib) {
    // does (a, ia) beat (b, ib)?  NaN beats everything (np.argmax), ties go to the lower index
    const bool na = isnan(a), nb = isnan(b);
    if (na || nb) return na && (!nb || ia < ib);
    return a > b || (a == b && ia < ib);
}
__device__ __forceinline__ Best warp_best(Best x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double v = __shfl_xor_sync(0xffffffffu, x.v, o);
        long long i = __shfl_xor_sync(0xffffffffu, x.i, o);
        if (i >= 0 && (x.i < 0 || better(v, i, x.v, x.i))) { x.v = v; x.i = i; }
    }
    return x;
}
__device__ __forceinline__ Best block_best(Best x, Best* sh) {
    x = warp_best(x);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) sh[warp] = x;
    __syncthreads();
    if (warp == 0) {
        Best y = (lane < (int)(blockDim.x >> 5)) ? sh[lane] : Best{0.0, -1};
        y = warp_best(y);
        if (lane == 0) sh[0] = y;
    }
    __syncthreads();
    return sh[0];
}

constexpr int kArgmaxBlocks = 592;   // 4 x 148 SMs

__global__ void __launch_bounds__(256) argmax_stage1(const double* __restrict__ dc, long long N, long long idx_base,
                                                     Best* __restrict__ partial) {
    __shared__ Best sh[8];
    Best x{0.0, -1};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
        double v = dc[i];
        if (x.i < 0 || better(v, i + idx_base, x.v, x.i)) { x.v = v; x.i = i + idx_base; }
    }
    x = block_best(x, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = x;
}
__global__ void __launch_bounds__(256) argmax_stage2(const Best* __restrict__ partial, int nparts, int first,
                                                     double* __restrict__ best_val, long long* __restrict__ best_idx) {
    __shared__ Best sh[8];
    Best x{0.0, -1};
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
        Best p = partial[i];
        if (p.i >= 0 && (x.i < 0 || better(p.v, p.i, x.v, x.i))) x = p;
    }
    x = block_best(x, sh);
    if (threadIdx.x == 0) {
        if (!first) {
            Best cur{*best_val, *best_idx};
            if (cur.i >= 0 && (x.i < 0 || better(cur.v, cur.i, x.v, x.i))) x = cur;
        }
        *best_val = x.v;
        *best_idx = x.i;
    }
}

size_t argmax_partial_bytes() { return sizeof(Best) * kArgmaxBlocks * kMaxKinds; }

// ---------------------------------------------------------------------------------------------
// Fused first-order post-processing of the scoring path (one launch each):
//   post1_kernel      raw contractions -> units -> S = diag(s2) + J Sigma J^T (clipped) -> + noise -> inverse, det
//                     written straight into the criterion workspace (SoA); replaces M finalize + M assemble +
//                     crit_prep launches.  Same arithmetic order as finalize_kernel / assemble_kernel.
//   crit_multi_kernel every requested criterion for a candidate in one thread + per-block (max, index)
//   argmax_final      one block per criterion folds the block partials into the running best
// ---------------------------------------------------------------------------------------------
template <int E>
__global__ void __launch_bounds__(128) post1_kernel(const PostModel* __restrict__ pms, const RawOut* __restrict__ raws,
                                                    int M, int N, int chunk_n, const double* __restrict__ noise,
                                                    CritWork w, int need_inv) {
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= (long long)N * M) return;
    const int m = (int)(tid / N), n = (int)(tid % N);
    const PostModel pm = pms[m];
    const ModelXform* __restrict__ xf = pm.xf;
    const RawOut& raw = raws[m];
    const int P = pm.P;
    double J[E][kMaxP], s2[E];
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const size_t row = (size_t)e * chunk_n + n;
        const double* c = raw.cmu + row * pm.ldraw;
        const double ystd = xf->ystd[e];
        const double* gr = raw.gram + row * pm.npart;
        double vv = gr[0];
        for (int k = 1; k < pm.npart; ++k) vv += gr[k];
        w.mu[((size_t)m * E + e) * N + n] = xf->ymean[e] + c[0] * ystd;
        s2[e] = fmax(raw.kss[row] - vv, xf->var_floor) * (ystd * ystd);
        for (int a = 0; a < P; ++a) J[e][a] = c[1 + a] * ystd / xf->pdiff[a];
    }
    double A[E][E];
#pragma unroll
    for (int e1 = 0; e1 < E; ++e1)
#pragma unroll
        for (int e2 = 0; e2 < E; ++e2) {
            double v = 0.0;
            for (int a = 0; a < P; ++a) {
                double t = 0.0;
                for (int b = 0; b < P; ++b) t = fma(xf->Sigma[a * P + b], J[e2][b], t);
                v = fma(J[e1][a], t, v);
            }
            if (e1 == e2) v = fmax(1e-15, s2[e1] + v);
            A[e1][e2] = v + noise[e1 * E + e2];
            w.Sn[((size_t)(m * E + e1) * E + e2) * N + n] = A[e1][e2];
        }
    if (!need_inv) return;
    const double det = inv_inplace<E>(A);
    w.dS[(size_t)m * N + n] = det;
#pragma unroll
    for (int a = 0; a < E; ++a)
#pragma unroll
        for (int b = 0; b < E; ++b) w.iS[((size_t)(m * E + a) * E + b) * N + n] = A[a][b];
}

struct KindList { int n; int kind[kMaxKinds]; double* dc[kMaxKinds]; };

template <int E>
__global__ void __launch_bounds__(128) crit_multi_kernel(KindList kl, int N, int M, const double* __restrict__ noise,
                                                         const double* __restrict__ pps, CritWork w, long long idx_base,
                                                         Best* __restrict__ partial) {
    __shared__ Best sh[8];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    for (int k = 0; k < kl.n; ++k) {
        Best x{0.0, -1};
        if (n < N) {
            const double v = crit_eval<E>(kl.kind[k], n, N, M, noise, pps, w);
            if (kl.dc[k]) kl.dc[k][n] = v;
            x.v = v;
            x.i = idx_base + n;
        }
        x = block_best(x, sh);
        if (threadIdx.x == 0) partial[(size_t)k * gridDim.x + blockIdx.x] = x;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) argmax_final(const Best* __restrict__ partial, int nparts, int first,
                                                    double* __restrict__ best_val, long long* __restrict__ best_idx) {
    __shared__ Best sh[8];
    const int k = blockIdx.x;
    Best x{0.0, -1};
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
        Best p = partial[(size_t)k * nparts + i];
        if (p.i >= 0 && (x.i < 0 || better(p.v, p.i, x.v, x.i))) x = p;
    }
    x = block_best(x, sh);
    if (threadIdx.x == 0) {
        if (!first) {
            Best cur{best_val[k], best_idx[k]};
            if (cur.i >= 0 && (x.i < 0 || better(cur.v, cur.i, x.v, x.i))) x = cur;
        }
        best_val[k] = x.v;
        best_idx[k] = x.i;
    }
}

size_t crit_partial_bytes(int N, int nk) { return sizeof(Best) * (size_t)((N + 127) / 128) * nk; }

static CritWork make_work(double* work, int N, int M, int E) {
    CritWork w;
    w.Sn = work;
    w.iS = w.Sn + (size_t)N * M * E * E;
    w.dS = w.iS + (size_t)N * M * E * E;
    w.mu = w.dS + (size_t)N * M;
    return w;
}

template <int E>
static int post1_dispatch(cudaStream_t s, const PostModel* pms, const RawOut* raws, int M, int N, int chunk_n,
                          const double* noise, double* work, int need_inv) {
    long long total = (long long)N * M;
    post1_kernel<E><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(pms, raws, M, N, chunk_n, noise,
                                                                    make_work(work, N, M, E), need_inv);
    GPD_CUDA(cudaGetLastError());
    return 0;
}
template <int E>
static int crit_multi_dispatch(cudaStream_t s, const KindList& kl, int N, int M, const double* noise, const double* pps,
                               double* work, long long idx_base, void* partial) {
    crit_multi_kernel<E><<<(N + 127) / 128, 128, 0, s>>>(kl, N, M, noise, pps, make_work(work, N, M, E), idx_base,
                                                         (Best*)partial);
    GPD_CUDA(cudaGetLastError());
    return 0;
}
template <int E>
static int crit_prep_dispatch(cudaStream_t s, const double* mu, const double* S, int N, int M, const double* noise,
                              double* work, int need_inv) {
    long long total = (long long)N * M;
    crit_prep_kernel<E><<<(unsigned)((total + 127) / 128), 128, 0, s>>>(mu, S, N, M, noise, make_work(work, N, M, E), need_inv);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

#define GPD_E_SWITCH(FN, ...)                                   \
    switch (E) {                                                \
        case 1: return FN<1>(__VA_ARGS__);                      \
        case 2: return FN<2>(__VA_ARGS__);                      \
        case 3: return FN<3>(__VA_ARGS__);                      \
        case 4: return FN<4>(__VA_ARGS__);                      \
        case 5: return FN<5>(__VA_ARGS__);                      \
        case 6: return FN<6>(__VA_ARGS__);                      \
        case 7: return FN<7>(__VA_ARGS__);                      \
        case 8: return FN<8>(__VA_ARGS__);                      \
        default:                                                \
            set_error("criteria support 1 <= E <= 8 outputs"); \
            return 3;                                           \
    }

static int need_inverse(const int* kinds, int nk) {
    for (int k = 0; k < nk; ++k)
        if (kinds[k] == 1 || kinds[k] == 3 || kinds[k] == 4) return 1;
    return 0;
}

int launch_post1(cudaStream_t s, const PostModel* pms_dev, const RawOut* raws_dev, int M, int E, int N, int chunk_n,
                 const double* noise_dev, double* work, const int* kinds, int nk) {
    const int inv = need_inverse(kinds, nk);
    GPD_E_SWITCH(post1_dispatch, s, pms_dev, raws_dev, M, N, chunk_n, noise_dev, work, inv)
}

int launch_crit_prep(cudaStream_t s, const double* mu, const double* S, int N, int M, int E, const double* noise_dev,
                     double* work, const int* kinds, int nk) {
    const int inv = need_inverse(kinds, nk);
    GPD_E_SWITCH(crit_prep_dispatch, s, mu, S, N, M, noise_dev, work, inv)
}

int launch_crit_multi(cudaStream_t s, const int* kinds, int nk, double* const* dc, int N, int M, int E,
                      const double* noise_dev, const double* pps_dev, double* work, long long idx_base, void* partial_dev,
                      int first, double* best_val_dev, long long* best_idx_dev) {
    KindList kl;
    kl.n = nk;
    for (int k = 0; k < nk; ++k) {
        kl.kind[k] = kinds[k];
        kl.dc[k] = dc ? dc[k] : nullptr;
    }
    {
        auto go = [&]() -> int { GPD_E_SWITCH(crit_multi_dispatch, s, kl, N, M, noise_dev, pps_dev, work, idx_base, partial_dev) };
        GPD_TRY(go());
    }
    argmax_final<<<nk, 256, 0, s>>>((const Best*)partial_dev, (N + 127) / 128, first, best_val_dev, best_idx_dev);
    GPD_CUDA(cudaGetLastError());
    return 0;
}
#undef GPD_E_SWITCH


int launch_argmax(cudaStream_t s, const double* dc, long long N, long long idx_base, double* best_val_dev,
                  long long* best_idx_dev, int first, void* partial_dev) {
    long long want = (N + 255) / 256;
    int blocks = (int)(want < kArgmaxBlocks ? (want > 0 ? want : 1) : kArgmaxBlocks);
    argmax_stage1<<<blocks, 256, 0, s>>>(dc, N, idx_base, (Best*)partial_dev);
    GPD_CUDA(cudaGetLastError());
    argmax_stage2<<<1, 256, 0, s>>>((const Best*)partial_dev, blocks, first, best_val_dev, best_idx_dev);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// sharded scoring (gpd_score_sharded): every rank packs its running (max, index) pairs, one ncclAllGather moves
// world x nk x 16 bytes, and every rank folds the gathered pairs with np.argmax semantics (NaN wins, ties go to
// the lowest global index; an empty shard publishes index -1 and never wins)
// ---------------------------------------------------------------------------------------------
__global__ void best_pack_kernel(const double* __restrict__ bv, const long long* __restrict__ bi, int nk, int empty,
                                 Best* __restrict__ out) {
    const int k = threadIdx.x;
    if (k < nk) out[k] = empty ? Best{0.0, -1} : Best{bv[k], bi[k]};
}
__global__ void best_reduce_kernel(const Best* __restrict__ all, int world, int nk, double* __restrict__ bv,
                                   long long* __restrict__ bi) {
    const int k = threadIdx.x;
    if (k >= nk) return;
    Best x{0.0, -1};
    for (int r = 0; r < world; ++r) {
        const Best p = all[(size_t)r * nk + k];
        if (p.i >= 0 && (x.i < 0 || better(p.v, p.i, x.v, x.i))) x = p;
    }
    bv[k] = x.v;
    bi[k] = x.i;
}
size_t best_pair_bytes() { return sizeof(Best); }
// copy the singular-matrix flag to (pinned) host memory and clear it, in stream order
int launch_singular_fetch(cudaStream_t s, int* host_out) {
    int* flag_dev = nullptr;
    GPD_CUDA(cudaGetSymbolAddress((void**)&flag_dev, g_singular_flag));
    GPD_CUDA(cudaMemcpyAsync(host_out, flag_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
    GPD_CUDA(cudaMemsetAsync(flag_dev, 0, sizeof(int), s));
    return 0;
}
int launch_best_pack(cudaStream_t s, const double* bv, const long long* bi, int nk, int empty, void* out) {
    best_pack_kernel<<<1, 32, 0, s>>>(bv, bi, nk, empty, (Best*)out);
    GPD_CUDA(cudaGetLastError());
    return 0;
}
int launch_best_reduce(cudaStream_t s, const void* all, int world, int nk, double* bv, long long* bi) {
    best_reduce_kernel<<<1, 32, 0, s>>>((const Best*)all, world, nk, bv, bi);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// counter-based candidates and the L2 flush
// ---------------------------------------------------------------------------------------------
__host__ __device__ inline unsigned long long splitmix64(unsigned long long z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__global__ void fill_uniform_kernel(double* __restrict__ X, long long N, int D, unsigned long long seed,
                                    long long first_row, const double* __restrict__ lo, const double* __restrict__ hi) {
    long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= N * D) return;
    const long long row = tid / D;
    const int d = (int)(tid % D);
    const unsigned long long ctr = (unsigned long long)(first_row + row) * (unsigned long long)D + d;
    const unsigned long long bits = splitmix64(splitmix64(seed) ^ ctr);
    const double u = (double)(bits >> 11) * (1.0 / 9007199254740992.0);
    X[tid] = lo[d] + u * (hi[d] - lo[d]);
}
int launch_fill_uniform(cudaStream_t s, double* X, long long N, int D, uint64_t seed, long long first_row,
                        const double* lo_dev, const double* hi_dev) {
    long long total = N * D;
    fill_uniform_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(X, N, D, seed, first_row, lo_dev, hi_dev);
    GPD_CUDA(cudaGetLastError());
    return 0;
}
__global__ void flush_kernel(double* buf, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        buf[i] = buf[i] * 0.5 + 1.0;
}
int launch_flush(cudaStream_t s, double* buf, size_t n) {
    flush_kernel<<<1184, 256, 0, s>>>(buf, n);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace gpd
