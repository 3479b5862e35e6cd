/*
 * The design step of GPdoemd's demos/case_study.py:279-293 from plain C, through the C ABI alone
 * (include/gpdoemd_b200.h): what a host without Python -- or a cgo / JNI / N-API binding -- would call.
 *
 *   gcc -std=c99 -Iinclude examples/score_c_abi.c -Lgpdoemd_b200 -lgpdoemd_b200 -Wl,-rpath,$PWD/gpdoemd_b200 -lm -o score_c_abi
 *   ./score_c_abi [n_gpus]
 *
 * Two rival models with two GP surrogates each are trained on synthetic data ON THE DEVICE (gpd_gp_build: covariance
 * matrix, Cholesky factor and alpha; hyper-parameters are given, as after GPModel.gp_load_hyp, gp_model.py:128-143), a
 * grid of candidate designs is scored with the Box-Hill criterion (gpd_score_multi), and -- with n_gpus > 1 -- the same
 * grid is cut into contiguous blocks over several GPUs, one host thread and one context per GPU, with the NCCL
 * all-gather of the per-block winners inside gpd_score_sharded.  Every rank must report the same global argmax.
 */
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gpdoemd_b200.h"

#define CHECK(call)                                                                  \
    do {                                                                             \
        int rc__ = (call);                                                           \
        if (rc__) {                                                                  \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, gpd_last_error());  \
            exit(1);                                                                 \
        }                                                                            \
    } while (0)

enum { D = 2, P = 2, E = 2, M = 2, NTRAIN = 300, GRID = 200, N = GRID * GRID };

static double urand(unsigned long long* s) {      /* splitmix64 -> [0, 1) */
    unsigned long long z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return (double)((z ^ (z >> 31)) >> 11) / 9007199254740992.0;
}

/* one rival model on one device: E exact GPs on (x, p) in [0,1]^(D+P), standardised targets */
static gpd_model* build_model(gpd_ctx* ctx, int which, gpd_gp* gps[E]) {
    static const int32_t x_active[D] = {0, 1};
    double* Z = (double*)malloc(sizeof(double) * NTRAIN * (D + P));
    double* Y = (double*)malloc(sizeof(double) * NTRAIN);
    unsigned long long seed = 1234 + which;
    for (int i = 0; i < NTRAIN * (D + P); ++i) Z[i] = urand(&seed);
    for (int e = 0; e < E; ++e) {
        double mean = 0, var = 0;
        for (int i = 0; i < NTRAIN; ++i) {
            const double* z = Z + i * (D + P);
            Y[i] = sin(3.0 * z[0] + (1 + which) * z[2]) + (e ? z[1] * z[3] : cos(2.0 * z[1]));
            mean += Y[i] / NTRAIN;
        }
        for (int i = 0; i < NTRAIN; ++i) var += (Y[i] - mean) * (Y[i] - mean) / NTRAIN;
        for (int i = 0; i < NTRAIN; ++i) Y[i] = (Y[i] - mean) / sqrt(var);
        double ls_x[D] = {0.6 + 0.1 * e, 0.9}, ls_p[P] = {1.1, 0.7 + 0.2 * which};
        gpd_gp_desc d;
        memset(&d, 0, sizeof(d));
        d.kernel = GPD_KERN_MATERN52;
        d.n = NTRAIN;
        d.dim_x = D;
        d.dim_p = P;
        d.n_x_active = D;
        d.x_active = x_active;
        d.var_x = 1.3;
        d.ls_x = ls_x;
        d.var_p = 0.9;
        d.ls_p = ls_p;
        d.power_x = d.power_p = 2.0;
        d.Z = Z;
        d.factor_kind = GPD_FACTOR_CHOL;
        CHECK(gpd_gp_build(ctx, &d, Y, 1e-4, &gps[e], NULL, NULL, NULL));
    }
    free(Z);
    free(Y);
    const double zero[2] = {0, 0}, one[2] = {1, 1};            /* transforms: data already in [0,1], y standardised */
    gpd_model* mo = NULL;
    CHECK(gpd_model_create(ctx, E, D, P, 0, NULL, gps, zero, one, zero, one, zero, one, &mo));
    const double pmean[P] = {0.4 + 0.1 * which, 0.55};
    const double Sigma[P * P] = {0.010, 0.002, 0.002, 0.020};
    CHECK(gpd_model_set_pmean_sigma(mo, pmean, Sigma));
    return mo;
}

typedef struct {
    int rank, world, device;
    const char* nccl_id;
    const double* X;
    double best_val;
    int64_t best_idx;
} rank_args;

static void* rank_main(void* p) {
    rank_args* a = (rank_args*)p;
    gpd_ctx* ctx = NULL;
    CHECK(gpd_create(a->device, &ctx));
    gpd_gp* gps[M][E];
    gpd_model* models[M];
    for (int m = 0; m < M; ++m) models[m] = build_model(ctx, m, gps[m]);      /* GP state replicated on every GPU */
    const int32_t kinds[1] = {GPD_BH};
    const double noise[E * E] = {0.01, 0, 0, 0.01}, pps[M] = {0.5, 0.5};
    if (a->world == 1) {
        CHECK(gpd_score_multi(ctx, models, M, a->X, 0, N, 1, kinds, 1, noise, pps, NULL, 0, &a->best_val, &a->best_idx, 0));
    } else {
        gpd_comm* comm = NULL;
        CHECK(gpd_comm_init(ctx, a->nccl_id, a->rank, a->world, &comm));       /* collective over the rank threads */
        const int64_t per = (N + a->world - 1) / a->world, lo = a->rank * per;
        const int64_t n_local = lo >= N ? 0 : (lo + per > N ? N - lo : per);
        CHECK(gpd_score_sharded(ctx, comm, models, M, a->X + lo * D, 0, n_local, 1, kinds, 1, noise, pps, NULL, 0,
                                &a->best_val, &a->best_idx, lo));
        gpd_comm_destroy(comm);
    }
    for (int m = 0; m < M; ++m) {
        gpd_model_free(models[m]);
        for (int e = 0; e < E; ++e) gpd_gp_free(gps[m][e]);
    }
    gpd_destroy(ctx);
    return NULL;
}

int main(int argc, char** argv) {
    const int world = argc > 1 ? atoi(argv[1]) : 1;
    double* X = (double*)malloc(sizeof(double) * N * D);
    for (int i = 0; i < GRID; ++i)
        for (int j = 0; j < GRID; ++j) {
            X[(i * GRID + j) * D + 0] = i / (GRID - 1.0);
            X[(i * GRID + j) * D + 1] = j / (GRID - 1.0);
        }
    char id[GPD_NCCL_ID_BYTES];
    memset(id, 0, sizeof(id));
    if (world > 1) CHECK(gpd_nccl_unique_id(id));
    rank_args args[16];
    pthread_t th[16];
    for (int r = 0; r < world; ++r) {
        args[r].rank = r;
        args[r].world = world;
        args[r].device = r;
        args[r].nccl_id = id;
        args[r].X = X;
        pthread_create(&th[r], NULL, rank_main, &args[r]);
    }
    for (int r = 0; r < world; ++r) pthread_join(th[r], NULL);
    for (int r = 0; r < world; ++r) {
        printf("rank %d: next design x = (%.4f, %.4f), BH = %.9g, index %lld\n", r, X[args[r].best_idx * D],
               X[args[r].best_idx * D + 1], args[r].best_val, (long long)args[r].best_idx);
        if (args[r].best_idx != args[0].best_idx || args[r].best_val != args[0].best_val) {
            fprintf(stderr, "ranks disagree on the global argmax\n");
            return 1;
        }
    }
    free(X);
    return 0;
}
