// Host side of libgpdoemd_b200.so: the C ABI of include/gpdoemd_b200.h on top of the kernels.
//
// Every entry point mirrors one reference interface (file:line in the header).  Nothing is computed on
// the CPU here: the host code only plans chunks, builds descriptor tables and launches kernels.
//
// Data flow of one candidate chunk (Nc candidates, all on ctx->stream):
//   route (binary combos) -> eval (kx rows, kx.C contractions, RHS panels) -> tri (V = L^-1 R on DMMA,
//   ||v||^2) [-> gram, tri backward (beta), beta contraction for the variance derivatives]
//   -> finalize (original units) [-> assemble (Taylor marginal) -> criterion -> argmax]
#include <cuda_runtime.h>
#include <math.h>
#include <nvtx3/nvToolsExt.h>     // header-only NVTX: ranges for `ncu --nvtx --nvtx-include "gpd/..."`
#include <string.h>

#include <algorithm>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/gpdoemd_b200.h"
#include "common.cuh"

namespace gpd {
static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return 0;
        if (p) GPD_CUDA(cudaFree(p));
        p = nullptr;
        cap = 0;
        GPD_CUDA(cudaMalloc(&p, bytes));
        cap = bytes;
        return 0;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

enum Family { F_EVAL = 0, F_TRI = 1, F_GRAM = 2, F_MARG = 3, F_CRIT = 4, F_ARGMAX = 5, F_SETUP = 6, F_TRI_HEAD = 7, F_COMM = 8 };
}  // namespace gpd

using namespace gpd;

struct gpd_ctx {
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr;          // high priority: every kernel of the library
    cudaStream_t stream2 = nullptr;         // low priority: only the evaluation overlapped with a tail wave (run_chunk)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    int tail_overlap = 1;                   // option "tail_overlap"
    int redzones = 0, redzone_selftest = 0; // option "scratch_redzones" (debug)
    std::vector<size_t> zones;              // offsets of the red zones of the current scratch layout
    int head_periods = 0, head_grid = 0;    // options "head_periods" / "head_grid": see plan_materialize (tail overlap)
    int tri_parts_auto = 0, tri_parts = 0;  // option "tri_parts": 0 (default) / -1 = whole items, 9 = by factor size, 1..8 forced
    int fused = 0;                          // option "fused": first-order path through the fused evaluation + contraction kernel
    int setup_variant = 1;                  // option "setup_variant": 1 = per-GP setup on the tensor pipe (kernels_inv.cu), 0 = round-1 FMA kernels
    uint64_t scratch_limit = 6ull << 30;
    DevBuf scratch, tables, flush;
    DevBuf setup_a, setup_dense;           // n x n work matrices of gpd_gp_build / gpd_gp_upload, kept between GPs (a model has E x R of them)
    uint64_t launches = 0;
    int tri_variant = 2;   // 0: 16 warps as 4 x 4 (32 x 32 warp tiles), 1: 8 warps as 2 x 4, 2: 16 warps as 2 x 8 (64 x 16)
    int tri_cps = 0;   // 0 = automatic (by factor size), 1 or 2 = forced chunks per pipeline stage
    int tri_trim = 1;  // skip the all-padding 8-row blocks of the last row block
    // profiling: event pairs recorded without synchronising, resolved in gpd_profile_read
    int profile = 0;
    struct Span { cudaEvent_t a, b; int fam; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> pool;
    double fam_ms[GPD_NFAMILIES] = {0}, fam_flops[GPD_NFAMILIES] = {0};
    uint64_t fam_launches[GPD_NFAMILIES] = {0};
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    double* best_val_dev = nullptr;
    long long* best_idx_dev = nullptr;
    void* argmax_partial = nullptr;
    void* result_pinned = nullptr;          // page-locked landing zone of (best value, best index) per criterion
    // Plan and descriptor tables of the last fused scoring call, still valid in the scratch region: a design loop
    // calls gpd_score* with the same models / shapes over and over, and planning + table uploads + small pageable
    // copies cost ~0.1 ms per call (1.5 % of a C2 step).  Any other use of the scratch region, a freed handle or a
    // changed option invalidates it.
    struct ScoreState* score_state = nullptr;
    // device buffers of freed resident marginals, kept for the next gpd_marginals call (a design loop asks for the same
    // sizes over and over; cudaMalloc + cudaFree per call cost ~0.5 ms and cudaFree synchronises the device)
    std::vector<std::pair<void*, size_t>> res_pool;
    uint64_t epoch = 1;                     // bumped by everything that invalidates cached plans
};

struct gpd_gp {
    gpd_ctx* ctx = nullptr;
    int kernel = 0, n = 0, n_pad = 0, nblk = 0, D = 0, P = 0, nx = 0, factor_kind = 0;
    int x_active[kMaxXDims];
    double ls_x[kMaxXDims];
    double var_x = 1, var_p = 1, power_x = 2, power_p = 2;
    double *Xs = nullptr, *XsE = nullptr, *Zp = nullptr, *alpha = nullptr, *ls_p = nullptr;
    double *C = nullptr, *G = nullptr, *H = nullptr;
    double *LinvF = nullptr, *LinvB = nullptr;
};

struct gpd_model {
    gpd_ctx* ctx = nullptr;
    int E = 0, D = 0, P = 0, nb = 0, R = 1;
    int binary[kMaxBinary];
    std::vector<gpd_gp*> gps;  // E * R
    std::vector<double> xmin, xmax, pmin, pmax, ymean, ystd, pmean, Sigma;
    bool has_pmean = false, has_sigma = false;
    ModelXform* xf_dev = nullptr;      // [0] real unit conversions, [1] identity (transformed-unit hooks)
    double* pstar_dev = nullptr;
    std::vector<double> pstar;         // p* currently folded into the GPs' C/G/H tables
    int *bcols_dev = nullptr, *weights_dev = nullptr;
};

namespace gpd {

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
constexpr double kRbfCoordScaleHost = kRbfCoordScale;          // common.cuh

struct NvtxRange {                // RAII range, domain-less, name prefix "gpd/"
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
};

struct Span {
    gpd_ctx* c;
    int fam;
    cudaEvent_t a = nullptr, b = nullptr;
    cudaStream_t st;
    Span(gpd_ctx* ctx, int family, uint64_t nlaunch, double flops, cudaStream_t on = nullptr)
        : c(ctx), fam(family), st(on ? on : ctx->stream) {
        c->launches += nlaunch;
        if (!c->profile) return;
        c->fam_launches[fam] += nlaunch;
        c->fam_flops[fam] += flops;
        a = take();
        b = take();
        cudaEventRecord(a, st);
    }
    ~Span() {
        if (!a) return;
        cudaEventRecord(b, st);
        c->spans.push_back({a, b, fam});
    }
    cudaEvent_t take() {
        cudaEvent_t e;
        if (!c->pool.empty()) {
            e = c->pool.back();
            c->pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        return e;
    }
};

static int resolve_spans(gpd_ctx* c) {
    if (c->spans.empty()) return 0;
    GPD_CUDA(cudaStreamSynchronize(c->stream));
    for (auto& s : c->spans) {
        float ms = 0.f;
        GPD_CUDA(cudaEventElapsedTime(&ms, s.a, s.b));
        c->fam_ms[s.fam] += ms;
        c->pool.push_back(s.a);
        c->pool.push_back(s.b);
    }
    c->spans.clear();
    return 0;
}

// bump allocator over the context scratch region
// Option "scratch_redzones" (debug; compute-sanitizer is not available on this pool): every sub-buffer carved out of
// the scratch region is followed by a 256-byte red zone filled with 0xCB in stream order before the call's kernels run;
// gpd_debug_check_scratch counts the bytes that no longer hold the pattern.  The scratch region is where every
// N-sized write of the library lands (panels, raw contractions, hook outputs, criterion workspace, work-item lists).
constexpr size_t kRedZone = 256;
struct Bump {
    char* base;
    size_t off = 0, cap;
    gpd_ctx* ctx;
    Bump(void* p, size_t c, gpd_ctx* owner = nullptr) : base((char*)p), cap(c), ctx(owner) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* r = (T*)(base + off);
        off += count * sizeof(T);
        if (ctx && ctx->redzones) {
            off = align_up(off, 256);
            if (off + kRedZone <= cap) {
                cudaMemsetAsync(base + off, 0xCB, kRedZone, ctx->stream);
                ctx->zones.push_back(off);
            }
            off += kRedZone;
        }
        return r;
    }
};

// ---- per-call plan ---------------------------------------------------------------------------
enum HookFlags { WANT_DMU = 1, WANT_D2MU = 2, WANT_DS2 = 4, WANT_D2S2 = 8 };

struct GpEntry {
    gpd_gp* gp;
    int model, e, r;
    long long panel_off;
};
struct Group {          // GPs evaluated by one eval launch: same kernel class, active dims count and P
    int kernel, nx, P;
    std::vector<int> entries;
    int item_begin = 0, item_count = 0;
    int head_begin = 0, head_count = 0, big_begin = 0, big_count = 0;   // in Plan::items2_dev (tail overlap)
    int fitem_begin = 0, fitem_count = 0;                               // in Plan::fitems_dev (fused path)
};
struct ModelBufs {
    RawOut raw;
    int ldraw = 0, ng = 0, ldmu = 0;
    double *mu = nullptr, *s2 = nullptr, *dmu = nullptr, *d2mu = nullptr, *ds2 = nullptr, *d2s2 = nullptr;
    int* idx = nullptr;
    int* counts = nullptr;
};
struct Plan {
    std::vector<gpd_model*> models;
    int flags = 0;
    bool transformed = false;       // X, outputs in the reference's transformed units (the `_`-hooks)
    bool need_W = false;            // variance derivatives: 1+P right-hand sides, V panels kept
    bool skip_finalize = false;     // the fused scoring path converts units inside post1_kernel
    int chunk = 0, tiles = 0;
    std::vector<GpEntry> entries;
    std::vector<Group> groups;
    std::vector<ModelBufs> mb;
    size_t per_cand_bytes = 0;      // scratch bytes per candidate (panels + raws + hook outputs)
    long long panel_unit = 0;       // sum over GPs of n_pad (doubles per candidate per RHS kind)
    int max_nrhs = 1;
    // device tables
    GpView* views_dev = nullptr;
    Item* items_dev = nullptr;
    Item* items2_dev = nullptr;     // the same items as [head tiles of every GP][remaining tiles of every GP]
    int head_tiles = 0, n_head = 0, n_big = 0;
    RawOut* raws_dev = nullptr;
    double *panels_in = nullptr, *panels_out = nullptr, *panels_beta = nullptr;
    int nitems = 0;
    Bump* bump = nullptr;
    // first-order contraction units: every (GP, tile) item cut into tri_parts row-block parts of equal triangular work
    // (large factors: the parts of a tile run on neighbouring CTAs and share its right-hand-side panel through L2)
    Item* titems_dev = nullptr;
    int tri_parts = 1, n_head_units = 0, n_units = 0;
    // fused evaluation + contraction (first order): no per-candidate panels, one private panel per CTA instead
    bool fused = false;
    FItem* fitems_dev = nullptr;
    double* priv = nullptr;
    long long priv_stride = 0;
    int fused_parts = 1;            // max parts an item is cut into (tail wave); ||v||^2 slots = 2 * fused_parts
    size_t fixed_bytes = 0;         // scratch that does not scale with the chunk (private panels, fused item list)
};

}  // namespace gpd
struct ScoreState {
    gpd::Plan pl;
    uint64_t epoch = 0;
    std::vector<gpd_model*> models;
    std::vector<int> kinds;
    std::vector<double> noise, pps;
    long long N = 0;
    int order = 0, x_on_device = 0, keep_dc = 0, dc_on_device = 0;
    void* scratch_p = nullptr;
    size_t scratch_cap = 0;
    double *X_stage = nullptr, *mu_all = nullptr, *S_all = nullptr, *dc_dev = nullptr, *work = nullptr, *tmpA = nullptr,
           *tmp_s2 = nullptr, *d_noise = nullptr, *d_pps = nullptr;
    void* partial = nullptr;
    gpd::PostModel* pms_dev = nullptr;
};
namespace gpd {

static int ncols_for(int flags, int P) {
    if (flags & WANT_D2MU) return 1 + P + P * (P + 1) / 2;
    if (flags & WANT_DMU) return 1 + P;
    return 1;
}

static int plan_build(gpd_ctx* c, gpd_model* const* models, int M, int flags, long long N, size_t extra_per_cand,
                      size_t extra_fixed, Plan& pl) {
    pl.models.assign(models, models + M);
    pl.flags = flags;
    pl.need_W = (flags & (WANT_DS2 | WANT_D2S2)) != 0;
    pl.entries.clear();
    pl.groups.clear();
    pl.panel_unit = 0;
    pl.max_nrhs = 1;
    size_t per_cand = extra_per_cand;
    std::map<std::tuple<int, int, int>, int> gidx;
    for (int m = 0; m < M; ++m) {
        gpd_model* mo = models[m];
        GPD_CHECK(mo && mo->ctx == c, "model does not belong to this context");
        GPD_CHECK(mo->has_pmean, "pmean is not set (surrogate_model.py:171)");
        const int P = mo->P, E = mo->E, npair = P * (P + 1) / 2;
        const int nrhs = pl.need_W ? 1 + P : 1;
        pl.max_nrhs = std::max(pl.max_nrhs, nrhs);
        for (int e = 0; e < E; ++e)
            for (int r = 0; r < mo->R; ++r) {
                gpd_gp* g = mo->gps[(size_t)e * mo->R + r];
                if (!g) continue;
                auto key = std::make_tuple(g->kernel, g->nx, P);
                auto it = gidx.find(key);
                int gi;
                if (it == gidx.end()) {
                    gi = (int)pl.groups.size();
                    gidx[key] = gi;
                    pl.groups.push_back(Group{g->kernel, g->nx, P, {}, 0, 0});
                } else {
                    gi = it->second;
                }
                pl.groups[gi].entries.push_back((int)pl.entries.size());
                pl.entries.push_back(GpEntry{g, m, e, r, 0});
                pl.panel_unit += g->n_pad;
            }
        // raws + hook outputs, bytes per candidate
        const int ncols = ncols_for(flags, P);
        const int ng = pl.need_W ? nrhs * (nrhs + 1) / 2 : 32;     // first order: up to 4 row groups x 8 parts
        size_t d = (size_t)E * (ncols + 1 + ng + ((flags & WANT_D2S2) ? npair : 0));
        d += (size_t)E * 2;
        if (flags & WANT_DMU) d += (size_t)E * P;
        if (flags & WANT_D2MU) d += (size_t)E * P * P;
        if (flags & WANT_DS2) d += (size_t)E * P;
        if (flags & WANT_D2S2) d += (size_t)E * P * P;
        per_cand += d * sizeof(double);
        if (mo->nb) per_cand += sizeof(int) * mo->R;
    }
    GPD_CHECK(!pl.entries.empty(), "no GP surrogates in the models");
    // first order (one right-hand side, mean + first derivative columns only): the fused kernel keeps the panel of the
    // item in flight in a private L2-resident region per CTA, nothing panel-sized is allocated per candidate
    long long n_pad_max = 0;
    pl.fused = c->fused && !pl.need_W && !(flags & WANT_D2MU);
    for (auto& grp : pl.groups)
        if (ncols_for(flags, grp.P) > 11 || grp.nx > kMaxXDims) pl.fused = false;
    for (auto& en : pl.entries) n_pad_max = std::max<long long>(n_pad_max, en.gp->n_pad);
    pl.priv_stride = n_pad_max * kFusedTile;      // per CTA: n_pad x tile candidates
    // Row-block parts of the (unfused) first-order contraction (option "tri_parts", off by default).  A tile's
    // right-hand-side panel is 8 n_pad x 128 bytes and row block I re-reads its first I + 1 k blocks: at n_pad = 4096
    // that is 68 MB per item out of a 4 MB panel, all of it from HBM (148 private panels = 592 MB in flight).  Cutting
    // every item into parts on neighbouring CTAs was meant to share the panel through L2; measured on B200 it is 9 %
    // SLOWER (C5: 337 -> 370 ms, C4: 540 -> 592 ms): between two uses of a k block the whole chip streams ~1 GB
    // through L2, so the panel is evicted either way, and the factor chunks lose their 148-way sharing
    // (profiles/r02_summary.md).  The mechanism stays because the fused kernel cuts its tail-wave items the same way.
    {
        int nblk_min = 1 << 30;
        for (auto& en : pl.entries) nblk_min = std::min(nblk_min, en.gp->nblk);
        pl.tri_parts = (!pl.need_W && !pl.fused && c->tri_parts_auto) ? std::max(1, std::min(8, nblk_min / 4)) : 1;
        if (c->tri_parts > 0 && !pl.need_W && !pl.fused) pl.tri_parts = std::max(1, std::min(std::min(8, nblk_min), c->tri_parts));
    }
    // panels: in (max_nrhs kinds), out (only when W is needed), beta (only for d2s2)
    int kinds = pl.max_nrhs + (pl.need_W ? pl.max_nrhs : 0) + ((flags & WANT_D2S2) ? 1 : 0);
    pl.fixed_bytes = 0;
    if (pl.fused) {
        pl.fixed_bytes = sizeof(double) * (size_t)pl.priv_stride * (c->fused == 2 ? 2 : kFusedCtasPerSm) * c->num_sms + sizeof(FItem) * 16 * (size_t)c->num_sms + (1 << 16);
        extra_fixed += pl.fixed_bytes;
        per_cand += (2 * sizeof(FItem) * pl.entries.size() + kBlk - 1) / kBlk; // one Item per (GP, tile)
    } else per_cand += (size_t)pl.panel_unit * kinds * sizeof(double);
    pl.per_cand_bytes = per_cand;
    const size_t fixed = extra_fixed + (size_t)(1 << 20) +
                         pl.entries.size() * (sizeof(GpView) + 64) +
                         (size_t)kBlk * per_cand;   // tile rounding slack
    GPD_CHECK(c->scratch_limit > fixed + (size_t)kBlk * per_cand,
              "scratch limit too small for one 128-candidate tile of these models; raise gpd_set_scratch_bytes");
    long long chunk = (long long)((c->scratch_limit - fixed) / per_cand);
    chunk = chunk / kBlk * kBlk;
    chunk = std::min<long long>(chunk, 1 << 22);
    GPD_CHECK(chunk >= kBlk, "scratch limit too small");
    const long long want = (N + kBlk - 1) / kBlk * kBlk;
    if (want <= chunk) {
        chunk = want;
    } else {
        // Several chunks.  The triangular kernel is persistent over num_sms CTAs and every work item costs the same,
        // so a chunk of t tiles takes ceil(t * GPs * RHS / num_sms) item-times forward (+ ceil(t * GPs / num_sms)
        // for the backward pass of the second variance derivative).  Pick the tile count in the upper half of the
        // capacity that minimises the total over all chunks including the remainder chunk.
        const long long T = want / kBlk, tmax = chunk / kBlk, tmin = std::max<long long>(1, tmax / 2);
        const long long S = c->num_sms, G = (long long)pl.entries.size(), R = pl.max_nrhs;
        const bool bwd = (flags & WANT_D2S2) != 0;
        auto cost = [&](long long t) -> double {
            if (t <= 0) return 0.0;
            double v = (double)((t * G * R + S - 1) / S);
            if (bwd) v += (double)((t * G + S - 1) / S);
            return v + 0.05;                       // per-chunk launch / pipeline fill overhead
        };
        double best = 1e300;
        long long best_t = tmax;
        for (long long t = tmax; t >= tmin; --t) {
            const double total = (double)(T / t) * cost(t) + cost(T % t);
            if (total < best - 1e-9) {
                best = total;
                best_t = t;
            }
        }
        chunk = best_t * kBlk;
    }
    pl.chunk = (int)chunk;
    pl.tiles = pl.chunk / kBlk;
    return 0;
}

// carve the scratch region and upload the static descriptor tables for this plan
// cut the row blocks [0, nblk) of a tail item into s contiguous parts of (nearly) equal cost; the cost of part [a, b) is
// its triangular tile count sum_{I=a}^{b-1} (I + 1) plus kE * b for re-evaluating the k rows it needs -> s + 1 bounds
static std::vector<int> split_rows(int nblk, int s, double kE) {
    s = std::max(1, std::min(s, nblk));
    auto cost = [&](int a, int b) { return 0.5 * ((double)b * (b + 1) - (double)a * (a + 1)) + kE * b; };
    std::vector<std::vector<double>> best(s + 1, std::vector<double>(nblk + 1, 1e300));
    std::vector<std::vector<int>> from(s + 1, std::vector<int>(nblk + 1, 0));
    best[0][0] = 0.0;
    for (int k = 1; k <= s; ++k)
        for (int b = k; b <= nblk; ++b)
            for (int a = k - 1; a < b; ++a) {
                const double v = std::max(best[k - 1][a], cost(a, b));
                if (v < best[k][b]) { best[k][b] = v; from[k][b] = a; }
            }
    std::vector<int> bounds(s + 1);
    bounds[s] = nblk;
    for (int k = s; k >= 1; --k) bounds[k - 1] = from[k][bounds[k]];
    return bounds;
}

// items of the fused kernel for one group: whole (GP, tile) items for the complete waves of the persistent grid, the
// remainder (the partial last wave, or everything when there is less than one wave) cut into row-block parts
static void fused_items_for(const gpd_ctx* c, const Plan& pl, const Group& grp, std::vector<FItem>& out, int* max_parts) {
    const int S = kFusedCtasPerSm * c->num_sms;
    std::vector<std::pair<int, int>> all;     // (entry, tile of kFusedTile candidates), tile-major: one wave touches every factor
    for (int t = 0; t < kFusedCtasPerSm * pl.tiles; ++t)
        for (int ei : grp.entries) all.push_back({ei, t});
    const int nit = (int)all.size(), n_full = nit / S * S, rem = nit - n_full;
    int s = rem > 0 ? std::min(4, S / rem) : 1;
    if (!c->tail_overlap) s = 1;              // the same switch disables the tail treatment of both paths
    for (int k = 0; k < nit; ++k) {
        const gpd_gp* g = pl.entries[all[k].first].gp;
        if (k < n_full || s < 2 || g->nblk < 2) {
            out.push_back(FItem{all[k].first, all[k].second, 0, (short)g->nblk, 0, 1});
            continue;
        }
        const std::vector<int> b = split_rows(g->nblk, s, 0.2);
        const int np = (int)b.size() - 1;
        *max_parts = std::max(*max_parts, np);
        for (int p = 0; p < np; ++p) out.push_back(FItem{all[k].first, all[k].second, (short)b[p], (short)b[p + 1], (short)p, (short)np});
    }
}

static int plan_materialize(gpd_ctx* c, Plan& pl, Bump& bump) {
    const int M = (int)pl.models.size();
    const int Nc = pl.chunk;
    std::vector<FItem> fitems;
    pl.fused_parts = 1;
    if (pl.fused)
        for (auto& grp : pl.groups) {
            grp.fitem_begin = (int)fitems.size();
            fused_items_for(c, pl, grp, fitems, &pl.fused_parts);
            grp.fitem_count = (int)fitems.size() - grp.fitem_begin;
        }
    pl.mb.assign(M, ModelBufs());
    std::vector<RawOut> raws(M);
    for (int m = 0; m < M; ++m) {
        gpd_model* mo = pl.models[m];
        const int P = mo->P, E = mo->E, npair = P * (P + 1) / 2;
        const int nrhs = pl.need_W ? 1 + P : 1;
        ModelBufs& b = pl.mb[m];
        b.ldraw = ncols_for(pl.flags, P);
        b.ng = pl.need_W ? nrhs * (nrhs + 1) / 2
                         : (pl.fused ? 2 * pl.fused_parts : tri_row_groups(c->tri_variant) * pl.tri_parts);
        b.raw.cmu = bump.take<double>((size_t)E * Nc * b.ldraw);
        b.raw.kss = bump.take<double>((size_t)E * Nc);
        b.raw.gram = bump.take<double>((size_t)E * Nc * b.ng);
        b.raw.bh = (pl.flags & WANT_D2S2) ? bump.take<double>((size_t)E * Nc * npair) : nullptr;
        b.mu = bump.take<double>((size_t)Nc * E);
        b.ldmu = E;
        b.s2 = bump.take<double>((size_t)Nc * E);
        if (pl.flags & WANT_DMU) b.dmu = bump.take<double>((size_t)Nc * E * P);
        if (pl.flags & WANT_D2MU) b.d2mu = bump.take<double>((size_t)Nc * E * P * P);
        if (pl.flags & WANT_DS2) b.ds2 = bump.take<double>((size_t)Nc * E * P);
        if (pl.flags & WANT_D2S2) b.d2s2 = bump.take<double>((size_t)Nc * E * P * P);
        if (mo->nb) {
            b.idx = bump.take<int>((size_t)mo->R * Nc);
            b.counts = bump.take<int>(mo->R);
        }
        raws[m] = b.raw;
    }
    const size_t panel_doubles = pl.fused ? 0 : (size_t)pl.panel_unit * Nc;
    pl.priv = pl.fused ? bump.take<double>((size_t)pl.priv_stride * (c->fused == 2 ? 2 : kFusedCtasPerSm) * c->num_sms) : nullptr;
    pl.fitems_dev = pl.fused ? bump.take<FItem>(fitems.size()) : nullptr;
    pl.panels_in = bump.take<double>(panel_doubles * pl.max_nrhs);
    pl.panels_out = pl.need_W ? bump.take<double>(panel_doubles * pl.max_nrhs) : nullptr;
    pl.panels_beta = (pl.flags & WANT_D2S2) ? bump.take<double>(panel_doubles) : nullptr;

    // views, grouped so that each eval launch sees a contiguous item range
    std::vector<GpView> views(pl.entries.size());
    std::vector<Item> items;
    long long off = 0, off_w = 0;
    for (auto& grp : pl.groups) {
        grp.item_begin = (int)items.size();
        for (int ei : grp.entries) {
            GpEntry& en = pl.entries[ei];
            gpd_gp* g = en.gp;
            gpd_model* mo = pl.models[en.model];
            en.panel_off = off;
            off += (long long)pl.tiles * g->n_pad * kBlk;
            const long long kinds_g = pl.need_W ? 1 + g->P : 1;     // right-hand-side kinds of this GP's model
            const long long my_off_w = off_w;
            off_w += (long long)pl.tiles * g->n_pad * kBlk * kinds_g;
            GpView& v = views[ei];
            memset(&v, 0, sizeof(v));
            v.kernel = g->kernel;
            v.n = g->n;
            v.n_pad = g->n_pad;
            v.nblk = g->nblk;
            v.nx = g->nx;
            v.P = g->P;
            v.kss = g->var_x * g->var_p;
            v.power_x = g->power_x;
            v.Xs = g->Xs;
            v.XsE = g->XsE;
            v.C = g->C;
            v.G = g->G;
            v.H = g->H;
            v.LinvF = g->LinvF;
            v.LinvB = g->LinvB;
            for (int d = 0; d < g->nx; ++d) {
                v.ls[d] = g->ls_x[d];
                v.x_active[d] = g->x_active[d];
                v.xmin[d] = pl.transformed ? 0.0 : mo->xmin[g->x_active[d]];
                v.xdiff[d] = pl.transformed ? 1.0 : mo->xmax[g->x_active[d]] - mo->xmin[g->x_active[d]];
                v.xscale[d] = 1.0 / (v.xdiff[d] * g->ls_x[d]) * (g->kernel == GPD_KERN_RBF ? kRbfCoordScaleHost : 1.0);
            }
            const ModelBufs& b = pl.mb[en.model];
            v.idx = mo->nb ? b.idx + (size_t)en.r * Nc : nullptr;
            v.count = mo->nb ? b.counts + en.r : nullptr;
            v.out_e = en.e;
            v.model = en.model;
            v.panel_off = en.panel_off;
            v.panel_off_w = my_off_w;
            for (int t = 0; t < pl.tiles; ++t) items.push_back(Item{ei, t});
        }
        grp.item_count = (int)items.size() - grp.item_begin;
    }
    pl.nitems = (int)items.size();
    // Tail overlap (first-order path): the persistent triangular kernel runs ceil(items / SMs) waves and the last,
    // partial one leaves most SMs idle for a whole item time (C2: 6256 items = 42.27 waves).  The first `head_tiles`
    // candidate tiles of every GP are therefore split off such that the REST fills complete waves; the head's
    // triangular launch (few CTAs) then runs concurrently with the evaluation of the rest on a second stream.
    std::vector<Item> items2;
    pl.head_tiles = 0;
    pl.n_head = pl.n_big = 0;
    if (c->tail_overlap && !pl.need_W && !pl.fused) {
        const long long S = c->num_sms, G = (long long)pl.entries.size() * pl.tri_parts;   // contraction units per tile
        long long a = S, b = G;
        while (b) { const long long t = a % b; a = b; b = t; }      // a = gcd(S, G)
        const long long period = S / a;                              // tiles per GP that make whole waves
        const long long h = pl.tiles % period;
        if (h > 0 && (long long)pl.tiles * G > S && h < pl.tiles) pl.head_tiles = (int)h;
        // option "head_periods": a LONGER head (+ k whole-wave periods of tiles) whose triangular launch, restricted to
        // "head_grid" CTAs, keeps part of the SMs on the tensor pipe for the whole evaluation of the rest (which is bound
        // by its panel write and tolerates fewer SMs)
        if (c->head_periods > 0 && (long long)pl.tiles * G > S) {
            const long long want = h + (long long)c->head_periods * period;
            if (want > 0 && want < pl.tiles) pl.head_tiles = (int)want;
        }
    }
    // contraction units of the first-order path: the items (head tiles first when the tail overlap is on) cut into parts
    std::vector<Item> titems;
    auto expand = [&](const std::vector<Item>& src) {
        for (const Item& it : src) {
            const gpd_gp* g = pl.entries[it.gp].gp;
            if (pl.tri_parts <= 1) { titems.push_back(it); continue; }
            const std::vector<int> bnd = split_rows(g->nblk, pl.tri_parts, 0.0);
            const int np = (int)bnd.size() - 1;
            for (int p = 0; p < np; ++p)
                titems.push_back(Item{it.gp, it.tile, (short)bnd[p], (short)bnd[p + 1], (short)p, (short)np});
        }
    };
    if (pl.head_tiles) {
        for (int pass = 0; pass < 2; ++pass)
            for (auto& grp : pl.groups) {
                const int begin = (int)items2.size();
                for (int ei : grp.entries)
                    for (int t = pass ? pl.head_tiles : 0; t < (pass ? pl.tiles : pl.head_tiles); ++t) items2.push_back(Item{ei, t});
                (pass ? grp.big_begin : grp.head_begin) = begin;
                (pass ? grp.big_count : grp.head_count) = (int)items2.size() - begin;
            }
        pl.n_head = (int)pl.entries.size() * pl.head_tiles;
        pl.n_big = (int)items2.size() - pl.n_head;
    }
    int n_head_units = 0;
    if (!pl.need_W && !pl.fused) {
        if (pl.head_tiles) {
            expand(std::vector<Item>(items2.begin(), items2.begin() + pl.n_head));
            n_head_units = (int)titems.size();
            expand(std::vector<Item>(items2.begin() + pl.n_head, items2.end()));
        } else {
            expand(items);
        }
    }
    pl.n_head_units = n_head_units;
    pl.n_units = (int)titems.size();
    pl.views_dev = bump.take<GpView>(views.size());
    pl.items_dev = bump.take<Item>(items.size());
    pl.items2_dev = items2.empty() ? nullptr : bump.take<Item>(items2.size());
    pl.titems_dev = titems.empty() ? nullptr : bump.take<Item>(titems.size());
    pl.raws_dev = bump.take<RawOut>(M);
    GPD_CHECK(bump.off <= bump.cap, "internal: scratch plan overflow");
    GPD_CUDA(cudaMemcpyAsync(pl.views_dev, views.data(), sizeof(GpView) * views.size(), cudaMemcpyHostToDevice, c->stream));
    GPD_CUDA(cudaMemcpyAsync(pl.items_dev, items.data(), sizeof(Item) * items.size(), cudaMemcpyHostToDevice, c->stream));
    if (pl.items2_dev)
        GPD_CUDA(cudaMemcpyAsync(pl.items2_dev, items2.data(), sizeof(Item) * items2.size(), cudaMemcpyHostToDevice, c->stream));
    if (pl.titems_dev)
        GPD_CUDA(cudaMemcpyAsync(pl.titems_dev, titems.data(), sizeof(Item) * titems.size(), cudaMemcpyHostToDevice, c->stream));
    GPD_CUDA(cudaMemcpyAsync(pl.raws_dev, raws.data(), sizeof(RawOut) * M, cudaMemcpyHostToDevice, c->stream));
    if (pl.fused)
        GPD_CUDA(cudaMemcpyAsync(pl.fitems_dev, fitems.data(), sizeof(FItem) * fitems.size(), cudaMemcpyHostToDevice, c->stream));
    // pageable-source cudaMemcpyAsync returns once the data is staged: the host vectors may go out of scope
    return 0;
}

// Pipeline-stage depth of tri_kernel.  Measured on B200 with the 2 x 8 warp layout (profiles/r01_summary.md): 3 stages of
// 32 k beat 4 stages of 16 k at n_pad = 512 (C2: 5.91 vs 6.00 ms) and at n_pad = 4096 (C5: 0.967 vs 0.956 of the peak);
// with the earlier 4 x 4 layout the large factors had preferred 4 x 16 k.
static int tri_cps_for(const gpd_ctx* c, const Plan& pl) {
    (void)pl;
    return c->tri_cps ? c->tri_cps : 2;
}

// run the GP part for one chunk of nc candidates (device X, row stride ldx); results in pl.mb[m]
// X_host != nullptr: the candidates of this chunk are still on the host (row-major, ldx doubles per row) and are
// copied into X_dev here, so that the copy of everything but the head tiles can ride on the second stream beside
// the head's evaluation and triangular launch.
static int run_chunk(gpd_ctx* c, Plan& pl, double* X_dev, int ldx, int nc, const double* X_host = nullptr) {
    cudaStream_t s = c->stream;
    const int M = (int)pl.models.size();
    const int Nc = pl.chunk;
    NvtxRange nvtx_chunk("gpd/run_chunk");
    bool any_binary = false;
    for (int m = 0; m < M; ++m) any_binary = any_binary || pl.models[m]->nb > 0;
    const bool overlap = !pl.fused && pl.head_tiles > 0 && !pl.need_W && nc > pl.head_tiles * kBlk;
    // rows the main stream needs before its first kernel: all of them when a routing pass reads X, else the head tiles
    const size_t head_rows = (X_host && overlap && !any_binary) ? (size_t)pl.head_tiles * kBlk : (size_t)nc;
    if (X_host)
        GPD_CUDA(cudaMemcpyAsync(X_dev, X_host, sizeof(double) * head_rows * ldx, cudaMemcpyHostToDevice, s));
    const int tiles_used = (nc + kBlk - 1) / kBlk;
    // routing + zero fill for models with binary variables (rows without a GP keep 0: gp_model.py:190-196)
    for (int m = 0; m < M; ++m) {
        gpd_model* mo = pl.models[m];
        if (!mo->nb) continue;
        ModelBufs& b = pl.mb[m];
        Span sp(c, F_SETUP, 1, 0);
        GPD_TRY(launch_route(s, X_dev, ldx, nc, mo->nb, mo->bcols_dev, mo->weights_dev, mo->R, b.idx, b.counts, Nc));
        GPD_CUDA(cudaMemsetAsync(b.raw.cmu, 0, sizeof(double) * (size_t)mo->E * Nc * b.ldraw, s));
        GPD_CUDA(cudaMemsetAsync(b.raw.kss, 0, sizeof(double) * (size_t)mo->E * Nc, s));
        GPD_CUDA(cudaMemsetAsync(b.raw.gram, 0, sizeof(double) * (size_t)mo->E * Nc * b.ng, s));
        if (b.raw.bh)
            GPD_CUDA(cudaMemsetAsync(b.raw.bh, 0, sizeof(double) * (size_t)mo->E * Nc * (mo->P * (mo->P + 1) / 2), s));
    }
    // The item list holds pl.tiles tiles per GP; tiles beyond the chunk's candidates exit on the device
    // count check (identity routing uses chunk_n = nc).
    (void)tiles_used;
    double n2sum = 0;   // algorithmic flops: n^2 per candidate per right-hand side
    if (pl.fused) {
        // K1 + K2 in one persistent launch per group; the panel of an item lives in the CTA's private L2-resident region
        if (pl.fused_parts > 1)     // slots of the items that are not cut into parts must read as zero
            for (int m = 0; m < M; ++m)
                if (!pl.models[m]->nb)
                    GPD_CUDA(cudaMemsetAsync(pl.mb[m].raw.gram, 0, sizeof(double) * (size_t)pl.models[m]->E * Nc * pl.mb[m].ng, s));
        for (auto& grp : pl.groups) {
            const int ncols = ncols_for(pl.flags, grp.P);
            double fl = 0;
            for (int ei : grp.entries) {
                const gpd_gp* g = pl.entries[ei].gp;
                fl += (double)nc * (double)g->n * g->n / pl.models[pl.entries[ei].model]->R;
            }
            FusedArgs fa;
            fa.gps = pl.views_dev;
            fa.items = pl.fitems_dev + grp.fitem_begin;
            fa.nitems = grp.fitem_count;
            fa.X = X_dev;
            fa.ldx = ldx;
            fa.chunk_n = nc;
            fa.priv = pl.priv;
            fa.priv_stride = pl.priv_stride;
            fa.raws = pl.raws_dev;
            fa.ldraw = pl.mb[pl.entries[grp.entries[0]].model].ldraw;
            fa.ncols = ncols;
            fa.ng = 2 * pl.fused_parts;
            fa.trim = c->tri_trim;
            const int grid = std::min(grp.fitem_count, (c->fused == 2 ? 1 : kFusedCtasPerSm) * c->num_sms);
            Span sp(c, F_TRI, 1, fl);
            const int rc = c->fused == 2 ? launch_fusedc(s, grp.kernel, grid, grp.nx, ncols, fa)
                                         : launch_fused(s, grp.kernel, grid, grp.nx, ncols, fa);
            GPD_CHECK(rc >= 0, "internal: no fused kernel variant for a shape the planner accepted");
            if (rc) return rc;
        }
    } else
    if (overlap) {
        GPD_CUDA(cudaEventRecord(c->ev_fork, s));
        GPD_CUDA(cudaStreamWaitEvent(c->stream2, c->ev_fork, 0));
        if (X_host && head_rows < (size_t)nc)
            GPD_CUDA(cudaMemcpyAsync(X_dev + head_rows * ldx, X_host + head_rows * ldx,
                                     sizeof(double) * ((size_t)nc - head_rows) * ldx, cudaMemcpyHostToDevice, c->stream2));
    }
    for (int pass = 0; pass < (pl.fused ? 0 : (overlap ? 2 : 1)); ++pass) {
        // pass 0: all items (or the head tiles) on the main, high-priority stream; pass 1: the remaining tiles on the
        // second stream, concurrently with the head's evaluation and triangular launch
        cudaStream_t es = pass ? c->stream2 : s;
        for (auto& grp : pl.groups) {
            const int nrhs = pl.need_W ? 1 + grp.P : 1;
            const int ncols = ncols_for(pl.flags, grp.P);
            const Item* its = !overlap ? pl.items_dev + grp.item_begin : pl.items2_dev + (pass ? grp.big_begin : grp.head_begin);
            const int nits = !overlap ? grp.item_count : (pass ? grp.big_count : grp.head_count);
            const double part = (double)nits / std::max(1, grp.item_count);
            double evflops = 0;
            for (int ei : grp.entries) {
                const gpd_gp* g = pl.entries[ei].gp;
                const double share = part / pl.models[pl.entries[ei].model]->R;
                evflops += share * (double)nc * g->n * (3.0 * g->nx + 45.0 + 2.0 * ncols + nrhs);
                n2sum += share * (double)nc * (double)g->n * g->n * nrhs;
            }
            Span sp(c, F_EVAL, 1, evflops, es);
            GPD_TRY(launch_eval(es, pl.views_dev, its, nits, X_dev, ldx, nc, nrhs, ncols, grp.kernel, grp.nx, pl.panels_in,
                                pl.raws_dev, pl.mb[pl.entries[grp.entries[0]].model].ldraw));
        }
        if (pass) GPD_CUDA(cudaEventRecord(c->ev_join, c->stream2));
    }
    if (overlap) {
        const double fh = (double)pl.n_head / (pl.n_head + pl.n_big);
        {
            Span sp(c, F_TRI_HEAD, 1, n2sum * fh);
            GPD_TRY(launch_tri(s, pl.views_dev, pl.titems_dev, pl.n_head_units, 1, 1, 1, 0, pl.panels_in, nullptr, pl.raws_dev,
                               tri_row_groups(c->tri_variant) * pl.tri_parts, nc,
                               c->head_grid > 0 ? std::min(c->head_grid, c->num_sms) : c->num_sms, c->tri_variant,
                               tri_cps_for(c, pl), c->tri_trim));
        }
        GPD_CUDA(cudaStreamWaitEvent(s, c->ev_join, 0));
        Span sp(c, F_TRI, 1, n2sum * (1.0 - fh));
        GPD_TRY(launch_tri(s, pl.views_dev, pl.titems_dev + pl.n_head_units, pl.n_units - pl.n_head_units, 1, 1, 1, 0, pl.panels_in,
                           nullptr, pl.raws_dev, tri_row_groups(c->tri_variant) * pl.tri_parts, nc, c->num_sms, c->tri_variant,
                           tri_cps_for(c, pl), c->tri_trim));
    } else
    if (pl.fused) {
        // evaluation and contraction already ran in the fused launch above
    } else if (!pl.need_W) {
        Span sp(c, F_TRI, 1, n2sum);
        GPD_TRY(launch_tri(s, pl.views_dev, pl.titems_dev, pl.n_units, 1, 1, 1, 0, pl.panels_in, nullptr, pl.raws_dev,
                           tri_row_groups(c->tri_variant) * pl.tri_parts, nc, c->num_sms, c->tri_variant, tri_cps_for(c, pl),
                           c->tri_trim));
    } else {
        for (auto& grp : pl.groups) {
            const int nrhs = 1 + grp.P;
            double fl = 0, flb = 0;
            for (int ei : grp.entries) {
                const gpd_gp* g = pl.entries[ei].gp;
                const double share = 1.0 / pl.models[pl.entries[ei].model]->R;
                fl += share * (double)nc * (double)g->n * g->n * nrhs;
                flb += share * (double)nc * (double)g->n * g->n;
            }
            {
                Span sp(c, F_TRI, 1, fl);
                GPD_TRY(launch_tri(s, pl.views_dev, pl.items_dev + grp.item_begin, grp.item_count, nrhs, nrhs, nrhs, 0,
                                   pl.panels_in, pl.panels_out, nullptr, 0, nc, c->num_sms, c->tri_variant, tri_cps_for(c, pl),
                                   c->tri_trim));
            }
            {
                Span sp(c, F_GRAM, 1, fl / (double)std::max(1, pl.entries[grp.entries[0]].gp->n) * (nrhs + 1));
                GPD_TRY(launch_gram(s, pl.views_dev, pl.items_dev + grp.item_begin, grp.item_count, nrhs, pl.panels_out,
                                    pl.raws_dev, nc));
            }
            if (pl.flags & WANT_D2S2) {
                {
                    Span sp(c, F_TRI, 1, flb);
                    GPD_TRY(launch_tri(s, pl.views_dev, pl.items_dev + grp.item_begin, grp.item_count, 1, nrhs, 1, 1,
                                       pl.panels_out, pl.panels_beta, nullptr, 0, nc, c->num_sms, c->tri_variant, tri_cps_for(c, pl),
                                       c->tri_trim));
                }
                Span sp(c, F_EVAL, 1, 0);
                GPD_TRY(launch_beta_contract(s, pl.views_dev, pl.items_dev + grp.item_begin, grp.item_count, X_dev, ldx, nc,
                                             grp.kernel, grp.nx, grp.P, pl.panels_beta, pl.raws_dev));
            }
        }
    }
    for (int m = 0; m < M && !pl.skip_finalize; ++m) {
        gpd_model* mo = pl.models[m];
        ModelBufs& b = pl.mb[m];
        Span sp(c, F_MARG, 1, 0);
        GPD_TRY(launch_finalize(s, mo->xf_dev + (pl.transformed ? 1 : 0), pl.raws_dev, m, mo->E, mo->P, nc, b.ldraw, pl.flags,
                                pl.need_W ? 1 + mo->P : 1, pl.need_W ? 0 : b.ng, b.mu, b.ldmu,
                                b.s2, b.dmu, b.d2mu, b.ds2, b.d2s2));
    }
    return 0;
}

static int ensure_scratch(gpd_ctx* c, size_t bytes) {
    GPD_CHECK(bytes <= c->scratch_limit + (64u << 20), "internal: plan exceeds the scratch limit");
    c->epoch++;                       // the caller is about to overwrite the scratch region
    c->zones.clear();
    return c->scratch.ensure(bytes + (c->redzones ? (1u << 20) : 0));
}

static int upload_model_state(gpd_model* mo);

// singular-matrix flag of the E x E inversions (kernels_crit.cu): fetched in stream order before the final
// synchronisation of an entry point, judged after it
static int* singular_slot(gpd_ctx* c) { return (int*)((char*)c->result_pinned + kMaxKinds * (sizeof(double) + sizeof(long long))); }
static int singular_fetch(gpd_ctx* c) { return launch_singular_fetch(c->stream, singular_slot(c)); }
static int singular_rc(gpd_ctx* c) {
    if (*singular_slot(c) == 0) return 0;
    *singular_slot(c) = 0;
    set_error("Singular matrix (np.linalg.inv of a covariance in design_criteria.py / scipy logpdf in "
              "discrimination_criteria.py raises LinAlgError here)");
    return GPD_ERR_SINGULAR;
}

}  // namespace gpd

// =================================================================================================
// context
// =================================================================================================
extern "C" {

const char* gpd_last_error(void) { return g_error.c_str(); }

int gpd_create(int device, gpd_ctx** out) {
    GPD_CHECK(out, "null out pointer");
    *out = nullptr;
    int count = 0;
    GPD_CUDA(cudaGetDeviceCount(&count));
    GPD_CHECK(count > 0 && device >= 0 && device < count, "no such CUDA device (this library has no CPU fallback)");
    GPD_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    GPD_CUDA(cudaGetDeviceProperties(&prop, device));
    GPD_CHECK(prop.major == 10, "libgpdoemd_b200 is built for sm_100a (Blackwell B200) only");
    gpd_ctx* c = new gpd_ctx();
    c->device = device;
    c->num_sms = prop.multiProcessorCount;
    int prio_least = 0, prio_greatest = 0;
    GPD_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
    GPD_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_greatest));
    GPD_CUDA(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prio_least));
    GPD_CUDA(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    GPD_CUDA(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
    GPD_CUDA(cudaEventCreate(&c->t0));
    GPD_CUDA(cudaEventCreate(&c->t1));
    GPD_CUDA(cudaMalloc(&c->best_val_dev, sizeof(double) * kMaxKinds));
    GPD_CUDA(cudaMalloc(&c->best_idx_dev, sizeof(long long) * kMaxKinds));
    GPD_CUDA(cudaMalloc(&c->argmax_partial, argmax_partial_bytes()));
    GPD_CUDA(cudaHostAlloc(&c->result_pinned, kMaxKinds * (sizeof(double) + sizeof(long long)) + 16, cudaHostAllocDefault));
    *out = c;
    return 0;
}

void gpd_destroy(gpd_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->scratch.release();
    c->tables.release();
    c->flush.release();
    c->setup_a.release();
    c->setup_dense.release();
    for (auto& e : c->res_pool) cudaFree(e.first);
    for (auto& s : c->spans) {
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    for (auto e : c->pool) cudaEventDestroy(e);
    cudaEventDestroy(c->t0);
    cudaEventDestroy(c->t1);
    cudaFree(c->best_val_dev);
    cudaFree(c->best_idx_dev);
    cudaFree(c->argmax_partial);
    cudaFreeHost(c->result_pinned);
    delete c->score_state;
    cudaEventDestroy(c->ev_fork);
    cudaEventDestroy(c->ev_join);
    cudaStreamDestroy(c->stream2);
    cudaStreamDestroy(c->stream);
    delete c;
}

int gpd_set_scratch_bytes(gpd_ctx* c, uint64_t bytes) {
    GPD_CHECK(c, "null context");
    GPD_CHECK(bytes >= (64ull << 20), "scratch must be at least 64 MiB");
    GPD_CUDA(cudaSetDevice(c->device));
    c->epoch++;
    c->scratch_limit = bytes;
    if (c->scratch.cap > bytes) {
        GPD_CUDA(cudaStreamSynchronize(c->stream));
        c->scratch.release();
    }
    return 0;
}

int gpd_set_option(gpd_ctx* c, const char* name, int64_t value) {
    GPD_CHECK(c && name, "null argument");
    c->epoch++;
    if (!strcmp(name, "tri_variant")) {
        GPD_CHECK(value >= 0 && value <= 2, "tri_variant is 0 (16 warps, 4x4), 1 (8 warps, 2x4) or 2 (16 warps, 2x8)");
        c->tri_variant = (int)value;
        return 0;
    }
    if (!strcmp(name, "tri_cps")) {
        GPD_CHECK(value >= 0 && value <= 2, "tri_cps is 0 (automatic), 1 or 2");
        c->tri_cps = (int)value;
        return 0;
    }
    if (!strcmp(name, "tail_overlap")) {
        GPD_CHECK(value == 0 || value == 1, "tail_overlap is 0 or 1");
        c->tail_overlap = (int)value;
        return 0;
    }
    if (!strcmp(name, "scratch_redzones")) {
        GPD_CHECK(value >= 0 && value <= 2, "scratch_redzones is 0, 1, or 2 (self-test: the next check first damages one byte)");
        if (value == 2) {
            c->redzone_selftest = 1;
            return 0;
        }
        c->redzones = (int)value;
        c->zones.clear();
        return 0;
    }
    if (!strcmp(name, "head_periods")) {
        GPD_CHECK(value >= 0 && value <= 16, "head_periods is 0..16");
        c->head_periods = (int)value;
        return 0;
    }
    if (!strcmp(name, "head_grid")) {
        GPD_CHECK(value >= 0 && value <= 1024, "head_grid is 0 (all SMs) or a CTA count");
        c->head_grid = (int)value;
        return 0;
    }
    if (!strcmp(name, "tri_parts")) {
        GPD_CHECK(value >= -1 && value <= 9, "tri_parts is 0 / -1 (whole items), 1..8, or 9 (by factor size)");
        c->tri_parts_auto = value == 9;
        c->tri_parts = (value > 0 && value < 9) ? (int)value : 0;
        return 0;
    }
    if (!strcmp(name, "fused")) {
        GPD_CHECK(value >= 0 && value <= 2, "fused is 0, 1 (phase-serial fused kernel) or 2 (producer warps inside the contraction CTA)");
        c->fused = (int)value;
        return 0;
    }
    if (!strcmp(name, "setup_variant")) {
        GPD_CHECK(value == 0 || value == 1, "setup_variant is 0 (FP64 FMA setup kernels) or 1 (tensor-pipe setup kernels)");
        c->setup_variant = (int)value;
        return 0;
    }
    if (!strcmp(name, "tri_trim")) {
        GPD_CHECK(value == 0 || value == 1, "tri_trim is 0 or 1");
        c->tri_trim = (int)value;
        return 0;
    }
    set_error(std::string("unknown option ") + name);
    return 2;
}

uint64_t gpd_launch_count(gpd_ctx* c) { return c ? c->launches : 0; }

int gpd_profile_enable(gpd_ctx* c, int on) {
    GPD_CHECK(c, "null context");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_TRY(resolve_spans(c));
    c->profile = on;
    for (int i = 0; i < GPD_NFAMILIES; ++i) {
        c->fam_ms[i] = 0;
        c->fam_flops[i] = 0;
        c->fam_launches[i] = 0;
    }
    return 0;
}

int gpd_profile_read(gpd_ctx* c, double ms_out[GPD_NFAMILIES], uint64_t launches_out[GPD_NFAMILIES],
                     double flops_out[GPD_NFAMILIES]) {
    GPD_CHECK(c, "null context");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_TRY(resolve_spans(c));
    for (int i = 0; i < GPD_NFAMILIES; ++i) {
        if (ms_out) ms_out[i] = c->fam_ms[i];
        if (launches_out) launches_out[i] = c->fam_launches[i];
        if (flops_out) flops_out[i] = c->fam_flops[i];
    }
    return 0;
}

int gpd_timer_start(gpd_ctx* c) {
    GPD_CHECK(c, "null context");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_CUDA(cudaEventRecord(c->t0, c->stream));
    return 0;
}
int gpd_timer_stop(gpd_ctx* c, double* ms_out) {
    GPD_CHECK(c && ms_out, "null argument");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_CUDA(cudaEventRecord(c->t1, c->stream));
    GPD_CUDA(cudaEventSynchronize(c->t1));
    float ms = 0.f;
    GPD_CUDA(cudaEventElapsedTime(&ms, c->t0, c->t1));
    *ms_out = ms;
    return 0;
}
int gpd_sync(gpd_ctx* c) {
    GPD_CHECK(c, "null context");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int gpd_probe_fp64_peak(gpd_ctx* c, double* tflops_out) {
    GPD_CHECK(c && tflops_out, "null argument");
    GPD_CUDA(cudaSetDevice(c->device));
    const int iters = 4096;
    double* sink = nullptr;
    GPD_CUDA(cudaMalloc(&sink, sizeof(double) * c->num_sms * 2 * 256));
    cudaEvent_t a, b;
    GPD_CUDA(cudaEventCreate(&a));
    GPD_CUDA(cudaEventCreate(&b));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        GPD_CUDA(cudaEventRecord(a, c->stream));
        GPD_TRY(launch_dmma_probe(c->stream, c->num_sms, iters, sink));
        c->launches += 1;
        GPD_CUDA(cudaEventRecord(b, c->stream));
        GPD_CUDA(cudaEventSynchronize(b));
        float ms = 0.f;
        GPD_CUDA(cudaEventElapsedTime(&ms, a, b));
        // per warp and iteration: 8 independent m8n8k4 DMMA = 8 * 512 flops
        const double flops = (double)c->num_sms * 2 * 8 /*warps*/ * (double)iters * 8 * 512;
        if (rep) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(sink);
    *tflops_out = best;
    return 0;
}

int gpd_flush_l2(gpd_ctx* c) {
    GPD_CHECK(c, "null context");
    GPD_CUDA(cudaSetDevice(c->device));
    const size_t n = (256ull << 20) / sizeof(double);   // 256 MiB > 126 MB L2
    if (!c->flush.p) {
        GPD_TRY(c->flush.ensure(n * sizeof(double)));
        GPD_CUDA(cudaMemsetAsync(c->flush.p, 0, n * sizeof(double), c->stream));
    }
    GPD_TRY(launch_flush(c->stream, (double*)c->flush.p, n));
    return 0;
}

// ---- device buffer helpers ----------------------------------------------------------------------
int gpd_dev_alloc(gpd_ctx* c, uint64_t bytes, void** out_dev) {
    GPD_CHECK(c && out_dev, "null argument");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_CUDA(cudaMalloc(out_dev, bytes ? bytes : 8));
    return 0;
}
int gpd_dev_free(gpd_ctx* c, void* dev) {
    GPD_CHECK(c, "null context");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_CUDA(cudaStreamSynchronize(c->stream));
    GPD_CUDA(cudaFree(dev));
    return 0;
}
int gpd_dev_upload(gpd_ctx* c, void* dst_dev, const void* src_host, uint64_t bytes) {
    GPD_CHECK(c, "null context");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, c->stream));
    GPD_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int gpd_dev_download(gpd_ctx* c, void* dst_host, const void* src_dev, uint64_t bytes) {
    GPD_CHECK(c, "null context");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, c->stream));
    GPD_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int gpd_host_alloc_pinned(gpd_ctx* c, uint64_t bytes, void** out_host) {
    GPD_CHECK(c && out_host, "null argument");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_CUDA(cudaHostAlloc(out_host, bytes ? bytes : 8, cudaHostAllocDefault));
    return 0;
}
int gpd_host_free_pinned(gpd_ctx* c, void* host) {
    GPD_CHECK(c, "null context");
    GPD_CUDA(cudaFreeHost(host));
    return 0;
}
int gpd_dev_fill_uniform(gpd_ctx* c, double* X_dev, int64_t N, int32_t D, uint64_t seed, int64_t first_row,
                         const double* lo, const double* hi) {
    GPD_CHECK(c && X_dev && lo && hi, "null argument");
    GPD_CHECK(D >= 1 && D <= 64 && N >= 0, "bad shape");
    GPD_CUDA(cudaSetDevice(c->device));
    if (N == 0) return 0;
    GPD_TRY(c->tables.ensure(sizeof(double) * 128));
    double* lo_dev = (double*)c->tables.p;
    double* hi_dev = lo_dev + 64;
    GPD_CUDA(cudaMemcpyAsync(lo_dev, lo, sizeof(double) * D, cudaMemcpyHostToDevice, c->stream));
    GPD_CUDA(cudaMemcpyAsync(hi_dev, hi, sizeof(double) * D, cudaMemcpyHostToDevice, c->stream));
    Span sp(c, F_SETUP, 1, 0);
    GPD_TRY(launch_fill_uniform(c->stream, X_dev, N, D, seed, first_row, lo_dev, hi_dev));
    return 0;
}

// =================================================================================================
// GP state
// =================================================================================================
void gpd_gp_free(gpd_gp* g) {
    if (!g) return;
    g->ctx->epoch++;
    cudaSetDevice(g->ctx->device);
    cudaStreamSynchronize(g->ctx->stream);
    cudaFree(g->Xs);
    cudaFree(g->XsE);
    cudaFree(g->Zp);
    cudaFree(g->alpha);
    cudaFree(g->ls_p);
    cudaFree(g->C);
    cudaFree(g->G);
    cudaFree(g->H);
    cudaFree(g->LinvF);
    cudaFree(g->LinvB);
    delete g;
}

}  // extern "C"

namespace gpd {
// Shared part of gpd_gp_upload / gpd_gp_build: validation, allocation, upload of the scaled training inputs.
// alpha and the packed factor are filled in by the caller.
static int gp_create_common(gpd_ctx* c, const gpd_gp_desc* d, gpd_gp** out) {
    *out = nullptr;
    GPD_CHECK(d->kernel >= 0 && d->kernel <= 5, "unknown kernel id");
    GPD_CHECK(d->n >= 1, "empty training set");
    GPD_CHECK(d->dim_p >= 1 && d->dim_p <= kMaxP, "dim_p out of range (1..16)");
    GPD_CHECK(d->n_x_active >= 1 && d->n_x_active <= kMaxXDims, "x-kernel acts on 1..8 non-binary design dimensions");
    GPD_CHECK(d->dim_x >= d->n_x_active, "dim_x smaller than the number of active dimensions");
    GPD_CHECK(d->x_active && d->ls_x && d->ls_p && d->Z, "null array in descriptor");
    GPD_CUDA(cudaSetDevice(c->device));
    gpd_gp* g = new gpd_gp();
    g->ctx = c;
    g->kernel = d->kernel;
    g->n = d->n;
    g->n_pad = (d->n + kBlk - 1) / kBlk * kBlk;
    g->nblk = g->n_pad / kBlk;
    g->D = d->dim_x;
    g->P = d->dim_p;
    g->nx = d->n_x_active;
    g->var_x = d->var_x;
    g->var_p = d->var_p;
    g->power_x = d->power_x;
    g->power_p = d->power_p;
    const int n = g->n, n_pad = g->n_pad, P = g->P, nx = g->nx, dz = g->D + g->P;
    for (int k = 0; k < nx; ++k) {
        g->x_active[k] = d->x_active[k];
        g->ls_x[k] = d->ls_x[k];
        if (!(d->x_active[k] >= 0 && d->x_active[k] < g->D) || !(d->ls_x[k] > 0)) {
            delete g;
            set_error("bad active dimension or non-positive lengthscale");
            return 2;
        }
    }
    for (int a = 0; a < P; ++a)
        if (!(d->ls_p[a] > 0)) {
            delete g;
            set_error("non-positive parameter lengthscale");
            return 2;
        }
    const int npair = P * (P + 1) / 2;
    std::vector<double> Xs((size_t)n_pad * nx, 0.0), Zp((size_t)n * P);
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < nx; ++k) {
            double v = d->Z[(size_t)i * dz + d->x_active[k]] / d->ls_x[k];
            if (d->kernel == GPD_KERN_RBF) v *= kRbfCoordScaleHost;      // see eval_impl.cuh: the squared distance in table units
            Xs[(size_t)i * nx + k] = v;
        }
        for (int a = 0; a < P; ++a) Zp[(size_t)i * P + a] = d->Z[(size_t)i * dz + g->D + a];
    }
    // RBF: the evaluation kernel forms the squared distance in expanded form from {-2 Xs, |Xs|^2} per training row
    std::vector<double> XsE;
    if (d->kernel == GPD_KERN_RBF) {
        XsE.assign((size_t)n_pad * (nx + 1), 0.0);
        for (int i = 0; i < n; ++i) {
            double sq = 0.0;
            for (int k = 0; k < nx; ++k) {
                const double v = Xs[(size_t)i * nx + k];
                XsE[(size_t)i * (nx + 1) + k] = -2.0 * v;
                sq = fma(v, v, sq);
            }
            XsE[(size_t)i * (nx + 1) + nx] = sq;
        }
    }
    cudaStream_t s = c->stream;
    const size_t pack_bytes = sizeof(double) * (size_t)packed_tiles(g->nblk) * kTileElems;
    cudaError_t e = cudaMalloc(&g->Xs, sizeof(double) * Xs.size());
    if (e == cudaSuccess && !XsE.empty()) e = cudaMalloc(&g->XsE, sizeof(double) * XsE.size());
    if (e == cudaSuccess && !XsE.empty()) e = cudaMemcpyAsync(g->XsE, XsE.data(), sizeof(double) * XsE.size(), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess) e = cudaMalloc(&g->Zp, sizeof(double) * Zp.size());
    if (e == cudaSuccess) e = cudaMalloc(&g->alpha, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMalloc(&g->ls_p, sizeof(double) * P);
    if (e == cudaSuccess) e = cudaMalloc(&g->C, sizeof(double) * (size_t)(1 + P + npair) * n_pad);
    if (e == cudaSuccess) e = cudaMalloc(&g->G, sizeof(double) * (size_t)(1 + P) * n_pad);
    if (e == cudaSuccess) e = cudaMalloc(&g->H, sizeof(double) * (size_t)std::max(1, npair) * n_pad);
    if (e == cudaSuccess) e = cudaMalloc(&g->LinvF, pack_bytes);
    if (e == cudaSuccess) e = cudaMalloc(&g->LinvB, pack_bytes);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g->Xs, Xs.data(), sizeof(double) * Xs.size(), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g->Zp, Zp.data(), sizeof(double) * Zp.size(), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g->ls_p, d->ls_p, sizeof(double) * P, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);      // Xs / Zp are host temporaries
    if (e != cudaSuccess) {
        set_error(std::string("GP state allocation / upload failed: ") + cudaGetErrorString(e));
        gpd_gp_free(g);
        return 1;
    }
    *out = g;
    return 0;
}

// L_dev (n x n row-major, lower triangle) -> packed forward / backward operands of the triangular kernel
static int gp_pack_factor(gpd_ctx* c, gpd_gp* g, const double* L_dev, int factor_kind) {
    if (c->setup_dense.ensure(sizeof(double) * (size_t)g->n_pad * g->n_pad)) {
        set_error("out of device memory for the factor inversion workspace");
        return 1;
    }
    double* dense = static_cast<double*>(c->setup_dense.p);
    uint64_t nl = 0;
    int rc = launch_invert_and_pack(c->stream, L_dev, g->n, g->n_pad, dense, g->LinvF, g->LinvB, &nl, factor_kind,
                                    c->setup_variant);
    c->launches += nl;
    if (c->profile) c->fam_launches[F_SETUP] += nl;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (rc) return rc;
    if (e != cudaSuccess) {
        set_error(std::string("factor inversion failed: ") + cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
}  // namespace gpd

extern "C" {

int gpd_gp_upload(gpd_ctx* c, const gpd_gp_desc* d, gpd_gp** out) {
    GPD_CHECK(c && d && out, "null argument");
    *out = nullptr;
    GPD_CHECK(d->factor_kind == GPD_FACTOR_CHOL || d->factor_kind == GPD_FACTOR_ROOT, "unknown factor_kind");
    GPD_CHECK(d->alpha && d->L, "null array in descriptor");
    gpd_gp* g = nullptr;
    GPD_TRY(gp_create_common(c, d, &g));
    g->factor_kind = d->factor_kind;
    const int n = g->n;
    cudaStream_t s = c->stream;
    cudaError_t e = c->setup_a.ensure(sizeof(double) * (size_t)n * n) ? cudaErrorMemoryAllocation : cudaSuccess;
    double* L_dev = static_cast<double*>(c->setup_a.p);
    if (e == cudaSuccess) e = cudaMemcpyAsync(g->alpha, d->alpha, sizeof(double) * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(L_dev, d->L, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice, s);
    int rc = 0;
    if (e == cudaSuccess) rc = gp_pack_factor(c, g, L_dev, d->factor_kind);
    else set_error(std::string("factor upload failed: ") + cudaGetErrorString(e));
    cudaStreamSynchronize(s);
    if (e != cudaSuccess || rc) {
        gpd_gp_free(g);
        return rc ? rc : 1;
    }
    *out = g;
    return 0;
}

int gpd_gp_build(gpd_ctx* c, const gpd_gp_desc* d, const double* Y, double noise_var, gpd_gp** out, double* alpha_out,
                 double* L_out, double* jitter_out) {
    GPD_CHECK(c && d && Y && out, "null argument");
    *out = nullptr;
    GPD_CHECK(d->factor_kind == GPD_FACTOR_CHOL, "gpd_gp_build forms the Cholesky factor of an exact GP");
    GPD_CHECK(noise_var >= 0.0, "negative noise variance");
    NvtxRange nvtx_build("gpd/gp_build");
    gpd_gp* g = nullptr;
    GPD_TRY(gp_create_common(c, d, &g));
    const int n = g->n, dz = g->D + g->P;
    cudaStream_t s = c->stream;
    double *A = nullptr, *Z_dev = nullptr, *y_dev = nullptr;
    int* info_dev = nullptr;
    int rc = 0;
    double jitter = 0.0;
    cudaError_t e = c->setup_a.ensure(sizeof(double) * (size_t)n * n) ? cudaErrorMemoryAllocation : cudaSuccess;
    A = static_cast<double*>(c->setup_a.p);
    if (e == cudaSuccess) e = cudaMalloc(&Z_dev, sizeof(double) * (size_t)n * dz);
    if (e == cudaSuccess) e = cudaMalloc(&y_dev, sizeof(double) * n);
    if (e == cudaSuccess) e = cudaMalloc(&info_dev, sizeof(int));
    if (e == cudaSuccess) e = cudaMemcpyAsync(Z_dev, d->Z, sizeof(double) * (size_t)n * dz, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(y_dev, Y, sizeof(double) * n, cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        KmatArgs ka;
        memset(&ka, 0, sizeof(ka));
        ka.kernel = g->kernel;
        ka.n = n;
        ka.dz = dz;
        ka.D = g->D;
        ka.P = g->P;
        ka.nx = g->nx;
        for (int k = 0; k < g->nx; ++k) {
            ka.x_active[k] = g->x_active[k];
            ka.ls_x[k] = g->ls_x[k];
        }
        for (int a = 0; a < g->P; ++a) ka.ls_p[a] = d->ls_p[a];
        ka.var_x = g->var_x;
        ka.var_p = g->var_p;
        ka.power_x = g->power_x;
        ka.power_p = g->power_p;
        // GPy jitchol: plain factorisation first, then up to 5 retries with
        // jitter = mean(diag) * 1e-6, * 10 per retry
        const double diag_mean = g->var_x * g->var_p + noise_var + 1e-8;
        for (int attempt = 0; attempt <= 5 && !rc && e == cudaSuccess; ++attempt) {
            jitter = attempt == 0 ? 0.0 : diag_mean * 1e-6 * pow(10.0, attempt - 1);
            ka.diag_add = noise_var + 1e-8 + jitter;
            const int init = n + 1;
            e = cudaMemcpyAsync(info_dev, &init, sizeof(int), cudaMemcpyHostToDevice, s);
            if (e != cudaSuccess) break;
            uint64_t nl = 1;
            rc = launch_kmat(s, ka, Z_dev, A);
            if (!rc) rc = launch_cholesky(s, A, n, info_dev, &nl, c->setup_variant);
            c->launches += nl;
            if (c->profile) c->fam_launches[F_SETUP] += nl;
            if (rc) break;
            int info = 0;
            e = cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess) e = cudaStreamSynchronize(s);
            if (e != cudaSuccess) break;
            if (info > n) break;                 // positive definite
            if (attempt == 5) {
                set_error("covariance matrix not positive definite, even with jitter (GPy jitchol gives up after 5 tries)");
                rc = 4;
            }
        }
    }
    if (e == cudaSuccess && !rc) {
        uint64_t nls = 0;
        rc = launch_chol_solve(s, A, n, y_dev, g->alpha, &nls, c->setup_variant);
        c->launches += nls;
        if (c->profile) c->fam_launches[F_SETUP] += nls;
        if (!rc && alpha_out) e = cudaMemcpyAsync(alpha_out, g->alpha, sizeof(double) * n, cudaMemcpyDeviceToHost, s);
        if (!rc && e == cudaSuccess && L_out)
            e = cudaMemcpyAsync(L_out, A, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost, s);
        if (!rc && e == cudaSuccess) rc = gp_pack_factor(c, g, A, GPD_FACTOR_CHOL);
    }
    if (e != cudaSuccess && !rc) set_error(std::string("device GP build failed: ") + cudaGetErrorString(e));
    cudaStreamSynchronize(s);
    cudaFree(Z_dev);
    cudaFree(y_dev);
    cudaFree(info_dev);
    if (e != cudaSuccess || rc) {
        gpd_gp_free(g);
        return rc ? rc : 1;
    }
    if (jitter_out) *jitter_out = jitter;
    *out = g;
    return 0;
}

// =================================================================================================
// model
// =================================================================================================
void gpd_model_free(gpd_model* mo) {
    if (!mo) return;
    mo->ctx->epoch++;
    cudaSetDevice(mo->ctx->device);
    cudaStreamSynchronize(mo->ctx->stream);
    cudaFree(mo->xf_dev);
    cudaFree(mo->pstar_dev);
    cudaFree(mo->bcols_dev);
    cudaFree(mo->weights_dev);
    delete mo;
}

int gpd_model_create(gpd_ctx* c, int32_t E, int32_t dim_x, int32_t dim_p, int32_t n_binary, const int32_t* binary,
                     gpd_gp* const* gps, const double* xmin, const double* xmax, const double* pmin,
                     const double* pmax, const double* ymean, const double* ystd, gpd_model** out) {
    GPD_CHECK(c && out && gps && xmin && xmax && pmin && pmax && ymean && ystd, "null argument");
    *out = nullptr;
    GPD_CHECK(E >= 1 && E <= kMaxE, "num_outputs out of range (1..8)");
    GPD_CHECK(dim_p >= 1 && dim_p <= kMaxP, "dim_p out of range (1..16)");
    GPD_CHECK(dim_x >= 1 && dim_x <= 64, "dim_x out of range");
    GPD_CHECK(n_binary >= 0 && n_binary <= kMaxBinary && (n_binary == 0 || binary), "bad binary variable list");
    GPD_CUDA(cudaSetDevice(c->device));
    gpd_model* mo = new gpd_model();
    mo->ctx = c;
    mo->E = E;
    mo->D = dim_x;
    mo->P = dim_p;
    mo->nb = n_binary;
    mo->R = 1 << n_binary;
    for (int v = 0; v < n_binary; ++v) mo->binary[v] = binary[v];
    mo->gps.assign(gps, gps + (size_t)E * mo->R);
    int live = 0;
    for (gpd_gp* g : mo->gps) {
        if (!g) continue;
        ++live;
        if (g->ctx != c || g->D != dim_x || g->P != dim_p) {
            delete mo;
            set_error("GP handle does not match the model (context, dim_x or dim_p)");
            return 2;
        }
    }
    if (!live) {
        delete mo;
        set_error("model without any GP");
        return 2;
    }
    mo->xmin.assign(xmin, xmin + dim_x);
    mo->xmax.assign(xmax, xmax + dim_x);
    mo->pmin.assign(pmin, pmin + dim_p);
    mo->pmax.assign(pmax, pmax + dim_p);
    mo->ymean.assign(ymean, ymean + E);
    mo->ystd.assign(ystd, ystd + E);
    mo->pmean.assign(dim_p, 0.0);
    mo->Sigma.assign((size_t)dim_p * dim_p, 0.0);
    cudaError_t err = cudaMalloc(&mo->xf_dev, 2 * sizeof(ModelXform));
    if (err == cudaSuccess) err = cudaMalloc(&mo->pstar_dev, sizeof(double) * dim_p);
    if (err == cudaSuccess && n_binary) {
        err = cudaMalloc(&mo->bcols_dev, sizeof(int) * n_binary);
        if (err == cudaSuccess) err = cudaMalloc(&mo->weights_dev, sizeof(int) * n_binary);
        if (err == cudaSuccess) {
            int w[kMaxBinary];
            for (int v = 0; v < n_binary; ++v) w[v] = binary_combo_weight(n_binary, v);
            err = cudaMemcpy(mo->bcols_dev, mo->binary, sizeof(int) * n_binary, cudaMemcpyHostToDevice);
            if (err == cudaSuccess) err = cudaMemcpy(mo->weights_dev, w, sizeof(int) * n_binary, cudaMemcpyHostToDevice);
        }
    }
    if (err != cudaSuccess) {
        set_error(std::string("model allocation failed: ") + cudaGetErrorString(err));
        gpd_model_free(mo);
        return 1;
    }
    *out = mo;
    return 0;
}

int gpd_model_set_pmean_sigma(gpd_model* mo, const double* pmean, const double* Sigma) {
    GPD_CHECK(mo && pmean, "null argument");
    gpd_ctx* c = mo->ctx;
    GPD_CUDA(cudaSetDevice(c->device));
    const int P = mo->P;
    mo->pmean.assign(pmean, pmean + P);
    mo->has_pmean = true;
    if (Sigma) {
        mo->Sigma.assign(Sigma, Sigma + (size_t)P * P);
        mo->has_sigma = true;
    }
    mo->pstar.resize(P);
    for (int a = 0; a < P; ++a)   // trans_p(pmean), surrogate_model.py:175
        mo->pstar[a] = (mo->pmean[a] - mo->pmin[a]) / (mo->pmax[a] - mo->pmin[a]);
    return upload_model_state(mo);
}

int gpd_model_set_pstar(gpd_model* mo, const double* pstar) {
    GPD_CHECK(mo && pstar, "null argument");
    GPD_CUDA(cudaSetDevice(mo->ctx->device));
    const int P = mo->P;
    mo->pstar.assign(pstar, pstar + P);
    for (int a = 0; a < P; ++a) mo->pmean[a] = mo->pmin[a] + pstar[a] * (mo->pmax[a] - mo->pmin[a]);
    mo->has_pmean = true;
    return upload_model_state(mo);
}

}  // extern "C"

namespace gpd {
static int upload_model_state(gpd_model* mo) {
    gpd_ctx* c = mo->ctx;
    const int P = mo->P, E = mo->E;
    ModelXform xf[2];
    memset(xf, 0, sizeof(xf));
    for (int k = 0; k < 2; ++k) {
        xf[k].E = E;
        xf[k].P = P;
    }
    for (int e = 0; e < E; ++e) {
        xf[0].ymean[e] = mo->ymean[e];
        xf[0].ystd[e] = mo->ystd[e];
        xf[1].ymean[e] = 0.0;
        xf[1].ystd[e] = 1.0;
    }
    for (int a = 0; a < P; ++a) {
        xf[0].pdiff[a] = mo->pmax[a] - mo->pmin[a];
        xf[1].pdiff[a] = 1.0;
    }
    for (int k = 0; k < P * P; ++k) xf[0].Sigma[k] = xf[1].Sigma[k] = mo->Sigma[k];
    bool sparse = false;
    for (gpd_gp* g : mo->gps) sparse = sparse || (g && g->factor_kind == GPD_FACTOR_ROOT);
    xf[0].var_floor = xf[1].var_floor = sparse ? 1e-15 : -HUGE_VAL;
    cudaStream_t s = c->stream;
    GPD_CUDA(cudaMemcpyAsync(mo->xf_dev, xf, sizeof(xf), cudaMemcpyHostToDevice, s));
    GPD_CUDA(cudaMemcpyAsync(mo->pstar_dev, mo->pstar.data(), sizeof(double) * P, cudaMemcpyHostToDevice, s));
    for (gpd_gp* g : mo->gps) {
        if (!g) continue;
        Span sp(c, F_SETUP, 1, 0);
        GPD_TRY(launch_pstate(s, g->kernel, g->n, g->n_pad, P, P, g->Zp, g->alpha, mo->pstar_dev, g->ls_p, g->var_x,
                              g->var_p, g->power_p, g->C, g->G, g->H));
    }
    GPD_CUDA(cudaStreamSynchronize(s));   // xf is a stack copy
    return 0;
}

// ---- shared driver for the hook-style entry points: host X in, host arrays out ------------------
struct HostOut {
    double *mu = nullptr, *s2 = nullptr, *dmu = nullptr, *d2mu = nullptr, *ds2 = nullptr, *d2s2 = nullptr;
};

static int hooks_host(gpd_model* mo, const double* X, int64_t N, int flags, const HostOut& o, bool transformed = false) {
    GPD_CHECK(mo && (X || N == 0), "null argument");
    GPD_CHECK(N >= 0, "negative N");
    gpd_ctx* c = mo->ctx;
    GPD_CUDA(cudaSetDevice(c->device));
    if (N == 0) return 0;
    Plan pl;
    pl.transformed = transformed;
    const int D = mo->D, E = mo->E, P = mo->P;
    GPD_TRY(plan_build(c, &mo, 1, flags, N, sizeof(double) * D, 0, pl));
    const size_t total = pl.per_cand_bytes * ((size_t)pl.chunk + kBlk) + (8u << 20) + pl.fixed_bytes + pl.entries.size() * sizeof(GpView) +
                         (size_t)pl.entries.size() * pl.tiles * sizeof(Item) * 10;
    GPD_TRY(ensure_scratch(c, total));
    Bump bump(c->scratch.p, c->scratch.cap, c);
    double* X_dev = bump.take<double>((size_t)pl.chunk * D);
    GPD_TRY(plan_materialize(c, pl, bump));
    cudaStream_t s = c->stream;
    for (int64_t start = 0; start < N; start += pl.chunk) {
        const int nc = (int)std::min<int64_t>(pl.chunk, N - start);
        GPD_TRY(run_chunk(c, pl, X_dev, D, nc, X + (size_t)start * D));
        const ModelBufs& b = pl.mb[0];
        auto down = [&](double* host, const double* dev, size_t per) -> cudaError_t {
            if (!host) return cudaSuccess;
            return cudaMemcpyAsync(host + (size_t)start * per, dev, sizeof(double) * (size_t)nc * per, cudaMemcpyDeviceToHost, s);
        };
        GPD_CUDA(down(o.mu, b.mu, E));
        GPD_CUDA(down(o.s2, b.s2, E));
        GPD_CUDA(down(o.dmu, b.dmu, (size_t)E * P));
        GPD_CUDA(down(o.d2mu, b.d2mu, (size_t)E * P * P));
        GPD_CUDA(down(o.ds2, b.ds2, (size_t)E * P));
        GPD_CUDA(down(o.d2s2, b.d2s2, (size_t)E * P * P));
        GPD_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}
}  // namespace gpd

extern "C" {

int gpd_predict(gpd_model* mo, const double* X, int64_t N, double* M_out, double* S_out) {
    HostOut o;
    o.mu = M_out;
    o.s2 = S_out;
    return hooks_host(mo, X, N, 0, o);
}
int gpd_dmu_dp(gpd_model* mo, const double* X, int64_t N, double* out) {
    GPD_CHECK(out || N == 0, "null output");
    HostOut o;
    o.dmu = out;
    return hooks_host(mo, X, N, WANT_DMU, o);
}
int gpd_d2mu_dp2(gpd_model* mo, const double* X, int64_t N, double* out) {
    GPD_CHECK(out || N == 0, "null output");
    HostOut o;
    o.d2mu = out;
    return hooks_host(mo, X, N, WANT_D2MU, o);
}
int gpd_ds2_dp(gpd_model* mo, const double* X, int64_t N, double* out) {
    GPD_CHECK(out || N == 0, "null output");
    HostOut o;
    o.ds2 = out;
    return hooks_host(mo, X, N, WANT_DS2, o);
}
int gpd_d2s2_dp2(gpd_model* mo, const double* X, int64_t N, double* out) {
    GPD_CHECK(out || N == 0, "null output");
    HostOut o;
    o.d2s2 = out;
    return hooks_host(mo, X, N, WANT_D2S2, o);
}
/* all hooks in one pass (any output may be NULL): what a caching host wrapper calls once per X */
int gpd_hooks(gpd_model* mo, const double* X, int64_t N, int32_t transformed, double* mu, double* s2, double* dmu,
              double* d2mu, double* ds2, double* d2s2) {
    HostOut o;
    o.mu = mu;
    o.s2 = s2;
    o.dmu = dmu;
    o.d2mu = d2mu;
    o.ds2 = ds2;
    o.d2s2 = d2s2;
    int flags = (dmu ? WANT_DMU : 0) | (d2mu ? WANT_D2MU : 0) | (ds2 ? WANT_DS2 : 0) | (d2s2 ? WANT_D2S2 : 0);
    return hooks_host(mo, X, N, flags, o, transformed != 0);
}

// =================================================================================================
// marginals, criteria, argmax
// =================================================================================================
static int order_flags(int order) { return order == 1 ? WANT_DMU : (WANT_DMU | WANT_D2MU | WANT_D2S2); }

int gpd_marginal(gpd_model* mo, const double* X, int64_t N, int order, double* mu_out, double* S_out) {
    GPD_CHECK(mo && (X || N == 0) && (N == 0 || (mu_out && S_out)), "null argument");
    GPD_CHECK(order == 1 || order == 2, "order must be 1 or 2");
    GPD_CHECK(mo->has_sigma, "Sigma is not set (taylor_first_order.py:38 reads model.Sigma)");
    gpd_ctx* c = mo->ctx;
    GPD_CUDA(cudaSetDevice(c->device));
    if (N == 0) return 0;
    const int D = mo->D, E = mo->E, P = mo->P;
    Plan pl;
    const size_t extra = sizeof(double) * ((size_t)D + (size_t)E * E + (order == 2 ? (size_t)E * P * P + E : 0));
    GPD_TRY(plan_build(c, &mo, 1, order_flags(order), N, extra, 0, pl));
    const size_t total = pl.per_cand_bytes * ((size_t)pl.chunk + kBlk) + (8u << 20) + pl.fixed_bytes + pl.entries.size() * sizeof(GpView) +
                         (size_t)pl.entries.size() * pl.tiles * sizeof(Item) * 10;
    GPD_TRY(ensure_scratch(c, total));
    Bump bump(c->scratch.p, c->scratch.cap, c);
    double* X_dev = bump.take<double>((size_t)pl.chunk * D);
    double* S_dev = bump.take<double>((size_t)pl.chunk * E * E);
    double* tmpA = order == 2 ? bump.take<double>((size_t)pl.chunk * E * P * P) : nullptr;
    double* tmp_s2 = order == 2 ? bump.take<double>((size_t)pl.chunk * E) : nullptr;
    GPD_TRY(plan_materialize(c, pl, bump));
    cudaStream_t s = c->stream;
    const double* Sigma_dev = (const double*)((const char*)mo->xf_dev + offsetof(ModelXform, Sigma));
    for (int64_t start = 0; start < N; start += pl.chunk) {
        const int nc = (int)std::min<int64_t>(pl.chunk, N - start);
        GPD_TRY(run_chunk(c, pl, X_dev, D, nc, X + (size_t)start * D));
        const ModelBufs& b = pl.mb[0];
        {
            Span sp(c, F_MARG, order == 2 ? 2 : 1, 0);
            GPD_TRY(launch_assemble(s, order, nc, E, P, b.mu, E, b.s2, b.dmu, b.d2mu, b.d2s2, Sigma_dev, S_dev, E * E, tmpA,
                                    tmp_s2));
        }
        GPD_CUDA(cudaMemcpyAsync(mu_out + (size_t)start * E, b.mu, sizeof(double) * (size_t)nc * E, cudaMemcpyDeviceToHost, s));
        GPD_CUDA(cudaMemcpyAsync(S_out + (size_t)start * E * E, S_dev, sizeof(double) * (size_t)nc * E * E,
                                 cudaMemcpyDeviceToHost, s));
        GPD_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int gpd_marginal_assemble(gpd_ctx* c, int order, int64_t N, int32_t E, int32_t P, double* mu_io, const double* s2,
                          const double* dmu, const double* ddmu, const double* dds2, const double* Sigma,
                          double* S_out) {
    GPD_CHECK(c && (N == 0 || (mu_io && s2 && dmu && Sigma && S_out)), "null argument");
    GPD_CHECK(order == 1 || order == 2, "order must be 1 or 2");
    GPD_CHECK(order == 1 || N == 0 || (ddmu && dds2), "second order needs ddmu and dds2");
    GPD_CHECK(E >= 1 && E <= 64 && P >= 1 && P <= 64, "bad shape");
    GPD_CUDA(cudaSetDevice(c->device));
    if (N == 0) return 0;
    const size_t per = sizeof(double) * ((size_t)2 * E + (size_t)E * P + (size_t)E * E +
                                         (order == 2 ? (size_t)3 * E * P * P + E : 0));
    int64_t chunk = std::min<int64_t>(N, std::max<int64_t>(1, (int64_t)((c->scratch_limit - (1 << 20)) / per)));
    chunk = std::min<int64_t>(chunk, 1 << 24);
    GPD_TRY(ensure_scratch(c, per * (size_t)chunk + (1 << 20)));
    Bump bump(c->scratch.p, c->scratch.cap, c);
    double* d_mu = bump.take<double>((size_t)chunk * E);
    double* d_s2 = bump.take<double>((size_t)chunk * E);
    double* d_dmu = bump.take<double>((size_t)chunk * E * P);
    double* d_S = bump.take<double>((size_t)chunk * E * E);
    double *d_ddmu = nullptr, *d_dds2 = nullptr, *tmpA = nullptr, *tmp_s2 = nullptr;
    if (order == 2) {
        d_ddmu = bump.take<double>((size_t)chunk * E * P * P);
        d_dds2 = bump.take<double>((size_t)chunk * E * P * P);
        tmpA = bump.take<double>((size_t)chunk * E * P * P);
        tmp_s2 = bump.take<double>((size_t)chunk * E);
    }
    double* d_Sigma = bump.take<double>((size_t)P * P);
    GPD_CHECK(bump.off <= bump.cap, "internal: scratch overflow");
    cudaStream_t s = c->stream;
    GPD_CUDA(cudaMemcpyAsync(d_Sigma, Sigma, sizeof(double) * P * P, cudaMemcpyHostToDevice, s));
    for (int64_t start = 0; start < N; start += chunk) {
        const int nc = (int)std::min<int64_t>(chunk, N - start);
        GPD_CUDA(cudaMemcpyAsync(d_mu, mu_io + (size_t)start * E, sizeof(double) * (size_t)nc * E, cudaMemcpyHostToDevice, s));
        GPD_CUDA(cudaMemcpyAsync(d_s2, s2 + (size_t)start * E, sizeof(double) * (size_t)nc * E, cudaMemcpyHostToDevice, s));
        GPD_CUDA(cudaMemcpyAsync(d_dmu, dmu + (size_t)start * E * P, sizeof(double) * (size_t)nc * E * P, cudaMemcpyHostToDevice, s));
        if (order == 2) {
            GPD_CUDA(cudaMemcpyAsync(d_ddmu, ddmu + (size_t)start * E * P * P, sizeof(double) * (size_t)nc * E * P * P,
                                     cudaMemcpyHostToDevice, s));
            GPD_CUDA(cudaMemcpyAsync(d_dds2, dds2 + (size_t)start * E * P * P, sizeof(double) * (size_t)nc * E * P * P,
                                     cudaMemcpyHostToDevice, s));
        }
        {
            Span sp(c, F_MARG, order == 2 ? 2 : 1, 0);
            GPD_TRY(launch_assemble(s, order, nc, E, P, d_mu, E, d_s2, d_dmu, d_ddmu, d_dds2, d_Sigma, d_S, E * E, tmpA, tmp_s2));
        }
        if (order == 2)
            GPD_CUDA(cudaMemcpyAsync(mu_io + (size_t)start * E, d_mu, sizeof(double) * (size_t)nc * E, cudaMemcpyDeviceToHost, s));
        GPD_CUDA(cudaMemcpyAsync(S_out + (size_t)start * E * E, d_S, sizeof(double) * (size_t)nc * E * E, cudaMemcpyDeviceToHost, s));
        GPD_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

int gpd_criterion(gpd_ctx* c, int kind, const double* mu, const double* S, int64_t N, int32_t M, int32_t E,
                  const double* noise, const double* pps, double* dc_out) {
    GPD_CHECK(c && (N == 0 || (mu && S && noise && pps && dc_out)), "null argument");
    GPD_CHECK(kind >= 0 && kind <= 4, "unknown criterion");
    GPD_CHECK(M >= 1 && M <= 64, "number of models out of range (1..64)");
    GPD_CHECK(E >= 1 && E <= kMaxE, "criteria support 1 <= E <= 8 outputs");
    GPD_CUDA(cudaSetDevice(c->device));
    if (N == 0) return 0;
    const size_t per = sizeof(double) * ((size_t)M * E + (size_t)M * E * E + 1) + sizeof(double) * criterion_work_doubles(1, M, E);
    int64_t chunk = std::min<int64_t>(N, std::max<int64_t>(1, (int64_t)((c->scratch_limit - (1 << 20)) / per)));
    chunk = std::min<int64_t>(chunk, 1 << 24);
    GPD_TRY(ensure_scratch(c, per * (size_t)chunk + (1 << 20)));
    Bump bump(c->scratch.p, c->scratch.cap, c);
    double* d_mu = bump.take<double>((size_t)chunk * M * E);
    double* d_S = bump.take<double>((size_t)chunk * M * E * E);
    double* d_dc = bump.take<double>((size_t)chunk);
    double* d_work = bump.take<double>(criterion_work_doubles((int)chunk, M, E));
    double* d_noise = bump.take<double>((size_t)E * E);
    double* d_pps = bump.take<double>(M);
    GPD_CHECK(bump.off <= bump.cap, "internal: scratch overflow");
    cudaStream_t s = c->stream;
    GPD_CUDA(cudaMemcpyAsync(d_noise, noise, sizeof(double) * E * E, cudaMemcpyHostToDevice, s));
    GPD_CUDA(cudaMemcpyAsync(d_pps, pps, sizeof(double) * M, cudaMemcpyHostToDevice, s));
    for (int64_t start = 0; start < N; start += chunk) {
        const int nc = (int)std::min<int64_t>(chunk, N - start);
        GPD_CUDA(cudaMemcpyAsync(d_mu, mu + (size_t)start * M * E, sizeof(double) * (size_t)nc * M * E, cudaMemcpyHostToDevice, s));
        GPD_CUDA(cudaMemcpyAsync(d_S, S + (size_t)start * M * E * E, sizeof(double) * (size_t)nc * M * E * E,
                                 cudaMemcpyHostToDevice, s));
        {
            Span sp(c, F_CRIT, 2, 0);
            GPD_TRY(launch_criterion(s, kind, d_mu, d_S, nc, M, E, d_noise, d_pps, d_dc, d_work));
        }
        GPD_CUDA(cudaMemcpyAsync(dc_out + start, d_dc, sizeof(double) * nc, cudaMemcpyDeviceToHost, s));
        GPD_TRY(singular_fetch(c));
        GPD_CUDA(cudaStreamSynchronize(s));
        GPD_TRY(singular_rc(c));
    }
    return 0;
}

int gpd_gauss_loglik(gpd_ctx* c, const double* Y, const double* M, const double* S, int64_t n, int32_t N, int32_t E,
                     double* logpdf_out, double* maha_out) {
    GPD_CHECK(c && (n == 0 || (Y && M && S && logpdf_out && maha_out)), "null argument");
    GPD_CHECK(n >= 0 && N >= 1 && N <= 4096, "bad shape");
    GPD_CHECK(E >= 1 && E <= kMaxE, "discrimination criteria support 1 <= E <= 8 outputs");
    GPD_CUDA(cudaSetDevice(c->device));
    if (n == 0) return 0;
    const size_t per = sizeof(double) * ((size_t)E + (size_t)N * (E + (size_t)E * E + 2));
    int64_t chunk = std::min<int64_t>(n, std::max<int64_t>(1, (int64_t)((c->scratch_limit - (1 << 20)) / per)));
    GPD_TRY(ensure_scratch(c, per * (size_t)chunk + (1 << 20)));
    Bump bump(c->scratch.p, c->scratch.cap, c);
    double* d_Y = bump.take<double>((size_t)chunk * E);
    double* d_M = bump.take<double>((size_t)chunk * N * E);
    double* d_S = bump.take<double>((size_t)chunk * N * E * E);
    double* d_l = bump.take<double>((size_t)chunk * N);
    double* d_q = bump.take<double>((size_t)chunk * N);
    GPD_CHECK(bump.off <= bump.cap, "internal: scratch overflow");
    cudaStream_t s = c->stream;
    for (int64_t start = 0; start < n; start += chunk) {
        const int64_t nc = std::min<int64_t>(chunk, n - start);
        GPD_CUDA(cudaMemcpyAsync(d_Y, Y + (size_t)start * E, sizeof(double) * (size_t)nc * E, cudaMemcpyHostToDevice, s));
        GPD_CUDA(cudaMemcpyAsync(d_M, M + (size_t)start * N * E, sizeof(double) * (size_t)nc * N * E, cudaMemcpyHostToDevice, s));
        GPD_CUDA(cudaMemcpyAsync(d_S, S + (size_t)start * N * E * E, sizeof(double) * (size_t)nc * N * E * E,
                                 cudaMemcpyHostToDevice, s));
        {
            Span sp(c, F_CRIT, 1, 0);
            GPD_TRY(launch_gauss_loglik(s, d_Y, d_M, d_S, nc, N, E, d_l, d_q));
        }
        GPD_CUDA(cudaMemcpyAsync(logpdf_out + (size_t)start * N, d_l, sizeof(double) * (size_t)nc * N, cudaMemcpyDeviceToHost, s));
        GPD_CUDA(cudaMemcpyAsync(maha_out + (size_t)start * N, d_q, sizeof(double) * (size_t)nc * N, cudaMemcpyDeviceToHost, s));
        GPD_TRY(singular_fetch(c));
        GPD_CUDA(cudaStreamSynchronize(s));
        GPD_TRY(singular_rc(c));
    }
    return 0;
}

int gpd_argmax(gpd_ctx* c, const double* dc, int64_t N, double* best_val, int64_t* best_idx) {
    GPD_CHECK(c && dc && best_val && best_idx, "null argument");
    GPD_CHECK(N >= 1, "argmax of an empty sequence");
    GPD_CUDA(cudaSetDevice(c->device));
    const int64_t chunk = std::min<int64_t>(N, (int64_t)std::min<uint64_t>(c->scratch_limit, 1ull << 30) / 8);
    GPD_TRY(ensure_scratch(c, (size_t)chunk * 8));
    double* d = (double*)c->scratch.p;
    cudaStream_t s = c->stream;
    for (int64_t start = 0; start < N; start += chunk) {
        const int64_t nc = std::min<int64_t>(chunk, N - start);
        GPD_CUDA(cudaMemcpyAsync(d, dc + start, sizeof(double) * nc, cudaMemcpyHostToDevice, s));
        Span sp(c, F_ARGMAX, 2, 0);
        GPD_TRY(launch_argmax(s, d, nc, start, c->best_val_dev, c->best_idx_dev, start == 0, c->argmax_partial));
    }
    long long bi = -1;
    GPD_CUDA(cudaMemcpyAsync(best_val, c->best_val_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
    GPD_CUDA(cudaMemcpyAsync(&bi, c->best_idx_dev, sizeof(long long), cudaMemcpyDeviceToHost, s));
    GPD_CUDA(cudaStreamSynchronize(s));
    *best_idx = bi;
    return 0;
}

// =================================================================================================
// fused scoring
// =================================================================================================
static int score_impl(gpd_ctx* c, gpd_model* const* models, int32_t M, const double* X, bool x_on_device, int64_t N,
                      int order, const int32_t* kinds, int nk, const double* noise, const double* pps, double* dc_out,
                      bool dc_on_device, double* best_val, int64_t* best_idx, int64_t idx_offset, bool finish = true) {
    GPD_CHECK(c && models && X && noise && pps && kinds, "null argument");
    GPD_CHECK(N >= 1, "no candidates");
    GPD_CHECK(M >= 1 && M <= 64, "number of models out of range (1..64)");
    GPD_CHECK(order == 1 || order == 2, "order must be 1 or 2");
    GPD_CHECK(nk >= 1 && nk <= kMaxKinds, "1..5 criteria per call");
    for (int k = 0; k < nk; ++k) GPD_CHECK(kinds[k] >= 0 && kinds[k] <= 4, "unknown criterion");
    GPD_CUDA(cudaSetDevice(c->device));
    const int E = models[0]->E, D = models[0]->D;
    int maxP = 0;
    for (int m = 0; m < M; ++m) {
        GPD_CHECK(models[m] && models[m]->E == E && models[m]->D == D, "rival models must share num_outputs and dim_x");
        GPD_CHECK(models[m]->has_sigma, "Sigma is not set on every model");
        maxP = std::max(maxP, models[m]->P);
    }
    NvtxRange nvtx_score("gpd/score");
    const bool keep_dc = dc_out != nullptr;
    // ---- plan: reuse the previous call's plan and device tables when nothing they depend on has changed ----
    ScoreState* st = c->score_state;
    bool hit = st && st->epoch == c->epoch && st->N == N && st->order == order && st->x_on_device == (int)x_on_device &&
               st->keep_dc == (int)keep_dc && st->dc_on_device == (int)dc_on_device && (int)st->models.size() == M &&
               (int)st->kinds.size() == nk && st->scratch_p == c->scratch.p && st->scratch_cap == c->scratch.cap;
    for (int m = 0; hit && m < M; ++m) hit = st->models[m] == models[m];
    for (int k = 0; hit && k < nk; ++k) hit = st->kinds[k] == kinds[k];
    if (hit) hit = memcmp(st->noise.data(), noise, sizeof(double) * E * E) == 0 && memcmp(st->pps.data(), pps, sizeof(double) * M) == 0;
    cudaStream_t s = c->stream;
    if (!hit) {
        delete c->score_state;
        c->score_state = st = new ScoreState();
        Plan& pl = st->pl;
        pl.skip_finalize = (order == 1);
        size_t extra = sizeof(double) * ((size_t)M * E + (size_t)M * E * E + nk) + sizeof(double) * criterion_work_doubles(1, M, E) +
                       (crit_partial_bytes(kBlk, nk) + kBlk - 1) / kBlk;
        if (!x_on_device) extra += sizeof(double) * D;
        if (order == 2) extra += sizeof(double) * ((size_t)E * maxP * maxP + E);
        GPD_TRY(plan_build(c, models, M, order_flags(order), N, extra, 0, pl));
        const size_t total = pl.per_cand_bytes * ((size_t)pl.chunk + kBlk) + (8u << 20) + pl.fixed_bytes + pl.entries.size() * sizeof(GpView) +
                             (size_t)pl.entries.size() * pl.tiles * sizeof(Item) * 10;
        GPD_TRY(ensure_scratch(c, total));
        Bump bump(c->scratch.p, c->scratch.cap, c);
        const int Nc0 = pl.chunk;
        st->X_stage = x_on_device ? nullptr : bump.take<double>((size_t)Nc0 * D);
        st->mu_all = order == 2 ? bump.take<double>((size_t)Nc0 * M * E) : nullptr;
        st->S_all = order == 2 ? bump.take<double>((size_t)Nc0 * M * E * E) : nullptr;
        st->dc_dev = (keep_dc && !dc_on_device) ? bump.take<double>((size_t)Nc0 * nk) : nullptr;
        st->work = bump.take<double>(criterion_work_doubles(Nc0, M, E));
        st->tmpA = order == 2 ? bump.take<double>((size_t)Nc0 * E * maxP * maxP) : nullptr;
        st->tmp_s2 = order == 2 ? bump.take<double>((size_t)Nc0 * E) : nullptr;
        st->d_noise = bump.take<double>((size_t)E * E);
        st->d_pps = bump.take<double>(M);
        st->partial = bump.take<char>(crit_partial_bytes(Nc0, nk));
        st->pms_dev = bump.take<PostModel>(M);
        GPD_TRY(plan_materialize(c, pl, bump));
        if (order == 2) {
            for (int m = 0; m < M; ++m) {   // finalize writes the means straight into the (N, M, E) criterion layout
                pl.mb[m].mu = st->mu_all + (size_t)m * E;
                pl.mb[m].ldmu = M * E;
            }
        } else {
            std::vector<PostModel> pms(M);
            for (int m = 0; m < M; ++m) pms[m] = PostModel{models[m]->xf_dev, models[m]->P, pl.mb[m].ldraw, pl.mb[m].ng};
            GPD_CUDA(cudaMemcpyAsync(st->pms_dev, pms.data(), sizeof(PostModel) * M, cudaMemcpyHostToDevice, s));
        }
        GPD_CUDA(cudaMemcpyAsync(st->d_noise, noise, sizeof(double) * E * E, cudaMemcpyHostToDevice, s));
        GPD_CUDA(cudaMemcpyAsync(st->d_pps, pps, sizeof(double) * M, cudaMemcpyHostToDevice, s));
        st->models.assign(models, models + M);
        st->kinds.assign(kinds, kinds + nk);
        st->noise.assign(noise, noise + (size_t)E * E);
        st->pps.assign(pps, pps + M);
        st->N = N;
        st->order = order;
        st->x_on_device = x_on_device;
        st->keep_dc = keep_dc;
        st->dc_on_device = dc_on_device;
        st->scratch_p = c->scratch.p;
        st->scratch_cap = c->scratch.cap;
        st->epoch = c->epoch;             // after ensure_scratch, which bumps it
    }
    Plan& pl = st->pl;
    const int Nc = pl.chunk;
    double *X_stage = st->X_stage, *mu_all = st->mu_all, *S_all = st->S_all, *dc_dev = st->dc_dev, *work = st->work,
           *tmpA = st->tmpA, *tmp_s2 = st->tmp_s2, *d_noise = st->d_noise, *d_pps = st->d_pps;
    void* partial = st->partial;
    PostModel* pms_dev = st->pms_dev;
    int kinds_h[kMaxKinds];
    for (int k = 0; k < nk; ++k) kinds_h[k] = kinds[k];
    for (int64_t start = 0; start < N; start += Nc) {
        const int nc = (int)std::min<int64_t>(Nc, N - start);
        if (x_on_device) GPD_TRY(run_chunk(c, pl, const_cast<double*>(X) + (size_t)start * D, D, nc));
        else GPD_TRY(run_chunk(c, pl, X_stage, D, nc, X + (size_t)start * D));
        if (order == 1) {
            Span sp(c, F_MARG, 1, 0);
            GPD_TRY(launch_post1(s, pms_dev, pl.raws_dev, M, E, nc, nc, d_noise, work, kinds_h, nk));
        } else {
            for (int m = 0; m < M; ++m) {
                gpd_model* mo = models[m];
                const ModelBufs& b = pl.mb[m];
                const double* Sigma_dev = (const double*)((const char*)mo->xf_dev + offsetof(ModelXform, Sigma));
                Span sp(c, F_MARG, 2, 0);
                GPD_TRY(launch_assemble(s, order, nc, E, mo->P, b.mu, b.ldmu, b.s2, b.dmu, b.d2mu, b.d2s2, Sigma_dev,
                                        S_all + (size_t)m * E * E, M * E * E, tmpA, tmp_s2));
            }
            Span sp(c, F_CRIT, 1, 0);
            GPD_TRY(launch_crit_prep(s, mu_all, S_all, nc, M, E, d_noise, work, kinds_h, nk));
        }
        double* dcp[kMaxKinds];
        for (int k = 0; k < nk; ++k)
            dcp[k] = !keep_dc ? nullptr : (dc_on_device ? dc_out + (size_t)k * N + start : dc_dev + (size_t)k * Nc);
        {
            Span sp(c, F_CRIT, 2, 0);
            GPD_TRY(launch_crit_multi(s, kinds_h, nk, dcp, nc, M, E, d_noise, d_pps, work, idx_offset + start, partial,
                                      start == 0, c->best_val_dev, c->best_idx_dev));
        }
        if (keep_dc && !dc_on_device)
            for (int k = 0; k < nk; ++k)
                GPD_CUDA(cudaMemcpyAsync(dc_out + (size_t)k * N + start, dcp[k], sizeof(double) * nc, cudaMemcpyDeviceToHost, s));
    }
    if (!finish) return 0;      // gpd_score_sharded: the running best stays on the device for the all-gather
    double* bv = (double*)c->result_pinned;
    long long* bi = (long long*)(bv + kMaxKinds);
    GPD_CUDA(cudaMemcpyAsync(bv, c->best_val_dev, sizeof(double) * nk, cudaMemcpyDeviceToHost, s));
    GPD_CUDA(cudaMemcpyAsync(bi, c->best_idx_dev, sizeof(long long) * nk, cudaMemcpyDeviceToHost, s));
    GPD_TRY(singular_fetch(c));
    GPD_CUDA(cudaStreamSynchronize(s));
    GPD_TRY(singular_rc(c));
    for (int k = 0; k < nk; ++k) {
        if (best_val) best_val[k] = bv[k];
        if (best_idx) best_idx[k] = bi[k];
    }
    return 0;
}

int gpd_score(gpd_ctx* c, gpd_model* const* models, int32_t M, const double* X, int64_t N, int order, int kind,
              const double* noise, const double* pps, double* dc_out, double* best_val, int64_t* best_idx,
              int64_t idx_offset) {
    const int32_t k = kind;
    return score_impl(c, models, M, X, false, N, order, &k, 1, noise, pps, dc_out, false, best_val, best_idx, idx_offset);
}
int gpd_score_dev(gpd_ctx* c, gpd_model* const* models, int32_t M, const double* X_dev, int64_t N, int order, int kind,
                  const double* noise, const double* pps, double* dc_out_dev, double* best_val, int64_t* best_idx,
                  int64_t idx_offset) {
    const int32_t k = kind;
    return score_impl(c, models, M, X_dev, true, N, order, &k, 1, noise, pps, dc_out_dev, true, best_val, best_idx, idx_offset);
}
int gpd_score_multi(gpd_ctx* c, gpd_model* const* models, int32_t M, const double* X, int32_t x_on_device, int64_t N,
                    int order, const int32_t* kinds, int32_t nkinds, const double* noise, const double* pps,
                    double* dc_out, int32_t dc_on_device, double* best_val, int64_t* best_idx, int64_t idx_offset) {
    return score_impl(c, models, M, X, x_on_device != 0, N, order, kinds, nkinds, noise, pps, dc_out, dc_on_device != 0,
                      best_val, best_idx, idx_offset);
}

// =================================================================================================
// design gradients: first-order Taylor marginal and its derivative with respect to x
// =================================================================================================
int gpd_marginal_dx(gpd_model* mo, const double* X, int64_t N, double* mu_out, double* S_out, double* dmu_dx_out,
                    double* dS_dx_out) {
    GPD_CHECK(mo && (X || N == 0), "null argument");
    GPD_CHECK(N >= 0, "negative N");
    GPD_CHECK(N == 0 || (mu_out && S_out && dmu_dx_out && dS_dx_out), "null output");
    GPD_CHECK(mo->has_pmean, "pmean is not set (surrogate_model.py:171)");
    GPD_CHECK(mo->has_sigma, "Sigma is not set (taylor_first_order.py:38 reads model.Sigma)");
    gpd_ctx* c = mo->ctx;
    GPD_CUDA(cudaSetDevice(c->device));
    if (N == 0) return 0;
    NvtxRange nvtx("gpd/marginal_dx");
    const int E = mo->E, R = mo->R, D = mo->D, P = mo->P, G = E * R;
    // every GP of a model shares the kernel class and the active design dimensions (gp_model.py:115-118)
    const gpd_gp* g0 = nullptr;
    long long n_pad_sum = 0, n_pad_max = 0;
    for (gpd_gp* g : mo->gps)
        if (g) {
            if (!g0) g0 = g;
            GPD_CHECK(g->nx == g0->nx, "GPs of one model must share the active design dimensions");
            n_pad_sum += g->n_pad;
            n_pad_max = std::max<long long>(n_pad_max, g->n_pad);
        }
    GPD_CHECK(g0, "model without any GP");
    const int nx = g0->nx, NR = 1 + nx, ncol = 1 + P, NG = NR * (NR + 1) / 2;
    // scratch per candidate: X, 2 panel sets (in / out), contractions, Gram rows, outputs
    const size_t per = sizeof(double) * ((size_t)D + 2 * (size_t)n_pad_sum * NR + (size_t)G * (NR * ncol + NG) +
                                         (size_t)E * (1 + E) * (1 + D)) +
                       ((size_t)G * sizeof(Item) + kBlk - 1) / kBlk;            // one work item per (GP, tile)
    GPD_CHECK(c->scratch_limit > (size_t)kBlk * per + (8u << 20) + (size_t)G * sizeof(GpView),
              "scratch limit too small for the design-gradient path");
    int64_t chunk = (int64_t)((c->scratch_limit - (8u << 20) - (size_t)G * sizeof(GpView)) / per) / kBlk * kBlk;
    chunk = std::min<int64_t>(chunk, 1 << 20);
    chunk = std::min<int64_t>(chunk, (N + kBlk - 1) / kBlk * kBlk);
    GPD_CHECK(chunk >= kBlk, "scratch limit too small");
    const int Nc = (int)chunk, tiles = Nc / kBlk;
    GPD_TRY(ensure_scratch(c, per * (size_t)Nc + (8u << 20) + (size_t)G * sizeof(GpView)));
    Bump bump(c->scratch.p, c->scratch.cap, c);
    double* X_dev = bump.take<double>((size_t)Nc * D);
    double* pan_in = bump.take<double>((size_t)n_pad_sum * NR * Nc);
    double* pan_out = bump.take<double>((size_t)n_pad_sum * NR * Nc);
    double* acc = bump.take<double>((size_t)G * Nc * NR * ncol);
    double* gram = bump.take<double>((size_t)G * Nc * NG);
    double* d_mu = bump.take<double>((size_t)Nc * E);
    double* d_S = bump.take<double>((size_t)Nc * E * E);
    double* d_dmu = bump.take<double>((size_t)Nc * E * D);
    double* d_dS = bump.take<double>((size_t)Nc * E * E * D);
    GpView* views_dev = bump.take<GpView>(G);
    Item* items_dev = bump.take<Item>((size_t)G * tiles);
    RawOut* raw_dev = bump.take<RawOut>(1);
    GPD_CHECK(bump.off <= bump.cap, "internal: scratch overflow");
    std::vector<GpView> views(G);
    std::vector<Item> items;
    long long off_w = 0;
    for (int gi = 0; gi < G; ++gi) {
        GpView& v = views[gi];
        memset(&v, 0, sizeof(v));
        const gpd_gp* g = mo->gps[gi];
        if (!g) continue;                       // n == 0 marks "no GP for this binary combination"
        v.kernel = g->kernel;
        v.n = g->n;
        v.n_pad = g->n_pad;
        v.nblk = g->nblk;
        v.nx = g->nx;
        v.P = g->P;
        v.kss = g->var_x * g->var_p;
        v.power_x = g->power_x;
        v.Xs = g->Xs;
        v.XsE = g->XsE;
        v.C = g->C;
        v.G = g->G;
        v.H = g->H;
        v.LinvF = g->LinvF;
        v.LinvB = g->LinvB;
        for (int d = 0; d < g->nx; ++d) {
            v.ls[d] = g->ls_x[d];
            v.x_active[d] = g->x_active[d];
            v.xmin[d] = mo->xmin[g->x_active[d]];
            v.xdiff[d] = mo->xmax[g->x_active[d]] - mo->xmin[g->x_active[d]];
            v.xscale[d] = 1.0 / (v.xdiff[d] * g->ls_x[d]) * (g->kernel == GPD_KERN_RBF ? kRbfCoordScaleHost : 1.0);
        }
        v.out_e = gi;                           // one Gram slot per (output, binary combination)
        v.model = 0;
        v.panel_off = 0;
        v.panel_off_w = off_w;
        off_w += (long long)tiles * g->n_pad * kBlk * NR;
        for (int t = 0; t < tiles; ++t) items.push_back(Item{gi, t});
    }
    RawOut raw;
    memset(&raw, 0, sizeof(raw));
    raw.gram = gram;
    cudaStream_t s = c->stream;
    GPD_CUDA(cudaMemcpyAsync(views_dev, views.data(), sizeof(GpView) * G, cudaMemcpyHostToDevice, s));
    GPD_CUDA(cudaMemcpyAsync(items_dev, items.data(), sizeof(Item) * items.size(), cudaMemcpyHostToDevice, s));
    GPD_CUDA(cudaMemcpyAsync(raw_dev, &raw, sizeof(RawOut), cudaMemcpyHostToDevice, s));
    XgradArgs a;
    memset(&a, 0, sizeof(a));
    a.E = E; a.R = R; a.D = D; a.P = P; a.nx = nx; a.nb = mo->nb; a.ngram = NG;
    for (int d = 0; d < nx; ++d) a.x_active[d] = g0->x_active[d];
    for (int v = 0; v < mo->nb; ++v) {
        a.bcols[v] = mo->binary[v];
        a.weights[v] = binary_combo_weight(mo->nb, v);
    }
    a.xf = mo->xf_dev;
    a.gps = views_dev;
    a.X = X_dev;
    a.ldx = D;
    a.acc = acc;
    a.gram = gram;
    a.mu = d_mu; a.S = d_S; a.dmu = d_dmu; a.dS = d_dS;
    for (int64_t start = 0; start < N; start += Nc) {
        const int nc = (int)std::min<int64_t>(Nc, N - start);
        a.chunk_n = nc;
        GPD_CUDA(cudaMemcpyAsync(X_dev, X + (size_t)start * D, sizeof(double) * (size_t)nc * D, cudaMemcpyHostToDevice, s));
        {
            Span sp(c, F_EVAL, 1, 0);
            GPD_TRY(launch_xgrad_eval(s, views_dev, G, tiles, X_dev, D, nc, pan_in, acc));
        }
        {
            Span sp(c, F_TRI, 1, 0);
            GPD_TRY(launch_tri(s, views_dev, items_dev, (int)items.size(), NR, NR, NR, 0, pan_in, pan_out, nullptr, 0, nc,
                               c->num_sms, c->tri_variant, tri_cps_for(c, Plan()), c->tri_trim));
        }
        {
            Span sp(c, F_GRAM, 1, 0);
            GPD_TRY(launch_gram(s, views_dev, items_dev, (int)items.size(), NR, pan_out, raw_dev, nc));
        }
        {
            Span sp(c, F_MARG, 1, 0);
            GPD_TRY(launch_xgrad_finalize(s, a));
        }
        GPD_CUDA(cudaMemcpyAsync(mu_out + (size_t)start * E, d_mu, sizeof(double) * (size_t)nc * E, cudaMemcpyDeviceToHost, s));
        GPD_CUDA(cudaMemcpyAsync(S_out + (size_t)start * E * E, d_S, sizeof(double) * (size_t)nc * E * E, cudaMemcpyDeviceToHost, s));
        GPD_CUDA(cudaMemcpyAsync(dmu_dx_out + (size_t)start * E * D, d_dmu, sizeof(double) * (size_t)nc * E * D,
                                 cudaMemcpyDeviceToHost, s));
        GPD_CUDA(cudaMemcpyAsync(dS_dx_out + (size_t)start * E * E * D, d_dS, sizeof(double) * (size_t)nc * E * E * D,
                                 cudaMemcpyDeviceToHost, s));
        GPD_CUDA(cudaStreamSynchronize(s));
    }
    return 0;
}

// =================================================================================================
// marginals of ALL rival models in one pass, kept resident for the criteria (the drop-in call pattern of
// demos/case_study.py:284-293 without the host round trip of mu / S between taylor_* and HR..JR)
// =================================================================================================
}  // extern "C"

struct gpd_resident {
    gpd_ctx* ctx = nullptr;
    int64_t N = 0;
    int M = 0, E = 0;
    double *mu_dev = nullptr, *S_dev = nullptr;     // (N, M, E) and (N, M, E, E), the layout the criteria take
};

extern "C" {

static void res_pool_put(gpd_ctx* c, void* p, size_t bytes) {
    if (!p) return;
    size_t held = bytes;
    for (auto& e : c->res_pool) held += e.second;
    if (c->res_pool.size() >= 8 || held > (4ull << 30)) {
        cudaFree(p);
        return;
    }
    c->res_pool.push_back({p, bytes});
}
static void* res_pool_take(gpd_ctx* c, size_t bytes) {
    for (size_t i = 0; i < c->res_pool.size(); ++i)
        if (c->res_pool[i].second == bytes) {
            void* p = c->res_pool[i].first;
            c->res_pool.erase(c->res_pool.begin() + i);
            return p;
        }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        for (auto& e : c->res_pool) cudaFree(e.first);     // make room and retry once
        c->res_pool.clear();
        if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
    }
    return p;
}

void gpd_resident_free(gpd_resident* r) {
    if (!r) return;
    cudaSetDevice(r->ctx->device);
    cudaStreamSynchronize(r->ctx->stream);
    res_pool_put(r->ctx, r->mu_dev, sizeof(double) * (size_t)r->N * r->M * r->E);
    res_pool_put(r->ctx, r->S_dev, sizeof(double) * (size_t)r->N * r->M * r->E * r->E);
    delete r;
}

int gpd_marginals(gpd_ctx* c, gpd_model* const* models, int32_t M, const double* X, int64_t N, int order,
                  double* mu_out, double* S_out, gpd_resident** keep) {
    GPD_CHECK(c && models && (X || N == 0), "null argument");
    GPD_CHECK(N >= 0, "negative N");
    GPD_CHECK(M >= 1 && M <= 64, "number of models out of range (1..64)");
    GPD_CHECK(order == 1 || order == 2, "order must be 1 or 2");
    GPD_CHECK(N == 0 || (mu_out && S_out) || keep, "nowhere to put the result");
    if (keep) *keep = nullptr;
    GPD_CUDA(cudaSetDevice(c->device));
    const int E = models[0]->E, D = models[0]->D;
    int maxP = 0;
    for (int m = 0; m < M; ++m) {
        GPD_CHECK(models[m] && models[m]->E == E && models[m]->D == D, "rival models must share num_outputs and dim_x");
        GPD_CHECK(models[m]->has_sigma, "Sigma is not set on every model (taylor_first_order.py:38 reads model.Sigma)");
        maxP = std::max(maxP, models[m]->P);
    }
    if (N == 0) return 0;
    NvtxRange nvtx("gpd/marginals");
    gpd_resident* res = nullptr;
    if (keep) {
        res = new gpd_resident();
        res->ctx = c;
        res->N = N;
        res->M = M;
        res->E = E;
        res->mu_dev = (double*)res_pool_take(c, sizeof(double) * (size_t)N * M * E);
        res->S_dev = (double*)res_pool_take(c, sizeof(double) * (size_t)N * M * E * E);
        if (!res->mu_dev || !res->S_dev) {
            set_error("resident marginals: out of device memory");
            gpd_resident_free(res);
            return 1;
        }
    }
    Plan pl;
    size_t extra = sizeof(double) * (size_t)D;
    if (!res) extra += sizeof(double) * ((size_t)M * E + (size_t)M * E * E);
    if (order == 2) extra += sizeof(double) * ((size_t)E * maxP * maxP + E);
    int rc = plan_build(c, models, M, order_flags(order), N, extra, 0, pl);
    const size_t total = pl.per_cand_bytes * ((size_t)pl.chunk + kBlk) + (8u << 20) + pl.fixed_bytes + pl.entries.size() * sizeof(GpView) +
                         (size_t)pl.entries.size() * pl.tiles * sizeof(Item) * 10;
    if (!rc) rc = ensure_scratch(c, total);
    if (rc) {
        gpd_resident_free(res);
        return rc;
    }
    Bump bump(c->scratch.p, c->scratch.cap, c);
    const int Nc = pl.chunk;
    double* X_dev = bump.take<double>((size_t)Nc * D);
    double* mu_tmp = res ? nullptr : bump.take<double>((size_t)Nc * M * E);
    double* S_tmp = res ? nullptr : bump.take<double>((size_t)Nc * M * E * E);
    double* tmpA = order == 2 ? bump.take<double>((size_t)Nc * E * maxP * maxP) : nullptr;
    double* tmp_s2 = order == 2 ? bump.take<double>((size_t)Nc * E) : nullptr;
    rc = plan_materialize(c, pl, bump);
    cudaStream_t s = c->stream;
    for (int64_t start = 0; start < N && !rc; start += Nc) {
        const int nc = (int)std::min<int64_t>(Nc, N - start);
        double* mu_c = res ? res->mu_dev + (size_t)start * M * E : mu_tmp;
        double* S_c = res ? res->S_dev + (size_t)start * M * E * E : S_tmp;
        for (int m = 0; m < M; ++m) {        // finalize writes the means straight into the (N, M, E) layout
            pl.mb[m].mu = mu_c + (size_t)m * E;
            pl.mb[m].ldmu = M * E;
        }
        rc = run_chunk(c, pl, X_dev, D, nc, X + (size_t)start * D);
        for (int m = 0; m < M && !rc; ++m) {
            gpd_model* mo = models[m];
            const ModelBufs& b = pl.mb[m];
            const double* Sigma_dev = (const double*)((const char*)mo->xf_dev + offsetof(ModelXform, Sigma));
            Span sp(c, F_MARG, order == 2 ? 2 : 1, 0);
            rc = launch_assemble(s, order, nc, E, mo->P, b.mu, b.ldmu, b.s2, b.dmu, b.d2mu, b.d2s2, Sigma_dev,
                                 S_c + (size_t)m * E * E, M * E * E, tmpA, tmp_s2);
        }
        if (!rc && mu_out) {
            cudaError_t e = cudaMemcpyAsync(mu_out + (size_t)start * M * E, mu_c, sizeof(double) * (size_t)nc * M * E,
                                            cudaMemcpyDeviceToHost, s);
            if (e == cudaSuccess)
                e = cudaMemcpyAsync(S_out + (size_t)start * M * E * E, S_c, sizeof(double) * (size_t)nc * M * E * E,
                                    cudaMemcpyDeviceToHost, s);
            if (e != cudaSuccess) {
                set_error(std::string("marginals download: ") + cudaGetErrorString(e));
                rc = 1;
            }
        }
        if (!rc && !res && cudaStreamSynchronize(s) != cudaSuccess) {      // the chunk buffers are reused
            set_error("marginals: stream synchronisation failed");
            rc = 1;
        }
    }
    if (!rc && cudaStreamSynchronize(s) != cudaSuccess) {
        set_error(std::string("marginals: ") + cudaGetErrorString(cudaGetLastError()));
        rc = 1;
    }
    if (rc) {
        gpd_resident_free(res);
        return rc;
    }
    if (keep) *keep = res;
    return 0;
}

int gpd_criterion_resident(gpd_ctx* c, int kind, const gpd_resident* r, const double* noise, const double* pps,
                           double* dc_out) {
    GPD_CHECK(c && r && r->ctx == c && noise && pps && dc_out, "null argument or resident marginals of another context");
    GPD_CHECK(kind >= 0 && kind <= 4, "unknown criterion");
    const int M = r->M, E = r->E;
    GPD_CHECK(E >= 1 && E <= kMaxE, "criteria support 1 <= E <= 8 outputs");
    GPD_CUDA(cudaSetDevice(c->device));
    const int64_t N = r->N;
    const size_t per = sizeof(double) + sizeof(double) * criterion_work_doubles(1, M, E);
    int64_t chunk = std::min<int64_t>(N, std::max<int64_t>(1, (int64_t)((c->scratch_limit - (1 << 20)) / per)));
    chunk = std::min<int64_t>(chunk, 1 << 24);
    GPD_TRY(ensure_scratch(c, per * (size_t)chunk + (1 << 20)));
    Bump bump(c->scratch.p, c->scratch.cap, c);
    double* d_dc = bump.take<double>((size_t)chunk);
    double* d_work = bump.take<double>(criterion_work_doubles((int)chunk, M, E));
    double* d_noise = bump.take<double>((size_t)E * E);
    double* d_pps = bump.take<double>(M);
    GPD_CHECK(bump.off <= bump.cap, "internal: scratch overflow");
    cudaStream_t s = c->stream;
    GPD_CUDA(cudaMemcpyAsync(d_noise, noise, sizeof(double) * E * E, cudaMemcpyHostToDevice, s));
    GPD_CUDA(cudaMemcpyAsync(d_pps, pps, sizeof(double) * M, cudaMemcpyHostToDevice, s));
    for (int64_t start = 0; start < N; start += chunk) {
        const int nc = (int)std::min<int64_t>(chunk, N - start);
        {
            Span sp(c, F_CRIT, 2, 0);
            GPD_TRY(launch_criterion(s, kind, r->mu_dev + (size_t)start * M * E, r->S_dev + (size_t)start * M * E * E, nc, M, E,
                                     d_noise, d_pps, d_dc, d_work));
        }
        GPD_CUDA(cudaMemcpyAsync(dc_out + start, d_dc, sizeof(double) * nc, cudaMemcpyDeviceToHost, s));
        GPD_TRY(singular_fetch(c));
        GPD_CUDA(cudaStreamSynchronize(s));
        GPD_TRY(singular_rc(c));
    }
    return 0;
}

// =================================================================================================
// multi-GPU: candidates sharded, GP state replicated, ONE collective per scoring call
// =================================================================================================
}  // extern "C"

#include <dlfcn.h>
#include <nccl.h>      // types and prototypes only: the library is bound at run time (no link dependency)

namespace gpd {
// NCCL is resolved with dlopen so that libgpdoemd_b200.so loads on a machine without NCCL; the sharded entry
// points then fail loudly.  If the process already holds a libnccl.so.2 (e.g. the one bundled with PyTorch) that
// instance is reused, so both sides talk to one NCCL.
struct NcclApi {
    void* handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclAllGather) AllGather = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
};
static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.handle ? &api : nullptr;
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the instance the process already holds, if any
    if (!h && getenv("GPDOEMD_NCCL_LIB")) h = dlopen(getenv("GPDOEMD_NCCL_LIB"), RTLD_NOW);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW);
    if (!h) return nullptr;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.GetVersion = (decltype(api.GetVersion))dlsym(h, "ncclGetVersion");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllGather || !api.GetErrorString) return nullptr;
    api.handle = h;
    return &api;
}
#define GPD_NCCL(api, call)                                                                            \
    do {                                                                                               \
        ncclResult_t r__ = (call);                                                                     \
        if (r__ != ncclSuccess) {                                                                      \
            ::gpd::set_error(std::string(#call) + ": " + (api)->GetErrorString(r__));                  \
            return 5;                                                                                  \
        }                                                                                              \
    } while (0)
}  // namespace gpd

struct gpd_comm {
    gpd_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    bool owned = false;             // created by gpd_comm_init (destroyed with the handle) or adopted from the caller
    int rank = 0, world = 1;
    void *send_dev = nullptr, *recv_dev = nullptr;     // Best[kMaxKinds], Best[world][kMaxKinds]
};

extern "C" {

int gpd_nccl_version(int32_t* version_out) {
    GPD_CHECK(version_out, "null argument");
    NcclApi* api = nccl_api();
    GPD_CHECK(api, "libnccl.so.2 cannot be loaded: the sharded entry points need NCCL (there is no fallback)");
    int v = 0;
    if (api->GetVersion) GPD_NCCL(api, api->GetVersion(&v));
    *version_out = v;
    return 0;
}

int gpd_nccl_unique_id(char id_out[GPD_NCCL_ID_BYTES]) {
    GPD_CHECK(id_out, "null argument");
    static_assert(sizeof(ncclUniqueId) == GPD_NCCL_ID_BYTES, "ncclUniqueId is 128 bytes");
    NcclApi* api = nccl_api();
    GPD_CHECK(api, "libnccl.so.2 cannot be loaded: the sharded entry points need NCCL (there is no fallback)");
    ncclUniqueId id;
    GPD_NCCL(api, api->GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

static int comm_alloc(gpd_ctx* c, gpd_comm* cm) {
    GPD_CUDA(cudaMalloc(&cm->send_dev, best_pair_bytes() * kMaxKinds));
    GPD_CUDA(cudaMalloc(&cm->recv_dev, best_pair_bytes() * kMaxKinds * (size_t)cm->world));
    (void)c;
    return 0;
}

int gpd_comm_init(gpd_ctx* c, const char id[GPD_NCCL_ID_BYTES], int32_t rank, int32_t world, gpd_comm** out) {
    GPD_CHECK(c && id && out, "null argument");
    *out = nullptr;
    GPD_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank / world size");
    NcclApi* api = nccl_api();
    GPD_CHECK(api, "libnccl.so.2 cannot be loaded: the sharded entry points need NCCL (there is no fallback)");
    GPD_CUDA(cudaSetDevice(c->device));
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    gpd_comm* cm = new gpd_comm();
    cm->ctx = c;
    cm->rank = rank;
    cm->world = world;
    cm->owned = true;
    ncclResult_t r = api->CommInitRank(&cm->comm, world, uid, rank);
    if (r != ncclSuccess) {
        set_error(std::string("ncclCommInitRank: ") + api->GetErrorString(r));
        delete cm;
        return 5;
    }
    if (int rc = comm_alloc(c, cm)) {
        gpd_comm_destroy(cm);
        return rc;
    }
    *out = cm;
    return 0;
}

int gpd_comm_adopt(gpd_ctx* c, void* nccl_comm, int32_t rank, int32_t world, gpd_comm** out) {
    GPD_CHECK(c && nccl_comm && out, "null argument");
    *out = nullptr;
    GPD_CHECK(world >= 1 && rank >= 0 && rank < world, "bad rank / world size");
    GPD_CHECK(nccl_api(), "libnccl.so.2 cannot be loaded: the sharded entry points need NCCL (there is no fallback)");
    GPD_CUDA(cudaSetDevice(c->device));
    gpd_comm* cm = new gpd_comm();
    cm->ctx = c;
    cm->comm = (ncclComm_t)nccl_comm;
    cm->rank = rank;
    cm->world = world;
    if (int rc = comm_alloc(c, cm)) {
        gpd_comm_destroy(cm);
        return rc;
    }
    *out = cm;
    return 0;
}

void gpd_comm_destroy(gpd_comm* cm) {
    if (!cm) return;
    cudaSetDevice(cm->ctx->device);
    cudaStreamSynchronize(cm->ctx->stream);
    NcclApi* api = nccl_api();
    if (cm->owned && cm->comm && api) api->CommDestroy(cm->comm);
    cudaFree(cm->send_dev);
    cudaFree(cm->recv_dev);
    delete cm;
}

// The data-parallel step of SURVEY 8(e) / BASELINE configs[4]: this rank scores ITS contiguous block of the candidates
// (N_local rows starting at global index idx_offset; N_local may be 0) exactly like gpd_score_multi, then the nkinds
// (max, global index) pairs of all ranks are all-gathered on the library stream (world x nkinds x 16 bytes, the only
// collective) and folded on the device with np.argmax semantics.  best_val / best_idx are the GLOBAL winners and are
// identical on every rank.
int gpd_score_sharded(gpd_ctx* c, gpd_comm* cm, gpd_model* const* models, int32_t M, const double* X, int32_t x_on_device,
                      int64_t N_local, int order, const int32_t* kinds, int32_t nkinds, const double* noise,
                      const double* pps, double* dc_out, int32_t dc_on_device, double* best_val, int64_t* best_idx,
                      int64_t idx_offset) {
    GPD_CHECK(c && cm && cm->ctx == c, "communicator does not belong to this context");
    GPD_CHECK(N_local >= 0, "negative shard size");
    GPD_CHECK(nkinds >= 1 && nkinds <= kMaxKinds, "1..5 criteria per call");
    NcclApi* api = nccl_api();
    GPD_CHECK(api, "libnccl.so.2 cannot be loaded");
    GPD_CUDA(cudaSetDevice(c->device));
    // A rank whose local scoring fails still takes part in the collective (as an empty shard) so that its peers do not
    // hang in the all-gather; it reports its own error afterwards.
    int local_rc = 0;
    std::string local_err;
    if (N_local > 0) {
        local_rc = score_impl(c, models, M, X, x_on_device != 0, N_local, order, kinds, nkinds, noise, pps, dc_out,
                              dc_on_device != 0, nullptr, nullptr, idx_offset, false);
        if (local_rc) local_err = g_error;
    }
    cudaStream_t s = c->stream;
    const size_t bytes = best_pair_bytes() * (size_t)nkinds;
    {
        Span sp(c, F_COMM, 2, 0);
        GPD_TRY(launch_best_pack(s, c->best_val_dev, c->best_idx_dev, nkinds, N_local == 0 || local_rc != 0, cm->send_dev));
        GPD_NCCL(api, api->AllGather(cm->send_dev, cm->recv_dev, bytes, ncclUint8, cm->comm, s));
        GPD_TRY(launch_best_reduce(s, cm->recv_dev, cm->world, nkinds, c->best_val_dev, c->best_idx_dev));
    }
    double* bv = (double*)c->result_pinned;
    long long* bi = (long long*)(bv + kMaxKinds);
    GPD_CUDA(cudaMemcpyAsync(bv, c->best_val_dev, sizeof(double) * nkinds, cudaMemcpyDeviceToHost, s));
    GPD_CUDA(cudaMemcpyAsync(bi, c->best_idx_dev, sizeof(long long) * nkinds, cudaMemcpyDeviceToHost, s));
    GPD_TRY(singular_fetch(c));
    GPD_CUDA(cudaStreamSynchronize(s));
    if (local_rc) {
        set_error(local_err);
        return local_rc;
    }
    GPD_TRY(singular_rc(c));
    for (int k = 0; k < nkinds; ++k) {
        if (best_val) best_val[k] = bv[k];
        if (best_idx) best_idx[k] = bi[k];
    }
    return 0;
}

/* debug: bytes of the scratch red zones (option "scratch_redzones") that were overwritten since the last planning */
int gpd_debug_check_scratch(gpd_ctx* c, int64_t* zones_out, int64_t* bad_bytes_out) {
    GPD_CHECK(c && zones_out && bad_bytes_out, "null argument");
    GPD_CUDA(cudaSetDevice(c->device));
    GPD_CUDA(cudaStreamSynchronize(c->stream));
    int64_t bad = 0;
    unsigned char host[kRedZone];
    if (c->redzone_selftest && !c->zones.empty()) {      // the detector must see a single damaged byte
        GPD_CUDA(cudaMemset((char*)c->scratch.p + c->zones.back() + 17, 0x00, 1));
        c->redzone_selftest = 0;
    }
    for (size_t off : c->zones) {
        if (off + kRedZone > c->scratch.cap) continue;
        GPD_CUDA(cudaMemcpy(host, (const char*)c->scratch.p + off, kRedZone, cudaMemcpyDeviceToHost));
        for (size_t k = 0; k < kRedZone; ++k) bad += host[k] != 0xCB;
    }
    *zones_out = (int64_t)c->zones.size();
    *bad_bytes_out = bad;
    return 0;
}

// ---- layout self-description --------------------------------------------------------------------
int32_t gpd_layout_a_index(int32_t ms, int32_t ks, int32_t lane) { return a_index(ms, ks, lane); }
int32_t gpd_layout_a_row(int32_t ms, int32_t lane) { return a_row(ms, lane); }
int32_t gpd_layout_a_col(int32_t ks, int32_t lane) { return a_col(ks, lane); }
int32_t gpd_layout_b_index(int32_t k, int32_t t) { return b_index(k, t); }
int64_t gpd_layout_tile_offset(int32_t nblk, int32_t I, int32_t kb, int32_t backward) {
    const int kfirst = tile_run_kfirst(I, backward);
    const int len = tile_run_len(nblk, I, backward);
    if (kb < kfirst || kb >= kfirst + len) return -1;
    return (tile_run_base(nblk, I, backward) + (kb - kfirst)) * (long long)kTileElems;
}
int32_t gpd_binary_combo_weight(int32_t n_binary, int32_t v) { return binary_combo_weight(n_binary, v); }

}  // extern "C"
