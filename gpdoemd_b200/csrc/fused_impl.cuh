// K1 + K2 fused for the first-order scoring path: per (GP, 128-candidate tile) one CTA evaluates the x-kernel rows,
// contracts them with the mean / derivative tables and writes the right-hand-side panel into a PRIVATE, L2-resident
// region (gridDim.x x n_pad x 128 doubles in total, rewritten for every item), then streams that panel back through the
// mbarrier ring of the FP64 tensor-core triangular contraction and leaves only ||v||^2 behind.
//
// Reference arithmetic: the same as eval_impl.cuh (GPy Stationary.K, gp_model.py:199,217) + kernels_tri.cu (GPy
// PosteriorExact._raw_predict `dtrtrs(woodbury_chol, Kx)`, gp_model.py:199).
//
// Why: with separate kernels the K* panel (8 n_pad bytes per candidate per GP: 3.3 GB per C2 step) makes a round trip
// through HBM and the evaluation kernel is bound by that write; here the panel never leaves L2,
// the scratch requirement drops from O(candidates x n) to O(SMs x n), and one launch replaces the head / rest
// two-stream choreography.  The evaluation's FP64 instructions still share the DFMA/DMMA pipe with the contraction,
// so the step gains the evaluation kernel's memory stall time, not its arithmetic.
//
// Work items carry a row-block range [I0, I1): the items of the last, partial wave of the persistent grid are cut
// into parts of equal triangular work so that all SMs stay busy (each part re-evaluates the k rows it needs; the
// part with I1 == nblk also publishes the mean contractions).  ||v||^2 partials go to slot part * 2 + row group.
#pragma once
#include "common.cuh"
#include "eval_impl.cuh"
#include "tri_core.cuh"

namespace gpd {

constexpr int kFusedThreads = 512 / kFusedCtasPerSm;
constexpr int kFusedRows = 256;        // training rows per staging block of the evaluation phase
constexpr int kFusedCps = (kFusedCtasPerSm == 1) ? 2 : 1;          // 16-k chunks per pipeline stage
constexpr int kFusedStages = (kFusedCtasPerSm == 1) ? 3 : 4;       // 3 x (32 + 32 KiB) per SM, or 4 x (16 + 8 KiB) per CTA
constexpr int kFusedAElems = kChunkElems * kFusedCps;              // 128 rows x 16 k per chunk
constexpr int kFusedBElems = kFusedTile * kChunkK * kFusedCps;     // 16 k x tile candidates per chunk

template <int NXB, int NCB>
struct FusedRec {
    static constexpr int RS = (NXB + NCB + 1 + 1) & ~1;      // [x (NXB) | C columns (NCB) | G0], even stride
};

// ---- evaluation phase: rows [0, kend) of the panel, optional mean contractions ------------------------------------
// CPT candidates per thread: 4 for the 5-column bucket (one warp spans the 128 candidates, 16 row groups), 2 for the
// 11-column bucket (two warps per row, 8 row groups: 4 x 11 accumulators would not fit the 128-register budget)
template <int KERN, int NXB, int NCB>
__device__ __forceinline__ void fused_eval(const GpView& g, const FusedArgs& a, int tile, int count, int kend,
                                           bool want_acc, double* __restrict__ panel, double* __restrict__ scratch,
                                           const double* __restrict__ sTab) {
    constexpr int CPT = (NCB <= 5) ? 4 : 2, TPG = kFusedTile / CPT, NW = kFusedThreads / TPG;
    constexpr int RS = FusedRec<NXB, NCB>::RS;
    const int tid = threadIdx.x, warp = tid / TPG, lane = tid % TPG;   // "warp" = row group, "lane" = position in it
    const int nx = g.nx, n_pad = g.n_pad, ncols = a.ncols;
    const double power = g.power_x;
    const int slot0 = tile * kFusedTile;
    double xs[CPT][NXB];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int slot = slot0 + eval_col<CPT, TPG>(lane, j);
        const bool valid = slot < count;
        const int cand = valid ? (g.idx ? g.idx[slot] : slot) : 0;
#pragma unroll
        for (int d = 0; d < NXB; ++d) {
            double v = 0.0;
            if (d < nx && valid) v = (a.X[(size_t)cand * a.ldx + g.x_active[d]] - g.xmin[d]) * g.xscale[d];
            xs[j][d] = v;
        }
    }
    double acc[CPT][NCB];
#pragma unroll
    for (int j = 0; j < CPT; ++j)
#pragma unroll
        for (int c = 0; c < NCB; ++c) acc[j][c] = 0.0;
    double* sRec = scratch;
    for (int i0 = 0; i0 < kend; i0 += kFusedRows) {
        const int rows = min(kFusedRows, kend - i0);
        __syncthreads();                                   // the previous block's records are consumed
        for (int k = tid; k < rows * NXB; k += kFusedThreads) {
            const int row = k / NXB, d = k % NXB;
            sRec[row * RS + d] = (d < nx) ? g.Xs[(size_t)(i0 + row) * nx + d] : 0.0;
        }
        for (int k = tid; k < rows * NCB; k += kFusedThreads) {
            const int c = k / rows, row = k % rows;
            sRec[row * RS + NXB + c] = (c < ncols) ? g.C[(size_t)c * n_pad + i0 + row] : 0.0;
        }
        for (int k = tid; k < rows; k += kFusedThreads) sRec[k * RS + NXB + NCB] = g.G[i0 + k];
        __syncthreads();
#pragma unroll 2
        for (int ii = warp; ii < rows; ii += NW) {
            const double* __restrict__ rec = sRec + ii * RS;
            double rv[RS];
#pragma unroll
            for (int d = 0; d < RS; d += 2) {
                const double2 v = *reinterpret_cast<const double2*>(rec + d);
                rv[d] = v.x;
                rv[d + 1] = v.y;
            }
            const int row = i0 + ii;
            const int sw = (row & 3) << 2;
            double kap[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) {
                double r2 = (KERN == 0 || KERN == 5) ? 0.0 : kTinyR2;   // sqrt_fast_pos never sees 0 (eval_impl.cuh)
#pragma unroll
                for (int d = 0; d < NXB; ++d) {
                    const double df = xs[j][d] - rv[d];
                    r2 = fma(df, df, r2);
                }
                kap[j] = kappa_of_r2<KERN>(r2, power, sTab);
            }
#pragma unroll
            for (int j = 0; j < CPT; ++j)
#pragma unroll
                for (int c = 0; c < NCB; ++c) acc[j][c] = fma(kap[j], rv[NXB + c], acc[j][c]);
            const double ga = rv[NXB + NCB];
            double* __restrict__ prow = panel + (size_t)row * kFusedTile;
#pragma unroll
            for (int jp = 0; jp < CPT / 2; ++jp)
                *reinterpret_cast<double2*>(prow + (eval_col<CPT, TPG>(lane, 2 * jp) ^ sw)) =
                    make_double2(kap[2 * jp] * ga, kap[2 * jp + 1] * ga);
        }
    }
    if (!want_acc) return;
    // ---- combine the 16 warps' partial contractions, then write the raw outputs ----
    __syncthreads();
    double* sRed = scratch;                                // [warp][column of the tile][c]
#pragma unroll
    for (int j = 0; j < CPT; ++j)
#pragma unroll
        for (int c = 0; c < NCB; ++c) sRed[((size_t)warp * kFusedTile + eval_col<CPT, TPG>(lane, j)) * NCB + c] = acc[j][c];
    __syncthreads();
    const RawOut& raw = a.raws[g.model];
    for (int o = tid; o < kFusedTile * NCB; o += kFusedThreads) {
        const int col = o / NCB, c = o % NCB;
        const int slot = slot0 + col;
        if (slot >= count || c >= ncols) continue;
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) s += sRed[((size_t)w * kFusedTile + col) * NCB + c];
        const int cand = g.idx ? g.idx[slot] : slot;
        const size_t orow = (size_t)g.out_e * a.chunk_n + cand;
        raw.cmu[orow * a.ldraw + c] = s;
        if (c == 0) raw.kss[orow] = g.kss;
    }
}

// ---- producer side of the contraction phase: one thread walks (I, q, chunk pair) of the item's row-block range ----
struct FusedStream {
    int j, jend, I, q, c, len;       // j: 16-k chunk index inside the packed factor
    const double* A;
    const double* B;
    int stage;
    uint32_t phase;
};

struct FusedSmem {
    double A[kFusedStages][kFusedAElems];
    double B[kFusedStages][kFusedBElems];
    unsigned long long full[kFusedStages];
    unsigned long long empty[kFusedStages];
    double tab[kKernTabSize];
};

__device__ __forceinline__ void fused_issue(FusedStream& ps, FusedSmem& sm) {
    constexpr uint32_t kABytes = kFusedAElems * sizeof(double), kBBytes = kFusedBElems * sizeof(double);
    if (ps.j >= ps.jend) return;
    mbar_wait(smem_u32(&sm.empty[ps.stage]), ps.phase ^ 1);
    const uint32_t fb = smem_u32(&sm.full[ps.stage]);
    mbar_expect_tx(fb, kABytes + kBBytes);
    bulk_g2s(smem_u32(sm.A[ps.stage]), ps.A + (size_t)ps.j * kChunkElems, kABytes, fb);
    bulk_g2s(smem_u32(sm.B[ps.stage]), ps.B + ((size_t)ps.q * kBlk + ps.c * kChunkK) * kFusedTile, kBBytes, fb);
    if (++ps.stage == kFusedStages) { ps.stage = 0; ps.phase ^= 1; }
    ps.j += kFusedCps;
    ps.c += kFusedCps;
    if (ps.c == kChunksPerBlk) {
        ps.c = 0;
        if (++ps.q == ps.len) {
            ps.q = 0;
            ++ps.I;
            ps.len = ps.I + 1;
        }
    }
}

template <int KERN, int NXB, int NCB>
__global__ void __launch_bounds__(kFusedThreads, kFusedCtasPerSm) fused_kernel(const FusedArgs args) {
    constexpr int WM = 2, WN = kFusedTile / 16, MB = 16 / WM, NB = kFusedTile / (8 * WN), kStages = kFusedStages;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FusedSmem& fs = *reinterpret_cast<FusedSmem*>(smem_raw);
    FusedSmem& sm = fs;
    double* scratch = &sm.A[0][0];                         // the ring is drained between items: evaluation scratch
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), WM * WN);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_kernel_tables<KERN>(fs.tab, threadIdx.x, kFusedThreads);
    __syncthreads();

    int stage = 0;
    uint32_t phase = 0;
    const bool is_producer = (threadIdx.x == 0);
    FusedStream ps;
    ps.stage = 0;
    ps.phase = 0;
    double* panel = args.priv + (size_t)blockIdx.x * args.priv_stride;
    const int wm = warp / WN, wn = warp % WN;
    const int lq = lane & 3, lr = lane >> 2;

    for (int w = blockIdx.x; w < args.nitems; w += gridDim.x) {
        const FItem it = args.items[w];
        const GpView& g = args.gps[it.gp];
        const int count = g.count ? *g.count : args.chunk_n;
        if (it.tile * kFusedTile >= count) continue;       // uniform over the CTA
        const int nblk = g.nblk;
        const int I0 = it.I0, I1 = it.I1 ? it.I1 : nblk;
        // ===== evaluation: panel rows [0, I1 * 128) =====
        fused_eval<KERN, NXB, NCB>(g, args, it.tile, count, I1 * kBlk, I1 == nblk, panel, scratch, fs.tab);
        // the panel was written through the generic proxy and is read back by bulk async copies
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();
        // ===== contraction: row blocks [I0, I1) =====
        if (is_producer) {
            ps.j = (int)tile_run_base(nblk, I0, 0) * kChunksPerBlk;
            ps.jend = (int)tile_run_base(nblk, I1, 0) * kChunksPerBlk;
            ps.I = I0;
            ps.q = 0;
            ps.c = 0;
            ps.len = I0 + 1;
            ps.A = g.LinvF;
            ps.B = panel;
#pragma unroll 1
            for (int k = 0; k < kStages - 1; ++k) fused_issue(ps, sm);
        }
        __syncwarp();
        double csq[NB][2];
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) csq[nb][0] = csq[nb][1] = 0.0;
#pragma unroll 1
        for (int I = I0; I < I1; ++I) {
            double acc[MB][NB][2];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) acc[mb][nb][0] = acc[mb][nb][1] = 0.0;
            const int len = I + 1;
            int mode_dense = 1;
            if (args.trim && I == nblk - 1) {
                const int left = g.n - I * kBlk - wm * (MB * 8);
                const int hi = left <= 0 ? 0 : (left >= MB * 8 ? MB : (left + 7) >> 3);
                mode_dense = hi == MB ? 1 : (hi == 0 ? 0 : 2 + hi);
            }
#pragma unroll 1
            for (int q = 0; q < len; ++q) {
                const bool diag = (q == len - 1);
#pragma unroll 1
                for (int c = 0; c < kChunksPerBlk; c += kFusedCps) {
                    mbar_wait(smem_u32(&sm.full[stage]), phase);
#pragma unroll
                    for (int cc = 0; cc < kFusedCps; ++cc) {
                        const double* sA = sm.A[stage] + cc * kChunkElems;
                        const double* sB = sm.B[stage] + cc * (kFusedTile * kChunkK);
                        int mode = mode_dense;
                        if (diag) {
                            const int rlo = GPD_BLK(wm, 0, WM, MB) * 8, rhi = GPD_BLK(wm, MB - 1, WM, MB) * 8 + 7;
                            const int klo = (c + cc) * kChunkK, khi = klo + kChunkK - 1;
                            if (klo > rhi) mode = 0;
                            else if (khi <= rlo) mode = 1;
                            else mode = 2;
                        }
                        if (mode == 1) compute_chunk<WM, WN, false, false, MB, kFusedTile, NB>(sA, sB, lane, wm, wn, c + cc, acc);
                        else if (mode == 2) compute_chunk<WM, WN, true, false, MB, kFusedTile, NB>(sA, sB, lane, wm, wn, c + cc, acc);
                        else if (mode >= 3) compute_trimmed<WM, WN, false, MB - 1, kFusedTile, NB>(mode - 2, sA, sB, lane, wm, wn, c + cc, acc);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&sm.empty[stage]));
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                    if (is_producer) fused_issue(ps, sm);
                    __syncwarp();
                }
            }
#pragma unroll
            for (int mb = 0; mb < MB; ++mb)
#pragma unroll
                for (int nb = 0; nb < NB; ++nb) {
                    const double v0 = acc[mb][nb][0], v1 = acc[mb][nb][1];
                    csq[nb][0] = fma(v0, v0, csq[nb][0]);
                    csq[nb][1] = fma(v1, v1, csq[nb][1]);
                }
        }
        {
            // ||v||^2 partial of this warp's rows (see tri_kernel): slot = part * WM + row group
            double mine = 0.0;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    double s = csq[nb][j];
                    s += __shfl_xor_sync(0xffffffffu, s, 4);
                    s += __shfl_xor_sync(0xffffffffu, s, 8);
                    s += __shfl_xor_sync(0xffffffffu, s, 16);
                    if (lr == nb * 2 + j) mine = s;
                }
            if (lr < 2 * NB) {
                const int t = (wn * NB + (lr >> 1)) * 8 + 2 * lq + (lr & 1);
                const int slot = it.tile * kFusedTile + t;
                if (slot < count) {
                    const int cand = g.idx ? g.idx[slot] : slot;
                    const RawOut& raw = args.raws[g.model];
                    raw.gram[((size_t)g.out_e * args.chunk_n + cand) * args.ng + it.part * WM + wm] = mine;
                }
            }
        }
        __syncthreads();      // every warp is done with the ring (next: evaluation scratch) and with this item's panel
    }
}

template <int KERN, int NXB, int NCB>
static int fused_launch_one(cudaStream_t s, int grid, const FusedArgs& args) {
    static bool attr_done_dev[64] = {};
    int dev = 0;
    GPD_CUDA(cudaGetDevice(&dev));
    if (!attr_done_dev[dev & 63]) {
        GPD_CUDA(cudaFuncSetAttribute(fused_kernel<KERN, NXB, NCB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(FusedSmem)));
        attr_done_dev[dev & 63] = true;
    }
    fused_kernel<KERN, NXB, NCB><<<grid, kFusedThreads, sizeof(FusedSmem), s>>>(args);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// buckets: exact nx = 3 or padded to 8; 5 columns (P <= 4) or 11 (P <= 10).  -1: no fused variant (caller falls back)
template <int KERN>
static int fused_dispatch(cudaStream_t s, int grid, int nx, int ncols, const FusedArgs& args) {
    if (ncols <= 5) {
        if (nx == 3) return fused_launch_one<KERN, 3, 5>(s, grid, args);
        return fused_launch_one<KERN, 8, 5>(s, grid, args);
    }
    if (nx == 3) return fused_launch_one<KERN, 3, 11>(s, grid, args);
    return fused_launch_one<KERN, 8, 11>(s, grid, args);
}

#define GPD_DEFINE_FUSED_TU(K)                                                                       \
    int launch_fused_k##K(cudaStream_t s, int grid, int nx, int ncols, const FusedArgs& args) {      \
        return fused_dispatch<K>(s, grid, nx, ncols, args);                                          \
    }

}  // namespace gpd
