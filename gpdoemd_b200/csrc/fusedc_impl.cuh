// K1 + K2 fused CONCURRENTLY (option "fused" = 2): the persistent triangular-contraction CTA (16 DMMA consumer warps,
// one elected thread issuing bulk copies into the mbarrier ring, exactly tri_kernel's structure) carries one more
// warpgroup of 4 PRODUCER warps that evaluate the kernel rows of the CTA's NEXT work item while the consumers contract
// the current one.
//
// Reference arithmetic: the same as eval_impl.cuh (GPy Stationary.K, gp_model.py:199,217) + kernels_tri.cu (GPy
// PosteriorExact._raw_predict `dtrtrs(woodbury_chol, Kx)`, gp_model.py:199).
//
// Why this and not the phase-serial fused_kernel (fused_impl.cuh, measured slower): DFMA and DMMA share one FP64 pipe, so
// the evaluation's arithmetic cannot be hidden -- but the contraction leaves 3 % (n = 4000) to 6 % (n = 500) of the pipe's
// issue slots idle (barrier skew in diagonal tiles, stage hand-over), and the separate evaluation kernel keeps the pipe
// only 58 - 73 % busy while it runs alone.  A producer warp per SM sub-partition feeds its DFMAs into exactly those idle
// slots; the evaluation then costs its pipe time (C2: 0.42 of 0.81 ms) minus the slots that were idle anyway.
//
// Data flow per CTA: two private panel buffers (n_pad x 128 doubles each, L2 resident).  Producers: wait until buffer
// b = seq & 1 is free (`pfree[b]`, 16 consumer-warp arrivals at the end of the item that used it), evaluate rows
// [0, I1 * 128) of item seq into it -- panel[i][t ^ swizzle] = kx_i(x_t) G0_i, mean / derivative contractions to the raw
// outputs when the item covers the last row block -- then `fence.proxy.async` (the panel is read back by bulk async
// copies) and arrive on `ready[b]`.  The copy-issuing thread waits on `ready[b]` when its chunk stream enters item seq.
// Registers: 640 threads are launched at 96 registers; the consumer warpgroups raise their limit to 104 and the
// producer warpgroup drops to 64 (setmaxnreg), 512 * 104 + 128 * 64 = 640 * 96: the increase is funded by the decrease inside the CTA pool.
#pragma once
#include "common.cuh"
#include "eval_impl.cuh"
#include "tri_core.cuh"

namespace gpd {

constexpr int kFcConsumers = 512, kFcProducers = 128, kFcThreads = kFcConsumers + kFcProducers;
constexpr int kFcRows = 128;           // training rows per producer staging block
constexpr int kFcStages = 3, kFcCps = 2;
constexpr int kFcStageElems = kChunkElems * kFcCps;

template <int NXB, int NCB>
struct FcRec {
    static constexpr int RS = (NXB + NCB + 1 + 1) & ~1;      // [x (NXB) | C columns (NCB) | G0], even stride
};

template <int NXB, int NCB>
struct FcSmem {
    double A[kFcStages][kFcStageElems];
    double B[kFcStages][kFcStageElems];
    unsigned long long full[kFcStages];
    unsigned long long empty[kFcStages];
    unsigned long long ready[2];
    unsigned long long pfree[2];
    double tab[kKernTabSize];
    double rec[kFcRows * FcRec<NXB, NCB>::RS > kBlk * NCB ? kFcRows * FcRec<NXB, NCB>::RS : kBlk * NCB];
};

// polling wait with back-off: the producers are idle for most of an item and must not spend issue slots on the poll
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
        __nanosleep(1000);
    }
}
__device__ __forceinline__ void producer_bar() { asm volatile("bar.sync 2, %0;" ::"n"(kFcProducers) : "memory"); }

// ---- producer warpgroup: panel rows [0, kend) of one item, optional mean contractions --------------------------------
template <int KERN, int NXB, int NCB>
__device__ __forceinline__ void fc_produce(const GpView& g, const FusedArgs& a, int tile, int count, int kend, bool want_acc,
                                           double* __restrict__ panel, double* __restrict__ sRec,
                                           const double* __restrict__ sTab, int ptid) {
    constexpr int CPT = (NCB <= 5) ? 2 : 1, TPG = kBlk / CPT, RG = kFcProducers / TPG;
    constexpr int RS = FcRec<NXB, NCB>::RS;
    const int rg = ptid / TPG, s = ptid % TPG;
    const int nx = g.nx, n_pad = g.n_pad, ncols = a.ncols;
    const double power = g.power_x;
    const int slot0 = tile * kBlk;
    double xs[CPT][NXB];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int slot = slot0 + eval_col<CPT, TPG>(s, j);
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
    for (int i0 = 0; i0 < kend; i0 += kFcRows) {
        producer_bar();                                    // the previous block's records are consumed
        for (int k = ptid; k < kFcRows * NXB; k += kFcProducers) {
            const int row = k / NXB, d = k % NXB;
            sRec[row * RS + d] = (d < nx) ? g.Xs[(size_t)(i0 + row) * nx + d] : 0.0;
        }
        for (int k = ptid; k < kFcRows * NCB; k += kFcProducers) {
            const int c = k / kFcRows, row = k % kFcRows;
            sRec[row * RS + NXB + c] = (c < ncols) ? g.C[(size_t)c * n_pad + i0 + row] : 0.0;
        }
        for (int k = ptid; k < kFcRows; k += kFcProducers) sRec[k * RS + NXB + NCB] = g.G[i0 + k];
        producer_bar();
#pragma unroll 2
        for (int ii = rg; ii < kFcRows; ii += RG) {
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
            double* __restrict__ prow = panel + (size_t)row * kBlk;
            if (CPT == 2)
                *reinterpret_cast<double2*>(prow + (eval_col<CPT, TPG>(s, 0) ^ sw)) = make_double2(kap[0] * ga, kap[CPT - 1] * ga);
            else
                prow[s ^ sw] = kap[0] * ga;
        }
    }
    if (!want_acc) return;
    // ---- combine the row groups (group 0 accumulates), then write the raw contractions ----
    if (RG > 1) {
        producer_bar();
        if (rg == 1) {
#pragma unroll
            for (int j = 0; j < CPT; ++j)
#pragma unroll
                for (int c = 0; c < NCB; ++c) sRec[eval_col<CPT, TPG>(s, j) * NCB + c] = acc[j][c];
        }
        producer_bar();
        if (rg == 0) {
#pragma unroll
            for (int j = 0; j < CPT; ++j)
#pragma unroll
                for (int c = 0; c < NCB; ++c) acc[j][c] += sRec[eval_col<CPT, TPG>(s, j) * NCB + c];
        }
    }
    if (rg != 0) return;
    const RawOut& raw = a.raws[g.model];
#pragma unroll
    for (int j = 0; j < CPT; ++j) {
        const int slot = slot0 + eval_col<CPT, TPG>(s, j);
        if (slot >= count) continue;
        const int cand = g.idx ? g.idx[slot] : slot;
        const size_t orow = (size_t)g.out_e * a.chunk_n + cand;
#pragma unroll
        for (int c = 0; c < NCB; ++c)
            if (c < ncols) raw.cmu[orow * a.ldraw + c] = acc[j][c];
        raw.kss[orow] = g.kss;
    }
}

// ---- chunk stream of the copy-issuing thread: flat over (item, row block, k block, stage), as in tri_kernel ----------
struct FcStream {
    int w, j, total, I, q, c, len, seq;
    const double* A;
    const double* B;
    int stage;
    uint32_t phase;
};

template <int NXB, int NCB>
__device__ __forceinline__ void fc_seek(FcStream& ps, const FusedArgs& args, FcSmem<NXB, NCB>& sm, double* priv) {
    for (; ps.w < args.nitems; ps.w += gridDim.x) {
        const FItem it = args.items[ps.w];
        const GpView& g = args.gps[it.gp];
        const int count = g.count ? *g.count : args.chunk_n;
        if (it.tile * kBlk >= count) continue;
        const int I0 = it.I0, I1 = it.I1 ? it.I1 : g.nblk;
        ps.A = g.LinvF;
        ps.total = (int)tile_run_base(g.nblk, I1, 0) * kChunksPerBlk;
        ps.j = (int)tile_run_base(g.nblk, I0, 0) * kChunksPerBlk;
        ps.I = I0;
        ps.q = 0;
        ps.c = 0;
        ps.len = I0 + 1;
        const int buf = ps.seq & 1;
        mbar_wait(smem_u32(&sm.ready[buf]), (ps.seq >> 1) & 1);      // the producers have finished this item's panel
        ps.B = priv + (size_t)buf * args.priv_stride;
        ++ps.seq;
        return;
    }
}

template <int NXB, int NCB>
__device__ __forceinline__ void fc_issue(FcStream& ps, const FusedArgs& args, FcSmem<NXB, NCB>& sm, double* priv) {
    constexpr uint32_t kStageBytes = kFcStageElems * sizeof(double);
    if (ps.w >= args.nitems) return;
    mbar_wait(smem_u32(&sm.empty[ps.stage]), ps.phase ^ 1);
    const uint32_t fb = smem_u32(&sm.full[ps.stage]);
    mbar_expect_tx(fb, 2 * kStageBytes);
    bulk_g2s(smem_u32(sm.A[ps.stage]), ps.A + (size_t)ps.j * kChunkElems, kStageBytes, fb);
    bulk_g2s(smem_u32(sm.B[ps.stage]), ps.B + ((size_t)ps.q * kBlk + ps.c * kChunkK) * kBlk, kStageBytes, fb);
    if (++ps.stage == kFcStages) { ps.stage = 0; ps.phase ^= 1; }
    ps.j += kFcCps;
    ps.c += kFcCps;
    if (ps.c == kChunksPerBlk) {
        ps.c = 0;
        if (++ps.q == ps.len) {
            ps.q = 0;
            ++ps.I;
            ps.len = ps.I + 1;
        }
    }
    if (ps.j == ps.total) {
        ps.w += gridDim.x;
        fc_seek<NXB, NCB>(ps, args, sm, priv);
    }
}

template <int KERN, int NXB, int NCB>
__global__ void __launch_bounds__(kFcThreads, 1) fusedc_kernel(const FusedArgs args) {
    constexpr int WM = 2, WN = 8, MB = 16 / WM, NB = 16 / WN;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FcSmem<NXB, NCB>& sm = *reinterpret_cast<FcSmem<NXB, NCB>*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kFcStages; ++s) {
            mbar_init(smem_u32(&sm.full[s]), 1);
            mbar_init(smem_u32(&sm.empty[s]), WM * WN);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&sm.ready[b]), kFcProducers);
            mbar_init(smem_u32(&sm.pfree[b]), WM * WN);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    load_kernel_tables<KERN>(sm.tab, threadIdx.x, kFcThreads);
    __syncthreads();                                       // the only CTA-wide barrier
    double* priv = args.priv + (size_t)blockIdx.x * 2 * args.priv_stride;

    if (warp >= kFcConsumers / 32) {
        // =================== producer warpgroup ===================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        const int ptid = threadIdx.x - kFcConsumers;
        int seq = 0;
        for (int w = blockIdx.x; w < args.nitems; w += gridDim.x) {
            const FItem it = args.items[w];
            const GpView& g = args.gps[it.gp];
            const int count = g.count ? *g.count : args.chunk_n;
            if (it.tile * kBlk >= count) continue;         // the consumers skip the same items
            const int I1 = it.I1 ? it.I1 : g.nblk;
            const int buf = seq & 1;
            mbar_wait_sleep(smem_u32(&sm.pfree[buf]), ((seq >> 1) & 1) ^ 1);   // the item that used this buffer is contracted
            fc_produce<KERN, NXB, NCB>(g, args, it.tile, count, I1 * kBlk, I1 == g.nblk, priv + (size_t)buf * args.priv_stride,
                                       sm.rec, sm.tab, ptid);
            asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes -> bulk async copies
            mbar_arrive(smem_u32(&sm.ready[buf]));
            ++seq;
        }
        return;
    }

    // =================== consumer warpgroups ===================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    int stage = 0;
    uint32_t phase = 0;
    const bool is_issuer = (threadIdx.x == 0);
    FcStream ps;
    ps.w = blockIdx.x;
    ps.stage = 0;
    ps.phase = 0;
    ps.seq = 0;
    if (is_issuer) {
        fc_seek<NXB, NCB>(ps, args, sm, priv);
#pragma unroll 1
        for (int k = 0; k < kFcStages - 1; ++k) fc_issue<NXB, NCB>(ps, args, sm, priv);
    }
    __syncwarp();
    const int wm = warp / WN, wn = warp % WN;
    const int lq = lane & 3, lr = lane >> 2;
    int seq = 0;
    for (int w = blockIdx.x; w < args.nitems; w += gridDim.x) {
        const FItem it = args.items[w];
        const GpView& g = args.gps[it.gp];
        const int count = g.count ? *g.count : args.chunk_n;
        if (it.tile * kBlk >= count) continue;
        const int nblk = g.nblk;
        const int I0 = it.I0, I1 = it.I1 ? it.I1 : nblk;
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
                for (int c = 0; c < kChunksPerBlk; c += kFcCps) {
                    mbar_wait(smem_u32(&sm.full[stage]), phase);
#pragma unroll
                    for (int cc = 0; cc < kFcCps; ++cc) {
                        const double* sA = sm.A[stage] + cc * kChunkElems;
                        const double* sB = sm.B[stage] + cc * kChunkElems;
                        int mode = mode_dense;
                        if (diag) {
                            const int rlo = GPD_BLK(wm, 0, WM, MB) * 8, rhi = GPD_BLK(wm, MB - 1, WM, MB) * 8 + 7;
                            const int klo = (c + cc) * kChunkK, khi = klo + kChunkK - 1;
                            if (klo > rhi) mode = 0;
                            else if (khi <= rlo) mode = 1;
                            else mode = 2;
                        }
                        if (mode == 1) compute_chunk<WM, WN, false, false>(sA, sB, lane, wm, wn, c + cc, acc);
                        else if (mode == 2) compute_chunk<WM, WN, true, false>(sA, sB, lane, wm, wn, c + cc, acc);
                        else if (mode >= 3) compute_trimmed<WM, WN, false, MB - 1>(mode - 2, sA, sB, lane, wm, wn, c + cc, acc);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(smem_u32(&sm.empty[stage]));
                    if (++stage == kFcStages) { stage = 0; phase ^= 1; }
                    if (is_issuer) fc_issue<NXB, NCB>(ps, args, sm, priv);
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
        // this warp has consumed every stage of the item: its panel buffer may be overwritten
        if (lane == 0) mbar_arrive(smem_u32(&sm.pfree[seq & 1]));
        ++seq;
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
                const int slot = it.tile * kBlk + t;
                if (slot < count) {
                    const int cand = g.idx ? g.idx[slot] : slot;
                    const RawOut& raw = args.raws[g.model];
                    raw.gram[((size_t)g.out_e * args.chunk_n + cand) * args.ng + it.part * WM + wm] = mine;
                }
            }
        }
    }
}

template <int KERN, int NXB, int NCB>
static int fusedc_launch_one(cudaStream_t s, int grid, const FusedArgs& args) {
    static bool attr_done_dev[64] = {};
    int dev = 0;
    GPD_CUDA(cudaGetDevice(&dev));
    if (!attr_done_dev[dev & 63]) {
        GPD_CUDA(cudaFuncSetAttribute(fusedc_kernel<KERN, NXB, NCB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)sizeof(FcSmem<NXB, NCB>)));
        attr_done_dev[dev & 63] = true;
    }
    fusedc_kernel<KERN, NXB, NCB><<<grid, kFcThreads, sizeof(FcSmem<NXB, NCB>), s>>>(args);
    GPD_CUDA(cudaGetLastError());
    return 0;
}

// buckets: exact nx = 3 or padded to 8; 5 columns (P <= 4) or 11 (P <= 10).  -1: no variant (caller falls back)
template <int KERN>
static int fusedc_dispatch(cudaStream_t s, int grid, int nx, int ncols, const FusedArgs& args) {
    if (ncols <= 5) {
        if (nx == 3) return fusedc_launch_one<KERN, 3, 5>(s, grid, args);
        return fusedc_launch_one<KERN, 8, 5>(s, grid, args);
    }
    if (nx == 3) return fusedc_launch_one<KERN, 3, 11>(s, grid, args);
    return fusedc_launch_one<KERN, 8, 11>(s, grid, args);
}

#define GPD_DEFINE_FUSEDC_TU(K)                                                                      \
    int launch_fusedc_k##K(cudaStream_t s, int grid, int nx, int ncols, const FusedArgs& args) {     \
        return fusedc_dispatch<K>(s, grid, nx, ncols, args);                                         \
    }

}  // namespace gpd
