// Device building blocks of the FP64 tensor-core triangular contraction, shared by tri_kernel (kernels_tri.cu) and
// the fused evaluation + contraction kernel (fused_impl.cuh): PTX wrappers (mbarrier, bulk async copy, mma.sync
// m8n8k4 f64), the pipeline-stage shared-memory layout and the per-chunk warp-tile update with the zero sub-blocks of
// the diagonal tiles skipped.  See kernels_tri.cu for the design notes.
#pragma once
#include <cuda_runtime.h>

#include "common.cuh"

namespace gpd {

#ifndef GPD_TRI_CYCLIC
#define GPD_TRI_CYCLIC 0
#endif
// 8-row block owned by (row group wm, slot mb): blocked (wm*MB + mb) or cyclic (mb*WM + wm)
#define GPD_BLK(wm, mb, WM, MB) (GPD_TRI_CYCLIC ? ((mb) * (WM) + (wm)) : ((wm) * (MB) + (mb)))

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void consumer_bar(int nthreads) {
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// ---------------------------------------------------------------------------------------------
// the persistent triangular-GEMM kernel
// ---------------------------------------------------------------------------------------------
// CPS = 16-k chunks per pipeline stage (consecutive chunks are contiguous in both operands):
//   CPS 1: 4 stages x (16 KiB A + 16 KiB B)      CPS 2: 3 stages x (32 KiB A + 32 KiB B)
template <int CPS>
struct TriCfg {
#ifndef GPD_TRI_STAGES1
#define GPD_TRI_STAGES1 4
#endif
    static constexpr int kStages = (CPS == 1) ? GPD_TRI_STAGES1 : 3;
    static constexpr int kStageElems = kChunkElems * CPS;
    static constexpr uint32_t kStageBytes = kStageElems * sizeof(double);
};

template <int CPS>
struct TriSmem {
    double A[TriCfg<CPS>::kStages][TriCfg<CPS>::kStageElems];
    double B[TriCfg<CPS>::kStages][TriCfg<CPS>::kStageElems];
    unsigned long long full[TriCfg<CPS>::kStages];
    unsigned long long empty[TriCfg<CPS>::kStages];
};

struct TriArgs {
    const GpView* gps;
    const Item* items;
    int nitems;
    int nwork_rhs;        // RHS kinds processed per item (a = 0 .. nwork_rhs-1)
    int in_stride;        // RHS kinds per tile in the input panel buffer
    int out_stride;       // RHS kinds per tile in the output panel buffer
    int backward;         // 0: V = Linv * R ; 1: beta = Linv^T * V
    const double* panels_in;
    double* panels_out;   // may be null
    const RawOut* raws;   // non-null: write column sums of squares into raw.gram[..][0]
    int ng;               // row length of raw.gram
    int chunk_n;
    int trim;             // skip the all-padding 8-row blocks of the last row block (forward, 16-warp variant)
};

// One k4 step for the 8-row blocks [LO, HI) of this warp's row group.
template <int MB, int NB, int LO, int HI>
__device__ __forceinline__ void kstep(double (&acc)[MB][NB][2], const double (&a)[MB], const double (&b)[NB]) {
#pragma unroll
    for (int mb = LO; mb < HI; ++mb)
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[mb], b[nb]);
}
// Diagonal tiles: the zero 8 x 4 sub-blocks of the triangular factor are skipped with REAL (warp-uniform)
// branches.  A predicated-off DMMA still occupies the FP64 tensor pipe for its full issue interval
// (measured: profiles/r01_tri_c2_ncu.md), so predication alone saves nothing.
//   forward  (lower): blocks mb >= v are non-zero  -> run [v, MB)
//   backward (upper): blocks mb <  v are non-zero  -> run [0, v)
template <int MB, int NB, bool BWD, int C>
__device__ __forceinline__ void kstep_bounded(int v, double (&acc)[MB][NB][2], const double (&a)[MB],
                                              const double (&b)[NB]) {
    if (v == C) {
        if (BWD) kstep<MB, NB, 0, C>(acc, a, b);
        else kstep<MB, NB, C, MB>(acc, a, b);
    } else if constexpr (C < MB) {
        kstep_bounded<MB, NB, BWD, C + 1>(v, acc, a, b);
    }
}

// BLD = row length of the B chunk in shared memory (candidates per tile: 128, or 64 in the two-CTA fused kernel)
template <int WM, int WN, bool DIAG, bool BWD, int HI = 16 / WM, int BLD = kBlk, int NBT = 16 / WN>
__device__ __forceinline__ void compute_chunk(const double* __restrict__ sA, const double* __restrict__ sB, int lane,
                                              int wm, int wn, int chunk_in_blk,
                                              double (&acc)[16 / WM][NBT][2]) {
    constexpr int MB = 16 / WM, NB = NBT;
    const int lq = lane & 3, lr = lane >> 2;
    if (DIAG) {
        // whole chunk zero for this warp's rows?  8-row blocks mb*WM + wm (cyclic), k in [16c, 16c+16)
        const int rlo = GPD_BLK(wm, 0, WM, MB) * 8, rhi = GPD_BLK(wm, MB - 1, WM, MB) * 8 + 7, klo = chunk_in_blk * kChunkK, khi = klo + kChunkK - 1;
        if (BWD ? (khi < rlo) : (klo > rhi)) return;
    }
    const double* __restrict__ pa = sA + lane * 2;
    // b_index(k, t), k = ks*4 + lq, t = (wn*NB + nb)*8 + lr: the swizzle (k & 3) << 2 == lq << 2 flips bits 2..3 of t,
    // i.e. it depends on lr and on the parity of nb only -> two lane-constant bases, even and odd nb
    static_assert(NB % 2 == 0, "NB must be even");
    const double* __restrict__ pb0 = sB + lq * BLD + ((wn * NB * 8 + lr) ^ (lq << 2));
    const double* __restrict__ pb1 = sB + lq * BLD + ((wn * NB * 8 + 8 + lr) ^ (lq << 2));
#pragma unroll
    for (int kp = 0; kp < 2; ++kp) {
        double2 a2[MB];
#pragma unroll
        for (int mb = 0; mb < MB; ++mb)
            if (mb < HI) a2[mb] = *reinterpret_cast<const double2*>(pa + (GPD_BLK(wm, mb, WM, MB) * 2 + kp) * 64);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int ks = kp * 2 + h;
            double a[MB], b[NB];
#pragma unroll
            for (int mb = 0; mb < MB; ++mb) a[mb] = (mb < HI) ? (h ? a2[mb].y : a2[mb].x) : 0.0;
#pragma unroll
            for (int nb = 0; nb < NB; ++nb) b[nb] = ((nb & 1) ? pb1 : pb0)[ks * 4 * BLD + (nb >> 1) * 16];
            if (DIAG) {
                const int k0 = chunk_in_blk * kChunkK + ks * 4;
                // block mb*WM + wm holds rows >= its first row; forward (lower) needs block >= k0/8,
                // backward (upper) block <= k0/8
                int v;
                if (GPD_TRI_CYCLIC) {
                    const int d = (k0 >> 3) - wm;
                    v = BWD ? (d < 0 ? 0 : d / WM + 1) : (d <= 0 ? 0 : (d + WM - 1) / WM);
                } else {
                    v = (k0 >> 3) - wm * MB + (BWD ? 1 : 0);
                    v = v < 0 ? 0 : v;
                }
                v = v > MB ? MB : v;
                kstep_bounded<MB, NB, BWD, 0>(v, acc, a, b);
            } else {
                kstep<MB, NB, 0, HI>(acc, a, b);
            }
        }
    }
}

// dense chunk restricted to the first hi (1 .. MB-1) 8-row blocks of the warp (padding rows trimmed, warp-uniform hi)
template <int WM, int WN, bool BWD, int HI, int BLD = kBlk, int NBT = 16 / WN>
__device__ __forceinline__ void compute_trimmed(int hi, const double* __restrict__ sA, const double* __restrict__ sB, int lane,
                                                int wm, int wn, int chunk_in_blk, double (&acc)[16 / WM][NBT][2]) {
    if (hi == HI) compute_chunk<WM, WN, false, BWD, HI, BLD, NBT>(sA, sB, lane, wm, wn, chunk_in_blk, acc);
    else if constexpr (HI > 1) compute_trimmed<WM, WN, BWD, HI - 1, BLD, NBT>(hi, sA, sB, lane, wm, wn, chunk_in_blk, acc);
}

}  // namespace gpd
