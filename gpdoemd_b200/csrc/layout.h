// Shared host/device index arithmetic for the DMMA (mma.sync m8n8k4 f64) operand layouts.
//
// The triangular contraction  V = L^-1 K*  is tiled as  O[128 x 128] += A[128 x 16] * B[16 x 128]
// chunks.  A (the pre-inverted factor) is static per GP, so it is packed ONCE into exactly the
// order the tensor-core fragments are read in; B (kernel rows of a candidate tile) is written by
// the evaluation kernel with an XOR swizzle so that dense 1-D bulk copies land conflict-free.
//
//  A chunk (128 rows x 16 k, 2048 doubles):  element (row = ms*8 + lane/4, k = ks*4 + lane%4)
//      lives at  ((ms*2 + ks/2)*32 + lane)*2 + (ks&1)      -> one LDS.128 per lane feeds two k4 steps
//  B chunk (16 k x 128 cols, 2048 doubles):  element (k, t) lives at  k*128 + (t ^ ((k&3)<<2))
//      -> the 16 lanes of a half warp (k = lane%4, t = t0 + lane/4) hit 16 distinct 8-byte banks.
//
// mma.m8n8k4 f64 fragment ownership (PTX ISA): A[row = lane/4][k = lane%4], B[k = lane%4][col = lane/4],
// C/D[row = lane/4][col = 2*(lane%4) + {0,1}].
#pragma once

#if defined(__CUDACC__)
#define GPD_HD __host__ __device__ __forceinline__
#else
#define GPD_HD inline
#endif

namespace gpd {

constexpr int kBlk = 128;          // rows per block of the factor / columns per candidate tile
constexpr int kChunkK = 16;        // k extent of one pipeline chunk
constexpr int kChunkElems = kBlk * kChunkK;      // 2048 doubles = 16 KiB, for A and for B
constexpr int kChunksPerBlk = kBlk / kChunkK;    // 8
constexpr int kTileElems = kBlk * kBlk;          // one packed 128x128 block of the factor

GPD_HD int a_index(int ms, int ks, int lane) { return (((ms * 2 + (ks >> 1)) * 32 + lane) << 1) + (ks & 1); }
GPD_HD int a_row(int ms, int lane) { return ms * 8 + (lane >> 2); }
GPD_HD int a_col(int ks, int lane) { return ks * 4 + (lane & 3); }
GPD_HD int b_index(int k, int t) { return k * kBlk + (t ^ ((k & 3) << 2)); }

// Packed factor: row block I owns a run of 128x128 tiles, one per k block it needs.
//   forward  (V = Linv * R,   Linv lower):  k blocks 0..I      ; diagonal tile is the LAST of the run
//   backward (B = Linv^T * V, upper)     :  k blocks I..nblk-1 ; diagonal tile is the FIRST of the run
GPD_HD long long tile_run_base(int nblk, int I, int backward) {
    return backward ? ((long long)I * nblk - (long long)I * (I - 1) / 2) : ((long long)I * (I + 1) / 2);
}
GPD_HD int tile_run_len(int nblk, int I, int backward) { return backward ? (nblk - I) : (I + 1); }
GPD_HD int tile_run_kfirst(int I, int backward) { return backward ? I : 0; }
GPD_HD long long packed_tiles(int nblk) { return (long long)nblk * (nblk + 1) / 2; }

// Combination index weight of binary variable v among nb binary variables, i.e. the row order of
// np.vstack([ravel(g) for g in np.meshgrid(*[[-1,1]]*nb)]).T  (GPdoemd/utils.py:38-39): meshgrid's
// default 'xy' indexing swaps the first two axes.
GPD_HD int binary_combo_weight(int nb, int v) {
    if (nb == 1) return 1;
    int axis = (v == 0) ? 1 : (v == 1 ? 0 : v);
    return 1 << (nb - 1 - axis);
}

}  // namespace gpd
