// fp32 SIMT tile GEMM used by the channel-mixing kernels (1x1 convs, dilated-conv taps).
//
//   out[m, c] = sum_k Wt[k, m] * Xs[k, c]        m < M (output channels), c < TT (time columns)
//
// Wt is the folded weight stored K-major-rows ([K][M], i.e. transposed at finalize) so that the
// 32/64 row groups of a warp read consecutive floats; Xs is the activation tile in shared
// memory ([K][TT + pad]).  256 threads = NCG column groups (4 consecutive columns each) x NRG row
// groups; one call handles a "chunk" of NRG*4 output rows: thread (rg, cg) owns rows
// m0 + rg + NRG*i (i < 4) and columns 4cg..4cg+3.
//
// This is the exact-fp32 baseline path (parity mode).  The tensor-core path replaces the
// inner product only; tiling, prologues and epilogues are shared.
#pragma once
#include "common.cuh"

namespace dwb {

constexpr int MIX_THREADS = 256;

template <int TT>
struct TileGeom {
    static constexpr int NCG = TT / 4;
    static constexpr int NRG = MIX_THREADS / NCG;
    static constexpr int CHUNK = NRG * 4;       // output rows per chunk
    static constexpr int XS = TT + 4;           // smem row stride (floats), keeps float4 alignment
};

// acc[i][j] += sum_k Wt[k*ldw + m0 + rg + NRG*i] * Xs[k*XS + 4*cg + j]
template <int TT>
__device__ __forceinline__ void tile_gemm_chunk(const float *__restrict__ Wt, int ldw, int K, int M, int m0,
                                                const float *Xs, float (&acc)[4][4], int rg, int cg) {
    using G = TileGeom<TT>;
    const float *xp = Xs + 4 * cg;
    bool ok[4];
    const float *wp[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + rg + G::NRG * i;
        ok[i] = m < M;
        wp[i] = Wt + (ok[i] ? m : 0);
    }
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        const float4 xv = *reinterpret_cast<const float4 *>(xp + (size_t)k * G::XS);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float w = ok[i] ? __ldg(wp[i] + (size_t)k * ldw) : 0.f;
            acc[i][0] = fmaf(w, xv.x, acc[i][0]);
            acc[i][1] = fmaf(w, xv.y, acc[i][1]);
            acc[i][2] = fmaf(w, xv.z, acc[i][2]);
            acc[i][3] = fmaf(w, xv.w, acc[i][3]);
        }
    }
}

__device__ __forceinline__ void zero_acc(float (&acc)[4][4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

// Column statistics over `rows` rows of a smem tile: (mean, rstd) with biased variance, no eps
// (TransposedLN, models/sashimi.py:17-19).  All 256 threads participate; result in stat_s[c*2..].
// scratch: >= 2 * (MIX_THREADS/TT) * TT floats.
template <int TT>
__device__ __forceinline__ void tile_col_stats(const float *Xs, int rows, float *scratch, float *stat_s, int tid) {
    using G = TileGeom<TT>;
    constexpr int NP = MIX_THREADS / TT;
    const int c = tid % TT, part = tid / TT;
    float sum = 0.f;
    for (int r = part; r < rows; r += NP) sum += Xs[(size_t)r * G::XS + c];
    scratch[part * TT + c] = sum;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int p = 0; p < NP; ++p) mean += scratch[p * TT + c];
    mean /= (float)rows;
    float var = 0.f;
    for (int r = part; r < rows; r += NP) {
        const float d = Xs[(size_t)r * G::XS + c] - mean;
        var = fmaf(d, d, var);
    }
    scratch[NP * TT + part * TT + c] = var;
    __syncthreads();
    if (part == 0) {
        float v = 0.f;
#pragma unroll
        for (int p = 0; p < NP; ++p) v += scratch[NP * TT + p * TT + c];
        v /= (float)rows;
        stat_s[2 * c] = mean;
        stat_s[2 * c + 1] = 1.0f / sqrtf(v);
    }
    __syncthreads();
}

}  // namespace dwb
