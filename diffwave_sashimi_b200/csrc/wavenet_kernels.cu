// WaveNet residual block, one fused kernel per layer.
//
// Reference: models/wavenet.py:82-121 (Residual_block.forward) + :160-162 (skip accumulation):
//   u = h + fc_t(emb)                      (bias BEFORE the zero padding of the dilated conv)
//   g = sum_{k in -1,0,1} W_k u[t + k d] + b  (+ mel features)       2C rows
//   o = tanh(g[:C]) * sigmoid(g[C:])
//   h' = (h + W_res o + b_res) * sqrt(1/2) ;  skip += W_skip o + b_skip
// The reference runs ~14 ATen launches per layer (3 of them re-normalising weights); here the
// three taps are three K-slabs of one implicit GEMM over a tile of all channels x TT time
// steps, the gate happens in registers and both 1x1 convs consume the gated tile from shared
// memory.  This file is the exact-fp32 SIMT implementation (parity mode).
#include "common.cuh"
#include "kernels.h"
#include "tile_gemm.cuh"

namespace dwb {

template <int TT>
__global__ void __launch_bounds__(MIX_THREADS)
wave_block_kernel(WaveBlockArgs a) {
    using G = TileGeom<TT>;
    extern __shared__ __align__(16) float smem[];
    const int C = a.C, S = a.S, L = a.L, d = a.dilation;
    float *Us = smem;                               // [3C][XS]  taps t-d, t, t+d
    float *Os = Us + (size_t)3 * C * G::XS;         // [C][XS]   gated activations
    const int tid = threadIdx.x, cg = tid % G::NCG, rg = tid / G::NCG;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    const float *hb = a.h + (size_t)b * C * L;
    const float *pt = a.part_t + (size_t)b * a.part_stride_b;

    for (int i = tid; i < 3 * C * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        const int tap = r / C, ch = r - tap * C;
        const int t = t0 + c + (tap - 1) * d;
        // outside [0, L) the conv sees zero padding, NOT the t-embedding bias (wavenet.py:91-95)
        Us[(size_t)r * G::XS + c] = (t >= 0 && t < L && t0 + c < L) ? hb[(size_t)ch * L + t] + pt[ch] : 0.f;
    }
    __syncthreads();

    for (int m0 = 0; m0 < C; m0 += G::CHUNK) {
        float acc_a[4][4], acc_b[4][4];
        zero_acc(acc_a);
        zero_acc(acc_b);
        tile_gemm_chunk<TT>(a.Wd_t, 2 * C, 3 * C, C, m0, Us, acc_a, rg, cg);
        tile_gemm_chunk<TT>(a.Wd_t + C, 2 * C, 3 * C, C, m0, Us, acc_b, rg, cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rg + G::NRG * i;
            if (m >= C) continue;
            const float ba = a.bd[m], bb = a.bd[C + m];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * cg + j;
                float ga = acc_a[i][j] + ba, gb = acc_b[i][j] + bb;
                if (a.cond && t0 + c < L) {
                    const float *cb = a.cond + (size_t)(a.cond_stride_b ? b : 0) * 2 * C * L;
                    ga += cb[(size_t)m * L + t0 + c];
                    gb += cb[(size_t)(C + m) * L + t0 + c];
                }
                Os[(size_t)m * G::XS + c] = tanhf(ga) * sigmoidf_(gb);
            }
        }
    }
    __syncthreads();

    const float rs = 0.70710678118654752440f;
    for (int m0 = 0; m0 < C; m0 += G::CHUNK) {
        float acc[4][4];
        zero_acc(acc);
        tile_gemm_chunk<TT>(a.Wr_t, C, C, C, m0, Os, acc, rg, cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rg + G::NRG * i;
            if (m >= C) continue;
            const float bv = a.br[m];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = t0 + 4 * cg + j;
                if (t < L) a.h_out[((size_t)b * C + m) * L + t] = (hb[(size_t)m * L + t] + acc[i][j] + bv) * rs;
            }
        }
    }
    for (int m0 = 0; m0 < S; m0 += G::CHUNK) {
        float acc[4][4];
        zero_acc(acc);
        tile_gemm_chunk<TT>(a.Ws_t, S, C, S, m0, Os, acc, rg, cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rg + G::NRG * i;
            if (m >= S) continue;
            const float bv = a.bs[m];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int t = t0 + 4 * cg + j;
                if (t < L) {
                    float *sp = a.skip + ((size_t)b * S + m) * L + t;
                    *sp = a.first ? acc[i][j] + bv : *sp + acc[i][j] + bv;
                }
            }
        }
    }
}

int wave_block_launch(const WaveBlockArgs &a, int B, cudaStream_t st) {
    constexpr int TT = 32;
    using G = TileGeom<TT>;
    const size_t sm = (size_t)4 * a.C * G::XS * sizeof(float);
    DWB_REQUIRE(sm <= 227 * 1024, DWB_ERR_UNSUPPORTED, "wavenet: res_channels=%d needs %zu B of shared memory", a.C, sm);
    if (sm > 48 * 1024)
        DWB_CUDA(cudaFuncSetAttribute(wave_block_kernel<TT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    dim3 grid(ceil_div(a.L, TT), B);
    wave_block_kernel<TT><<<grid, MIX_THREADS, sm, st>>>(a);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

}  // namespace dwb
