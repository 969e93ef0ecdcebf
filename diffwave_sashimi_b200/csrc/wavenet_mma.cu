// WaveNet residual block on tensor cores (split-bf16 mma.sync, fp32 accumulate) — same fusion as
// wave_block_kernel (wavenet_kernels.cu): the three dilated taps are three K-slabs of one implicit
// GEMM (K = 3C) over a tile of all channels x TT time steps, the tanh*sigmoid gate is applied to
// the accumulator fragments in registers, and the residual / skip 1x1 convolutions consume the
// gated tile from shared memory.  Reference: models/wavenet.py:82-121, :160-162.
#include "common.cuh"
#include "kernels.h"
#include "mma_split.cuh"

namespace dwb {

template <int TT>
__global__ void __launch_bounds__(256)
wave_block_mma_kernel(WaveBlockArgs a) {
    constexpr int TTP = TT + 8, NT = TT / 8, NW = 8;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int C = a.C, S = a.S, L = a.L, d = a.dilation;
    __nv_bfloat16 *Uhi = reinterpret_cast<__nv_bfloat16 *>(smraw);      // [3C][TTP]
    __nv_bfloat16 *Ulo = Uhi + (size_t)3 * C * TTP;
    __nv_bfloat16 *Ohi = Ulo + (size_t)3 * C * TTP;                     // [C][TTP]
    __nv_bfloat16 *Olo = Ohi + (size_t)C * TTP;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    const float *hb = a.h + (size_t)b * C * L;
    const float *pt = a.part_t + (size_t)b * a.part_stride_b;

    // taps t-d, t, t+d of u = h + fc_t(emb); zero (not the bias) outside [0, L)   (wavenet.py:91-95)
    for (int i0 = tid; i0 < 3 * C * TT; i0 += 256 * 4) {
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256;
            const int r = i / TT, c = i - r * TT;
            const int tap = r / C, ch = r - tap * C;
            const int t = t0 + c + (tap - 1) * d;
            v[u] = (i < 3 * C * TT && t >= 0 && t < L && t0 + c < L) ? hb[(size_t)ch * L + t] + pt[ch] : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * 256;
            const int r = i / TT, c = i - r * TT;
            if (i < 3 * C * TT) split_store(Uhi, Ulo, (size_t)r * TTP + c, v[u]);
        }
    }
    __syncthreads();

    // dilated conv (K = 3C) + gate -> O
    for (int p = warp; p < C / 16; p += NW) {
        int tiles[2] = {p, p + C / 16};
        float acc[2][NT][4];
        zero3(acc);
        gemm_split_bf16<2, NT, TTP>(a.Wd_fh, a.Wd_fl, 3 * C / 16, tiles, Uhi, Ulo, 0, acc, lane);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int m = p * 16 + g + half * 8;
            const float ba = a.bd[m], bb = a.bd[C + m];
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                const int c = n * 8 + 2 * tq;
                float ga0 = acc[0][n][half * 2] + ba, ga1 = acc[0][n][half * 2 + 1] + ba;
                float gb0 = acc[1][n][half * 2] + bb, gb1 = acc[1][n][half * 2 + 1] + bb;
                if (a.cond) {
                    const float *cb = a.cond + (size_t)(a.cond_stride_b ? b : 0) * 2 * C * L;
                    if (t0 + c < L) { ga0 += cb[(size_t)m * L + t0 + c]; gb0 += cb[(size_t)(C + m) * L + t0 + c]; }
                    if (t0 + c + 1 < L) { ga1 += cb[(size_t)m * L + t0 + c + 1]; gb1 += cb[(size_t)(C + m) * L + t0 + c + 1]; }
                }
                split_store2(Ohi, Olo, (size_t)m * TTP + c, tanhf(ga0) * sigmoidf_(gb0), tanhf(ga1) * sigmoidf_(gb1));
            }
        }
    }
    __syncthreads();

    // residual 1x1: h' = (h + W_res o + b_res) sqrt(1/2)
    const float rs = 0.70710678118654752440f;
    for (int mt = warp; mt < C / 16; mt += NW) {
        int tiles[1] = {mt};
        float acc[1][NT][4];
        zero3(acc);
        gemm_split_bf16<1, NT, TTP>(a.Wr_fh, a.Wr_fl, C / 16, tiles, Ohi, Olo, 0, acc, lane);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int m = mt * 16 + g + half * 8;
            const float bv = a.br[m];
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int t = t0 + n * 8 + 2 * tq + j;
                    if (t < L) a.h_out[((size_t)b * C + m) * L + t] = (hb[(size_t)m * L + t] + acc[0][n][half * 2 + j] + bv) * rs;
                }
        }
    }
    // skip 1x1, accumulated in place
    for (int mt = warp; mt < S / 16; mt += NW) {
        int tiles[1] = {mt};
        float acc[1][NT][4];
        zero3(acc);
        gemm_split_bf16<1, NT, TTP>(a.Ws_fh, a.Ws_fl, C / 16, tiles, Ohi, Olo, 0, acc, lane);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int m = mt * 16 + g + half * 8;
            const float bv = a.bs[m];
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int t = t0 + n * 8 + 2 * tq + j;
                    if (t < L) {
                        float *sp = a.skip + ((size_t)b * S + m) * L + t;
                        *sp = a.first ? acc[0][n][half * 2 + j] + bv : *sp + acc[0][n][half * 2 + j] + bv;
                    }
                }
        }
    }
}

bool wave_mma_supported(int C, int S) { return C % 16 == 0 && S % 16 == 0 && C >= 32 && (size_t)8 * C * (32 + 8) * 2 <= 227 * 1024; }

template <int TT>
static int launch_wave(const WaveBlockArgs &a, int B, size_t sm, cudaStream_t st) {
    auto k = wave_block_mma_kernel<TT>;
    if (sm > 48 * 1024) DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    dim3 grid(ceil_div(a.L, TT), B);
    k<<<grid, 256, sm, st>>>(a);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

int wave_block_mma_launch(const WaveBlockArgs &a, int B, cudaStream_t st) {
    auto smem = [&](int TT) { return (size_t)8 * a.C * (TT + 8) * 2; };   // (3C + C) rows x 2 halves x bf16
    if (smem(64) <= 160 * 1024) return launch_wave<64>(a, B, smem(64), st);
    return launch_wave<32>(a, B, smem(32), st);
}

}  // namespace dwb
