// DiffWaveBlock channel mixing on tensor cores (split-bf16, fp32 accumulate).
//
// Same computation, tiling and epilogues as sashimi_mix_kernel (sashimi_kernels.cu):
//   q = Wo g + bo ; y = q[:H] * sigmoid(q[H:]) (+cond) ; x1 = x + y
//   x2 = x1 + W2 gelu(W1 LN2(x1) + b1) + b2 (+skip) ; stats(x2)
// but the three channel contractions (12 H^2 flop per time step, ~27 GFLOP per clip-step at
// d_model 64: 10x more than an fp32 SIMT pipe can do inside the HBM time of the block) run on
// the tensor cores.  fp32 operands are split x = hi + lo into two bf16 halves and each product
// is evaluated as hi*hi + lo*hi + hi*lo with fp32 accumulation (relative error ~2^-17 per
// product, i.e. 1e-5 — the parity budget is 1e-3 and single-pass bf16/tf32 would consume all of
// it).
//
// Layout: weights are pre-split and stored in mma.m16n8k16 A-fragment order at finalize
// (one coalesced 16-byte load per lane per fragment, no shared-memory staging, no ldmatrix);
// activations sit in shared memory as [k][t] bf16 (t contiguous, exactly how they stream in from
// HBM) and reach the B fragments through ldmatrix.trans.
#include "common.cuh"
#include "kernels.h"
#include "mma_split.cuh"
#include "tile_gemm.cuh"

namespace dwb {

// ---------------------------------------------------------------------------------------
// finalize: folded fp32 weight Wt [K][M] (transposed) -> hi / lo bf16 A fragments
//   frag[((mt*KT + kt)*32 + lane)*4 + r]: r0={a0,a1} r1={a2,a3} r2={a4,a5} r3={a6,a7}
// ---------------------------------------------------------------------------------------
__global__ void frag_pack_kernel(const float *__restrict__ Wt, int M, int K, uint32_t *__restrict__ fhi,
                                 uint32_t *__restrict__ flo) {
    const int KT = K / 16;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;   // one (mt, kt, lane, r)
    const size_t total = (size_t)(M / 16) * KT * 32 * 4;
    if (idx >= total) return;
    const int r = idx & 3, lane = (idx >> 2) & 31;
    const size_t tile = idx >> 7;
    const int kt = tile % KT, mt = tile / KT;
    const int g = lane >> 2, t = lane & 3;
    const int m = mt * 16 + g + ((r & 1) ? 8 : 0);
    const int k = kt * 16 + 2 * t + ((r & 2) ? 8 : 0);
    const float w0 = Wt[(size_t)k * M + m], w1 = Wt[(size_t)(k + 1) * M + m];
    const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(w0 - __bfloat162float(h0));
    const __nv_bfloat16 l1 = __float2bfloat16_rn(w1 - __bfloat162float(h1));
    fhi[idx] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    flo[idx] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
}

int frag_pack(const float *Wt, int M, int K, uint32_t *fhi, uint32_t *flo, cudaStream_t st) {
    DWB_REQUIRE(M % 16 == 0 && K % 16 == 0, DWB_ERR_INVALID, "frag_pack: M=%d K=%d must be multiples of 16", M, K);
    const size_t total = (size_t)(M / 16) * (K / 16) * 128;
    frag_pack_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(Wt, M, K, fhi, flo);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// H: block width; FM: F / H; TT: time columns per CTA; WM x WN warps (WM*WN = 8)
template <int H, int FM, int TT, int WM, int WN>
struct MmaCfg {
    static constexpr int F = FM * H;
    static constexpr int TTP = TT + 8;                 // bf16 row pitch: odd multiple of 16 B
    static constexpr int XS = TT + 4;                  // fp32 row pitch
    static constexpr int NT = TT / 8 / WN;             // n-tiles per warp
    static constexpr int PAIRS = H / 16 / WM;          // GLU tile pairs per warp (G1)
    static constexpr int MT2 = F / 16 / WM;            // m-tiles per warp (G2)
    static constexpr int MT3 = H / 16 / WM;            // m-tiles per warp (G3)
    static constexpr size_t SMEM = (size_t)H * XS * 4            // X1 fp32
                                   + (size_t)2 * H * TTP * 2      // g / z split
                                   + (size_t)2 * F * TTP * 2      // hidden split
                                   + (2 * MIX_THREADS + 2 * TT) * 4;
    static_assert(WM * WN == 8 && PAIRS >= 1 && MT3 >= 1 && NT >= 2, "bad tiling");
};

template <int H, int FM, int TT, int WM, int WN>
__global__ void __launch_bounds__(MIX_THREADS)
sashimi_mix_mma_kernel(MixArgs a) {
    using C = MmaCfg<H, FM, TT, WM, WN>;
    constexpr int F = C::F, TTP = C::TTP, XS = C::XS, NT = C::NT;
    extern __shared__ __align__(16) unsigned char smraw[];
    float *X1 = reinterpret_cast<float *>(smraw);                                  // [H][XS]
    __nv_bfloat16 *Ghi = reinterpret_cast<__nv_bfloat16 *>(X1 + (size_t)H * XS);   // [H][TTP]
    __nv_bfloat16 *Glo = Ghi + (size_t)H * TTP;
    __nv_bfloat16 *Hhi = Glo + (size_t)H * TTP;                                    // [F][TTP]
    __nv_bfloat16 *Hlo = Hhi + (size_t)F * TTP;
    float *scratch = reinterpret_cast<float *>(Hlo + (size_t)F * TTP);
    float *stat_s = scratch + 2 * MIX_THREADS;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp % WM, wn = warp / WM;
    const int col0 = wn * (TT / WN);
    const int g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.y, t0 = blockIdx.x * TT, l = a.l;
    const size_t boff = (size_t)b * H * l;

    // ---- load g (split) and x (fp32); l and t0 are even so float2 accesses stay in range pairwise
    {
        constexpr int ITEMS = H * (TT / 2), U = 4;          // U independent (g, x) loads in flight per thread
        static_assert(ITEMS % (MIX_THREADS * U) == 0, "tile load tiling");
        for (int i0 = tid; i0 < ITEMS; i0 += MIX_THREADS * U) {
            float2 gv[U], xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * MIX_THREADS;
                const int r = i / (TT / 2), c = 2 * (i - r * (TT / 2));
                const size_t gi = boff + (size_t)r * l + t0 + c;
                gv[u] = xv[u] = make_float2(0.f, 0.f);
                if (t0 + c + 1 < l) {
                    gv[u] = *reinterpret_cast<const float2 *>(a.g + gi);
                    xv[u] = *reinterpret_cast<const float2 *>(a.x + gi);
                } else if (t0 + c < l) {
                    gv[u].x = a.g[gi];
                    xv[u].x = a.x[gi];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * MIX_THREADS;
                const int r = i / (TT / 2), c = 2 * (i - r * (TT / 2));
                split_store2(Ghi, Glo, (size_t)r * TTP + c, gv[u].x, gv[u].y);
                *reinterpret_cast<float2 *>(X1 + (size_t)r * XS + c) = xv[u];
            }
        }
    }
    __syncthreads();

    // ---- G1: output_linear + GLU + residual -> X1
    constexpr int PG = C::PAIRS > 2 ? 2 : C::PAIRS;   // at most 4 m-tiles of accumulators live at once
    for (int grp = 0; grp < C::PAIRS / PG; ++grp) {
        constexpr int MT = 2 * PG;
        int tiles[MT];
#pragma unroll
        for (int p = 0; p < PG; ++p) {
            tiles[2 * p] = wm * C::PAIRS + grp * PG + p;
            tiles[2 * p + 1] = tiles[2 * p] + H / 16;
        }
        float acc[MT][NT][4];
        zero3(acc);
        gemm_split_bf16<MT, NT, TTP>(a.Wo_fh, a.Wo_fl, H / 16, tiles, Ghi, Glo, col0, acc, lane);
#pragma unroll
        for (int p = 0; p < PG; ++p)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int h = tiles[2 * p] * 16 + g + half * 8;
                const float ba = a.bo[h], bb = a.bo[H + h];
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int c = col0 + n * 8 + 2 * tq + j;
                        float y = (acc[2 * p][n][half * 2 + j] + ba) * sigmoidf_(acc[2 * p + 1][n][half * 2 + j] + bb);
                        if (a.cond && t0 + c < l) y += a.cond[((size_t)(a.cond_stride_b ? b : 0) * H + h) * l + t0 + c];
                        X1[(size_t)h * XS + c] += y;
                    }
            }
    }
    __syncthreads();

    // ---- LN2 -> z (split) over the g tile
    tile_col_stats<TT>(X1, H, scratch, stat_s, tid);
    for (int i = tid; i < H * (TT / 2); i += MIX_THREADS) {
        const int r = i / (TT / 2), c = 2 * (i - r * (TT / 2));
        const float2 xv = *reinterpret_cast<const float2 *>(X1 + (size_t)r * XS + c);
        const float z0 = (a.ln2_s * stat_s[2 * c + 1]) * (xv.x - stat_s[2 * c] + a.ln2_m);
        const float z1 = (a.ln2_s * stat_s[2 * c + 3]) * (xv.y - stat_s[2 * c + 2] + a.ln2_m);
        split_store2(Ghi, Glo, (size_t)r * TTP + c, z0, z1);
    }
    __syncthreads();

    // ---- G2: hidden = gelu(W1 z + b1) (split)
    constexpr int G2 = C::MT2 > 4 ? 4 : C::MT2;
    for (int grp = 0; grp < C::MT2 / G2; ++grp) {
        constexpr int MT = G2;
        int tiles[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) tiles[i] = wm * C::MT2 + grp * G2 + i;
        float acc[MT][NT][4];
        zero3(acc);
        gemm_split_bf16<MT, NT, TTP>(a.W1_fh, a.W1_fl, H / 16, tiles, Ghi, Glo, col0, acc, lane);
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int m = tiles[i] * 16 + g + half * 8;
                const float bv = a.b1[m];
#pragma unroll
                for (int n = 0; n < NT; ++n) {
                    const int c = col0 + n * 8 + 2 * tq;
                    split_store2(Hhi, Hlo, (size_t)m * TTP + c, gelu_fast(acc[i][n][half * 2] + bv),
                                 gelu_fast(acc[i][n][half * 2 + 1] + bv));
                }
            }
    }
    __syncthreads();

    // ---- G3: x2 = x1 + W2 hidden + b2 (+skip) -> X1 in place
    constexpr int G3 = C::MT3 > 4 ? 4 : C::MT3;
    for (int grp = 0; grp < C::MT3 / G3; ++grp) {
        constexpr int MT = G3;
        int tiles[MT];
#pragma unroll
        for (int i = 0; i < MT; ++i) tiles[i] = wm * C::MT3 + grp * G3 + i;
        float acc[MT][NT][4];
        zero3(acc);
        gemm_split_bf16<MT, NT, TTP>(a.W2_fh, a.W2_fl, F / 16, tiles, Hhi, Hlo, col0, acc, lane);
#pragma unroll
        for (int i = 0; i < MT; ++i)
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                const int h = tiles[i] * 16 + g + half * 8;
                const float bv = a.b2[h];
#pragma unroll
                for (int n = 0; n < NT; ++n)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const int c = col0 + n * 8 + 2 * tq + j;
                        float v = X1[(size_t)h * XS + c] + acc[i][n][half * 2 + j] + bv;
                        if (a.skip && t0 + c < l) v += a.skip[boff + (size_t)h * l + t0 + c];
                        X1[(size_t)h * XS + c] = v;
                    }
            }
    }
    __syncthreads();

    // ---- statistics for the next norm, store
    tile_col_stats<TT>(X1, H, scratch, stat_s, tid);
    for (int i = tid; i < H * (TT / 2); i += MIX_THREADS) {
        const int r = i / (TT / 2), c = 2 * (i - r * (TT / 2));
        const float2 v = *reinterpret_cast<const float2 *>(X1 + (size_t)r * XS + c);
        const size_t gi = boff + (size_t)r * l + t0 + c;
        if (t0 + c + 1 < l) *reinterpret_cast<float2 *>(a.out + gi) = v;
        else if (t0 + c < l) a.out[gi] = v.x;
    }
    if (tid < TT && t0 + tid < l) {
        a.stats_out[((size_t)b * l + t0 + tid) * 2] = stat_s[2 * tid];
        a.stats_out[((size_t)b * l + t0 + tid) * 2 + 1] = stat_s[2 * tid + 1];
    }
}

template <int H, int FM, int TT, int WM, int WN>
static int launch_mma(const MixArgs &a, int B, cudaStream_t st) {
    using C = MmaCfg<H, FM, TT, WM, WN>;
    auto k = sashimi_mix_mma_kernel<H, FM, TT, WM, WN>;
    static_assert(C::SMEM <= 227 * 1024, "tile does not fit shared memory");
    if (C::SMEM > 48 * 1024) DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    dim3 grid(ceil_div(a.l, TT), B);
    k<<<grid, MIX_THREADS, C::SMEM, st>>>(a);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

bool mix_mma_supported(int H, int F, int l) {
    return F == 2 * H && (l % 2 == 0) && (H == 32 || H == 64 || H == 128 || H == 256 || H == 512);
}

int mix_mma_launch(const MixArgs &a, int B, cudaStream_t st) {
    switch (a.H) {
        case 32: return launch_mma<32, 2, 64, 2, 4>(a, B, st);
        case 64: return launch_mma<64, 2, 64, 4, 2>(a, B, st);
        case 128: return launch_mma<128, 2, 64, 8, 1>(a, B, st);
        case 256: return launch_mma<256, 2, 32, 8, 1>(a, B, st);
        case 512: return launch_mma<512, 2, 16, 8, 1>(a, B, st);
    }
    set_error("mix_mma: H=%d unsupported", a.H);
    return DWB_ERR_UNSUPPORTED;
}

// ---------------------------------------------------------------------------------------
// pools on the same split-bf16 path                               (models/sashimi.py:23-58)
// ---------------------------------------------------------------------------------------
// Both kernels: the fp32 output staging tile aliases the split-bf16 operand tile (a barrier separates the
// last MMA from the first staging store), so the tile can be twice as wide in the same shared memory and
// the weight fragments (read from L2 by every CTA) are fetched half as often; a warp owns MTW m-tiles
// that share every B fragment it loads.
//
// down(s): x'[h*s+j][c] = x[b, h, (t0+c)*s + j];  out = W x' + bias  (K = Hi*s -> Ho), + stats
template <int TT, int MTW>
__global__ void __launch_bounds__(MIX_THREADS)
down_pool_mma_kernel(PoolArgs a) {
    constexpr int TTP = TT + 8, XS = TT + 4, NT = TT / 8;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int Hi = a.Hi, Ho = a.Ho, s = a.s, li = a.li, lo = li / s, K = Hi * s;
    const size_t ubytes = max((size_t)Ho * XS * 4, (size_t)2 * K * TTP * 2);
    float *Os = reinterpret_cast<float *>(smraw);                                  // [Ho][XS]   (after the GEMM)
    __nv_bfloat16 *Bhi = reinterpret_cast<__nv_bfloat16 *>(smraw);                 // [K][TTP]   (before)
    __nv_bfloat16 *Blo = Bhi + (size_t)K * TTP;
    float *scratch = reinterpret_cast<float *>(smraw + ((ubytes + 15) & ~(size_t)15));
    float *stat_s = scratch + 2 * MIX_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    for (int i = tid; i < Hi * TT * s; i += MIX_THREADS) {
        const int h = i / (TT * s), r = i - h * (TT * s);    // r = c*s + j: consecutive input samples
        const int c = r / s, j = r - c * s;
        const float v = (t0 + c < lo) ? a.x[((size_t)b * Hi + h) * li + (size_t)t0 * s + r] : 0.f;
        split_store(Bhi, Blo, (size_t)(h * s + j) * TTP + c, v);
    }
    __syncthreads();
    int tiles[MTW];
#pragma unroll
    for (int i = 0; i < MTW; ++i) tiles[i] = min(warp + 8 * i, Ho / 16 - 1);     // clamped duplicates are not stored
    float acc[MTW][NT][4];
    zero3(acc);
    gemm_split_bf16<MTW, NT, TTP>(a.W_fh, a.W_fl, K / 16, tiles, Bhi, Blo, 0, acc, lane);
    __syncthreads();                                                               // operands dead: Os may overwrite them
#pragma unroll
    for (int i = 0; i < MTW; ++i) {
        if (warp + 8 * i >= Ho / 16) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int m = tiles[i] * 16 + g + half * 8;
            const float bv = a.bias[m];
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int j = 0; j < 2; ++j) Os[(size_t)m * XS + n * 8 + 2 * tq + j] = acc[i][n][half * 2 + j] + bv;
        }
    }
    __syncthreads();
    tile_col_stats<TT>(Os, Ho, scratch, stat_s, tid);
    for (int i = tid; i < Ho * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        if (t0 + c < lo) a.out[((size_t)b * Ho + r) * lo + t0 + c] = Os[(size_t)r * XS + c];
    }
    if (tid < TT && t0 + tid < lo) {
        a.stats_out[((size_t)b * lo + t0 + tid) * 2] = stat_s[2 * tid];
        a.stats_out[((size_t)b * lo + t0 + tid) * 2 + 1] = stat_s[2 * tid + 1];
    }
}

// up(s): y = W x + bias (Hi -> Ho*s);  out[b, h, (t0+c)*s + j] = y[h*s+j][c] (+ skip), + stats
template <int S, int TT, int MTW>
__global__ void __launch_bounds__(MIX_THREADS)
up_pool_mma_kernel(PoolArgs a) {
    constexpr int TTP = TT + 8, NT = TT / 8, TTO = TT * S, XSO = TTO + 4;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int Hi = a.Hi, Ho = a.Ho, li = a.li, lo = li * S, M = Ho * S;
    const size_t ubytes = max((size_t)Ho * XSO * 4, (size_t)2 * Hi * TTP * 2);
    float *Os = reinterpret_cast<float *>(smraw);                                   // [Ho][XSO]  (after the GEMM)
    __nv_bfloat16 *Bhi = reinterpret_cast<__nv_bfloat16 *>(smraw);                  // [Hi][TTP]  (before)
    __nv_bfloat16 *Blo = Bhi + (size_t)Hi * TTP;
    float *scratch = reinterpret_cast<float *>(smraw + ((ubytes + 15) & ~(size_t)15));
    float *stat_s = scratch + 2 * MIX_THREADS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    for (int i = tid; i < Hi * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        split_store(Bhi, Blo, (size_t)r * TTP + c, (t0 + c < li) ? a.x[((size_t)b * Hi + r) * li + t0 + c] : 0.f);
    }
    __syncthreads();
    int tiles[MTW];
#pragma unroll
    for (int i = 0; i < MTW; ++i) tiles[i] = min(warp + 8 * i, M / 16 - 1);
    float acc[MTW][NT][4];
    zero3(acc);
    gemm_split_bf16<MTW, NT, TTP>(a.W_fh, a.W_fl, Hi / 16, tiles, Bhi, Blo, 0, acc, lane);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < MTW; ++i) {
        if (warp + 8 * i >= M / 16) continue;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int m = tiles[i] * 16 + g + half * 8;
            const int h = m / S, j = m - h * S;
            const float bv = a.bias[m];
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int jj = 0; jj < 2; ++jj) Os[(size_t)h * XSO + (n * 8 + 2 * tq + jj) * S + j] = acc[i][n][half * 2 + jj] + bv;
        }
    }
    __syncthreads();
    if (a.skip) {
        for (int i = tid; i < Ho * TTO; i += MIX_THREADS) {
            const int r = i / TTO, c = i - r * TTO;
            if (t0 * S + c < lo) Os[(size_t)r * XSO + c] += a.skip[((size_t)b * Ho + r) * lo + (size_t)t0 * S + c];
        }
        __syncthreads();
    }
    tile_col_stats<TTO>(Os, Ho, scratch, stat_s, tid);
    for (int i = tid; i < Ho * TTO; i += MIX_THREADS) {
        const int r = i / TTO, c = i - r * TTO;
        if (t0 * S + c < lo) a.out[((size_t)b * Ho + r) * lo + (size_t)t0 * S + c] = Os[(size_t)r * XSO + c];
    }
    if (tid < TTO && t0 * S + tid < lo) {
        a.stats_out[((size_t)b * lo + (size_t)t0 * S + tid) * 2] = stat_s[2 * tid];
        a.stats_out[((size_t)b * lo + (size_t)t0 * S + tid) * 2 + 1] = stat_s[2 * tid + 1];
    }
}

bool pool_mma_supported(int Hi, int Ho, int s, bool up) {
    if (up) return (s == 2 || s == 4) && Hi % 16 == 0 && (Ho * s) % 16 == 0;
    return (Hi * s) % 16 == 0 && Ho % 16 == 0;
}

template <typename KernelT, typename ArgsT>
static int launch_pool(KernelT k, const ArgsT &a, dim3 grid, size_t sm, cudaStream_t st) {
    DWB_REQUIRE(sm <= 227 * 1024, DWB_ERR_UNSUPPORTED, "pool tile needs %zu B of shared memory", sm);
    if (sm > 48 * 1024) DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k<<<grid, MIX_THREADS, sm, st>>>(a);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// ---------------------------------------------------------------------------------------
// output head on the same path: final LN, C -> C + ReLU, C -> 1, optional DDPM update
// (models/sashimi.py:310-311, models/wavenet.py:198-208, generate.py:52-54)
// 8 warps = 4 m-tile groups x 2 column halves; eps[t] = bz + sum_m wz[m] relu((Wf y)[m][t] + bf[m])
// ---------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(MIX_THREADS)
head_mma_kernel(HeadArgs a) {
    constexpr int TTP = TT + 8, NT = TT / 16;
    extern __shared__ __align__(16) unsigned char smraw[];
    const int C = a.C, l = a.l;
    __nv_bfloat16 *Bhi = reinterpret_cast<__nv_bfloat16 *>(smraw);                 // [C][TTP]
    __nv_bfloat16 *Blo = Bhi + (size_t)C * TTP;
    float *red = reinterpret_cast<float *>(Blo + (size_t)C * TTP);                  // [4][TT]
    float *sc = red + 4 * TT, *sh = sc + TT;                                        // LN scale / shift per column
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tq = lane & 3;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    if (tid < TT) {
        float s = a.prescale, h = 0.f;
        if (a.stats && t0 + tid < l) {
            const float2 ms = *reinterpret_cast<const float2 *>(a.stats + ((size_t)b * l + t0 + tid) * 2);
            s = a.ln_s * ms.y * a.prescale;          // v = (ln_s rstd)(x - mean + ln_m) prescale
            h = (a.ln_m - ms.x) * s;
        }
        sc[tid] = s;
        sh[tid] = h;
    }
    __syncthreads();
    for (int i = tid; i < C * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        float v = 0.f;
        if (t0 + c < l) v = fmaf(a.x[((size_t)b * C + r) * l + t0 + c], sc[c], sh[c]);
        split_store(Bhi, Blo, (size_t)r * TTP + c, v);
    }
    __syncthreads();
    const int wm = warp >> 1, col0 = (warp & 1) * (TT / 2);
    float part[NT][2];
#pragma unroll
    for (int n = 0; n < NT; ++n) part[n][0] = part[n][1] = 0.f;
    for (int mt = wm; mt < C / 16; mt += 4) {
        int tiles[1] = {mt};
        float acc[1][NT][4];
        zero3(acc);
        gemm_split_bf16<1, NT, TTP>(a.Wf_fh, a.Wf_fl, C / 16, tiles, Bhi, Blo, col0, acc, lane);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int m = mt * 16 + g + half * 8;
            const float bv = a.bf[m], wz = a.wz[m];
#pragma unroll
            for (int n = 0; n < NT; ++n)
#pragma unroll
                for (int j = 0; j < 2; ++j) part[n][j] = fmaf(wz, fmaxf(acc[0][n][half * 2 + j] + bv, 0.f), part[n][j]);
        }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float v = part[n][j];
            v += __shfl_xor_sync(0xffffffffu, v, 4);
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            if (g == 0) red[wm * TT + col0 + n * 8 + 2 * tq + j] = v;
        }
    __syncthreads();
    if (tid < TT && t0 + tid < l) {
        const float e = a.bz + ((red[tid] + red[TT + tid]) + (red[2 * TT + tid] + red[3 * TT + tid]));
        const size_t o = (size_t)b * l + t0 + tid;
        if (a.upd_x) {
            // x <- (x - c1 eps) / sqrt(alpha) (+ sigma z)           generate.py:52-54
            const float c1 = __ldg(a.ctl), sqrt_alpha = __ldg(a.ctl + 1), sigma = __ldg(a.ctl + 2);
            const int slot = __float_as_int(__ldg(a.ctl + 3));
            float xn = (a.upd_x[o] - c1 * e) / sqrt_alpha;
            if (slot >= 0) xn += sigma * (*a.noise_base)[(size_t)slot * gridDim.y * l + o];
            a.out[o] = xn;
        } else {
            a.out[o] = e;
        }
    }
}

int head_mma_launch(const HeadArgs &a, int B, cudaStream_t st) {
    constexpr int TT = 64;
    DWB_REQUIRE(a.Wf_fh && a.Wf_fl && a.C % 16 == 0, DWB_ERR_INVALID, "head_mma: C=%d needs packed weights", a.C);
    const size_t sm = (size_t)2 * a.C * (TT + 8) * 2 + (size_t)6 * TT * 4;
    return launch_pool(head_mma_kernel<TT>, a, dim3(ceil_div(a.l, TT), B), sm, st);
}

static size_t down_pool_smem(const PoolArgs &a, int TT) {
    const size_t u = std::max((size_t)a.Ho * (TT + 4) * 4, (size_t)2 * a.Hi * a.s * (TT + 8) * 2);
    return ((u + 15) & ~(size_t)15) + (size_t)(2 * MIX_THREADS + 2 * TT) * 4;
}
static size_t up_pool_smem(const PoolArgs &a, int TT) {
    const size_t u = std::max((size_t)a.Ho * (TT * a.s + 4) * 4, (size_t)2 * a.Hi * (TT + 8) * 2);
    return ((u + 15) & ~(size_t)15) + (size_t)(2 * MIX_THREADS + 2 * TT * a.s) * 4;
}

// widest tile with two CTAs per SM (<= 113 KB each) and <= 16 accumulator tiles per warp (MTW * TT/8)
template <int TT>
static int launch_down(const PoolArgs &a, int B, int mtw, cudaStream_t st) {
    const dim3 grid(ceil_div(a.li / a.s, TT), B);
    const size_t sm = down_pool_smem(a, TT);
    if constexpr (TT <= 64) if (mtw == 1) return launch_pool(down_pool_mma_kernel<TT, 1>, a, grid, sm, st);
    if constexpr (TT <= 64) if (mtw == 2) return launch_pool(down_pool_mma_kernel<TT, 2>, a, grid, sm, st);
    if constexpr (TT <= 32) if (mtw <= 4) return launch_pool(down_pool_mma_kernel<TT, 4>, a, grid, sm, st);
    if constexpr (TT <= 16) if (mtw <= 8) return launch_pool(down_pool_mma_kernel<TT, 8>, a, grid, sm, st);
    set_error("down_pool_mma: Ho=%d needs %d m-tiles per warp at TT=%d", a.Ho, mtw, TT);
    return DWB_ERR_UNSUPPORTED;
}

int down_pool_mma_launch(const PoolArgs &a, int B, cudaStream_t st) {
    const int mtw = ceil_div(a.Ho / 16, 8);
    const size_t cap = 113 * 1024;
    if (mtw <= 2 && down_pool_smem(a, 64) <= cap) return launch_down<64>(a, B, mtw, st);
    if (mtw <= 4 && down_pool_smem(a, 32) <= cap) return launch_down<32>(a, B, mtw, st);
    return launch_down<16>(a, B, mtw, st);
}

template <int S, int TT>
static int launch_up(const PoolArgs &a, int B, int mtw, cudaStream_t st) {
    const dim3 grid(ceil_div(a.li, TT), B);
    const size_t sm = up_pool_smem(a, TT);
    if constexpr (TT <= 64 && TT * S <= MIX_THREADS) if (mtw <= 2) return launch_pool(up_pool_mma_kernel<S, TT, 2>, a, grid, sm, st);
    if constexpr (TT <= 32) if (mtw <= 4) return launch_pool(up_pool_mma_kernel<S, TT, 4>, a, grid, sm, st);
    if constexpr (TT <= 16) if (mtw <= 8) return launch_pool(up_pool_mma_kernel<S, TT, 8>, a, grid, sm, st);
    set_error("up_pool_mma: Ho*s=%d needs %d m-tiles per warp at TT=%d", a.Ho * a.s, mtw, TT);
    return DWB_ERR_UNSUPPORTED;
}

template <int S>
static int up_pool_pick(const PoolArgs &a, int B, cudaStream_t st) {
    const int mtw = ceil_div(a.Ho * S / 16, 8);
    // measured (B200, unet d64, B = 32): wider tiles LOSE here (174 us at TT = 16 vs 212 / 341 us at 32 / 64): the kernel is
    // bound by its staging / statistics / scatter phases, not by the weight fragments, and wants many small CTAs
    return launch_up<S, 16>(a, B, mtw, st);
}

int up_pool_mma_launch(const PoolArgs &a, int B, cudaStream_t st) {
    return a.s == 2 ? up_pool_pick<2>(a, B, st) : up_pool_pick<4>(a, B, st);
}

}  // namespace dwb
