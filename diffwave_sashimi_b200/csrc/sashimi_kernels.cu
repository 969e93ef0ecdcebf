// Per-step kernels of the SaShiMi backbone other than the FFT convolution, plus the pieces
// shared with WaveNet (t-embedding MLP, 1->C input conv, output head, DDPM update, weight fold).
//
// Data layout: every activation is (B, H, l) fp32, time contiguous, exactly as the reference's
// tensors.  Every kernel that produces a stream tensor also emits the TransposedLN statistics
// (mean, rstd over channels) of what it wrote, (B, l, 2), so the next block's FFT-conv prologue
// applies norm1 without another pass over the data.
//
// Reference: models/sashimi.py:11-20 (TransposedLN), :23-58 (pools), :60-75 (FF), :143-184
// (DiffWaveBlock), :277-313 (Sashimi.forward); models/s4.py:1435 (output_linear + GLU);
// models/utils.py:4-29 + sashimi.py:287-289 (t-embedding); generate.py:52-54 (update).
#include "common.cuh"
#include "kernels.h"
#include "tile_gemm.cuh"

namespace dwb {

// ---------------------------------------------------------------------------------------
// weight preparation (finalize)
// ---------------------------------------------------------------------------------------
// w[m, :] = g[m] * v[m, :] / ||v[m, :]||, written transposed: out[(k*taps_stride...)].
// v is (M, Kin, taps); out is [(tap*Kin + kin)][M]  (K-major rows for tile_gemm).
// g == nullptr: plain (not weight-normed) weight.
__global__ void fold_weight_kernel(const float *__restrict__ v, const float *__restrict__ g, int M, int Kin, int taps,
                                   float *__restrict__ out) {
    const int m = blockIdx.x;
    const int n = Kin * taps;
    __shared__ float red[32];
    float ss = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const float a = v[(size_t)m * n + i];
        ss = fmaf(a, a, ss);
    }
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int i = 0; i < (blockDim.x + 31) / 32; ++i) tot += red[i];
    const float scale = g ? g[m] / sqrtf(tot) : 1.0f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int kin = i / taps, tap = i - kin * taps;
        out[(size_t)(tap * Kin + kin) * M + m] = v[(size_t)m * n + i] * scale;
    }
}

int fold_weight(const float *v, const float *g, int M, int Kin, int taps, float *out, cudaStream_t st) {
    fold_weight_kernel<<<M, 128, 0, st>>>(v, g, M, Kin, taps, out);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// ---------------------------------------------------------------------------------------
// mel conditioning front end inside the model (t-independent, once per utterance):
// ConvTranspose2d(1, 1, (3, 2s), stride (1, s), padding (1, s/2)) + leaky-ReLU(0.4), twice, then a
// 1x1 mel_bands -> H conv on the first l samples           (models/sashimi.py:133-141,160-175,
// models/wavenet.py:62-70,98-111).  out[f][x] = b + sum_{kf<3} sum_{kx = (x + s/2) mod s (+ s)}
// in[f + 1 - kf][(x + s/2 - kx) / s] w[kf][kx]: six taps per output sample.
// ---------------------------------------------------------------------------------------
__global__ void mel_upsample_kernel(const float *__restrict__ in, int F, int Win, int s, const float *__restrict__ w,
                                    const float *__restrict__ bias, float *__restrict__ out, long long total) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int Wout = Win * s;
    const int x = (int)(idx % Wout);
    const long long r = idx / Wout;
    const int f = (int)(r % F);
    const long long b = r / F;
    const int k0 = (x + s / 2) % s;
    float acc = bias[0];
#pragma unroll
    for (int kf = 0; kf < 3; ++kf) {
        const int fi = f + 1 - kf;
        if (fi < 0 || fi >= F) continue;
        const float *row = in + ((size_t)b * F + fi) * Win;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int kx = k0 + q * s;
            const int xi = (x + s / 2 - kx) / s;          // exact: x + s/2 - kx is a multiple of s (may be negative)
            if (x + s / 2 - kx >= 0 && xi < Win) acc = fmaf(row[xi], w[kf * 2 * s + kx], acc);
        }
    }
    out[idx] = acc > 0.f ? acc : 0.4f * acc;
}

// out[b][m][t] = bias[m] + sum_k Wt[k][m] u[b][k][t],  t < l <= Wu
__global__ void mel_conv_kernel(const float *__restrict__ u, int K, int Wu, const float *__restrict__ Wt,
                                const float *__restrict__ bias, int Hc, int l, float *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int m = blockIdx.y, b = blockIdx.z;
    if (t >= l) return;
    const float *ub = u + (size_t)b * K * Wu + t;
    float acc = bias[m];
    for (int k = 0; k < K; ++k) acc = fmaf(Wt[(size_t)k * Hc + m], ub[(size_t)k * Wu], acc);
    out[((size_t)b * Hc + m) * l + t] = acc;
}

int mel_upsample_launch(const float *in, int rows, int F, int Win, int s, const float *w, const float *bias, float *out,
                        cudaStream_t st) {
    const long long total = (long long)rows * F * Win * s;
    mel_upsample_kernel<<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(in, F, Win, s, w, bias, out, total);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

int mel_conv_launch(const float *u, int rows, int K, int Wu, const float *Wt, const float *bias, int Hc, int l, float *out,
                    cudaStream_t st) {
    DWB_REQUIRE(Hc <= 65535 && rows <= 65535, DWB_ERR_UNSUPPORTED, "mel_conv: Hc=%d rows=%d exceed the grid", Hc, rows);
    mel_conv_kernel<<<dim3(ceil_div(l, 128), Hc, rows), 128, 0, st>>>(u, K, Wu, Wt, bias, Hc, l, out);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// ---------------------------------------------------------------------------------------
// diffusion-step embedding: t -> [sin(t f), cos(t f)] -> swish(fc1) -> swish(fc2) -> all per-layer fc_t
// ---------------------------------------------------------------------------------------
// one CTA per embedding row (batch element, or step index when building the per-step table)
__global__ void __launch_bounds__(256)
embed_mlp_kernel(const float *__restrict__ t, int E_in, int E_mid, int E_out, const float *__restrict__ W1,
                 const float *__restrict__ b1, const float *__restrict__ W2, const float *__restrict__ b2,
                 float *__restrict__ emb /* (rows, E_out) */) {
    extern __shared__ float sm[];
    float *e0 = sm, *e1 = sm + E_in;
    const int r = blockIdx.x;
    const float tv = t[r];
    const int half = E_in / 2;
    const float c = (float)(log(10000.0) / (double)(half - 1));   // python double rounded to fp32, like torch's scalar multiply
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const float f = expf((float)i * -c);
        const float a = tv * f;
        e0[i] = sinf(a);
        e0[half + i] = cosf(a);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int m = warp; m < E_mid; m += nw) {
        float s = 0.f;
        for (int k = lane; k < E_in; k += 32) s = fmaf(W1[(size_t)m * E_in + k], e0[k], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            s += b1[m];
            e1[m] = s / (1.0f + expf(-s));
        }
    }
    __syncthreads();
    for (int m = warp; m < E_out; m += nw) {
        float s = 0.f;
        for (int k = lane; k < E_mid; k += 32) s = fmaf(W2[(size_t)m * E_mid + k], e1[k], s);
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) {
            s += b2[m];
            emb[(size_t)r * E_out + m] = s / (1.0f + expf(-s));
        }
    }
}

// part[r, m] = Wt_all[m, :] . emb[r, :] + bt_all[m]; one warp per output, m over the stacked fc_t rows
__global__ void __launch_bounds__(256)
embed_fc_kernel(const float *__restrict__ emb, int E_out, const float *__restrict__ Wt, const float *__restrict__ bt,
                int Mtot, float *__restrict__ part /* (rows, Mtot) */) {
    const int r = blockIdx.y;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m = blockIdx.x * (blockDim.x >> 5) + warp;
    if (m >= Mtot) return;
    float s = 0.f;
    for (int k = lane; k < E_out; k += 32) s = fmaf(__ldg(Wt + (size_t)m * E_out + k), emb[(size_t)r * E_out + k], s);
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) part[(size_t)r * Mtot + m] = s + bt[m];
}

int embed_launch(const float *t, int rows, int E_in, int E_mid, int E_out, const float *W1, const float *b1,
                 const float *W2, const float *b2, const float *Wt, const float *bt, int Mtot, float *emb, float *part,
                 cudaStream_t st) {
    embed_mlp_kernel<<<rows, 256, (E_in + E_mid) * sizeof(float), st>>>(t, E_in, E_mid, E_out, W1, b1, W2, b2, emb);
    DWB_LAUNCH_CHECK();
    dim3 grid(ceil_div(Mtot, 8), rows);
    embed_fc_kernel<<<grid, 256, 0, st>>>(emb, E_out, Wt, bt, Mtot, part);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// ---------------------------------------------------------------------------------------
// input conv 1 -> C, ReLU (+ channel statistics)          wavenet.py:184,206 / sashimi.py:209,281
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
init_conv_kernel(const float *__restrict__ x, const float *__restrict__ w, const float *__restrict__ bias, int C, int l,
                 float *__restrict__ out, float *__restrict__ stats) {
    extern __shared__ float sw[];
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        sw[i] = w[i];
        sw[C + i] = bias[i];
    }
    __syncthreads();
    const int b = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= l) return;
    const float xv = x[(size_t)b * l + t];
    float *o = out + (size_t)b * C * l + t;
    float sum = 0.f;
    for (int c = 0; c < C; ++c) {
        const float v = fmaxf(fmaf(sw[c], xv, sw[C + c]), 0.f);
        o[(size_t)c * l] = v;
        sum += v;
    }
    if (stats) {
        const float mean = sum / (float)C;
        float var = 0.f;
        for (int c = 0; c < C; ++c) {
            const float d = fmaxf(fmaf(sw[c], xv, sw[C + c]), 0.f) - mean;
            var = fmaf(d, d, var);
        }
        var /= (float)C;
        stats[((size_t)b * l + t) * 2] = mean;
        stats[((size_t)b * l + t) * 2 + 1] = 1.0f / sqrtf(var);
    }
}

int init_conv_launch(const float *x, const float *w, const float *bias, int B, int C, int l, float *out, float *stats,
                     cudaStream_t st) {
    dim3 grid(ceil_div(l, 256), B);
    init_conv_kernel<<<grid, 256, 2 * C * sizeof(float), st>>>(x, w, bias, C, l, out, stats);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// ---------------------------------------------------------------------------------------
// DiffWaveBlock, second half: everything after the S4 activation
//   q = Wo g + bo ; y = q[:H] * sigmoid(q[H:]) (+ cond) ; x1 = x + y
//   x2 = x1 + W2 gelu(W1 LN2(x1) + b1) + b2 (+ skip) ; stats(x2)
// tile = all channels x TT time steps of one batch element
// ---------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(MIX_THREADS)
sashimi_mix_kernel(MixArgs a) {
    using G = TileGeom<TT>;
    extern __shared__ __align__(16) float smem[];
    const int H = a.H, F = a.F, l = a.l;
    float *Gs = smem;                               // [H][XS]   g, later LN2(x1), later x2
    float *Xs = Gs + (size_t)H * G::XS;             // [H][XS]   x, then x1
    float *Hs = Xs + (size_t)H * G::XS;             // [F][XS]   hidden
    float *scratch = Hs + (size_t)F * G::XS;        // 2*NP*TT
    float *stat_s = scratch + 2 * MIX_THREADS;      // 2*TT
    const int tid = threadIdx.x;
    const int cg = tid % G::NCG, rg = tid / G::NCG;
    const int b = blockIdx.y;
    const int t0 = blockIdx.x * TT;
    const size_t boff = (size_t)b * H * l;

    // ---- load g and x tiles (zero-filled past l)
    for (int i = tid; i < H * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        const bool ok = t0 + c < l;
        const size_t gi = boff + (size_t)r * l + t0 + c;
        Gs[(size_t)r * G::XS + c] = ok ? a.g[gi] : 0.f;
        Xs[(size_t)r * G::XS + c] = ok ? a.x[gi] : 0.f;
    }
    __syncthreads();

    // ---- output_linear (H -> 2H) + GLU + residual
    for (int m0 = 0; m0 < H; m0 += G::CHUNK) {
        float acc_a[4][4], acc_b[4][4];
        zero_acc(acc_a);
        zero_acc(acc_b);
        tile_gemm_chunk<TT>(a.Wo_t, 2 * H, H, H, m0, Gs, acc_a, rg, cg);
        tile_gemm_chunk<TT>(a.Wo_t + H, 2 * H, H, H, m0, Gs, acc_b, rg, cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rg + G::NRG * i;
            if (m >= H) continue;
            const float ba = a.bo[m], bb = a.bo[H + m];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * cg + j;
                float y = (acc_a[i][j] + ba) * sigmoidf_(acc_b[i][j] + bb);
                if (a.cond && t0 + c < l) y += a.cond[((size_t)(a.cond_stride_b ? b : 0) * H + m) * l + t0 + c];
                Xs[(size_t)m * G::XS + c] += y;
            }
        }
    }
    __syncthreads();

    // ---- LN2 over channels, written over the g tile
    tile_col_stats<TT>(Xs, H, scratch, stat_s, tid);
    for (int i = tid; i < H * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        Gs[(size_t)r * G::XS + c] = (a.ln2_s * stat_s[2 * c + 1]) * (Xs[(size_t)r * G::XS + c] - stat_s[2 * c] + a.ln2_m);
    }
    __syncthreads();

    // ---- FF: hidden = gelu(W1 z + b1)
    for (int m0 = 0; m0 < F; m0 += G::CHUNK) {
        float acc[4][4];
        zero_acc(acc);
        tile_gemm_chunk<TT>(a.W1_t, F, H, F, m0, Gs, acc, rg, cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rg + G::NRG * i;
            if (m >= F) continue;
            const float bv = a.b1[m];
#pragma unroll
            for (int j = 0; j < 4; ++j) Hs[(size_t)m * G::XS + 4 * cg + j] = gelu_erf(acc[i][j] + bv);
        }
    }
    __syncthreads();

    // ---- x2 = x1 + W2 hidden + b2 (+ skip) -> g tile (z is dead)
    for (int m0 = 0; m0 < H; m0 += G::CHUNK) {
        float acc[4][4];
        zero_acc(acc);
        tile_gemm_chunk<TT>(a.W2_t, H, F, H, m0, Hs, acc, rg, cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rg + G::NRG * i;
            if (m >= H) continue;
            const float bv = a.b2[m];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = 4 * cg + j;
                float v = Xs[(size_t)m * G::XS + c] + acc[i][j] + bv;
                if (a.skip && t0 + c < l) v += a.skip[boff + (size_t)m * l + t0 + c];
                Gs[(size_t)m * G::XS + c] = v;
            }
        }
    }
    __syncthreads();

    // ---- statistics of the block output for the next norm, then store
    tile_col_stats<TT>(Gs, H, scratch, stat_s, tid);
    for (int i = tid; i < H * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        if (t0 + c < l) a.out[boff + (size_t)r * l + t0 + c] = Gs[(size_t)r * G::XS + c];
    }
    if (tid < TT && t0 + tid < l) {
        a.stats_out[((size_t)b * l + t0 + tid) * 2] = stat_s[2 * tid];
        a.stats_out[((size_t)b * l + t0 + tid) * 2 + 1] = stat_s[2 * tid + 1];
    }
}

template <int TT>
static size_t mix_smem(int H, int F) {
    using G = TileGeom<TT>;
    return ((size_t)(2 * H + F) * G::XS + 2 * MIX_THREADS + 2 * TT) * sizeof(float);
}

template <typename KernelT>
static int set_smem(KernelT k, size_t bytes) {
    DWB_REQUIRE(bytes <= 227 * 1024, DWB_ERR_UNSUPPORTED, "tile needs %zu B of shared memory (> 227 KB)", bytes);
    if (bytes > 48 * 1024) DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return DWB_OK;
}

int mix_launch(const MixArgs &a, int B, cudaStream_t st) {
    int rc;
    if (mix_smem<32>(a.H, a.F) <= 200 * 1024) {
        const size_t sm = mix_smem<32>(a.H, a.F);
        if ((rc = set_smem(sashimi_mix_kernel<32>, sm)) != DWB_OK) return rc;
        dim3 grid(ceil_div(a.l, 32), B);
        sashimi_mix_kernel<32><<<grid, MIX_THREADS, sm, st>>>(a);
    } else {
        const size_t sm = mix_smem<16>(a.H, a.F);
        if ((rc = set_smem(sashimi_mix_kernel<16>, sm)) != DWB_OK) return rc;
        dim3 grid(ceil_div(a.l, 16), B);
        sashimi_mix_kernel<16><<<grid, MIX_THREADS, sm, st>>>(a);
    }
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// ---------------------------------------------------------------------------------------
// pools                                                         sashimi.py:23-58
//   down(s): x'[b, h*s+j, l'] = x[b, h, l'*s+j] ; out = W x' + bias          (H*s -> Ho)
//   up(s)  : y = W x + bias (Hi -> Ho*s) ; out[b, h, l*s+j] = y[b, h*s+j, l] (+ skip)
// ---------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(MIX_THREADS)
down_pool_kernel(PoolArgs a) {
    using G = TileGeom<TT>;
    extern __shared__ __align__(16) float smem[];
    const int Hi = a.Hi, Ho = a.Ho, s = a.s, li = a.li, lo = li / s, K = Hi * s;
    float *Xs = smem;                                // [K][XS]
    float *Os = Xs + (size_t)K * G::XS;              // [Ho][XS]
    float *scratch = Os + (size_t)Ho * G::XS;
    float *stat_s = scratch + 2 * MIX_THREADS;
    const int tid = threadIdx.x, cg = tid % G::NCG, rg = tid / G::NCG;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    // gather: consecutive threads read consecutive input samples of one channel
    for (int i = tid; i < Hi * TT * s; i += MIX_THREADS) {
        const int h = i / (TT * s), r = i - h * (TT * s);    // r = c*s + j
        const int c = r / s, j = r - c * s;
        const bool ok = t0 + c < lo;
        Xs[(size_t)(h * s + j) * G::XS + c] = ok ? a.x[((size_t)b * Hi + h) * li + (size_t)t0 * s + r] : 0.f;
    }
    __syncthreads();
    for (int m0 = 0; m0 < Ho; m0 += G::CHUNK) {
        float acc[4][4];
        zero_acc(acc);
        tile_gemm_chunk<TT>(a.W_t, Ho, K, Ho, m0, Xs, acc, rg, cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rg + G::NRG * i;
            if (m >= Ho) continue;
            const float bv = a.bias[m];
#pragma unroll
            for (int j = 0; j < 4; ++j) Os[(size_t)m * G::XS + 4 * cg + j] = acc[i][j] + bv;
        }
    }
    __syncthreads();
    tile_col_stats<TT>(Os, Ho, scratch, stat_s, tid);
    for (int i = tid; i < Ho * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        if (t0 + c < lo) a.out[((size_t)b * Ho + r) * lo + t0 + c] = Os[(size_t)r * G::XS + c];
    }
    if (tid < TT && t0 + tid < lo) {
        a.stats_out[((size_t)b * lo + t0 + tid) * 2] = stat_s[2 * tid];
        a.stats_out[((size_t)b * lo + t0 + tid) * 2 + 1] = stat_s[2 * tid + 1];
    }
}

// up pool: tile of TT input columns -> TT*s output columns; TTO = TT*s must be <= 64 here
template <int TT, int S>
__global__ void __launch_bounds__(MIX_THREADS)
up_pool_kernel(PoolArgs a) {
    using G = TileGeom<TT>;
    constexpr int TTO = TT * S;
    using GO = TileGeom<TTO>;
    extern __shared__ __align__(16) float smem[];
    const int Hi = a.Hi, Ho = a.Ho, li = a.li, lo = li * S;
    float *Xs = smem;                                 // [Hi][XS]
    float *Os = Xs + (size_t)Hi * G::XS;              // [Ho][GO::XS]
    float *scratch = Os + (size_t)Ho * GO::XS;
    float *stat_s = scratch + 2 * MIX_THREADS;
    const int tid = threadIdx.x, cg = tid % G::NCG, rg = tid / G::NCG;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    for (int i = tid; i < Hi * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        Xs[(size_t)r * G::XS + c] = (t0 + c < li) ? a.x[((size_t)b * Hi + r) * li + t0 + c] : 0.f;
    }
    __syncthreads();
    const int M = Ho * S;
    for (int m0 = 0; m0 < M; m0 += G::CHUNK) {
        float acc[4][4];
        zero_acc(acc);
        tile_gemm_chunk<TT>(a.W_t, M, Hi, M, m0, Xs, acc, rg, cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rg + G::NRG * i;
            if (m >= M) continue;
            const int h = m / S, j = m - h * S;
            const float bv = a.bias[m];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int c = 4 * cg + jj;
                Os[(size_t)h * GO::XS + c * S + j] = acc[i][jj] + bv;
            }
        }
    }
    __syncthreads();
    if (a.skip) {
        for (int i = tid; i < Ho * TTO; i += MIX_THREADS) {
            const int r = i / TTO, c = i - r * TTO;
            if (t0 * S + c < lo) Os[(size_t)r * GO::XS + c] += a.skip[((size_t)b * Ho + r) * lo + (size_t)t0 * S + c];
        }
        __syncthreads();
    }
    tile_col_stats<TTO>(Os, Ho, scratch, stat_s, tid);
    for (int i = tid; i < Ho * TTO; i += MIX_THREADS) {
        const int r = i / TTO, c = i - r * TTO;
        if (t0 * S + c < lo) a.out[((size_t)b * Ho + r) * lo + (size_t)t0 * S + c] = Os[(size_t)r * GO::XS + c];
    }
    if (tid < TTO && t0 * S + tid < lo) {
        a.stats_out[((size_t)b * lo + (size_t)t0 * S + tid) * 2] = stat_s[2 * tid];
        a.stats_out[((size_t)b * lo + (size_t)t0 * S + tid) * 2 + 1] = stat_s[2 * tid + 1];
    }
}

int down_pool_launch(const PoolArgs &a, int B, cudaStream_t st) {
    using G = TileGeom<16>;
    DWB_REQUIRE(a.li % a.s == 0, DWB_ERR_INVALID, "down pool: length %d not divisible by %d", a.li, a.s);
    const size_t sm = ((size_t)(a.Hi * a.s + a.Ho) * G::XS + 2 * MIX_THREADS + 2 * 16) * sizeof(float);
    int rc = set_smem(down_pool_kernel<16>, sm);
    if (rc != DWB_OK) return rc;
    dim3 grid(ceil_div(a.li / a.s, 16), B);
    down_pool_kernel<16><<<grid, MIX_THREADS, sm, st>>>(a);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

template <int S>
static int up_pool_launch_s(const PoolArgs &a, int B, cudaStream_t st) {
    constexpr int TT = 16;
    using G = TileGeom<TT>;
    using GO = TileGeom<TT * S>;
    const size_t sm = ((size_t)a.Hi * G::XS + (size_t)a.Ho * GO::XS + 2 * MIX_THREADS + 2 * TT * S) * sizeof(float);
    int rc = set_smem(up_pool_kernel<TT, S>, sm);
    if (rc != DWB_OK) return rc;
    dim3 grid(ceil_div(a.li, TT), B);
    up_pool_kernel<TT, S><<<grid, MIX_THREADS, sm, st>>>(a);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

int up_pool_launch(const PoolArgs &a, int B, cudaStream_t st) {
    switch (a.s) {
        case 2: return up_pool_launch_s<2>(a, B, st);
        case 4: return up_pool_launch_s<4>(a, B, st);
        default:
            set_error("up pool factor %d unsupported (2 or 4)", a.s);
            return DWB_ERR_UNSUPPORTED;
    }
}

// ---------------------------------------------------------------------------------------
// output head: [LN] -> (prescale) -> Wf (C->C) + ReLU -> wz (C->1)  [-> DDPM update]
//   sashimi.py:310-311 / wavenet.py:165,198-200,208 ; generate.py:52-54
// ---------------------------------------------------------------------------------------
template <int TT>
__global__ void __launch_bounds__(MIX_THREADS)
head_kernel(HeadArgs a) {
    using G = TileGeom<TT>;
    extern __shared__ __align__(16) float smem[];
    const int C = a.C, l = a.l;
    float *Xs = smem;                               // [C][XS]
    float *red = Xs + (size_t)C * G::XS;            // [NRG][TT]
    const int tid = threadIdx.x, cg = tid % G::NCG, rg = tid / G::NCG;
    const int b = blockIdx.y, t0 = blockIdx.x * TT;
    for (int i = tid; i < C * TT; i += MIX_THREADS) {
        const int r = i / TT, c = i - r * TT;
        float v = 0.f;
        if (t0 + c < l) {
            v = a.x[((size_t)b * C + r) * l + t0 + c];
            if (a.stats) {
                const float mean = a.stats[((size_t)b * l + t0 + c) * 2], rstd = a.stats[((size_t)b * l + t0 + c) * 2 + 1];
                v = (a.ln_s * rstd) * (v - mean + a.ln_m);
            }
            v *= a.prescale;
        }
        Xs[(size_t)r * G::XS + c] = v;
    }
    __syncthreads();
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    for (int m0 = 0; m0 < C; m0 += G::CHUNK) {
        float acc[4][4];
        zero_acc(acc);
        tile_gemm_chunk<TT>(a.Wf_t, C, C, C, m0, Xs, acc, rg, cg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int m = m0 + rg + G::NRG * i;
            if (m >= C) continue;
            const float bv = a.bf[m], wz = a.wz[m];
#pragma unroll
            for (int j = 0; j < 4; ++j) part[j] = fmaf(wz, fmaxf(acc[i][j] + bv, 0.f), part[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) red[rg * TT + 4 * cg + j] = part[j];
    __syncthreads();
    if (tid < TT && t0 + tid < l) {
        float e = a.bz;
        for (int r = 0; r < G::NRG; ++r) e += red[r * TT + tid];
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

int head_launch(const HeadArgs &a, int B, cudaStream_t st) {
    constexpr int TT = 32;
    using G = TileGeom<TT>;
    const size_t sm = ((size_t)a.C * G::XS + (size_t)G::NRG * TT) * sizeof(float);
    int rc = set_smem(head_kernel<TT>, sm);
    if (rc != DWB_OK) return rc;
    dim3 grid(ceil_div(a.l, TT), B);
    head_kernel<TT><<<grid, MIX_THREADS, sm, st>>>(a);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

}  // namespace dwb
