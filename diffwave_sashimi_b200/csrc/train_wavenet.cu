// Training step of the WaveNet backbone: loss + every parameter gradient + fused Adam,
// as hand-written fp32 kernels behind the C ABI.  First correct path: exact-fp32 SIMT GEMM tiles (the same
// arithmetic class as the reference's autograd), no tensor cores yet.
//
// Reference call sites this file stands in for:
//   train.py:198-222   training_loss(): x_t = sqrt(abar_t) x + sqrt(1 - abar_t) z;  MSE(net((x_t, t)), z)
//   train.py:137-143   optimizer.zero_grad(); loss.backward(); optimizer.step()   (torch.optim.Adam, train.py:92)
//   models/wavenet.py:82-121,149-165,202-210   the network whose backward is written out below
//
// Memory model: parameters, gradients and both Adam moments are FOUR FLAT fp32 buffers owned by the caller, laid
// out in net.parameters() order (dwb_trainer_layout).  The host mirror makes every nn.Parameter a view of the flat
// parameter buffer and every .grad a view of the flat gradient buffer, so state_dict()/checkpoints keep the
// reference's keys, the data-parallel gradient exchange is ONE all-reduce of one contiguous buffer (the reference
// flattens and unflattens per step, distributed_util.py:119-138) and Adam is one launch over the whole model.
//
// Backward, per residual layer n (d = 2^(n mod cycle), u = h_n + p_n inside [0,L), 0 outside - bias before padding):
//   forward   G = sum_k W_k u[. + (k-1)d] + b;  o = tanh(G_a) * sigmoid(G_b);  h_{n+1} = (h_n + W_r o + b_r) sqrt(1/2);
//             s += W_s o + b_s
//   given dh = dL/dh_{n+1} and ds = dL/ds (the same tensor for every layer):
//   do = sqrt(1/2) W_r^T dh + W_s^T ds;  dG_a = do * sg * (1 - th^2);  dG_b = do * th * sg * (1 - sg)
//   du[l] = sum_k W_k^T dG[l - (k-1)d];  dp_n = sum_l du;  dL/dh_n = du + sqrt(1/2) dh
//   dW_k = sum_{b,l} dG[l] u[l + (k-1)d]^T;  dW_r = sqrt(1/2) sum dh o^T;  dW_s = sum ds o^T;  biases = row sums
// Weight norm W = g v / |v| (per output row) is folded before the forward and un-folded after the backward:
//   dg = <dW, v> / |v|;  dv = (g / |v|) dW - (g <dW, v> / |v|^3) v.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include <cuda_bf16.h>

#include "common.cuh"

namespace dwb {
namespace train {

constexpr int TK = 16, NT = 256;
// Tiles are (64 W) x (64 W) outputs per 256-thread CTA, (4 W) x (4 W) per thread as W x W blocks of 4 x 4 spaced 64 apart
// (conflict-free float4 shared-memory reads).  W = 2 is the default (16 FMAs per shared-memory load instead of 8);
// DWB_TRAIN_TILE=64 selects W = 1.
// DWB_TRAIN_GEMM=mma|simt selects the GEMM family of the training step (tensor-core split-bf16 / exact fp32)
constexpr bool kTrainMmaDefault = false;
static bool train_mma() {
    static const bool on = [] {
        const char *e = getenv("DWB_TRAIN_GEMM");
        return e ? strcmp(e, "mma") == 0 : kTrainMmaDefault;
    }();
    return on;
}
static int tile_w() {
    static const int w = [] { const char *e = getenv("DWB_TRAIN_TILE"); return (e && atoi(e) == 64) ? 1 : 2; }();
    return w;
}

// Y[b,m,l] = alpha * sum_tap sum_k A(tap,m,k) X'[b,k,l+shift_tap] + bias_scale * bias[m] + beta * R[b,m,l]   (then ReLU)
// X'[b,k,l] = X[b,k,l] + rowadd[b,k] for 0 <= l < L, 0 outside.  A(tap,m,k) = A[tap*a_tap + m*a_m + k*a_k].
struct GemmArgs {
    const float *A;
    long long a_tap, a_m, a_k;
    const float *X, *rowadd, *bias, *R;
    float *Y;
    int M, K, L, ntap, shift[3];
    float alpha, bias_scale, beta;
    int relu;
};

template <int W>
__device__ __forceinline__ void tile_fma(const float (*As)[64 * W + 4], const float (*Bs)[64 * W + 4], int ty, int tx,
                                         float (&acc)[4 * W][4 * W]) {
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
        float av[4 * W], bv[4 * W];
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const float4 a = *reinterpret_cast<const float4 *>(&As[kk][w * 64 + ty * 4]);
            const float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][w * 64 + tx * 4]);
            av[4 * w] = a.x; av[4 * w + 1] = a.y; av[4 * w + 2] = a.z; av[4 * w + 3] = a.w;
            bv[4 * w] = b.x; bv[4 * w + 1] = b.y; bv[4 * w + 2] = b.z; bv[4 * w + 3] = b.w;
        }
#pragma unroll
        for (int i = 0; i < 4 * W; ++i)
#pragma unroll
            for (int j = 0; j < 4 * W; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}
// row / column of accumulator element i of thread coordinate t inside the tile
__device__ __forceinline__ int tile_pos(int t, int i) { return (i >> 2) * 64 + t * 4 + (i & 3); }

template <int W>
__global__ void __launch_bounds__(NT) cgemm_kernel(GemmArgs p) {
    constexpr int T = 64 * W;
    __shared__ __align__(16) float As[TK][T + 4];
    __shared__ __align__(16) float Bs[TK][T + 4];
    const int b = blockIdx.z, m0 = blockIdx.y * T, l0 = blockIdx.x * T;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4 * W][4 * W] = {};
    const float *Xb = p.X + (size_t)b * p.K * p.L;
    const float *ra = p.rowadd ? p.rowadd + (size_t)b * p.K : nullptr;
    for (int tap = 0; tap < p.ntap; ++tap) {
        const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
        for (int k0 = 0; k0 < p.K; k0 += TK) {
            for (int e = tid; e < TK * T; e += NT) {
                const int kk = e & (TK - 1), mm = e / TK, m = m0 + mm, k = k0 + kk;
                As[kk][mm] = (m < p.M && k < p.K) ? p.A[tap * p.a_tap + m * p.a_m + k * p.a_k] : 0.f;
            }
            for (int e = tid; e < TK * T; e += NT) {
                const int ll = e & (T - 1), kk = e / T, k = k0 + kk, l = l0 + ll + sh;
                float v = 0.f;
                if (k < p.K && l >= 0 && l < p.L) {
                    v = Xb[(size_t)k * p.L + l];
                    if (ra) v += ra[k];
                }
                Bs[kk][ll] = v;
            }
            __syncthreads();
            tile_fma<W>(As, Bs, ty, tx, acc);
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4 * W; ++i) {
        const int m = m0 + tile_pos(ty, i);
        if (m >= p.M) continue;
        const float bm = p.bias ? p.bias_scale * p.bias[m] : 0.f;
#pragma unroll
        for (int j = 0; j < 4 * W; ++j) {
            const int l = l0 + tile_pos(tx, j);
            if (l >= p.L) continue;
            const size_t idx = ((size_t)b * p.M + m) * p.L + l;
            float v = fmaf(p.alpha, acc[i][j], bm);
            if (p.R) v = fmaf(p.beta, p.R[idx], v);
            if (p.relu) v = fmaxf(v, 0.f);
            p.Y[idx] = v;
        }
    }
}

// dW(tap,m,k) += alpha * sum_{l in chunk} dY[b,m,l] X'[b,k,l+shift_tap]      (split over batch and time, fp32 atomics)
struct WgradArgs {
    const float *dY, *X, *rowadd;
    float *dW;
    long long o_tap, o_m, o_k;
    int M, K, L, ntap, shift[3], lchunk;
    float alpha;
};

template <int W>
__global__ void __launch_bounds__(NT) wgrad_kernel(WgradArgs p) {
    constexpr int T = 64 * W;
    __shared__ __align__(16) float As[TK][T + 4];
    __shared__ __align__(16) float Bs[TK][T + 4];
    const int nchunk = (p.L + p.lchunk - 1) / p.lchunk;
    const int b = blockIdx.x / nchunk, ch = blockIdx.x % nchunk;
    const int tiles_k = (p.K + T - 1) / T;
    const int m0 = (blockIdx.y / tiles_k) * T, k0 = (blockIdx.y % tiles_k) * T;
    const int tap = blockIdx.z, sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lbeg = ch * p.lchunk, lend = min(p.L, lbeg + p.lchunk);
    const float *dYb = p.dY + (size_t)b * p.M * p.L;
    const float *Xb = p.X + (size_t)b * p.K * p.L;
    const float *ra = p.rowadd ? p.rowadd + (size_t)b * p.K : nullptr;
    float acc[4 * W][4 * W] = {};
    for (int l0 = lbeg; l0 < lend; l0 += TK) {
        for (int e = tid; e < TK * T; e += NT) {
            const int ll = e & (TK - 1), mm = e / TK, m = m0 + mm, l = l0 + ll;
            As[ll][mm] = (m < p.M && l < lend) ? dYb[(size_t)m * p.L + l] : 0.f;
        }
        for (int e = tid; e < TK * T; e += NT) {
            const int ll = e & (TK - 1), kk = e / TK, k = k0 + kk, l = l0 + ll, ls = l + sh;
            float v = 0.f;
            if (k < p.K && l < lend && ls >= 0 && ls < p.L) {
                v = Xb[(size_t)k * p.L + ls];
                if (ra) v += ra[k];
            }
            Bs[ll][kk] = v;
        }
        __syncthreads();
        tile_fma<W>(As, Bs, ty, tx, acc);
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4 * W; ++i) {
        const int m = m0 + tile_pos(ty, i);
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4 * W; ++j) {
            const int k = k0 + tile_pos(tx, j);
            if (k >= p.K) continue;
            atomicAdd(&p.dW[tap * p.o_tap + m * p.o_m + k * p.o_k], p.alpha * acc[i][j]);
        }
    }
}

// ---- software-pipelined forms (default): the global loads of reduction step i+1 are issued into registers before the FMAs of
// step i run from shared memory.  profiles/ncu_r2_train.md: without this the 1x1 convolutions spend 3 issue slots waiting on
// `long_scoreboard` per instruction issued (FMA pipe 33-35 % active) - load, barrier, compute, barrier with nothing in flight.
// DWB_TRAIN_PREFETCH=0 keeps the plain forms above.
static bool train_prefetch() {
    static const bool on = [] { const char *e = getenv("DWB_TRAIN_PREFETCH"); return e ? atoi(e) != 0 : true; }();
    return on;
}

template <int W>
__global__ void __launch_bounds__(NT, 2) cgemm_pf_kernel(GemmArgs p) {
    constexpr int T = 64 * W, NV = TK * T / NT;
    __shared__ __align__(16) float As[TK][T + 4];
    __shared__ __align__(16) float Bs[TK][T + 4];
    const int b = blockIdx.z, m0 = blockIdx.y * T, l0 = blockIdx.x * T;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4 * W][4 * W] = {};
    const float *Xb = p.X + (size_t)b * p.K * p.L;
    const float *ra = p.rowadd ? p.rowadd + (size_t)b * p.K : nullptr;
    const int ksteps = (p.K + TK - 1) / TK, nsteps = p.ntap * ksteps;
    float va[NV], vb[NV];
    auto fetch = [&](int step) {
        const int tap = step / ksteps, k0 = (step - tap * ksteps) * TK;
        const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = tid + i * NT;
            const int kk = e & (TK - 1), mm = e / TK, m = m0 + mm, k = k0 + kk;
            va[i] = (m < p.M && k < p.K) ? p.A[tap * p.a_tap + m * p.a_m + k * p.a_k] : 0.f;
        }
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = tid + i * NT;
            const int ll = e & (T - 1), kk = e / T, k = k0 + kk, l = l0 + ll + sh;
            float v = 0.f;
            if (k < p.K && l >= 0 && l < p.L) {
                v = Xb[(size_t)k * p.L + l];
                if (ra) v += ra[k];
            }
            vb[i] = v;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = tid + i * NT;
            As[e & (TK - 1)][e / TK] = va[i];
            Bs[e / T][e & (T - 1)] = vb[i];
        }
    };
    fetch(0);
    stage();
    __syncthreads();
    for (int step = 0; step < nsteps; ++step) {
        const bool more = step + 1 < nsteps;
        if (more) fetch(step + 1);
        tile_fma<W>(As, Bs, ty, tx, acc);
        __syncthreads();
        if (more) {
            stage();
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4 * W; ++i) {
        const int m = m0 + tile_pos(ty, i);
        if (m >= p.M) continue;
        const float bm = p.bias ? p.bias_scale * p.bias[m] : 0.f;
#pragma unroll
        for (int j = 0; j < 4 * W; ++j) {
            const int l = l0 + tile_pos(tx, j);
            if (l >= p.L) continue;
            const size_t idx = ((size_t)b * p.M + m) * p.L + l;
            float v = fmaf(p.alpha, acc[i][j], bm);
            if (p.R) v = fmaf(p.beta, p.R[idx], v);
            if (p.relu) v = fmaxf(v, 0.f);
            p.Y[idx] = v;
        }
    }
}

template <int W>
__global__ void __launch_bounds__(NT, 2) wgrad_pf_kernel(WgradArgs p) {
    constexpr int T = 64 * W, NV = TK * T / NT;
    __shared__ __align__(16) float As[TK][T + 4];
    __shared__ __align__(16) float Bs[TK][T + 4];
    const int nchunk = (p.L + p.lchunk - 1) / p.lchunk;
    const int b = blockIdx.x / nchunk, ch = blockIdx.x % nchunk;
    const int tiles_k = (p.K + T - 1) / T;
    const int m0 = (blockIdx.y / tiles_k) * T, k0 = (blockIdx.y % tiles_k) * T;
    const int tap = blockIdx.z, sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lbeg = ch * p.lchunk, lend = min(p.L, lbeg + p.lchunk);
    const float *dYb = p.dY + (size_t)b * p.M * p.L;
    const float *Xb = p.X + (size_t)b * p.K * p.L;
    const float *ra = p.rowadd ? p.rowadd + (size_t)b * p.K : nullptr;
    float acc[4 * W][4 * W] = {};
    float va[NV], vb[NV];
    auto fetch = [&](int l0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = tid + i * NT;
            const int ll = e & (TK - 1), r = e / TK, l = l0 + ll, m = m0 + r, k = k0 + r, ls = l + sh;
            va[i] = (m < p.M && l < lend) ? dYb[(size_t)m * p.L + l] : 0.f;
            float v = 0.f;
            if (k < p.K && l < lend && ls >= 0 && ls < p.L) {
                v = Xb[(size_t)k * p.L + ls];
                if (ra) v += ra[k];
            }
            vb[i] = v;
        }
    };
    auto stage = [&]() {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int e = tid + i * NT;
            As[e & (TK - 1)][e / TK] = va[i];
            Bs[e & (TK - 1)][e / TK] = vb[i];
        }
    };
    fetch(lbeg);
    stage();
    __syncthreads();
    for (int l0 = lbeg; l0 < lend; l0 += TK) {
        const bool more = l0 + TK < lend;
        if (more) fetch(l0 + TK);
        tile_fma<W>(As, Bs, ty, tx, acc);
        __syncthreads();
        if (more) {
            stage();
            __syncthreads();
        }
    }
#pragma unroll
    for (int i = 0; i < 4 * W; ++i) {
        const int m = m0 + tile_pos(ty, i);
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4 * W; ++j) {
            const int k = k0 + tile_pos(tx, j);
            if (k >= p.K) continue;
            atomicAdd(&p.dW[tap * p.o_tap + m * p.o_m + k * p.o_k], p.alpha * acc[i][j]);
        }
    }
}

// ---- tensor-core forms of the two GEMM kernels ------------------------------------------------------------------
// Same contracts as cgemm_kernel / wgrad_kernel.  fp32 operands are split x = hi + lo into bf16 halves while they are
// staged into shared memory and every product is hi*hi + lo*hi + hi*lo on mma.sync.m16n8k16 with fp32 accumulation
// (the inference kernels' precision choice, DESIGN.md section 4: ~2^-17 per product).  One CTA = 128 x 128 outputs, 8 warps as
// 2 (rows) x 4 (columns), a warp = 64 x 32 = 4 x 4 MMA tiles; the reduction dimension advances 32 at a time.
constexpr int MT_ = 128, MK = 32, MSA = MK + 8, MSB = MT_ + 8;      // MSA / MSB: bf16 row strides that keep ldmatrix conflict free

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void *p) {
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two consecutive elements of a row -> packed bf16 hi and lo words
__device__ __forceinline__ void split_pair(float v0, float v1, __nv_bfloat16 *hi, __nv_bfloat16 *lo) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
    *reinterpret_cast<__nv_bfloat162 *>(hi) = __halves2bfloat162(h0, h1);
    *reinterpret_cast<__nv_bfloat162 *>(lo) = __halves2bfloat162(__float2bfloat16_rn(v0 - __bfloat162float(h0)),
                                                                 __float2bfloat16_rn(v1 - __bfloat162float(h1)));
}
// acc += A[128 x 32] * B[32 x 128] for this warp's 64 x 32 block.  A tiles: [row][k] (k contiguous).
// BT = false: B tiles are [k][col] (col contiguous, ldmatrix.trans);  BT = true: B tiles are [col][k] (k contiguous).
template <bool BT>
__device__ __forceinline__ void warp_mma_tile(const __nv_bfloat16 *Ah, const __nv_bfloat16 *Al, const __nv_bfloat16 *Bh,
                                              const __nv_bfloat16 *Bl, int wm, int wn, int lane, float (&acc)[4][4][4]) {
#pragma unroll
    for (int ks = 0; ks < MK / 16; ++ks) {
        uint32_t bh[2][4], bl[2][4];
#pragma unroll
        for (int n2 = 0; n2 < 2; ++n2) {
            if (BT) {       // matrices: (cols 0-7, k 0-7), (cols 0-7, k 8-15), (cols 8-15, k 0-7), (cols 8-15, k 8-15)
                const int col = wn * 32 + n2 * 16 + (lane & 7) + (lane >> 4) * 8, k = ks * 16 + ((lane >> 3) & 1) * 8;
                ldsm_x4(bh[n2], Bh + col * MSA + k);
                ldsm_x4(bl[n2], Bl + col * MSA + k);
            } else {        // rows k 0-7 / 8-15 at cols 0-7, then at cols 8-15, transposed on the way in
                const int k = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = wn * 32 + n2 * 16 + (lane >> 4) * 8;
                uint32_t a = (uint32_t)__cvta_generic_to_shared(Bh + k * MSB + col);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                             : "=r"(bh[n2][0]), "=r"(bh[n2][1]), "=r"(bh[n2][2]), "=r"(bh[n2][3]) : "r"(a));
                a = (uint32_t)__cvta_generic_to_shared(Bl + k * MSB + col);
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                             : "=r"(bl[n2][0]), "=r"(bl[n2][1]), "=r"(bl[n2][2]), "=r"(bl[n2][3]) : "r"(a));
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t ah[4], al[4];      // rows 0-7 / 8-15 at k 0-7, then at k 8-15
            const int row = wm * 64 + i * 16 + (lane & 15), k = ks * 16 + (lane >> 4) * 8;
            ldsm_x4(ah, Ah + row * MSA + k);
            ldsm_x4(al, Al + row * MSA + k);
#pragma unroll
            for (int n2 = 0; n2 < 2; ++n2)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    mma16816(acc[i][2 * n2 + h], ah, bh[n2][2 * h], bh[n2][2 * h + 1]);
                    mma16816(acc[i][2 * n2 + h], al, bh[n2][2 * h], bh[n2][2 * h + 1]);
                    mma16816(acc[i][2 * n2 + h], ah, bl[n2][2 * h], bl[n2][2 * h + 1]);
                }
        }
    }
}

__global__ void __launch_bounds__(NT) cgemm_mma_kernel(GemmArgs p) {
    __shared__ __align__(16) __nv_bfloat16 Ah[MT_ * MSA], Al[MT_ * MSA], Bh[MK * MSB], Bl[MK * MSB];
    const int b = blockIdx.z, m0 = blockIdx.y * MT_, l0 = blockIdx.x * MT_;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
    float acc[4][4][4] = {};
    const float *Xb = p.X + (size_t)b * p.K * p.L;
    const float *ra = p.rowadd ? p.rowadd + (size_t)b * p.K : nullptr;
    for (int tap = 0; tap < p.ntap; ++tap) {
        const int sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
        for (int k0 = 0; k0 < p.K; k0 += MK) {
            for (int e = tid; e < MT_ * MK / 2; e += NT) {             // A: weights, pairs along k
                const int kk = (e & (MK / 2 - 1)) * 2, mm = e / (MK / 2), m = m0 + mm, k = k0 + kk;
                float v0 = 0.f, v1 = 0.f;
                if (m < p.M) {
                    const float *a = p.A + tap * p.a_tap + m * p.a_m + k * p.a_k;
                    if (k < p.K) v0 = a[0];
                    if (k + 1 < p.K) v1 = a[p.a_k];
                }
                split_pair(v0, v1, Ah + mm * MSA + kk, Al + mm * MSA + kk);
            }
            for (int e = tid; e < MK * MT_ / 2; e += NT) {             // B: activations, pairs along time
                const int ll = (e & (MT_ / 2 - 1)) * 2, kk = e / (MT_ / 2), k = k0 + kk, l = l0 + ll + sh;
                float v0 = 0.f, v1 = 0.f;
                if (k < p.K) {
                    const float *x = Xb + (size_t)k * p.L;
                    const float r = ra ? ra[k] : 0.f;
                    if (l >= 0 && l < p.L) v0 = x[l] + r;
                    if (l + 1 >= 0 && l + 1 < p.L) v1 = x[l + 1] + r;
                }
                split_pair(v0, v1, Bh + kk * MSB + ll, Bl + kk * MSB + ll);
            }
            __syncthreads();
            warp_mma_tile<false>(Ah, Al, Bh, Bl, wm, wn, lane, acc);
            __syncthreads();
        }
    }
    const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int m = m0 + wm * 64 + i * 16 + g + hrow * 8;
            if (m >= p.M) continue;
            const float bm = p.bias ? p.bias_scale * p.bias[m] : 0.f;
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int l = l0 + wn * 32 + n * 8 + t2 + c;
                    if (l >= p.L) continue;
                    const size_t idx = ((size_t)b * p.M + m) * p.L + l;
                    float v = fmaf(p.alpha, acc[i][n][hrow * 2 + c], bm);
                    if (p.R) v = fmaf(p.beta, p.R[idx], v);
                    if (p.relu) v = fmaxf(v, 0.f);
                    p.Y[idx] = v;
                }
        }
}

__global__ void __launch_bounds__(NT) wgrad_mma_kernel(WgradArgs p) {
    __shared__ __align__(16) __nv_bfloat16 Ah[MT_ * MSA], Al[MT_ * MSA], Bh[MT_ * MSA], Bl[MT_ * MSA];
    const int nchunk = (p.L + p.lchunk - 1) / p.lchunk;
    const int b = blockIdx.x / nchunk, ch = blockIdx.x % nchunk;
    const int tiles_k = (p.K + MT_ - 1) / MT_;
    const int m0 = (blockIdx.y / tiles_k) * MT_, k0 = (blockIdx.y % tiles_k) * MT_;
    const int tap = blockIdx.z, sh = tap == 0 ? p.shift[0] : (tap == 1 ? p.shift[1] : p.shift[2]);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
    const int lbeg = ch * p.lchunk, lend = min(p.L, lbeg + p.lchunk);
    const float *dYb = p.dY + (size_t)b * p.M * p.L;
    const float *Xb = p.X + (size_t)b * p.K * p.L;
    const float *ra = p.rowadd ? p.rowadd + (size_t)b * p.K : nullptr;
    float acc[4][4][4] = {};
    for (int l0 = lbeg; l0 < lend; l0 += MK) {
        for (int e = tid; e < MT_ * MK / 2; e += NT) {
            const int ll = (e & (MK / 2 - 1)) * 2, r = e / (MK / 2), l = l0 + ll;
            float a0 = 0.f, a1 = 0.f, x0 = 0.f, x1 = 0.f;
            const int m = m0 + r, k = k0 + r;
            if (m < p.M) {
                if (l < lend) a0 = dYb[(size_t)m * p.L + l];
                if (l + 1 < lend) a1 = dYb[(size_t)m * p.L + l + 1];
            }
            if (k < p.K) {
                const float *x = Xb + (size_t)k * p.L;
                const float rr = ra ? ra[k] : 0.f;
                const int ls = l + sh;
                if (l < lend && ls >= 0 && ls < p.L) x0 = x[ls] + rr;
                if (l + 1 < lend && ls + 1 >= 0 && ls + 1 < p.L) x1 = x[ls + 1] + rr;
            }
            split_pair(a0, a1, Ah + r * MSA + ll, Al + r * MSA + ll);
            split_pair(x0, x1, Bh + r * MSA + ll, Bl + r * MSA + ll);
        }
        __syncthreads();
        warp_mma_tile<true>(Ah, Al, Bh, Bl, wm, wn, lane, acc);
        __syncthreads();
    }
    const int g = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            const int m = m0 + wm * 64 + i * 16 + g + hrow * 8;
            if (m >= p.M) continue;
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const int k = k0 + wn * 32 + n * 8 + t2 + c;
                    if (k >= p.K) continue;
                    atomicAdd(&p.dW[tap * p.o_tap + m * p.o_m + k * p.o_k], p.alpha * acc[i][n][hrow * 2 + c]);
                }
        }
}

__device__ __forceinline__ float block_sum(float v) {
    __shared__ float red[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    __syncthreads();
    if (ln == 0) red[w] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
    if (w == 0)
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;      // valid in thread 0
}

// out[r] = sum_l in[r,l]
__global__ void __launch_bounds__(NT) rowsum_kernel(const float *in, int L, float *out) {
    const float *row = in + (size_t)blockIdx.x * L;
    float s = 0.f;
    for (int l = threadIdx.x; l < L; l += NT) s += row[l];
    s = block_sum(s);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}
// out[m] += alpha * sum_b in[b,m]
__global__ void batchsum_kernel(const float *in, int B, int M, float alpha, float *out) {
    const int m = blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += in[(size_t)b * M + m];
    out[m] += alpha * s;
}

// G (B,2C,L): rows [0,C) -> tanh, rows [C,2C) -> sigmoid, in place; O = tanh * sigmoid
__global__ void gate_fwd_kernel(float *G, float *O, int C, int L, size_t n) {
    const size_t CL = (size_t)C * L;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / CL, r = i % CL;
        float *ga = G + b * 2 * CL + r, *gb = ga + CL;
        const float th = tanhf(*ga), sg = 1.0f / (1.0f + expf(-*gb));
        *ga = th;
        *gb = sg;
        O[i] = th * sg;
    }
}
__global__ void gate_bwd_kernel(const float *dO, const float *TS, float *dG, int C, int L, size_t n) {
    const size_t CL = (size_t)C * L;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / CL, r = i % CL;
        const float th = TS[b * 2 * CL + r], sg = TS[b * 2 * CL + CL + r], d = dO[i];
        dG[b * 2 * CL + r] = d * sg * (1.0f - th * th);
        dG[b * 2 * CL + CL + r] = d * th * sg * (1.0f - sg);
    }
}
__global__ void relu_mask_kernel(float *d, const float *act, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        if (!(act[i] > 0.f)) d[i] = 0.f;
}
__global__ void axpby_kernel(float *y, const float *x, float a, size_t n) {      // y = a*y + x
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = fmaf(a, y[i], x[i]);
}
// x_t = coef[b,0] * audio + coef[b,1] * z                                  (train.py:219)
__global__ void diffuse_kernel(const float *audio, const float *z, const float *coef, float *xt, int L, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / L;
        xt[i] = coef[2 * b] * audio[i] + coef[2 * b + 1] * z[i];
    }
}
// loss += sum (y - z)^2 / n;  dy = 2 (y - z) / n                           (nn.MSELoss, train.py:138)
__global__ void __launch_bounds__(NT) mse_kernel(const float *y, const float *z, float *dy, float *loss, size_t n) {
    float s = 0.f;
    const float inv = 1.0f / (float)n;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float d = y[i] - z[i];
        s = fmaf(d, d, s);
        dy[i] = 2.0f * d * inv;
    }
    s = block_sum(s);
    if (threadIdx.x == 0) atomicAdd(loss, s * inv);
}
// models/utils.py:4-29: e = [sin(t f), cos(t f)], f_i = exp(-i ln(1e4) / (half - 1))
__global__ void embed_kernel(const float *t, int B, int E, float *e0) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x, half = E / 2;
    if (i >= B * half) return;
    const int b = i / half, j = i % half;
    const float f = expf((float)j * -(logf(10000.0f) / (float)(half - 1)));
    const float a = t[b] * f;
    e0[(size_t)b * E + j] = sinf(a);
    e0[(size_t)b * E + half + j] = cosf(a);
}
// z = x W^T + b;  y = swish(z) or z
__global__ void linear_fwd_kernel(const float *x, const float *W, const float *bias, int B, int I, int O, float *z, float *y, int swish) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * O) return;
    const int b = i / O, o = i % O;
    float s = bias[o];
    for (int k = 0; k < I; ++k) s = fmaf(x[(size_t)b * I + k], W[(size_t)o * I + k], s);
    if (z) z[i] = s;
    y[i] = swish ? s / (1.0f + expf(-s)) : s;
}
__global__ void swish_bwd_kernel(float *d, const float *z, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float sg = 1.0f / (1.0f + expf(-z[i]));
    d[i] *= sg * (1.0f + z[i] * (1.0f - sg));
}
// dW[o,k] += sum_b dy[b,o] x[b,k];  db[o] += sum_b dy[b,o]
__global__ void linear_bwd_w_kernel(const float *dy, const float *x, int B, int I, int O, float *dW, float *db) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= O * I) return;
    const int o = i / I, k = i % I;
    float s = 0.f, sb = 0.f;
    for (int b = 0; b < B; ++b) {
        s = fmaf(dy[(size_t)b * O + o], x[(size_t)b * I + k], s);
        sb += dy[(size_t)b * O + o];
    }
    dW[i] += s;
    if (k == 0) db[o] += sb;
}
// dx[b,k] (+)= sum_o dy[b,o] W[o,k]
__global__ void linear_bwd_x_kernel(const float *dy, const float *W, int B, int I, int O, float *dx, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * I) return;
    const int b = i / I, k = i % I;
    float s = accumulate ? dx[i] : 0.f;
    for (int o = 0; o < O; ++o) s = fmaf(dy[(size_t)b * O + o], W[(size_t)o * I + k], s);
    dx[i] = s;
}
// weight norm, one CTA per output row (torch.nn.utils.weight_norm, dim = 0)
__global__ void __launch_bounds__(NT) wn_fwd_kernel(const float *v, const float *g, int rowlen, float *W) {
    const float *vr = v + (size_t)blockIdx.x * rowlen;
    float s = 0.f;
    for (int i = threadIdx.x; i < rowlen; i += NT) s = fmaf(vr[i], vr[i], s);
    __shared__ float scale;
    s = block_sum(s);
    if (threadIdx.x == 0) scale = g[blockIdx.x] / sqrtf(s);
    __syncthreads();
    for (int i = threadIdx.x; i < rowlen; i += NT) W[(size_t)blockIdx.x * rowlen + i] = scale * vr[i];
}
__global__ void __launch_bounds__(NT) wn_bwd_kernel(const float *v, const float *g, const float *dW, int rowlen, float *dv, float *dg) {
    const float *vr = v + (size_t)blockIdx.x * rowlen, *dr = dW + (size_t)blockIdx.x * rowlen;
    float s = 0.f, d = 0.f;
    for (int i = threadIdx.x; i < rowlen; i += NT) {
        s = fmaf(vr[i], vr[i], s);
        d = fmaf(dr[i], vr[i], d);
    }
    __shared__ float sh[2];
    s = block_sum(s);
    if (threadIdx.x == 0) sh[0] = s;
    d = block_sum(d);
    if (threadIdx.x == 0) sh[1] = d;
    __syncthreads();
    const float nrm = sqrtf(sh[0]), dot = sh[1], gg = g[blockIdx.x];
    const float a = gg / nrm, c = gg * dot / (nrm * nrm * nrm);
    if (threadIdx.x == 0) dg[blockIdx.x] += dot / nrm;
    for (int i = threadIdx.x; i < rowlen; i += NT) dv[(size_t)blockIdx.x * rowlen + i] += a * dr[i] - c * vr[i];
}
// torch.optim.Adam (amsgrad off, weight_decay 0), one launch over the flat buffers; bc1 = 1 - beta1^step,
// bc2s = sqrt(1 - beta2^step) computed on the host in double
__global__ void adam_kernel(float *p, const float *g, float *m, float *v, long long n, float gscale, float lr_over_bc1, float b1,
                            float b2, float bc2s, float eps) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * gscale;
        const float mi = fmaf(b1, m[i], (1.0f - b1) * gi);           // exp_avg.lerp_(grad, 1 - beta1)
        const float vi = fmaf(b2, v[i], (1.0f - b2) * gi * gi);      // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2s + eps;
        p[i] -= lr_over_bc1 * (mi / denom);
    }
}

// ---- parameter layout (net.parameters() order of models/wavenet.py) ------------------------------------------
struct PEntry {
    std::string name;
    int64_t off, numel;
};
struct WN {
    int64_t bias, g, v;
    int rows, rowlen;
};
struct Lin {
    int64_t w, b;
    int O, I;
};
struct Layer {
    Lin fct;
    WN dil, res, skip;
};
struct Layout {
    std::vector<PEntry> e;
    int64_t total = 0;
    WN init, fin0;
    Lin fc1, fc2, finz;
    std::vector<Layer> layers;
    int64_t add(const std::string &name, int64_t n) {
        e.push_back({name, total, n});
        total += n;
        return total - n;
    }
    WN wn(const std::string &p, int rows, int rowlen) {       // registration order after weight_norm: bias, weight_g, weight_v
        WN w;
        w.rows = rows;
        w.rowlen = rowlen;
        w.bias = add(p + ".bias", rows);
        w.g = add(p + ".weight_g", rows);
        w.v = add(p + ".weight_v", (int64_t)rows * rowlen);
        return w;
    }
    Lin lin(const std::string &p, int O, int I) {
        Lin l;
        l.O = O;
        l.I = I;
        l.w = add(p + ".weight", (int64_t)O * I);
        l.b = add(p + ".bias", O);
        return l;
    }
};

static int check_cfg(const dwb_config *cfg) {
    DWB_REQUIRE(cfg != nullptr, DWB_ERR_INVALID, "null config");
    DWB_REQUIRE(cfg->model == DWB_MODEL_WAVENET, DWB_ERR_UNSUPPORTED,
                "the training step is implemented for model=wavenet only (SaShiMi backward is not built)");
    DWB_REQUIRE(cfg->unconditional == 1, DWB_ERR_UNSUPPORTED, "the training step is implemented for unconditional models only");
    DWB_REQUIRE(cfg->res_channels > 0 && cfg->skip_channels > 0 && cfg->num_res_layers > 0 && cfg->dilation_cycle > 0 &&
                    cfg->embed_in >= 4 && cfg->embed_in % 2 == 0 && cfg->embed_mid > 0 && cfg->embed_out > 0,
                DWB_ERR_INVALID, "bad wavenet config");
    return DWB_OK;
}

static Layout build_layout(const dwb_config *cfg) {
    Layout lo;
    const int C = cfg->res_channels, S = cfg->skip_channels;
    lo.init = lo.wn("init_conv.0.conv", C, 1);
    lo.fc1 = lo.lin("residual_layer.fc_t1", cfg->embed_mid, cfg->embed_in);
    lo.fc2 = lo.lin("residual_layer.fc_t2", cfg->embed_out, cfg->embed_mid);
    for (int n = 0; n < cfg->num_res_layers; ++n) {
        const std::string p = "residual_layer.residual_blocks." + std::to_string(n);
        Layer ly;
        ly.fct = lo.lin(p + ".fc_t", C, cfg->embed_out);
        ly.dil = lo.wn(p + ".dilated_conv_layer.conv", 2 * C, 3 * C);
        ly.res = lo.wn(p + ".res_conv", C, C);
        ly.skip = lo.wn(p + ".skip_conv", S, C);
        lo.layers.push_back(ly);
    }
    lo.fin0 = lo.wn("final_conv.0.conv", S, S);
    lo.finz = lo.lin("final_conv.2.conv", 1, S);
    return lo;
}

struct Trainer {
    dwb_config cfg;
    int device, B, L;
    Layout lo;
    float *ws = nullptr;
    size_t ws_floats = 0;
    // activations kept for the backward
    float *H, *TS, *O, *SK, *F, *XT, *Y;
    // gradients in flight
    float *dY, *dF, *dS, *dH, *dO, *dU, *dG;
    // folded weights and their gradients (same offsets as the flat parameter buffer)
    float *eff, *deff;
    // step embedding
    float *e0, *z1, *e1, *z2, *e2, *part, *dpart, *de2, *de1, *tmp, *rs_dS;
    int64_t launches = 0;
    bool mma = false;       // GEMM family: split-bf16 tensor cores / exact fp32 (dwb_trainer_set_gemm, DWB_TRAIN_GEMM)
};

static inline int ew_grid(size_t n) { return (int)std::min<size_t>((n + 255) / 256, 148 * 16); }

}  // namespace train
}  // namespace dwb

using namespace dwb;
using namespace dwb::train;

#define TR_LAUNCH(tr, ...)        \
    do {                          \
        __VA_ARGS__;              \
        DWB_LAUNCH_CHECK();       \
        ++(tr)->launches;         \
    } while (0)

static int run_gemm(Trainer *tr, cudaStream_t st, const float *A, long long a_tap, long long a_m, long long a_k, const float *X,
                    const float *rowadd, const float *bias, const float *R, float *Y, int M, int K, int ntap, int s0, int s1,
                    int s2, float alpha, float bias_scale, float beta, int relu) {
    GemmArgs p;
    p.A = A; p.a_tap = a_tap; p.a_m = a_m; p.a_k = a_k;
    p.X = X; p.rowadd = rowadd; p.bias = bias; p.R = R; p.Y = Y;
    p.M = M; p.K = K; p.L = tr->L; p.ntap = ntap;
    p.shift[0] = s0; p.shift[1] = s1; p.shift[2] = s2;
    p.alpha = alpha; p.bias_scale = bias_scale; p.beta = beta; p.relu = relu;
    if (tr->mma && M >= 8 && K >= 8) {
        dim3 grid(ceil_div(tr->L, MT_), ceil_div(M, MT_), tr->B);
        TR_LAUNCH(tr, cgemm_mma_kernel<<<grid, NT, 0, st>>>(p));
        return DWB_OK;
    }
    // narrow outputs (the 1-channel head, tiny test models) keep the 64-wide tile
    const int W = (tile_w() == 2 && M > 64) ? 2 : 1, T = 64 * W;
    dim3 grid(ceil_div(tr->L, T), ceil_div(M, T), tr->B);
    if (train_prefetch()) {
        if (W == 2)
            TR_LAUNCH(tr, cgemm_pf_kernel<2><<<grid, NT, 0, st>>>(p));
        else
            TR_LAUNCH(tr, cgemm_pf_kernel<1><<<grid, NT, 0, st>>>(p));
    } else if (W == 2)
        TR_LAUNCH(tr, cgemm_kernel<2><<<grid, NT, 0, st>>>(p));
    else
        TR_LAUNCH(tr, cgemm_kernel<1><<<grid, NT, 0, st>>>(p));
    return DWB_OK;
}

static int run_wgrad(Trainer *tr, cudaStream_t st, const float *dY, const float *X, const float *rowadd, float *dW,
                     long long o_tap, long long o_m, long long o_k, int M, int K, int ntap, int s0, int s1, int s2, float alpha) {
    WgradArgs p;
    p.dY = dY; p.X = X; p.rowadd = rowadd; p.dW = dW;
    p.o_tap = o_tap; p.o_m = o_m; p.o_k = o_k;
    p.M = M; p.K = K; p.L = tr->L; p.ntap = ntap;
    p.shift[0] = s0; p.shift[1] = s1; p.shift[2] = s2;
    p.alpha = alpha;
    const bool mma = tr->mma && M >= 8 && K >= 8;
    const int W = (mma || (tile_w() == 2 && M > 64 && K > 64)) ? 2 : 1, T = 64 * W;
    // time chunk per CTA: at least two waves of CTAs (2 resident per SM x 148 SMs) however few output tiles the weight has
    // (a 128 x 128 res_conv is ONE tile), at most 2048 samples so that the atomics stay a small share
    const int per_chunk = tr->B * ceil_div(M, T) * ceil_div(K, T) * ntap;
    const int want_chunks = std::max(1, ceil_div(592, per_chunk));
    p.lchunk = std::min(2048, std::max(128, ceil_div(ceil_div(tr->L, want_chunks), MK) * MK));
    dim3 grid(tr->B * ceil_div(tr->L, p.lchunk), ceil_div(M, T) * ceil_div(K, T), ntap);
    if (mma)
        TR_LAUNCH(tr, wgrad_mma_kernel<<<grid, NT, 0, st>>>(p));
    else if (train_prefetch()) {
        if (W == 2)
            TR_LAUNCH(tr, wgrad_pf_kernel<2><<<grid, NT, 0, st>>>(p));
        else
            TR_LAUNCH(tr, wgrad_pf_kernel<1><<<grid, NT, 0, st>>>(p));
    } else if (W == 2)
        TR_LAUNCH(tr, wgrad_kernel<2><<<grid, NT, 0, st>>>(p));
    else
        TR_LAUNCH(tr, wgrad_kernel<1><<<grid, NT, 0, st>>>(p));
    return DWB_OK;
}

// db[m] += alpha * sum_{b,l} d[b,m,l]
static int run_bias_grad(Trainer *tr, cudaStream_t st, const float *d, int M, float alpha, float *db) {
    TR_LAUNCH(tr, rowsum_kernel<<<tr->B * M, NT, 0, st>>>(d, tr->L, tr->tmp));
    TR_LAUNCH(tr, batchsum_kernel<<<ceil_div(M, 128), 128, 0, st>>>(tr->tmp, tr->B, M, alpha, db));
    return DWB_OK;
}

#define TR_CALL(expr)                 \
    do {                              \
        int _rc = (expr);             \
        if (_rc != DWB_OK) return _rc; \
    } while (0)

extern "C" {

int dwb_trainer_layout(const dwb_config *cfg, int index, char *name, int name_cap, int64_t *offset, int64_t *numel, int *n_params,
                       int64_t *total) {
    TR_CALL(check_cfg(cfg));
    const Layout lo = build_layout(cfg);
    if (n_params) *n_params = (int)lo.e.size();
    if (total) *total = lo.total;
    if (index >= 0) {
        DWB_REQUIRE(index < (int)lo.e.size(), DWB_ERR_INVALID, "parameter index %d out of range (%d parameters)", index, (int)lo.e.size());
        const PEntry &e = lo.e[index];
        if (name) {
            DWB_REQUIRE((int)e.name.size() < name_cap, DWB_ERR_INVALID, "name buffer too small");
            strcpy(name, e.name.c_str());
        }
        if (offset) *offset = e.off;
        if (numel) *numel = e.numel;
    }
    return DWB_OK;
}

int dwb_trainer_create(const dwb_config *cfg, int device, int B, int L, dwb_trainer **out) {
    DWB_REQUIRE(out != nullptr, DWB_ERR_INVALID, "null out pointer");
    *out = nullptr;
    TR_CALL(check_cfg(cfg));
    DWB_REQUIRE(B > 0 && L > 0, DWB_ERR_INVALID, "bad batch %d / length %d", B, L);
    DWB_CUDA(cudaSetDevice(device));
    Trainer *tr = new Trainer();
    tr->cfg = *cfg;
    tr->device = device;
    tr->B = B;
    tr->L = L;
    tr->lo = build_layout(cfg);
    tr->mma = train_mma();
    const size_t C = cfg->res_channels, S = cfg->skip_channels, N = cfg->num_res_layers, BL = (size_t)B * L;
    const size_t Ei = cfg->embed_in, Em = cfg->embed_mid, Eo = cfg->embed_out, P = tr->lo.total;
    const size_t widest = std::max<size_t>(2 * C, S);
    size_t need = 0;
    auto take = [&](size_t n) { size_t o = need; need += (n + 63) & ~(size_t)63; return o; };
    const size_t oH = take((N + 1) * C * BL), oTS = take(N * 2 * C * BL), oO = take(N * C * BL), oSK = take(S * BL), oF = take(S * BL);
    const size_t oXT = take(BL), oY = take(BL), odY = take(BL), odF = take(S * BL), odS = take(S * BL), odH = take(C * BL);
    const size_t odO = take(C * BL), odU = take(C * BL), odG = take(2 * C * BL), oeff = take(P), odeff = take(P);
    const size_t oe0 = take(B * Ei), oz1 = take(B * Em), oe1 = take(B * Em), oz2 = take(B * Eo), oe2 = take(B * Eo);
    const size_t opart = take(N * B * C), odpart = take(N * B * C), ode2 = take(B * Eo), ode1 = take(B * Em);
    const size_t otmp = take(B * widest), ors = take(B * S);
    cudaError_t e = cudaMalloc(&tr->ws, need * sizeof(float));
    if (e != cudaSuccess) {
        delete tr;
        return cuda_fail(e, "cudaMalloc(training workspace)", __FILE__, __LINE__);
    }
    tr->ws_floats = need;
    float *w = tr->ws;
    tr->H = w + oH; tr->TS = w + oTS; tr->O = w + oO; tr->SK = w + oSK; tr->F = w + oF; tr->XT = w + oXT; tr->Y = w + oY;
    tr->dY = w + odY; tr->dF = w + odF; tr->dS = w + odS; tr->dH = w + odH; tr->dO = w + odO; tr->dU = w + odU; tr->dG = w + odG;
    tr->eff = w + oeff; tr->deff = w + odeff;
    tr->e0 = w + oe0; tr->z1 = w + oz1; tr->e1 = w + oe1; tr->z2 = w + oz2; tr->e2 = w + oe2;
    tr->part = w + opart; tr->dpart = w + odpart; tr->de2 = w + ode2; tr->de1 = w + ode1; tr->tmp = w + otmp; tr->rs_dS = w + ors;
    *out = reinterpret_cast<dwb_trainer *>(tr);
    return DWB_OK;
}

int dwb_trainer_destroy(dwb_trainer *t) {
    Trainer *tr = reinterpret_cast<Trainer *>(t);
    if (!tr) return DWB_OK;
    cudaFree(tr->ws);
    delete tr;
    return DWB_OK;
}

int dwb_trainer_set_gemm(dwb_trainer *t, int tensor_cores) {
    Trainer *tr = reinterpret_cast<Trainer *>(t);
    DWB_REQUIRE(tr != nullptr, DWB_ERR_INVALID, "null trainer");
    tr->mma = tensor_cores != 0;
    return DWB_OK;
}

int dwb_trainer_info(dwb_trainer *t, int64_t *workspace_bytes, int64_t *launches) {
    Trainer *tr = reinterpret_cast<Trainer *>(t);
    DWB_REQUIRE(tr != nullptr, DWB_ERR_INVALID, "null trainer");
    if (workspace_bytes) *workspace_bytes = (int64_t)(tr->ws_floats * sizeof(float));
    if (launches) *launches = tr->launches;
    return DWB_OK;
}

int dwb_trainer_loss_backward(dwb_trainer *t, const float *params, float *grads, const float *audio, const float *z,
                              const float *steps, const float *coef, float *eps_out, float *loss, void *stream) {
    Trainer *tr = reinterpret_cast<Trainer *>(t);
    DWB_REQUIRE(tr && params && grads && audio && z && steps && coef && loss, DWB_ERR_INVALID, "null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    DWB_CUDA(cudaSetDevice(tr->device));
    const Layout &lo = tr->lo;
    const dwb_config &cfg = tr->cfg;
    const int B = tr->B, L = tr->L, C = cfg.res_channels, S = cfg.skip_channels, N = cfg.num_res_layers;
    const size_t BL = (size_t)B * L, CBL = (size_t)C * BL, SBL = (size_t)S * BL;
    const float rs2 = (float)sqrt(0.5), cN = (float)sqrt(1.0 / N);
    const float *P = params;
    float *G = grads, *eff = tr->eff, *deff = tr->deff;

    DWB_CUDA(cudaMemsetAsync(G, 0, lo.total * sizeof(float), st));
    DWB_CUDA(cudaMemsetAsync(deff, 0, lo.total * sizeof(float), st));
    DWB_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    DWB_CUDA(cudaMemsetAsync(tr->SK, 0, SBL * sizeof(float), st));
    DWB_CUDA(cudaMemsetAsync(tr->dH, 0, CBL * sizeof(float), st));
    DWB_CUDA(cudaMemsetAsync(tr->de2, 0, (size_t)B * cfg.embed_out * sizeof(float), st));

    // ---- fold weight norm ----
    auto fold = [&](const WN &w) -> int {
        TR_LAUNCH(tr, wn_fwd_kernel<<<w.rows, NT, 0, st>>>(P + w.v, P + w.g, w.rowlen, eff + w.v));
        return DWB_OK;
    };
    TR_CALL(fold(lo.init));
    TR_CALL(fold(lo.fin0));
    for (const Layer &ly : lo.layers) {
        TR_CALL(fold(ly.dil));
        TR_CALL(fold(ly.res));
        TR_CALL(fold(ly.skip));
    }

    // ---- forward ----
    TR_LAUNCH(tr, diffuse_kernel<<<ew_grid(BL), 256, 0, st>>>(audio, z, coef, tr->XT, L, BL));
    TR_LAUNCH(tr, embed_kernel<<<ceil_div(B * cfg.embed_in / 2, 128), 128, 0, st>>>(steps, B, cfg.embed_in, tr->e0));
    TR_LAUNCH(tr, linear_fwd_kernel<<<ceil_div(B * cfg.embed_mid, 128), 128, 0, st>>>(tr->e0, P + lo.fc1.w, P + lo.fc1.b, B, cfg.embed_in,
                                                                                 cfg.embed_mid, tr->z1, tr->e1, 1));
    TR_LAUNCH(tr, linear_fwd_kernel<<<ceil_div(B * cfg.embed_out, 128), 128, 0, st>>>(tr->e1, P + lo.fc2.w, P + lo.fc2.b, B, cfg.embed_mid,
                                                                                 cfg.embed_out, tr->z2, tr->e2, 1));
    for (int n = 0; n < N; ++n)
        TR_LAUNCH(tr, linear_fwd_kernel<<<ceil_div(B * C, 128), 128, 0, st>>>(tr->e2, P + lo.layers[n].fct.w, P + lo.layers[n].fct.b, B,
                                                                         cfg.embed_out, C, nullptr, tr->part + (size_t)n * B * C, 0));
    // h_0 = ReLU(W_in x_t + b)                                               (wavenet.py:181,205)
    TR_CALL(run_gemm(tr, st, eff + lo.init.v, 0, 1, 1, tr->XT, nullptr, P + lo.init.bias, nullptr, tr->H, C, 1, 1, 0, 0, 0, 1.f, 1.f, 0.f, 1));
    for (int n = 0; n < N; ++n) {
        const Layer &ly = lo.layers[n];
        const int d = 1 << (n % cfg.dilation_cycle);
        float *Hn = tr->H + (size_t)n * CBL, *Hn1 = Hn + CBL, *TSn = tr->TS + (size_t)n * 2 * CBL, *On = tr->O + (size_t)n * CBL;
        const float *pn = tr->part + (size_t)n * B * C;
        TR_CALL(run_gemm(tr, st, eff + ly.dil.v, 1, 3LL * C, 3, Hn, pn, P + ly.dil.bias, nullptr, TSn, 2 * C, C, 3, -d, 0, d, 1.f, 1.f, 0.f, 0));
        TR_LAUNCH(tr, gate_fwd_kernel<<<ew_grid(CBL), 256, 0, st>>>(TSn, On, C, L, CBL));
        TR_CALL(run_gemm(tr, st, eff + ly.res.v, 0, C, 1, On, nullptr, P + ly.res.bias, Hn, Hn1, C, C, 1, 0, 0, 0, rs2, rs2, rs2, 0));
        TR_CALL(run_gemm(tr, st, eff + ly.skip.v, 0, C, 1, On, nullptr, P + ly.skip.bias, tr->SK, tr->SK, S, C, 1, 0, 0, 0, 1.f, 1.f, 1.f, 0));
    }
    // f = ReLU(W_f (s sqrt(1/N)) + b_f);  y = W_z f + b_z                     (wavenet.py:165,194-196)
    TR_CALL(run_gemm(tr, st, eff + lo.fin0.v, 0, S, 1, tr->SK, nullptr, P + lo.fin0.bias, nullptr, tr->F, S, S, 1, 0, 0, 0, cN, 1.f, 0.f, 1));
    TR_CALL(run_gemm(tr, st, P + lo.finz.w, 0, S, 1, tr->F, nullptr, P + lo.finz.b, nullptr, tr->Y, 1, S, 1, 0, 0, 0, 1.f, 1.f, 0.f, 0));
    if (eps_out) DWB_CUDA(cudaMemcpyAsync(eps_out, tr->Y, BL * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TR_LAUNCH(tr, mse_kernel<<<ew_grid(BL), NT, 0, st>>>(tr->Y, z, tr->dY, loss, BL));

    // ---- backward: output head ----
    TR_CALL(run_wgrad(tr, st, tr->dY, tr->F, nullptr, G + lo.finz.w, 0, S, 1, 1, S, 1, 0, 0, 0, 1.f));
    TR_CALL(run_bias_grad(tr, st, tr->dY, 1, 1.f, G + lo.finz.b));
    TR_CALL(run_gemm(tr, st, P + lo.finz.w, 0, 1, 0, tr->dY, nullptr, nullptr, nullptr, tr->dF, S, 1, 1, 0, 0, 0, 1.f, 0.f, 0.f, 0));
    TR_LAUNCH(tr, relu_mask_kernel<<<ew_grid(SBL), 256, 0, st>>>(tr->dF, tr->F, SBL));
    TR_CALL(run_wgrad(tr, st, tr->dF, tr->SK, nullptr, deff + lo.fin0.v, 0, S, 1, S, S, 1, 0, 0, 0, cN));
    TR_CALL(run_bias_grad(tr, st, tr->dF, S, 1.f, G + lo.fin0.bias));
    TR_CALL(run_gemm(tr, st, eff + lo.fin0.v, 0, 1, S, tr->dF, nullptr, nullptr, nullptr, tr->dS, S, S, 1, 0, 0, 0, cN, 0.f, 0.f, 0));
    TR_LAUNCH(tr, rowsum_kernel<<<B * S, NT, 0, st>>>(tr->dS, L, tr->rs_dS));

    // ---- backward: residual layers ----
    for (int n = N - 1; n >= 0; --n) {
        const Layer &ly = lo.layers[n];
        const int d = 1 << (n % cfg.dilation_cycle);
        const float *Hn = tr->H + (size_t)n * CBL, *TSn = tr->TS + (size_t)n * 2 * CBL, *On = tr->O + (size_t)n * CBL;
        const float *pn = tr->part + (size_t)n * B * C;
        // do = sqrt(1/2) W_r^T dh + W_s^T ds
        TR_CALL(run_gemm(tr, st, eff + ly.res.v, 0, 1, C, tr->dH, nullptr, nullptr, nullptr, tr->dO, C, C, 1, 0, 0, 0, rs2, 0.f, 0.f, 0));
        TR_CALL(run_gemm(tr, st, eff + ly.skip.v, 0, 1, C, tr->dS, nullptr, nullptr, tr->dO, tr->dO, C, S, 1, 0, 0, 0, 1.f, 0.f, 1.f, 0));
        TR_CALL(run_wgrad(tr, st, tr->dH, On, nullptr, deff + ly.res.v, 0, C, 1, C, C, 1, 0, 0, 0, rs2));
        TR_CALL(run_bias_grad(tr, st, tr->dH, C, rs2, G + ly.res.bias));
        TR_CALL(run_wgrad(tr, st, tr->dS, On, nullptr, deff + ly.skip.v, 0, C, 1, S, C, 1, 0, 0, 0, 1.f));
        TR_LAUNCH(tr, batchsum_kernel<<<ceil_div(S, 128), 128, 0, st>>>(tr->rs_dS, B, S, 1.f, G + ly.skip.bias));
        TR_LAUNCH(tr, gate_bwd_kernel<<<ew_grid(CBL), 256, 0, st>>>(tr->dO, TSn, tr->dG, C, L, CBL));
        // du[l] = sum_k W_k^T dG[l - (k-1) d]
        TR_CALL(run_gemm(tr, st, eff + ly.dil.v, 1, 3, 3LL * C, tr->dG, nullptr, nullptr, nullptr, tr->dU, C, 2 * C, 3, d, 0, -d, 1.f, 0.f, 0.f, 0));
        TR_CALL(run_wgrad(tr, st, tr->dG, Hn, pn, deff + ly.dil.v, 1, 3LL * C, 3, 2 * C, C, 3, -d, 0, d, 1.f));
        TR_CALL(run_bias_grad(tr, st, tr->dG, 2 * C, 1.f, G + ly.dil.bias));
        TR_LAUNCH(tr, rowsum_kernel<<<B * C, NT, 0, st>>>(tr->dU, L, tr->dpart + (size_t)n * B * C));
        TR_LAUNCH(tr, axpby_kernel<<<ew_grid(CBL), 256, 0, st>>>(tr->dH, tr->dU, rs2, CBL));
    }

    // ---- backward: input conv ----
    TR_LAUNCH(tr, relu_mask_kernel<<<ew_grid(CBL), 256, 0, st>>>(tr->dH, tr->H, CBL));
    TR_CALL(run_wgrad(tr, st, tr->dH, tr->XT, nullptr, deff + lo.init.v, 0, 1, 1, C, 1, 1, 0, 0, 0, 1.f));
    TR_CALL(run_bias_grad(tr, st, tr->dH, C, 1.f, G + lo.init.bias));

    // ---- backward: step embedding ----
    for (int n = 0; n < N; ++n) {
        const Lin &f = lo.layers[n].fct;
        const float *dp = tr->dpart + (size_t)n * B * C;
        TR_LAUNCH(tr, linear_bwd_w_kernel<<<ceil_div(f.O * f.I, 128), 128, 0, st>>>(dp, tr->e2, B, f.I, f.O, G + f.w, G + f.b));
        TR_LAUNCH(tr, linear_bwd_x_kernel<<<ceil_div(B * f.I, 128), 128, 0, st>>>(dp, P + f.w, B, f.I, f.O, tr->de2, 1));
    }
    TR_LAUNCH(tr, swish_bwd_kernel<<<ceil_div(B * cfg.embed_out, 128), 128, 0, st>>>(tr->de2, tr->z2, B * cfg.embed_out));
    TR_LAUNCH(tr, linear_bwd_w_kernel<<<ceil_div(lo.fc2.O * lo.fc2.I, 128), 128, 0, st>>>(tr->de2, tr->e1, B, lo.fc2.I, lo.fc2.O, G + lo.fc2.w, G + lo.fc2.b));
    TR_LAUNCH(tr, linear_bwd_x_kernel<<<ceil_div(B * lo.fc2.I, 128), 128, 0, st>>>(tr->de2, P + lo.fc2.w, B, lo.fc2.I, lo.fc2.O, tr->de1, 0));
    TR_LAUNCH(tr, swish_bwd_kernel<<<ceil_div(B * cfg.embed_mid, 128), 128, 0, st>>>(tr->de1, tr->z1, B * cfg.embed_mid));
    TR_LAUNCH(tr, linear_bwd_w_kernel<<<ceil_div(lo.fc1.O * lo.fc1.I, 128), 128, 0, st>>>(tr->de1, tr->e0, B, lo.fc1.I, lo.fc1.O, G + lo.fc1.w, G + lo.fc1.b));

    // ---- un-fold weight norm ----
    auto unfold = [&](const WN &w) -> int {
        TR_LAUNCH(tr, wn_bwd_kernel<<<w.rows, NT, 0, st>>>(P + w.v, P + w.g, deff + w.v, w.rowlen, G + w.v, G + w.g));
        return DWB_OK;
    };
    TR_CALL(unfold(lo.init));
    TR_CALL(unfold(lo.fin0));
    for (const Layer &ly : lo.layers) {
        TR_CALL(unfold(ly.dil));
        TR_CALL(unfold(ly.res));
        TR_CALL(unfold(ly.skip));
    }
    return DWB_OK;
}

int dwb_adam_step(float *params, const float *grads, float *exp_avg, float *exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                  float eps, int64_t step, float grad_scale, void *stream) {
    DWB_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0 && step >= 1, DWB_ERR_INVALID, "bad dwb_adam_step argument");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    adam_kernel<<<ew_grid((size_t)n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(params, grads, exp_avg, exp_avg_sq, (long long)n, grad_scale,
                                                                                   (float)((double)lr / bc1), beta1, beta2, (float)sqrt(bc2), eps);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

}  // extern "C"
