// The rest of the reference's native FFI surface (extensions/cauchy/cauchy.cpp:86-95): non-symmetric forward and
// both backward ops, complex64 as interleaved float pairs.  dwb_cauchy_sym_fwd lives in s4_kernelgen.cu.
//
//   fwd      out[b,l] = sum_n v[b,n] / (z[l] - w[b,n])                                   (cauchy_cuda.cu:66-137)
//   bwd      dv[b,n] = sum_l dout[b,l] / conj(z[l] - w[b,n])
//            dw[b,n] = conj(v[b,n]) sum_l dout[b,l] / conj(z[l] - w[b,n])^2              (cauchy_cuda.cu:139-239)
//   sym_bwd  with r1 = 1/conj(z - w), r2 = 1/(z - conj w):
//            dv = sum_l dout r1 + conj(dout) r2 ;  dw = conj(v) sum_l dout r1^2 + conj(dout) r2^2   (:377-487)
// (PyTorch's complex autograd convention: the returned gradients are conjugate Wirtinger derivatives.)
//
// Backward layout: one CTA per (batch row, NPB consecutive states); its 256 threads stride over l, so dout and z are
// read once per NPB states instead of once per state, partial sums stay in registers, warp shuffles + one shared
// exchange finish the reduction.  No atomics, no (batch, N, chunks) temporary + .sum(-1) as in the reference.
#include "common.cuh"

namespace dwb {

constexpr int CB_THREADS = 256;
constexpr int CB_NPB = 4;

__device__ __forceinline__ float2 crecip(float ax, float ay) {     // 1 / (ax + i ay); MUFU.RCP (1 ulp), no IEEE division sequence
    float inv;
    asm("rcp.approx.f32 %0, %1;" : "=f"(inv) : "f"(fmaf(ax, ax, ay * ay)));
    return make_float2(ax * inv, -ay * inv);
}

template <bool SYM>
__global__ void __launch_bounds__(CB_THREADS)
cauchy_bwd_kernel(const float2 *__restrict__ v, const float2 *__restrict__ z, const float2 *__restrict__ w,
                  const float2 *__restrict__ dout, float2 *__restrict__ dv, float2 *__restrict__ dw, int N, int L) {
    __shared__ float2 red[CB_THREADS / 32][CB_NPB][2];
    const int b = blockIdx.y, n0 = blockIdx.x * CB_NPB;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float2 wn[CB_NPB];
#pragma unroll
    for (int i = 0; i < CB_NPB; ++i) wn[i] = n0 + i < N ? w[(size_t)b * N + n0 + i] : make_float2(1e30f, 0.f);
    float2 av[CB_NPB], aw[CB_NPB];
#pragma unroll
    for (int i = 0; i < CB_NPB; ++i) av[i] = aw[i] = make_float2(0.f, 0.f);
    const float2 *dr = dout + (size_t)b * L;
    for (int l = tid; l < L; l += CB_THREADS) {
        const float2 d = dr[l], zz = z[l];
#pragma unroll
        for (int i = 0; i < CB_NPB; ++i) {
            // r1 = 1 / conj(z - w)
            const float2 r1 = crecip(zz.x - wn[i].x, -(zz.y - wn[i].y));
            const float2 t1 = cmul(d, r1);
            av[i] = cadd(av[i], t1);
            aw[i] = cadd(aw[i], cmul(t1, r1));
            if (SYM) {
                // r2 = 1 / (z - conj w)
                const float2 r2 = crecip(zz.x - wn[i].x, zz.y + wn[i].y);
                const float2 t2 = cmul(cconj(d), r2);
                av[i] = cadd(av[i], t2);
                aw[i] = cadd(aw[i], cmul(t2, r2));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < CB_NPB; ++i) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            av[i].x += __shfl_down_sync(0xffffffffu, av[i].x, off);
            av[i].y += __shfl_down_sync(0xffffffffu, av[i].y, off);
            aw[i].x += __shfl_down_sync(0xffffffffu, aw[i].x, off);
            aw[i].y += __shfl_down_sync(0xffffffffu, aw[i].y, off);
        }
        if (lane == 0) {
            red[warp][i][0] = av[i];
            red[warp][i][1] = aw[i];
        }
    }
    __syncthreads();
    if (tid < CB_NPB && n0 + tid < N) {
        float2 sv = make_float2(0.f, 0.f), sw = make_float2(0.f, 0.f);
#pragma unroll
        for (int q = 0; q < CB_THREADS / 32; ++q) {
            sv = cadd(sv, red[q][tid][0]);
            sw = cadd(sw, red[q][tid][1]);
        }
        const size_t o = (size_t)b * N + n0 + tid;
        dv[o] = sv;
        dw[o] = cmul(sw, cconj(v[o]));
    }
}

// LPL lanes share one output l (state dimension split, warp-shuffle reduced) - same scheme as cauchy_sym_fwd_kernel
template <int LPL>
__global__ void __launch_bounds__(CB_THREADS)
cauchy_fwd_kernel(const float2 *__restrict__ v, const float2 *__restrict__ z, const float2 *__restrict__ w,
                  float2 *__restrict__ out, int N, int L) {
    __shared__ float2 sv[256], sw[256];
    const int b = blockIdx.y;
    const int lane = threadIdx.x % LPL;
    const int li = blockIdx.x * (CB_THREADS / LPL) + threadIdx.x / LPL;
    const bool live = li < L;
    const float2 zz = live ? z[li] : make_float2(0.f, 0.f);
    float2 acc = make_float2(0.f, 0.f);
    for (int c0 = 0; c0 < N; c0 += 256) {
        const int cnt = min(256, N - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < cnt; i += CB_THREADS) {
            sv[i] = v[(size_t)b * N + c0 + i];
            sw[i] = w[(size_t)b * N + c0 + i];
        }
        __syncthreads();
        if (live)
            for (int i = lane; i < cnt; i += LPL) acc = cadd(acc, cmul(sv[i], crecip(zz.x - sw[i].x, zz.y - sw[i].y)));
    }
#pragma unroll
    for (int off = LPL / 2; off > 0; off >>= 1) {
        acc.x += __shfl_down_sync(0xffffffffu, acc.x, off, LPL);
        acc.y += __shfl_down_sync(0xffffffffu, acc.y, off, LPL);
    }
    if (live && lane == 0) out[(size_t)b * L + li] = acc;
}

static int bwd_launch(bool sym, const float *v, const float *z, const float *w, const float *dout, float *dv, float *dw,
                      int batch, int N, int L, cudaStream_t st, const char *who) {
    DWB_REQUIRE(v && z && w && dout && dv && dw, DWB_ERR_INVALID, "%s: null pointer", who);
    DWB_REQUIRE(batch >= 0 && N >= 1 && L >= 0, DWB_ERR_INVALID, "%s: bad sizes batch=%d N=%d L=%d", who, batch, N, L);
    DWB_REQUIRE(batch <= 65535, DWB_ERR_UNSUPPORTED, "%s: batch %d > 65535", who, batch);
    if (batch == 0) return DWB_OK;
    const dim3 grid(ceil_div(N, CB_NPB), batch);
    if (sym)
        cauchy_bwd_kernel<true><<<grid, CB_THREADS, 0, st>>>((const float2 *)v, (const float2 *)z, (const float2 *)w,
                                                            (const float2 *)dout, (float2 *)dv, (float2 *)dw, N, L);
    else
        cauchy_bwd_kernel<false><<<grid, CB_THREADS, 0, st>>>((const float2 *)v, (const float2 *)z, (const float2 *)w,
                                                             (const float2 *)dout, (float2 *)dv, (float2 *)dw, N, L);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

}  // namespace dwb

using namespace dwb;

extern "C" int dwb_cauchy_fwd(const float *v, const float *z, const float *w, float *out, int batch, int N, int L,
                              void *stream) {
    DWB_REQUIRE(v && z && w && out, DWB_ERR_INVALID, "dwb_cauchy_fwd: null pointer");
    DWB_REQUIRE(batch >= 0 && N >= 1 && L >= 0, DWB_ERR_INVALID, "dwb_cauchy_fwd: bad sizes batch=%d N=%d L=%d", batch, N, L);
    DWB_REQUIRE(batch <= 65535, DWB_ERR_UNSUPPORTED, "dwb_cauchy_fwd: batch %d > 65535", batch);
    if (batch == 0 || L == 0) return DWB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const float2 *v2 = (const float2 *)v, *z2 = (const float2 *)z, *w2 = (const float2 *)w;
    if (N >= 16)
        cauchy_fwd_kernel<4><<<dim3(ceil_div(L, CB_THREADS / 4), batch), CB_THREADS, 0, st>>>(v2, z2, w2, (float2 *)out, N, L);
    else
        cauchy_fwd_kernel<1><<<dim3(ceil_div(L, CB_THREADS), batch), CB_THREADS, 0, st>>>(v2, z2, w2, (float2 *)out, N, L);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

extern "C" int dwb_cauchy_bwd(const float *v, const float *z, const float *w, const float *dout, float *dv, float *dw,
                              int batch, int N, int L, void *stream) {
    return bwd_launch(false, v, z, w, dout, dv, dw, batch, N, L, (cudaStream_t)stream, "dwb_cauchy_bwd");
}

extern "C" int dwb_cauchy_sym_bwd(const float *v, const float *z, const float *w, const float *dout, float *dv,
                                  float *dw, int batch, int N, int L, void *stream) {
    return bwd_launch(true, v, z, w, dout, dv, dw, batch, N, L, (cudaStream_t)stream, "dwb_cauchy_sym_bwd");
}
