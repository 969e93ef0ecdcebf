// Per-step S4 long convolution: LayerNorm-apply + t-embedding prologue, FFT -> spectrum
// product -> inverse FFT, GELU epilogue.  One CTA per (batch, channel) row; the whole
// M = n/2 point complex transform lives in shared memory, in place.
//
// Reference: models/sashimi.py:148-152 (norm1 + fc_t), models/s4.py:1391-1411 (two-sided kernel,
// rfft/irfft product at n = 2l, D skip), :1430 (GELU).  The reference calls cuFFT three times per
// layer per step and regenerates rfft(k) every time; here the spectrum is cached
// (s4_kernelgen.cu) with D folded in, n is padded to a power of two >= 2l (wrapped kernel
// layout, identical result), and no cuFFT is involved.
//
// HBM traffic per row: read x (4l B) + write g (4l B); stats (8l B per batch element) and the
// spectrum (8(M+1) B per channel) are shared by H resp. B rows and stay in L2 (rows of one
// channel are adjacent in the grid).
#include <mutex>

#include "common.cuh"
#include "fft_plan.cuh"

namespace dwb {

// ---- in-register radix-R DFT, natural order in and out --------------------------------
template <int R, bool INV>
struct Radix {
    static __device__ __forceinline__ void run(float2 *x) {
        // omega_16^q = exp(-2 pi i q / 16), q = 0..7
        constexpr float WR[8] = {1.0f, 0.92387953251128674f, 0.70710678118654752f, 0.38268343236508977f,
                                 0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f};
        constexpr float WI[8] = {0.0f, -0.38268343236508977f, -0.70710678118654752f, -0.92387953251128674f,
                                 -1.0f, -0.92387953251128674f, -0.70710678118654752f, -0.38268343236508977f};
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) {
            e[i] = x[2 * i];
            o[i] = x[2 * i + 1];
        }
        Radix<R / 2, INV>::run(e);
        Radix<R / 2, INV>::run(o);
#pragma unroll
        for (int q = 0; q < R / 2; ++q) {
            constexpr int step = 16 / R;
            const float wr = WR[q * step], wi = INV ? -WI[q * step] : WI[q * step];
            const float2 t = make_float2(o[q].x * wr - o[q].y * wi, o[q].x * wi + o[q].y * wr);
            x[q] = make_float2(e[q].x + t.x, e[q].y + t.y);
            x[q + R / 2] = make_float2(e[q].x - t.x, e[q].y - t.y);
        }
    }
};
template <bool INV>
struct Radix<1, INV> {
    static __device__ __forceinline__ void run(float2 *) {}
};

// w[q] = w1^q, q = 1..R-1, by squaring/products of depth log2 R (a serial chain w *= w1 puts R-1
// dependent complex multiplies on the critical path of every butterfly)
template <int R>
__device__ __forceinline__ void twiddle_powers(float2 w1, float2 (&w)[R]) {
    w[0] = make_float2(1.f, 0.f);
    if (R > 1) w[1] = w1;
#pragma unroll
    for (int q = 2; q < R; ++q) {
        int hb = 1;
        while (hb * 2 <= q) hb *= 2;
        w[q] = (q == hb) ? cmul(w[q / 2], w[q / 2]) : cmul(w[hb], w[q - hb]);
    }
}

// ---- one in-place pass over the shared array ---------------------------------------------
// forward (DIF): u_q = sum_p x[j + p sub] w_R^{pq};  store u_q W_S^{jq} at j + q sub
// inverse (DIT): y_q = s[j + q sub] conj(W_S^{jq});  x_p = sum_q y_q w_R^{-pq} at j + p sub
template <int LOG2R, bool INV, int LOG2M, int NT>
__device__ __forceinline__ void fft_pass(float2 *s, const float2 *__restrict__ tw, int log2S, int tid) {
    constexpr int R = 1 << LOG2R, M = 1 << LOG2M;
    const int log2sub = log2S - LOG2R;
    const int sub = 1 << log2sub;
    for (int bi = tid; bi < M / R; bi += NT) {
        const int j = bi & (sub - 1);
        const int base = ((bi >> log2sub) << log2S) + j;
        float2 x[R];
#pragma unroll
        for (int p = 0; p < R; ++p) x[p] = s[fft_pad(base + (p << log2sub))];
        float2 w[R];
        if (log2sub > 0) {
            float2 w1 = tw[j << (LOG2M - log2S)];   // W_S^j = W_M^{j M / S}, shared-memory table
            if (INV) w1.y = -w1.y;
            twiddle_powers<R>(w1, w);
        }
        if (INV && log2sub > 0) {
#pragma unroll
            for (int q = 1; q < R; ++q) x[q] = cmul(x[q], w[q]);
        }
        Radix<R, INV>::run(x);
        if (!INV && log2sub > 0) {
#pragma unroll
            for (int q = 1; q < R; ++q) x[q] = cmul(x[q], w[q]);
        }
#pragma unroll
        for (int p = 0; p < R; ++p) s[fft_pad(base + (p << log2sub))] = x[p];
    }
    __syncthreads();
}

template <int LOG2M, int NT, bool INV, int PASS>
__device__ __forceinline__ void fft_run_pass(float2 *s, const float2 *__restrict__ tw, int tid) {
    constexpr int lr = fft_radix_log2(LOG2M, PASS);
    if constexpr (lr > 0) {
        // span of pass p = M / (R_0 ... R_{p-1})
        int log2S = LOG2M;
#pragma unroll
        for (int q = 0; q < PASS; ++q) log2S -= fft_radix_log2(LOG2M, q);
        fft_pass<lr, INV, LOG2M, NT>(s, tw, log2S, tid);
    }
}

template <int LOG2M>
struct FftCfg {
    static constexpr int M = 1 << LOG2M;
    static constexpr int NT = (M / 16 > 512) ? 512 : ((M / 16 < 64) ? 64 : M / 16);
    static constexpr int NTW = M / 16;                         // W_M^j, j < M/16: enough for every twiddled pass
    static constexpr int SDATA = M + M / 16 + 1;               // padded data array (float2)
    static constexpr int SMEM = (SDATA + NTW) * (int)sizeof(float2);
};

template <int LOG2M>
__global__ void __launch_bounds__(FftCfg<LOG2M>::NT)
fftconv_kernel(const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ part_t,
               long long part_stride_b, float ln_m, float ln_s, const float2 *__restrict__ kf,
               const float2 *__restrict__ tw /* W_n^i, i < M */, const float2 *__restrict__ twpos /* W_n^{freq(p)} */,
               float *__restrict__ g, int B, int H, int l) {
    constexpr int M = 1 << LOG2M, NT = FftCfg<LOG2M>::NT;
    extern __shared__ float2 s[];
    float2 *stw = s + FftCfg<LOG2M>::SDATA;
    const int tid = threadIdx.x;
    for (int j = tid; j < FftCfg<LOG2M>::NTW; j += NT) stw[j] = tw[2 * j];
    const int row = blockIdx.x;
    const int h = row / B, b = row - h * B;
    const size_t off = ((size_t)b * H + h) * (size_t)l;
    const float *xr = x + off;
    float *gr = g + off;
    const float pt = part_t ? part_t[(size_t)b * part_stride_b + h] : 0.f;
    const float *st = stats ? stats + (size_t)b * l * 2 : nullptr;

    // ---- prologue: y = (ln_s rstd)(x - mean + ln_m) + part_t, packed z[j] = y[2j] + i y[2j+1]
    const bool vec = ((l & 1) == 0);
    if (vec && st) {
        // hot case: batches of 4 independent (x, stats) loads in flight per thread
        const int half = l >> 1;
        const float2 *x2 = reinterpret_cast<const float2 *>(xr);
        const float4 *s4 = reinterpret_cast<const float4 *>(st);
        for (int j0 = tid; j0 < M; j0 += 4 * NT) {
            float2 xv[4];
            float4 sv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * NT;
                if (j < half) {
                    xv[u] = x2[j];
                    sv[u] = s4[j];
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j0 + u * NT;
                if (j < M) {
                    float2 v = make_float2(0.f, 0.f);
                    if (j < half) {
                        v.x = (ln_s * sv[u].y) * (xv[u].x - sv[u].x + ln_m) + pt;
                        v.y = (ln_s * sv[u].w) * (xv[u].y - sv[u].z + ln_m) + pt;
                    }
                    s[fft_pad(j)] = v;
                }
            }
        }
    } else
    for (int j = tid; j < M; j += NT) {
        float2 v = make_float2(0.f, 0.f);
        const int t0 = 2 * j;
        if (vec) {
            if (t0 < l) {
                const float2 xv = *reinterpret_cast<const float2 *>(xr + t0);
                v.x = xv.x + pt;
                v.y = xv.y + pt;
            }
        } else {
            if (t0 < l) v.x = st ? (ln_s * st[2 * t0 + 1]) * (xr[t0] - st[2 * t0] + ln_m) + pt : xr[t0] + pt;
            if (t0 + 1 < l)
                v.y = st ? (ln_s * st[2 * t0 + 3]) * (xr[t0 + 1] - st[2 * t0 + 2] + ln_m) + pt : xr[t0 + 1] + pt;
        }
        s[fft_pad(j)] = v;
    }
    __syncthreads();

    // ---- forward passes (natural -> digit reversed)
    fft_run_pass<LOG2M, NT, false, 0>(s, stw, tid);
    fft_run_pass<LOG2M, NT, false, 1>(s, stw, tid);
    fft_run_pass<LOG2M, NT, false, 2>(s, stw, tid);
    fft_run_pass<LOG2M, NT, false, 3>(s, stw, tid);

    // ---- untangle the real spectrum, multiply by the cached kernel spectrum, re-tangle.
    // Slot p holds Z[k], k = fft_freq(p); its partner Z[M-k] sits in slot fft_pos(M-k).
    // Leaders are the slots whose k < M/2: bit (lrl-1) of p clear (the last pass's digit is the
    // most significant digit of k).  Leader index 0 is k = 0 (DC + Nyquist, both real).
    {
        constexpr int np = fft_num_passes(LOG2M);
        constexpr int lrl = fft_radix_log2(LOG2M, np - 1);
        const float2 *kfr = kf + (size_t)h * (M + 1);
        for (int idx = tid; idx <= M / 2; idx += NT) {
            if (idx == 0) {
                const float2 a = s[0];
                const float y0 = 2.f * (a.x + a.y), yM = 2.f * (a.x - a.y);
                const float p0 = y0 * kfr[0].x, pM = yM * kfr[M].x;
                s[0] = make_float2(p0 + pM, p0 - pM);
                continue;
            }
            int p, p2;
            float2 w;
            if (idx == M / 2) {          // k = M/2, self-paired, W_n^{M/2} = -i
                p = p2 = 1 << (lrl - 1);
                w = make_float2(0.f, -1.f);
            } else {
                p = ((idx >> (lrl - 1)) << lrl) | (idx & ((1 << (lrl - 1)) - 1));
                const int k = fft_freq(p, LOG2M);
                p2 = fft_pos(M - k, LOG2M);
                w = twpos[p];
            }
            const float2 A = s[fft_pad(p)], Bv = s[fft_pad(p2)];
            const float2 K1 = kfr[p], K2 = kfr[p2];
            const float2 S = make_float2(A.x + Bv.x, A.y - Bv.y);
            const float2 Dm = make_float2(A.x - Bv.x, A.y + Bv.y);
            const float2 WD = cmul(w, Dm);
            const float2 T = make_float2(WD.y, -WD.x);                     // -i W Dm
            const float2 Yk = cadd(S, T);
            const float2 Yk2 = make_float2(S.x - T.x, -(S.y - T.y));       // conj(S - T)
            const float2 P1 = cmul(Yk, K1), P2 = cmul(Yk2, K2);
            const float2 S2 = make_float2(P1.x + P2.x, P1.y - P2.y);
            const float2 D2 = make_float2(P1.x - P2.x, P1.y + P2.y);
            const float2 CD = cmul_conj(D2, w);                            // conj(W) D2
            const float2 T2 = make_float2(-CD.y, CD.x);                    // i conj(W) D2
            s[fft_pad(p)] = cadd(S2, T2);
            s[fft_pad(p2)] = make_float2(S2.x - T2.x, -(S2.y - T2.y));     // conj(S2 - T2)
        }
    }
    __syncthreads();

    // ---- inverse passes (digit reversed -> natural), mirror order
    fft_run_pass<LOG2M, NT, true, 3>(s, stw, tid);
    fft_run_pass<LOG2M, NT, true, 2>(s, stw, tid);
    fft_run_pass<LOG2M, NT, true, 1>(s, stw, tid);
    fft_run_pass<LOG2M, NT, true, 0>(s, stw, tid);

    // ---- epilogue: first l samples, GELU
    if (vec) {
        for (int j = tid; j < l / 2; j += NT) {
            const float2 v = s[fft_pad(j)];
            *reinterpret_cast<float2 *>(gr + 2 * j) = make_float2(gelu_fast(v.x), gelu_fast(v.y));
        }
    } else {
        for (int j = tid; 2 * j < l; j += NT) {
            const float2 v = s[fft_pad(j)];
            gr[2 * j] = gelu_fast(v.x);
            if (2 * j + 1 < l) gr[2 * j + 1] = gelu_fast(v.y);
        }
    }
}

// ---- twiddle tables: immutable, per (device, log2M), created on first use -----------------
__global__ void twiddle_kernel(float2 *tw, float2 *twpos, int log2M) {
    const int M = 1 << log2M, n = 2 * M;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double sn, cs;
    sincospi(-2.0 * (double)i / (double)n, &sn, &cs);
    tw[i] = make_float2((float)cs, (float)sn);
    const int k = fft_freq(i, log2M);
    sincospi(-2.0 * (double)k / (double)n, &sn, &cs);
    twpos[i] = make_float2((float)cs, (float)sn);
}

struct TwiddleCache {
    std::mutex mu;
    float2 *tw[16][FFT_MAX_LOG2M + 1] = {};
    float2 *twpos[16][FFT_MAX_LOG2M + 1] = {};
};
static TwiddleCache g_tw;

int fft_twiddles(int log2M, cudaStream_t st, const float2 **tw, const float2 **twpos) {
    int dev = 0;
    DWB_CUDA(cudaGetDevice(&dev));
    DWB_REQUIRE(dev >= 0 && dev < 16, DWB_ERR_UNSUPPORTED, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_tw.mu);
    if (!g_tw.tw[dev][log2M]) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        DWB_CUDA(cudaStreamIsCapturing(st, &cs));
        DWB_REQUIRE(cs == cudaStreamCaptureStatusNone, DWB_ERR_STATE,
                    "fft twiddle table for M=2^%d must be created before stream capture", log2M);
        const int M = 1 << log2M;
        float2 *a = nullptr, *bq = nullptr;
        DWB_CUDA(cudaMalloc(&a, (size_t)M * sizeof(float2)));
        DWB_CUDA(cudaMalloc(&bq, (size_t)M * sizeof(float2)));
        twiddle_kernel<<<ceil_div(M, 256), 256, 0, st>>>(a, bq, log2M);
        DWB_LAUNCH_CHECK();
        DWB_CUDA(cudaStreamSynchronize(st));
        g_tw.tw[dev][log2M] = a;
        g_tw.twpos[dev][log2M] = bq;
    }
    *tw = g_tw.tw[dev][log2M];
    *twpos = g_tw.twpos[dev][log2M];
    return DWB_OK;
}

template <int LOG2M>
static int launch_fftconv(const float *x, const float *stats, const float *part_t, long long psb, float ln_m,
                          float ln_s, const float *kf, const float2 *tw, const float2 *twpos, float *g, int B, int H,
                          int l, cudaStream_t st) {
    using Cfg = FftCfg<LOG2M>;
    static bool attr_set[16] = {};
    int dev = 0;
    DWB_CUDA(cudaGetDevice(&dev));
    if (Cfg::SMEM > 48 * 1024 && !attr_set[dev & 15]) {
        DWB_CUDA(cudaFuncSetAttribute(fftconv_kernel<LOG2M>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_set[dev & 15] = true;
    }
    fftconv_kernel<LOG2M><<<B * H, Cfg::NT, Cfg::SMEM, st>>>(x, stats, part_t, psb, ln_m, ln_s, (const float2 *)kf, tw,
                                                            twpos, g, B, H, l);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

int fftconv_launch(const float *x, const float *stats, const float *part_t, long long psb, float ln_m, float ln_s,
                   const float *kf, float *g, int B, int H, int l, cudaStream_t st) {
    const int lg = fft_log2m_for(l);
    DWB_REQUIRE(lg > 0, DWB_ERR_UNSUPPORTED, "fftconv: stage length %d unsupported (max %d)", l, 1 << FFT_MAX_LOG2M);
    const float2 *tw, *twpos;
    int rc = fft_twiddles(lg, st, &tw, &twpos);
    if (rc != DWB_OK) return rc;
#define DWB_FFT_CASE(LG) \
    case LG:             \
        return launch_fftconv<LG>(x, stats, part_t, psb, ln_m, ln_s, kf, tw, twpos, g, B, H, l, st);
    switch (lg) {
        DWB_FFT_CASE(4) DWB_FFT_CASE(5) DWB_FFT_CASE(6) DWB_FFT_CASE(7) DWB_FFT_CASE(8) DWB_FFT_CASE(9)
        DWB_FFT_CASE(10) DWB_FFT_CASE(11) DWB_FFT_CASE(12) DWB_FFT_CASE(13) DWB_FFT_CASE(14)
    }
#undef DWB_FFT_CASE
    set_error("fftconv: no kernel for log2M=%d", lg);
    return DWB_ERR_UNSUPPORTED;
}

}  // namespace dwb

extern "C" int dwb_fftconv(const float *x, const float *stats, const float *part_t, int64_t part_stride_b, float ln_m,
                           float ln_s, const float *kf, float *g, int B, int H, int l, void *stream) {
    DWB_REQUIRE(x && kf && g, DWB_ERR_INVALID, "dwb_fftconv: null pointer");
    DWB_REQUIRE(B >= 1 && H >= 1 && l >= 1, DWB_ERR_INVALID, "dwb_fftconv: bad sizes B=%d H=%d l=%d", B, H, l);
    return dwb::fftconv_launch(x, stats, part_t, (long long)part_stride_b, ln_m, ln_s, kf, g, B, H, l,
                               (cudaStream_t)stream);
}
