// Per-step S4 long convolution: LayerNorm-apply + t-embedding prologue, FFT -> spectrum
// product -> inverse FFT, GELU epilogue.  One CTA per (batch, channel) row; the whole
// M = n/2 point complex transform lives in shared memory, in place.
//
// Reference: models/sashimi.py:148-152 (norm1 + fc_t), models/s4.py:1391-1411 (two-sided kernel,
// rfft/irfft product at n = 2l, D skip), :1430 (GELU).  The reference calls cuFFT three times per
// layer per step and regenerates rfft(k) every time; here the kernel spectrum is cached
// (s4_kernelgen.cu) with D folded in, n is padded to a power of two >= 2l (wrapped kernel
// layout, identical result), and no cuFFT is involved.
//
// Shared-memory traffic and barriers, not flops, bound this kernel, so the passes are fused at
// both ends and in the middle:
//   * pass 0 forward reads its 16 inputs straight from HBM (LN apply + t-embedding in registers;
//     the upper half of the zero-padded transform is never materialised: x[8..15] = 0 is folded
//     into the butterfly) and the last inverse pass writes GELU(y) straight to HBM (only the
//     first l outputs are computed);
//   * when the last forward pass has radix 2/4/8 it is done in registers together with the
//     real-FFT untangle + spectrum product + re-tangle and the first inverse pass: the outputs
//     are stored bit-reversed per pass, so the partner X[M-k] of slot p sits in slot
//     p ^ ((1 << msb(p)) - 1) and whole radix groups pair up (g, partner(g));
//   * untangle, product and re-tangle are one 2x2 complex map per pair whose four coefficients are
//     precomputed at weight load (kcoef_kernel): 16 FMA per pair instead of ~40 flops + 3 loads.
//
// HBM traffic per row: read x (4l B) + write g (4l B); stats (8l B per batch element) and the
// coefficient table (16(M+2) B per channel) are shared by H resp. B rows and stay in L2 (rows of
// one channel are adjacent in the grid).
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "fft_plan.cuh"
#include "fft_radix.cuh"
#include "kernels.h"

namespace dwb {

// ---- one in-place pass over the shared array ---------------------------------------------
// forward (DIF): u_q = sum_p x[j + p sub] w_R^{pq};  store u_q W_S^{jq} at j + brev(q) sub
// inverse (DIT): y_q = s[j + brev(q) sub] conj(W_S^{jq});  x_p = sum_q y_q w_R^{-pq} at j + p sub
template <int LR, bool INV, int LOG2M, int LOG2S, int NT>
__device__ __forceinline__ void fft_pass(float2 *s, const float2 *__restrict__ stw, int tid) {
    constexpr int R = 1 << LR, M = 1 << LOG2M, log2sub = LOG2S - LR, sub = 1 << log2sub;
#pragma unroll 1
    for (int bi = tid; bi < M / R; bi += NT) {
        const int j = bi & (sub - 1);
        const int base = ((bi >> log2sub) << LOG2S) + j;
        float2 x[R];
        float2 w1 = make_float2(1.f, 0.f);
        if (log2sub > 0) {
            w1 = stw[j << (LOG2M - LOG2S)];          // W_S^j = W_M^{j M / S}
            if (INV) w1.y = -w1.y;
        }
        if (!INV) {
#pragma unroll
            for (int p = 0; p < R; ++p) x[p] = s[fft_pad(base + (p << log2sub))];
            Radix<R, false>::run(x);
            if (log2sub > 0) apply_twiddles<R>(x, w1);
#pragma unroll
            for (int q = 0; q < R; ++q) s[fft_pad(base + (fft_brev(q, LR) << log2sub))] = x[q];
        } else {
#pragma unroll
            for (int q = 0; q < R; ++q) x[q] = s[fft_pad(base + (fft_brev(q, LR) << log2sub))];
            if (log2sub > 0) apply_twiddles<R>(x, w1);
            Radix<R, true>::run(x);
#pragma unroll
            for (int p = 0; p < R; ++p) s[fft_pad(base + (p << log2sub))] = x[p];
        }
    }
    __syncthreads();
}

// middle passes p = FIRST .. LAST (forward ascending, inverse descending); span of pass p = M / 16^p
template <int LOG2M, int NT, bool INV, int P, int LAST>
__device__ __forceinline__ void fft_mid_passes(float2 *s, const float2 *__restrict__ stw, int tid) {
    if constexpr (P <= LAST) {
        constexpr int LR = fft_radix_log2(LOG2M, P), LOG2S = LOG2M - 4 * P;
        if constexpr (!INV) {
            fft_pass<LR, false, LOG2M, LOG2S, NT>(s, stw, tid);
            fft_mid_passes<LOG2M, NT, false, P + 1, LAST>(s, stw, tid);
        } else {
            fft_mid_passes<LOG2M, NT, true, P + 1, LAST>(s, stw, tid);
            fft_pass<LR, true, LOG2M, LOG2S, NT>(s, stw, tid);
        }
    }
}

// generic (shared-memory) form for leader entry idx in [0, M/2]
template <int LOG2M>
__device__ __forceinline__ void pointwise_smem(float2 *s, int idx, const float4 c, const float4 c1) {
    constexpr int M = 1 << LOG2M;
    if (idx == 0) {                          // k = 0: DC and Nyquist, both real
        const float2 a = s[0];
        const float p0 = 2.f * (a.x + a.y) * c.x, pM = 2.f * (a.x - a.y) * c.y;
        s[0] = make_float2(p0 + pM, p0 - pM);
        return;
    }
    const int p = (idx == M / 2) ? 1 : 2 * idx;                    // slot 1 is k = M/2, self-paired
    const int p2 = p ^ ((1 << (31 - __clz(p))) - 1);
    float2 a = s[fft_pad(p)], b = s[fft_pad(p2)];
    pair_map(a, b, c, c1);
    s[fft_pad(p)] = a;
    if (p2 != p) s[fft_pad(p2)] = b;
}

template <int LOG2M>
struct FftCfg {
    static constexpr int M = 1 << LOG2M;
    static constexpr int NT = (M / 16 > 512) ? 512 : ((M / 16 < 64) ? 64 : M / 16);
    static constexpr int NTW = M / 16;                         // W_M^j, j < M/16: enough for every twiddled pass
    static constexpr int SDATA = M + M / 16 + 1;               // padded data array (float2)
    static constexpr int SMEM = (SDATA + NTW) * (int)sizeof(float2);
    static constexpr int NP = fft_num_passes(LOG2M);
    static constexpr int RL = fft_radix_log2(LOG2M, NP - 1);   // log2 radix of the last forward pass
    static constexpr bool FUSED = RL != 4;                     // last pass done in registers with the pointwise stage
};

// Overlap-save mode (sequences longer than the kernel, models/s4.py:1387 with L > l_max): the row has r samples,
// the two-sided kernel Lk taps per direction (n = 2M >= 2 Lk), P = n - Lk outputs per block; blockIdx.y = block j.
// The two directions are separate passes over separate windows because one window of n samples would leave only
// n - 2 Lk valid outputs:  anticausal pass (anti = 1): window [jP, jP + n), outputs at window positions [0, P) are
// written raw to `g` (the partial sums);  causal pass: window [jP - Lk, jP - Lk + n), outputs at positions
// [Lk, n) = times [jP, jP + P), g = gelu(value + partial[t]).  D and the 1/(4M) scale live in the causal table.
struct OlsArgs {
    const float *partial;
    int r, Lk, P, anti;
};

template <int LOG2M, bool OLS>
__device__ __forceinline__ void
fftconv_body(const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ part_t,
             long long part_stride_b, float ln_m, float ln_s, const float4 *__restrict__ kc,
             const float2 *__restrict__ tw /* W_n^i, i < M */, float *__restrict__ g, int B, int H, int l, const OlsArgs ols) {
    using Cfg = FftCfg<LOG2M>;
    constexpr int M = 1 << LOG2M, NT = Cfg::NT, NP = Cfg::NP, RL = Cfg::RL;
    constexpr int log2sub0 = LOG2M - 4, sub0 = 1 << log2sub0;      // pass 0: radix 16, span M
    extern __shared__ float2 s[];
    float2 *stw = s + Cfg::SDATA;
    const int tid = threadIdx.x;
    const int row = blockIdx.x;
    const int h = row / B, b = row - h * B;
    const size_t off = ((size_t)b * H + h) * (size_t)l;
    const float *xr = x + off;
    float *gr = g + off;
    const float pt = part_t ? part_t[(size_t)b * part_stride_b + h] : 0.f;
    const float *st = stats ? stats + (size_t)b * l * 2 : nullptr;
    const float4 *kcr = kc + (size_t)h * (M + 2);                  // (M/2 + 1) entries of two float4
    const float lns = st ? ln_s : 1.f, lnm = st ? ln_m : 0.f;
    const bool vec = ((l & 1) == 0);
    const int half = l >> 1;

    // packed input z[i] = y[2i] + i y[2i+1] as raw (x pair, statistics of the two time steps)
    auto load_in = [&](int i, float2 &xv, float4 &sv) {
        xv = make_float2(0.f, 0.f);
        sv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec) {
            if (i < half) {
                xv = reinterpret_cast<const float2 *>(xr)[i];
                sv = st ? reinterpret_cast<const float4 *>(st)[i] : make_float4(0.f, 1.f, 0.f, 1.f);
            }
        } else {
            const int t0 = 2 * i;
            if (t0 < l) {
                xv.x = xr[t0];
                sv.x = st ? st[2 * t0] : 0.f;
                sv.y = st ? st[2 * t0 + 1] : 1.f;
            }
            if (t0 + 1 < l) {
                xv.y = xr[t0 + 1];
                sv.z = st ? st[2 * t0 + 2] : 0.f;
                sv.w = st ? st[2 * t0 + 3] : 1.f;
            }
        }
    };
    // y = (ln_s rstd)(x - mean + ln_m) + part_t inside the row, 0 in the zero padding
    auto apply_in = [&](int i, float2 xv, float4 sv) {
        const int t0 = 2 * i;
        return make_float2((lns * sv.y) * (xv.x - sv.x + lnm) + (t0 < l ? pt : 0.f),
                           (lns * sv.w) * (xv.y - sv.z + lnm) + (t0 + 1 < l ? pt : 0.f));
    };

    // overlap-save: first sample of this block's window (may be negative: zero padding)
    const int w0 = OLS ? (ols.anti ? (int)blockIdx.y * ols.P : (int)blockIdx.y * ols.P - ols.Lk) : 0;
    auto ols_in = [&](int i) {                                   // y at window positions 2i, 2i+1 (0 outside the row)
        float2 y = make_float2(0.f, 0.f);
        const int t0 = w0 + 2 * i;
        if (t0 >= 0 && t0 < l) y.x = (lns * (st ? st[2 * t0 + 1] : 1.f)) * (xr[t0] - (st ? st[2 * t0] : 0.f) + lnm) + pt;
        if (t0 + 1 >= 0 && t0 + 1 < l)
            y.y = (lns * (st ? st[2 * t0 + 3] : 1.f)) * (xr[t0 + 1] - (st ? st[2 * t0 + 2] : 0.f) + lnm) + pt;
        return y;
    };

    if constexpr (OLS) {
        // ---- pass 0 forward on a FULL window: all 16 inputs of a butterfly are live
        for (int j = tid; j < Cfg::NTW; j += NT) stw[j] = tw[2 * j];
        __syncthreads();
#pragma unroll 1
        for (int bi = tid; bi < sub0; bi += NT) {
            float2 xx[16];
#pragma unroll
            for (int p = 0; p < 16; ++p) xx[p] = ols_in(bi + (p << log2sub0));
            Radix<16, false, false>::run(xx);
            if (log2sub0 > 0) apply_twiddles<16>(xx, stw[bi]);
#pragma unroll
            for (int q = 0; q < 16; ++q) s[fft_pad(bi + (fft_brev(q, 4) << log2sub0))] = xx[q];
        }
        __syncthreads();
    } else
    // ---- pass 0 forward, fused with the prologue: inputs i = j + p sub0; p >= 8 lies in the zero padding
    {
        float2 xv[8];
        float4 sv[8];
        int bi = tid;
        if (bi < sub0) {
#pragma unroll
            for (int p = 0; p < 8; ++p) load_in(bi + (p << log2sub0), xv[p], sv[p]);
        }
        for (int j = tid; j < Cfg::NTW; j += NT) stw[j] = tw[2 * j];
        __syncthreads();
#pragma unroll 1
        for (; bi < sub0; bi += NT) {
            float2 xx[16];
#pragma unroll
            for (int p = 0; p < 8; ++p) xx[p] = apply_in(bi + (p << log2sub0), xv[p], sv[p]);
            if constexpr (sub0 > NT) {            // next butterfly's loads go in flight under this one's arithmetic
                if (bi + NT < sub0) {
#pragma unroll
                    for (int p = 0; p < 8; ++p) load_in(bi + NT + (p << log2sub0), xv[p], sv[p]);
                }
            }
            Radix<16, false, true>::run(xx);
            if (log2sub0 > 0) apply_twiddles<16>(xx, stw[bi]);
#pragma unroll
            for (int q = 0; q < 16; ++q) s[fft_pad(bi + (fft_brev(q, 4) << log2sub0))] = xx[q];
        }
        __syncthreads();
    }

    // ---- middle forward passes; the last one stays separate only when the centre is not fused
    fft_mid_passes<LOG2M, NT, false, 1, Cfg::FUSED ? NP - 2 : NP - 1>(s, stw, tid);

    // ---- centre: (last forward pass +) untangle, product, re-tangle (+ first inverse pass)
    if constexpr (Cfg::FUSED) {
        constexpr int R = 1 << RL, NITEM = M / (2 * R);
        // item = radix groups (ga, gb = partner(ga)); its coefficients are entries (R g + c)/2, c even: R float4 per group.
        // They come from L2, so the next item's set is loaded while the current item is being computed.
        auto partner = [](int ga) { return ga == 0 ? 1 : ga ^ ((1 << (31 - __clz(ga))) - 1); };
        auto load_coef = [&](int item, float4 (&ca)[R], float4 (&cb)[R]) {
            const float4 *ka = kcr + (size_t)R * (2 * item), *kb = kcr + (size_t)R * partner(2 * item);
#pragma unroll
            for (int i = 0; i < R; ++i) {
                ca[i] = __ldg(ka + i);
                cb[i] = __ldg(kb + i);
            }
        };
        float4 ca[R], cb[R];
        if (tid < NITEM) load_coef(tid, ca, cb);
#pragma unroll 1
        for (int item = tid; item < NITEM; item += NT) {
            const int ga = 2 * item, gb = partner(ga);
            float2 xa[R], xb[R];
#pragma unroll
            for (int c = 0; c < R; ++c) {
                xa[c] = s[fft_pad(R * ga + c)];
                xb[c] = s[fft_pad(R * gb + c)];
            }
            float4 na[R], nb[R];
            if (item + NT < NITEM) load_coef(item + NT, na, nb);
            Radix<R, false>::run(xa);     // xa[q] = output digit q, whose slot is R ga + brev(q)
            Radix<R, false>::run(xb);
            if (item == 0) {
                // groups 0 and 1 pair inside themselves: slot 0 = DC/Nyquist (entry 0), slot 1 = k = M/2 paired
                // with itself (entry M/2), even slots c >= 2 of group 0 with c ^ ((1 << msb(c)) - 1), group 1 like
                // any other group but with itself.  Everything in registers; digit q sits in slot brev(q).
                const float4 s0 = __ldg(kcr + M), s1 = __ldg(kcr + M + 1);
                {
                    const float2 a = xa[0];
                    const float p0 = 2.f * (a.x + a.y) * ca[0].x, pM = 2.f * (a.x - a.y) * ca[0].y;
                    xa[0] = make_float2(p0 + pM, p0 - pM);
                    float2 d = xa[R / 2];
                    pair_map(xa[R / 2], d, s0, s1);
                }
#pragma unroll
                for (int c = 2; c < R; c += 2) {
                    int msb = 0;
                    while ((2 << msb) <= c) ++msb;
                    const int c2 = c ^ ((1 << msb) - 1);
                    pair_map(xa[fft_brev(c, RL)], xa[fft_brev(c2, RL)], ca[c], ca[c + 1]);
                }
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const int e = fft_brev(q, RL) >> 1;
                    pair_map(xb[q], xb[q ^ (R - 1)], cb[2 * e], cb[2 * e + 1]);
                }
            } else {
                // slot (ga, c) pairs with (gb, c ^ (R-1)); in digit order q <-> q ^ (R-1); leaders have q < R/2
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const int e = fft_brev(q, RL) >> 1;
                    pair_map(xa[q], xb[q ^ (R - 1)], ca[2 * e], ca[2 * e + 1]);
                    pair_map(xb[q], xa[q ^ (R - 1)], cb[2 * e], cb[2 * e + 1]);
                }
            }
            Radix<R, true>::run(xa);      // first inverse pass: input digit q is already in place
            Radix<R, true>::run(xb);
#pragma unroll
            for (int p = 0; p < R; ++p) {
                s[fft_pad(R * ga + p)] = xa[p];
                s[fft_pad(R * gb + p)] = xb[p];
            }
#pragma unroll
            for (int i = 0; i < R; ++i) {
                ca[i] = na[i];
                cb[i] = nb[i];
            }
        }
        __syncthreads();
    } else {
        // coefficients come from L2 (32 B per pair): GRP iterations' worth are requested at once, so a thread
        // exposes (M/2)/(NT GRP) round trips instead of (M/2)/NT
        constexpr int GRP = 4;
#pragma unroll 1
        for (int idx0 = tid; idx0 <= M / 2; idx0 += GRP * NT) {
            float4 c0[GRP], c1[GRP];
#pragma unroll
            for (int u = 0; u < GRP; ++u) {
                const int idx = idx0 + u * NT;
                c0[u] = c1[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (idx <= M / 2) {
                    c0[u] = __ldg(kcr + 2 * idx);
                    c1[u] = __ldg(kcr + 2 * idx + 1);
                }
            }
#pragma unroll
            for (int u = 0; u < GRP; ++u) {
                const int idx = idx0 + u * NT;
                if (idx <= M / 2) pointwise_smem<LOG2M>(s, idx, c0[u], c1[u]);
            }
        }
        __syncthreads();
    }

    // ---- middle inverse passes (mirror order)
    fft_mid_passes<LOG2M, NT, true, 1, Cfg::FUSED ? NP - 2 : NP - 1>(s, stw, tid);

    // ---- last inverse pass (pass 0), fused with the epilogue: only outputs p < 8 can fall inside the row
#pragma unroll 1
    for (int bi = tid; bi < sub0; bi += NT) {
        float2 xx[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) xx[q] = s[fft_pad(bi + (fft_brev(q, 4) << log2sub0))];
        if (log2sub0 > 0) {
            const float2 w1 = stw[bi];
            apply_twiddles<16>(xx, make_float2(w1.x, -w1.y));
        }
        Radix<16, true>::run(xx);
        if constexpr (OLS) {
            const int tb = (int)blockIdx.y * ols.P - (ols.anti ? 0 : ols.Lk);      // time of window position 0
            const int lo = ols.anti ? 0 : ols.Lk, hi = ols.anti ? ols.P : 2 * M;   // valid window positions [lo, hi)
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int pos = 2 * (bi + (p << log2sub0));
                const float v[2] = {xx[p].x, xx[p].y};
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int t = tb + pos + e;
                    if (pos + e >= lo && pos + e < hi && t < l)
                        gr[t] = ols.anti ? v[e] : gelu_fast(v[e] + (ols.partial ? ols.partial[off + t] : 0.f));
                }
            }
        } else {
#pragma unroll
            for (int p = 0; p < 8; ++p) {
                const int i = bi + (p << log2sub0);
                if (vec) {
                    if (i < half) reinterpret_cast<float2 *>(gr)[i] = make_float2(gelu_fast(xx[p].x), gelu_fast(xx[p].y));
                } else {
                    if (2 * i < l) gr[2 * i] = gelu_fast(xx[p].x);
                    if (2 * i + 1 < l) gr[2 * i + 1] = gelu_fast(xx[p].y);
                }
            }
        }
    }
}

template <int LOG2M>
__global__ void __launch_bounds__(FftCfg<LOG2M>::NT)
fftconv_kernel(const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ part_t,
               long long part_stride_b, float ln_m, float ln_s, const float4 *__restrict__ kc,
               const float2 *__restrict__ tw /* W_n^i, i < M */, float *__restrict__ g, int B, int H, int l) {
    pdl_trigger();          // launched through launch_pdl: let the next kernel's CTAs set up early, ...
    pdl_wait();             // ... and wait for the previous kernel before touching activations
    fftconv_body<LOG2M, false>(x, stats, part_t, part_stride_b, ln_m, ln_s, kc, tw, g, B, H, l, OlsArgs{});
}

template <int LOG2M>
__global__ void __launch_bounds__(FftCfg<LOG2M>::NT)
fftconv_ols_kernel(const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ part_t,
                   long long part_stride_b, float ln_m, float ln_s, const float4 *__restrict__ kc,
                   const float2 *__restrict__ tw, float *g, int B, int H, OlsArgs ols) {
    fftconv_body<LOG2M, true>(x, stats, part_t, part_stride_b, ln_m, ln_s, kc, tw, g, B, H, ols.r, ols);
}

// =====================================================================================================
// v2: split transform.  The packed row z[i] (i < M) is zero for i >= M/2, so its M-point spectrum is two
// independent Mh = M/2 point transforms:  Z[2k'] = FFT_Mh(z)[k']  and  Z[2k'+1] = FFT_Mh(z[i] W_M^i)[k'].
// The conjugate pairs (k, M-k) of the real-FFT untangle never mix the two (k and M-k have the same
// parity), and the wanted outputs are z'[i] = a[i] + W_M^{-i} b[i], i < Mh, with a, b the inverse
// transforms of the two halves.  One CTA runs the two halves one after the other in Mh complex of shared
// memory (half of v1's footprint, so two rows are resident per SM and one row's shared-memory phases
// overlap the other's arithmetic); `a` is parked in the output row (written and read back by the same
// thread) and the W_M^{+-i} rotations are folded into the outer radix-16 passes (a constant rotation
// W_32^p of the butterfly inputs/outputs and odd instead of even twiddle powers).
// =====================================================================================================

template <int LOG2M>
struct Fft2Cfg {
    static constexpr int LH = LOG2M - 1, Mh = 1 << LH;
    static constexpr int NT = (Mh / 32 > 256) ? 256 : ((Mh / 32 < 64) ? 64 : Mh / 32);
    static constexpr int NTW = Mh / 16;                        // W_Mh^j and W_M^j, j < Mh/16
    static constexpr int SDATA = Mh + Mh / 16 + 1;
    static constexpr int SMEM = (SDATA + 2 * NTW) * (int)sizeof(float2);
    static constexpr int NP = fft_num_passes(LH);
    static constexpr int RL = fft_radix_log2(LH, NP - 1);      // centre radix: 2, 4 or 8 (fft_use_v2)
    static_assert(LH >= 8 && RL >= 1 && RL <= 3, "v2 needs a radix-2/4/8 tail");
};

template <int LOG2M>
__global__ void __launch_bounds__(Fft2Cfg<LOG2M>::NT, 512 / Fft2Cfg<LOG2M>::NT)
fftconv2_kernel(const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ part_t,
                long long part_stride_b, float ln_m, float ln_s, const float4 *__restrict__ kc,
                const float2 *__restrict__ tw /* W_n^i, i < M */, float *g, int B, int H, int l) {
    using Cfg = Fft2Cfg<LOG2M>;
    constexpr int LH = Cfg::LH, Mh = Cfg::Mh, NT = Cfg::NT, NP = Cfg::NP, RL = Cfg::RL;
    constexpr int log2sub0 = LH - 4, sub0 = 1 << log2sub0;         // outer pass: radix 16, span Mh
    extern __shared__ float2 s[];
    float2 *stwA = s + Cfg::SDATA;                                 // W_Mh^j
    float2 *stwB = stwA + Cfg::NTW;                                // W_M^j
    const int tid = threadIdx.x;
    const int row = blockIdx.x;
    const int h = row / B, b = row - h * B;
    const size_t off = ((size_t)b * H + h) * (size_t)l;
    const float *xr = x + off;
    float *gr = g + off;
    const float pt = part_t ? part_t[(size_t)b * part_stride_b + h] : 0.f;
    const float *st = stats ? stats + (size_t)b * l * 2 : nullptr;
    const float4 *kcr = kc + (size_t)h * (2 * Mh + 2);             // (Mh + 1) entries of two float4
    const float lns = st ? ln_s : 1.f, lnm = st ? ln_m : 0.f;
    const bool vec = ((l & 1) == 0);
    const int half = l >> 1;

    auto load_in = [&](int i, float2 &xv, float4 &sv) {
        xv = make_float2(0.f, 0.f);
        sv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec) {
            if (i < half) {
                xv = reinterpret_cast<const float2 *>(xr)[i];
                sv = st ? reinterpret_cast<const float4 *>(st)[i] : make_float4(0.f, 1.f, 0.f, 1.f);
            }
        } else {
            const int t0 = 2 * i;
            if (t0 < l) {
                xv.x = xr[t0];
                sv.x = st ? st[2 * t0] : 0.f;
                sv.y = st ? st[2 * t0 + 1] : 1.f;
            }
            if (t0 + 1 < l) {
                xv.y = xr[t0 + 1];
                sv.z = st ? st[2 * t0 + 2] : 0.f;
                sv.w = st ? st[2 * t0 + 3] : 1.f;
            }
        }
    };
    auto apply_in = [&](int i, float2 xv, float4 sv) {
        const int t0 = 2 * i;
        return make_float2((lns * sv.y) * (xv.x - sv.x + lnm) + (t0 < l ? pt : 0.f),
                           (lns * sv.w) * (xv.y - sv.z + lnm) + (t0 + 1 < l ? pt : 0.f));
    };

    // L1 prefetches (no destination registers): inputs of the next outer butterfly / the parked a[i]
    auto prefetch_in = [&](int bi) {
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int i = bi + (p << log2sub0);
            if (2 * i < l) {
                prefetch_l1(xr + 2 * i);
                if (st) prefetch_l1(st + 4 * i);
            }
        }
    };
    auto prefetch_out = [&](int bi) {
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int i = bi + (p << log2sub0);
            if (2 * i < l) prefetch_l1(gr + 2 * i);
        }
    };

    for (int j = tid; j < Cfg::NTW; j += NT) {
        stwA[j] = tw[4 * j];
        stwB[j] = tw[2 * j];
    }
    __syncthreads();

#pragma unroll 1
    for (int odd = 0; odd < 2; ++odd) {       // one copy of the code for both halves: instruction-cache footprint
        // ---- outer forward pass, fused with the prologue: inputs i = j + p sub0, p < 16
#pragma unroll 1
        for (int bi = tid; bi < sub0; bi += NT) {
            float2 xx[16];
            if (bi + NT < sub0) prefetch_in(bi + NT);
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {          // two batches of 8 loads bound the registers in flight
                float2 xv[8];
                float4 sv[8];
#pragma unroll
                for (int p = 0; p < 8; ++p) load_in(bi + ((8 * hb + p) << log2sub0), xv[p], sv[p]);
#pragma unroll
                for (int p = 0; p < 8; ++p) xx[8 * hb + p] = apply_in(bi + ((8 * hb + p) << log2sub0), xv[p], sv[p]);
            }
            if (odd) rotate_w32<false>(xx);
            Radix<16, false>::run(xx);
            apply_twiddles16<true>(xx, odd ? stwB[bi] : make_float2(1.f, 0.f), stwA[bi]);
#pragma unroll
            for (int q = 0; q < 16; ++q) s[fft_pad(bi + (fft_brev(q, 4) << log2sub0))] = xx[q];
        }
        __syncthreads();

        fft_mid_passes<LH, NT, false, 1, NP - 2>(s, stwA, tid);

        // ---- centre: last forward pass + untangle/product/re-tangle + first inverse pass, in registers
        {
            constexpr int R = 1 << RL, G = Mh / R, NITEM = Mh / (2 * R);
            const float4 *kh = odd ? kcr + (Mh + 2) : kcr;          // odd half: entries Mh/2 + 1 ...
            auto partner = [&](int ga) {
                return odd ? (ga ^ (G - 1)) : (ga == 0 ? 1 : ga ^ ((1 << (31 - __clz(ga))) - 1));
            };
            auto load_coef = [&](int item, float4 (&ca)[R], float4 (&cb)[R]) {
                const float4 *ka = kh + (size_t)R * (2 * item), *kb = kh + (size_t)R * partner(2 * item);
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    ca[i] = __ldg(ka + i);
                    cb[i] = __ldg(kb + i);
                }
            };
            float4 ca[R], cb[R];
            if (tid < NITEM) load_coef(tid, ca, cb);
#pragma unroll 1
            for (int item = tid; item < NITEM; item += NT) {
                const int ga = 2 * item, gb = partner(ga);
                float2 xa[R], xb[R];
#pragma unroll
                for (int c = 0; c < R; ++c) {
                    xa[c] = s[fft_pad(R * ga + c)];
                    xb[c] = s[fft_pad(R * gb + c)];
                }
                float4 na[R], nb[R];
                if (item + NT < NITEM) load_coef(item + NT, na, nb);
                Radix<R, false>::run(xa);
                Radix<R, false>::run(xb);
                if (!odd && item == 0) {
                    const float4 s0 = __ldg(kcr + Mh), s1 = __ldg(kcr + Mh + 1);
                    {
                        const float2 a = xa[0];
                        const float p0 = 2.f * (a.x + a.y) * ca[0].x, pM = 2.f * (a.x - a.y) * ca[0].y;
                        xa[0] = make_float2(p0 + pM, p0 - pM);
                        float2 d = xa[R / 2];
                        pair_map(xa[R / 2], d, s0, s1);
                    }
#pragma unroll
                    for (int c = 2; c < R; c += 2) {
                        int msb = 0;
                        while ((2 << msb) <= c) ++msb;
                        const int c2 = c ^ ((1 << msb) - 1);
                        pair_map(xa[fft_brev(c, RL)], xa[fft_brev(c2, RL)], ca[c], ca[c + 1]);
                    }
#pragma unroll
                    for (int q = 0; q < R / 2; ++q) {
                        const int e = fft_brev(q, RL) >> 1;
                        pair_map(xb[q], xb[q ^ (R - 1)], cb[2 * e], cb[2 * e + 1]);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < R / 2; ++q) {
                        const int e = fft_brev(q, RL) >> 1;
                        pair_map(xa[q], xb[q ^ (R - 1)], ca[2 * e], ca[2 * e + 1]);
                        pair_map(xb[q], xa[q ^ (R - 1)], cb[2 * e], cb[2 * e + 1]);
                    }
                }
                Radix<R, true>::run(xa);
                Radix<R, true>::run(xb);
#pragma unroll
                for (int p = 0; p < R; ++p) {
                    s[fft_pad(R * ga + p)] = xa[p];
                    s[fft_pad(R * gb + p)] = xb[p];
                }
#pragma unroll
                for (int i = 0; i < R; ++i) {
                    ca[i] = na[i];
                    cb[i] = nb[i];
                }
            }
            __syncthreads();
        }

        if (odd && tid < sub0) prefetch_out(tid);
        fft_mid_passes<LH, NT, true, 1, NP - 2>(s, stwA, tid);

        // ---- outer inverse pass: even half parks a[i] in the output row; odd half reads it back (same
        //      thread, same addresses), adds W_M^{-i} b[i] and writes GELU
#pragma unroll 1
        for (int bi = tid; bi < sub0; bi += NT) {
            if (bi + NT < sub0) {
                if (odd) prefetch_out(bi + NT);
            } else if (!odd) {
                prefetch_in(tid);               // first outer butterfly of the odd half
            }
            float2 xx[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) xx[q] = s[fft_pad(bi + (fft_brev(q, 4) << log2sub0))];
            apply_twiddles16<true>(xx, odd ? cconj(stwB[bi]) : make_float2(1.f, 0.f), cconj(stwA[bi]));
            Radix<16, true>::run(xx);
            if (odd) {
                rotate_w32<true>(xx);
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    const int i = bi + (p << log2sub0);
                    float2 av = make_float2(0.f, 0.f);          // a[i]: parked by this thread, prefetched to L1
                    if (vec) {
                        if (i < half) av = reinterpret_cast<const float2 *>(gr)[i];
                    } else {
                        if (2 * i < l) av.x = gr[2 * i];
                        if (2 * i + 1 < l) av.y = gr[2 * i + 1];
                    }
                    xx[p] = make_float2(gelu_fast(av.x + xx[p].x), gelu_fast(av.y + xx[p].y));
                }
            }
#pragma unroll
            for (int p = 0; p < 16; ++p) {
                const int i = bi + (p << log2sub0);
                if (vec) {
                    if (i < half) reinterpret_cast<float2 *>(gr)[i] = xx[p];
                } else {
                    if (2 * i < l) gr[2 * i] = xx[p].x;
                    if (2 * i + 1 < l) gr[2 * i + 1] = xx[p].y;
                }
            }
        }
        __syncthreads();          // the next half overwrites the shared array
    }
}

// ---- twiddle table W_n^i, i < M: immutable, per (device, log2M), created on first use -------
__global__ void twiddle_kernel(float2 *tw, int log2M) {
    const int M = 1 << log2M, n = 2 * M;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    double sn, cs;
    sincospi(-2.0 * (double)i / (double)n, &sn, &cs);
    tw[i] = make_float2((float)cs, (float)sn);
}

struct TwiddleCache {
    std::mutex mu;
    float2 *tw[16][FFT_MAX_LOG2M + 1] = {};
    float2 *tw2[16][FFT_MAX_LOG2M + 1] = {};      // per-item pair twiddles of the compact pointwise table
};
static TwiddleCache g_tw;

int fft_twiddles(int log2M, cudaStream_t st, const float2 **tw) {
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
        float2 *a = nullptr;
        DWB_CUDA(cudaMalloc(&a, (size_t)M * sizeof(float2)));
        twiddle_kernel<<<ceil_div(M, 256), 256, 0, st>>>(a, log2M);
        DWB_LAUNCH_CHECK();
        DWB_CUDA(cudaStreamSynchronize(st));
        g_tw.tw[dev][log2M] = a;
    }
    *tw = g_tw.tw[dev][log2M];
    return DWB_OK;
}

int fft_pair_twiddles_launch(float2 *tw2, int log2M, cudaStream_t st);   // s4_kernelgen.cu

int fft_pair_twiddles(int log2M, cudaStream_t st, const float2 **tw2) {
    int dev = 0;
    DWB_CUDA(cudaGetDevice(&dev));
    DWB_REQUIRE(dev >= 0 && dev < 16, DWB_ERR_UNSUPPORTED, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lock(g_tw.mu);
    if (!g_tw.tw2[dev][log2M]) {
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        DWB_CUDA(cudaStreamIsCapturing(st, &cs));
        DWB_REQUIRE(cs == cudaStreamCaptureStatusNone, DWB_ERR_STATE,
                    "fft pair-twiddle table for M=2^%d must be created before stream capture", log2M);
        float2 *a = nullptr;
        DWB_CUDA(cudaMalloc(&a, (size_t)((1 << log2M) / 8) * sizeof(float2)));
        int rc = fft_pair_twiddles_launch(a, log2M, st);
        if (rc != DWB_OK) return rc;
        DWB_CUDA(cudaStreamSynchronize(st));
        g_tw.tw2[dev][log2M] = a;
    }
    *tw2 = g_tw.tw2[dev][log2M];
    return DWB_OK;
}

template <int LOG2M>
static int launch_fftconv(const float *x, const float *stats, const float *part_t, long long psb, float ln_m,
                          float ln_s, const float *kc, const float2 *tw, float *g, int B, int H, int l, cudaStream_t st) {
    using Cfg = FftCfg<LOG2M>;
    static bool attr_set[16] = {};
    int dev = 0;
    DWB_CUDA(cudaGetDevice(&dev));
    if (Cfg::SMEM > 48 * 1024 && !attr_set[dev & 15]) {
        DWB_CUDA(cudaFuncSetAttribute(fftconv_kernel<LOG2M>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_set[dev & 15] = true;
    }
    DWB_CUDA(launch_pdl(fftconv_kernel<LOG2M>, dim3(B * H), dim3(Cfg::NT), Cfg::SMEM, st, x, stats, part_t, psb, ln_m, ln_s, (const float4 *)kc, tw, g,
                        B, H, l));
    return DWB_OK;
}

template <int LOG2M>
static int launch_fftconv2(const float *x, const float *stats, const float *part_t, long long psb, float ln_m,
                           float ln_s, const float *kc, const float2 *tw, float *g, int B, int H, int l, cudaStream_t st) {
    using Cfg = Fft2Cfg<LOG2M>;
    static bool attr_set[16] = {};
    int dev = 0;
    DWB_CUDA(cudaGetDevice(&dev));
    if (Cfg::SMEM > 48 * 1024 && !attr_set[dev & 15]) {
        DWB_CUDA(cudaFuncSetAttribute(fftconv2_kernel<LOG2M>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_set[dev & 15] = true;
    }
    fftconv2_kernel<LOG2M><<<B * H, Cfg::NT, Cfg::SMEM, st>>>(x, stats, part_t, psb, ln_m, ln_s, (const float4 *)kc, tw, g, B, H, l);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

template <int LOG2M>
static int launch_fftconv_ols(const float *x, const float *stats, const float *part_t, long long psb, float ln_m,
                              float ln_s, const float *kc, const float2 *tw, float *g, int B, int H, const OlsArgs &o,
                              cudaStream_t st) {
    using Cfg = FftCfg<LOG2M>;
    static bool attr_set[16] = {};
    int dev = 0;
    DWB_CUDA(cudaGetDevice(&dev));
    if (Cfg::SMEM > 48 * 1024 && !attr_set[dev & 15]) {
        DWB_CUDA(cudaFuncSetAttribute(fftconv_ols_kernel<LOG2M>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
        attr_set[dev & 15] = true;
    }
    const dim3 grid(B * H, ceil_div(o.r, o.P));
    fftconv_ols_kernel<LOG2M><<<grid, Cfg::NT, Cfg::SMEM, st>>>(x, stats, part_t, psb, ln_m, ln_s, (const float4 *)kc, tw, g, B, H, o);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// rows of r > Lk samples against the Lk-tap two-sided kernel: anticausal pass into `partial`, causal pass adds it,
// applies D and GELU.  kc_c / kc_a: v1-layout (mode 0) tables of (k0, 0, D) and (0, k1, no D) at n = 2 M(Lk).
int fftconv_ols_launch(const float *x, const float *stats, const float *part_t, long long psb, float ln_m, float ln_s,
                       const float *kc_c, const float *kc_a, float *g, float *partial, int B, int H, int r, int Lk,
                       cudaStream_t st) {
    const int lg = fft_log2m_for(Lk);
    DWB_REQUIRE(lg > 0, DWB_ERR_UNSUPPORTED, "fftconv: kernel length %d unsupported (max %d)", Lk, 1 << FFT_MAX_LOG2M);
    DWB_REQUIRE(partial && partial != g && partial != x, DWB_ERR_INVALID, "fftconv_ols: needs a distinct partial-sum buffer");
    DWB_REQUIRE((int64_t)B * H <= 0x7fffffff && ceil_div(r, (2 << lg) - Lk) <= 65535, DWB_ERR_UNSUPPORTED, "fftconv_ols: grid too large");
    const float2 *tw;
    int rc = fft_twiddles(lg, st, &tw);
    if (rc != DWB_OK) return rc;
    OlsArgs o{};
    o.r = r; o.Lk = Lk; o.P = (2 << lg) - Lk;
#define DWB_OLS_CASE(LG)                                                                                             \
    case LG:                                                                                                         \
        o.anti = 1; o.partial = nullptr;                                                                             \
        rc = launch_fftconv_ols<LG>(x, stats, part_t, psb, ln_m, ln_s, kc_a, tw, partial, B, H, o, st);              \
        if (rc != DWB_OK) return rc;                                                                                 \
        o.anti = 0; o.partial = partial;                                                                             \
        return launch_fftconv_ols<LG>(x, stats, part_t, psb, ln_m, ln_s, kc_c, tw, g, B, H, o, st);
    switch (lg) {
        DWB_OLS_CASE(4) DWB_OLS_CASE(5) DWB_OLS_CASE(6) DWB_OLS_CASE(7) DWB_OLS_CASE(8) DWB_OLS_CASE(9)
        DWB_OLS_CASE(10) DWB_OLS_CASE(11) DWB_OLS_CASE(12) DWB_OLS_CASE(13) DWB_OLS_CASE(14)
    }
#undef DWB_OLS_CASE
    set_error("fftconv_ols: no kernel for log2M=%d", lg);
    return DWB_ERR_UNSUPPORTED;
}

int fftconv_launch(const float *x, const float *stats, const float *part_t, long long psb, float ln_m, float ln_s,
                   const float *kc, float *g, int B, int H, int l, cudaStream_t st, float *scratch) {
    const int lg = fft_log2m_for(l);
    DWB_REQUIRE(lg > 0, DWB_ERR_UNSUPPORTED, "fftconv: stage length %d unsupported (max %d)", l, 1 << FFT_MAX_LOG2M);
    const float2 *tw;
    int rc = fft_twiddles(lg, st, &tw);
    if (rc != DWB_OK) return rc;
    if (fft_use_v2(lg)) {
        const int mode = fft_table_mode(lg, l);
        if (mode == 2) {
            DWB_REQUIRE(fftconv3_supported(lg, x, stats, g, l), DWB_ERR_UNSUPPORTED,
                        "fftconv: rows of the n = %d stage must be 16-byte aligned", 2 << lg);
            const float2 *tw2;
            rc = fft_pair_twiddles(lg, st, &tw2);
            if (rc != DWB_OK) return rc;
            return fftconv3_launch(lg, x, stats, part_t, psb, ln_m, ln_s, kc, tw, tw2, g, scratch, B, H, l, st);
        }
        if (fft_forced_variant() != 2 && fftconv3_supported(lg, x, stats, g, l))
            return fftconv3_launch(lg, x, stats, part_t, psb, ln_m, ln_s, kc, tw, nullptr, g, scratch, B, H, l, st);
        switch (lg) {
            case 12: return launch_fftconv2<12>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, g, B, H, l, st);
            case 14: return launch_fftconv2<14>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, g, B, H, l, st);
        }
    }
    {
        static const bool v5 = getenv("DWB_FFT5") != nullptr;       // experiment: unsplit packed kernel at n <= 8192
        if (v5 && !fft_use_v2(lg) && fft_forced_variant() != 1 && fftconv5_supported(lg, x, stats, g, l))
            return fftconv5_launch(lg, x, stats, part_t, psb, ln_m, ln_s, kc, tw, g, B, H, l, st);
    }
#define DWB_FFT_CASE(LG) \
    case LG:             \
        return launch_fftconv<LG>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, g, B, H, l, st);
    switch (lg) {
        DWB_FFT_CASE(4) DWB_FFT_CASE(5) DWB_FFT_CASE(6) DWB_FFT_CASE(7) DWB_FFT_CASE(8) DWB_FFT_CASE(9)
        DWB_FFT_CASE(10) DWB_FFT_CASE(11) DWB_FFT_CASE(12) DWB_FFT_CASE(13) DWB_FFT_CASE(14)
    }
#undef DWB_FFT_CASE
    set_error("fftconv: no kernel for log2M=%d", lg);
    return DWB_ERR_UNSUPPORTED;
}

}  // namespace dwb

// scratch row store for the standalone op (the plan passes the block's still-unused output buffer instead):
// grown on demand per device, never during stream capture (then the kernel simply re-reads its input)
static float *op_scratch(size_t floats, cudaStream_t st) {
    static std::mutex mu;
    static float *buf[16] = {};
    static size_t cap[16] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return nullptr;
    std::lock_guard<std::mutex> lock(mu);
    if (cap[dev] < floats) {
        if (buf[dev]) {
            cudaDeviceSynchronize();
            cudaFree(buf[dev]);
        }
        buf[dev] = nullptr;
        cap[dev] = 0;
        if (cudaMalloc(&buf[dev], floats * sizeof(float)) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        cap[dev] = floats;
    }
    return buf[dev];
}

extern "C" int dwb_fftconv(const float *x, const float *stats, const float *part_t, int64_t part_stride_b, float ln_m,
                           float ln_s, const float *kf, float *g, int B, int H, int l, void *stream) {
    DWB_REQUIRE(x && kf && g, DWB_ERR_INVALID, "dwb_fftconv: null pointer");
    DWB_REQUIRE(B >= 1 && H >= 1 && l >= 1, DWB_ERR_INVALID, "dwb_fftconv: bad sizes B=%d H=%d l=%d", B, H, l);
    float *scratch = nullptr;
    if (dwb::fft_use_v2(dwb::fft_log2m_for(l))) scratch = op_scratch((size_t)B * H * l, (cudaStream_t)stream);
    return dwb::fftconv_launch(x, stats, part_t, (long long)part_stride_b, ln_m, ln_s, kf, g, B, H, l,
                               (cudaStream_t)stream, scratch);
}
