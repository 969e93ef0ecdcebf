// fftconv v3: the split transform of fftconv2_kernel (fftconv.cu) on packed fp32 arithmetic.
//
// Same algorithm as v2 — the half-empty M-point packed row is two independent Mh = M/2 point
// transforms (even / odd output frequencies) run one after the other in Mh complex of shared memory,
// `a` parked in the output row — but every radix-16 pass processes TWO butterflies per thread in the two
// lanes of FADD2/FMUL2/FFMA2 (fft_simd2.cuh).  For that the shared array is kept as separate re / im
// planes: the butterflies j and j+1 of a pass touch adjacent plane entries, so one 64-bit LDS/STS moves
// the same element of both butterflies and lands directly in a register pair.  The radix-2/4/8 centre
// (last forward pass + untangle/product/re-tangle + first inverse pass) stays scalar: its conjugate-pair
// partner map reverses index order, which would need lane swaps.
//
// Reference: models/s4.py:1403-1411 (rfft/irfft product at n = 2l, D skip), :1430 (GELU),
// models/sashimi.py:148-152 (norm1 + fc_t).  Requires l % 4 == 0 and 16-byte aligned rows.
#include <stdint.h>

#include <stdlib.h>

#include <algorithm>

#include "common.cuh"
#include "fft_plan.cuh"
#include "fft_radix.cuh"
#include "fft_simd2.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace dwb {
using s2::C2;
using s2::V2;

// plane padding: two floats per 32 keep 64-bit alignment and make the stride-32 accesses of the last
// radix-16 pass (one 32-float block per thread) conflict free
__host__ __device__ constexpr int padf(int i) { return i + 2 * (i >> 5); }

template <int LOG2M>
struct Fft3Cfg {
    static constexpr int LH = LOG2M - 1, Mh = 1 << LH;
    static constexpr int NT = Mh / 32;                       // one butterfly PAIR per thread in every radix-16 pass
    static constexpr int NP = fft_num_passes(LH);            // radix-16 passes + the radix-2/4/8 tail
    static constexpr int RL = fft_radix_log2(LH, NP - 1);
    static constexpr int PLANE = Mh + Mh / 16 + 2;           // floats per plane
    static constexpr int NTW0 = Mh / 16;                     // W_Mh^j and W_M^j, j < Mh/16
    static constexpr int mid_off(int P) {                    // twiddles of pass P (1..NP-2): W_{S_P}^j, j < S_P/16
        int o = 0;
        for (int q = 1; q < P; ++q) o += Mh >> (4 * q + 4);
        return o;
    }
    static constexpr int NMID = mid_off(NP - 1);
    static constexpr int TWBYTES = (4 * NTW0 + 2 * NMID) * (int)sizeof(float);     // the six twiddle tables, contiguous after the planes
    static constexpr int SMEM = 2 * PLANE * (int)sizeof(float) + TWBYTES;
    static_assert(TWBYTES % 16 == 0, "bulk-copied twiddle image");
    static constexpr int MINB = (512 / NT > 8) ? 8 : 512 / NT;
    static_assert(RL >= 1 && RL <= 3 && NP >= 3 && NT >= 32, "v3 plan");
};

// the twiddle region in its shared-memory layout: W_Mh^j (re | im), W_M^j (re | im), j < Mh/16, then per middle pass
// W_{S_P}^j (all re | all im); dst = shared memory (per CTA) or the global image built once per (device, log2M)
template <int LOG2M>
__device__ __forceinline__ void fft3_fill_twiddles(const float2 *__restrict__ tw, float *dst, int tid0, int nthreads = Fft3Cfg<LOG2M>::NT) {
    using Cfg = Fft3Cfg<LOG2M>;
    float *twAr = dst, *twAi = twAr + Cfg::NTW0, *twBr = twAi + Cfg::NTW0, *twBi = twBr + Cfg::NTW0;
    float *midr = twBi + Cfg::NTW0, *midi = midr + Cfg::NMID;
    for (int j = tid0; j < Cfg::NTW0; j += nthreads) {
        const float2 a = tw[4 * j], bb = tw[2 * j];
        twAr[j] = a.x;
        twAi[j] = a.y;
        twBr[j] = bb.x;
        twBi[j] = bb.y;
    }
    int o = 0;
#pragma unroll
    for (int P = 1; P <= Cfg::NP - 2; ++P) {                        // W_{S_P}^j = W_n^{j 2^(2 + 4P)}
        const int subP = Cfg::Mh >> (4 * P + 4);
        for (int j = tid0; j < subP; j += nthreads) {
            const float2 a = tw[j << (2 + 4 * P)];
            midr[o + j] = a.x;
            midi[o + j] = a.y;
        }
        o += subP;
    }
}
template <int LOG2M>
__global__ void fft3_twimg_kernel(const float2 *__restrict__ tw, float *img) {
    fft3_fill_twiddles<LOG2M>(tw, img, threadIdx.x, blockDim.x);
}

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ V2 ld2(const float *p) { return V2(*reinterpret_cast<const float2 *>(p)); }
__device__ __forceinline__ void st2(float *p, const V2 &a) { *reinterpret_cast<float2 *>(p) = a.v; }

// one in-place radix-16 pass P (span S = Mh / 16^P) over the planes, butterflies 2 tid and 2 tid + 1
template <int LH, int P, bool INV>
__device__ __forceinline__ void mid_pass(float *re, float *im, const float *twr, const float *twi, int tid) {
    constexpr int LOG2S = LH - 4 * P, log2sub = LOG2S - 4, sub = 1 << log2sub;
    static_assert(log2sub >= 1, "pairs of adjacent butterflies need sub >= 2");
    const int bi = 2 * tid;
    const int j = bi & (sub - 1);
    const int base = ((bi >> log2sub) << LOG2S) + j;
    const int pb = padf(base);
    C2 w1;
    w1.x = ld2(twr + j);
    w1.y = ld2(twi + j);
    if (INV) w1.y = -w1.y;
    C2 one;
    one.x = V2(1.f);
    one.y = V2(0.f);
    C2 x[16];
    if (!INV) {
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int o = p * sub + 2 * ((p * sub) >> 5);
            x[p].x = ld2(re + pb + o);
            x[p].y = ld2(im + pb + o);
        }
        s2::RadixS<16, false>::run(x);
        s2::apply_twiddles16<false>(x, one, w1);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int c = fft_brev(q, 4), o = c * sub + 2 * ((c * sub) >> 5);
            st2(re + pb + o, x[q].x);
            st2(im + pb + o, x[q].y);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int c = fft_brev(q, 4), o = c * sub + 2 * ((c * sub) >> 5);
            x[q].x = ld2(re + pb + o);
            x[q].y = ld2(im + pb + o);
        }
        s2::apply_twiddles16<false>(x, one, w1);
        s2::RadixS<16, true>::run(x);
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int o = p * sub + 2 * ((p * sub) >> 5);
            st2(re + pb + o, x[p].x);
            st2(im + pb + o, x[p].y);
        }
    }
    __syncthreads();
}

template <int LH, bool INV, int P, int LAST, int NMIDOFF>
__device__ __forceinline__ void mid_passes(float *re, float *im, const float *mr, const float *mi, int tid) {
    if constexpr (P <= LAST) {
        constexpr int sub = 1 << (LH - 4 * P - 4);
        if constexpr (!INV) {
            mid_pass<LH, P, false>(re, im, mr + NMIDOFF, mi + NMIDOFF, tid);
            mid_passes<LH, false, P + 1, LAST, NMIDOFF + sub>(re, im, mr, mi, tid);
        } else {
            mid_passes<LH, true, P + 1, LAST, NMIDOFF + sub>(re, im, mr, mi, tid);
            mid_pass<LH, P, true>(re, im, mr + NMIDOFF, mi + NMIDOFF, tid);
        }
    }
}

// Scalar centre on the planes with the 32 B/pair table: last forward pass (radix 2^RL, sub = 1) + untangle/product/
// re-tangle + first inverse pass of an N-point (half) transform.  kh: table base of this transform (two float4 per
// leader slot pair, R float4 per group); kspecial: its self-paired entry N/2; odd: odd-half pairing (complement
// groups, no DC / self-paired slots).
template <int N, int RL, int NT>
__device__ __forceinline__ void centre_wide(float *re, float *im, const float4 *kh, const float4 *kspecial, const bool odd,
                                            const int tid) {
    constexpr int R = 1 << RL, G = N / R, NITEM = N / (2 * R);
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
    // coefficients come from L2 (16 B per point, shared by the rows of a channel): GRP items' worth (64
    // registers) are requested at once, so a thread exposes IPT / GRP round trips instead of IPT
    constexpr int IPT = NITEM / NT, GRP = 8 / R;
    static_assert(IPT % GRP == 0, "centre grouping");
#pragma unroll 1
    for (int k0 = 0; k0 < IPT; k0 += GRP) {
        float4 ca[GRP][R], cb[GRP][R];
#pragma unroll
        for (int u = 0; u < GRP; ++u) load_coef(tid + (k0 + u) * NT, ca[u], cb[u]);
#pragma unroll
        for (int u = 0; u < GRP; ++u) {
            const int item = tid + (k0 + u) * NT;
            const int ga = 2 * item, gb = partner(ga);
            const int pa = padf(R * ga), pg = padf(R * gb);
            float2 xa[R], xb[R];
#pragma unroll
            for (int k = 0; k < R / 2; ++k) {
                const float2 ar = *reinterpret_cast<const float2 *>(re + pa + 2 * k);
                const float2 ai = *reinterpret_cast<const float2 *>(im + pa + 2 * k);
                const float2 br = *reinterpret_cast<const float2 *>(re + pg + 2 * k);
                const float2 bi = *reinterpret_cast<const float2 *>(im + pg + 2 * k);
                xa[2 * k] = make_float2(ar.x, ai.x);
                xa[2 * k + 1] = make_float2(ar.y, ai.y);
                xb[2 * k] = make_float2(br.x, bi.x);
                xb[2 * k + 1] = make_float2(br.y, bi.y);
            }
            Radix<R, false>::run(xa);
            Radix<R, false>::run(xb);
            if (!odd && item == 0) {
                const float4 s0 = __ldg(kspecial), s1 = __ldg(kspecial + 1);
                {
                    const float2 a = xa[0];
                    const float p0 = 2.f * (a.x + a.y) * ca[u][0].x, pM = 2.f * (a.x - a.y) * ca[u][0].y;
                    xa[0] = make_float2(p0 + pM, p0 - pM);
                    float2 d = xa[R / 2];
                    pair_map(xa[R / 2], d, s0, s1);
                }
#pragma unroll
                for (int c = 2; c < R; c += 2) {
                    int msb = 0;
                    while ((2 << msb) <= c) ++msb;
                    const int c2 = c ^ ((1 << msb) - 1);
                    pair_map(xa[fft_brev(c, RL)], xa[fft_brev(c2, RL)], ca[u][c], ca[u][c + 1]);
                }
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const int e = fft_brev(q, RL) >> 1;
                    pair_map(xb[q], xb[q ^ (R - 1)], cb[u][2 * e], cb[u][2 * e + 1]);
                }
            } else {
#pragma unroll
                for (int q = 0; q < R / 2; ++q) {
                    const int e = fft_brev(q, RL) >> 1;
                    pair_map(xa[q], xb[q ^ (R - 1)], ca[u][2 * e], ca[u][2 * e + 1]);
                    pair_map(xb[q], xa[q ^ (R - 1)], cb[u][2 * e], cb[u][2 * e + 1]);
                }
            }
            Radix<R, true>::run(xa);
            Radix<R, true>::run(xb);
#pragma unroll
            for (int k = 0; k < R / 2; ++k) {
                *reinterpret_cast<float2 *>(re + pa + 2 * k) = make_float2(xa[2 * k].x, xa[2 * k + 1].x);
                *reinterpret_cast<float2 *>(im + pa + 2 * k) = make_float2(xa[2 * k].y, xa[2 * k + 1].y);
                *reinterpret_cast<float2 *>(re + pg + 2 * k) = make_float2(xb[2 * k].x, xb[2 * k + 1].x);
                *reinterpret_cast<float2 *>(im + pg + 2 * k) = make_float2(xb[2 * k].y, xb[2 * k + 1].y);
            }
        }
    }
    __syncthreads();
}

// TPARK (n = 32768 only): the two rows a CTA parks between its halves - the LayerNorm-applied input for the odd half and
// the even half's result a[i] - live in TENSOR MEMORY instead of global memory.  The kernel issues no MMA, so the SM's
// 256 KB of TMEM are idle; each of the two resident CTAs allocates 256 columns, a thread owns 128 words of its lane
// (64 for y, 64 for a), writes them with tcgen05.st and reads them back itself with tcgen05.ld.  That removes
// 4 x 64 KB per row of global stores and loads (and their L2 / DRAM traffic: 1.69x -> ~1.0x the algorithmic bytes).
template <int LOG2M, bool COMPACT, bool TPARK>
__global__ void __launch_bounds__(Fft3Cfg<LOG2M>::NT, Fft3Cfg<LOG2M>::MINB)
fftconv3_kernel(const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ part_t,
                long long part_stride_b, float ln_m, float ln_s, const float4 *__restrict__ kc,
                const float2 *__restrict__ tw /* W_n^i, i < M */, const float2 *__restrict__ tw2, float *g, float *scratch, int B, int H,
                int l, int resident, int stagger_ns, const float *__restrict__ twimg) {
    using Cfg = Fft3Cfg<LOG2M>;
    constexpr int LH = Cfg::LH, Mh = Cfg::Mh, NT = Cfg::NT, NP = Cfg::NP, RL = Cfg::RL;
    constexpr int log2sub0 = LH - 4, sub0 = 1 << log2sub0;         // outer pass: radix 16, span Mh
    extern __shared__ float sm[];
    float *re = sm, *im = sm + Cfg::PLANE;
    float *twAr = im + Cfg::PLANE, *twAi = twAr + Cfg::NTW0;       // W_Mh^j
    float *twBr = twAi + Cfg::NTW0, *twBi = twBr + Cfg::NTW0;      // W_M^j
    float *midr = twBi + Cfg::NTW0, *midi = midr + Cfg::NMID;
    pdl_trigger();
    const int tid = threadIdx.x;
    uint32_t tpark = 0;                                            // this thread's 128 TMEM words
    if constexpr (TPARK) {
        static_assert(!TPARK || (NT == 256 && Cfg::MINB == 2), "TMEM parking: 8 warps x 128 columns, two CTAs per SM");
        uint32_t *tptr = reinterpret_cast<uint32_t *>(sm + Cfg::SMEM / sizeof(float));
        if (tid < 32) umma::tmem_alloc(tptr, 256);
        umma::tc_fence_before();
        __syncthreads();
        umma::tc_fence_after();
        const int warp = tid >> 5;
        tpark = *tptr + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(warp >> 2) * 128u;
    }
    const float lns = stats ? ln_s : 1.f, lnm = stats ? ln_m : 0.f;
    const int half = l >> 1;                                       // complex entries of the packed row (even)

    // twiddle tables, once per CTA: with a prebuilt image (twimg: the six tables in their shared-memory layout) one bulk
    // copy that lands under the prologue's loads and is awaited just before the first twiddle is read; otherwise
    // gathered from the W_n^i table
    uint64_t *twbar = reinterpret_cast<uint64_t *>(sm + Cfg::SMEM / sizeof(float) + 4);
    if (twimg) {
        if (tid == 0) {
            umma::mbar_init(twbar, 1);
            umma::fence_mbar_init();
            umma::mbar_arrive_expect_tx(twbar, Cfg::TWBYTES);
            umma::bulk_g2s(twAr, twimg, Cfg::TWBYTES, twbar);
        }
    } else {
        fft3_fill_twiddles<LOG2M>(tw, twAr, tid);
    }
    __syncthreads();
    bool tw_pending = twimg != nullptr;
    pdl_wait();             // TMEM and twiddles are set up; the rows below read the previous kernel's output

    const int j0 = 2 * tid;                                         // outer-pass butterflies j0, j0 + 1
    const int pb0 = padf(j0);

    // persistent grid: the second CTA of an SM starts half a row period late, so that the memory phases of one
    // (prologue, parked rows, epilogue) fall into the arithmetic phases of the other
    if (stagger_ns > 0 && blockIdx.x >= gridDim.x / 2)
        for (int t = 0; t < stagger_ns; t += 1000) __nanosleep(1000);
#pragma unroll 1
    for (int row = blockIdx.x; row < B * H; row += gridDim.x) {
    const int h = row / B, b = row - h * B;
    const size_t off = ((size_t)b * H + h) * (size_t)l;
    const float4 *xr4 = reinterpret_cast<const float4 *>(x + off);
    float *gr = g + off;
    float *ys = scratch ? scratch + off : nullptr;               // parks the LN-applied input for the odd half
    const float pt = part_t ? part_t[(size_t)b * part_stride_b + h] : 0.f;
    const float4 *st4 = stats ? reinterpret_cast<const float4 *>(stats + (size_t)b * l * 2) : nullptr;
    const float4 *kcr = kc + (size_t)h * (2 * Mh + 2);             // (Mh + 1) entries of two float4
#pragma unroll 1
    for (int odd = 0; odd < 2; ++odd) {
        if (odd) {
            // warm L2 with the row this CTA (persistent grid) or the CTA two waves from now (one row per CTA) takes next:
            // its outer pass then waits for L2 instead of HBM
            const int nrow = row + resident;
            if (nrow < B * H) {
                const int nh = nrow / B, nb = nrow - nh * B;
                const char *nx = reinterpret_cast<const char *>(x + ((size_t)nb * H + nh) * (size_t)l);
                for (int o = tid * 128; o < l * 4; o += NT * 128) prefetch_l2(nx + o);
            }
        }
        // ---- outer forward pass, fused with the prologue: packed inputs i = j + p sub0 (+1 in lane 1), p < 16
        {
            C2 xx[16];
            if (TPARK && odd) {
                // the even half parked y in tensor memory, in register order
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v[16];
                    umma::tmem_ld16(tpark + 16 * c, v);
                    umma::tmem_wait_ld();
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        xx[4 * c + p].x = V2(v[4 * p], v[4 * p + 1]);
                        xx[4 * c + p].y = V2(v[4 * p + 2], v[4 * p + 3]);
                    }
                }
            } else if (odd && ys) {
                // the even half parked y as (re_i, re_i+1, im_i, im_i+1): one round trip, no statistics, no LN
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    const int i = j0 + (p << log2sub0);
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < half) a = *reinterpret_cast<const float4 *>(ys + 2 * i);
                    xx[p].x = V2(a.x, a.y);
                    xx[p].y = V2(a.z, a.w);
                }
            } else {
#pragma unroll
                for (int hb = 0; hb < 4; ++hb) {                    // batches of 4 bound the registers in flight
                    float4 xv[4], sa[4], sb[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int i = j0 + ((4 * hb + p) << log2sub0);
                        xv[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                        sa[p] = make_float4(0.f, 1.f, 0.f, 1.f);
                        sb[p] = sa[p];
                        if (i < half) {
                            xv[p] = __ldg(xr4 + (i >> 1));
                            if (st4) {
                                sa[p] = __ldg(st4 + i);
                                sb[p] = __ldg(st4 + i + 1);
                            }
                        }
                    }
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const int i = j0 + ((4 * hb + p) << log2sub0);
                        const bool in = i < half;
                        const float add = in ? pt : 0.f;
                        const float y0 = (lns * sa[p].y) * (xv[p].x - sa[p].x + lnm) + add;
                        const float y1 = (lns * sa[p].w) * (xv[p].y - sa[p].z + lnm) + add;
                        const float y2 = (lns * sb[p].y) * (xv[p].z - sb[p].x + lnm) + add;
                        const float y3 = (lns * sb[p].w) * (xv[p].w - sb[p].z + lnm) + add;
                        xx[4 * hb + p].x = V2(in ? y0 : 0.f, in ? y2 : 0.f);
                        xx[4 * hb + p].y = V2(in ? y1 : 0.f, in ? y3 : 0.f);
                        if (!TPARK && !odd && ys && in) *reinterpret_cast<float4 *>(ys + 2 * i) = make_float4(y0, y2, y1, y3);
                    }
                    if (TPARK && !odd) {
                        float v[16];
#pragma unroll
                        for (int p = 0; p < 4; ++p) {
                            v[4 * p] = xx[4 * hb + p].x.v.x;
                            v[4 * p + 1] = xx[4 * hb + p].x.v.y;
                            v[4 * p + 2] = xx[4 * hb + p].y.v.x;
                            v[4 * p + 3] = xx[4 * hb + p].y.v.y;
                        }
                        umma::tmem_st16(tpark + 16 * hb, v);
                    }
                }
                if (TPARK && !odd) umma::tmem_wait_st();
            }
            if (odd) s2::rotate_w32<false>(xx);
            s2::RadixS<16, false>::run(xx);
            if (tw_pending) {
                umma::mbar_wait(twbar, 0);
                tw_pending = false;
            }
            C2 v, u0;
            v.x = ld2(twAr + j0);
            v.y = ld2(twAi + j0);
            u0.x = odd ? ld2(twBr + j0) : V2(1.f);
            u0.y = odd ? ld2(twBi + j0) : V2(0.f);
            s2::apply_twiddles16<true>(xx, u0, v);
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int c = fft_brev(q, 4), o = c * sub0 + 2 * ((c * sub0) >> 5);
                st2(re + pb0 + o, xx[q].x);
                st2(im + pb0 + o, xx[q].y);
            }
        }
        __syncthreads();

        mid_passes<LH, false, 1, NP - 2, 0>(re, im, midr, midi, tid);

        // ---- centre (scalar): last forward pass + untangle/product/re-tangle + first inverse pass
        if constexpr (COMPACT) {
            // compact table (16 B per pair): all of a thread's items are requested in one round trip
            static_assert(!COMPACT || RL == 1, "compact table needs the radix-2 centre");
            constexpr int R = 2, G = Mh / R, NITEM = Mh / (2 * R), IPT = NITEM / NT;
            const float4 *kh = kcr + (odd ? 2 * NITEM : 0);
            const float2 w1 = __ldg(tw + 1);                        // W_n^1: odd half frequencies are k + 1
            float4 sa[IPT], sb[IPT];
            float2 wv[IPT];
#pragma unroll
            for (int u = 0; u < IPT; ++u) {
                const int item = tid + u * NT;
                sa[u] = __ldg(kh + 2 * item);
                sb[u] = __ldg(kh + 2 * item + 1);
                wv[u] = __ldg(tw2 + item);
            }
#pragma unroll
            for (int u = 0; u < IPT; ++u) {
                const int item = tid + u * NT;
                const int ga = 2 * item;
                const int gb = odd ? (ga ^ (G - 1)) : (ga == 0 ? 1 : ga ^ ((1 << (31 - __clz(ga))) - 1));
                const int pa = padf(R * ga), pg = padf(R * gb);
                const float2 ar = *reinterpret_cast<const float2 *>(re + pa), ai = *reinterpret_cast<const float2 *>(im + pa);
                const float2 br = *reinterpret_cast<const float2 *>(re + pg), bi = *reinterpret_cast<const float2 *>(im + pg);
                float2 xa[2] = {make_float2(ar.x, ai.x), make_float2(ar.y, ai.y)};
                float2 xb[2] = {make_float2(br.x, bi.x), make_float2(br.y, bi.y)};
                Radix<2, false>::run(xa);
                Radix<2, false>::run(xb);
                if (!odd && item == 0) {
                    const float4 *sp = kcr + 4 * NITEM;             // 32 B forms: DC/Nyquist, slot 1, group 1
                    const float4 dc = __ldg(sp);
                    const float2 a = xa[0];
                    const float p0 = 2.f * (a.x + a.y) * dc.x, pM = 2.f * (a.x - a.y) * dc.y;
                    xa[0] = make_float2(p0 + pM, p0 - pM);
                    float2 d = xa[1];
                    pair_map(xa[1], d, __ldg(sp + 1), __ldg(sp + 2));
                    pair_map(xb[0], xb[1], __ldg(sp + 3), __ldg(sp + 4));
                } else {
                    const float2 wa = odd ? cmul(wv[u], w1) : wv[u];
                    // pair A: w = wa;  pair B: w = -i conj(wa) = (-wa.y, -wa.x)
                    {
                        const float4 c0 = make_float4(fmaf(wa.y, sa[u].z, sa[u].x), fmaf(wa.y, sa[u].w, sa[u].y), -wa.x * sa[u].w, wa.x * sa[u].z);
                        const float4 c1 = make_float4(-c0.z, -c0.w, fmaf(-wa.y, sa[u].z, sa[u].x), fmaf(-wa.y, sa[u].w, sa[u].y));
                        pair_map(xa[0], xb[1], c0, c1);
                    }
                    {
                        const float wc = -wa.y, ws = -wa.x;
                        const float4 c0 = make_float4(fmaf(ws, sb[u].z, sb[u].x), fmaf(ws, sb[u].w, sb[u].y), -wc * sb[u].w, wc * sb[u].z);
                        const float4 c1 = make_float4(-c0.z, -c0.w, fmaf(-ws, sb[u].z, sb[u].x), fmaf(-ws, sb[u].w, sb[u].y));
                        pair_map(xb[0], xa[1], c0, c1);
                    }
                }
                Radix<2, true>::run(xa);
                Radix<2, true>::run(xb);
                *reinterpret_cast<float2 *>(re + pa) = make_float2(xa[0].x, xa[1].x);
                *reinterpret_cast<float2 *>(im + pa) = make_float2(xa[0].y, xa[1].y);
                *reinterpret_cast<float2 *>(re + pg) = make_float2(xb[0].x, xb[1].x);
                *reinterpret_cast<float2 *>(im + pg) = make_float2(xb[0].y, xb[1].y);
            }
            __syncthreads();
        } else {
            centre_wide<Mh, RL, NT>(re, im, odd ? kcr + (Mh + 2) : kcr, kcr + Mh, odd != 0, tid);
        }

        mid_passes<LH, true, 1, NP - 2, 0>(re, im, midr, midi, tid);

        // ---- outer inverse pass + epilogue.  even half: park a[i] in the output row as (re_i, re_i+1, im_i, im_i+1);
        //      odd half: read it back (same thread, same addresses), add W_M^{-i} b[i], GELU, store y[2i .. 2i+3]
        {
            C2 xx[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) {
                const int c = fft_brev(q, 4), o = c * sub0 + 2 * ((c * sub0) >> 5);
                xx[q].x = ld2(re + pb0 + o);
                xx[q].y = ld2(im + pb0 + o);
            }
            C2 v, u0;
            v.x = ld2(twAr + j0);
            v.y = -ld2(twAi + j0);
            u0.x = odd ? ld2(twBr + j0) : V2(1.f);
            u0.y = odd ? -ld2(twBi + j0) : V2(0.f);
            s2::apply_twiddles16<true>(xx, u0, v);
            s2::RadixS<16, true>::run(xx);
            if (TPARK && odd) {
                s2::rotate_w32<true>(xx);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v[16];
                    umma::tmem_ld16(tpark + 64 + 16 * c, v);
                    umma::tmem_wait_ld();
#pragma unroll
                    for (int pp = 0; pp < 4; ++pp) {
                        const int p = 4 * c + pp, i = j0 + (p << log2sub0);
                        const V2 r = s2::gelu_fast2(xx[p].x + V2(v[4 * pp], v[4 * pp + 1])), q = s2::gelu_fast2(xx[p].y + V2(v[4 * pp + 2], v[4 * pp + 3]));
                        if (i < half) *reinterpret_cast<float4 *>(gr + 2 * i) = make_float4(r.v.x, q.v.x, r.v.y, q.v.y);
                    }
                }
            } else if (TPARK) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float v[16];
#pragma unroll
                    for (int pp = 0; pp < 4; ++pp) {
                        v[4 * pp] = xx[4 * c + pp].x.v.x;
                        v[4 * pp + 1] = xx[4 * c + pp].x.v.y;
                        v[4 * pp + 2] = xx[4 * c + pp].y.v.x;
                        v[4 * pp + 3] = xx[4 * c + pp].y.v.y;
                    }
                    umma::tmem_st16(tpark + 64 + 16 * c, v);
                }
                umma::tmem_wait_st();
            } else if (odd) {
                s2::rotate_w32<true>(xx);
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    const int i = j0 + (p << log2sub0);
                    if (i < half) {
                        float4 *dst = reinterpret_cast<float4 *>(gr + 2 * i);
                        const float4 a = *dst;
                        const V2 r = s2::gelu_fast2(xx[p].x + V2(a.x, a.y)), q = s2::gelu_fast2(xx[p].y + V2(a.z, a.w));
                        *dst = make_float4(r.v.x, q.v.x, r.v.y, q.v.y);
                    }
                }
            } else {
#pragma unroll
                for (int p = 0; p < 16; ++p) {
                    const int i = j0 + (p << log2sub0);
                    if (i < half)
                        *reinterpret_cast<float4 *>(gr + 2 * i) =
                            make_float4(xx[p].x.v.x, xx[p].x.v.y, xx[p].y.v.x, xx[p].y.v.y);
                }
            }
        }
        __syncthreads();          // the next half overwrites the planes
    }
    }
    if constexpr (TPARK) {
        umma::tc_fence_before();
        __syncthreads();
        if (tid < 32) {
            umma::tc_fence_after();
            umma::tmem_dealloc(tpark, 256);             // warp 0: lane field 0, column offset 0 = the allocation base
        }
    }
}

// =====================================================================================================
// v5: the UNSPLIT transform on packed fp32 for n <= 8192 (M = 1024, 4096), where shared memory is not the
// occupancy limit: v1's algorithm (one M-point transform per row; the zero upper half of the packed row folded
// into the first radix-16 pass, only the first l outputs computed, no parked rows) with v3's arithmetic (two
// butterflies per thread in the lanes of FADD2/FMUL2/FFMA2 over re/im planes).  Plan 16 . 16 (. 4) + the scalar
// radix-4 centre on the v1 table (fft_table_mode 0).
// =====================================================================================================
template <int LOG2M>
struct Fft5Cfg {
    static_assert(LOG2M == 10 || LOG2M == 12, "v5 plan");
    static constexpr int M = 1 << LOG2M;
    static constexpr int NT = M / 32;                        // one butterfly pair per thread in the radix-16 passes
    static constexpr bool MID4 = (LOG2M == 12);              // M = 4096: a radix-4 pass between pass 1 and the centre
    static constexpr int PLANE = M + M / 16 + 2;
    static constexpr int NTW0 = M / 16, NTW1 = M / 256, NTW2 = 4;            // W_M^j, W_{M/16}^j, W_16^j
    static constexpr int SMEM = (2 * PLANE + 2 * (NTW0 + NTW1 + NTW2)) * (int)sizeof(float);
    static constexpr int MINB = (512 / NT > 16) ? 16 : 512 / NT;
};

// in-place radix-4 pass of span 16 (sub = 4) over the planes: M/8 butterfly pairs, (M/8)/NT per thread
template <int LOG2M, bool INV>
__device__ __forceinline__ void mid4_pass(float *re, float *im, const float *twr, const float *twi, int tid) {
    constexpr int M = 1 << LOG2M, NT = M / 32, NPAIR = M / 8;
#pragma unroll 1
    for (int pi = tid; pi < NPAIR; pi += NT) {
        const int bi = 2 * pi, j = bi & 3;
        const int pb = padf(((bi >> 2) << 4) + j);           // + 4 p stays inside the 16-float block: no extra padding
        C2 w1;
        w1.x = ld2(twr + j);
        w1.y = ld2(twi + j);
        if (INV) w1.y = -w1.y;
        C2 x[4];
        if (!INV) {
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                x[p].x = ld2(re + pb + 4 * p);
                x[p].y = ld2(im + pb + 4 * p);
            }
            s2::RadixS<4, false>::run(x);
            s2::apply_twiddles4(x, w1);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                st2(re + pb + 4 * fft_brev(q, 2), x[q].x);
                st2(im + pb + 4 * fft_brev(q, 2), x[q].y);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                x[q].x = ld2(re + pb + 4 * fft_brev(q, 2));
                x[q].y = ld2(im + pb + 4 * fft_brev(q, 2));
            }
            s2::apply_twiddles4(x, w1);
            s2::RadixS<4, true>::run(x);
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                st2(re + pb + 4 * p, x[p].x);
                st2(im + pb + 4 * p, x[p].y);
            }
        }
    }
    __syncthreads();
}

template <int LOG2M>
__global__ void __launch_bounds__(Fft5Cfg<LOG2M>::NT, Fft5Cfg<LOG2M>::MINB)
fftconv5_kernel(const float *__restrict__ x, const float *__restrict__ stats, const float *__restrict__ part_t,
                long long part_stride_b, float ln_m, float ln_s, const float4 *__restrict__ kc,
                const float2 *__restrict__ tw /* W_n^i, i < M */, float *__restrict__ g, int B, int H, int l) {
    using Cfg = Fft5Cfg<LOG2M>;
    constexpr int M = Cfg::M, NT = Cfg::NT;
    constexpr int log2sub0 = LOG2M - 4, sub0 = 1 << log2sub0;      // outer pass: radix 16, span M
    extern __shared__ float sm[];
    float *re = sm, *im = sm + Cfg::PLANE;
    float *t0r = im + Cfg::PLANE, *t0i = t0r + Cfg::NTW0;
    float *t1r = t0i + Cfg::NTW0, *t1i = t1r + Cfg::NTW1;
    float *t2r = t1i + Cfg::NTW1, *t2i = t2r + Cfg::NTW2;
    const int tid = threadIdx.x;
    const int row = blockIdx.x;
    const int h = row / B, b = row - h * B;
    const size_t off = ((size_t)b * H + h) * (size_t)l;
    const float4 *xr4 = reinterpret_cast<const float4 *>(x + off);
    float *gr = g + off;
    const float pt = part_t ? part_t[(size_t)b * part_stride_b + h] : 0.f;
    const float4 *st4 = stats ? reinterpret_cast<const float4 *>(stats + (size_t)b * l * 2) : nullptr;
    const float4 *kcr = kc + (size_t)h * (M + 2);                  // (M/2 + 1) entries of two float4 (v1 layout)
    const float lns = st4 ? ln_s : 1.f, lnm = st4 ? ln_m : 0.f;
    const int half = l >> 1;

    for (int j = tid; j < Cfg::NTW0; j += NT) {
        const float2 a = tw[2 * j];                                // W_M^j = W_n^{2j}
        t0r[j] = a.x;
        t0i[j] = a.y;
    }
    for (int j = tid; j < Cfg::NTW1; j += NT) {
        const float2 a = tw[32 * j];                               // W_{M/16}^j
        t1r[j] = a.x;
        t1i[j] = a.y;
    }
    for (int j = tid; j < Cfg::NTW2; j += NT) {
        const float2 a = tw[j * (M / 8)];                          // W_16^j
        t2r[j] = a.x;
        t2i[j] = a.y;
    }
    __syncthreads();

    const int j0 = 2 * tid, pb0 = padf(j0);
    C2 one;
    one.x = V2(1.f);
    one.y = V2(0.f);
    // ---- outer forward pass + prologue: packed inputs i = j + p sub0 (+1 in lane 1); p >= 8 lies in the zero padding
    {
        C2 xx[16];
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
            float4 xv[4], sa[4], sb[4];
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int i = j0 + ((4 * hb + p) << log2sub0);
                xv[p] = make_float4(0.f, 0.f, 0.f, 0.f);
                sa[p] = make_float4(0.f, 1.f, 0.f, 1.f);
                sb[p] = sa[p];
                if (i < half) {
                    xv[p] = __ldg(xr4 + (i >> 1));
                    if (st4) {
                        sa[p] = __ldg(st4 + i);
                        sb[p] = __ldg(st4 + i + 1);
                    }
                }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p) {
                const int i = j0 + ((4 * hb + p) << log2sub0);
                const bool in = i < half;
                const float add = in ? pt : 0.f;
                const float y0 = (lns * sa[p].y) * (xv[p].x - sa[p].x + lnm) + add;
                const float y1 = (lns * sa[p].w) * (xv[p].y - sa[p].z + lnm) + add;
                const float y2 = (lns * sb[p].y) * (xv[p].z - sb[p].x + lnm) + add;
                const float y3 = (lns * sb[p].w) * (xv[p].w - sb[p].z + lnm) + add;
                xx[4 * hb + p].x = V2(in ? y0 : 0.f, in ? y2 : 0.f);
                xx[4 * hb + p].y = V2(in ? y1 : 0.f, in ? y3 : 0.f);
            }
        }
        s2::RadixS<16, false, true>::run(xx);
        C2 v;
        v.x = ld2(t0r + j0);
        v.y = ld2(t0i + j0);
        s2::apply_twiddles16<false>(xx, one, v);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int c = fft_brev(q, 4), o = c * sub0 + 2 * ((c * sub0) >> 5);
            st2(re + pb0 + o, xx[q].x);
            st2(im + pb0 + o, xx[q].y);
        }
    }
    __syncthreads();
    mid_pass<LOG2M, 1, false>(re, im, t1r, t1i, tid);
    if constexpr (Cfg::MID4) mid4_pass<LOG2M, false>(re, im, t2r, t2i, tid);
    centre_wide<M, 2, NT>(re, im, kcr, kcr + M, false, tid);
    if constexpr (Cfg::MID4) mid4_pass<LOG2M, true>(re, im, t2r, t2i, tid);
    mid_pass<LOG2M, 1, true>(re, im, t1r, t1i, tid);
    // ---- outer inverse pass + epilogue: only outputs p < 8 can fall inside the row
    {
        C2 xx[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int c = fft_brev(q, 4), o = c * sub0 + 2 * ((c * sub0) >> 5);
            xx[q].x = ld2(re + pb0 + o);
            xx[q].y = ld2(im + pb0 + o);
        }
        C2 v;
        v.x = ld2(t0r + j0);
        v.y = -ld2(t0i + j0);
        s2::apply_twiddles16<false>(xx, one, v);
        s2::RadixS<16, true>::run(xx);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const int i = j0 + (p << log2sub0);
            if (i < half) {
                const V2 r = s2::gelu_fast2(xx[p].x), q = s2::gelu_fast2(xx[p].y);
                *reinterpret_cast<float4 *>(gr + 2 * i) = make_float4(r.v.x, q.v.x, r.v.y, q.v.y);
            }
        }
    }
}

template <int LOG2M>
static int launch_fftconv5(const float *x, const float *stats, const float *part_t, long long psb, float ln_m, float ln_s,
                           const float *kc, const float2 *tw, float *g, int B, int H, int l, cudaStream_t st) {
    using Cfg = Fft5Cfg<LOG2M>;
    fftconv5_kernel<LOG2M><<<B * H, Cfg::NT, Cfg::SMEM, st>>>(x, stats, part_t, psb, ln_m, ln_s, (const float4 *)kc, tw, g, B, H, l);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

bool fftconv5_supported(int lg, const float *x, const float *stats, const float *g, int l) {
    const uintptr_t a = (uintptr_t)x | (uintptr_t)stats | (uintptr_t)g;
    return (lg == 10 || lg == 12) && (l % 4) == 0 && (a & 15) == 0;
}

int fftconv5_launch(int lg, const float *x, const float *stats, const float *part_t, long long psb, float ln_m, float ln_s,
                    const float *kc, const float2 *tw, float *g, int B, int H, int l, cudaStream_t st) {
    switch (lg) {
        case 10: return launch_fftconv5<10>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, g, B, H, l, st);
        case 12: return launch_fftconv5<12>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, g, B, H, l, st);
    }
    set_error("fftconv5: no kernel for log2M=%d", lg);
    return DWB_ERR_UNSUPPORTED;
}

template <int LOG2M, bool COMPACT, bool TPARK>
static int launch_fftconv3_t(const float *x, const float *stats, const float *part_t, long long psb, float ln_m,
                             float ln_s, const float *kc, const float2 *tw, const float2 *tw2, float *g, float *scratch, int B, int H,
                             int l, cudaStream_t st) {
    using Cfg = Fft3Cfg<LOG2M>;
    constexpr int SMEM = Cfg::SMEM + 32;                        // + the TMEM base address word, the twiddle-copy barrier
    // image of the shared-memory twiddle tables: immutable, per (device, log2M), built on first use (outside stream capture)
    static float *twimg[16] = {};
    {
        int dev0 = 0;
        DWB_CUDA(cudaGetDevice(&dev0));
        if (!twimg[dev0 & 15]) {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            cudaStreamIsCapturing(st, &cs);
            if (cs == cudaStreamCaptureStatusNone) {
                float *d = nullptr;
                DWB_CUDA(cudaMalloc(&d, Cfg::TWBYTES));
                fft3_twimg_kernel<LOG2M><<<1, 256, 0, st>>>(tw, d);
                DWB_LAUNCH_CHECK();
                twimg[dev0 & 15] = d;
            }
        }
    }
    static bool attr_set[16] = {};
    int dev = 0;
    DWB_CUDA(cudaGetDevice(&dev));
    if (SMEM > 48 * 1024 && !attr_set[dev & 15]) {
        DWB_CUDA(cudaFuncSetAttribute((fftconv3_kernel<LOG2M, COMPACT, TPARK>), cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_set[dev & 15] = true;
    }
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    // One CTA per row by default.  DWB_FFT_PERS=1 launches a persistent grid instead (one CTA per resident slot walks rows
    // slot, slot + grid, ...; twiddle tables built once per CTA): measured SLOWER on B200 at B = 64, 549 us against 504 us,
    // with or without a half-period start offset between the two CTAs of an SM (DWB_FFT_STAGGER, ns) - the hardware's
    // dynamic CTA dispatch keeps the rows in flight a contiguous window of one or two channels' tables.
    static const bool pers = [] { const char *e = getenv("DWB_FFT_PERS"); return e && atoi(e) == 1; }();
    static const int stagger = [] { const char *e = getenv("DWB_FFT_STAGGER"); return e ? atoi(e) : 0; }();
    const int resident = nsm * Cfg::MINB, grid = pers ? std::min(B * H, resident) : B * H;
    DWB_CUDA(launch_pdl(fftconv3_kernel<LOG2M, COMPACT, TPARK>, dim3(grid, 1, 1), dim3(Cfg::NT), SMEM, st, x, stats, part_t, psb, ln_m, ln_s,
                        (const float4 *)kc, tw, tw2, g, scratch, B, H, l, resident, pers ? stagger : 0, (const float *)twimg[dev & 15]));
    return DWB_OK;
}

template <int LOG2M, bool COMPACT>
static int launch_fftconv3(const float *x, const float *stats, const float *part_t, long long psb, float ln_m,
                           float ln_s, const float *kc, const float2 *tw, const float2 *tw2, float *g, float *scratch, int B, int H,
                           int l, cudaStream_t st) {
    if constexpr (LOG2M == 14) {
        // rows parked in tensor memory (DWB_FFT_TPARK=0: in global memory, as in round 1)
        static const bool tpark = [] { const char *e = getenv("DWB_FFT_TPARK"); return !(e && atoi(e) == 0); }();
        if (tpark) return launch_fftconv3_t<LOG2M, COMPACT, true>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, tw2, g, scratch, B, H, l, st);
    }
    return launch_fftconv3_t<LOG2M, COMPACT, false>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, tw2, g, scratch, B, H, l, st);
}

bool fftconv3_supported(int lg, const float *x, const float *stats, const float *g, int l) {
    const uintptr_t a = (uintptr_t)x | (uintptr_t)stats | (uintptr_t)g;
    return (lg == 12 || lg == 14) && (l % 4) == 0 && (a & 15) == 0;
}

int fftconv3_launch(int lg, const float *x, const float *stats, const float *part_t, long long psb, float ln_m, float ln_s,
                    const float *kc, const float2 *tw, const float2 *tw2, float *g, float *scratch, int B, int H, int l,
                    cudaStream_t st) {
    if (((uintptr_t)scratch & 15) != 0) scratch = nullptr;
    if (tw2) {
        DWB_REQUIRE(lg == 14, DWB_ERR_UNSUPPORTED, "fftconv3: compact table only at n = 32768");
        return launch_fftconv3<14, true>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, tw2, g, scratch, B, H, l, st);
    }
    switch (lg) {
        case 12: return launch_fftconv3<12, false>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, nullptr, g, scratch, B, H, l, st);
        case 14: return launch_fftconv3<14, false>(x, stats, part_t, psb, ln_m, ln_s, kc, tw, nullptr, g, scratch, B, H, l, st);
    }
    set_error("fftconv3: no kernel for log2M=%d", lg);
    return DWB_ERR_UNSUPPORTED;
}

}  // namespace dwb
