// WaveNet residual layer on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM),
// split-bf16 operands, fp32 accumulate.                 models/wavenet.py:82-121, :160-162
//
//   u = h + fc_t(emb)                    (bias BEFORE the zero padding of the dilated conv, :89-95)
//   g = sum_{tap in -1,0,1} W_tap u[t + tap d] + b (+ mel features)         2C rows
//   o = tanh(g[:C]) * sigmoid(g[C:])
//   h' = (h + W_res o + b_res) sqrt(1/2) ;  skip += W_skip o + b_skip
//
// One CTA owns 128 consecutive time steps of one clip and ALL channels; time is the MMA M
// dimension (128 TMEM lanes), output channels are TMEM columns (same orientation as mix_umma.cu).
//
// Phase 1 - dilated conv as ONE implicit GEMM, K = 3C (tap-major), N = 2C.  The A operand of K chunk
//   (tap, 64 channels) is the [128 x 64] slab u[c][t0 + r + (tap-1) d]: eight loader warps read it
//   from global memory (coalesced along time, zero outside [0, L)), split it into bf16 hi/lo and
//   write it K-major / SW128 into a two-slot ring.  Weights are pre-packed at finalize into the exact
//   shared-memory image (split, swizzled, consumption order) and streamed by 1-D bulk async copies
//   (TMA engine) through a four-stage mbarrier ring.  All 2C accumulator columns stay live, so every
//   input slab and every weight byte is fetched exactly once per tile.
// E1 - gate in place: an epilogue thread owns one time step; it reads 16 tanh and 16 sigmoid columns,
//   and writes o as packed bf16 (8 columns hi, 8 columns lo) over the tanh columns it just consumed.
// Phase 2 - res + skip 1x1 as the second GEMM with the gated tile as TMEM A operand (TS form), N in
//   128-column chunks double-buffered in columns [256, 512) (for C = 256 these are the sigmoid
//   columns E1 has drained), so chunk j's epilogue (residual add / skip accumulate, global I/O)
//   runs under chunk j+1's MMAs.
//
// Roles: warps 0-7 = slab loaders, then epilogue (thread = time step r, column group cg = warp / 4);
//        warp 8 = weight producer (one lane); warp 9 = TMEM owner + MMA issuer (one lane).
#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace dwb {
using namespace umma;

constexpr int WU_TT = 128;               // time steps per tile = MMA M
constexpr int WU_STAGE = 32768;          // phase-2 weight stage: [128 rows x 64 k] hi (16 KB) + lo (16 KB)
constexpr int WU_STAGE1 = 65536;         // phase-1 weight stage: [256 rows x 64 k] hi (32 KB) + lo (32 KB)
constexpr int WU_SLAB = 32768;           // input slab:   [128 steps x 64 ch] hi + lo

template <int C, int S>
struct WCfg {
    static constexpr int KC = C / 64;                 // K chunks per tap = K chunks of phase 2
    static constexpr int KC1 = 3 * KC;                // K chunks of phase 1
    static constexpr int NH1 = 2 * C / 256;           // 256-column accumulator blocks of phase 1 (one N = 256 MMA each)
    static constexpr int NJ = (C + S) / 128;          // 128-row output chunks of phase 2
    static constexpr int NP1 = KC1 * NH1;             // 64 KB weight stages of phase 1
    static constexpr int NP2 = NJ * KC;               // 32 KB weight stages of phase 2
    // one 128 KB ring: two 64 KB stages in phase 1 (N = 256 per MMA: the A slab is read once per 256 columns, which
    // keeps the shared-memory bandwidth of MMA reads + bulk-copy writes + slab stores under the tensor pipe's pace),
    // four 32 KB stages in phase 2
    static constexpr int NSW = 4, NSU = 2;
    static constexpr int D2 = 256;                    // first column of the phase-2 accumulators
    // first o chunk that must be complete before phase 2 may overwrite buffer 0 (sigmoid columns of
    // channels [256 - C, 384 - C) live there when 2C > 256)
    static constexpr int KMIN = (2 * C > 256) ? (384 - C) / 64 - 1 : 0;
    static constexpr int EPI = 256, NTHREADS = EPI + 64;
    static constexpr int NBIAS = 4 * C + S;           // bd (2C) | br (C) | bs (S) | fc_t part of this clip (C)
    static constexpr int OFF_SLAB = 0;
    static constexpr int OFF_RING = NSU * WU_SLAB;
    static constexpr int OFF_BIAS = OFF_RING + NSW * WU_STAGE;
    static constexpr int OFF_BAR = OFF_BIAS + NBIAS * 4;
    static constexpr int NBAR = 4 + 2 * NSW + 2 * NSU + 1 + KC + NJ + 2;
    static constexpr int OFF_TPTR = OFF_BAR + NBAR * 8;
    static constexpr int SMEM = OFF_TPTR + 16 + 1024;
    static constexpr size_t IMG_BYTES = (size_t)NP1 * WU_STAGE1 + (size_t)NP2 * WU_STAGE;
    static_assert(C % 128 == 0 && C <= 256 && S % 128 == 0, "wave_umma: C in {128, 256}, S a multiple of 128");
    static_assert(NP1 >= 2 && SMEM <= 227 * 1024, "shared memory");
};

// tanh(a) * sigmoid(b) = (1 - E) / ((1 + E)(1 + F)),  E = exp(-2a), F = exp(-b); arguments clamped so
// the product stays finite.  |error| ~ 1e-7 absolute (2 MUFU.EX2 + 1 MUFU.RCP).
__device__ __forceinline__ float gate_fast(float a, float b) {
    a = fminf(fmaxf(a, -15.f), 15.f);
    b = fminf(fmaxf(b, -30.f), 30.f);
    float E, F, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(E) : "f"(a * -2.8853900817779268f));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(F) : "f"(b * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"((1.0f + E) * (1.0f + F)));
    return (1.0f - E) * r;
}

template <int C, int S, bool COND>
__global__ void __launch_bounds__(WCfg<C, S>::NTHREADS, 1)
wave_block_umma_kernel(WaveBlockArgs a) {
    using W = WCfg<C, S>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *slabs = sm + W::OFF_SLAB, *ring = sm + W::OFF_RING;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + W::OFF_BAR);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(sm + W::OFF_TPTR);
    float *bias_s = reinterpret_cast<float *>(sm + W::OFF_BIAS);
    uint64_t *p1full = bars, *p1empty = p1full + 2, *wfull = p1empty + 2, *wempty = wfull + W::NSW, *ufull = wempty + W::NSW,
             *uempty = ufull + W::NSU,
             *acc1_ready = uempty + W::NSU, *o_ready = acc1_ready + 1, *d2_ready = o_ready + W::KC,
             *d2_free = d2_ready + W::NJ;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, t0 = blockIdx.x * WU_TT, L = a.L, d = a.dilation;
    long long *trace = a.trace ? a.trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
#define WU_TRACE(slot) do { if (trace && tid == 0) trace[slot] = clock64(); } while (0)
#define WU_TRACE_MMA(slot) do { if (trace && lane == 0) trace[slot] = clock64(); } while (0)
    WU_TRACE(0);
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(p1full + i, 1);
            mbar_init(p1empty + i, 1);
        }
        for (int i = 0; i < W::NSW; ++i) {
            mbar_init(wfull + i, 1);
            mbar_init(wempty + i, 1);
        }
        for (int i = 0; i < W::NSU; ++i) {
            mbar_init(ufull + i, 128);
            mbar_init(uempty + i, 1);
        }
        mbar_init(acc1_ready, 1);
        for (int i = 0; i < W::KC; ++i) mbar_init(o_ready + i, W::EPI);
        for (int i = 0; i < W::NJ; ++i) mbar_init(d2_ready + i, 1);
        mbar_init(d2_free, W::EPI);
        mbar_init(d2_free + 1, W::EPI);
        fence_mbar_init();
    }
    // loader / epilogue geometry: one time step per thread, column group cg
    const int q = warp & 3, cg = warp >> 2;
    const int r = 32 * q + lane, t = t0 + r;
    const bool valid = t < L;
    const float *hb = a.h + (size_t)b * C * L;
    // 64 channels of tap (kc / KC) at this thread's time step, raw (the fc_t part is added from shared memory later)
    auto load_slab = [&](int kc, float (&v)[64], bool &inb) {
        const int tap = kc / W::KC, c0 = (kc % W::KC) * 64;
        const int ts = t + (tap - 1) * d;
        inb = valid && ts >= 0 && ts < L;       // zero padding, NOT the t-embedding bias, outside [0, L)   (wavenet.py:91-95)
        const float *hp = hb + (size_t)c0 * L + (inb ? ts : 0);
#pragma unroll
        for (int i = 0; i < 64; ++i, hp += L) v[i] = inb ? __ldg(hp) : 0.f;
    };
    float v0[64];
    bool inb0 = false;
    if (warp < 8) load_slab(cg, v0, inb0);      // first slab of this group: in flight across the setup

    for (int i = tid; i < W::NBIAS; i += W::NTHREADS)
        bias_s[i] = i < 2 * C ? a.bd[i] : (i < 3 * C ? a.br[i - 2 * C] : (i < 3 * C + S ? a.bs[i - 3 * C]
                              : a.part_t[(size_t)b * a.part_stride_b + i - 3 * C - S]));
    if (warp == 9) tmem_alloc(tptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    WU_TRACE(1);

    if (warp == 8) {
        // ================= weight producer =====================================================
        if (lane == 0) {
            const uint8_t *img = a.Wimg;
            for (int i = 0; i < W::NP1; ++i, img += WU_STAGE1) {
                const int s = i & 1, n = i >> 1;
                mbar_wait(p1empty + s, (n & 1) ^ 1);
                mbar_arrive_expect_tx(p1full + s, WU_STAGE1);
                bulk_g2s(ring + (size_t)s * WU_STAGE1, img, WU_STAGE1, p1full + s);
            }
            for (int i = 0; i < W::NP2; ++i, img += WU_STAGE) {
                const int s = i % W::NSW, n = i / W::NSW;
                if (n == 0) {       // first phase-2 use of this 32 KB slot: the 64 KB phase-1 stage under it must be drained
                    const int q = s >> 1, uses = (W::NP1 - q + 1) / 2;
                    mbar_wait(p1empty + q, (uses - 1) & 1);
                } else
                    mbar_wait(wempty + s, (n & 1) ^ 1);
                mbar_arrive_expect_tx(wfull + s, WU_STAGE);
                bulk_g2s(ring + (size_t)s * WU_STAGE, img, WU_STAGE, wfull + s);
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer ==========================================================
        {   // all 32 lanes run the loops; the *_w forms elect the issuing lane
            const uint32_t slab0 = smem_u32(slabs), ring0 = smem_u32(ring);
            constexpr uint32_t idesc = idesc_bf16(128, 128), idesc1 = idesc_bf16(128, 256);
            int i = 0;                                    // phase-2 weight stage counter
            auto next_stage = [&]() {
                const int s = i % W::NSW;
                mbar_wait(wfull + s, (i / W::NSW) & 1);
                tc_fence_after();
                return ring0 + s * WU_STAGE;
            };
            auto done_stage = [&]() {
                mma_commit_w(wempty + (i % W::NSW));
                ++i;
            };
            // ---- phase 1: D[:, 0:2C) = sum over (tap, channel chunk) slabs, one N = 256 MMA per 256 columns
            int i1 = 0;
#pragma unroll 1
            for (int kc = 0; kc < W::KC1; ++kc) {
                const int us = kc % W::NSU;
                mbar_wait(ufull + us, (kc / W::NSU) & 1);
                tc_fence_after();
                if (kc == 0) WU_TRACE_MMA(13);
                const uint32_t abase = slab0 + us * WU_SLAB;
#pragma unroll 1
                for (int nh = 0; nh < W::NH1; ++nh, ++i1) {
                    const int s = i1 & 1;
                    mbar_wait(p1full + s, (i1 >> 1) & 1);
                    tc_fence_after();
                    if (i1 == 0) WU_TRACE_MMA(14);
                    const uint32_t bbase = ring0 + s * WU_STAGE1;
                    if (elect_one()) {      // one election per block of 12 MMAs
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const uint32_t ao = abase + (term == 1 ? WU_SLAB / 2 : 0), bo = bbase + (term == 2 ? WU_STAGE1 / 2 : 0);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_bf16_ss(tmem + nh * 256, smem_desc_sw128(ao + ks * 32), smem_desc_sw128(bo + ks * 32), idesc1,
                                            (kc > 0 || term > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    __syncwarp();
                    mma_commit_w(p1empty + s);
                }
                mma_commit_w(uempty + us);
                if (kc == 0) WU_TRACE_MMA(8);
            }
            mma_commit_w(acc1_ready);
            WU_TRACE_MMA(9);
            // ---- phase 2: [W_res; W_skip] o, 128 output rows at a time, A = packed o in TMEM
#pragma unroll 1
            for (int j = 0; j < W::NJ; ++j) {
                const int buf = j & 1;
                if (j >= 2) {
                    mbar_wait(d2_free + buf, ((j - 2) >> 1) & 1);
                    tc_fence_after();
                }
                const uint32_t dcol = tmem + W::D2 + 128 * buf;
#pragma unroll 1
                for (int kc = 0; kc < W::KC; ++kc) {
                    mbar_wait(o_ready + (kc > W::KMIN ? kc : W::KMIN), 0);
                    mbar_wait(o_ready + kc, 0);
                    tc_fence_after();
                    const uint32_t bbase = next_stage();
                    if (elect_one()) {      // one election per block of 12 MMAs
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const uint32_t bo = bbase + (term == 2 ? WU_STAGE / 2 : 0);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_bf16_ts(dcol, tmem + kc * 64 + ks * 16 + (term == 1 ? 8 : 0), smem_desc_sw128(bo + ks * 32), idesc,
                                            (kc > 0 || term > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    __syncwarp();
                    done_stage();
                }
                mma_commit_w(d2_ready + j);
                if (j == 0) WU_TRACE_MMA(10);
            }
            WU_TRACE_MMA(11);
        }
    } else {
        // ================= loaders, then epilogue: one time step per thread =====================
        const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
        const float *bd_s = bias_s, *br_s = bias_s + 2 * C, *bs_s = bias_s + 3 * C, *pt_s = bias_s + 3 * C + S;

        // ---- slabs kc = cg, cg + 2, ...
#pragma unroll 1
        for (int kc = cg; kc < W::KC1; kc += W::NSU) {
            const int c0 = (kc % W::KC) * 64;
            if (kc != cg) load_slab(kc, v0, inb0);
            if (inb0) {
#pragma unroll
                for (int i4 = 0; i4 < 16; ++i4) {
                    const float4 pv = *reinterpret_cast<const float4 *>(pt_s + c0 + 4 * i4);
                    v0[4 * i4] += pv.x; v0[4 * i4 + 1] += pv.y; v0[4 * i4 + 2] += pv.z; v0[4 * i4 + 3] += pv.w;
                }
            }
            mbar_wait(uempty + cg, ((kc / W::NSU) & 1) ^ 1);
            uint8_t *slab = slabs + cg * WU_SLAB;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                uint4 hi, lo;
                split8(v0 + 8 * c8, hi, lo);
                const uint32_t off = sw128_off(r, c8);
                *reinterpret_cast<uint4 *>(slab + off) = hi;
                *reinterpret_cast<uint4 *>(slab + WU_SLAB / 2 + off) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(ufull + cg);
            if (kc == 0) WU_TRACE(15);
        }

        // ---- E1: o = tanh(ga) sigmoid(gb) -> packed bf16 hi/lo over the consumed tanh columns
        WU_TRACE(2);
        mbar_wait(acc1_ready, 0);
        tc_fence_after();
        WU_TRACE(3);
        const float *cb = COND ? a.cond + (size_t)(a.cond_stride_b ? b : 0) * 2 * C * L + (valid ? t : 0) : nullptr;
#pragma unroll 1
        for (int kc = 0; kc < W::KC; ++kc) {
#pragma unroll 1
            for (int sc = 0; sc < 2; ++sc) {
                const int c0 = kc * 64 + cg * 32 + sc * 16;
                float av[16], gv[16];
                tmem_ld16(tl + c0, av);
                tmem_ld16(tl + C + c0, gv);
                float ba[16], bb[16];
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float4 x4 = *reinterpret_cast<const float4 *>(bd_s + c0 + 4 * i4);
                    const float4 y4 = *reinterpret_cast<const float4 *>(bd_s + C + c0 + 4 * i4);
                    ba[4 * i4] = x4.x; ba[4 * i4 + 1] = x4.y; ba[4 * i4 + 2] = x4.z; ba[4 * i4 + 3] = x4.w;
                    bb[4 * i4] = y4.x; bb[4 * i4 + 1] = y4.y; bb[4 * i4 + 2] = y4.z; bb[4 * i4 + 3] = y4.w;
                }
                if (COND && valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        ba[i] += __ldg(cb + (size_t)(c0 + i) * L);
                        bb[i] += __ldg(cb + (size_t)(C + c0 + i) * L);
                    }
                }
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) av[i] = gate_fast(av[i] + ba[i], gv[i] + bb[i]);
                uint4 h0, l0, h1, l1;
                split8(av, h0, l0);
                split8(av + 8, h1, l1);
                tmem_st8(tl + c0, h0, h1);
                tmem_st8(tl + c0 + 8, l0, l1);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(o_ready + kc);
        }

        WU_TRACE(4);
        // ---- E2: per 128-row chunk of [res; skip], in units of 32 columns per thread: the global inputs of unit u + 1
        //      (h for the residual rows, the running skip sum for the skip rows) are in flight while unit u is processed
        const float rs = 0.70710678118654752440f;
        auto prefetch32 = [&](int u, float (&p)[32]) {
            const int n0 = (u >> 1) * 128 + cg * 64 + (u & 1) * 32;
            if (n0 < C) {
                const float *hp = hb + (size_t)n0 * L + (valid ? t : 0);
#pragma unroll
                for (int i = 0; i < 32; ++i, hp += L) p[i] = valid ? __ldg(hp) : 0.f;
            } else if (!a.first) {
                const float *sp = a.skip + ((size_t)b * S + (n0 - C)) * L + (valid ? t : 0);
#pragma unroll
                for (int i = 0; i < 32; ++i, sp += L) p[i] = valid ? *sp : 0.f;
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) p[i] = 0.f;
            }
        };
        auto process32 = [&](int u, const float (&p)[32]) {
            const int j = u >> 1, hf = u & 1, buf = j & 1, n0 = j * 128 + cg * 64 + hf * 32;
            const bool res = n0 < C;
            if (hf == 0) {
                mbar_wait(d2_ready + j, 0);
                tc_fence_after();
                if (j == 0) WU_TRACE(5);
                if (j == W::NJ - 1) WU_TRACE(6);
            }
            const uint32_t col = tl + W::D2 + 128 * buf + cg * 64 + hf * 32;
            const float *bias = res ? br_s + n0 : bs_s + (n0 - C);
            float *op = res ? a.h_out + ((size_t)b * C + n0) * L + (valid ? t : 0)
                            : a.skip + ((size_t)b * S + (n0 - C)) * L + (valid ? t : 0);
#pragma unroll
            for (int sc = 0; sc < 2; ++sc) {
                float v[16];
                tmem_ld16(col + sc * 16, v);
                tmem_wait_ld();
                if (hf == 1 && sc == 1) {
                    tc_fence_before();
                    mbar_arrive(d2_free + buf);            // accumulator drained: chunk j + 2 may be issued
                }
#pragma unroll
                for (int i4 = 0; i4 < 4; ++i4) {
                    const float4 b4 = *reinterpret_cast<const float4 *>(bias + sc * 16 + 4 * i4);
                    const float bq[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int i = 4 * i4 + e;
                        const float sum = p[sc * 16 + i] + v[i] + bq[e];
                        v[i] = res ? sum * rs : sum;
                    }
                }
                if (valid) {
                    float *oq = op + (size_t)(sc * 16) * L;
#pragma unroll
                    for (int i = 0; i < 16; ++i, oq += L) *oq = v[i];
                }
            }
        };
        float pa[32], pb[32];
        prefetch32(0, pa);
#pragma unroll 1
        for (int u = 0; u < 2 * W::NJ; u += 2) {
            prefetch32(u + 1, pb);
            process32(u, pa);
            if (u + 2 < 2 * W::NJ) prefetch32(u + 2, pa);
            process32(u + 1, pb);
        }
    }
    WU_TRACE(7);
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
    WU_TRACE(12);
}

// finalize: folded fp32 weights (transposed [K][M]) -> the streamed shared-memory image.
// Stage i < KC1*NQ1: phase 1, (kc, nq) = (i / NQ1, i % NQ1): rows nq*128.. of the 2C conv outputs, K chunk kc of
// the tap-major 3C inputs.  Then phase 2: (j, kc): rows j*128.. of [W_res; W_skip], K chunk kc of C.
template <int C, int S>
__global__ void wave_umma_pack_kernel(const float *__restrict__ Wd_t, const float *__restrict__ Wr_t,
                                      const float *__restrict__ Ws_t, uint8_t *__restrict__ img) {
    using W = WCfg<C, S>;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one (stage, row, 16-byte chunk)
    const size_t n1 = (size_t)W::NP1 * 256 * 8, n2 = (size_t)W::NP2 * 128 * 8;
    if (idx >= n1 + n2) return;
    const float *Wt;
    int M, n, k0, row, j8, half;
    size_t base;
    if (idx < n1) {         // phase 1, stage (kc, nh): rows nh*256.. of the 2C conv outputs, K chunk kc of the tap-major 3C inputs
        const int stage = idx / (256 * 8), rem = idx % (256 * 8);
        row = rem / 8; j8 = rem % 8;
        const int kc = stage / W::NH1, nh = stage % W::NH1;
        Wt = Wd_t; M = 2 * C; n = nh * 256 + row; k0 = kc * 64 + j8 * 8;
        base = (size_t)stage * WU_STAGE1; half = WU_STAGE1 / 2;
    } else {                // phase 2, stage (j, kc): rows j*128.. of [W_res; W_skip], K chunk kc of C
        const size_t i2 = idx - n1;
        const int stage = i2 / (128 * 8), rem = i2 % (128 * 8);
        row = rem / 8; j8 = rem % 8;
        const int j = stage / W::KC, kc = stage % W::KC, nn = j * 128 + row;
        k0 = kc * 64 + j8 * 8;
        if (nn < C) { Wt = Wr_t; M = C; n = nn; }
        else { Wt = Ws_t; M = S; n = nn - C; }
        base = (size_t)W::NP1 * WU_STAGE1 + (size_t)stage * WU_STAGE; half = WU_STAGE / 2;
    }
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float w0 = Wt[(size_t)(k0 + 2 * e) * M + n], w1 = Wt[(size_t)(k0 + 2 * e + 1) * M + n];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(w0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(w1 - __bfloat162float(h1));
        hp[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lp[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t off = base + (size_t)row * 128 + ((j8 ^ (row & 7)) << 4);
    *reinterpret_cast<uint4 *>(img + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4 *>(img + off + half) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
}

bool wave_umma_supported(int C, int S) { return (C == 128 || C == 256) && (S == 128 || S == 256); }

template <int C, int S>
static int pack_wave_umma(const float *Wd_t, const float *Wr_t, const float *Ws_t, uint8_t *img, cudaStream_t st) {
    const unsigned grid = (unsigned)ceil_div64((int64_t)WCfg<C, S>::NP1 * 256 * 8 + (int64_t)WCfg<C, S>::NP2 * 128 * 8, 256);
    wave_umma_pack_kernel<C, S><<<grid, 256, 0, st>>>(Wd_t, Wr_t, Ws_t, img);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

size_t wave_umma_image_bytes(int C, int S) {
    if (!wave_umma_supported(C, S)) return 0;
    return (size_t)(3 * (C / 64) * (2 * C / 128) + ((C + S) / 128) * (C / 64)) * WU_STAGE;
}

int wave_umma_pack(int C, int S, const float *Wd_t, const float *Wr_t, const float *Ws_t, uint8_t *img, cudaStream_t st) {
    if (C == 128 && S == 128) return pack_wave_umma<128, 128>(Wd_t, Wr_t, Ws_t, img, st);
    if (C == 128 && S == 256) return pack_wave_umma<128, 256>(Wd_t, Wr_t, Ws_t, img, st);
    if (C == 256 && S == 128) return pack_wave_umma<256, 128>(Wd_t, Wr_t, Ws_t, img, st);
    if (C == 256 && S == 256) return pack_wave_umma<256, 256>(Wd_t, Wr_t, Ws_t, img, st);
    set_error("wave_umma_pack: C=%d S=%d unsupported", C, S);
    return DWB_ERR_UNSUPPORTED;
}

template <int C, int S>
static int launch_wave_umma(const WaveBlockArgs &a, int B, cudaStream_t st) {
    using W = WCfg<C, S>;
    auto k = a.cond ? wave_block_umma_kernel<C, S, true> : wave_block_umma_kernel<C, S, false>;
    DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)W::SMEM));
    k<<<dim3(ceil_div(a.L, WU_TT), B), W::NTHREADS, W::SMEM, st>>>(a);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

int wave_block_umma_launch(const WaveBlockArgs &a, int B, cudaStream_t st) {
    DWB_REQUIRE(a.Wimg, DWB_ERR_STATE, "wave_umma: weights were not packed");
    DWB_REQUIRE((int64_t)2 * a.C * a.L < (int64_t)1 << 31, DWB_ERR_UNSUPPORTED, "wave_umma: C*L too large");
    if (a.C == 128 && a.S == 128) return launch_wave_umma<128, 128>(a, B, st);
    if (a.C == 128 && a.S == 256) return launch_wave_umma<128, 256>(a, B, st);
    if (a.C == 256 && a.S == 128) return launch_wave_umma<256, 128>(a, B, st);
    if (a.C == 256 && a.S == 256) return launch_wave_umma<256, 256>(a, B, st);
    set_error("wave_umma: C=%d S=%d unsupported", a.C, a.S);
    return DWB_ERR_UNSUPPORTED;
}

}  // namespace dwb
