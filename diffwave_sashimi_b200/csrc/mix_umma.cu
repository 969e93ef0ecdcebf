// DiffWaveBlock channel mixing on the 5th-generation tensor cores (tcgen05.mma, accumulators in
// TMEM), split-bf16 operands, fp32 accumulate.                  models/sashimi.py:157-182
//
//   q = Wo g + bo ; y = q[:H] * sigmoid(q[H:]) (+cond) ; x1 = x + y          (s4.py:1435, sashimi.py:177)
//   x2 = x1 + W2 gelu(W1 LN2(x1) + b1) + b2 (+skip) ; stats(x2)              (sashimi.py:179-182)
//
// Orientation.  One CTA owns a tile of 128 consecutive time steps of one clip and ALL channels.
// The three contractions are issued as  D[time][out_ch] = sum_k Act[time][k] * W[out_ch][k]:
// time is the MMA M dimension (128 = the TMEM lanes), output channels are TMEM columns.  An
// epilogue thread therefore owns ONE time step and sees every channel of it in its own TMEM lane:
// the TransposedLN statistics over channels, the GLU pairing (q[h], q[H+h]) and both residual adds
// are thread-local, and x1 simply stays in the TMEM columns that the third contraction later
// accumulates into (x2 = x1 + W2 hidden falls out of the MMA, no separate add).
//
// Operands.  fp32 values are split v = hi + lo (two bf16) and each product is three MMAs
// (hi*hi + lo*hi + hi*lo), ~2^-17 relative: single-pass bf16/tf32 would eat the whole 1e-3 parity
// budget.  Activations are written by the epilogue threads straight into
// the K-major SW128 layout the MMA reads (one 16-byte chunk = 8 channels of one time step);
// weights are packed once at finalize into the exact shared-memory image (split, swizzled, in
// consumption order) and streamed per tile with 1-D bulk async copies through a small ring.
//
// Roles: warps [0, 4*CS) epilogue (CS threads per time step, each a contiguous column group),
// warp 4*CS = weight producer (one lane), warp 4*CS+1 = TMEM owner + MMA issuer (one lane).
#include <cuda.h>
#include <stdlib.h>

#include <string.h>

#include <algorithm>
#include <string>

#include "common.cuh"
#include "kernels.h"
#include "fft_simd2.cuh"
#include "umma.cuh"

namespace dwb {
using namespace umma;

constexpr int UM_TT = 128;                // time steps per tile = MMA M
constexpr int UM_STAGE = 32768;           // weight ring stage
constexpr int UM_SLOT = 32768;            // activation operand slot: [128 x 64] hi (16 KB) + lo (16 KB)

template <int H, int CS>
struct UCfg {
    static constexpr int F = 2 * H;
    static constexpr int EPI = 128 * CS;                 // epilogue threads
    static constexpr int NTHREADS = EPI + 64;
    static constexpr int KC1 = H / 64;                   // K chunks of G1 / G2
    static constexpr int NC1 = 2 * H / 128;              // N chunks (128 columns) of G1 / G2
    static constexpr int KC3 = F / 64;                   // K chunks of G3
    static constexpr int NR3 = H < 128 ? H : 128;        // rows per weight block of G3
    static constexpr int NC3 = H / NR3;
    static constexpr int BPS3 = UM_STAGE / (NR3 * 256);  // G3 weight blocks per ring stage
    static constexpr int NSTG = 2 * NC1 * KC1 + (KC3 * NC3) / BPS3;
    static constexpr int NSLOT = (H == 64) ? 2 : 4;
    static constexpr int NS = (H == 64) ? 1 : 2;         // ring depth
    static constexpr int TMEM_COLS = (3 * H <= 256) ? 256 : 512;
    static constexpr int R3 = 2 * H;                     // first TMEM column of x1 / acc3
    static constexpr int HID_ARRIVE = 128 * (CS >= 2 ? CS / 2 : 1);
    static constexpr int NBAR = 2 * NS + 2 + KC3 + 2 * NC1 + 1;
    // shared memory carve-up (bytes from the 1024-aligned base)
    static constexpr int OFF_SLOT = 0;
    static constexpr int OFF_RING = OFF_SLOT + NSLOT * UM_SLOT;
    static constexpr int OFF_BIAS = OFF_RING + NS * UM_STAGE;          // bo' (2H) b1 (F) b2 (H)
    static constexpr int OFF_EX = OFF_BIAS + 5 * H * 4;                // 2 exchanges x CS x 128 x (mean, M2)
    static constexpr int OFF_BAR = OFF_EX + 2 * CS * 128 * 8;
    static constexpr int OFF_TPTR = OFF_BAR + NBAR * 8;
    static constexpr int SMEM = OFF_TPTR + 16 + 1024;                  // + alignment slack
    static constexpr size_t IMG_BYTES = (size_t)NSTG * UM_STAGE;
    static_assert(H % 64 == 0 && (KC3 * NC3) % BPS3 == 0, "tiling");
    static_assert(3 * H <= 512, "TMEM budget: acc (2H) + x1/acc3 (H) columns");
    static constexpr int hid_slot(int kc) { return H == 64 ? kc : (kc + 2) % 4; }
};

// running (mean, M2) over n values + a chunk of 16 -> n + 16 values (Chan et al. pairwise update)
__device__ __forceinline__ void stat_merge16(const float (&v)[16], int n, float &mean, float &M2) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += v[i];
    const float cm = s * (1.0f / 16.0f);
    float c2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float d = v[i] - cm;
        c2 = fmaf(d, d, c2);
    }
    // n == 0 with (mean, M2) = (0, 0) reduces to (cm, c2)
    const float delta = cm - mean, w = __fdividef(16.0f, (float)(n + 16));
    mean = fmaf(delta, w, mean);
    M2 += c2 + delta * delta * ((float)n * w);
}


// debug: a per-thread pseudo-random sleep of up to `ns` nanoseconds (MixArgs::jitter), keyed on (thread, point)
__device__ __forceinline__ void jitter_sleep(int ns, unsigned key) {
    if (ns > 0) {
        unsigned h = (threadIdx.x * 2654435761u) ^ (key * 40503u + blockIdx.x * 9973u);
        h ^= h >> 13;
        h *= 0x5bd1e995u;
        h ^= h >> 15;
        __nanosleep(h % (unsigned)ns);
    }
}

// ---- epilogue helpers shared by the fused kernels --------------------------------------------
// channel stride in floats: a compile-time constant for the shapes of the BASELINE configs (the loads
// and stores of a thread's channels then need no address arithmetic at all: [base + immediate]),
// the runtime l otherwise (one IMAD.WIDE per access)
template <int LC>
__device__ __forceinline__ const float *chan(const float *base, int i, int l) {
    if (LC) return base + (size_t)i * LC;
    return reinterpret_cast<const float *>(reinterpret_cast<const char *>(base) + (unsigned long long)(unsigned)(4 * l) * (unsigned)i);
}
template <int LC>
__device__ __forceinline__ float *chan(float *base, int i, int l) {
    if (LC) return base + (size_t)i * LC;
    return reinterpret_cast<float *>(reinterpret_cast<char *>(base) + (unsigned long long)(unsigned)(4 * l) * (unsigned)i);
}

// running sums of d = v - pivot and d^2 over 16 values, two values per packed instruction.  The pivot
// (the thread's first value) keeps the final M2 = sum d^2 - (sum d)^2 / n free of the cancellation a raw
// sum of squares has when |mean| >> std.
__device__ __forceinline__ void stat_acc16(const float (&v)[16], const s2::V2 &piv, s2::V2 &sd, s2::V2 &sq) {
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
        const s2::V2 d = s2::V2(v[i], v[i + 1]) - piv;
        sd = sd + d;
        sq = s2::fma(d, d, sq);
    }
}
__device__ __forceinline__ void stat_finish(const s2::V2 &sd, const s2::V2 &sq, float piv, int n, float &mean, float &M2) {
    const float S = sd.v.x + sd.v.y, Q = sq.v.x + sq.v.y, inv = 1.0f / (float)n;
    mean = fmaf(S, inv, piv);
    M2 = fmaxf(fmaf(-S * inv, S, Q), 0.f);
}

// fp32 -> (hi, lo) bf16 split of 8 consecutive K elements with the residuals as packed subtractions
__device__ __forceinline__ void split8p(const float *v, uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        uint32_t hp;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hp) : "f"(b), "f"(a));     // low half = a, high half = b
        const s2::V2 d = s2::V2(a, b) - s2::V2(__uint_as_float(hp << 16), __uint_as_float(hp & 0xFFFF0000u));
        uint32_t lp;
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(lp) : "f"(d.v.y), "f"(d.v.x));
        h[i] = hp;
        l[i] = lp;
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int H, int CS>
__global__ void __launch_bounds__(UCfg<H, CS>::NTHREADS, (H == 64) ? 2 : 1)
sashimi_mix_umma_kernel(MixArgs a) {
    using C = UCfg<H, CS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *slots = sm + C::OFF_SLOT, *ring = sm + C::OFF_RING;
    float *bias_s = reinterpret_cast<float *>(sm + C::OFF_BIAS);
    float2 *ex = reinterpret_cast<float2 *>(sm + C::OFF_EX);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(sm + C::OFF_TPTR);
    uint64_t *wfull = bars, *wempty = bars + C::NS, *g_ready = bars + 2 * C::NS, *z_ready = g_ready + 1,
             *hid_ready = z_ready + 1, *acc1_ready = hid_ready + C::KC3, *acc2_ready = acc1_ready + C::NC1,
             *acc3_ready = acc2_ready + C::NC1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = a.rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y, t0 = (a.rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * UM_TT, l = a.l;
    long long *trace = a.trace ? a.trace + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 16 : nullptr;
#define UM_TRACE(slot) do { if (trace && tid == 0) trace[slot] = clock64(); } while (0)
#define UM_TRACE_MMA(slot) do { if (trace && lane == 0) trace[slot] = clock64(); } while (0)
    UM_TRACE(0);

    // epilogue threads: one time step each (row r of the tile), column group cg
    const bool is_epi = warp < 4 * CS;
    const int q = warp & 3, cg = warp >> 2;
    const int r = 32 * q + lane, t = t0 + r;
    const bool valid = is_epi && t < l;
    const size_t brow = (size_t)b * H * l + (valid ? t : 0);
    constexpr int PER = H / CS;                        // channels per epilogue thread
    float gin[PER];                                    // g of this thread's channels: in flight across the setup
    if (is_epi) {
        const float *gp = a.g + brow + cg * PER * l;
#pragma unroll
        for (int i = 0; i < PER; ++i, gp += l) gin[i] = valid ? __ldg(gp) : 0.f;
    }

    if (tid == 0) {
        for (int i = 0; i < C::NS; ++i) {
            mbar_init(wfull + i, 1);
            mbar_init(wempty + i, 1);
        }
        mbar_init(g_ready, C::EPI);
        mbar_init(z_ready, C::EPI);
        for (int i = 0; i < C::KC3; ++i) mbar_init(hid_ready + i, C::HID_ARRIVE);
        for (int i = 0; i < C::NC1; ++i) {
            mbar_init(acc1_ready + i, 1);
            mbar_init(acc2_ready + i, 1);
        }
        mbar_init(acc3_ready, 1);
        fence_mbar_init();
    }
    for (int i = tid; i < 5 * H; i += C::NTHREADS) bias_s[i] = a.bimg[i];
    if (warp == 4 * CS + 1) tmem_alloc(tptr, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    UM_TRACE(1);

    if (warp == 4 * CS) {
        // ================= weight producer: the packed image, one ring stage at a time =========
        if (lane == 0) {
            for (int i = 0; i < C::NSTG; ++i) {
                const int s = i % C::NS, n = i / C::NS;
                mbar_wait(wempty + s, (n & 1) ^ 1);
                mbar_arrive_expect_tx(wfull + s, UM_STAGE);
                bulk_g2s(ring + (size_t)s * UM_STAGE, a.Wimg + (size_t)i * UM_STAGE, UM_STAGE, wfull + s);
            }
        }
    } else if (warp == 4 * CS + 1) {
        // ================= MMA issuer ==========================================================
        {   // all 32 lanes run the loops; the *_w forms elect the issuing lane
            const uint32_t slot0 = smem_u32(slots), ring0 = smem_u32(ring);
            // one [128 x NR] x K=64 block: 3 split terms x 4 k-steps
            auto issue_block = [&](uint32_t d, uint32_t abase, uint32_t bbase, int NR, bool acc0) {
                const uint32_t idesc = idesc_bf16(128, NR);
                if (elect_one()) {      // one election per block of 12 MMAs
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t ao = abase + (term == 1 ? UM_SLOT / 2 : 0);
                        const uint32_t bo = bbase + (term == 2 ? NR * 128 : 0);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_bf16_ss(d, smem_desc_sw128(ao + ks * 32), smem_desc_sw128(bo + ks * 32), idesc,
                                        (acc0 || term > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                __syncwarp();
            };
            int i = 0;
#pragma unroll 1
            for (int gemm = 0; gemm < 2; ++gemm) {
                mbar_wait(gemm == 0 ? g_ready : z_ready, 0);
                tc_fence_after();
                UM_TRACE_MMA(12 + gemm);
#pragma unroll 1
                for (int nc = 0; nc < C::NC1; ++nc) {
#pragma unroll 1
                    for (int kc = 0; kc < C::KC1; ++kc, ++i) {
                        const int s = i % C::NS;
                        mbar_wait(wfull + s, (i / C::NS) & 1);
                        tc_fence_after();
                        issue_block(tmem + nc * 128, slot0 + kc * UM_SLOT, ring0 + s * UM_STAGE, 128, kc > 0);
                        mma_commit_w(wempty + s);
                    }
                    mma_commit_w((gemm == 0 ? acc1_ready : acc2_ready) + nc);
                }
            }
#pragma unroll 1
            for (int kc = 0; kc < C::KC3; ++kc) {
                mbar_wait(hid_ready + kc, 0);
                tc_fence_after();
#pragma unroll 1
                for (int nc = 0; nc < C::NC3; ++nc) {
                    const int j = kc * C::NC3 + nc, s = i % C::NS;
                    if (j % C::BPS3 == 0) {
                        mbar_wait(wfull + s, (i / C::NS) & 1);
                        tc_fence_after();
                    }
                    issue_block(tmem + C::R3 + nc * 128, slot0 + C::hid_slot(kc) * UM_SLOT,
                                ring0 + s * UM_STAGE + (j % C::BPS3) * C::NR3 * 256, C::NR3, true);
                    if (j % C::BPS3 == C::BPS3 - 1) {
                        mma_commit_w(wempty + s);
                        ++i;
                    }
                }
            }
            mma_commit_w(acc3_ready);
            UM_TRACE_MMA(14);
        }
    } else {
        // ================= epilogue threads ======================================================
        const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
        const float *xp = a.x + brow;
        float *op = a.out + brow;
        const float *bo_s = bias_s, *b1_s = bias_s + 2 * H, *b2_s = bias_s + 4 * H;
        constexpr int PP = 64 / CS;                        // GLU pairs per thread per N chunk

        // ---- g: split, store as the A operand of G1
        {
#pragma unroll
            for (int c8 = 0; c8 < PER / 8; ++c8) {
                const int h0 = cg * PER + c8 * 8;
                uint4 hi, lo;
                split8(gin + 8 * c8, hi, lo);
                uint8_t *slot = slots + (h0 >> 6) * UM_SLOT;
                const uint32_t off = sw128_off(r, (h0 & 63) >> 3);
                *reinterpret_cast<uint4 *>(slot + off) = hi;
                *reinterpret_cast<uint4 *>(slot + UM_SLOT / 2 + off) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(g_ready);
            UM_TRACE(2);
        }
        // ---- x of this thread's channels -> TMEM R3 while G1 runs (E1 turns it into x1 in place)
        {
            float xin[PER];
#pragma unroll
            for (int nc = 0; nc < C::NC1; ++nc) {
                const float *xq = xp + (nc * 64 + cg * PP) * l;
#pragma unroll
                for (int i = 0; i < PP; ++i, xq += l) xin[nc * PP + i] = valid ? __ldg(xq) : 0.f;
            }
#pragma unroll
            for (int nc = 0; nc < C::NC1; ++nc)
#pragma unroll
                for (int sc = 0; sc < PP / 16; ++sc) {
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = xin[nc * PP + sc * 16 + i];
                    tmem_st16(tl + C::R3 + nc * 64 + cg * PP + sc * 16, v);
                }
            tmem_wait_st();
            UM_TRACE(3);
        }

        // ---- E1: GLU + residual -> x1 (TMEM R3), LN2 statistics
        float mean = 0.f, M2 = 0.f;
        {
            int n = 0;
#pragma unroll 1
            for (int nc = 0; nc < C::NC1; ++nc) {
                mbar_wait(acc1_ready + nc, 0);
                tc_fence_after();
                if (nc == 0) UM_TRACE(4);
#pragma unroll 1
                for (int sc = 0; sc < PP / 16; ++sc) {
                    const int p0 = cg * PP + sc * 16, h0 = nc * 64 + p0;
                    float xv[16], av[16], gv[16];
                    tmem_ld16(tl + nc * 128 + p0, av);
                    tmem_ld16(tl + nc * 128 + 64 + p0, gv);
                    tmem_ld16(tl + C::R3 + h0, xv);
                    tmem_wait_ld();
                    const float *ba = bo_s + nc * 128 + p0;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        // two channels per packed instruction (FADD2 / FMUL2 / FFMA2)
                        s2::V2 y = (s2::V2(av[i], av[i + 1]) + s2::V2(ba[i], ba[i + 1])) *
                                   s2::sigmoid_scaled2(s2::V2(gv[i], gv[i + 1]), s2::V2(ba[64 + i], ba[64 + i + 1]));
                        y = y + s2::V2(xv[i], xv[i + 1]);
                        xv[i] = y.v.x;
                        xv[i + 1] = y.v.y;
                        if (a.cond && valid) {
                            xv[i] += __ldg(a.cond + (size_t)(a.cond_stride_b ? b : 0) * H * l + t + (h0 + i) * l);
                            xv[i + 1] += __ldg(a.cond + (size_t)(a.cond_stride_b ? b : 0) * H * l + t + (h0 + i + 1) * l);
                        }
                    }
                    stat_merge16(xv, n, mean, M2);
                    n += 16;
                    tmem_st16(tl + C::R3 + h0, xv);
                }
            }
            tmem_wait_st();
            UM_TRACE(5);
        }
        if (CS > 1) {
            ex[cg * 128 + r] = make_float2(mean, M2);
            asm volatile("bar.sync 1, %0;" ::"n"(C::EPI) : "memory");
            float ms = 0.f;
#pragma unroll
            for (int c = 0; c < CS; ++c) ms += ex[c * 128 + r].x;
            const float mt = ms * (1.0f / CS);
            float m2 = 0.f;
#pragma unroll
            for (int c = 0; c < CS; ++c) {
                const float2 e = ex[c * 128 + r];
                const float d = e.x - mt;
                m2 += e.y + d * d * (float)(H / CS);
            }
            mean = mt;
            M2 = m2;
        }
        // ---- z = LN2(x1), split, store as the A operand of G2 (the slots of g: G1 has completed)
        {
            const float rstd = valid ? rsqrtf(M2 * (1.0f / H)) : 0.f;
            const float sc_a = a.ln2_s * rstd, sh = a.ln2_m - mean;
#pragma unroll 1
            for (int it = 0; it < PER / 16; ++it) {
                const int nc = it / (PP / 16), sc = it % (PP / 16);
                const int h0 = nc * 64 + cg * PP + sc * 16;
                float v[16];
                tmem_ld16(tl + C::R3 + h0, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = sc_a * (v[i] + sh);
                uint8_t *slot = slots + (h0 >> 6) * UM_SLOT;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint4 hi, lo;
                    split8(v + 8 * hh, hi, lo);
                    const uint32_t off = sw128_off(r, ((h0 & 63) >> 3) + hh);
                    *reinterpret_cast<uint4 *>(slot + off) = hi;
                    *reinterpret_cast<uint4 *>(slot + UM_SLOT / 2 + off) = lo;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(z_ready);
            UM_TRACE(6);
        }

        // ---- E2: hidden = gelu(W1 z + b1), split, store as the A operand of G3
        {
            constexpr int PERF = 128 / CS;                 // f columns per thread per N chunk
#pragma unroll 1
            for (int nc = 0; nc < C::NC1; ++nc) {
                mbar_wait(acc2_ready + nc, 0);
                tc_fence_after();
                if (nc == 0) UM_TRACE(7);
#pragma unroll 1
                for (int sc = 0; sc < PERF / 16; ++sc) {
                    const int col = cg * PERF + sc * 16, f0 = nc * 128 + col;
                    float v[16];
                    tmem_ld16(tl + nc * 128 + col, v);
                    tmem_wait_ld();
                    const float *bb = b1_s + f0;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const s2::V2 r = s2::gelu_fast2(s2::V2(v[i], v[i + 1]) + s2::V2(bb[i], bb[i + 1]));
                        v[i] = r.v.x;
                        v[i + 1] = r.v.y;
                    }
                    const int kc = f0 >> 6;
                    uint8_t *slot = slots + C::hid_slot(kc) * UM_SLOT;
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        uint4 hi, lo;
                        split8(v + 8 * hh, hi, lo);
                        const uint32_t off = sw128_off(r, ((f0 & 63) >> 3) + hh);
                        *reinterpret_cast<uint4 *>(slot + off) = hi;
                        *reinterpret_cast<uint4 *>(slot + UM_SLOT / 2 + off) = lo;
                    }
                    // last 16 columns this thread contributes to K chunk kc
                    if (((f0 + 16) & 63) == 0 || sc == PERF / 16 - 1) {
                        fence_proxy_async_smem();
                        tc_fence_before();
                        mbar_arrive(hid_ready + kc);
                    }
                }
            }
        }

        // ---- E3: x2 = acc3 (= x1 + W2 hidden) + b2 (+skip); store; statistics for the next norm
        {
            float sk[PER];
            if (a.skip) {
                const float *sp = a.skip + brow + cg * PER * l;
#pragma unroll
                for (int i = 0; i < PER; ++i, sp += l) sk[i] = valid ? __ldg(sp) : 0.f;
            }
            UM_TRACE(8);
            mbar_wait(acc3_ready, 0);
            tc_fence_after();
            UM_TRACE(9);
            int n = 0;
            mean = 0.f;
            M2 = 0.f;
#pragma unroll
            for (int sc = 0; sc < PER / 16; ++sc) {
                const int h0 = cg * PER + sc * 16;
                float v[16];
                tmem_ld16(tl + C::R3 + h0, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    v[i] += b2_s[h0 + i];
                    if (a.skip) v[i] += sk[sc * 16 + i];
                }
                if (valid) {
                    float *oq = op + h0 * l;
#pragma unroll
                    for (int i = 0; i < 16; ++i, oq += l) *oq = v[i];
                }
                stat_merge16(v, n, mean, M2);
                n += 16;
            }
            if (CS > 1) {
                float2 *ex3 = ex + CS * 128;
                ex3[cg * 128 + r] = make_float2(mean, M2);
                asm volatile("bar.sync 1, %0;" ::"n"(C::EPI) : "memory");
                float ms = 0.f;
#pragma unroll
                for (int c = 0; c < CS; ++c) ms += ex3[c * 128 + r].x;
                const float mt = ms * (1.0f / CS);
                float m2 = 0.f;
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                    const float2 e = ex3[c * 128 + r];
                    const float d = e.x - mt;
                    m2 += e.y + d * d * (float)(H / CS);
                }
                mean = mt;
                M2 = m2;
            }
            if (cg == 0 && valid)
                *reinterpret_cast<float2 *>(a.stats_out + ((size_t)b * l + t) * 2) = make_float2(mean, rsqrtf(M2 * (1.0f / H)));
        }
    }
    UM_TRACE(10);
    tc_fence_before();
    __syncthreads();
    UM_TRACE(11);
    if (warp == 4 * CS + 1) {
        tc_fence_after();
        tmem_dealloc(tmem, C::TMEM_COLS);
    }
}

// =========================================================================================
// Persistent variant: one CTA per SM walks a strided list of tiles; NG independent groups of
// epilogue warps (each with its own MMA-issuer warp, barriers and TMEM columns) keep NG tiles in
// flight, so the waits of one group are filled by the other's arithmetic.  Barriers, TMEM and
// (H = 64) the whole 96 KB weight image are set up once per CTA; H = 128 has TMEM for one group
// only and streams its 384 KB weight image through a ring.
//
// No activation operand touches shared memory: the A operands of all three GEMMs live in TMEM
// (tcgen05.mma TS form).  Per group:  ACC [0, 2H) accumulator of G1 / G2, then - written in place
// by E2 over the columns it has just read - the split hidden tile (per 16 f-columns: 8 packed hi
// columns, 8 packed lo columns = one K = 16 step each);  R3 [2H, 3H) x1 -> x2 (G3 accumulates onto
// x1);  AOP [3H, 4H) the split g tile, then the split z = LN2(x1) tile, same [8 hi | 8 lo] layout.
// Shared memory holds the weights and, per group, two fp32 [H x 128] staging buffers that a
// producer warp fills one tile ahead with 1-D bulk copies (one 512-byte row per channel): XB with
// the next tile's x, GB with its g.  The epilogue threads read them with plain LDS, so nothing
// they execute waits on global memory except the optional skip tensor, and up to 4 x 32 KB of
// input are in flight per SM at any time (register prefetch kept ~40 KB in flight and paced the
// kernel at ~3 TB/s: Little's law).  Tiles are software-pipelined inside a group: the next tile's
// g operand is written as soon as G3 has been awaited, so G1(next) runs under E3(this).
// STAGE = false (sequence lengths outside the BASELINE set, unaligned tensors) takes x and g
// through registers with plain loads instead; everything else is identical.
// =========================================================================================
template <int H, int CS>
struct PCfg {
    using U = UCfg<H, CS>;
    static constexpr int NG = (H == 64) ? 2 : 1;
    static constexpr bool RESIDENT = (H == 64);
    static constexpr int GW = 4 * CS, EPI = 128 * CS;
    // warps: NG epilogue groups | NG MMA issuers | weight producer | (streaming weights only) input-staging producer;
    // with resident weights the weight producer is idle after its first copies and stages the inputs itself
    static constexpr int NTHREADS = NG * EPI + NG * 32 + 32 + (RESIDENT ? 0 : 32);
    static constexpr int NS = RESIDENT ? U::NSTG : 3;          // weight buffers (RESIDENT: one per stage)
    // biases in shared memory.  (H = 128 measured with a third ring stage in their place and the biases through L1:
    // 191 us against 167 us - the uniform bias loads sit on the epilogues' critical path.)
    static constexpr bool BIAS_SMEM = true;
    static constexpr int GCOLS = 512 / NG;                     // TMEM columns per group
    static constexpr int R3 = 2 * H, AOP = 3 * H;
    static constexpr int STG = H * 512;                        // one staged fp32 [H x 128] tile
    static constexpr int NBAR_G = 2 + U::KC3 + 2 * U::NC1 + 1 + 4;      // g z hid[] acc1[] acc2[] acc3 | xfull xempty gfull gempty
    static constexpr int NBAR = NG * NBAR_G + 2 * NS;
    static constexpr int OFF_STG = 0;
    static constexpr int OFF_W = OFF_STG + NG * 2 * STG;
    static constexpr int OFF_BIAS = OFF_W + NS * UM_STAGE;
    static constexpr int OFF_BAR = OFF_BIAS + (BIAS_SMEM ? 5 * H * 4 : 0);
    static constexpr int OFF_TPTR = OFF_BAR + NBAR * 8;
    // no alignment slack: the third ring stage of H = 128 leaves 328 bytes.  Dynamic shared memory starts 1024-aligned when a
    // kernel has no static shared memory; the kernel checks it and traps otherwise
    static constexpr int SMEM = OFF_TPTR + 16;
    static_assert(4 * H <= GCOLS, "TMEM budget: accumulator (2H) + x1 (H) + A operand (H) columns per group");
    static_assert(SMEM <= 227 * 1024, "persistent tile set does not fit shared memory");
};

template <int H, int CS, int LC, bool STAGE>
__global__ void __launch_bounds__(PCfg<H, CS>::NTHREADS, 1)
sashimi_mix_umma_pers_kernel(MixArgs a, int B, int stagger_ns, int rev, const __grid_constant__ CUtensorMap tm_x,
                             const __grid_constant__ CUtensorMap tm_g) {
    using P = PCfg<H, CS>;
    using C = UCfg<H, CS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw;
    if (smem_u32(smem_raw) & 1023u) __trap();                   // SW128 operands and TMA boxes need the 1024-byte alignment
    uint8_t *wbuf = sm + P::OFF_W;
    float *bias_s = reinterpret_cast<float *>(sm + P::OFF_BIAS);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + P::OFF_BAR);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(sm + P::OFF_TPTR);
    uint64_t *wfull = bars + P::NG * P::NBAR_G, *wempty = wfull + P::NS;

    pdl_trigger();          // the next kernel's CTAs may be scheduled as SMs free up (they block in their own pdl_wait)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int l = LC ? LC : a.l, ntx = ceil_div(l, UM_TT), ntiles = B * ntx;
    const int stride = gridDim.x * P::NG;
    long long *trace = a.trace ? a.trace + (size_t)blockIdx.x * 16 : nullptr;

    if (tid == 0) {
        for (int g = 0; g < P::NG; ++g) {
            uint64_t *gb = bars + g * P::NBAR_G;
            mbar_init(gb + 0, P::EPI);                                    // g_ready
            mbar_init(gb + 1, P::EPI);                                    // z_ready
            for (int i = 0; i < C::KC3; ++i) mbar_init(gb + 2 + i, P::EPI);      // every thread contributes to every K chunk
            for (int i = 0; i < 2 * C::NC1 + 1; ++i) mbar_init(gb + 2 + C::KC3 + i, 1);   // acc1[], acc2[], acc3
            uint64_t *sb = gb + P::NBAR_G - 4;
            mbar_init(sb + 0, 1);                                         // xfull (transaction bytes)
            mbar_init(sb + 1, P::EPI);                                    // xempty
            mbar_init(sb + 2, 1);                                         // gfull
            mbar_init(sb + 3, P::EPI);                                    // gempty
        }
        for (int i = 0; i < P::NS; ++i) {
            mbar_init(wfull + i, 1);
            mbar_init(wempty + i, 1);
        }
        fence_mbar_init();
    }
    // biases: bo' (2H; per 128: 64 value | 64 gate pre-multiplied by -log2 e) b1 (F) b2 (H): in shared memory next to resident
    // weights; with streamed weights the space goes to a third ring stage and the biases come through L1
    if (P::BIAS_SMEM)
        for (int i = tid; i < 5 * H; i += P::NTHREADS) bias_s[i] = a.bimg[i];
    constexpr int EPI_WARPS = P::NG * P::GW;
    if (warp == EPI_WARPS) tmem_alloc(tptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem0 = *tptr;
    // Everything above (and the resident weight copies below) is independent of the previous kernel's output; from here on
    // every warp touches activations.  The weight producer of the resident form waits after its copies are in flight.
    if (!(P::RESIDENT && warp == EPI_WARPS + P::NG)) pdl_wait();

    if (STAGE && tid == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_g);
    }
    // ---- input staging producer: stream 2g = x tiles of group g, 2g + 1 = its g tiles; transfer k of a stream is the
    // group's k-th tile and is issued as soon as transfer k - 1 has been taken out of the buffer
    auto stage_producer = [&]() {
        constexpr int NSTR = 2 * P::NG;
        int k[NSTR], total[P::NG];
        bool more = false;
#pragma unroll
        for (int g = 0; g < P::NG; ++g) {
            const int first = blockIdx.x * P::NG + g;
            k[2 * g] = k[2 * g + 1] = 0;
            total[g] = first < ntiles ? (ntiles - 1 - first) / stride + 1 : 0;
            more |= total[g] > 0;
        }
        while (more) {
            more = false;
#pragma unroll
            for (int sidx = 0; sidx < NSTR; ++sidx) {
                const int g = sidx >> 1, isg = sidx & 1;
                if (k[sidx] >= total[g]) continue;
                more = true;
                uint64_t *sb = bars + g * P::NBAR_G + P::NBAR_G - 4 + 2 * isg;       // full, empty of this stream
                if (k[sidx] > 0) {
                    const int ok = lane == 0 ? (int)mbar_try(sb + 1, (uint32_t)((k[sidx] - 1) & 1)) : 0;
                    if (!__shfl_sync(0xffffffffu, ok, 0)) continue;
                }
                const int tile_ = blockIdx.x * P::NG + g + k[sidx] * stride;
                const int pt_ = rev ? ntiles - 1 - tile_ : tile_;
                const int b_ = pt_ / ntx, t0_ = (pt_ - b_ * ntx) * UM_TT;
                if (lane == 0) {
                    // one box = all H channel rows x 128 steps of clip b_ (columns past l arrive as zeros and count as bytes)
                    mbar_arrive_expect_tx(sb, P::STG);
                    tma_load_2d(sm + P::OFF_STG + (2 * g + isg) * P::STG, isg ? &tm_g : &tm_x, t0_, b_ * H, sb);
                }
                ++k[sidx];
            }
        }
    };

    if (STAGE && !P::RESIDENT && warp == EPI_WARPS + P::NG + 1) {
        stage_producer();
    } else if (warp == EPI_WARPS + P::NG) {
        // ================= weight producer =====================================================
        if (P::RESIDENT) {
            if (lane == 0)
                for (int i = 0; i < C::NSTG; ++i) {
                    mbar_arrive_expect_tx(wfull + i, UM_STAGE);
                    bulk_g2s(wbuf + (size_t)i * UM_STAGE, a.Wimg + (size_t)i * UM_STAGE, UM_STAGE, wfull + i);
                }
            __syncwarp();
            pdl_wait();
            if (STAGE) stage_producer();
        } else {
            // streamed weights: all lanes run the loop (uniform addresses), one elected lane issues the copy
            int cnt = 0;
            for (int tile = blockIdx.x * P::NG; tile < ntiles; tile += stride)
                for (int i = 0; i < C::NSTG; ++i, ++cnt) {
                    const int s = cnt % P::NS;
                    mbar_wait(wempty + s, ((cnt / P::NS) & 1) ^ 1);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(wfull + s, UM_STAGE);
                        bulk_g2s(wbuf + (size_t)s * UM_STAGE, a.Wimg + (size_t)i * UM_STAGE, UM_STAGE, wfull + s);
                    }
                    __syncwarp();
                }
        }
    } else if (warp >= EPI_WARPS + P::NG) {
        // (staging warp of an instantiation without staging: nothing to do)
    } else if (warp >= EPI_WARPS) {
        // ================= MMA issuer of group grp (all 32 lanes run the loops; one elected lane issues) ==========
        const int grp = warp - EPI_WARPS;
        {
            uint64_t *gb = bars + grp * P::NBAR_G;
            uint64_t *g_ready = gb, *z_ready = gb + 1, *hid_ready = gb + 2, *acc1_ready = hid_ready + C::KC3,
                     *acc2_ready = acc1_ready + C::NC1, *acc3_ready = acc2_ready + C::NC1;
            const uint32_t w0 = smem_u32(wbuf);
            const uint32_t tmem = tmem0 + grp * P::GCOLS;
            // one [128 x NR] x K = 64 block, A operand in TMEM ([8 hi | 8 lo] columns per K = 16 step): 3 split terms x 4 k-steps
            auto issue_block = [&](uint32_t d, uint32_t a_tmem, uint32_t bbase, int NR, bool acc0) {
                const uint32_t idesc = idesc_bf16(128, NR);
                if (elect_one()) {
#pragma unroll
                    for (int term = 0; term < 3; ++term) {
                        const uint32_t ao = a_tmem + (term == 1 ? 8 : 0);
                        const uint32_t bo = bbase + (term == 2 ? NR * 128 : 0);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            mma_bf16_ts(d, ao + ks * 16, smem_desc_sw128(bo + ks * 32), idesc, (acc0 || term > 0 || ks > 0) ? 1u : 0u);
                    }
                }
                __syncwarp();
            };
            auto commit = [&](uint64_t *bar) {
                if (elect_one()) mma_commit(bar);
                __syncwarp();
            };
            int cnt = 0;     // weight stages consumed so far (ring position when streaming)
            uint32_t ph = 0;
            // debug trace (a.trace): cycles the issuer of group 0 spends waiting, second tile: [10] g/z operands, [11] weights, [12] hidden chunks, [13] whole tile
            const bool mtr = trace && grp == 0;
            long long w_op = 0, w_wt = 0, w_hid = 0, t_tile = 0;
            int mit = 0;
#define MW(acc, stmt) do { if (mtr && mit == 1) { const long long c0_ = clock64(); stmt; acc += clock64() - c0_; } else { stmt; } } while (0)
#pragma unroll 1
            for (int tile = blockIdx.x * P::NG + grp; tile < ntiles; tile += stride, ph ^= 1, ++mit) {
                if (mtr && mit == 1) t_tile = clock64();
                int i = 0;   // stage index inside the tile
#pragma unroll 1
                for (int gemm = 0; gemm < 2; ++gemm) {
                    MW(w_op, mbar_wait(gemm == 0 ? g_ready : z_ready, ph));
                    tc_fence_after();
#pragma unroll 1
                    for (int nc = 0; nc < C::NC1; ++nc) {
#pragma unroll 1
                        for (int kc = 0; kc < C::KC1; ++kc, ++i, ++cnt) {
                            const int s = P::RESIDENT ? i : cnt % P::NS;
                            MW(w_wt, mbar_wait(wfull + s, P::RESIDENT ? 0u : (uint32_t)((cnt / P::NS) & 1)));
                            tc_fence_after();
                            issue_block(tmem + nc * 128, tmem + P::AOP + 64 * kc, w0 + s * UM_STAGE, 128, kc > 0);
                            if (!P::RESIDENT) commit(wempty + s);
                        }
                        commit((gemm == 0 ? acc1_ready : acc2_ready) + nc);
                    }
                }
#pragma unroll 1
                for (int kc = 0; kc < C::KC3; ++kc) {
                    MW(w_hid, mbar_wait(hid_ready + kc, ph));
                    tc_fence_after();
#pragma unroll 1
                    for (int nc = 0; nc < C::NC3; ++nc) {
                        const int j = kc * C::NC3 + nc;
                        const int s = P::RESIDENT ? i : cnt % P::NS;
                        if (j % C::BPS3 == 0) {
                            MW(w_wt, mbar_wait(wfull + s, P::RESIDENT ? 0u : (uint32_t)((cnt / P::NS) & 1)));
                            tc_fence_after();
                        }
                        issue_block(tmem + P::R3 + nc * 128, tmem + 64 * kc, w0 + s * UM_STAGE + (j % C::BPS3) * C::NR3 * 256, C::NR3, true);
                        if (j % C::BPS3 == C::BPS3 - 1) {
                            if (!P::RESIDENT) commit(wempty + s);
                            ++i;
                            ++cnt;
                        }
                    }
                }
                commit(acc3_ready);
                if (mtr && mit == 1 && lane == 0) {
                    trace[10] = w_op;
                    trace[11] = w_wt;
                    trace[12] = w_hid;
                    trace[13] = clock64() - t_tile;
                }
            }
#undef MW
        }
    } else {
        // ================= epilogue threads of group grp =========================================
        const int grp = warp / P::GW, cg = (warp % P::GW) >> 2, q = warp & 3;
        const int r = 32 * q + lane;
        uint64_t *gb = bars + grp * P::NBAR_G;
        uint64_t *g_ready = gb, *z_ready = gb + 1, *hid_ready = gb + 2, *acc1_ready = hid_ready + C::KC3,
                 *acc2_ready = acc1_ready + C::NC1, *acc3_ready = acc2_ready + C::NC1;
        uint64_t *xfull = gb + P::NBAR_G - 4, *xempty = xfull + 1, *gfull = xfull + 2, *gempty = xfull + 3;
        const uint32_t tl = tmem0 + grp * P::GCOLS + ((uint32_t)(32 * q) << 16);
        const float *bsrc = P::BIAS_SMEM ? bias_s : a.bimg;
        const float *bo_s = bsrc, *b1_s = bsrc + 2 * H, *b2_s = bsrc + 4 * H;
        constexpr int PER = H / CS, PP = 64 / CS;
        const int etid = tid - grp * P::EPI;         // thread index inside the group
        const bool tracer = trace && grp == 0 && etid == 0;
#define PT(slot) do { if (tracer && it == 1) trace[slot] = clock64(); } while (0)

        // statistics exchange between the CS column groups of a time step (partners share the TMEM lane) through two
        // columns of this thread's OWN range that it has already consumed (col + cg * stride: a partner may still be
        // reading its own columns of the same region).  `release`: a second barrier after the partners' values have been
        // read, needed when the next writer of those columns is another epilogue thread rather than an MMA behind an mbarrier.
        auto exchange = [&](uint32_t col, int stride, bool release, float &mean, float &M2) {
            if (CS > 1) {
                jitter_sleep(a.jitter, col + 1);
                tmem_st2(tl + col + cg * stride, mean, M2);
                tmem_wait_st();
                tc_fence_before();
                asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(P::EPI) : "memory");
                tc_fence_after();
                jitter_sleep(a.jitter, col + 2);
                float pm[CS], p2[CS];
#pragma unroll
                for (int c = 0; c < CS; ++c) tmem_ld2(tl + col + c * stride, pm[c], p2[c]);
                tmem_wait_ld();
                if (release) {
                    tc_fence_before();
                    asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(P::EPI) : "memory");
                    tc_fence_after();
                }
                float ms = 0.f;
#pragma unroll
                for (int c = 0; c < CS; ++c) ms += pm[c];
                const float mt = ms * (1.0f / CS);
                float m2 = 0.f;
#pragma unroll
                for (int c = 0; c < CS; ++c) {
                    const float d = pm[c] - mt;
                    m2 += p2[c] + d * d * (float)(H / CS);
                }
                mean = mt;
                M2 = m2;
            }
        };
        auto tile_row = [&](int tile_, bool &valid_) -> size_t {
            const int pt_ = rev ? ntiles - 1 - tile_ : tile_;           // rev: walk the batch from its end (see mix_umma_launch)
            const int b_ = pt_ / ntx, t_ = (pt_ - b_ * ntx) * UM_TT + r;
            valid_ = tile_ < ntiles && t_ < l;
            return (size_t)b_ * H * l + (valid_ ? t_ : 0);
        };
        // this thread's channels: g / skip / out use h = cg*PER + i; x uses the GLU pairing h = nc*64 + cg*PP + i
        // (the columns E1 produces).  par = parity of the stream's transfer (the group's tile counter & 1).
        const float *xb = reinterpret_cast<const float *>(sm + P::OFF_STG + (2 * grp) * P::STG) + r;
        const float *gbuf = reinterpret_cast<const float *>(sm + P::OFF_STG + (2 * grp + 1) * P::STG) + r;
        auto take_x = [&](float (&xin)[PER], int tile_, uint32_t par) {
            bool v_;
            const size_t row = tile_row(tile_, v_);
            if (STAGE) {
                mbar_wait(xfull, par);
#pragma unroll
                for (int nc = 0; nc < C::NC1; ++nc)
#pragma unroll
                    for (int i = 0; i < PP; ++i) xin[nc * PP + i] = v_ ? xb[(nc * 64 + cg * PP + i) * UM_TT] : 0.f;
                mbar_arrive(xempty);
            } else {
                const float *xp = a.x + row + (size_t)cg * PP * l;
#pragma unroll
                for (int nc = 0; nc < C::NC1; ++nc)
#pragma unroll
                    for (int i = 0; i < PP; ++i) xin[nc * PP + i] = v_ ? __ldg(chan<LC>(xp, nc * 64 + i, l)) : 0.f;
            }
        };
        auto load_g = [&](float (&gin)[PER], int tile_) {       // !STAGE: request only
            bool v_;
            const float *gp = a.g + tile_row(tile_, v_) + (size_t)cg * PER * l;
#pragma unroll
            for (int i = 0; i < PER; ++i) gin[i] = v_ ? __ldg(chan<LC>(gp, i, l)) : 0.f;
        };

        // 16 channels starting at h0 -> the A operand columns of their K = 16 step: [8 packed hi | 8 packed lo]
        auto store_aop = [&](const float *v, int h0) {
            uint4 hi0, lo0, hi1, lo1;
            split8p(v, hi0, lo0);
            split8p(v + 8, hi1, lo1);
            tmem_st8(tl + P::AOP + h0, hi0, hi1);
            tmem_st8(tl + P::AOP + h0 + 8, lo0, lo1);
        };
        // staged g tile -> A operand of G1, 16 channels at a time (keeps the live registers low next to the skip prefetch)
        auto take_store_g = [&](int tile_, uint32_t par) {
            bool v_;
            tile_row(tile_, v_);
            mbar_wait(gfull, par);
#pragma unroll
            for (int c = 0; c < PER / 16; ++c) {
                float gin[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) gin[i] = v_ ? gbuf[(cg * PER + 16 * c + i) * UM_TT] : 0.f;
                store_aop(gin, cg * PER + 16 * c);
            }
            mbar_arrive(gempty);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(g_ready);
        };
        auto store_g = [&](const float (&gin)[PER]) {
#pragma unroll
            for (int c = 0; c < PER / 16; ++c) store_aop(gin + 16 * c, cg * PER + 16 * c);
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(g_ready);
        };

        int tile = blockIdx.x * P::NG + grp;
        float xin[PER];                               // x of this tile, then x1 (E1 -> LN2)
        if (tile < ntiles) {
            if (grp > 0 && stagger_ns > 0) __nanosleep(stagger_ns);     // start the groups out of phase
            if (STAGE) {
                take_store_g(tile, 0);
            } else {
                float gin[PER];
                take_x(xin, tile, 0);
                load_g(gin, tile);
                store_g(gin);
            }
        }
        uint32_t ph = 0;
        int it = 0;
#pragma unroll 1
        for (; tile < ntiles; tile += stride, ph ^= 1, ++it) {
            const int ptile = rev ? ntiles - 1 - tile : tile;
            const int b = ptile / ntx, t = (ptile - b * ntx) * UM_TT + r;
            const bool valid = t < l;
            const size_t brow = (size_t)b * H * l + (valid ? t : 0);
            const int nt = tile + stride;
            jitter_sleep(a.jitter, 16 * it + 3);
            if (STAGE) take_x(xin, tile, (uint32_t)(it & 1));      // staged a tile ago; x is not live across E2 / E3
            PT(0);
            // ---- E1: GLU + residual -> x1 (TMEM R3 for G3's accumulation, registers for LN2), LN2 statistics
            float mean, M2;
            {
                s2::V2 sd(0.f), sq(0.f);
                float piv = 0.f;
#pragma unroll
                for (int nc = 0; nc < C::NC1; ++nc) {
                    mbar_wait(acc1_ready + nc, ph);
                    tc_fence_after();
                    if (nc == 0) PT(1);
#pragma unroll
                    for (int sc = 0; sc < PP / 16; ++sc) {
                        const int p0 = cg * PP + sc * 16, h0 = nc * 64 + p0;
                        float av[16], gv[16];
                        tmem_ld16(tl + nc * 128 + p0, av);
                        tmem_ld16(tl + nc * 128 + 64 + p0, gv);
                        tmem_wait_ld();
                        const float *ba = bo_s + nc * 128 + p0;
                        float(&xv)[16] = *reinterpret_cast<float(*)[16]>(xin + nc * PP + sc * 16);
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            // two channels per packed instruction: sigmoid(gate) = 1 / (1 + 2^(-log2e gate - log2e b))
                            const s2::V2 sg = s2::sigmoid_scaled2(s2::V2(gv[i], gv[i + 1]), s2::V2(ba[64 + i], ba[64 + i + 1]));
                            const s2::V2 y = s2::fma(s2::V2(av[i], av[i + 1]) + s2::V2(ba[i], ba[i + 1]), sg, s2::V2(xv[i], xv[i + 1]));
                            xv[i] = y.v.x;
                            xv[i + 1] = y.v.y;
                        }
                        if (a.cond && valid) {
                            const float *cp = a.cond + (size_t)(a.cond_stride_b ? b : 0) * H * l + t + (size_t)h0 * l;
#pragma unroll
                            for (int i = 0; i < 16; ++i) xv[i] += __ldg(chan<LC>(cp, i, l));
                        }
                        if (nc == 0 && sc == 0) piv = xv[0];
                        stat_acc16(xv, s2::V2(piv), sd, sq);
                        tmem_st16(tl + P::R3 + h0, xv);
                    }
                }
                stat_finish(sd, sq, piv, PER, mean, M2);
                tmem_wait_st();
            }
            PT(2);
            // through accumulator columns this thread has consumed (its first value columns); their next writer is G2, issued
            // only after every thread has arrived on z_ready, i.e. after its exchange reads
            exchange(0, PP, false, mean, M2);
            // ---- z = LN2(x1), split -> AOP: the A operand of G2 (G1 has completed: acc1 was awaited)
            {
                const float rstd = valid ? rsqrtf(M2 * (1.0f / H)) : 0.f;
                const float sc_a = a.ln2_s * rstd;
                const s2::V2 sc2(sc_a), sh2(sc_a * (a.ln2_m - mean));
#pragma unroll
                for (int k = 0; k < PER / 16; ++k) {
                    const int nc = k / (PP / 16), sc = k % (PP / 16);
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const s2::V2 z = s2::fma(s2::V2(xin[k * 16 + i], xin[k * 16 + i + 1]), sc2, sh2);
                        v[i] = z.v.x;
                        v[i + 1] = z.v.y;
                    }
                    store_aop(v, nc * 64 + cg * PP + sc * 16);
                }
                tmem_wait_st();
                tc_fence_before();
                mbar_arrive(z_ready);
            }
            jitter_sleep(a.jitter, 16 * it + 4);
            PT(3);
            // ---- E2: hidden = gelu(W1 z + b1), split, written in place over the accumulator columns as the A operand of G3.
            //      The TMEM load of the next 16 columns is issued before the arithmetic of the current ones.
            {
                constexpr int PERF = 128 / CS, NSC = PERF / 16, SPK = PERF / 32;      // SPK: 16-column steps per K chunk and thread
                // a thread's columns are spread over both 64-wide K chunks of an N chunk, so the first chunk is complete -
                // and its part of G3 issued - when E2 is half way through the second
                auto e2col = [&](int sc) { return (sc / SPK) * 64 + cg * (PERF / 2) + (sc % SPK) * 16; };
#pragma unroll 1
                for (int nc = 0; nc < C::NC1; ++nc) {
                    mbar_wait(acc2_ready + nc, ph);
                    tc_fence_after();
                    if (nc == 0) PT(4);
                    float vbuf[2][16];
                    tmem_ld16(tl + nc * 128 + e2col(0), vbuf[0]);
#pragma unroll
                    for (int sc = 0; sc < NSC; ++sc) {
                        float(&v)[16] = vbuf[sc & 1];
                        const int f0 = nc * 128 + e2col(sc);
                        tmem_wait_ld();
                        if (sc + 1 < NSC) tmem_ld16(tl + nc * 128 + e2col(sc + 1), vbuf[(sc + 1) & 1]);
                        const float *bb = b1_s + f0;
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const s2::V2 rr = s2::gelu_fast2(s2::V2(v[i], v[i + 1]) + s2::V2(bb[i], bb[i + 1]));
                            v[i] = rr.v.x;
                            v[i + 1] = rr.v.y;
                        }
                        const int kc = f0 >> 6;
                        uint4 h0, l0, h1, l1;
                        split8p(v, h0, l0);
                        split8p(v + 8, h1, l1);
                        tmem_st8(tl + f0, h0, h1);
                        tmem_st8(tl + f0 + 8, l0, l1);
                        if (sc % SPK == SPK - 1) {
                            tmem_wait_st();
                            tc_fence_before();
                            mbar_arrive(hid_ready + kc);
                        }
                    }
                }
            }
            PT(5);
            // ---- once G3 has been awaited the A operand columns are free: the next tile's g goes in and G1(next) runs under E3
            // (the UNet skip is loaded 16 channels at a time inside E3: requesting the whole thread's share before the wait for G3
            //  hides its latency but costs ~23 spilled registers under the 96-register cap of this 608-thread CTA - measured on one
            //  box with tools/ab_lib.py: 19.44 clips/s with the prefetch, 20.10 without)
            const float *sp = a.skip ? a.skip + brow + (size_t)cg * PER * l : nullptr;
            if (STAGE) {
                mbar_wait(acc3_ready, ph);
                tc_fence_after();
                PT(6);
                if (nt < ntiles) take_store_g(nt, (uint32_t)((it + 1) & 1));
            } else {
                float gin[PER];
                load_g(gin, nt);
                mbar_wait(acc3_ready, ph);
                tc_fence_after();
                PT(6);
                if (nt < ntiles) store_g(gin);
                take_x(xin, nt, 0);               // requested here, lands during E3
            }
            jitter_sleep(a.jitter, 16 * it + 5);
            PT(7);
            // ---- E3: x2 = acc3 (= x1 + W2 hidden) + b2 (+skip); store; statistics for the next norm
            {
                s2::V2 sd(0.f), sq(0.f);
                float piv = 0.f;
                float *op = a.out + brow + (size_t)cg * PER * l;
#pragma unroll
                for (int sc = 0; sc < PER / 16; ++sc) {
                    const int h0 = cg * PER + sc * 16;
                    float v[16];
                    tmem_ld16(tl + P::R3 + h0, v);
                    const float *bb = b2_s + h0;
                    if (sp) {
                        float sk[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) sk[i] = valid ? __ldg(chan<LC>(sp, sc * 16 + i, l)) : 0.f;
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const s2::V2 y = (s2::V2(v[i], v[i + 1]) + s2::V2(bb[i], bb[i + 1])) + s2::V2(sk[i], sk[i + 1]);
                            v[i] = y.v.x;
                            v[i + 1] = y.v.y;
                        }
                    } else {
                        tmem_wait_ld();
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            const s2::V2 y = s2::V2(v[i], v[i + 1]) + s2::V2(bb[i], bb[i + 1]);
                            v[i] = y.v.x;
                            v[i + 1] = y.v.y;
                        }
                    }
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) *chan<LC>(op, sc * 16 + i, l) = v[i];
                    }
                    if (sc == 0) piv = v[0];
                    stat_acc16(v, s2::V2(piv), sd, sq);
                }
                stat_finish(sd, sq, piv, PER, mean, M2);
                // through x2 columns this thread has read; their next writer is a partner's E1 of the next tile: second barrier
                exchange(P::R3, PER, true, mean, M2);
                if (cg == 0 && valid)
                    *reinterpret_cast<float2 *>(a.stats_out + ((size_t)b * l + t) * 2) = make_float2(mean, rsqrtf(M2 * (1.0f / H)));
            }
            PT(8);
        }
#undef PT
    }
    tc_fence_before();
    __syncthreads();
    if (warp == EPI_WARPS) {
        tc_fence_after();
        tmem_dealloc(tmem0, 512);
    }
}

// =========================================================================================
// H = 256 (F = 512): the operands of one 128-step tile no longer fit the per-tile scheme
// (z alone is 128 KB, the weights 1.5 MB).  Same orientation and epilogues, but
//   * G1/G2 run in N chunks of 128 columns through TWO accumulator regions R1a / R1b used alternately, so the
//     epilogue of chunk n (GLU, or GELU + split) runs under the MMAs of chunk n + 1,
//   * the hidden activations never touch shared memory: E2 writes them (split bf16, per 16 f-columns 8 packed
//     hi columns then 8 packed lo columns) in place over the accumulator columns it has just read, and G3
//     takes its A operand from there (tcgen05.mma TS form).  The tensor pipe executes in issue order, so the
//     G2 chunk that next overwrites a region is simply issued after the G3 blocks that read it,
//   * biases come through L1 (uniform __ldg), the column-group statistics exchange through idle R1a columns,
//     so shared memory holds only g / z (4 slots) and a 3-stage weight ring.
// TMEM: [0,256) x1 / acc3 | [256,384) R1a | [384,512) R1b.
// Weight image = 48 stages of 32 KB in consumption order:
//   G1 (nc, kc) x16 | G2(0) x4 | for nc = 0..3: { G2(nc+1) x4 (nc < 3) | G3(2nc) x2 | G3(2nc+1) x2 }.
// =========================================================================================
struct U256 {
    static constexpr int H = 256, F = 512;
    static constexpr int NC1 = 4, KC1 = 4, KC3 = 8, NB3 = 2, NSTG = 48, NS = 3, NSLOT = 4;
    static constexpr int R3 = 0, R1 = 256;
    __host__ __device__ static constexpr int r1(int nc) { return R1 + 128 * (nc & 1); }
    static constexpr int OFF_SLOT = 0;
    static constexpr int OFF_RING = NSLOT * UM_SLOT;
    static constexpr int OFF_BAR = OFF_RING + NS * UM_STAGE;
    static constexpr int NBAR = 2 * NS + 2 + 2 + KC3 + 2 * NC1 + 1;
    static constexpr int OFF_TPTR = OFF_BAR + NBAR * 8;
    static constexpr int SMEM = OFF_TPTR + 16 + 1024;
    static constexpr size_t IMG_BYTES = (size_t)NSTG * UM_STAGE;
    static_assert(SMEM <= 227 * 1024, "H=256 tile does not fit shared memory");
    // stage i of the image -> (gemm, first output row of the 128-row block before permutation, K chunk)
    __host__ __device__ static void stage_info(int i, int &gemm, int &nblk, int &kc) {
        if (i < 16) { gemm = 0; nblk = i / 4; kc = i % 4; return; }
        if (i < 20) { gemm = 1; nblk = 0; kc = i - 16; return; }
        const int j = i - 20;
        if (j < 24) {
            const int nc = j / 8, r = j % 8;
            if (r < 4) { gemm = 1; nblk = nc + 1; kc = r; }
            else { gemm = 2; kc = 2 * nc + (r - 4) / 2; nblk = (r - 4) % 2; }
        } else {
            const int r = j - 24;
            gemm = 2; kc = 6 + r / 2; nblk = r % 2;
        }
    }
};

// CS = epilogue threads per time step (2 or 4): each owns 64 / CS GLU pairs of every 64-channel block
template <int CS>
__global__ void __launch_bounds__(128 * CS + 64, 1)
sashimi_mix_umma256_kernel(MixArgs a) {
    using C = U256;
    constexpr int H = C::H, EPI = 128 * CS, PP = 64 / CS, PER = H / CS, PERF = 128 / CS;
    constexpr int PW = 4 * CS, MW = 4 * CS + 1;             // producer / MMA issuer warps
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *slots = sm + C::OFF_SLOT, *ring = sm + C::OFF_RING;
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + C::OFF_BAR);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(sm + C::OFF_TPTR);
    uint64_t *wfull = bars, *wempty = wfull + C::NS, *g_ready = wempty + C::NS, *z_ready = g_ready + 1, *r1free = z_ready + 1,
             *hid_ready = r1free + 2, *acc1_ready = hid_ready + C::KC3, *acc2_ready = acc1_ready + C::NC1,
             *acc3_ready = acc2_ready + C::NC1;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = a.rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y, t0 = (a.rev ? gridDim.x - 1 - blockIdx.x : blockIdx.x) * UM_TT, l = a.l;
    if (tid == 0) {
        for (int i = 0; i < C::NS; ++i) {
            mbar_init(wfull + i, 1);
            mbar_init(wempty + i, 1);
        }
        mbar_init(g_ready, EPI);
        mbar_init(z_ready, EPI);
        mbar_init(r1free, EPI);
        mbar_init(r1free + 1, EPI);
        for (int i = 0; i < C::KC3; ++i) mbar_init(hid_ready + i, 64 * CS);
        for (int i = 0; i < 2 * C::NC1 + 1; ++i) mbar_init(acc1_ready + i, 1);
        fence_mbar_init();
    }
    pdl_trigger();
    if (warp == MW) tmem_alloc(tptr, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    pdl_wait();             // set-up done; everything below reads or writes activations

    if (warp == PW) {
        // ================= weight producer (all lanes run the loop, one elected lane issues the copy) ==========
        for (int i = 0; i < C::NSTG; ++i) {
            const int s = i % C::NS, n = i / C::NS;
            mbar_wait(wempty + s, (n & 1) ^ 1);
            if (elect_one()) {
                mbar_arrive_expect_tx(wfull + s, UM_STAGE);
                bulk_g2s(ring + (size_t)s * UM_STAGE, a.Wimg + (size_t)i * UM_STAGE, UM_STAGE, wfull + s);
            }
            __syncwarp();
        }
    } else if (warp == MW) {
        // ================= MMA issuer ==========================================================
        {   // all 32 lanes run the loops; the *_w forms elect the issuing lane
            const uint32_t slot0 = smem_u32(slots), ring0 = smem_u32(ring);
            constexpr uint32_t idesc = idesc_bf16(128, 128);
            int i = 0;                                    // weight stage counter
            auto next_stage = [&]() {
                const int s = i % C::NS;
                mbar_wait(wfull + s, (i / C::NS) & 1);
                tc_fence_after();
                return ring0 + s * UM_STAGE;
            };
            auto done_stage = [&]() {
                mma_commit_w(wempty + (i % C::NS));
                ++i;
            };
            // D[R1(nc)] = A[slots, all K] x stage blocks (SS form)
            auto gemm_ss = [&](int nc, uint64_t *ready) {
#pragma unroll 1
                for (int kc = 0; kc < C::KC1; ++kc) {
                    const uint32_t bbase = next_stage(), abase = slot0 + kc * UM_SLOT;
                    if (elect_one()) {      // one election per block of 12 MMAs
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const uint32_t ao = abase + (term == 1 ? UM_SLOT / 2 : 0), bo = bbase + (term == 2 ? UM_STAGE / 2 : 0);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_bf16_ss(tmem + C::r1(nc), smem_desc_sw128(ao + ks * 32), smem_desc_sw128(bo + ks * 32), idesc,
                                            (kc > 0 || term > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    __syncwarp();
                    done_stage();
                }
                mma_commit_w(ready);
            };
            // R3[nb] += hidden K chunk kc (in place in R1(kc / 2), columns 64 (kc & 1) ..) x stage (TS form: A from TMEM)
            auto gemm_ts = [&](int kc) {
#pragma unroll 1
                for (int nb = 0; nb < C::NB3; ++nb) {
                    const uint32_t bbase = next_stage(), abase = tmem + C::r1(kc >> 1) + 64 * (kc & 1);
                    if (elect_one()) {      // one election per block of 12 MMAs
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const uint32_t ao = abase + (term == 1 ? 8 : 0), bo = bbase + (term == 2 ? UM_STAGE / 2 : 0);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_bf16_ts(tmem + C::R3 + nb * 128, ao + ks * 16, smem_desc_sw128(bo + ks * 32), idesc, 1u);
                        }
                    }
                    __syncwarp();
                    done_stage();
                }
            };
            mbar_wait(g_ready, 0);
            tc_fence_after();
#pragma unroll 1
            for (int nc = 0; nc < C::NC1; ++nc) {
                if (nc >= 2) {                            // E1 has drained this accumulator region (chunk nc - 2)
                    mbar_wait(r1free + (nc & 1), 0);
                    tc_fence_after();
                }
                gemm_ss(nc, acc1_ready + nc);
            }
            mbar_wait(z_ready, 0);                        // (every E1 chunk has been read: both regions are free)
            tc_fence_after();
            gemm_ss(0, acc2_ready + 0);
#pragma unroll 1
            for (int nc = 0; nc < C::NC1; ++nc) {
                // region (nc + 1) & 1 held hidden chunk nc - 1: its G3 blocks were issued in the previous iteration
                if (nc + 1 < C::NC1) gemm_ss(nc + 1, acc2_ready + nc + 1);
#pragma unroll 1
                for (int cg = 0; cg < 2; ++cg) {
                    mbar_wait(hid_ready + 2 * nc + cg, 0);
                    tc_fence_after();
                    gemm_ts(2 * nc + cg);
                }
            }
            mma_commit_w(acc3_ready);
        }
    } else {
        // ================= epilogue threads: one time step each, column group cg ================
        const int q = warp & 3, cg = warp >> 2;
        const int r = 32 * q + lane, t = t0 + r;
        const bool valid = t < l;
        const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
        const size_t brow = (size_t)b * H * l + (valid ? t : 0);
        const float *bo_g = a.bimg, *b1_g = a.bimg + 2 * H, *b2_g = a.bimg + 4 * H;
        // this thread's channels: h = nc * 64 + cg * PP + i, nc < 4, i < PP (the same set in every phase)
        constexpr int NCR = 64 / PP, NRD = PER / 64;        // 64-channel blocks per round of 64 loads, rounds

        // ---- g -> slots (A operand of G1), rounds of 64 loads in flight
#pragma unroll 1
        for (int rd = 0; rd < NRD; ++rd) {
            float v[64];
#pragma unroll
            for (int kk = 0; kk < NCR; ++kk) {
                const float *gp = a.g + brow + (size_t)((NCR * rd + kk) * 64 + cg * PP) * l;
#pragma unroll
                for (int i = 0; i < PP; ++i) v[kk * PP + i] = valid ? __ldg(chan<0>(gp, i, l)) : 0.f;
            }
#pragma unroll
            for (int kk = 0; kk < NCR; ++kk) {
                uint8_t *slot = slots + (NCR * rd + kk) * UM_SLOT;
#pragma unroll
                for (int c8 = 0; c8 < PP / 8; ++c8) {
                    uint4 hi, lo;
                    split8p(v + kk * PP + 8 * c8, hi, lo);
                    const uint32_t off = sw128_off(r, cg * (PP / 8) + c8);
                    *reinterpret_cast<uint4 *>(slot + off) = hi;
                    *reinterpret_cast<uint4 *>(slot + UM_SLOT / 2 + off) = lo;
                }
            }
        }
        fence_proxy_async_smem();
        mbar_arrive(g_ready);
        // ---- x -> TMEM R3 while G1 runs
#pragma unroll 1
        for (int rd = 0; rd < NRD; ++rd) {
            float v[64];
#pragma unroll
            for (int kk = 0; kk < NCR; ++kk) {
                const float *xp = a.x + brow + (size_t)((NCR * rd + kk) * 64 + cg * PP) * l;
#pragma unroll
                for (int i = 0; i < PP; ++i) v[kk * PP + i] = valid ? __ldg(chan<0>(xp, i, l)) : 0.f;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float w[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) w[i] = v[c * 16 + i];
                tmem_st16(tl + C::R3 + (NCR * rd + (c * 16) / PP) * 64 + cg * PP + (c * 16) % PP, w);
            }
        }
        tmem_wait_st();

        // through R1a columns of this thread's OWN value range (all of its E1 reads are done; a partner may still be reading
        // its own columns of R1a); the next writer of R1a is G2(0), issued after every thread has arrived on z_ready
        auto exchange = [&](float &mean, float &M2) {
            jitter_sleep(a.jitter, 1);
            tmem_st2(tl + C::R1 + cg * PP, mean, M2);
            tmem_wait_st();
            tc_fence_before();
            asm volatile("bar.sync 1, %0;" ::"n"(EPI) : "memory");
            tc_fence_after();
            jitter_sleep(a.jitter, 2);
            float pm[CS], p2[CS];
#pragma unroll
            for (int c = 0; c < CS; ++c) tmem_ld2(tl + C::R1 + c * PP, pm[c], p2[c]);
            tmem_wait_ld();
            float ms = 0.f;
#pragma unroll
            for (int c = 0; c < CS; ++c) ms += pm[c];
            const float mt = ms * (1.0f / CS);
            float m2 = 0.f;
#pragma unroll
            for (int c = 0; c < CS; ++c) {
                const float d = pm[c] - mt;
                m2 += p2[c] + d * d * (float)PER;
            }
            mean = mt;
            M2 = m2;
        };

        // ---- E1: GLU + residual -> x1 (R3), LN2 statistics
        float mean, M2;
        {
            s2::V2 sd(0.f), sq(0.f);
            float piv = 0.f;
#pragma unroll 1
            for (int nc = 0; nc < C::NC1; ++nc) {
                jitter_sleep(a.jitter, 4 + nc);
                mbar_wait(acc1_ready + nc, 0);
                tc_fence_after();
                const uint32_t r1 = tl + C::r1(nc);
#pragma unroll 1
                for (int sc = 0; sc < PP / 16; ++sc) {
                    const int p0 = cg * PP + sc * 16, h0 = nc * 64 + p0;
                    float xv[16], av[16], gv[16];
                    tmem_ld16(r1 + p0, av);
                    tmem_ld16(r1 + 64 + p0, gv);
                    tmem_ld16(tl + C::R3 + h0, xv);
                    tmem_wait_ld();
                    const float *ba = bo_g + nc * 128 + p0;
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        // two channels per packed instruction (FADD2 / FMUL2 / FFMA2)
                        const s2::V2 y = s2::fma(s2::V2(av[i], av[i + 1]) + s2::V2(__ldg(ba + i), __ldg(ba + i + 1)),
                                                 s2::sigmoid_scaled2(s2::V2(gv[i], gv[i + 1]), s2::V2(__ldg(ba + 64 + i), __ldg(ba + 64 + i + 1))),
                                                 s2::V2(xv[i], xv[i + 1]));
                        xv[i] = y.v.x;
                        xv[i + 1] = y.v.y;
                    }
                    if (a.cond && valid) {
                        const float *cp = a.cond + (size_t)(a.cond_stride_b ? b : 0) * H * l + t + (size_t)h0 * l;
#pragma unroll
                        for (int i = 0; i < 16; ++i) xv[i] += __ldg(chan<0>(cp, i, l));
                    }
                    if (nc == 0 && sc == 0) piv = xv[0];
                    stat_acc16(xv, s2::V2(piv), sd, sq);
                    tmem_st16(tl + C::R3 + h0, xv);
                }
                tc_fence_before();
                mbar_arrive(r1free + (nc & 1));         // this accumulator region is drained: chunk nc + 2 may be issued
            }
            stat_finish(sd, sq, piv, PER, mean, M2);
            tmem_wait_st();
        }
        exchange(mean, M2);
        // ---- z = LN2(x1) -> slots (A operand of G2)
        {
            const float rstd = valid ? rsqrtf(M2 * (1.0f / H)) : 0.f;
            const float sc_a = a.ln2_s * rstd;
            const s2::V2 sc2(sc_a), sh2(sc_a * (a.ln2_m - mean));
#pragma unroll 1
            for (int k = 0; k < PER / 16; ++k) {
                const int nc = k / (PP / 16), sc = k % (PP / 16);
                const int h0 = nc * 64 + cg * PP + sc * 16;
                float v[16];
                tmem_ld16(tl + C::R3 + h0, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const s2::V2 z = s2::fma(s2::V2(v[i], v[i + 1]), sc2, sh2);
                    v[i] = z.v.x;
                    v[i + 1] = z.v.y;
                }
                uint8_t *slot = slots + nc * UM_SLOT;
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    uint4 hi, lo;
                    split8p(v + 8 * hh, hi, lo);
                    const uint32_t off = sw128_off(r, ((h0 & 63) >> 3) + hh);
                    *reinterpret_cast<uint4 *>(slot + off) = hi;
                    *reinterpret_cast<uint4 *>(slot + UM_SLOT / 2 + off) = lo;
                }
            }
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(z_ready);
        }
        // ---- E2: hidden = gelu(W1 z + b1), split, in place over the accumulator columns (A operand of G3, K chunk 2 nc + cg)
#pragma unroll 1
        for (int nc = 0; nc < C::NC1; ++nc) {
            jitter_sleep(a.jitter, 8 + nc);
            mbar_wait(acc2_ready + nc, 0);
            tc_fence_after();
            const uint32_t r1 = tl + C::r1(nc);
#pragma unroll 1
            for (int sc = 0; sc < PERF / 16; ++sc) {
                const int col = cg * PERF + sc * 16, f0 = nc * 128 + col;
                float v[16];
                tmem_ld16(r1 + col, v);
                tmem_wait_ld();
                const float *bb = b1_g + f0;
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const s2::V2 rr = s2::gelu_fast2(s2::V2(v[i], v[i + 1]) + s2::V2(__ldg(bb + i), __ldg(bb + i + 1)));
                    v[i] = rr.v.x;
                    v[i + 1] = rr.v.y;
                }
                uint4 h0, l0, h1, l1;
                split8p(v, h0, l0);
                split8p(v + 8, h1, l1);
                tmem_st8(r1 + col, h0, h1);
                tmem_st8(r1 + col + 8, l0, l1);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(hid_ready + 2 * nc + (cg * PERF) / 64);
        }
        // ---- E3: x2 = acc3 (= x1 + W2 hidden) + b2 (+skip); store; statistics for the next norm
        {
            mbar_wait(acc3_ready, 0);
            tc_fence_after();
            s2::V2 sd(0.f), sq(0.f);
            float piv = 0.f;
            float *op = a.out + brow;
#pragma unroll 1
            for (int k = 0; k < PER / 16; ++k) {
                const int h0 = (k / (PP / 16)) * 64 + cg * PP + (k % (PP / 16)) * 16;
                float v[16];
                tmem_ld16(tl + C::R3 + h0, v);
                const float *bb = b2_g + h0;
                if (a.skip) {
                    float sk[16];
                    const float *sp = a.skip + brow + (size_t)h0 * l;
#pragma unroll
                    for (int i = 0; i < 16; ++i) sk[i] = valid ? __ldg(chan<0>(sp, i, l)) : 0.f;
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const s2::V2 y = (s2::V2(v[i], v[i + 1]) + s2::V2(__ldg(bb + i), __ldg(bb + i + 1))) + s2::V2(sk[i], sk[i + 1]);
                        v[i] = y.v.x;
                        v[i + 1] = y.v.y;
                    }
                } else {
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; i += 2) {
                        const s2::V2 y = s2::V2(v[i], v[i + 1]) + s2::V2(__ldg(bb + i), __ldg(bb + i + 1));
                        v[i] = y.v.x;
                        v[i + 1] = y.v.y;
                    }
                }
                if (valid) {
                    float *oq = op + (size_t)h0 * l;
#pragma unroll
                    for (int i = 0; i < 16; ++i) *chan<0>(oq, i, l) = v[i];
                }
                if (k == 0) piv = v[0];
                stat_acc16(v, s2::V2(piv), sd, sq);
            }
            stat_finish(sd, sq, piv, PER, mean, M2);
            exchange(mean, M2);
            if (cg == 0 && valid)
                *reinterpret_cast<float2 *>(a.stats_out + ((size_t)b * l + t) * 2) = make_float2(mean, rsqrtf(M2 * (1.0f / H)));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MW) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// weight image of the H = 256 kernel (stage order: U256::stage_info) + biases (bo permuted | b1 | b2)
__global__ void umma_pack256_kernel(const float *__restrict__ Wo_t, const float *__restrict__ W1_t,
                                    const float *__restrict__ W2_t, const float *__restrict__ bo, const float *__restrict__ b1,
                                    const float *__restrict__ b2, uint8_t *__restrict__ img, float *__restrict__ bimg) {
    constexpr int H = 256, F = 512;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one (stage, row, 16-byte chunk)
    if (idx < (size_t)5 * H) {
        float v;
        if (idx < (size_t)2 * H) {
            const int nc = idx / 128, i = idx % 128;
            // value half as is; gate half pre-multiplied by -log2(e): the sigmoid exponent is one FFMA2 of the accumulator
            v = i < 64 ? bo[nc * 64 + i] : -1.4426950408889634f * bo[H + nc * 64 + (i - 64)];
        } else if (idx < (size_t)4 * H)
            v = b1[idx - 2 * H];
        else
            v = b2[idx - 4 * H];
        bimg[idx] = v;
    }
    if (idx >= (size_t)U256::NSTG * 128 * 8) return;
    const int stage = idx / (128 * 8), rem = idx % (128 * 8), row = rem / 8, j = rem % 8;
    int gemm, nblk, kc;
    U256::stage_info(stage, gemm, nblk, kc);
    const float *Wt = gemm == 0 ? Wo_t : (gemm == 1 ? W1_t : W2_t);
    const int M = gemm == 2 ? H : F;
    const int n = gemm == 0 ? (row < 64 ? nblk * 64 + row : H + nblk * 64 + (row - 64)) : nblk * 128 + row;
    const int k0 = kc * 64 + j * 8;
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float w0 = Wt[(size_t)(k0 + 2 * e) * M + n], w1 = Wt[(size_t)(k0 + 2 * e + 1) * M + n];
        const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(w0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(w1 - __bfloat162float(h1));
        hp[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lp[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t off = (size_t)stage * UM_STAGE + (size_t)row * 128 + ((j ^ (row & 7)) << 4);
    *reinterpret_cast<uint4 *>(img + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4 *>(img + off + UM_STAGE / 2) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
}

// ---------------------------------------------------------------------------------------
// finalize: folded fp32 weights (transposed [K][M]) -> the streamed shared-memory image
// ---------------------------------------------------------------------------------------
template <int H>
__global__ void umma_pack_kernel(const float *__restrict__ Wo_t, const float *__restrict__ W1_t,
                                 const float *__restrict__ W2_t, const float *__restrict__ bo, const float *__restrict__ b1,
                                 const float *__restrict__ b2, uint8_t *__restrict__ img, float *__restrict__ bimg) {
    using C = UCfg<H, 1>;
    constexpr int F = 2 * H;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one 16-byte chunk of one part pair
    if (idx < (size_t)5 * H) {
        float v;
        if (idx < (size_t)2 * H) {
            const int nc = idx / 128, i = idx % 128;
            // value half as is; gate half pre-multiplied by -log2(e): the sigmoid exponent is one FFMA2 of the accumulator
            v = i < 64 ? bo[nc * 64 + i] : -1.4426950408889634f * bo[H + nc * 64 + (i - 64)];
        } else if (idx < (size_t)4 * H)
            v = b1[idx - 2 * H];
        else
            v = b2[idx - 4 * H];
        bimg[idx] = v;
    }
    // enumerate (block, row, logical chunk j): hi and lo chunks are written together
    const size_t g12 = (size_t)2 * C::NC1 * C::KC1 * 128 * 8;             // chunks in G1 + G2 (hi part)
    const size_t g3 = (size_t)C::KC3 * C::NC3 * C::NR3 * 8;
    if (idx >= g12 + g3) return;
    const float *Wt;
    int M, n, k0;
    size_t base;            // byte offset of the block
    int NR, row, j;
    if (idx < g12) {
        const int blk = idx / (128 * 8), rem = idx % (128 * 8);
        row = rem / 8;
        j = rem % 8;
        const int gemm = blk / (C::NC1 * C::KC1), bb = blk % (C::NC1 * C::KC1), nc = bb / C::KC1, kc = bb % C::KC1;
        Wt = gemm == 0 ? Wo_t : W1_t;
        M = F;
        n = gemm == 0 ? (row < 64 ? nc * 64 + row : H + nc * 64 + (row - 64)) : nc * 128 + row;
        k0 = kc * 64 + j * 8;
        base = (size_t)blk * UM_STAGE;
        NR = 128;
    } else {
        const size_t i3 = idx - g12;
        const int blk = i3 / (C::NR3 * 8), rem = i3 % (C::NR3 * 8);
        row = rem / 8;
        j = rem % 8;
        const int kc = blk / C::NC3, nc = blk % C::NC3;
        Wt = W2_t;
        M = H;
        n = nc * C::NR3 + row;
        k0 = kc * 64 + j * 8;
        base = (size_t)2 * C::NC1 * C::KC1 * UM_STAGE + (size_t)blk * C::NR3 * 256;
        NR = C::NR3;
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
    const size_t off = base + (size_t)row * 128 + ((j ^ (row & 7)) << 4);
    *reinterpret_cast<uint4 *>(img + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4 *>(img + off + (size_t)NR * 128) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
}

bool mix_umma_supported(int H, int F, int l) { return F == 2 * H && (H == 64 || H == 128 || H == 256) && l >= 1; }

size_t mix_umma_image_bytes(int H) {
    return H == 64 ? UCfg<64, 2>::IMG_BYTES : (H == 128 ? UCfg<128, 2>::IMG_BYTES : (H == 256 ? U256::IMG_BYTES : 0));
}

int mix_umma_pack(int H, const float *Wo_t, const float *W1_t, const float *W2_t, const float *bo, const float *b1,
                  const float *b2, uint8_t *img, float *bimg, cudaStream_t st) {
    const size_t total = (size_t)(24 * H * H) / 32 + 5 * H;     // >= chunk pairs (24 H^2 bytes / 32) and bias entries
    const unsigned grid = (unsigned)ceil_div64((int64_t)total, 256);
    if (H == 64) umma_pack_kernel<64><<<grid, 256, 0, st>>>(Wo_t, W1_t, W2_t, bo, b1, b2, img, bimg);
    else if (H == 128) umma_pack_kernel<128><<<grid, 256, 0, st>>>(Wo_t, W1_t, W2_t, bo, b1, b2, img, bimg);
    else if (H == 256) umma_pack256_kernel<<<grid, 256, 0, st>>>(Wo_t, W1_t, W2_t, bo, b1, b2, img, bimg);
    else {
        set_error("mix_umma_pack: H=%d unsupported", H);
        return DWB_ERR_UNSUPPORTED;
    }
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

template <int H, int CS>
static int launch_umma(const MixArgs &a, int B, cudaStream_t st) {
    using C = UCfg<H, CS>;
    auto k = sashimi_mix_umma_kernel<H, CS>;
    static_assert(C::SMEM <= 227 * 1024, "tile does not fit shared memory");
    DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
    dim3 grid(ceil_div(a.l, UM_TT), B);
    k<<<grid, C::NTHREADS, C::SMEM, st>>>(a);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

// tensor map of a row-major (rows, cols) fp32 matrix for boxes of [box_rows x box_cols] (dense rows in shared memory).
// The driver entry point is resolved through the runtime, so libdwb.so does not link libcuda.
static int make_tmap_rows(CUtensorMap *tm, const float *base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_cols) {
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
            fn = nullptr;
        return (EncodeFn)fn;
    }();
    DWB_REQUIRE(encode, DWB_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t gdim[2] = {cols, rows}, gstride[1] = {cols * sizeof(float)};
    const cuuint32_t box[2] = {box_cols, box_rows}, estr[2] = {1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), gdim, gstride, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DWB_REQUIRE(r == CUDA_SUCCESS, DWB_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d (rows %llu, cols %llu)", (int)r,
                (unsigned long long)rows, (unsigned long long)cols);
    return DWB_OK;
}

template <int H, int CS, int LC, bool STAGE>
static int launch_umma_pers(const MixArgs &a, int B, cudaStream_t st) {
    using P = PCfg<H, CS>;
    auto k = sashimi_mix_umma_pers_kernel<H, CS, LC, STAGE>;
    static int sms[16] = {};
    int dev = 0;
    DWB_CUDA(cudaGetDevice(&dev));
    if (!sms[dev & 15]) {
        DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM));
        DWB_CUDA(cudaDeviceGetAttribute(&sms[dev & 15], cudaDevAttrMultiProcessorCount, dev));
    }
    const int ntiles = B * ceil_div(a.l, UM_TT);
    const int grid = std::min(sms[dev & 15], ceil_div(ntiles, P::NG));
    static const int stagger = [] { const char *e = getenv("DWB_UMMA_STAGGER"); return e ? atoi(e) : 0; }();
    CUtensorMap tm_x, tm_g;
    memset(&tm_x, 0, sizeof(tm_x));
    memset(&tm_g, 0, sizeof(tm_g));
    if (STAGE) {        // staged inputs: (B*H, l) fp32 row-major, boxes of [H rows x 128 steps]
        int rc = make_tmap_rows(&tm_x, a.x, (uint64_t)B * H, (uint64_t)a.l, H, UM_TT);
        if (rc == DWB_OK) rc = make_tmap_rows(&tm_g, a.g, (uint64_t)B * H, (uint64_t)a.l, H, UM_TT);
        if (rc != DWB_OK) return rc;
    }
    DWB_CUDA(launch_pdl(k, dim3(grid), dim3(P::NTHREADS), P::SMEM, st, a, B, stagger, mix_reverse_order() ? 1 : 0, tm_x, tm_g));
    return DWB_OK;
}
// Inputs are staged by TMA whenever the rows are 16-byte aligned (l % 4 == 0, aligned base pointers); the stage lengths of the
// BASELINE configs additionally get a compile-time channel stride (LC) for the remaining strided accesses (skip, output).
template <int H, int CS>
static int launch_umma_pers_l(const MixArgs &a, int B, cudaStream_t st) {
    static const bool nostage = [] { const char *e = getenv("DWB_UMMA_STAGE"); return e && atoi(e) == 0; }();
    if (nostage || (a.l & 3) || a.l < UM_TT || ((reinterpret_cast<uintptr_t>(a.g) | reinterpret_cast<uintptr_t>(a.x)) & 15))
        return launch_umma_pers<H, CS, 0, false>(a, B, st);
    switch (a.l) {
        case 16000: return launch_umma_pers<H, CS, 16000, true>(a, B, st);
        case 4000: return launch_umma_pers<H, CS, 4000, true>(a, B, st);
        case 1000: return launch_umma_pers<H, CS, 1000, true>(a, B, st);
    }
    return launch_umma_pers<H, CS, 0, true>(a, B, st);
}

// Serpentine order between consecutive kernels: the S4 convolution walks the batch forwards, the mixing kernel
// that follows walks it backwards, so each kernel starts on the clips its predecessor touched last - the part of
// the (B,H,l) tensors that is still in the 126 MB L2 (DWB_SERPENTINE=0 turns it off for the A/B).
bool mix_reverse_order() {
    static const bool on = [] { const char *e = getenv("DWB_SERPENTINE"); return !(e && atoi(e) == 0); }();
    return on;
}

int mix_umma_launch(const MixArgs &a_in, int B, cudaStream_t st) {
    MixArgs a = a_in;
    a.rev = mix_reverse_order() ? 1 : 0;
    static const int jitter = [] { const char *e = getenv("DWB_DEBUG_JITTER"); return e ? atoi(e) : 0; }();
    a.jitter = jitter;
    DWB_REQUIRE(a.Wimg && a.bimg, DWB_ERR_STATE, "mix_umma: weights were not packed");
    DWB_REQUIRE((int64_t)a.H * a.l < (int64_t)1 << 31, DWB_ERR_UNSUPPORTED, "mix_umma: H*l = %lld needs 64-bit channel offsets",
                (long long)a.H * a.l);
    // The persistent kernel is the default for H = 64 and H = 128 at every batch size (round 2, B = 64: 243 / 167 us per
    // launch against 370 / 229 us for the per-tile kernel); DWB_UMMA=tile forces the per-tile form, which keeps its
    // operands in shared memory and serves as the second implementation in the variant tests.
    static const int mode = [] {
        const char *e = getenv("DWB_UMMA");
        return !e ? 0 : (std::string(e) == "tile" ? 1 : (std::string(e) == "pers" ? 2 : 0));
    }();
    if (a.H == 256) {
        static const int cs = [] { const char *e = getenv("DWB_UMMA256_CS"); return e ? atoi(e) : 4; }();
        auto k = cs == 2 ? sashimi_mix_umma256_kernel<2> : sashimi_mix_umma256_kernel<4>;
        DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)U256::SMEM));
        DWB_CUDA(launch_pdl(k, dim3(ceil_div(a.l, UM_TT), B), dim3(128 * (cs == 2 ? 2 : 4) + 64), U256::SMEM, st, a));
        return DWB_OK;
    }
    // (H = 128 with four epilogue threads per step or a third weight-ring stage measured no gain: 173 / 172 us against
    // 167-174 us - its tile time is set by the MMAs waiting for the 384 KB weight image that every tile re-streams from L2)
    const bool pers = mode != 1;
    if (pers) switch (a.H) {
        case 64: return launch_umma_pers_l<64, 2>(a, B, st);
        case 128: return launch_umma_pers_l<128, 2>(a, B, st);
    }
    switch (a.H) {
        case 64: return launch_umma<64, 2>(a, B, st);
        case 128: return launch_umma<128, 2>(a, B, st);
    }
    set_error("mix_umma: H=%d unsupported", a.H);
    return DWB_ERR_UNSUPPORTED;
}

}  // namespace dwb
