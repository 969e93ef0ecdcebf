// SaShiMi DownPool / UpPool on the 5th-generation tensor cores.                    models/sashimi.py:23-58
//
//   down(s): x'[h*s+j][c] = x[b, h, c*s + j];  out = W x' + bias                     (K = Hi*s -> M = Ho), + stats
//   up(s):   y = W x + bias (K = Hi -> M = Ho*s);  out[b, h, c*s + j] = y[h*s+j][c] (+ skip),          + stats
//
// Same orientation as the mixing kernels: one CTA = 128 steps c (the MMA M dimension = the TMEM lanes) x ALL M
// output columns (M / 128 accumulators of 128 TMEM columns), K streamed in chunks of 64: eight loader warps read
// the A chunk from global memory with the einops rearrange folded into the addressing (down: one float4 / float2 =
// the s samples of a channel per thread, up: one sample per channel, coalesced along time either way), split it
// into bf16 hi / lo and store it K-major / SW128 into a two-slot ring; the weights come pre-packed (split,
// swizzled, consumption order (kc, n-tile)) through a three-stage bulk-copy ring; three MMAs per product
// (hi*hi + lo*hi + hi*lo).  Because a CTA holds every output channel of its steps, the epilogue thread that owns a
// step computes the next block's TransposedLN statistics itself (up: s statistics per thread, one per sub-step).
//
// The output head (models/sashimi.py:310-312, wavenet.py:205-209) is the third mode of the same kernel:
//   eps = wz . relu(Wf LN(x) prescale + bf) + bz,  then the DDPM update  x <- (x - c1 eps)/sqrt(alpha) (+ sigma z)
// (generate.py:52-54): the loader applies the final LayerNorm, the epilogue thread reduces its columns against wz.
//
// Modes GLU / GELU / RES are the three contractions of a DiffWaveBlock whose width does not fit the fused mixing kernels
// (H = 512, the centre stage of unet d128; models/sashimi.py:157-182): the same CTA shape with 512 output columns per CTA
// (blockIdx.z selects the column half of the 2H / F wide products), so an A chunk is loaded and split once for four
// accumulators instead of once per 128 columns as in mix_gemm_umma.cu, and G3 - which sees all H channels - writes the
// next LayerNorm's statistics itself.
#include "common.cuh"
#include "fft_simd2.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace dwb {
using namespace umma;

constexpr int PU_STAGE = 32768, PU_SLAB = 32768, PU_NSW = 3, PU_NSU = 2, PU_THREADS = 320;
constexpr int PU_OFF_RING = PU_NSU * PU_SLAB;
constexpr int PU_OFF_EX = PU_OFF_RING + PU_NSW * PU_STAGE;          // statistics exchange: [2 halves][128 steps][4] float2
constexpr int PU_OFF_BAR = PU_OFF_EX + 2 * 128 * 4 * 8;
constexpr int PU_NBAR = 2 * PU_NSW + 2 * PU_NSU + 1;
constexpr int PU_SMEM = PU_OFF_BAR + PU_NBAR * 8 + 16 + 1024;

struct PoolUmmaArgs {
    const float *x;                // down: (B,Hi,li)          up / head: (B,Hi,li)
    const float *skip;             // up only: (B,Ho,li*s) or null
    const uint8_t *Wimg;           // stages (kc, n-tile), 32 KB each
    const float *bias;             // (M)
    float *out, *stats_out;        // down: (B,Ho,li/s), (B,li/s,2)   up: (B,Ho,li*s), (B,li*s,2)   head: (B,li), unused
    int Hi, Ho, li, K, M;
    // block GEMMs (GLU / GELU / RES)
    int Mc;                        // output columns of one CTA (<= 512); blockIdx.z = column group; M = all columns
    const float *res;              // GLU: block input x (B,H,l)   RES: x1 (B,H,l) (= out, in place)
    const float *cond;             // GLU: (cond_batch,H,l) or null
    int cond_stride_b;
    // head and GELU
    const float *stats;            // (B,li,2) LayerNorm statistics applied to A while loading, or null
    float ln_m, ln_s, prescale;
    const float *wz;               // (M)
    float bz;
    const float *upd_x, *ctl;      // fused DDPM update (HeadArgs)
    const float *const *noise_base;
};

enum { PU_DOWN = 0, PU_UP = 1, PU_HEAD = 2, PU_GLU = 3, PU_GELU = 4, PU_RES = 5 };

template <int MODE, int S>
__global__ void __launch_bounds__(PU_THREADS, 1)
pool_umma_kernel(PoolUmmaArgs a) {
    constexpr bool UP = MODE != PU_DOWN;          // A operand = one sample per input channel (up pool, head)
    extern __shared__ uint8_t smem_raw[];
    uint8_t *sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t *slabs = sm, *ring = sm + PU_OFF_RING;
    float2 *ex = reinterpret_cast<float2 *>(sm + PU_OFF_EX);
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + PU_OFF_BAR);
    uint32_t *tptr = reinterpret_cast<uint32_t *>(bars + PU_NBAR);
    uint64_t *wfull = bars, *wempty = wfull + PU_NSW, *ufull = wempty + PU_NSW, *uempty = ufull + PU_NSU, *acc_ready = uempty + PU_NSU;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = blockIdx.y, t0 = blockIdx.x * 128;
    const int li = a.li, lo = UP ? li * S : li / S, lc = UP ? li : lo;      // lc: steps of the GEMM's M dimension
    constexpr bool BLK = MODE >= PU_GLU;            // block GEMM: a.Mc of the a.M columns per CTA
    const int Mcta = BLK ? a.Mc : a.M, zc = BLK ? (int)blockIdx.z : 0;
    const int KC = a.K / 64, NC = (Mcta + 127) / 128;
    const int ncols = Mcta < 128 ? Mcta : 128;      // columns of one accumulator (head with C = 64: one 64-wide tile)
    if (tid == 0) {
        for (int i = 0; i < PU_NSW; ++i) {
            mbar_init(wfull + i, 1);
            mbar_init(wempty + i, 1);
        }
        for (int i = 0; i < PU_NSU; ++i) {
            mbar_init(ufull + i, 128);
            mbar_init(uempty + i, 1);
        }
        mbar_init(acc_ready, 1);
        fence_mbar_init();
    }
    const uint32_t tcols = (uint32_t)(NC * 128 > 64 ? NC * 128 : 64);       // a power of two >= 32
    pdl_trigger();
    if (warp == 9) tmem_alloc(tptr, tcols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tptr;
    if (warp != 8) pdl_wait();        // (the weight producer touches packed weights only)

    if (warp == 8) {
        // ================= weight producer =====================================================
        if (lane == 0) {
            const int n = KC * NC;
            for (int i = 0; i < n; ++i) {
                const int s = i % PU_NSW, ph = i / PU_NSW;
                mbar_wait(wempty + s, (ph & 1) ^ 1);
                mbar_arrive_expect_tx(wfull + s, PU_STAGE);
                bulk_g2s(ring + (size_t)s * PU_STAGE, a.Wimg + ((size_t)zc * n + i) * PU_STAGE, PU_STAGE, wfull + s);
            }
        }
    } else if (warp == 9) {
        // ================= MMA issuer ==========================================================
        {   // all 32 lanes run the loops; the *_w forms elect the issuing lane
            const uint32_t slab0 = smem_u32(slabs), ring0 = smem_u32(ring);
            const uint32_t idesc = idesc_bf16(128, ncols);
            int i = 0;
#pragma unroll 1
            for (int kc = 0; kc < KC; ++kc) {
                const int us = kc % PU_NSU;
                mbar_wait(ufull + us, (kc / PU_NSU) & 1);
                tc_fence_after();
                const uint32_t abase = slab0 + us * PU_SLAB;
#pragma unroll 1
                for (int nt = 0; nt < NC; ++nt, ++i) {
                    const int s = i % PU_NSW;
                    mbar_wait(wfull + s, (i / PU_NSW) & 1);
                    tc_fence_after();
                    const uint32_t bbase = ring0 + s * PU_STAGE;
                    if (elect_one()) {      // one election per block of 12 MMAs
#pragma unroll
                        for (int term = 0; term < 3; ++term) {
                            const uint32_t ao = abase + (term == 1 ? PU_SLAB / 2 : 0), bo = bbase + (term == 2 ? PU_STAGE / 2 : 0);
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                mma_bf16_ss(tmem + nt * 128, smem_desc_sw128(ao + ks * 32), smem_desc_sw128(bo + ks * 32), idesc,
                                            (kc > 0 || term > 0 || ks > 0) ? 1u : 0u);
                        }
                    }
                    __syncwarp();
                    mma_commit_w(wempty + s);
                }
                mma_commit_w(uempty + us);
            }
            mma_commit_w(acc_ready);
        }
    } else {
        // ================= loaders, then epilogue: one step per thread ===========================
        const int q = warp & 3, cg = warp >> 2;
        const int r = 32 * q + lane, c = t0 + r;
        const bool valid = c < lc;
        const uint32_t tl = tmem + ((uint32_t)(32 * q) << 16);
        const size_t cc = valid ? c : 0;
        float hsc = a.prescale, hsh = 0.f;
        if ((MODE == PU_HEAD || MODE == PU_GELU) && a.stats && valid) {
            const float2 ms = *reinterpret_cast<const float2 *>(a.stats + ((size_t)b * li + c) * 2);
            hsc = a.ln_s * ms.y * a.prescale;
            hsh = (a.ln_m - ms.x) * hsc;
        }
#pragma unroll 1
        for (int kc = cg; kc < KC; kc += PU_NSU) {
            float v[64];
            if (UP) {
                const char *ap = reinterpret_cast<const char *>(a.x + ((size_t)b * a.Hi + (size_t)kc * 64) * li + cc);
                const unsigned rowb = 4u * (unsigned)li;                     // one IMAD.WIDE per load instead of a live pointer each
#pragma unroll
                for (int i = 0; i < 64; ++i)
                    v[i] = valid ? __ldg(reinterpret_cast<const float *>(ap + (unsigned long long)rowb * (unsigned)i)) : 0.f;
                if (MODE == PU_HEAD || MODE == PU_GELU) {           // v = (ln_s rstd)(x - mean + ln_m) prescale
#pragma unroll
                    for (int i = 0; i < 64; ++i) v[i] = valid ? fmaf(v[i], hsc, hsh) : 0.f;
                }
            } else {
                // k = h*S + j: the S consecutive samples of input channel h that feed output step c
                constexpr int CH = 64 / S;
                const float *ap = a.x + ((size_t)b * a.Hi + (size_t)kc * CH) * li + cc * S;
#pragma unroll
                for (int hh = 0; hh < CH; ++hh, ap += li) {
                    if (S == 4) {
                        const float4 w = valid ? __ldg(reinterpret_cast<const float4 *>(ap)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        v[4 * hh] = w.x; v[4 * hh + 1] = w.y; v[4 * hh + 2] = w.z; v[4 * hh + 3] = w.w;
                    } else {
                        const float2 w = valid ? __ldg(reinterpret_cast<const float2 *>(ap)) : make_float2(0.f, 0.f);
                        v[2 * hh] = w.x; v[2 * hh + 1] = w.y;
                    }
                }
            }
            mbar_wait(uempty + cg, ((kc / PU_NSU) & 1) ^ 1);
            uint8_t *slab = slabs + cg * PU_SLAB;
#pragma unroll
            for (int c8 = 0; c8 < 8; ++c8) {
                uint4 hi, lo;
                split8(v + 8 * c8, hi, lo);
                const uint32_t off = sw128_off(r, c8);
                *reinterpret_cast<uint4 *>(slab + off) = hi;
                *reinterpret_cast<uint4 *>(slab + PU_SLAB / 2 + off) = lo;
            }
            fence_proxy_async_smem();
            mbar_arrive(ufull + cg);
        }

        mbar_wait(acc_ready, 0);
        tc_fence_after();
        const int MH = Mcta / 2, n0 = cg * MH;                 // this thread's output columns [n0, n0 + MH)
        if (MODE == PU_GLU) {
            // accumulator nt holds [64 value | 64 gate] columns of channels 64 (zc NC + nt) ..; this thread: NC / 2 accumulators
            const unsigned rowb = 4u * (unsigned)li;
#pragma unroll 1
            for (int nt = cg * (NC / 2); nt < (cg + 1) * (NC / 2); ++nt) {
                const int h0 = (zc * NC + nt) * 64;
                const char *xp = reinterpret_cast<const char *>(a.res + ((size_t)b * a.Ho + h0) * li + cc);
                char *op = reinterpret_cast<char *>(a.out + ((size_t)b * a.Ho + h0) * li + cc);
                const char *cb = a.cond ? reinterpret_cast<const char *>(a.cond + ((size_t)(a.cond_stride_b ? b : 0) * a.Ho + h0) * li + cc) : nullptr;
#pragma unroll 1
                for (int sc = 0; sc < 4; ++sc) {
                    float av[16], gv[16], xv[16];
                    tmem_ld16(tl + nt * 128 + sc * 16, av);
                    tmem_ld16(tl + nt * 128 + 64 + sc * 16, gv);
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        xv[i] = valid ? __ldg(reinterpret_cast<const float *>(xp + (unsigned long long)rowb * (unsigned)(sc * 16 + i))) : 0.f;
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int h = h0 + sc * 16 + i;
                        float y = (av[i] + __ldg(a.bias + h)) * __fdividef(1.0f, 1.0f + __expf(-(gv[i] + __ldg(a.bias + a.Ho + h))));
                        if (cb && valid) y += __ldg(reinterpret_cast<const float *>(cb + (unsigned long long)rowb * (unsigned)(sc * 16 + i)));
                        if (valid) *reinterpret_cast<float *>(op + (unsigned long long)rowb * (unsigned)(sc * 16 + i)) = xv[i] + y;
                    }
                }
            }
        } else if (MODE == PU_GELU) {
            const unsigned rowb = 4u * (unsigned)li;
            const int f0 = zc * Mcta + n0;                      // first of this thread's hidden channels
            char *op = reinterpret_cast<char *>(a.out + ((size_t)b * a.M + f0) * li + cc);
#pragma unroll 1
            for (int sc = 0; sc < MH / 16; ++sc) {
                float v[16];
                tmem_ld16(tl + n0 + sc * 16, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float z = gelu_fast(v[i] + __ldg(a.bias + f0 + sc * 16 + i));
                    if (valid) *reinterpret_cast<float *>(op + (unsigned long long)rowb * (unsigned)(sc * 16 + i)) = z;
                }
            }
        } else if (MODE == PU_RES) {
            // all H channels of the step in this CTA (Mc = M = H): x2 = x1 + (W2 hid + b2) (+skip), statistics for the next norm
            const unsigned rowb = 4u * (unsigned)li;
            const char *xp = reinterpret_cast<const char *>(a.res + ((size_t)b * a.M + n0) * li + cc);
            const char *sp = a.skip ? reinterpret_cast<const char *>(a.skip + ((size_t)b * a.M + n0) * li + cc) : nullptr;
            char *op = reinterpret_cast<char *>(a.out + ((size_t)b * a.M + n0) * li + cc);
            float sd = 0.f, sq = 0.f, piv = 0.f;
#pragma unroll 1
            for (int sc = 0; sc < MH / 16; ++sc) {
                float v[16], pre[16];
                tmem_ld16(tl + n0 + sc * 16, v);
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const unsigned long long o = (unsigned long long)rowb * (unsigned)(sc * 16 + i);
                    pre[i] = valid ? *reinterpret_cast<const float *>(xp + o) : 0.f;            // x1 (written by G1: plain loads)
                    if (sp) pre[i] += valid ? __ldg(reinterpret_cast<const float *>(sp + o)) : 0.f;
                }
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += pre[i] + __ldg(a.bias + n0 + sc * 16 + i);
                if (sc == 0) piv = v[0];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float d = v[i] - piv;
                    sd += d;
                    sq = fmaf(d, d, sq);
                    if (valid) *reinterpret_cast<float *>(op + (unsigned long long)rowb * (unsigned)(sc * 16 + i)) = v[i];
                }
            }
            const float inv = 1.0f / (float)MH;
            ex[cg * 128 + r] = make_float2(fmaf(sd, inv, piv), fmaxf(fmaf(-sd * inv, sd, sq), 0.f));
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (cg == 0 && valid) {
                const float2 e0 = ex[r], e1 = ex[128 + r];
                const float mt = 0.5f * (e0.x + e1.x), d0 = e0.x - mt, d1 = e1.x - mt;
                const float m2 = e0.y + e1.y + (d0 * d0 + d1 * d1) * (float)MH;
                *reinterpret_cast<float2 *>(a.stats_out + ((size_t)b * li + c) * 2) = make_float2(mt, rsqrtf(m2 / (float)a.M));
            }
        } else if (MODE == PU_HEAD) {
            float part = 0.f;
#pragma unroll 1
            for (int sc = 0; sc < MH / 16; ++sc) {
                float v[16];
                tmem_ld16(tl + n0 + sc * 16, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    part = fmaf(__ldg(a.wz + n0 + sc * 16 + i), fmaxf(v[i] + __ldg(a.bias + n0 + sc * 16 + i), 0.f), part);
            }
            reinterpret_cast<float *>(ex)[cg * 128 + r] = part;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (cg == 0 && valid) {
                const float e = a.bz + (reinterpret_cast<float *>(ex)[r] + reinterpret_cast<float *>(ex)[128 + r]);
                const size_t o = (size_t)b * li + c;
                if (a.upd_x) {
                    // x <- (x - c1 eps) / sqrt(alpha) (+ sigma z)           generate.py:52-54
                    const float c1 = __ldg(a.ctl), sqrt_alpha = __ldg(a.ctl + 1), sigma = __ldg(a.ctl + 2);
                    const int slot = __float_as_int(__ldg(a.ctl + 3));
                    float xn = (a.upd_x[o] - c1 * e) / sqrt_alpha;
                    if (slot >= 0) xn += sigma * (*a.noise_base)[(size_t)slot * gridDim.y * li + o];
                    a.out[o] = xn;
                } else {
                    a.out[o] = e;
                }
            }
        } else if (!UP) {
            float *op = a.out + ((size_t)b * a.Ho + n0) * lo + cc;
            float sd = 0.f, sq = 0.f, piv = 0.f;
#pragma unroll 1
            for (int sc = 0; sc < MH / 16; ++sc) {
                float v[16];
                tmem_ld16(tl + n0 + sc * 16, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __ldg(a.bias + n0 + sc * 16 + i);
                if (sc == 0) piv = v[0];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float d = v[i] - piv;
                    sd += d;
                    sq = fmaf(d, d, sq);
                }
                if (valid) {
                    float *oq = op + (size_t)(sc * 16) * lo;
#pragma unroll
                    for (int i = 0; i < 16; ++i, oq += lo) *oq = v[i];
                }
            }
            const float inv = 1.0f / (float)MH;
            const float mean = fmaf(sd, inv, piv), M2 = fmaxf(fmaf(-sd * inv, sd, sq), 0.f);
            ex[cg * 128 + r] = make_float2(mean, M2);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (cg == 0 && valid) {
                const float2 e0 = ex[r], e1 = ex[128 + r];
                const float mt = 0.5f * (e0.x + e1.x), d0 = e0.x - mt, d1 = e1.x - mt;
                const float m2 = e0.y + e1.y + (d0 * d0 + d1 * d1) * (float)MH;
                *reinterpret_cast<float2 *>(a.stats_out + ((size_t)b * lo + c) * 2) = make_float2(mt, rsqrtf(m2 / (float)a.M));
            }
        } else {
            // columns n = h*S + j: 16 columns = 16/S output channels x the S sub-steps of this thread's step
            constexpr int HPC = 16 / S;
            const int h0 = n0 / S, nh = MH / S;                 // this thread's output channels [h0, h0 + nh)
            float *op = a.out + ((size_t)b * a.Ho + h0) * lo + cc * S;
            const float *sp = a.skip ? a.skip + ((size_t)b * a.Ho + h0) * lo + cc * S : nullptr;
            float sd[S], sq[S], piv[S];
#pragma unroll
            for (int j = 0; j < S; ++j) sd[j] = sq[j] = piv[j] = 0.f;
#pragma unroll 1
            for (int sc = 0; sc < MH / 16; ++sc) {
                float v[16], sk[16];
                tmem_ld16(tl + n0 + sc * 16, v);
                if (sp) {
#pragma unroll
                    for (int hh = 0; hh < HPC; ++hh) {
                        const float *sq_ = sp + (size_t)(sc * HPC + hh) * lo;
                        if (S == 4) {
                            const float4 w = valid ? __ldg(reinterpret_cast<const float4 *>(sq_)) : make_float4(0.f, 0.f, 0.f, 0.f);
                            sk[4 * hh] = w.x; sk[4 * hh + 1] = w.y; sk[4 * hh + 2] = w.z; sk[4 * hh + 3] = w.w;
                        } else {
                            const float2 w = valid ? __ldg(reinterpret_cast<const float2 *>(sq_)) : make_float2(0.f, 0.f);
                            sk[2 * hh] = w.x; sk[2 * hh + 1] = w.y;
                        }
                    }
                }
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] += __ldg(a.bias + n0 + sc * 16 + i) + (sp ? sk[i] : 0.f);
                if (sc == 0) {
#pragma unroll
                    for (int j = 0; j < S; ++j) piv[j] = v[j];
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const float d = v[i] - piv[i % S];
                    sd[i % S] += d;
                    sq[i % S] = fmaf(d, d, sq[i % S]);
                }
                if (valid) {
#pragma unroll
                    for (int hh = 0; hh < HPC; ++hh) {
                        float *oq = op + (size_t)(sc * HPC + hh) * lo;
                        if (S == 4) *reinterpret_cast<float4 *>(oq) = make_float4(v[4 * hh], v[4 * hh + 1], v[4 * hh + 2], v[4 * hh + 3]);
                        else *reinterpret_cast<float2 *>(oq) = make_float2(v[2 * hh], v[2 * hh + 1]);
                    }
                }
            }
            const float inv = 1.0f / (float)nh;
#pragma unroll
            for (int j = 0; j < S; ++j)
                ex[(cg * 128 + r) * 4 + j] = make_float2(fmaf(sd[j], inv, piv[j]), fmaxf(fmaf(-sd[j] * inv, sd[j], sq[j]), 0.f));
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (cg == 0 && valid) {
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const float2 e0 = ex[r * 4 + j], e1 = ex[(128 + r) * 4 + j];
                    const float mt = 0.5f * (e0.x + e1.x), d0 = e0.x - mt, d1 = e1.x - mt;
                    const float m2 = e0.y + e1.y + (d0 * d0 + d1 * d1) * (float)nh;
                    *reinterpret_cast<float2 *>(a.stats_out + ((size_t)b * lo + (size_t)c * S + j) * 2) = make_float2(mt, rsqrtf(m2 / (float)a.Ho));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        tc_fence_after();
        tmem_dealloc(tmem, tcols);
    }
}

// image of a transposed weight Wt [K][M]: stages (kc, n-tile) = [128 rows x 64 k] hi block, then lo block (K-major SW128)
// nsplit column groups of M / nsplit columns each (stage order: group, kc, n-tile inside the group); glu: tile T of 128
// rows = value rows 64 T .. | gate rows M/2 + 64 T .. (the pairing the GLU epilogue expects)
__global__ void pool_umma_pack_kernel(const float *__restrict__ Wt, int K, int M, uint8_t *__restrict__ img, int nsplit, int glu) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;     // one (stage, row, 16-byte chunk)
    const int NCT = (M + 127) / 128, NC = NCT / nsplit, KC = K / 64;
    if (idx >= (size_t)KC * NCT * 128 * 8) return;
    const size_t stage = idx / (128 * 8);
    const int rem = idx % (128 * 8), row = rem / 8, j8 = rem % 8;
    const int z = (int)(stage / ((size_t)KC * NC)), kc = (int)((stage / NC) % KC), nt = (int)(stage % NC);
    const int T = z * NC + nt;
    const int n = glu ? (row < 64 ? T * 64 + row : M / 2 + T * 64 + (row - 64)) : T * 128 + row, k0 = kc * 64 + j8 * 8;
    uint32_t hp[4], lp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float w0 = n < M ? Wt[(size_t)(k0 + 2 * e) * M + n] : 0.f, w1 = n < M ? Wt[(size_t)(k0 + 2 * e + 1) * M + n] : 0.f;
        const __nv_bfloat16 h0 = __float2bfloat16_rn(w0), h1 = __float2bfloat16_rn(w1);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(w0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(w1 - __bfloat162float(h1));
        hp[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lp[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
    const size_t off = stage * PU_STAGE + (size_t)row * 128 + ((j8 ^ (row & 7)) << 4);
    *reinterpret_cast<uint4 *>(img + off) = make_uint4(hp[0], hp[1], hp[2], hp[3]);
    *reinterpret_cast<uint4 *>(img + off + PU_STAGE / 2) = make_uint4(lp[0], lp[1], lp[2], lp[3]);
}

// M = all output columns of a step in one CTA's TMEM (a power of two <= 512), K in 64-wide chunks, float4 / float2 rows
bool pool_umma_supported(int Hi, int Ho, int s, bool up, int li) {
    const int M = up ? Ho * s : Ho, K = up ? Hi : Hi * s;
    if (!(s == 2 || s == 4)) return false;
    if (!(M == 128 || M == 256 || M == 512) || K % 64 != 0) return false;
    return up ? true : (li % s == 0 && (li % 4) == 0);
}

size_t pool_umma_image_bytes(int Hi, int Ho, int s, bool up) {
    const size_t M = up ? (size_t)Ho * s : Ho, K = up ? Hi : (size_t)Hi * s;
    return (K / 64) * (M / 128) * PU_STAGE;
}

int pool_umma_pack(int Hi, int Ho, int s, bool up, const float *W_t, uint8_t *img, cudaStream_t st) {
    const int M = up ? Ho * s : Ho, K = up ? Hi : Hi * s;
    const size_t total = (size_t)(K / 64) * (M / 128) * 128 * 8;
    pool_umma_pack_kernel<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(W_t, K, M, img, 1, 0);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}

template <int MODE, int S>
static int launch_pool_umma(const PoolUmmaArgs &g, int B, cudaStream_t st) {
    auto k = pool_umma_kernel<MODE, S>;
    DWB_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PU_SMEM));
    const int lc = MODE != PU_DOWN ? g.li : g.li / S;
    const int nz = MODE >= PU_GLU ? g.M / g.Mc : 1;
    DWB_CUDA(launch_pdl(k, dim3(ceil_div(lc, 128), B, nz), dim3(PU_THREADS), PU_SMEM, st, g));
    return DWB_OK;
}

int pool_umma_launch(const PoolArgs &a, const uint8_t *Wimg, bool up, int B, cudaStream_t st) {
    DWB_REQUIRE(Wimg, DWB_ERR_STATE, "pool_umma: weights were not packed");
    const uintptr_t al = (uintptr_t)a.x | (uintptr_t)a.out | (uintptr_t)(a.skip ? a.skip : a.x);
    DWB_REQUIRE((al & 15) == 0 && B <= 65535, DWB_ERR_UNSUPPORTED, "pool_umma: unaligned tensors");
    PoolUmmaArgs g{};
    g.x = a.x; g.skip = a.skip; g.Wimg = Wimg; g.bias = a.bias; g.out = a.out; g.stats_out = a.stats_out;
    g.Hi = a.Hi; g.Ho = a.Ho; g.li = a.li;
    g.M = up ? a.Ho * a.s : a.Ho;
    g.K = up ? a.Hi : a.Hi * a.s;
    if (up) return a.s == 4 ? launch_pool_umma<PU_UP, 4>(g, B, st) : launch_pool_umma<PU_UP, 2>(g, B, st);
    return a.s == 4 ? launch_pool_umma<PU_DOWN, 4>(g, B, st) : launch_pool_umma<PU_DOWN, 2>(g, B, st);
}

// ---- output head: C -> C (+ReLU) -> 1 (+ DDPM update); C in {64, 128, 256, 512}
bool head_umma_supported(int C) { return C == 64 || C == 128 || C == 256 || C == 512; }
size_t head_umma_image_bytes(int C) { return (size_t)(C / 64) * ((C + 127) / 128) * PU_STAGE; }
int head_umma_pack(int C, const float *Wf_t, uint8_t *img, cudaStream_t st) {
    const size_t total = (size_t)(C / 64) * ((C + 127) / 128) * 128 * 8;
    pool_umma_pack_kernel<<<(unsigned)ceil_div64((int64_t)total, 256), 256, 0, st>>>(Wf_t, C, C, img, 1, 0);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}
int head_umma_launch(const HeadArgs &h, const uint8_t *Wimg, int B, cudaStream_t st) {
    DWB_REQUIRE(Wimg && B <= 65535, DWB_ERR_STATE, "head_umma: weights were not packed");
    PoolUmmaArgs g{};
    g.x = h.x; g.Wimg = Wimg; g.bias = h.bf; g.out = h.out;
    g.Hi = h.C; g.Ho = h.C; g.li = h.l; g.K = h.C; g.M = h.C;
    g.stats = h.stats; g.ln_m = h.ln_m; g.ln_s = h.ln_s; g.prescale = h.prescale; g.wz = h.wz; g.bz = h.bz;
    g.upd_x = h.upd_x; g.ctl = h.ctl; g.noise_base = h.noise_base;
    return launch_pool_umma<PU_HEAD, 1>(g, B, st);
}

// ---- DiffWaveBlock channel mixing as three of these launches (H = 512): images G1 (2H columns, GLU pairing, two column
// groups) | G2 (F columns, two groups) | G3 (H columns, one group)
bool mix_gemm2_supported(int H, int F, int l) { return H == 512 && F == 2 * H && l >= 1; }
size_t mix_gemm2_image_bytes(int H, int F) { return ((size_t)(H / 64) * (2 * H / 128) + (size_t)(H / 64) * (F / 128) + (size_t)(F / 64) * (H / 128)) * PU_STAGE; }
int mix_gemm2_pack(int H, int F, const float *Wo_t, const float *W1_t, const float *W2_t, uint8_t *img, cudaStream_t st) {
    const size_t n1 = (size_t)(H / 64) * (2 * H / 128), n2 = (size_t)(H / 64) * (F / 128), n3 = (size_t)(F / 64) * (H / 128);
    pool_umma_pack_kernel<<<(unsigned)ceil_div64((int64_t)n1 * 128 * 8, 256), 256, 0, st>>>(Wo_t, H, 2 * H, img, 2 * H / 512, 1);
    pool_umma_pack_kernel<<<(unsigned)ceil_div64((int64_t)n2 * 128 * 8, 256), 256, 0, st>>>(W1_t, H, F, img + n1 * PU_STAGE, F / 512, 0);
    pool_umma_pack_kernel<<<(unsigned)ceil_div64((int64_t)n3 * 128 * 8, 256), 256, 0, st>>>(W2_t, F, H, img + (n1 + n2) * PU_STAGE, 1, 0);
    DWB_LAUNCH_CHECK();
    return DWB_OK;
}
// a.Wimg = the three images, a.bo / a.b1 / a.b2 fp32 biases; hid = (B, F, l) workspace; 4 launches (G1, stats(x1), G2, G3 + stats)
int mix_gemm2_launch(const MixArgs &a, float *hid, int B, cudaStream_t st) {
    DWB_REQUIRE(a.Wimg && hid && B <= 65535, DWB_ERR_STATE, "mix_gemm2: weights were not packed / no workspace");
    const int H = a.H, F = a.F, l = a.l;
    const size_t n1 = (size_t)(H / 64) * (2 * H / 128), n2 = (size_t)(H / 64) * (F / 128);
    PoolUmmaArgs g{};
    g.li = l; g.Hi = H; g.Ho = H;
    // G1: x1 = x + GLU(Wo g + bo) (+cond)
    g.x = a.g; g.Wimg = a.Wimg; g.bias = a.bo; g.res = a.x; g.cond = a.cond; g.cond_stride_b = a.cond_stride_b; g.out = a.out;
    g.K = H; g.M = 2 * H; g.Mc = 512;
    int rc = launch_pool_umma<PU_GLU, 1>(g, B, st);
    if (rc != DWB_OK) return rc;
    rc = mix_gemm_channel_stats(a.out, a.stats_out, H, l, B, st);
    if (rc != DWB_OK) return rc;
    // G2: hid = gelu(W1 LN2(x1) + b1)
    g.x = a.out; g.stats = a.stats_out; g.ln_m = a.ln2_m; g.ln_s = a.ln2_s; g.prescale = 1.0f; g.Wimg = a.Wimg + n1 * PU_STAGE; g.bias = a.b1;
    g.res = nullptr; g.cond = nullptr; g.out = hid; g.K = H; g.M = F; g.Mc = 512;
    rc = launch_pool_umma<PU_GELU, 1>(g, B, st);
    if (rc != DWB_OK) return rc;
    // G3: x2 = x1 + W2 hid + b2 (+skip), statistics of x2
    g.x = hid; g.stats = nullptr; g.Wimg = a.Wimg + (n1 + n2) * PU_STAGE; g.bias = a.b2; g.res = a.out; g.skip = a.skip;
    g.out = a.out; g.stats_out = a.stats_out; g.Hi = F; g.K = F; g.M = H; g.Mc = H;
    return launch_pool_umma<PU_RES, 1>(g, B, st);
}

}  // namespace dwb
